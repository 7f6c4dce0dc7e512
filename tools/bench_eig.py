"""Times ppx_sym_eig_topk on the Gram of a random s x m matrix (Tucker HOOI shape: s=800, m=1600, r=40)."""
import importlib, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ppx = importlib.import_module("pairwise-perturbation_b200")
import torch
s = int(sys.argv[1]) if len(sys.argv) > 1 else 800
m = int(sys.argv[2]) if len(sys.argv) > 2 else 2 * s
r = int(sys.argv[3]) if len(sys.argv) > 3 else 40
ctx = ppx.Ctx(0, workspace_bytes=1 << 30)
rng = np.random.default_rng(0)
Y = 0.5 + 0.5 * rng.random((s, m))
A = Y @ Y.T
Ad = ctx.to_device(np.asfortranarray(A))
U = ctx.empty(s * r)
ev = ctx.empty(r)
for rep in range(3):
    ctx.sync(); t0 = time.perf_counter()
    ctx.sym_eig_topk(Ad, s, r, U, ev)
    ctx.sync(); t1 = time.perf_counter()
    print(f"rep {rep}: {1e3 * (t1 - t0):.2f} ms")
basis = ctx.empty(s * s)
ctx.sync(); t0 = time.perf_counter()
ctx.sym_eig_topk(Ad, s, r, U, ev, basis=basis, basis_valid=False)
ctx.sync(); print(f"cold with basis output: {1e3 * (time.perf_counter() - t0):.2f} ms")
for eps in (1e-1, 1e-2, 1e-3):
    Y2 = Y + eps * (rng.random((s, m)) - 0.5)
    A2 = Y2 @ Y2.T
    A2d = ctx.to_device(np.asfortranarray(A2))
    b2 = basis.clone()
    ctx.sync(); t0 = time.perf_counter()
    ctx.sym_eig_topk(A2d, s, r, U, ev, basis=b2, basis_valid=True)
    ctx.sync(); t1 = time.perf_counter()
    w2, v2 = np.linalg.eigh(A2)
    Uh = ctx.to_host(U, (s, r))
    print(f"warm, perturbation {eps:g}: {1e3 * (t1 - t0):.2f} ms, projector err",
          np.abs(Uh @ Uh.T - v2[:, -r:] @ v2[:, -r:].T).max())
w, v = np.linalg.eigh(A)
ctx.sym_eig_topk(Ad, s, r, U, ev)
Uh = ctx.to_host(U, (s, r))
P = Uh @ Uh.T
Pr = v[:, -r:] @ v[:, -r:].T
print("projector err", np.abs(P - Pr).max(), "eval err", np.abs(ctx.to_host(ev, (r,)) - w[::-1][:r]).max() / w[-1])
