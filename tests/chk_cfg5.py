import importlib, sys, numpy as np
sys.path.insert(0, "/root/repo")
from oracle import pp_oracle as o
H = importlib.import_module("pairwise-perturbation_b200.host_api")
world = H.World(0, workspace_bytes=1 << 30)
for lens, R, maxiter in [((3, 32, 32, 200), 10, 80), ((3, 128, 128, 72), 10, 60)]:
    V, _ = o.make_tensor_r(lens, R)
    W, G = o.init_factors(lens, R), o.init_grad(lens, R)
    vnorm = np.linalg.norm(V)
    W_ref, G_ref = [w.copy() for w in W], [g.copy() for g in G]
    _, tr = o.alsCP_PP(V, W_ref, G_ref, 1e-10 * vnorm, 0.05, maxiter, resprint=10)
    Vd = H.Tensor.from_numpy(world, V)
    Wd = [H.Tensor.from_numpy(world, w, matrix=True) for w in W]
    Gd = [H.Tensor.from_numpy(world, g, matrix=True) for g in G]
    Fd = [H.Matrix(world, w.shape[0], w.shape[1]) for w in W]
    with H.Trace(quiet=True) as t:
        H.alsCP_PP(world, Vd, Wd, Gd, Fd, 1e-10 * vnorm, 0.05, maxiter, resprint=10)
    print(lens, "events equal:", t.events == [(0 if k == "DT" else 1, it) for k, it in tr.events], len(t.events))
    for rg, rr in zip(t.rows, tr.rows):
        print("  iter", int(rg[0]), "grad", rg[1], rr[1], "res", rg[3], rr[3])
    print("  factor err", max(np.abs(wd.numpy() - wr).max() / max(1, np.abs(wr).max()) for wd, wr in zip(Wd, W_ref)))
