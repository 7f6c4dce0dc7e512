"""Multi-GPU parity check (run under torch.distributed.run on a box with >= 2 GPUs; not part of the single-GPU
`-m gpu` suite):  the tensor is sharded along mode 0, alsCP_DT and alsCP_PP run through the C++ drivers with the NCCL
all-reduces, and every rank compares its rows of W_0 / the replicated W_j and the logged fitness with the CPU oracle;
then hosvd + alsTucker_DT / alsTucker_PP on a sharded tensor (replicated factors) against the oracle.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/check_multi_gpu.py
"""
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pp_oracle as o  # noqa: E402  (checker)

rank, nranks, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ppx = importlib.import_module("pairwise-perturbation_b200")
H = importlib.import_module("pairwise-perturbation_b200.host_api")

world = H.World(local, solver=0, use_graph=True, workspace_bytes=256 << 20)
idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    idt = torch.tensor(list(ppx.comm_unique_id()), dtype=torch.uint8, device="cuda")
dist.broadcast(idt, 0)



def check_allreduce():
    """ppx_allreduce_packed alone (the one-shot peer-memory kernel when the ranks could map each other, else NCCL):
    integer-valued doubles, so the sum is exact whatever the order -- bit-exact against the closed form, over sizes
    from one word to beyond the staging capacity, packed calls, and 60 back-to-back calls (flag epochs, double
    buffering)."""
    import ctypes as C
    lib = ppx.load_library()
    h = world.ctx_handle()
    p2p = bool(lib.ppx_comm_p2p(h))
    good = True
    sizes = [1, 2, 7, 50 * 50, 300 * 50, 2049, 96 * 1024, 96 * 1024 + 3, 200000]
    for rep in range(60):
        group = [sizes[(rep + k) % len(sizes)] for k in range(1 + rep % 3)]
        host = [np.arange(n, dtype=np.float64) % 11 + (rank + 1) * (rep + 1 + k) for k, n in enumerate(group)]
        bufs = [H.Tensor.from_numpy(world, x.reshape(-1, 1), matrix=True) for x in host]
        bp = (C.c_void_p * len(bufs))(*[b.data_ptr() for b in bufs])
        bn = (C.c_int64 * len(bufs))(*group)
        assert lib.ppx_allreduce_packed(h, bp, bn, len(bufs)) == 0
        for k, (b_, n) in enumerate(zip(bufs, group)):
            want = nranks * (np.arange(n, dtype=np.float64) % 11) + (rep + 1 + k) * nranks * (nranks + 1) / 2
            good &= bool(np.array_equal(b_.numpy().reshape(-1), want))
            b_.free()
    print(f"rank {rank} all-reduce ({'peer memory' if p2p else 'NCCL'}): {'OK' if good else 'MISMATCH'}", flush=True)
    return good


ok_all = True
allreduce_checked = False
for lens, R, maxiter, tol_init in [((13, 12, 11, 10), 4, 40, 0.1), ((9, 10, 11), 3, 30, 0.1), ((6, 7, 6, 5, 6, 7), 3, 40, 0.1),
                                   ((37, 9, 8, 7), 3, 24, 0.1)]:
    N = len(lens)
    if lens[0] < nranks:  # every rank needs at least one row of the sharded mode
        lens = (nranks,) + tuple(lens[1:])
    b, e = ppx.shard_range(lens[0], nranks, rank)
    if world.np == 1:
        world.comm_init(bytes(idt.cpu().tolist()), nranks, rank, 0, lens[0], b, e)  # NCCL communicator, once
    else:
        world.set_shard(0, lens[0], b, e)  # only the shard layout changes between problems
    if not allreduce_checked:
        allreduce_checked = True
        ok_all &= check_allreduce()
    V, _ = o.make_tensor_r(lens, R)
    W, G = o.init_factors(lens, R), o.init_grad(lens, R)
    vnorm = np.linalg.norm(V)
    for driver in ("DT", "PP", "PPpart", "plain"):
        W_ref, G_ref = [w.copy() for w in W], [g.copy() for g in G]
        if driver == "plain":
            # alsCP (als_CP.cxx:20-115): a full MTTKRP per mode, gradient_CP from scratch at iterations 0 and maxiter --
            # the partial MTTKRPs of gradient_CP must be reduced too, or the ranks disagree on the stopping test and
            # the next collective hangs (round-1 advisor finding).  Same normal equations as the DT sweeps.
            o.alsCP_DT(V, W_ref, G_ref, 0.0, 5, resprint=100, want_residual=False)
            Vd = H.Tensor.from_numpy(world, np.ascontiguousarray(V[b:e]))
            Wd = [H.Tensor.from_numpy(world, (w[b:e] if i == 0 else w), matrix=True) for i, w in enumerate(W)]
            Gd = [H.Tensor.from_numpy(world, (g[b:e] if i == 0 else g), matrix=True) for i, g in enumerate(G)]
            Fd = [H.Matrix(world, w.lens[0], R) for w in Wd]
            with H.Trace(quiet=True) as t:
                H.alsCP(world, Vd, Wd, Gd, Fd, 0.0, 5)
            worst_fac = 0.0
            for i in range(N):
                ref = W_ref[i][b:e] if i == 0 else W_ref[i]
                worst_fac = max(worst_fac, float(np.abs(Wd[i].numpy() - ref).max() / max(1.0, np.abs(ref).max())))
            # the projected-gradient norms alsCP logs are global quantities: every rank must have logged the same ones
            pn = torch.tensor([r[1] for r in t.rows], dtype=torch.float64, device="cuda")
            lo, hi = pn.clone(), pn.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            ok = worst_fac <= 1e-8 and len(t.rows) == 2 and bool(torch.all(hi - lo <= 1e-12 * hi.abs().clamp(min=1.0)))
            print(f"rank {rank} lens {lens} R {R} plain alsCP: {'OK' if ok else 'MISMATCH'} factor_err {worst_fac:.2e} "
                  f"projnorm {[float(x) for x in pn]}", flush=True)
            ok_all &= ok
            for x in [Vd] + Wd + Gd + Fd:
                x.free()
            continue
        if driver == "DT":
            _, tr = o.alsCP_DT(V, W_ref, G_ref, 1e-10 * vnorm, 12, resprint=4, F=None)
        elif driver == "PP":
            _, tr = o.alsCP_PP(V, W_ref, G_ref, 1e-10 * vnorm, tol_init, maxiter, resprint=4)
        else:
            _, tr = o.alsCP_PP_partupdate(V, W_ref, G_ref, 1e-10 * vnorm, tol_init, maxiter, update_percentage=1.0,
                                          resprint=4)
        Vd = H.Tensor.from_numpy(world, np.ascontiguousarray(V[b:e]))
        Wd = [H.Tensor.from_numpy(world, (w[b:e] if i == 0 else w), matrix=True) for i, w in enumerate(W)]
        Gd = [H.Tensor.from_numpy(world, (g[b:e] if i == 0 else g), matrix=True) for i, g in enumerate(G)]
        Fd = [H.Matrix(world, w.lens[0], R) for w in Wd]
        with H.Trace(quiet=True) as t:
            if driver == "DT":
                H.alsCP_DT(world, Vd, Wd, Gd, Fd, 1e-10 * vnorm, 12, resprint=4)
            elif driver == "PP":
                H.alsCP_PP(world, Vd, Wd, Gd, Fd, 1e-10 * vnorm, tol_init, maxiter, resprint=4)
            else:
                H.alsCP_PP_partupdate(world, Vd, Wd, Gd, Fd, 1e-10 * vnorm, tol_init, maxiter, update_percentage=1.0,
                                      resprint=4)
        ok = len(t.rows) == len(tr.rows)
        worst_fit = worst_fac = 0.0
        for rg, rr in zip(t.rows, tr.rows):
            ok &= int(rg[0]) == rr[0] and abs(rg[3] - rr[3]) <= 1e-10 * vnorm and abs(rg[1] - rr[1]) <= 1e-9 * max(rr[1], 1e-6 * vnorm)
            worst_fit = max(worst_fit, abs(rg[3] - rr[3]) / vnorm)
        if driver != "DT":
            ok &= t.events == [(0 if k == "DT" else 1, it) for k, it in tr.events]
        for i in range(N):
            ref = W_ref[i][b:e] if i == 0 else W_ref[i]
            err = float(np.abs(Wd[i].numpy() - ref).max() / max(1.0, np.abs(ref).max()))
            worst_fac = max(worst_fac, err)
            ok &= err <= 1e-8
        print(f"rank {rank} lens {lens} R {R} {driver}: {'OK' if ok else 'MISMATCH'} rows {len(t.rows)} "
              f"fit_err {worst_fit:.2e} factor_err {worst_fac:.2e} events {t.events if driver != 'DT' else ''}", flush=True)
        ok_all &= ok
        for x in [Vd] + Wd + Gd + Fd:
            x.free()
# ---- Tucker: hosvd + alsTucker_DT / alsTucker_PP on the sharded tensor (factors replicated) -------------
def proj_err(A, B):
    return float(np.abs(A @ A.T - B @ B.T).max())


for lens, R, tol_init in [((12, 13, 14), 3, 0.3), ((9, 10, 8, 7), 3, 0.3), ((3, 13, 14), 3, 0.3), ((3, 7, 6, 5), 2, 0.3),
                          ((6, 13, 14), 3, 0.3)]:
    N = len(lens)
    if lens[0] < nranks:
        lens = (nranks,) + tuple(lens[1:])
    b, e = ppx.shard_range(lens[0], nranks, rank)
    world.set_shard(0, lens[0], b, e)
    V = o.make_tensor_r2(lens)
    vnorm = np.linalg.norm(V)
    core_ref, W_ref = o.hosvd(V, [R] * N)
    Vd = H.Tensor.from_numpy(world, np.ascontiguousarray(V[b:e]))
    Wd = [H.Matrix(world, lens[i], R) for i in range(N)]
    cored = H.Tensor(world, (R,) * N)
    H.hosvd(world, Vd, cored, Wd, [R] * N)
    hosvd_err = [proj_err(Wd[i].numpy(), W_ref[i]) for i in range(N)]
    ok = all(x < 1e-8 for x in hosvd_err)
    ok &= abs(np.linalg.norm(cored.numpy()) - np.linalg.norm(core_ref)) < 1e-10 * vnorm
    if not ok and rank == 0:
        print(f"  Tucker lens {lens}: hosvd itself differs: proj_err {hosvd_err} core norm {np.linalg.norm(cored.numpy())} "
              f"vs {np.linalg.norm(core_ref)}", flush=True)
    W2 = [w.copy() for w in W_ref]
    ok_ref, rows_ref, _ = o.alsTucker_DT(V, core_ref, W2, 1e-10 * vnorm, 12, resprint=4)
    with H.Trace(quiet=True) as t:
        okd = H.alsTucker_DT(world, Vd, cored, Wd, 1e-10 * vnorm, 12, resprint=4)
    ok &= okd == ok_ref and len(t.rows) == len(rows_ref)
    worst = 0.0
    for rg, rr in zip(t.rows, rows_ref):
        ok &= int(rg[0]) == rr[0] and abs(rg[1] - rr[1]) <= 1e-9 * vnorm and abs(rg[3] - rr[2]) <= 1e-10 * vnorm
        worst = max(worst, abs(rg[3] - rr[2]) / vnorm)
    ok &= all(proj_err(Wd[i].numpy(), W2[i]) < 1e-7 for i in range(N))
    print(f"rank {rank} Tucker lens {lens} R {R} hosvd+DT: {'OK' if ok else 'MISMATCH'} rows {len(t.rows)} "
          f"fit_err {worst:.2e}", flush=True)
    if not ok and rank == 0:  # what differs: the HOSVD subspaces, the core, or the logged rows
        print("  hosvd proj_err", [proj_err(Wd[i].numpy(), W2[i]) for i in range(N)], "rows gpu", [tuple(r[:4]) for r in t.rows],
              "rows ref", rows_ref, flush=True)
    ok_all &= ok
    for x in Wd + [cored]:
        x.free()
    # PP from the oracle's HOSVD factors (column signs agree from the first sweep on)
    Wd = [H.Tensor.from_numpy(world, w, matrix=True) for w in W_ref]
    cored = H.Tensor.from_numpy(world, core_ref)
    W2 = [w.copy() for w in W_ref]
    ok_ref, rows_ref, ev_ref, _, _ = o.alsTucker_PP(V, core_ref, W2, 1e-10 * vnorm, tol_init, 30, resprint=5)
    with H.Trace(quiet=True) as t:
        okp = H.alsTucker_PP(world, Vd, cored, Wd, 1e-10 * vnorm, tol_init, 30, resprint=5)
    ok = okp == ok_ref and t.events == [(0 if k == "DT" else 1, it) for k, it in ev_ref] and len(t.rows) == len(rows_ref)
    worst = 0.0
    for rg, rr in zip(t.rows, rows_ref):
        ok &= int(rg[0]) == rr[0] and int(rg[2]) == rr[2] and abs(rg[1] - rr[1]) <= 1e-9 * vnorm
        ok &= abs(rg[3] - rr[3]) <= 1e-10 * vnorm
        worst = max(worst, abs(rg[3] - rr[3]) / vnorm)
    ok &= all(proj_err(Wd[i].numpy(), W2[i]) < 1e-7 for i in range(N))
    print(f"rank {rank} Tucker lens {lens} R {R} PP: {'OK' if ok else 'MISMATCH'} rows {len(t.rows)} fit_err {worst:.2e} "
          f"events {t.events}", flush=True)
    ok_all &= ok
    for x in [Vd] + Wd + [cored]:
        x.free()

flag = torch.tensor([1 if ok_all else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("MULTI-GPU PARITY", "PASS" if int(flag.item()) == 1 else "FAIL", flush=True)
world.close()
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 1 else 1)
