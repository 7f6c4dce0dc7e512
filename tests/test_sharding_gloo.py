"""World-size-2 test of the multi-GPU design on CPU (gloo): the tensor is sharded along mode 0, every rank runs the
oracle's contractions on its slab, and the ONLY exchanges are the all-reduce of the s x R partial MTTKRPs of the
non-sharded modes and of the R x R Gram of the sharded factor.  The sharded sweep must
reproduce the unsharded oracle sweep.  Shard ranges come from the C ABI's host function ppx_shard_range."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _sharded_sweep(rank, nranks, port, lens, R, out_dir):
    sys.path.insert(0, ROOT)
    from oracle import pp_oracle as o

    ppx = importlib.import_module("pairwise-perturbation_b200")
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=nranks)
    N = len(lens)
    V, _ = o.make_tensor_r(lens, R)
    W = o.init_factors(lens, R)
    b, e = ppx.shard_range(lens[0], nranks, rank)
    Vl = V[b:e]                       # local slab of mode 0
    Wl = [w.copy() for w in W]
    Wl[0] = W[0][b:e].copy()          # rows of W_0 are sharded the same way; the other factors are replicated

    def allreduce(x):
        t = torch.from_numpy(np.ascontiguousarray(x))
        dist.all_reduce(t)
        return t.numpy()

    parent, sibling = {}, {}
    o.construct_dimension_tree(parent, sibling, 0, N - 1)
    grams = [w.T @ w for w in Wl]
    grams[0] = allreduce(grams[0])
    n_allreduce = 1
    for sweep in range(2):
        mm = {}
        for i in range(N):
            M = o._leaf_M(mm, parent, sibling, Vl, Wl, i)
            if i != 0:
                M = allreduce(M)      # partial sums over the local slab
                n_allreduce += 1
            S = np.ones((R, R))
            for j in range(N):
                if j != i:
                    S = S * grams[j]
            Wl[i] = o.cholesky_solve(M, S)
            grams[i] = Wl[i].T @ Wl[i]
            if i == 0:
                grams[0] = allreduce(grams[0])
                n_allreduce += 1
        norms = [np.sqrt(np.trace(g)) for g in grams]   # global norms from the (reduced) Grams
        gm = np.prod(norms) ** (1.0 / N)
        for i in range(N):
            Wl[i] = Wl[i] * (gm / norms[i])
            grams[i] = grams[i] * (gm / norms[i]) ** 2
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), b=b, e=e, n_allreduce=n_allreduce,
             **{"W%d" % i: w for i, w in enumerate(Wl)})
    dist.destroy_process_group()


@pytest.mark.parametrize("lens,R", [((9, 6, 5, 4), 3), ((7, 5, 6), 2)])
def test_mode0_sharded_sweep_equals_unsharded(tmp_path, lens, R):
    sys.path.insert(0, ROOT)
    from oracle import pp_oracle as o

    nranks = 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_sharded_sweep, args=(nranks, port, lens, R, str(tmp_path)), nprocs=nranks, join=True)
    N = len(lens)
    V, _ = o.make_tensor_r(lens, R)
    W = o.init_factors(lens, R)
    parent, sibling = {}, {}
    o.construct_dimension_tree(parent, sibling, 0, N - 1)
    for sweep in range(2):
        mm = {}
        for i in range(N):
            M = o._leaf_M(mm, parent, sibling, V, W, i)
            S = o.gram_hadamard(W, i)
            W[i] = o.cholesky_solve(M, S)
        o.normalize(W)
    parts = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % r)) for r in range(nranks)]
    W0 = np.concatenate([p["W0"] for p in parts], axis=0)
    assert np.abs(W0 - W[0]).max() < 1e-10
    for i in range(1, N):
        for p in parts:
            assert np.abs(p["W%d" % i] - W[i]).max() < 1e-10   # replicated factors agree on every rank
    # N-1 MTTKRP all-reduces + 1 Gram all-reduce per sweep (+1 initial Gram)
    assert int(parts[0]["n_allreduce"]) == 1 + 2 * N
    assert sum(int(p["e"]) - int(p["b"]) for p in parts) == lens[0]


# ---- Tucker: V sharded along mode 0, factors replicated -------------------------------------------
def _sharded_tucker(rank, nranks, port, lens, R, out_dir):
    sys.path.insert(0, ROOT)
    from oracle import pp_oracle as o

    ppx = importlib.import_module("pairwise-perturbation_b200")
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=nranks)
    N = len(lens)
    V = o.make_tensor_r2(lens)
    b, e = ppx.shard_range(lens[0], nranks, rank)
    Vl = V[b:e]
    n_allreduce = 0

    def allreduce(x):
        nonlocal n_allreduce
        n_allreduce += 1
        t = torch.from_numpy(np.ascontiguousarray(x))
        dist.all_reduce(t)
        return t.numpy()

    def gather_mode0(Yl):     # zero-padded all-reduce, exactly what the host code does
        full = np.zeros((lens[0],) + Yl.shape[1:])
        full[b:e] = Yl
        return allreduce(full)

    ttm_saved = o.ttm

    def local_ttm(T, x, Wx):  # a contraction of mode 0 sums over the local rows only
        return ttm_saved(T, x, Wx[b:e] if x == 0 else Wx)

    # HOSVD: modes != 0 all-reduce the local Gram.  Mode 0 (gram_mode0_sharded in host/als_Tucker.cxx): ONE personalised
    # exchange -- rank k receives, from every rank, that rank's rows of k's share [tb_k, te_k) of the LAST mode -- then
    # the Gram of the full-height panel and one all-reduce of the partial Grams
    n_exchange = 0
    W = []
    for i in range(N):
        if i != 0:
            MTM = allreduce(o.unroll_tensor_contraction(Vl, i))
        else:
            tr = [ppx.shard_range(lens[-1], nranks, k) for k in range(nranks)]
            rr = [ppx.shard_range(lens[0], nranks, k) for k in range(nranks)]
            send = [torch.from_numpy(np.ascontiguousarray(Vl[..., tr[k][0]:tr[k][1]])) for k in range(nranks)]
            recv = [torch.empty((rr[k][1] - rr[k][0],) + tuple(lens[1:-1]) + (tr[rank][1] - tr[rank][0],),
                                dtype=torch.float64) for k in range(nranks)]
            reqs = []
            for k in range(nranks):
                if k == rank:
                    recv[k].copy_(send[k])
                else:
                    reqs.append(dist.isend(send[k], k))
                    reqs.append(dist.irecv(recv[k], k))
            for q in reqs:
                q.wait()
            n_exchange += 1
            full = np.concatenate([t.numpy() for t in recv], axis=0)   # all rows of my share of the last mode
            MTM = allreduce(o.unroll_tensor_contraction(full, 0))
        W.append(o.top_left_singular(MTM, R))
    # two HOOI sweeps with the dimension tree
    parent, sibling = {}, {}
    o.construct_dimension_tree(parent, sibling, 0, N - 1)
    o.ttm = local_ttm         # the tree code calls o.ttm for every edge
    try:
        for sweep in range(2):
            tm = {}
            for i in range(N):
                Y = o._tucker_leaf_Y(tm, parent, sibling, Vl, W, i)
                Y = allreduce(Y) if i != 0 else gather_mode0(Y)
                W[i] = o.top_left_singular(o.unroll_tensor_contraction(Y, i), R)
    finally:
        o.ttm = ttm_saved
    np.savez(os.path.join(out_dir, "trank%d.npz" % rank), n_allreduce=n_allreduce, n_exchange=n_exchange,
             **{"W%d" % i: w for i, w in enumerate(W)})
    dist.destroy_process_group()


@pytest.mark.parametrize("lens,R", [((8, 7, 6), 2), ((6, 5, 4, 5), 2)])
def test_mode0_sharded_hooi_equals_unsharded(tmp_path, lens, R):
    sys.path.insert(0, ROOT)
    from oracle import pp_oracle as o

    nranks = 2
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_sharded_tucker, args=(nranks, port, lens, R, str(tmp_path)), nprocs=nranks, join=True)
    N = len(lens)
    V = o.make_tensor_r2(lens)
    _, W = o.hosvd(V, [R] * N)
    parent, sibling = {}, {}
    o.construct_dimension_tree(parent, sibling, 0, N - 1)
    for sweep in range(2):
        tm = {}
        for i in range(N):
            Y = o._tucker_leaf_Y(tm, parent, sibling, V, W, i)
            W[i] = o.top_left_singular(o.unroll_tensor_contraction(Y, i), R)
    parts = [np.load(os.path.join(str(tmp_path), "trank%d.npz" % r)) for r in range(nranks)]
    for i in range(N):
        for p in parts:       # replicated factors: same subspace on every rank as the unsharded run
            A, B = p["W%d" % i], W[i]
            assert np.abs(A @ A.T - B @ B.T).max() < 1e-9
    # HOSVD: N Gram all-reduces + ONE personalised exchange for mode 0; HOOI: one exchange per mode update
    assert int(parts[0]["n_allreduce"]) == N + 2 * N
    assert int(parts[0]["n_exchange"]) == 1
