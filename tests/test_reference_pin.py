"""Pins oracle/pp_oracle.py against THE REFERENCE'S OWN CODE: oracle/_ref/ holds the reference's unmodified sources
(common.cxx, als_CP.cxx, als_Tucker.cxx, test_ALS.cxx, run.cxx, src/**, tests/test_decomposition.cxx) compiled against
oracle/ctf_standin/ctf.hpp, a dense single-process stand-in for the CTF/MPI API they use (`make -C oracle ref`).

Same bytes in (tensor, initial factors, initial gradient), then: per-print gradient norm and residual within 1e-10
relative to ||V||, factors within 1e-8, identical DT<->PP switching iterations -- the bar BASELINE.json sets.
What this pins: control flow, contraction strings and layouts, the quirks listed in DESIGN.md section 2, solves,
normalisation, the switching logic.  What it cannot pin: CTF's own arithmetic (summation order, ScaLAPACK's SVD).

CPU only; needs the prebuilt oracle/_ref/ (it travels with the repo), never /root/reference at run time."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pp_oracle as o  # noqa: E402
from oracle import ref_harness as rh  # noqa: E402

pytestmark = pytest.mark.skipif(not rh.available(), reason="oracle/_ref not built (needs /root/reference once)")

FIT_RTOL, FACTOR_TOL = 1e-10, 1e-8


def problem(lens, R, seed=1):
    V, _ = o.make_tensor_r(lens, R, seed=seed)
    return V, o.init_factors(lens, R, seed=seed + 1), o.init_grad(lens, R, seed=seed + 2)


def check_rows(ref_rows, rows, vnorm):
    assert len(ref_rows) == len(rows) and len(rows) > 0
    a, b = np.array(ref_rows, dtype=float), np.array([(r[0], r[1], r[2], r[3]) for r in rows], dtype=float)
    assert np.array_equal(a[:, 0], b[:, 0]) and np.array_equal(a[:, 2], b[:, 2])  # iterations, pp_update flags
    assert np.allclose(a[:, 3], b[:, 3], rtol=0, atol=FIT_RTOL * vnorm)  # residual ("fitness")
    assert np.allclose(a[:, 1], b[:, 1], rtol=1e-9, atol=1e-9 * vnorm)  # gradient norm


def check_factors(Wref, W):
    for a, b in zip(Wref, W):
        assert np.abs(a - b).max() < FACTOR_TOL


# ---- the pieces (one call each of the reference's functions) ----------------------------------------------------
@pytest.mark.parametrize("lens,R", [((7, 6, 5, 4), 3), ((6, 6, 6, 6), 3), ((4, 3, 5, 3, 4, 3), 2),
                                    ((5, 4, 3, 4, 3, 2, 3), 2)])
def test_reference_pieces_match_oracle(lens, R):
    """mttkrp_map_DT, Build_mttkrp_map, KhatriRao_contract, unroll_tensor_contraction, TTMc, build_V, Normalize,
    SVD_solve, cholesky_solve of the reference against the restatement (orders 4, 6, 7)."""
    N = len(lens)
    V = o.fill_uniform(lens, 7, 0, -1.0, 1.0)
    W = [o.fill_uniform((lens[i], R), 8, i) for i in range(N)]
    ref = rh.run_driver("kernels", V, W)
    f = ref["files"]
    parent, sibling = {}, {}
    o.construct_dimension_tree(parent, sibling, 0, N - 1)
    mp = {}
    n_tree = 0
    for name in parent:
        if len(name) == N or len(name) < 2:
            continue
        o.mttkrp_map_DT(mp, parent, sibling, V, W, name)
    for name, T in mp.items():
        assert np.allclose(f["tree_" + name], T.ravel(order="F"), rtol=1e-12, atol=1e-12), name
        n_tree += 1
    assert n_tree == len([k for k in f if k.startswith("tree_")]) and n_tree >= 2
    ops = o.build_pp_operators(V, W)
    n_pp = 0
    for name, T in ops.items():
        assert np.allclose(f["pp_" + name], T.ravel(order="F"), rtol=1e-12, atol=1e-12), name
        n_pp += 1
    assert n_pp == len([k for k in f if k.startswith("pp_")])
    for i in range(N):
        assert np.allclose(f["gram_%d" % i], o.unroll_tensor_contraction(V, i).ravel(order="F"), rtol=1e-12)
        assert np.allclose(f["ttmc_%d" % i], o.TTMc(V, W, i).ravel(order="F"), rtol=1e-12, atol=1e-12)
        if "krc_%d" % i in f:  # uniform extents only: the reference's callers mis-size lens_H otherwise
            index = [m for m in range(N) if m != i] + [i]
            assert np.allclose(f["krc_%d" % i], o.KhatriRao_contract(V, W, index).ravel(order="F"), rtol=1e-12,
                               atol=1e-12)
    assert np.allclose(f["build_V"], o.build_V(W).ravel(order="F"), rtol=1e-13)
    S = o.gram_hadamard(W, 0)
    assert np.allclose(f["S"], S.ravel(order="F"), rtol=1e-13)
    assert np.allclose(f["svd_solve"], o.SVD_solve(W[0], S).ravel(order="F"), rtol=1e-8)
    assert np.allclose(f["cholesky_solve"], o.cholesky_solve(W[0], S).ravel(order="F"), rtol=1e-8)
    Wn = [w.copy() for w in W]
    o.normalize(Wn)
    check_factors(ref["normalized"], Wn)


# ---- the drivers ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("lens,R,lam,maxiter", [((9, 8, 7, 6), 3, 0.0, 25), ((6, 5, 4, 5, 4, 3), 2, 1e-3, 15)])
def test_reference_alsCP_DT(lens, R, lam, maxiter):
    V, W, G = problem(lens, R)
    vnorm = np.linalg.norm(V)
    ref = rh.run_driver("alsCP_DT", V, W, G, tol=1e-10 * vnorm, maxiter=maxiter, lambda_=lam, resprint=5)
    tr = o.Trace()
    o.alsCP_DT(V, W, G, 1e-10 * vnorm, maxiter, lam=lam, resprint=5, trace=tr)
    check_rows(ref["rows"], tr.rows, vnorm)
    check_factors(ref["W"], W)
    check_factors(ref["grad"], G)


@pytest.mark.parametrize("lens,R,tol_init,ratio,lam,maxiter",
                         [((10, 10, 10, 10), 3, 0.1, 1.0, 0.0, 60), ((9, 14, 5, 11), 4, 0.05, 1.0, 0.0, 40),
                          ((8, 7, 9, 6), 3, 0.2, 0.8, 1e-4, 40), ((5, 4, 5, 4, 5, 4), 2, 0.1, 1.0, 0.0, 30)])
def test_reference_alsCP_PP(lens, R, tol_init, ratio, lam, maxiter):
    """The switching driver: identical `DT starts from` / `pairwise perturbation starts from` iterations."""
    V, W, G = problem(lens, R)
    vnorm = np.linalg.norm(V)
    ref = rh.run_driver("alsCP_PP", V, W, G, tol=1e-10 * vnorm, tol_init=tol_init, maxiter=maxiter, lambda_=lam,
                        ratio_step=ratio, resprint=5)
    _, tr = o.alsCP_PP(V, W, G, 1e-10 * vnorm, tol_init, maxiter, lam=lam, ratio_step=ratio, resprint=5)
    assert ref["events"] == tr.events and any(k == "PP" for k, _ in tr.events)
    check_rows(ref["rows"], tr.rows, vnorm)
    check_factors(ref["W"], W)


@pytest.mark.parametrize("lens,R,pct", [((9, 8, 10, 7), 3, 1.0), ((9, 8, 10, 7), 3, 0.5), ((5, 4, 5, 4, 5, 4), 2, 0.7)])
def test_reference_alsCP_PP_partupdate(lens, R, pct):
    V, W, G = problem(lens, R)
    vnorm = np.linalg.norm(V)
    ref = rh.run_driver("alsCP_PP_partupdate", V, W, G, tol=1e-10 * vnorm, tol_init=0.1, maxiter=30, update_pct=pct,
                        resprint=5)
    _, tr = o.alsCP_PP_partupdate(V, W, G, 1e-10 * vnorm, 0.1, 30, update_percentage=pct, resprint=5)
    assert ref["events"] == tr.events
    check_rows(ref["rows"], tr.rows, vnorm)
    check_factors(ref["W"], W)


def projector_err(A, B):
    return np.abs(A @ A.T - B @ B.T).max()


@pytest.mark.parametrize("lens,ranks", [((9, 8, 7, 6), (3, 3, 3, 3)), ((8, 9, 6, 7), (2, 3, 2, 4))])
def test_reference_tucker(lens, ranks):
    """hosvd, alsTucker_DT, alsTucker_PP.  Singular vectors carry an arbitrary sign, so factors are compared through
    their projectors W W^T and the printed core-norm differences / residuals directly."""
    V = o.make_tensor_r2(lens, seed=1)
    vnorm = np.linalg.norm(V)
    ref = rh.run_driver("hosvd", V, ranks=ranks)
    core, W = o.hosvd(V, list(ranks))
    for a, b in zip(ref["W"], W):
        assert projector_err(a, b) < 1e-9
    assert abs(np.linalg.norm(ref["core"]) - np.linalg.norm(core)) < 1e-10 * vnorm

    ref = rh.run_driver("alsTucker_DT", V, ranks=ranks, tol=1e-10 * vnorm, maxiter=12, resprint=4)
    core, W = o.hosvd(V, list(ranks))
    _, rows, core = o.alsTucker_DT(V, core, W, 1e-10 * vnorm, 12, resprint=4)
    assert [r[0] for r in ref["rows"]] == [r[0] for r in rows]
    assert np.allclose([r[3] for r in ref["rows"]], [r[2] for r in rows], rtol=0, atol=FIT_RTOL * vnorm)
    assert np.allclose([r[1] for r in ref["rows"]], [r[1] for r in rows], rtol=0, atol=1e-9 * vnorm)
    for a, b in zip(ref["W"], W):
        assert projector_err(a, b) < 1e-7

    ref = rh.run_driver("alsTucker_PP", V, ranks=ranks, tol=1e-10 * vnorm, tol_init=0.3, maxiter=16, resprint=4)
    core, W = o.hosvd(V, list(ranks))
    _, rows, events, sweeps, core = o.alsTucker_PP(V, core, W, 1e-10 * vnorm, 0.3, 16, resprint=4)
    assert ref["events"] == events and any(k == "PP" for k, _ in events)
    assert [r[0] for r in ref["rows"]] == [r[0] for r in rows]
    assert np.allclose([r[3] for r in ref["rows"]], [r[-1] for r in rows], rtol=0, atol=FIT_RTOL * vnorm)
    for a, b in zip(ref["W"], W):
        assert projector_err(a, b) < 1e-7


# ---- the reference's own mains -----------------------------------------------------------------------------------
def cli_fills(N, tensor="r"):
    """fill_random calls of test_ALS.cxx in order: `r` -> W_true[0..N-1] (:279-282), `r2` -> V (:272); then W[i],
    grad_W[i] interleaved (:337-338).  Mapped to the (seed, id) pairs our CLI and the oracle use."""
    head = [(1, i) for i in range(N)] if tensor == "r" else [(1, 100)]
    return head + [p for i in range(N) for p in ((2, i), (3, i))]


@pytest.mark.parametrize("N,s,R,pp,maxiter", [(4, 10, 3, 1, 60), (4, 12, 4, 0, 20), (6, 5, 2, 1, 30), (4, 9, 3, 2, 30)])
def test_reference_test_ALS_main_matches_oracle(N, s, R, pp, maxiter):
    """The reference's test_ALS executable, flags as in BASELINE.json configs, against the restatement."""
    lens = (s,) * N
    ref = rh.run_cli("test_ALS", ["-model", "CP", "-tensor", "r", "-dim", N, "-size", s, "-rank", R, "-pp", pp,
                                  "-maxiter", maxiter, "-pp_res_tol", 0.1, "-resprint", 5], fills=cli_fills(N))
    V, W, G = problem(lens, R)
    vnorm = np.linalg.norm(V)
    if pp == 0:
        tr = o.Trace()
        o.alsCP_DT(V, W, G, 1e-10 * vnorm, maxiter, resprint=5, trace=tr)
    elif pp == 1:
        _, tr = o.alsCP_PP(V, W, G, 1e-10 * vnorm, 0.1, maxiter, resprint=5)
    else:
        _, tr = o.alsCP_PP_partupdate(V, W, G, 1e-10 * vnorm, 0.1, maxiter, update_percentage=1.0, resprint=5)
    assert ref["events"] == tr.events
    a = np.array(ref["rows"])
    b = np.array([(r[0], r[1], r[2], r[3]) for r in tr.rows])
    assert a.shape == b.shape
    # the main prints 13 significant digits
    assert np.array_equal(a[:, 0], b[:, 0]) and np.array_equal(a[:, 2], b[:, 2])
    assert np.allclose(a[:, 3], b[:, 3], rtol=1e-9, atol=FIT_RTOL * vnorm)
    assert np.allclose(a[:, 1], b[:, 1], rtol=1e-8, atol=1e-9 * vnorm)


def test_reference_test_ALS_tucker_main_matches_oracle():
    N, s, R = 4, 8, 3
    ref = rh.run_cli("test_ALS", ["-model", "Tucker", "-tensor", "r2", "-dim", N, "-size", s, "-rank", R, "-pp", 1,
                                  "-maxiter", 16, "-pp_res_tol", 0.3, "-resprint", 4], fills=cli_fills(N, "r2"))
    V = o.make_tensor_r2((s,) * N, seed=1)
    vnorm = np.linalg.norm(V)
    core, W = o.hosvd(V, [R] * N)
    _, rows, events, sweeps, core = o.alsTucker_PP(V, core, W, 1e-10 * vnorm, 0.3, 16, resprint=4)
    assert ref["events"] == events
    assert [r[0] for r in ref["rows"]] == [r[0] for r in rows]
    assert np.allclose([r[3] for r in ref["rows"]], [r[-1] for r in rows], rtol=1e-9, atol=FIT_RTOL * vnorm)


def test_reference_own_test_runs():
    """tests/test_decomposition.cxx, the reference's single test (asserts on order / rank, then runs CPD::als with
    each optimizer): must run to completion on the stand-in."""
    out = rh.run_cli("test_decomposition", [])
    assert "Iters" in out["stdout"] or "sweeps" in out["stdout"]


# ---- the OO path: the reference's `run` main (src/CP.cxx + src/optimizer/**) --------------------------------------
@pytest.mark.parametrize("N,s,R,pp,extra", [(4, 8, 3, 0, []), (3, 9, 3, 0, []), (4, 8, 3, 1, []), (5, 5, 2, 1, []),
                                            (4, 7, 3, 4, []), (4, 8, 4, 2, ["-updaterank", 2]),
                                            (4, 8, 4, 3, ["-updaterank", 2])])
def test_reference_run_main_matches_oracle(N, s, R, pp, extra):
    """run -pp 0/1/4/2/3 = CPD<double, CPDTOptimizer | CPMSDTOptimizer | CPSimpleOptimizer | CPDTLROptimizer |
    CPMSDTLROptimizer>::als (run.cxx:387-414).  fill_random calls: W_true[i] (run.cxx:313), W[i] (:374), then
    grad_W[i] in CPD::Init (src/CP.cxx:78)."""
    fills = [(1, i) for i in range(N)] + [(2, i) for i in range(N)] + [(3, i) for i in range(N)]
    maxiter = 12
    ref = rh.run_cli("run", ["-model", "CP", "-tensor", "r", "-dim", N, "-size", s, "-rank", R, "-pp", pp,
                             "-maxiter", maxiter, "-resprint", 3] + extra, fills=fills)
    lens = (s,) * N
    V, _ = o.make_tensor_r(lens, R, seed=1)
    W = o.init_factors(lens, R, seed=2)
    vnorm = np.linalg.norm(V)
    cls = {0: o.CPDTOptimizer, 1: o.CPMSDTOptimizer, 4: o.CPSimpleOptimizer, 2: o.CPDTLROptimizer,
           3: o.CPMSDTLROptimizer}[pp]
    args = (2, 0) if pp in (2, 3) else ()
    d = o.CPD(N, s, R, cls, *args)
    d.Init(V, W, grad_W=o.init_grad(lens, R, seed=3))
    _, rows = d.als(1e-10 * vnorm, maxiter, 3)
    a, b = np.array(ref["rows"])[:, [0, 1, 3]], np.array(rows)
    assert a.shape == b.shape and len(rows) >= 3
    assert np.allclose(a[:, 0], b[:, 0])  # sweep counts (fractions: 0.5 / (N-1)/N / 1 per step)
    assert np.allclose(a[:, 2], b[:, 2], rtol=1e-9, atol=FIT_RTOL * vnorm)
    assert np.allclose(a[:, 1], b[:, 1], rtol=1e-8, atol=1e-9 * vnorm)


# ---- the input generators of test_ALS.cxx (-tensor p / p2 / c), through the reference's main ------------------------
@pytest.mark.parametrize("tensor,dim,size,R", [("p2", 4, 5, 3), ("p", 8, 3, 3), ("c", 4, 7, 3)])
def test_reference_generators_match_oracle(tensor, dim, size, R):
    """laplacian_tensor + fold_unfold (common.cxx:575-642, 870-882), Gen_collinearity + noise (common.cxx:361-423,
    test_ALS.cxx:245-261): the reference's main and the restatement must print the same ||V|| and the same ALS trace.
    Gen_collinearity redraws vectors until their collinearity fits, so the number of fill_random calls is data
    dependent: the restatement counts its draws and the schedule handed to the stand-in follows it."""
    if tensor == "c":
        lens = (size,) * dim
        _, vec = o.gen_collinearity(lens, R, 0.5, 0.9, seed=1)
        V = o.make_tensor_c(lens, R, 0.5, 0.9, 0.01, seed=1)
        # count the draws the restatement made: rerun with a counting generator
        n_draws = [0]
        orig = o.u01

        def counting(seed, tid, n, start=0):
            if tid >= 1000:
                n_draws[0] += 1
            return orig(seed, tid, n, start)

        o.u01 = counting
        try:
            o.gen_collinearity(lens, R, 0.5, 0.9, seed=1)
        finally:
            o.u01 = orig
        head = [(1, 1000 + k) for k in range(n_draws[0])] + [(1, 101)]
    elif tensor == "p2":
        V = o.make_tensor_p(dim, size, folded=False)
        head = []
    else:
        V = o.make_tensor_p(dim, size, folded=True)
        head = []
    N = V.ndim
    lens = V.shape
    fills = head + [p for i in range(N) for p in ((2, i), (3, i))]
    ref = rh.run_cli("test_ALS", ["-model", "CP", "-tensor", tensor, "-dim", dim, "-size", size, "-rank", R, "-pp", 0,
                                  "-maxiter", 10, "-resprint", 5], fills=fills)
    vnorm = np.linalg.norm(V)
    import re
    printed = float(re.search(r"Vnorm= (\S+)", ref["stdout"]).group(1))
    assert abs(printed - vnorm) < 1e-5 * vnorm  # the main prints 6 digits
    W, G = o.init_factors(lens, R), o.init_grad(lens, R)
    tr = o.Trace()
    o.alsCP_DT(V, W, G, 1e-10 * vnorm, 10, resprint=5, trace=tr)
    a = np.array(ref["rows"])
    b = np.array([(r[0], r[1], r[2], r[3]) for r in tr.rows])
    assert a.shape == b.shape
    assert np.allclose(a[:, 3], b[:, 3], rtol=1e-9, atol=FIT_RTOL * vnorm)
    assert np.allclose(a[:, 1], b[:, 1], rtol=1e-8, atol=1e-9 * vnorm)


# ---- who checks the checker: the CTF stand-in on its own against NumPy ---------------------------------------------
def test_ctf_standin_against_numpy(tmp_path):
    """oracle/_ref/standin_selftest (oracle/standin_selftest.cxx) runs expressions of every kind the reference uses --
    tensor-times-matrix, Hadamard-batched contraction, multi-term sums with += / -= and scalar factors, a right-hand
    side that mentions the output, diagonal assignment / extraction, Transform, scalar conversion, the unfolding Gram with
    the '^' '&' index characters, svd (tall and wide), qr, cholesky, solve_tri in its four variants -- through
    oracle/ctf_standin/ctf.hpp alone; NumPy recomputes them from the same counter-based inputs."""
    import subprocess
    exe = os.path.join(rh.REF_DIR, "standin_selftest")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/standin_selftest not built")
    pre = os.path.join(str(tmp_path), "st")
    subprocess.run([exe, pre], check=True, env=dict(os.environ, CTF_STANDIN_SEED="11"), capture_output=True, timeout=60)

    def ld(name, shape):
        return np.fromfile(pre + "." + name + ".bin").reshape(shape, order="F")

    V = o.fill_uniform((5, 4, 3, 6), 11, 0, -1.0, 1.0)
    A, B, C, D = (o.fill_uniform(sh, 11, k) for k, sh in ((1, (5, 3)), (2, (4, 3)), (3, (3, 3)), (4, (6, 3))))
    assert np.array_equal(ld("V", V.shape), V)
    T1 = np.einsum("abcd,cr->abdr", V, C)
    assert np.allclose(ld("ttm", T1.shape), T1, rtol=1e-13, atol=1e-14)
    T2 = np.einsum("abdr,dr->abr", T1, D)
    assert np.allclose(ld("mttv", T2.shape), T2, rtol=1e-13, atol=1e-14)
    M = np.einsum("abr,br->ar", T2, B) + 2.0 * A
    M = M - 0.5 * (A - M)
    assert np.allclose(ld("chain", M.shape), M, rtol=1e-13)
    S = (A.T @ A) * (B.T @ B) + 0.25 * np.eye(3)
    assert np.allclose(ld("S", (3, 3)), S, rtol=1e-13)
    dg = np.where(np.diag(S) > 2.0, 1.0, -1.0)
    assert np.array_equal(ld("diag", (3, 3)), np.diag(dg))
    sc = ld("scalars", (2,))
    assert np.isclose(sc[0], np.sum(A * A), rtol=1e-14) and np.isclose(sc[1], np.linalg.norm(V), rtol=1e-14)
    G = np.einsum("ipkl,iqkl->pq", V, V)
    assert np.allclose(ld("unfold_gram", (4, 4)), G, rtol=1e-13)
    for tag, shape in (("svd", (6, 4)), ("svdw", (3, 5))):
        X = ld(tag + "_in", shape)
        U, s, VT = ld(tag + "_U", (shape[0], 3)), ld(tag + "_s", (3,)), ld(tag + "_VT", (3, shape[1]))
        sref = np.linalg.svd(X, compute_uv=False)[:3]
        assert np.allclose(s, sref, rtol=1e-12)
        assert np.abs(U.T @ U - np.eye(3)).max() < 1e-12 and np.abs(VT @ VT.T - np.eye(3)).max() < 1e-12
        Ur, sr, VTr = np.linalg.svd(X, full_matrices=False)
        assert np.allclose((U * s) @ VT, (Ur[:, :3] * sr[:3]) @ VTr[:3], atol=1e-12)  # best rank-3 approximation
    Wd = ld("svd_in", (6, 4))
    Q, R = ld("qr_Q", (6, 4)), ld("qr_R", (4, 4))
    assert np.allclose(Q @ R, Wd, atol=1e-13) and np.abs(Q.T @ Q - np.eye(4)).max() < 1e-13
    assert np.abs(np.tril(R, -1)).max() == 0.0
    L = ld("chol_L", (3, 3))
    assert np.allclose(L, np.linalg.cholesky(S), rtol=1e-13)
    assert np.allclose(ld("tri_right_T", (5, 3)) @ L.T, A, atol=1e-13)
    assert np.allclose(ld("tri_right_N", (5, 3)) @ L, A, atol=1e-13)
    assert np.allclose(L @ ld("tri_left_N", (3, 5)), A.T, atol=1e-13)
    assert np.allclose(L.T @ ld("tri_left_T", (3, 5)), A.T, atol=1e-13)


# ---- the plain (tree-less) drivers of the reference: same normal equations, so the same iterates as the DT drivers ----
def test_reference_plain_alsCP_equals_dt_sweeps():
    """alsCP (als_CP.cxx:20-115: one KhatriRao_contract per mode, SVD_solve, Normalize) produces the factors of the
    dimension-tree sweeps.  Equal extents only: its callers mis-size lens_H for ragged tensors."""
    lens, R = (9, 9, 9, 9), 3
    V, W, G = problem(lens, R)
    ref = rh.run_driver("alsCP", V, W, G, tol=0.0, maxiter=5)
    o.alsCP_DT(V, W, G, 0.0, 5, resprint=100, want_residual=False)
    check_factors(ref["W"], W)


def test_reference_plain_alsTucker_equals_dt_sweeps():
    """alsTucker (als_Tucker.cxx:112-172: TTMc per mode, no tree) against the restated alsTucker_DT: same subspaces."""
    lens, ranks = (8, 9, 7, 6), (3, 3, 3, 3)
    V = o.make_tensor_r2(lens, seed=1)
    vnorm = np.linalg.norm(V)
    ref = rh.run_driver("alsTucker", V, ranks=ranks, tol=1e-10 * vnorm, maxiter=6)
    core, W = o.hosvd(V, list(ranks))
    o.alsTucker_DT(V, core, W, 1e-10 * vnorm, 6, resprint=100, want_residual=False)
    for a, b in zip(ref["W"], W):
        assert projector_err(a, b) < 1e-7


def test_reference_tree_code_cannot_do_order_3(tmp_path):
    """DESIGN.md section 2 / SURVEY 8a: for N = 3 the leaf `c` hangs off the root, mttkrp_map_DT asks for the root's
    parent and recurses for ever (common.cxx:29,89-91) -- the reference's test_ALS overflows its stack, while its OO
    path (run, CPDTOptimizer) handles order 3.  This is why the order-3 fixtures come from the oracle alone and why our
    drivers treat a leaf whose parent is the root by the first-level rule."""
    import subprocess
    args = ["-model", "CP", "-tensor", "r", "-dim", "3", "-size", "6", "-rank", "2", "-pp", "0", "-maxiter", "2"]
    bad = subprocess.run([os.path.join(rh.REF_DIR, "test_ALS")] + args, capture_output=True, timeout=120,
                         cwd=str(tmp_path))
    assert bad.returncode != 0  # SIGSEGV from the unbounded recursion
    good = rh.run_cli("run", args)
    assert "Iters" in good["stdout"]


def test_standin_dgemm_path_matches_the_plain_loops(tmp_path):
    """CTF_STANDIN_BLAS switches the stand-in's contraction engine to DGEMM for products that fold into a matrix product
    (bench.py's CPU arm times the reference's own main that way).  Same main, same inputs, both engines: identical
    switching iterations, printed values equal to rounding."""
    import glob

    hits = glob.glob(os.path.join(os.path.dirname(np.__file__), "..", "numpy.libs", "libscipy_openblas*.so"))
    if not hits:
        pytest.skip("no OpenBLAS shared library beside NumPy")
    N, s, R = 4, 16, 4
    args = ["-model", "CP", "-tensor", "r", "-dim", str(N), "-size", str(s), "-rank", str(R), "-pp", "1", "-maxiter", "30",
            "-pp_res_tol", "0.1", "-resprint", "3", "-filename", os.path.join(str(tmp_path), "a.csv")]
    fills = [(1, i) for i in range(N)] + [p for i in range(N) for p in ((2, i), (3, i))]
    plain = rh.run_cli("test_ALS", args, fills=fills)
    os.environ["CTF_STANDIN_BLAS"] = os.path.abspath(hits[0])
    try:
        fast = rh.run_cli("test_ALS", args, fills=fills)
    finally:
        del os.environ["CTF_STANDIN_BLAS"]
    assert fast["events"] == plain["events"] and len(fast["rows"]) == len(plain["rows"]) >= 5
    for a, b in zip(fast["rows"], plain["rows"]):
        assert a[0] == b[0] and a[2] == b[2]
        assert abs(a[1] - b[1]) <= 1e-9 * max(abs(b[1]), 1.0) and abs(a[3] - b[3]) <= 1e-9 * max(abs(b[3]), 1.0)
    # Tucker too (TTMs with the rank in place, the unfolding Gram with its ^ & indices)
    targs = ["-model", "Tucker", "-tensor", "r2", "-dim", "4", "-size", "9", "-rank", "3", "-pp", "1", "-maxiter", "12",
             "-pp_res_tol", "0.3", "-resprint", "3", "-filename", os.path.join(str(tmp_path), "t.csv")]
    tf = [(1, 100)] + [p for i in range(4) for p in ((2, i), (3, i))]
    plain = rh.run_cli("test_ALS", targs, fills=tf)
    os.environ["CTF_STANDIN_BLAS"] = os.path.abspath(hits[0])
    try:
        fast = rh.run_cli("test_ALS", targs, fills=tf)
    finally:
        del os.environ["CTF_STANDIN_BLAS"]
    assert fast["events"] == plain["events"] and len(fast["rows"]) == len(plain["rows"]) >= 3
    for a, b in zip(fast["rows"], plain["rows"]):
        assert abs(a[1] - b[1]) <= 1e-8 * max(abs(b[1]), 1.0) and abs(a[3] - b[3]) <= 1e-9 * max(abs(b[3]), 1.0)
