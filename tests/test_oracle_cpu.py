"""CPU tests (no GPU): the NumPy oracle against its independent C restatement, its own known-answer tests
 and the committed golden fixtures; the C-ABI library loads and exports every declared symbol."""
import ctypes as C
import importlib
import os
import re
import subprocess

import numpy as np
import pytest

from oracle import pp_oracle as o

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def naive():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_build", "libnaive.so"))
    lib.naive_u01.restype = C.c_double
    lib.naive_u01.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64]
    lib.naive_cp_residual.restype = C.c_double
    return lib


def f(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64).ravel(order="F"))


def p(a):
    return a.ctypes.data_as(C.c_void_p)


def lens_arr(lens):
    return (C.c_int64 * len(lens))(*lens)


def test_generator_matches_c(naive):
    ref = o.u01(5, 7, 64, start=3)
    got = np.array([naive.naive_u01(5, 7, 3 + i) for i in range(64)])
    assert np.array_equal(ref, got)
    assert ref.min() >= 0.0 and ref.max() < 1.0


@pytest.mark.parametrize("lens,x,R", [((5, 4, 6), 0, 3), ((5, 4, 6), 1, 3), ((5, 4, 6), 2, 3), ((4, 3, 5, 2), 2, 2)])
def test_first_contraction_and_mttv_vs_c(naive, lens, x, R):
    N = len(lens)
    V = o.fill_uniform(lens, 1, 0, -1, 1)
    W = o.fill_uniform((lens[x], R), 1, 1, -1, 1)
    idx = o.letters(N)
    rest = idx.replace(idx[x], "")
    ref = o.contract(rest + "*", V, idx, W, idx[x] + "*")
    out = np.zeros(ref.size)
    naive.naive_ttm_first(p(f(V)), lens_arr(lens), N, x, p(f(W)), R, p(out))
    assert np.allclose(out.reshape(ref.shape, order="F"), ref, rtol=0, atol=1e-13)
    T = o.fill_uniform(tuple(lens) + (R,), 1, 2, -1, 1)
    ref2 = o.contract(rest + "*", T, idx + "*", W, idx[x] + "*")
    out2 = np.zeros(ref2.size)
    naive.naive_mttv(p(f(T)), lens_arr(lens), N, x, p(f(W)), R, p(out2))
    assert np.allclose(out2.reshape(ref2.shape, order="F"), ref2, rtol=0, atol=1e-13)
    ref3 = o.ttm(V, x, W)
    out3 = np.zeros(ref3.size)
    naive.naive_ttm(p(f(V)), lens_arr(lens), N, x, p(f(W)), R, p(out3))
    assert np.allclose(out3.reshape(ref3.shape, order="F"), ref3, rtol=0, atol=1e-13)
    ref4 = o.unroll_tensor_contraction(V, x)
    out4 = np.zeros(ref4.size)
    naive.naive_unfold_gram(p(f(V)), lens_arr(lens), N, x, p(out4))
    assert np.allclose(out4.reshape(ref4.shape, order="F"), ref4, rtol=0, atol=1e-12)


def test_pp_correction_and_residual_vs_c(naive):
    lens, R = (5, 4, 6, 3), 2
    N = len(lens)
    V = o.fill_uniform(lens, 2, 0, -1, 1)
    W = [o.fill_uniform((lens[i], R), 2, 1 + i) for i in range(N)]
    dW = [0.01 * o.fill_uniform((lens[i], R), 2, 10 + i, -1, 1) for i in range(N)]
    ops = o.build_pp_operators(V, W)
    seq = o.letters(N)
    for i in range(N):
        ref = o.pp_corrected_M(ops, dW, i, N)
        M0 = f(ops["".join(c for k, c in enumerate(seq) if k != i)])
        arrs, which, dws, so = [], [], [], []
        for j in range(N):
            if j == i:
                continue
            arrs.append(f(ops["".join(c for k, c in enumerate(seq) if k not in (i, j))]))
            which.append(0 if j < i else 1)
            dws.append(f(dW[j]))
            so.append(lens[j])
        n = len(arrs)
        out = np.zeros(ref.size)
        naive.naive_pp_correct(p(M0), (C.c_void_p * n)(*[a.ctypes.data for a in arrs]), (C.c_int * n)(*which),
                               (C.c_void_p * n)(*[a.ctypes.data for a in dws]), lens_arr(so), n, C.c_int64(lens[i]), R,
                               p(out))
        assert np.allclose(out.reshape(ref.shape, order="F"), ref, rtol=0, atol=1e-12)
    Wf = [f(w) for w in W]
    r = naive.naive_cp_residual(p(f(V)), lens_arr(lens), N, (C.c_void_p * N)(*[a.ctypes.data for a in Wf]), R)
    assert abs(r - o.cp_residual(V, W)) < 1e-12 * np.linalg.norm(V)


# ---- known-answer tests of the algorithm -----------------------------------------------------------
def test_kat_exact_tensor_truth_start_is_a_fixed_point():
    lens, R = (7, 6, 5, 4), 3
    V, Wt = o.make_tensor_r(lens, R)
    W = [w.copy() for w in Wt]
    o.normalize(W)
    W0 = [w.copy() for w in W]
    G = o.init_grad(lens, R)
    _, tr = o.alsCP_DT(V, W, G, 0.0, 1, resprint=1)
    assert tr.rows[0][3] <= 1e-12 * np.linalg.norm(V)
    for a, b in zip(W, W0):
        assert np.abs(a - b).max() < 1e-10


def test_kat_pp_with_zero_dw_is_exact_mttkrp():
    lens, R = (6, 5, 7, 4), 3
    V, _ = o.make_tensor_r(lens, R)
    W = o.init_factors(lens, R)
    ops = o.build_pp_operators(V, W)
    zero = [np.zeros_like(w) for w in W]
    for i in range(4):
        seq = list(range(4))
        seq[i], seq[3] = seq[3], seq[i]
        assert np.allclose(o.pp_corrected_M(ops, zero, i, 4), o.KhatriRao_contract(V, W, seq), rtol=1e-13)


@pytest.mark.parametrize("N,s,R", [(3, 7, 3), (4, 6, 3), (6, 4, 2)])
def test_kat_optimizers_agree_and_tree_equals_krp(N, s, R):
    lens = (s,) * N
    V, _ = o.make_tensor_r(lens, R)
    res = {}
    for name, cls, steps in [("simple", o.CPSimpleOptimizer, 2), ("dt", o.CPDTOptimizer, 4)]:
        c = o.CPD(N, s, R, cls)
        c.Init(V, o.init_factors(lens, R))
        for _ in range(steps):
            c.optimizer.step()
        res[name] = c.W
    for a, b in zip(res["simple"], res["dt"]):
        assert np.abs(a - b).max() < 1e-10
    # MSDT: N steps update every mode N-1 times
    c = o.CPD(N, s, R, o.CPMSDTOptimizer)
    c.Init(V, o.init_factors(lens, R))
    total = sum(c.optimizer.step() for _ in range(N))
    assert abs(total - (N - 1)) < 1e-12
    parent, sibling = {}, {}
    o.construct_dimension_tree(parent, sibling, 0, N - 1)
    W = o.init_factors(lens, R)
    mm = {}
    for i in range(N):
        seq = list(range(N))
        seq[i], seq[N - 1] = seq[N - 1], seq[i]
        assert np.allclose(o._leaf_M(mm, parent, sibling, V, W, i), o.KhatriRao_contract(V, W, seq), rtol=1e-12)


def test_kat_dimension_tree_shapes():
    parent, sibling = {}, {}
    o.construct_dimension_tree(parent, sibling, 0, 3)
    assert parent == {"ab": "abcd", "cd": "abcd", "a": "ab", "b": "ab", "c": "cd", "d": "cd"}
    assert sibling["ab"] == "cd" and sibling["a"] == "b"
    parent, sibling = {}, {}
    o.construct_dimension_tree(parent, sibling, 0, 5)
    assert parent["abc"] == "abcdef" and parent["ab"] == "abc" and parent["c"] == "abc" and parent["f"] == "def"
    assert sibling["c"] == "ab" and sibling["de"] == "f"


def test_kat_residual_monotone_for_exact_als():
    lens, R = (8, 7, 6, 5), 3
    V, _ = o.make_tensor_r(lens, R)
    W, G = o.init_factors(lens, R), o.init_grad(lens, R)
    _, tr = o.alsCP_DT(V, W, G, 0.0, 15, resprint=1)
    res = [r[3] for r in tr.rows][1:]
    assert all(b <= a * (1 + 1e-12) for a, b in zip(res, res[1:]))


def test_kat_tucker_orthonormal_and_core_norm_nondecreasing():
    lens, R = (8, 7, 6), 3
    V = o.make_tensor_r2(lens)
    core, W = o.hosvd(V, [R] * 3)
    norms = [np.linalg.norm(core)]
    for _ in range(4):
        _, rows, core = o.alsTucker_DT(V, core, W, 0.0, 0, resprint=1)
        norms.append(np.linalg.norm(core))
        for w in W:
            assert np.abs(w.T @ w - np.eye(R)).max() < 1e-12
    assert all(b >= a - 1e-12 for a, b in zip(norms, norms[1:]))
    # PP with dW = 0 equals the exact TTMc
    ops = o.build_tucker_pp_operators(V, W)
    for i in range(3):
        assert np.allclose(o.tucker_pp_corrected_Y(ops, [np.zeros_like(w) for w in W], i, 3), o.TTMc(V, W, i), rtol=1e-12)


# ---- golden fixtures (generated by tests/golden/make_golden.py from this oracle; they pin it against drift) --------
@pytest.mark.parametrize("name", sorted(x[:-4] for x in os.listdir(GOLD) if x.endswith(".npz")) if os.path.isdir(GOLD) else [])
def test_oracle_reproduces_golden(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    lens, R = tuple(int(v) for v in g["lens"]), int(g["R"])
    V, _ = o.make_tensor_r(lens, R)
    W, G = o.init_factors(lens, R), o.init_grad(lens, R)
    vnorm = np.linalg.norm(V)
    assert abs(vnorm - float(g["vnorm"])) < 1e-12 * vnorm
    _, tr = o.alsCP_PP(V, W, G, 1e-10 * vnorm, float(g["tol_init"]), int(g["maxiter"]), resprint=int(g["resprint"]))
    assert [(0 if k == "DT" else 1, it) for k, it in tr.events] == [tuple(r) for r in g["events"].tolist()]
    rows = np.array([(r[0], r[1], r[2], r[3]) for r in tr.rows])
    assert rows.shape == g["rows"].shape
    assert np.allclose(rows[:, 1], g["rows"][:, 1], rtol=1e-9, atol=1e-9 * vnorm)
    assert np.allclose(rows[:, 3], g["rows"][:, 3], rtol=0, atol=1e-10 * vnorm)
    for i in range(len(lens)):
        assert np.abs(W[i] - g["W%d" % i]).max() < 1e-8


# ---- the C ABI ------------------------------------------------------------------------------------------------------
def test_library_exports_every_declared_symbol():
    ppx = importlib.import_module("pairwise-perturbation_b200")
    lib = ppx.load_library()
    header = open(os.path.join(ROOT, "include", "ppx.h")).read()
    declared = set(re.findall(r"\b(ppx_[a-z0-9_]+)\s*\(", header))
    declared.discard("ppx_ctx")
    assert len(declared) >= 50
    for name in declared:
        assert hasattr(lib, name), name
    assert declared <= set(ppx.SIGNATURES), declared - set(ppx.SIGNATURES)
    assert b"sm_100a" in lib.ppx_version()


def test_host_library_loads_and_fails_loudly_without_gpu():
    import torch

    ppx = importlib.import_module("pairwise-perturbation_b200")
    H = importlib.import_module("pairwise-perturbation_b200.host_api")
    H.load_host_library()
    assert ppx.shard_range(300, 8, 0) == (0, 38) and ppx.shard_range(300, 8, 7) == (263, 300)
    assert sum(ppx.shard_range(300, 8, r)[1] - ppx.shard_range(300, 8, r)[0] for r in range(8)) == 300
    if not torch.cuda.is_available():
        with pytest.raises(ppx.PpxError):  # no CPU fallback anywhere
            ppx.Ctx(0)
        with pytest.raises(ppx.PpxError):
            H.World(0)


# ---- input generators 'p' / 'p2' / 'c' (test_ALS.cxx:222-261) ----------------------------------------------------
@pytest.mark.parametrize("N,s", [(4, 3), (6, 3), (4, 5)])
def test_laplacian_tensor_closed_form(N, s):
    """The identity-tensor construction of common.cxx:575-642 equals the closed form the CUDA generator fills:
    no index pair off the diagonal -> 2d, exactly one pair (a,b) off -> D[a,b], more -> 0."""
    V = o.laplacian_tensor(N, s)
    d = N // 2
    W = np.zeros((s,) * N)
    for idx in np.ndindex(*W.shape):
        off = [(idx[2 * m], idx[2 * m + 1]) for m in range(d) if idx[2 * m] != idx[2 * m + 1]]
        if not off:
            W[idx] = 2.0 * d
        elif len(off) == 1:
            W[idx] = -1.0 if abs(off[0][0] - off[0][1]) == 1 else 0.0
    assert np.array_equal(V, W)
    # as an operator on d-dimensional grids it is the 2d-point Laplacian: fold (a_m) -> row, (b_m) -> column
    A = np.transpose(V, list(range(0, N, 2)) + list(range(1, N, 2))).reshape(s ** d, s ** d)
    D = 2.0 * np.eye(s) - np.eye(s, k=1) - np.eye(s, k=-1)
    L = sum(np.kron(np.kron(np.eye(s ** m), D), np.eye(s ** (d - 1 - m))) for m in range(d))
    assert np.array_equal(np.sort(np.linalg.eigvalsh(A)), np.sort(np.linalg.eigvalsh(A.T)))
    assert np.allclose(np.sort(np.linalg.eigvalsh(A)), np.sort(np.linalg.eigvalsh(L)))
    # 'p' is the same buffer seen as d modes of size s*s
    P = o.make_tensor_p(N, s, folded=True)
    assert P.shape == (s * s,) * d and np.array_equal(P.ravel(order="F"), V.ravel(order="F"))


def test_gen_collinearity_respects_the_range_and_noise_level():
    lens, R = (9, 8, 7), 4
    X, vec = o.gen_collinearity(lens, R, 0.6, 0.85, seed=1)
    for j in range(len(lens)):
        for i in range(1, R):
            for k in range(i):
                assert 0.6 <= o.collinearity(vec[i][j], vec[k][j]) <= 0.85
    lam = [0.2 + 0.6 / R * (i + 1) for i in range(R)]
    ref = sum(lam[i] * np.einsum("a,b,c->abc", vec[i][0], vec[i][1], vec[i][2]) for i in range(R))
    assert np.allclose(X, ref, rtol=1e-13, atol=1e-14)
    V = o.make_tensor_c(lens, R, 0.6, 0.85, ratio_noise=0.05, seed=1)
    assert abs(np.linalg.norm(V - X) / np.linalg.norm(X) - 0.05) < 1e-12


# ---- low-rank-update optimizers (run.cxx -pp 2 / 3) -----------------------------------------------------------------
def _run_opt(cls, lens, R, steps, *opt_args):
    V, _ = o.make_tensor_r(lens, R)
    W, G = o.init_factors(lens, R), o.init_grad(lens, R)
    opt = cls(len(lens), R, *opt_args)
    opt.configure(V, W, G, 0.0)
    for _ in range(steps):
        opt.step()
    return V, opt


@pytest.mark.parametrize("lens,R", [((7, 6, 5, 6), 3), ((6, 5, 7), 2)])
def test_lr_optimizers_with_full_update_rank_are_the_exact_ones(lens, R):
    """A rank-R 'low-rank' update is the whole update M S^-1 - W, and the patched root tensors are then the exact
    first contractions: CPDTLROptimizer == CPDTOptimizer and CPMSDTLROptimizer == CPMSDTOptimizer step for step."""
    for exact, lr in ((o.CPDTOptimizer, o.CPDTLROptimizer), (o.CPMSDTOptimizer, o.CPMSDTLROptimizer)):
        _, a = _run_opt(exact, lens, R, 14)
        _, b = _run_opt(lr, lens, R, 14, R, 0)
        for wa, wb in zip(a.W, b.W):
            assert np.abs(wa - wb).max() < 1e-11 * max(1.0, np.abs(wa).max())


def test_rank_update_is_the_best_rank_r_correction_in_the_S_norm():
    """get_rankR_update_cholesky: U s VT is the rank-r truncation of (M - A S) L^-T mapped back by L^-1, so
    A + U s VT -> M S^-1 as r -> R, and the patched root equals V x (A + U s VT)."""
    rng = np.random.default_rng(3)
    s_, R = 11, 5
    A, M = rng.random((s_, R)), rng.random((s_, R))
    B = rng.random((20, R))
    S = B.T @ B
    exact = M @ np.linalg.inv(S)
    errs = []
    for r in range(1, R + 1):
        Us, VT = o.get_rankR_update_cholesky(r, M, A, S)
        assert Us.shape == (s_, r) and VT.shape == (r, R)
        L = np.linalg.cholesky(S)
        errs.append(np.linalg.norm((A + Us @ VT - exact) @ L))   # error in the S-weighted norm
    assert all(errs[i + 1] <= errs[i] + 1e-12 for i in range(R - 1)) and errs[-1] < 1e-10
    # randomized range finder: exact when r = R (the range is everything)
    Us, VT = o.get_rankR_update_cholesky(R, M, A, S, random=True)
    assert np.abs(A + Us @ VT - exact).max() < 1e-9


# ---- the BLAS formulation the oracle switches to for large arrays (only bench.py's CPU baseline reaches that size) ----
@pytest.mark.parametrize("shape,iT,x", [((4, 30, 20, 25), "abcd", "c"), ((4, 30, 20, 25), "abcd", "a"),
                                        ((4, 30, 20, 25), "abcd", "d"), ((6, 17, 20, 9), "abd*", "d"),
                                        ((6, 17, 20, 9), "abd*", "a"), ((6, 17, 20, 9), "abd*", "b"), ((40, 7), "a*", "a")])
def test_contract_blas_path_equals_einsum(shape, iT, x, monkeypatch):
    T = o.fill_uniform(shape, 21, 0, -1.0, 1.0)
    batched = iT.endswith("*")
    modes = iT[:-1] if batched else iT
    R = shape[-1] if batched else 6
    W = o.fill_uniform((shape[modes.index(x)], R), 21, 1, -1.0, 1.0)
    out = modes.replace(x, "") + "*"
    monkeypatch.setattr(o, "_FAST_MIN_ELEMS", 1)
    fast = o.contract(out, T, iT, W, x + "*")
    assert o._contract_big(out, T, iT, W, x + "*") is not None  # the pattern is recognised
    monkeypatch.setattr(o, "_FAST_MIN_ELEMS", 10 ** 15)
    ref = o.contract(out, T, iT, W, x + "*")
    assert fast.shape == ref.shape and np.abs(fast - ref).max() <= 1e-13 * np.abs(ref).max()


# ---- the closed-form known answers used for the full-size GPU parity checks ---------------------------------------
@pytest.mark.parametrize("lens,R", [((7, 6, 5, 4), 3), ((5, 4, 6), 2), ((4, 3, 4, 3, 4, 3), 2)])
def test_closed_form_known_answers_match_the_oracle(lens, R):
    """pairwise-perturbation_b200/kat.py (MTTKRP and PP pair operators of an exact CP tensor from the small matrices
    A_m^T W_m) against the oracle's contractions of the materialised tensor, and on a mode-0 shard."""
    kat = importlib.import_module("pairwise-perturbation_b200.kat")
    N = len(lens)
    A = [o.fill_uniform((l, R), 1, i) for i, l in enumerate(lens)]
    W = [o.fill_uniform((l, R), 2, i) for i, l in enumerate(lens)]
    V = o.build_V(A)
    C = kat.cross_grams(A, W)
    seq = o.letters(N)
    ops = o.build_pp_operators(V, W)
    for i in range(N):
        ref = o.KhatriRao_contract(V, W, [j for j in range(N) if j != i] + [i])
        assert kat.max_rel_err(kat.mttkrp(A, C, i), ref) < 1e-13
        assert kat.max_rel_err(kat.mttkrp(A, C, i), ops[seq.replace(seq[i], "")]) < 1e-13
        for j in range(i + 1, N):
            key = seq.replace(seq[i], "").replace(seq[j], "")
            assert kat.max_rel_err(kat.pair_operator(A, C, i, j), ops[key]) < 1e-13
    b, e = 1, lens[0] - 1
    assert np.array_equal(kat.mttkrp(A, C, 0, rows=(b, e)), kat.mttkrp(A, C, 0)[b:e])
    assert np.allclose(kat.pair_operator(A, C, 0, 1, rows_i=(b, e)), kat.pair_operator(A, C, 0, 1)[b:e], rtol=1e-15)
