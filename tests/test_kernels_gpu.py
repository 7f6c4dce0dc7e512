"""Parity of every CUDA kernel behind include/ppx.h against the CPU oracle (oracle/pp_oracle.py), through the C ABI.

Tolerances: all arithmetic is FP64; a contraction of length K differs from the oracle only by summation order, so
results must agree to ~K*eps relative to the magnitude of the result (checked as 1e-12 of the max-abs).  The solve
is compared at 1e-9 (conditioning of S), well inside the north-star's 1e-8 factor tolerance."""
import numpy as np
import pytest

from oracle import pp_oracle as o

pytestmark = pytest.mark.gpu


def rel_err(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def rnd(shape, seed):
    return o.fill_uniform(shape, 7, seed, -1.0, 1.0)


def test_fill_uniform_matches_oracle(ctx):
    out = ctx.empty(1000)
    ctx.fill_uniform(out, seed=2, tensor_id=3, start=17, lo=0.5, hi=1.0)
    ref = 0.5 + 0.5 * o.u01(2, 3, 1000, start=17)
    assert np.array_equal(ctx.to_host(out, (1000,)), ref)  # bit exact: integer hash + one fma-free affine map


TTM_CASES = [
    # lens, x, R
    ((13, 9, 11), 0, 5), ((13, 9, 11), 1, 5), ((13, 9, 11), 2, 5),
    ((12, 10, 8, 6), 2, 4), ((12, 10, 8, 6), 0, 4), ((12, 10, 8, 6), 1, 10), ((12, 10, 8, 6), 3, 3),
    ((7, 5, 6, 4, 5, 3), 3, 3), ((7, 5, 6, 4, 5, 3), 0, 3),
    ((40, 40, 40), 1, 10), ((300, 37), 0, 50), ((37, 300), 1, 50), ((64, 33, 20), 1, 70),
    ((200, 3, 50), 1, 10), ((3, 128, 16), 0, 10), ((129, 17), 1, 1), ((1, 9, 5), 1, 2), ((130,), 0, 7),
    # TMA-eligible shapes (even leading extent, rows fill the 128-row tiles): M-major 3-D maps and k-major 2-D maps
    ((256, 24, 6), 1, 5), ((512, 17, 3), 1, 50), ((1024, 20), 1, 10), ((128, 8, 4, 6), 1, 7), ((384, 33, 2), 1, 64),
    ((300, 150), 0, 50), ((18, 700), 0, 3), ((64, 129), 0, 56), ((2560, 9), 1, 9),
    # leftover rank columns (R mod 8 in 1..4) run as DFMA beside the DMMA tiles: every tail width, both layouts
    ((40, 260), 0, 9), ((40, 260), 0, 10), ((34, 300), 0, 11), ((34, 300), 0, 12), ((34, 300), 0, 13),
    ((50, 131), 0, 17), ((20, 400), 0, 27), ((20, 400), 0, 36), ((22, 129), 0, 60),
    ((256, 19, 3), 1, 25), ((256, 19, 3), 1, 26), ((384, 10, 2), 1, 27), ((128, 33, 4), 1, 28), ((256, 9, 5), 1, 33),
    ((128, 21, 3), 1, 41), ((640, 7, 2), 1, 58), ((256, 12, 3), 1, 59),
    # M-major with L % 16 != 0: eight 16 x 16 boxes per stage instead of the single 4-D box
    ((250, 12, 3), 1, 50), ((378, 9, 2), 1, 33), ((122, 40, 2), 1, 26),
    # L % 16 != 0 with several row tiles (a mode-0 shard of 37 rows gives L = 37 * s): one 4-D box per stage for the tiles
    # made of whole groups of 16 l, eight 3-D boxes for the last tile of each t
    ((1110, 24), 1, 50), ((1110, 12, 3), 1, 50), ((1140, 9, 2), 1, 26), ((270, 6, 2), 1, 64), ((130, 16, 2), 1, 33),
    # M-major with a short L: 16 l x 8 t row tiles (one 3-D box per stage)
    ((300, 20, 16), 1, 50), ((46, 17, 40), 1, 27), ((300, 7, 4, 6), 1, 35), ((94, 33, 8), 1, 64), ((16, 40, 24), 1, 25),
    # streaming kernels (X <= 64 and R <= 16, k1_ttm_stream.cu): lanes along l with one or two rows per thread
    # (L odd / even), the slab-staged kernel for L == 1, every register-block width, ragged last block
    ((40, 40, 40), 1, 4), ((40, 40, 40), 2, 16), ((40, 24, 30), 1, 1), ((41, 17, 29), 1, 10), ((25, 64, 33), 1, 12),
    ((1600, 40, 3), 1, 10), ((16, 3, 50), 1, 5), ((18, 1, 40), 1, 3),
    ((40, 700), 0, 10), ((3, 5000), 0, 10), ((64, 150), 0, 16), ((7, 333, 3), 0, 8), ((1, 600), 0, 4), ((33, 1027), 0, 13),
    ((3, 128, 300), 1, 10), ((2, 200, 150), 1, 16), ((5, 256, 120), 1, 7), ((15, 70, 40), 1, 3),
]


@pytest.mark.parametrize("lens,x,R", TTM_CASES)
def test_ttm_first(ctx, lens, x, R):
    N = len(lens)
    V = rnd(lens, 1)
    W = rnd((lens[x], R), 2)
    idx = o.letters(N)
    ref = o.contract(idx.replace(idx[x], "") + "*", V, idx, W, idx[x] + "*")
    out = ctx.empty(ref.size)
    ctx.ttm_first(ctx.to_device(V), lens, x, ctx.to_device(W), R, out)
    got = ctx.to_host(out, ref.shape)
    assert rel_err(got, ref) < 1e-12


def test_ttm_first_unaligned_and_ldw(ctx):
    # odd base address (8-byte aligned only) and a padded leading dimension for W
    lens, x, R = (14, 10, 6), 1, 6
    V = rnd(lens, 3)
    W = rnd((lens[x], R), 4)
    ref = o.contract("ac*", V, "abc", W, "b*")
    Vd = ctx.empty(V.size + 1)
    Vd[1:].copy_(ctx.to_device(V))
    ldw = lens[x] + 3
    Wp = np.zeros((ldw, R))
    Wp[: lens[x]] = W
    out = ctx.empty(ref.size)
    ctx.ttm_first(Vd[1:], lens, x, ctx.to_device(Wp), R, out, ldw=ldw)
    assert rel_err(ctx.to_host(out, ref.shape), ref) < 1e-12


MTTV_CASES = [((13, 9, 11), 0, 5), ((13, 9, 11), 1, 5), ((13, 9, 11), 2, 5), ((300, 300), 1, 50), ((300, 300), 0, 50),
              ((40, 33), 0, 10), ((3, 50, 7), 1, 4), ((70, 5, 9, 4), 2, 3), ((6,), 0, 3), ((100, 64, 3), 1, 2),
              # one thread per output for X <= 64 (lanes along l, L even / odd; slab-staged for L == 1)
              ((40, 40, 12), 1, 10), ((41, 17, 9), 1, 5), ((1600, 40), 1, 10), ((16, 64, 8), 1, 3), ((25, 3, 30), 1, 7),
              ((40, 130), 0, 10), ((3, 999), 0, 4), ((64, 77), 0, 2), ((17, 40, 5), 0, 6),
              # few outputs, long contracted mode: x split over CTAs (coil-shaped leaves)
              ((3, 7200), 1, 10), ((128, 7200), 1, 10), ((5, 2, 1500), 2, 3), ((1000,), 0, 10), ((40, 600), 1, 2),
              ((700, 6), 0, 5)]


@pytest.mark.parametrize("lens,x,R", MTTV_CASES)
def test_mttv(ctx, lens, x, R):
    k = len(lens)
    T = rnd(tuple(lens) + (R,), 5)
    W = rnd((lens[x], R), 6)
    idx = o.letters(k)
    ref = o.contract(idx.replace(idx[x], "") + "*", T, idx + "*", W, idx[x] + "*")
    out = ctx.empty(ref.size)
    ctx.mttv(ctx.to_device(T), lens, x, ctx.to_device(W), R, out)
    assert rel_err(ctx.to_host(out, ref.shape), ref) < 1e-12


@pytest.mark.parametrize("lens,R,which", [
    ((300, 31, 29), 17, (1, 1, 1)),  # one-pass kernel: two x-tiles (23 + 8), rows not a multiple of the warp
    ((300, 31, 29), 17, (1, 1, 0)),  # the level-1 tensor with two consumers (PP build, mode c)
    ((192, 40, 37), 15, (0, 1, 1)), ((192, 40, 37), 15, (1, 0, 1)),
    ((320, 20, 66), 10, (1, 1, 1)),  # the longest row the kernel takes; one x-tile: no partial sums
    ((64, 70, 60), 16, (1, 1, 1)), ((38, 101, 90), 13, (1, 1, 1)),  # short rows: the separate kernels
    ((37, 80, 75), 20, (1, 1, 1)),   # odd row extent: the separate kernels
    ((9, 8, 7), 4, (1, 1, 1)), ((9, 8, 7), 4, (0, 1, 0))])
def test_mttv3_all_hadamard_contractions_in_one_pass(ctx, lens, R, which):
    """ppx_mttv3 = the Hadamard contractions of one level-1 tensor along each of its three modes (als_CP.cxx:394-408),
    every requested output against the einsum of the oracle."""
    T = rnd(tuple(lens) + (R,), 21)
    Ws = [rnd((lens[i], R), 22 + i) for i in range(3)]
    refs = [o.contract("abc".replace("abc"[i], "") + "*", T, "abc*", Ws[i], "abc"[i] + "*") for i in range(3)]
    outs = [ctx.empty(refs[i].size) if which[i] else None for i in range(3)]
    dW = [ctx.to_device(w) if which[i] else None for i, w in enumerate(Ws)]
    ctx.mttv3(ctx.to_device(T), lens, dW[0], dW[1], dW[2], R, outs[0], outs[1], outs[2])
    for i in range(3):
        if which[i]:
            assert rel_err(ctx.to_host(outs[i], refs[i].shape), refs[i]) < 1e-12, "output %d" % i


@pytest.mark.parametrize("lens,x1,x2,R", [((9, 8, 7), 0, 1, 4), ((9, 8, 7), 1, 2, 4), ((9, 8, 7), 0, 2, 4),
                                          ((40, 40, 40), 0, 1, 10)])
def test_mttv2(ctx, lens, x1, x2, R):
    T = rnd(tuple(lens) + (R,), 8)
    W1, W2 = rnd((lens[x1], R), 9), rnd((lens[x2], R), 10)
    idx = "abc"
    keep = [c for c in idx if c not in (idx[x1], idx[x2])][0]
    ref = o.contract(keep + "*", T, "abc*", W1, idx[x1] + "*", W2, idx[x2] + "*")
    out = ctx.empty(ref.size)
    ctx.mttv2(ctx.to_device(T), lens, x1, ctx.to_device(W1), x2, ctx.to_device(W2), R, out)
    assert rel_err(ctx.to_host(out, ref.shape), ref) < 1e-12


def test_ttm_first_mttv(ctx):
    lens, R = (12, 10, 8, 6), 5
    V = rnd(lens, 11)
    Wc, Wd = rnd((8, R), 12), rnd((6, R), 13)
    ref = o.contract("ab*", V, "abcd", Wc, "c*", Wd, "d*")
    out = ctx.empty(ref.size)
    ctx.ttm_first_mttv(ctx.to_device(V), lens, 2, ctx.to_device(Wc), 3, ctx.to_device(Wd), R, out)
    assert rel_err(ctx.to_host(out, ref.shape), ref) < 1e-12


@pytest.mark.parametrize("lens,R", [((12, 10, 8, 6), 4), ((33, 40, 35), 7), ((5, 6, 4, 5, 3, 4), 3),
                                    # few rows against long operators: the contracted index is split over gridDim.z
                                    ((3, 16, 600), 10), ((4, 301, 20, 6), 3), ((40, 3, 517), 5)])
def test_pp_correct_matches_reference_formula(ctx, lens, R):
    N = len(lens)
    V = rnd(lens, 14)
    W = [rnd((lens[i], R), 20 + i) for i in range(N)]
    dW = [1e-2 * rnd((lens[i], R), 40 + i) for i in range(N)]
    ops = o.build_pp_operators(V, W)
    seq = o.letters(N)
    dev_ops = {k: ctx.to_device(v) for k, v in ops.items()}
    dev_dW = [ctx.to_device(d) for d in dW]
    for i in range(N):
        ref = o.pp_corrected_M(ops, dW, i, N)
        m0 = dev_ops["".join(c for k, c in enumerate(seq) if k != i)]
        oplist, which, dws, sother = [], [], [], []
        for j in range(N):
            if j == i:
                continue
            oplist.append(dev_ops["".join(c for k, c in enumerate(seq) if k not in (i, j))])
            which.append(0 if j < i else 1)
            dws.append(dev_dW[j])
            sother.append(lens[j])
        out = ctx.empty(ref.size)
        ctx.pp_correct(m0, oplist, which, dws, sother, lens[i], R, out)
        assert rel_err(ctx.to_host(out, ref.shape), ref) < 1e-12


def test_pp_correct_zero_dw_is_identity(ctx):
    # PP with dW = 0 reproduces the exact MTTKRP (als_CP.cxx:778)
    s, R = 37, 5
    M0 = rnd((s, R), 50)
    op = rnd((s, s, R), 51)
    z = ctx.zeros(s * R)
    out = ctx.empty(s * R)
    ctx.pp_correct(ctx.to_device(M0), [ctx.to_device(op)] * 2, [0, 1], [z, z], [s, s], s, R, out)
    assert np.array_equal(ctx.to_host(out, (s, R)), M0)


@pytest.mark.parametrize("s,R", [(300, 50), (13, 5), (7200, 10), (40, 10), (5, 1)])
def test_gram_and_hadamard(ctx, s, R):
    Ws = [rnd((s + i, R), 60 + i) for i in range(4)]
    Gs = []
    for i, w in enumerate(Ws):
        G = ctx.empty(R * R)
        ctx.gram(ctx.to_device(w), s + i, R, G)
        assert rel_err(ctx.to_host(G, (R, R)), w.T @ w) < 1e-12
        Gs.append(G)
    S = ctx.empty(R * R)
    ctx.hadamard_grams(Gs, 1, R, 0.25, S)
    ref = (Ws[0].T @ Ws[0]) * (Ws[2].T @ Ws[2]) * (Ws[3].T @ Ws[3]) + 0.25 * np.eye(R)
    assert rel_err(ctx.to_host(S, (R, R)), ref) < 1e-12


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("s,R", [(300, 50), (13, 5), (40, 10), (33, 7), (9, 1)])
def test_solve_update(ctx, ppx, mode, s, R):
    N = 4
    Wf = [o.fill_uniform((s, R), 3, 70 + i) for i in range(N)]
    S = o.gram_hadamard(Wf, 0)
    M = rnd((s, R), 80)
    W_old = rnd((s, R), 81)
    W_init = rnd((s, R), 82)
    grad_ref = -M + W_old @ S
    for ratio in (1.0, 0.7):
        W_ref, dW_ref = o.SVD_solve_mod(M, W_init, S, ratio)
        Wd = ctx.to_device(W_old)
        grad, dW, sq = ctx.empty(s * R), ctx.empty(s * R), ctx.empty(3)
        ctx.solve_update(ctx.to_device(M), ctx.to_device(S), Wd, s, R, W_init=ctx.to_device(W_init), ratio_step=ratio,
                         mode=mode, grad=grad, dW=dW, sq_norms=sq)
        assert rel_err(ctx.to_host(Wd, (s, R)), W_ref) < 1e-9
        assert rel_err(ctx.to_host(dW, (s, R)), dW_ref) < 1e-9
        assert rel_err(ctx.to_host(grad, (s, R)), grad_ref) < 1e-12
        sqh = ctx.to_host(sq, (3,))
        ref_sq = np.array([np.sum(W_ref**2), np.sum(dW_ref**2), np.sum(grad_ref**2)])
        assert np.allclose(sqh, ref_sq, rtol=1e-9)
    # plain solve (no W_init): W = M S^-1, Cholesky and SVD semantics agree with the oracle's two solvers
    Wd = ctx.to_device(W_old)
    ctx.solve_update(ctx.to_device(M), ctx.to_device(S), Wd, s, R, mode=mode)
    assert rel_err(ctx.to_host(Wd, (s, R)), o.SVD_solve(M, S)) < 1e-9
    assert rel_err(ctx.to_host(Wd, (s, R)), o.cholesky_solve(M, S)) < 1e-9


@pytest.mark.parametrize("R", [1, 2, 3, 7, 8, 9, 10, 16, 17, 23, 24, 25, 32, 33, 40, 41, 48, 49, 50, 56, 57, 64, 65, 77, 100, 112])
def test_spd_inverse(ctx, R):
    """S^-1 of the Gram-Hadamard matrix (cholesky_solve semantics, common.cxx:727-737) by the register-resident sweep
    kernel (R <= 64, every 8-column class and its edges) and the one-CTA LDL^T kernel (R > 64, every register-tile class), against NumPy, with and without lambda, and on a matrix with a wide
    spectrum (Hadamard product of three Grams of nearly collinear factors)."""
    s = 3 * R + 5
    for trial, lam in enumerate((0.0, 1e-3, 0.0)):
        Ws = [rnd((s, R), 400 + 7 * trial + j) for j in range(4)]
        if trial == 2:  # nearly collinear columns: condition number ~1e7..1e9
            Ws = [0.05 * w + np.ones((s, 1)) * rnd((1, R), 440 + j) for j, w in enumerate(Ws)]
        Gs = []
        for w in Ws:
            G = ctx.empty(R * R)
            ctx.gram(ctx.to_device(w), s, R, G)
            Gs.append(G)
        S = np.ones((R, R))
        for j in (0, 2, 3):
            S = S * (Ws[j].T @ Ws[j])
        S = S + lam * np.eye(R)
        S_out, Sinv = ctx.empty(R * R), ctx.empty(R * R)
        ctx.spd_inverse_g(Gs, 1, lam, R, 0, S_out, Sinv)
        got = ctx.to_host(Sinv, (R, R))
        assert rel_err(ctx.to_host(S_out, (R, R)), S) < 1e-12
        cond = np.linalg.cond(S)
        assert np.abs(got @ S - np.eye(R)).max() < 1e-14 * max(cond, 1e3) * R
        assert np.abs(got - got.T).max() <= 1e-15 * np.abs(got).max()
        ref = np.linalg.inv(S)
        assert rel_err(got, ref) < 1e-13 * max(cond, 1e3)


def test_normalize(ctx):
    sizes, R = [13, 40, 7, 300], 6
    W = [rnd((s, R), 90 + i) * (i + 1) for i, s in enumerate(sizes)]
    ref = [w.copy() for w in W]
    o.normalize(ref)
    dW = [ctx.to_device(w) for w in W]
    Gs = []
    for w, s in zip(dW, sizes):
        G = ctx.empty(R * R)
        ctx.gram(w, s, R, G)
        Gs.append(G)
    ctx.normalize(dW, sizes, R, Gs)
    for i, s in enumerate(sizes):
        assert rel_err(ctx.to_host(dW[i], (s, R)), ref[i]) < 1e-13
        assert rel_err(ctx.to_host(Gs[i], (R, R)), ref[i].T @ ref[i]) < 1e-12


@pytest.mark.parametrize("sizes", [[4096, 8, 8], [13, 40, 7, 300], [3, 128, 128, 7200]])
def test_normalize_norms_ragged(ctx, sizes):
    """Normalize + switching norms in one call (als_CP.cxx:657-664, :825; common.cxx:680-688) with very unequal factor
    sizes, repeated: the block of a small factor finishes (and rescales its Gram) long before the block of a large one
    starts -- every block must still see the traces of the UNSCALED Grams (round-1 advisor finding)."""
    R = 6
    W = [rnd((s, R), 190 + i) * (i + 1) for i, s in enumerate(sizes)]
    D = [rnd((s, R), 290 + i) for i, s in enumerate(sizes)]
    ref = [w.copy() for w in W]
    o.normalize(ref)
    for rep in range(20):
        dW = [ctx.to_device(w) for w in W]
        dD = [ctx.to_device(d) for d in D]
        Gs = []
        for w, s in zip(dW, sizes):
            G = ctx.empty(R * R)
            ctx.gram(w, s, R, G)
            Gs.append(G)
        sq = ctx.empty(2 * len(sizes))
        ctx.normalize_norms(dW, dD, sizes, R, Gs, sq)
        h = ctx.to_host(sq, (2 * len(sizes),))
        for i, s in enumerate(sizes):
            assert rel_err(ctx.to_host(dW[i], (s, R)), ref[i]) < 1e-13
            assert rel_err(ctx.to_host(Gs[i], (R, R)), ref[i].T @ ref[i]) < 1e-12
            assert abs(h[2 * i] - np.sum(D[i] ** 2)) <= 1e-12 * np.sum(D[i] ** 2)
            assert abs(h[2 * i + 1] - np.sum(ref[i] ** 2)) <= 1e-12 * np.sum(ref[i] ** 2)


def test_sqnorms_and_diff_update(ctx):
    a, b = rnd((1234,), 95), rnd((1234,), 96)
    out = ctx.empty(2)
    ctx.sqnorms([ctx.to_device(a), ctx.to_device(b)], out)
    assert np.allclose(ctx.to_host(out, (2,)), [np.sum(a * a), np.sum(b * b)], rtol=1e-13)
    big, small = rnd((5_000_001,), 97), rnd((77,), 98)  # tensor-sized arrays take the grid-wide two-stage sum
    ctx.sqnorms([ctx.to_device(big), ctx.to_device(small)], out)
    assert np.allclose(ctx.to_host(out, (2,)), [np.sum(big * big), np.sum(small * small)], rtol=1e-13)
    Wp, dW, sq = ctx.to_device(b), ctx.empty(1234), ctx.empty(2)
    ctx.diff_update(ctx.to_device(a), Wp, dW, sq)
    assert np.array_equal(ctx.to_host(dW, (1234,)), a - b)
    assert np.array_equal(ctx.to_host(Wp, (1234,)), a)
    assert np.allclose(ctx.to_host(sq, (2,)), [np.sum((a - b) ** 2), np.sum(a * a)], rtol=1e-13)


@pytest.mark.parametrize("lens,R", [((12, 10, 8, 6), 4), ((13, 9, 11), 5), ((5, 6, 4, 5, 3, 4), 3), ((130, 35), 50),
                                    ((7, 200), 3),
                                    # >= 65536 elements: the tensor-pipe kernel (128-row tiles, 64-column chunks of the last
                                    # mode, ragged in both; R padded to a multiple of 4; every leading dimension class)
                                    ((40, 30, 20, 35), 50), ((64, 33, 50), 10), ((130, 35, 300), 3), ((7, 9, 11, 13, 9), 27),
                                    ((300, 250), 64), ((50, 40, 60), 100), ((129, 8, 8, 70), 1), ((33, 2000), 12),
                                    ((50, 40, 60), 101),
                                    # the strip kernel's register classes (3 / 7 / 13 k-steps), strips that cross a
                                    # boundary of the leading mode, two super-chunks of the last factor
                                    ((17, 19, 23, 40), 28), ((300, 30, 304), 52), ((20, 20, 20, 20), 13),
                                    ((18, 5000), 9), ((16, 16, 16, 16, 16), 5)])
def test_cp_residual_and_reconstruct(ctx, lens, R):
    N = len(lens)
    W = [rnd((lens[i], R), 100 + i) for i in range(N)]
    V = rnd(lens, 110)
    dW = [ctx.to_device(w) for w in W]
    Vhat = ctx.empty(V.size)
    ctx.cp_reconstruct(lens, dW, R, Vhat)
    ref = o.build_V(W)
    assert rel_err(ctx.to_host(Vhat, lens), ref) < 1e-13
    sq = ctx.empty(1)
    ctx.cp_residual(ctx.to_device(V), lens, dW, R, sq)
    assert abs(np.sqrt(ctx.to_host(sq, (1,))[0]) - o.cp_residual(V, W)) < 1e-11 * np.linalg.norm(V)
    # known answer: exact rank-R tensor, W = truth -> residual ~ 0 (SURVEY 8c KAT 1)
    ctx.cp_residual(Vhat, lens, dW, R, sq)
    assert np.sqrt(ctx.to_host(sq, (1,))[0]) <= 1e-12 * np.linalg.norm(ref)


@pytest.mark.parametrize("nranks", [2, 3, 4, 8])
@pytest.mark.parametrize("lens,Q", [((12, 13, 14), 3), ((4, 13, 14), 3), ((9, 10, 8, 7), 3), ((6, 13, 14), 3), ((4, 7, 6, 5), 2),
                                    ((300, 20, 30), 10), ((37, 300, 6), 50)])
def test_mode0_contraction_on_shards_sums_to_the_whole(ctx, ppx, lens, Q, nranks):
    """What every rank of a sharded run does with mode 0 (host/als_Tucker.cxx ttm_mode, host/als_CP.cxx): contract its
    slab T[b:e] with rows b..e of the REPLICATED factor -- the factor's pointer advanced by b, its leading dimension the
    global extent -- for the Tucker TTM (rank in place), the CP first contraction (rank last) and the Hadamard-batched
    one.  The sum of the partial results over the ranks must be the contraction of the whole tensor.  Shards of one, two
    and three rows, odd offsets (round 2: the 4- and 8-GPU parity runs found the sum wrong for some of these)."""
    N = len(lens)
    if lens[0] < nranks:
        pytest.skip("fewer rows than ranks")
    T = rnd(lens, 700)
    W = rnd((lens[0], Q), 701)
    Wd = ctx.to_device(W)
    rest = lens[1:]
    ref_ttm = o.ttm(T, 0, W)
    ref_first = o.contract(o.letters(N)[1:] + "*", T, o.letters(N), W, "a*")
    Th = rnd(tuple(lens) + (Q,), 702)
    ref_mttv = o.contract(o.letters(N)[1:] + "*", Th, o.letters(N) + "*", W, "a*")
    acc_ttm, acc_first, acc_mttv = np.zeros_like(ref_ttm), np.zeros_like(ref_first), np.zeros_like(ref_mttv)
    for r in range(nranks):
        b, e = ppx.shard_range(lens[0], nranks, r)
        ll = (e - b,) + tuple(rest)
        Wr = Wd[b:]  # rows b.. of every column: the same buffer, pointer advanced by b, leading dimension lens[0]
        out = ctx.empty(Q * int(np.prod(rest)))
        ctx.ttm(ctx.to_device(T[b:e]), ll, 0, Wr, Q, out, ldw=lens[0])
        acc_ttm += ctx.to_host(out, (Q,) + tuple(rest))
        out1 = ctx.empty(Q * int(np.prod(rest)))
        ctx.ttm_first(ctx.to_device(T[b:e]), ll, 0, Wr, Q, out1, ldw=lens[0])
        acc_first += ctx.to_host(out1, tuple(rest) + (Q,))
        out2 = ctx.empty(Q * int(np.prod(rest)))
        ctx.mttv(ctx.to_device(Th[b:e]), ll, 0, Wr, Q, out2, ldw=lens[0])
        acc_mttv += ctx.to_host(out2, tuple(rest) + (Q,))
    assert rel_err(acc_ttm, ref_ttm) < 1e-12
    assert rel_err(acc_first, ref_first) < 1e-12
    assert rel_err(acc_mttv, ref_mttv) < 1e-12


@pytest.mark.parametrize("lens,x,Q", [((13, 9, 11), 0, 4), ((13, 9, 11), 1, 4), ((13, 9, 11), 2, 4),
                                      ((12, 10, 8, 6), 1, 3), ((40, 7, 40), 2, 40), ((5, 1, 6), 1, 2),
                                      # TMA path with the rank written in place (+ DFMA tail columns)
                                      ((256, 14, 5), 1, 26), ((22, 130), 0, 27), ((128, 9, 3, 2), 1, 35),
                                      ((300, 14, 8), 1, 26), ((46, 9, 16), 1, 40),
                                      # streaming kernels with the rank written in place (+ accumulate)
                                      ((40, 24, 30), 1, 10), ((41, 17, 29), 1, 3), ((40, 700), 0, 10), ((3, 900, 2), 0, 16),
                                      ((64, 20, 5), 2, 12)])
def test_tucker_ttm_and_acc(ctx, lens, x, Q):
    T = rnd(lens, 120)
    W = rnd((lens[x], Q), 121)
    ref = o.ttm(T, x, W)
    out = ctx.empty(ref.size)
    ctx.ttm(ctx.to_device(T), lens, x, ctx.to_device(W), Q, out)
    assert rel_err(ctx.to_host(out, ref.shape), ref) < 1e-12
    ctx.ttm(ctx.to_device(T), lens, x, ctx.to_device(W), Q, out, acc=True)
    assert rel_err(ctx.to_host(out, ref.shape), 2 * ref) < 1e-12


@pytest.mark.parametrize("lens,i", [((13, 9, 11), 0), ((13, 9, 11), 1), ((13, 9, 11), 2), ((70, 5, 66), 2),
                                    ((130, 40, 3), 0), ((4, 150, 5, 3), 1),
                                    # 128 x 128 tile kernel (X >= 192 and >= 4096 columns): mode first / middle / last,
                                    # ragged tiles, columns not a multiple of the chunk
                                    ((200, 70, 61), 0), ((67, 257, 63), 1), ((90, 47, 193), 2), ((384, 4100), 0)])
def test_unfold_gram(ctx, lens, i):
    T = rnd(lens, 130)
    ref = o.unroll_tensor_contraction(T, i)
    out = ctx.empty(ref.size)
    ctx.unfold_gram(ctx.to_device(T), lens, i, out)
    assert rel_err(ctx.to_host(out, ref.shape), ref) < 1e-12


@pytest.mark.parametrize("s,r", [(12, 3), (13, 5), (64, 8), (150, 10)])
def test_sym_eig_topk(ctx, s, r):
    A = rnd((s, 2 * s), 140)
    MTM = A @ A.T
    U, ev = ctx.empty(s * r), ctx.empty(r)
    ctx.sym_eig_topk(ctx.to_device(MTM), s, r, U, ev)
    Uh, evh = ctx.to_host(U, (s, r)), ctx.to_host(ev, (r,))
    w, Q = np.linalg.eigh(MTM)
    w, Q = w[::-1][:r], Q[:, ::-1][:, :r]
    assert np.allclose(evh, w, rtol=1e-12)
    assert np.abs(Uh.T @ Uh - np.eye(r)).max() < 1e-12
    # same vectors as LAPACK up to the sign of each column (what MTM.svd(U,S,VT,r) returns)
    assert np.abs(np.abs(np.sum(Uh * Q, axis=0)) - 1.0).max() < 1e-9
    ref = o.top_left_singular(MTM, r)
    assert np.abs(np.abs(np.sum(Uh * ref, axis=0)) - 1.0).max() < 1e-9


def test_sign_align(ctx):
    s, r = 30, 6
    U, Uref = rnd((s, r), 150), rnd((s, r), 151)
    Ud = ctx.to_device(U)
    ctx.sign_align(Ud, ctx.to_device(Uref), s, r)
    assert np.array_equal(ctx.to_host(Ud, (s, r)), o.sign_align(U, Uref))
    # zero reference -> every column flips (b > 0 ? 1 : -1 with b == 0; als_Tucker.cxx:636-641)
    Ud = ctx.to_device(U)
    ctx.sign_align(Ud, ctx.zeros(s * r), s, r)
    assert np.array_equal(ctx.to_host(Ud, (s, r)), -U)


def test_errors_are_reported_not_swallowed(ctx, ppx):
    out = ctx.empty(10)
    with pytest.raises(ppx.PpxError):
        ctx.ttm_first(out, (5, 2), 3, out, 2, out, ldw=5)  # x out of range
    with pytest.raises(ppx.PpxError):
        ctx.solve_update(out, out, out, 2, 5, mode=7)


MULTI_CASES = [
    # lens, x_first, n_modes, R
    ((12, 10, 8, 6), 2, 2, 4), ((12, 10, 8, 6), 0, 2, 4), ((12, 10, 8, 6), 1, 2, 5), ((12, 10, 8, 6), 1, 3, 3),
    ((7, 5, 6, 4, 5, 3), 3, 3, 3), ((7, 5, 6, 4, 5, 3), 0, 3, 3), ((13, 9, 11), 1, 2, 5), ((13, 9, 11), 0, 2, 5),
    ((20, 64, 70), 1, 2, 10),   # few rows, deep K: exercises the K split (20 rows, K = 4480)
    ((64, 70, 3), 0, 2, 50),    # k-major, 3 rows, K = 4480
    ((9, 7, 5), 0, 3, 2),       # everything contracted: a single row
    ((256, 6, 5, 3), 1, 2, 4), ((128, 30, 40), 1, 2, 50), ((10, 12, 130, 3), 0, 2, 50),  # TMA-eligible fused cases
    ((640, 4, 5, 6), 1, 3, 10), ((128, 9, 11), 1, 2, 26), ((12, 14, 140), 0, 2, 11), ((256, 5, 4, 3), 1, 3, 43),
    ((46, 6, 5, 16), 1, 2, 26),
    # the 8-GPU shard shape in small: L = 37 * 30 = 1110 rows (L % 16 != 0), two trailing modes fused, split K
    ((37, 30, 40, 24), 2, 2, 50), ((38, 30, 24, 16), 2, 2, 26), ((37, 30, 20, 24), 0, 2, 50),
]


@pytest.mark.parametrize("lens,x_first,n,R", MULTI_CASES)
def test_ttm_multi(ctx, lens, x_first, n, R):
    N = len(lens)
    V = rnd(lens, 160)
    Ws = [rnd((lens[x_first + j], R), 161 + j) for j in range(n)]
    idx = o.letters(N)
    ops = [V, idx]
    keep = idx
    for j in range(n):
        c = idx[x_first + j]
        ops += [Ws[j], c + "*"]
        keep = keep.replace(c, "")
    ref = o.contract(keep + "*", *ops)
    out = ctx.empty(ref.size)
    ctx.ttm_multi(ctx.to_device(V), lens, x_first, [ctx.to_device(w) for w in Ws], R, out)
    got = ctx.to_host(out, ref.shape if ref.ndim else (1,))
    assert rel_err(got.reshape(ref.shape), ref) < 1e-12


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("s,R", [(300, 50), (13, 5), (33, 7), (200, 96)])
def test_solve_update_g_fuses_hadamard(ctx, mode, s, R):
    N = 4
    Wf = [o.fill_uniform((s + j, R), 3, 170 + j) for j in range(N)]
    Gs = []
    for j, w in enumerate(Wf):
        G = ctx.empty(R * R)
        ctx.gram(ctx.to_device(w), s + j, R, G)
        Gs.append(G)
    skip, lam = 1, 1e-3
    S = np.ones((R, R))
    for j in range(N):
        if j != skip:
            S = S * (Wf[j].T @ Wf[j])
    S = S + lam * np.eye(R)
    M, W_old, W_init = rnd((s, R), 180), rnd((s, R), 181), rnd((s, R), 182)
    W_ref, dW_ref = o.SVD_solve_mod(M, W_init, S, 1.0)
    Wd, grad, dW = ctx.to_device(W_old), ctx.empty(s * R), ctx.empty(s * R)
    ctx.solve_update_g(ctx.to_device(M), Gs, skip, lam, Wd, s, R, W_init=ctx.to_device(W_init), mode=mode, grad=grad,
                       dW=dW)
    assert rel_err(ctx.to_host(Wd, (s, R)), W_ref) < 1e-9
    assert rel_err(ctx.to_host(dW, (s, R)), dW_ref) < 1e-9
    assert rel_err(ctx.to_host(grad, (s, R)), -M + W_old @ S) < 1e-12


@pytest.mark.parametrize("d,s", [(1, 7), (2, 5), (3, 4)])
def test_fill_laplacian(ctx, d, s):
    ref = o.laplacian_tensor(2 * d, s) if d > 1 else (2.0 * np.eye(s) - np.eye(s, k=1) - np.eye(s, k=-1))
    out = ctx.empty(ref.size)
    ctx.fill_laplacian(out, d, s)
    assert np.array_equal(ctx.to_host(out, ref.shape), ref)


@pytest.mark.parametrize("R", [1, 5, 16, 17, 50, 70])
def test_spd_factor_inverse(ctx, R):
    rng = np.random.default_rng(R)
    B = rng.random((3 * R + 4, R))
    S = np.asfortranarray(B.T @ B + 0.1 * np.eye(R))
    Z = ctx.empty(R * R)
    ctx.spd_factor_inverse(ctx.to_device(S), R, Z)
    ref = np.linalg.inv(np.linalg.cholesky(S))
    got = ctx.to_host(Z, (R, R))
    assert np.abs(np.triu(got, 1)).max() == 0.0
    assert rel_err(got, ref) < 1e-9


@pytest.mark.parametrize("ta,tb,m,n,k", [(0, 0, 13, 7, 5), (1, 0, 4, 9, 11), (0, 1, 300, 50, 50), (1, 1, 6, 6, 6),
                                         (0, 0, 5, 3, 0)])
def test_gemm_small(ctx, ta, tb, m, n, k):
    A = rnd((k, m) if ta else (m, k), 500)
    B = rnd((n, k) if tb else (k, n), 501)
    Cm = rnd((m, n), 502)
    ref = 0.75 * (A.T if ta else A) @ (B.T if tb else B) - 0.5 * Cm
    Cd = ctx.to_device(Cm)
    ctx.gemm_small(ta, tb, m, n, k, 0.75, ctx.to_device(A), max(1, A.shape[0]), ctx.to_device(B), max(1, B.shape[0]),
                   -0.5, Cd, m)
    assert rel_err(ctx.to_host(Cd, (m, n)), ref) < 1e-13


@pytest.mark.parametrize("Mtot,r,R", [(1000, 1, 7), (333, 3, 50), (70000, 5, 10), (17, 16, 3)])
def test_rank_expand_acc(ctx, Mtot, r, R):
    T, VT, out = rnd((Mtot, r), 510), rnd((r, R), 511), rnd((Mtot, R), 512)
    ref = out + T @ VT
    od = ctx.to_device(out)
    ctx.rank_expand_acc(ctx.to_device(T), Mtot, r, ctx.to_device(VT), r, R, od)
    assert rel_err(ctx.to_host(od, (Mtot, R)), ref) < 1e-13


# ---- leading eigenvectors at sizes where the Chebyshev-filtered subspace iteration takes over (n >= 384, r <= n/4) ----
def _spectrum_case(kind, n, rng):
    if kind == "noise+outlier":      # Gram of a positive random matrix: one eigenvalue 10^5-10^6 x the bulk
        Y = 0.5 + 0.5 * rng.random((n, 2 * n))
        return Y @ Y.T
    if kind == "noise":              # Gram of a centred random matrix: Marchenko-Pastur bulk, gaps of a fraction of a %
        Y = rng.standard_normal((n, 2 * n))
        return Y @ Y.T
    Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    if kind == "decay":              # geometric decay over 8 orders of magnitude
        lam = 10.0 ** (-8.0 * np.arange(n) / n)
    elif kind == "lowrank":          # exact rank 12 (fewer non-zero eigenvalues than asked for when r > 12)
        lam = np.r_[np.linspace(5.0, 1.0, 12), np.zeros(n - 12)]
    else:                            # clusters of equal eigenvalues straddling nothing: subspaces well defined as a whole
        lam = np.r_[np.repeat([9.0, 7.0, 5.0], 4), np.linspace(1.0, 0.1, n - 12)]
    return (Q * lam) @ Q.T


@pytest.mark.parametrize("kind,n,r", [("noise+outlier", 400, 8), ("noise+outlier", 800, 40), ("noise", 512, 24),
                                      ("decay", 448, 16), ("lowrank", 400, 10), ("lowrank", 400, 20),
                                      ("clusters", 384, 12), ("noise", 390, 88), ("noise+outlier", 448, 100)])
def test_sym_eig_topk_large(ctx, kind, n, r):
    rng = np.random.default_rng(7)
    A = _spectrum_case(kind, n, rng)
    A = np.asfortranarray(0.5 * (A + A.T))
    U, ev = ctx.empty(n * r), ctx.empty(r)
    basis = ctx.empty(n * n)
    w, Q = np.linalg.eigh(A)
    w, Q = w[::-1], Q[:, ::-1]
    for warm in (False, True):   # cold, then warm started from its own block (must give the same answer)
        ctx.sym_eig_topk(ctx.to_device(A), n, r, U, ev, basis=basis, basis_valid=warm)
        Uh, evh = ctx.to_host(U, (n, r)), ctx.to_host(ev, (r,))
        assert np.abs(Uh.T @ Uh - np.eye(r)).max() < 1e-11
        assert np.abs(evh - w[:r]).max() <= 1e-11 * w[0]
        # invariant-subspace test that is meaningful with repeated / zero eigenvalues: A U = U (U^T A U)
        resid = A @ Uh - Uh @ (Uh.T @ A @ Uh)
        assert np.abs(resid).max() <= 1e-10 * w[0]
        k = int(np.sum(w[:r] > 1e-9 * w[0]))
        if kind in ("noise+outlier", "noise", "decay"):   # simple eigenvalues: the projector is unique
            gap = np.min(w[:k] - w[1:k + 1]) if k < n else 1.0
            P, Pr = Uh[:, :k] @ Uh[:, :k].T, Q[:, :k] @ Q[:, :k].T
            assert np.abs(P - Pr).max() <= 1e-9 * w[0] / max(gap, 1e-300) * 1e-3 + 1e-9


def test_sym_eig_topk_ignores_a_foreign_basis(ctx):
    """The `basis` buffer is shared by the two solvers behind ppx_sym_eig_topk_warm: the subspace iteration keeps an
    n x p block there, the Jacobi fallback an n x n matrix.  A buffer marked valid that was not written by the
    subspace iteration -- here the eigenvectors in INCREASING order, whose first columns are orthogonal to everything
    wanted, as a Jacobi run leaves them -- must be recognised (tag behind the block) and a cold start taken."""
    n, r = 512, 16
    rng = np.random.default_rng(7)
    A = _spectrum_case("noise+outlier", n, rng)
    A = np.asfortranarray(0.5 * (A + A.T))
    w, Q = np.linalg.eigh(A)  # increasing
    basis = ctx.to_device(np.asfortranarray(Q))
    U, ev = ctx.empty(n * r), ctx.empty(r)
    ctx.sym_eig_topk(ctx.to_device(A), n, r, U, ev, basis=basis, basis_valid=True)
    evh, Uh = ctx.to_host(ev, (r,)), ctx.to_host(U, (n, r))
    assert np.abs(evh - w[::-1][:r]).max() <= 1e-11 * w[-1]
    assert np.abs(A @ Uh - Uh @ (Uh.T @ A @ Uh)).max() <= 1e-10 * w[-1]
    # and the block it left is its own: a warm restart reproduces the answer
    ctx.sym_eig_topk(ctx.to_device(A), n, r, U, ev, basis=basis, basis_valid=True)
    assert np.abs(ctx.to_host(ev, (r,)) - w[::-1][:r]).max() <= 1e-11 * w[-1]
