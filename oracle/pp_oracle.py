"""CPU oracle for the ALS / pairwise-perturbation (PP) sweeps of CP and Tucker decomposition.

TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import it, and only as the checker / CPU baseline.

PARITY UNPINNED: the reference (LinjianMa/pairwise-perturbation) cannot be compiled here (it needs Cyclops CTF +
MPI + ScaLAPACK, none installed, no network) and its single test (tests/test_decomposition.cxx:38-66) asserts only
`order` and `rank`; it ships no golden vectors.  This file is therefore a line-by-line NumPy float64 restatement of
the reference's algorithm (every function cites the reference file:line it follows); oracle/naive.c is an
independent loop-based restatement of the contractions that cross-checks it.

Conventions: tensors are numpy arrays indexed [a,b,c,...] = modes 0,1,2,...; the reference's
(CTF) global element order is first-index-fastest, i.e. `x.ravel(order="F")` is the raw buffer.  Factor matrices
W[i] are s_i x R.  In index strings '*' is the rank index; an index present in both inputs AND the output is a
Hadamard (batch) index, an index missing from the output is summed  -- exactly numpy.einsum semantics.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

_MASK = (1 << 64) - 1


# ----------------------------------------------------------------------------------------------------------------
# deterministic counter-based generator (replaces CTF fill_random, test_ALS.cxx:282,337-338)
# ----------------------------------------------------------------------------------------------------------------
def u01(seed: int, tensor_id: int, n: int, start: int = 0) -> np.ndarray:
    """u(seed, tensor_id, linear_index) in [0,1): SplitMix64 finaliser of a counter, top 53 bits.

    The same function is implemented in the CUDA library (ppx_fill_uniform) and in oracle/naive.c; linear_index
    is the position in the global first-index-fastest order, so any shard can generate its slice on its own."""
    with np.errstate(over="ignore"):
        idx = np.arange(start, start + n, dtype=np.uint64)
        base = np.uint64((seed * 0x9E3779B97F4A7C15 + tensor_id * 0xD1B54A32D192ED03) & _MASK)
        z = idx + base
        z = z + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def fill_uniform(shape, seed, tensor_id, lo=0.0, hi=1.0):
    n = int(np.prod(shape))
    return (lo + (hi - lo) * u01(seed, tensor_id, n)).reshape(shape, order="F")


# ----------------------------------------------------------------------------------------------------------------
# small einsum helper following CTF's index-string semantics
# ----------------------------------------------------------------------------------------------------------------
def _es(s: str) -> str:
    return s.replace("*", "z").replace("^", "y").replace("&", "x")


def contract(out: str, *ops):
    """contract("abd*", V, "abcd", W, "c*")  ==  CTF  out[abd*] = V[abcd]*W[c*]."""
    arrs = ops[0::2]
    idxs = ops[1::2]
    return np.einsum(",".join(_es(i) for i in idxs) + "->" + _es(out), *arrs, optimize=True)


def letters(n):
    return "".join(chr(ord("a") + i) for i in range(n))


# ----------------------------------------------------------------------------------------------------------------
# common.cxx
# ----------------------------------------------------------------------------------------------------------------
def construct_dimension_tree(parent: dict, sibling: dict, start: int, end: int):
    """common.cxx:225-270  balanced binary tree over modes [start,end]; nodes are mode-letter strings."""
    if end == start:
        return
    full = "".join(chr(ord("a") + i) for i in range(start, end + 1))
    if end == start + 1:
        a, b = full[0], full[1]
        parent[a] = full
        parent[b] = full
        sibling[a] = b
        sibling[b] = a
        return
    middle = (start + end) // 2
    left = "".join(chr(ord("a") + i) for i in range(start, middle + 1))
    right = "".join(chr(ord("a") + i) for i in range(middle + 1, end + 1))
    parent[left] = full
    sibling[left] = right
    construct_dimension_tree(parent, sibling, start, middle)
    sibling[right] = left
    parent[right] = full
    construct_dimension_tree(parent, sibling, middle + 1, end)


def mttkrp_map_DT(mttkrp_map: dict, parent, sibling, V, W, args: str):
    """common.cxx:20-133.  Fills mttkrp_map[args] = tensor that KEEPS modes `args` (+ rank last).

    First-level nodes (children of the root; the reference tests this as len(args) in {N/2, N/2+1}, common.cxx:29,
    which is the same set for every order the reference supports: 4, 6, 7, 8, ...) start from V and contract the
    sibling modes left to right: the first is the tensor-times-matrix GEMM (common.cxx:56), the others are
    Hadamard-batched (common.cxx:83).  Deeper nodes start from the cached parent (common.cxx:89-128)."""
    if args in mttkrp_map:
        return
    N = V.ndim
    par = parent[args]
    if len(par) == N:  # first level  (common.cxx:29-88)
        cur_idx = letters(N)
        cur = V
        for j, x in enumerate(sibling[args]):
            out_idx = cur_idx.replace("*", "").replace(x, "") + "*"
            cur = contract(out_idx, cur, cur_idx, W[ord(x) - 97], x + "*")
            cur_idx = out_idx
        mttkrp_map[args] = cur
        return
    if par not in mttkrp_map:  # common.cxx:89-91
        mttkrp_map_DT(mttkrp_map, parent, sibling, V, W, par)
    cur = mttkrp_map[par]
    cur_idx = par + "*"
    for x in sibling[args]:  # common.cxx:105-130
        out_idx = cur_idx.replace("*", "").replace(x, "") + "*"
        cur = contract(out_idx, cur, cur_idx, W[ord(x) - 97], x + "*")
        cur_idx = out_idx
    mttkrp_map[args] = cur


def build_V(W):
    """common.cxx:135-197  V = [[W_0 .. W_{N-1}]] built left to right (KRP, then the last-mode GEMM :194)."""
    N = len(W)
    cur = W[0]
    idx = "a*"
    for i in range(1, N - 1):
        out = letters(i + 1) + "*"
        cur = contract(out, cur, idx, W[i], chr(97 + i) + "*")
        idx = out
    return contract(letters(N), cur, idx, W[N - 1], chr(97 + N - 1) + "*")


def unroll_tensor_contraction(T, i):
    """common.cxx:205-223  MTM[p,q] = sum_rest T[..p..] T[..q..] (Gram of the mode-i unfolding)."""
    Ti = np.moveaxis(T, i, 0).reshape(T.shape[i], -1)
    return Ti @ Ti.T


def normalize(W):
    """common.cxx:680-688  every factor rescaled to the geometric-mean Frobenius norm (one scalar per factor)."""
    N = len(W)
    norm = 1.0
    for i in range(N):
        norm = norm * np.linalg.norm(W[i])
    norm = norm ** (1.0 / N)
    for i in range(N):
        W[i] = (norm / np.linalg.norm(W[i])) * W[i]


def svd_inverse(S):
    """common.cxx:715-722  S = U diag(s) VT ; S^-1 = V diag(1/s) U^T  (no truncation)."""
    U, s, VT = np.linalg.svd(S)
    return (VT.T * (1.0 / s)) @ U.T


def SVD_solve(M, S):
    """common.cxx:710-725  W = M S^-1."""
    return M @ svd_inverse(S)


def cholesky_solve(M, S):
    """common.cxx:727-737  S = L L^T; T = M L^-T; W = T L^-1."""
    import scipy.linalg as sla

    L = np.linalg.cholesky(S)
    T = sla.solve_triangular(L, M.T, lower=True).T  # T L^T = M
    return sla.solve_triangular(L.T, T.T, lower=False).T  # W L = T


def SVD_solve_mod(M, W_init, S, ratio_step):
    """common.cxx:739-758  returns (W, dW)."""
    W = M @ svd_inverse(S)
    dW = ratio_step * (W - W_init)
    if ratio_step != 1.0:
        W = W_init + dW
    return W, dW


def gram_hadamard(W, skip, lam=0.0, always_regul=False):
    """als_CP.cxx:288-292 / 573-579 / 796-802; cp_als_optimizer.cxx:20-38.
    S = Hadamard_{j != skip} W_j^T W_j, in increasing j (the order of `index[]`)."""
    N = len(W)
    idx = [j for j in range(N) if j != skip]
    # the reference orders index[] by swapping mode i with the last one (als_CP.cxx:219-231); Hadamard products
    # commute exactly in floating point only up to ordering, so reproduce that order:
    seq = list(range(N))
    seq[skip], seq[N - 1] = seq[N - 1], seq[skip]
    idx = seq[: N - 1]
    S = W[idx[0]].T @ W[idx[0]]
    for j in idx[1:]:
        S = S * (W[j].T @ W[j])
    if always_regul or lam != 0:
        S = S + lam * np.eye(S.shape[0])
    return S


def KhatriRao_contract(V, W, index):
    """common.cxx:931-997  contract V with W[index[0]] (GEMM :963) then W[index[1..N-2]] Hadamard-batched (:992)."""
    N = V.ndim
    cur = V
    cur_idx = letters(N)
    for j in range(N - 1):
        x = chr(97 + index[j])
        out_idx = "".join(chr(97 + index[jj]) for jj in range(j + 1, N)) + "*"
        cur = contract(out_idx, cur, cur_idx, W[index[j]], x + "*")
        cur_idx = out_idx
    return cur


# ----------------------------------------------------------------------------------------------------------------
# als_CP.cxx
# ----------------------------------------------------------------------------------------------------------------
@dataclass
class Trace:
    """What the reference prints / writes to the CSV (als_CP.cxx:193-198) plus the switching markers (:1110,:1118)."""

    rows: list = field(default_factory=list)  # (iter, gradnorm, pp_update, diffV)
    events: list = field(default_factory=list)  # ("DT", iter) / ("PP", iter)
    sweeps: list = field(default_factory=list)  # ("DT"|"PP"|"PPinit", iter) one entry per executed sweep


def cp_residual(V, W):
    """als_CP.cxx:183-187."""
    return float(np.linalg.norm((V - build_V(W)).ravel()))


def _leaf_M(mttkrp_map, parent, sibling, V, W, i):
    """als_CP.cxx:236-284 (same at :520-569).  Leaf MTTKRP from the cached parent.

    Extension for order 3 (the reference recurses forever there): a leaf whose parent is the root is
    contracted directly from V with the first-level rule."""
    N = V.ndim
    a = chr(97 + i)
    par = parent[a]
    if len(par) == N:
        tmp = {}
        mttkrp_map_DT(tmp, parent, sibling, V, W, a)
        return tmp[a]
    if par not in mttkrp_map:
        mttkrp_map_DT(mttkrp_map, parent, sibling, V, W, par)
    T = mttkrp_map[par]
    sib = [c for c in par if c != a]
    ops = [T, par + "*"]
    for c in sib:
        ops += [W[ord(c) - 97], c + "*"]
    return contract(a + "*", *ops)


def alsCP_DT(V, W, grad_W, tol, maxiter, lam=0.0, resprint=10, F=None, trace=None, want_residual=True):
    """als_CP.cxx:127-320.  Returns (stopped_early, trace).  W, grad_W lists are updated in place."""
    N = V.ndim
    trace = trace if trace is not None else Trace()
    parent, sibling = {}, {}
    construct_dimension_tree(parent, sibling, 0, N - 1)
    it = 0
    projnorm = 0.0
    for it in range(0, maxiter + 2):
        if it > maxiter:
            break
        if it % resprint == 0 or it == maxiter:
            projnorm = math.sqrt(sum(np.linalg.norm(g) ** 2 for g in grad_W))
            diff = cp_residual(V, W) if want_residual else float("nan")
            trace.rows.append((it, projnorm, 0, diff))
            if projnorm < tol:
                break
        mttkrp_map = {}
        for i in range(N):
            M = _leaf_M(mttkrp_map, parent, sibling, V, W, i)
            S = gram_hadamard(W, i, lam, always_regul=True)
            if F is not None:
                M = M + F[i]
            grad_W[i] = -M + W[i] @ S
            W[i] = SVD_solve(M, S)
        normalize(W)
        trace.sweeps.append(("DT", it))
    return (it != maxiter + 1), trace


def stringbuilder_mttkrp(seq, N):
    """als_CP.cxx:323-350  contracted modes -> kept modes + '*'."""
    return "".join(c for c in letters(N) if c not in seq) + "*"


def Build_mttkrp_map(mttkrp_map, V, W, seq):
    """als_CP.cxx:352-409.  Key = contracted modes (increasing)."""
    N = V.ndim
    if len(seq) == 1:
        mttkrp_map[seq] = contract(stringbuilder_mttkrp(seq, N), V, letters(N), W[ord(seq) - 97], seq + "*")
        return
    seq2 = seq[:-1]
    if seq2 not in mttkrp_map:
        Build_mttkrp_map(mttkrp_map, V, W, seq2)
    x = seq[-1]
    mttkrp_map[seq] = contract(stringbuilder_mttkrp(seq, N), mttkrp_map[seq2], stringbuilder_mttkrp(seq2, N),
                               W[ord(x) - 97], x + "*")


def build_pp_operators(V, W):
    """als_CP.cxx:676-694  all pair operators, then all singles."""
    N = V.ndim
    seq = letters(N)
    m = {}
    for ii in range(N):
        for jj in range(ii + 1, N):
            Build_mttkrp_map(m, V, W, "".join(c for k, c in enumerate(seq) if k not in (ii, jj)))
    for ii in range(N):
        Build_mttkrp_map(m, V, W, "".join(c for k, c in enumerate(seq) if k != ii))
    return m


def pp_corrected_M(mttkrp_map, dW, i, N):
    """als_CP.cxx:774-794."""
    seq = letters(N)
    M = mttkrp_map["".join(c for k, c in enumerate(seq) if k != i)].copy()
    for ii in range(i):
        key = "".join(c for k, c in enumerate(seq) if k not in (ii, i))
        M = M + contract("jk", mttkrp_map[key], "ijk", dW[ii], "ik")
    for ii in range(i + 1, N):
        key = "".join(c for k, c in enumerate(seq) if k not in (ii, i))
        M = M + contract("ik", mttkrp_map[key], "ijk", dW[ii], "jk")
    return M


class _State:
    pass


def alsCP_DT_sub(V, W, grad_W, dW, st, tol, tol_init, maxiter, lam, resprint, trace, want_residual=True):
    """als_CP.cxx:418-612.  st.iter / st.projnorm are the by-reference `iter` / `projnorm`."""
    N = V.ndim
    W_prev = [np.zeros_like(w) for w in W]  # :428-431 zero-initialised
    parent, sibling = {}, {}
    construct_dimension_tree(parent, sibling, 0, N - 1)
    while st.iter <= maxiter:
        it = st.iter
        if it % resprint == 0 or it == maxiter:
            st.projnorm = math.sqrt(sum(np.linalg.norm(g) ** 2 for g in grad_W))
            diff = cp_residual(V, W) if want_residual else float("nan")
            trace.rows.append((it, st.projnorm, 0, diff))
            if st.projnorm < tol:
                break
        mttkrp_map = {}
        for i in range(N):
            M = _leaf_M(mttkrp_map, parent, sibling, V, W, i)
            S = gram_hadamard(W, i, lam, always_regul=False)
            grad_W[i] = -M + W[i] @ S
            W[i] = SVD_solve(M, S)
        normalize(W)
        trace.sweeps.append(("DT", it))
        num_dw_break = 0
        for i in range(N):
            dW[i] = W[i] - W_prev[i]
            W_prev[i] = W[i].copy()
            if abs(np.linalg.norm(dW[i]) / np.linalg.norm(W[i])) < tol_init:
                num_dw_break += 1
        if num_dw_break == N:
            return  # :604-605 returns BEFORE iter++ of this sweep
        st.iter += 1


def alsCP_PP_sub(V, W, grad_W, dW, st, tol, tol_init, maxiter, lam, ratio_step, resprint, bench, trace,
                 want_residual=True):
    """als_CP.cxx:621-833."""
    N = V.ndim
    init_iter = st.iter
    W_init = [None] * N
    mttkrp_map = {}
    while st.iter <= maxiter:
        it = st.iter
        num_dw_break = 0
        if not bench:
            for i in range(N):
                if abs(np.linalg.norm(dW[i]) / np.linalg.norm(W[i])) > tol_init:
                    num_dw_break += 1
        if (it - init_iter) % 15 == 0 or num_dw_break > 0:
            if num_dw_break > 0 or it != init_iter:
                return
            for j in range(N):
                W_init[j] = W[j].copy()
                dW[j] = np.zeros_like(W[j])
            mttkrp_map = build_pp_operators(V, W)
            trace.sweeps.append(("PPinit", it))
        if it % resprint == 0 or it == maxiter or it == init_iter:
            st.projnorm = math.sqrt(sum(np.linalg.norm(g) ** 2 for g in grad_W))
            diff = cp_residual(V, W) if want_residual else float("nan")
            trace.rows.append((it, st.projnorm, 1, diff))
            if st.projnorm < tol:
                break
        for i in range(N):
            M = pp_corrected_M(mttkrp_map, dW, i, N)
            S = gram_hadamard(W, i, lam, always_regul=False)
            grad_W[i] = -M + W[i] @ S
            W[i], dW[i] = SVD_solve_mod(M, W_init[i], S, ratio_step)
        normalize(W)
        trace.sweeps.append(("PP", it))
        st.iter += 1
    if bench:
        st.iter += 1


def alsCP_PP(V, W, grad_W, tol, tol_init, maxiter, lam=0.0, ratio_step=1.0, resprint=10, bench=False,
             want_residual=True):
    """als_CP.cxx:1082-1137.  Returns (stopped_early, trace)."""
    N = V.ndim
    trace = Trace()
    st = _State()
    st.iter = 0
    st.projnorm = 10.0
    dW = [np.zeros_like(w) for w in W]
    while st.projnorm > tol and st.iter <= maxiter:
        if not bench:
            trace.events.append(("DT", st.iter))
            alsCP_DT_sub(V, W, grad_W, dW, st, tol, tol_init, maxiter, lam, resprint, trace, want_residual)
        trace.events.append(("PP", st.iter))
        alsCP_PP_sub(V, W, grad_W, dW, st, tol, tol_init, maxiter, lam, ratio_step, resprint, bench, trace,
                     want_residual)
    return (st.iter != maxiter + 1), trace


def sort_indexes(v):
    """als_CP.cxx:835-843  indices sorted by decreasing value (std::sort, not stable; ties keep index order here)."""
    return sorted(range(len(v)), key=lambda k: (-v[k], k))


def alsCP_PP_partupdate_sub(V, W, grad_W, dW, st, tol, tol_init, maxiter, update_percentage, lam, ratio_step,
                            resprint, bench, trace, want_residual=True):
    """als_CP.cxx:852-1073."""
    N = V.ndim
    seq = letters(N)
    init_iter = st.iter
    W_init = [None] * N
    mttkrp_map = {}
    dM = [np.zeros_like(w) for w in W]
    M = [np.zeros_like(w) for w in W]
    rel = [0.0] * N
    update_size = int(N * update_percentage)
    while st.iter <= maxiter:
        it = st.iter
        num_dw_break = 0
        if not bench:
            for i in range(N):
                if abs(np.linalg.norm(dW[i]) / np.linalg.norm(W[i])) > tol_init:
                    num_dw_break += 1
        if (it - init_iter) % 15 == 0 or num_dw_break > 0:
            if num_dw_break > 0 or it != init_iter:
                return
            for j in range(N):
                W_init[j] = W[j].copy()
                dW[j] = np.zeros_like(W[j])
            mttkrp_map = build_pp_operators(V, W)
            trace.sweeps.append(("PPinit", it))
        if it % resprint == 0 or it == maxiter or it == init_iter:
            st.projnorm = math.sqrt(sum(np.linalg.norm(g) ** 2 for g in grad_W))
            diff = cp_residual(V, W) if want_residual else float("nan")
            trace.rows.append((it, st.projnorm, 1, diff))
            if st.projnorm < tol:
                break
        order = sort_indexes(rel)
        for i in order[:update_size]:
            M[i] = mttkrp_map["".join(c for k, c in enumerate(seq) if k != i)] + dM[i]  # :1025
            S = gram_hadamard(W, i, lam, always_regul=False)
            grad_W[i] = -M[i] + W[i] @ S
            W[i], dW[i] = SVD_solve_mod(M[i], W_init[i], S, ratio_step)
            dM[i] = np.zeros_like(dM[i])
            for ii in range(i):  # :1038-1045
                key = "".join(c for k, c in enumerate(seq) if k not in (ii, i))
                dM[ii] = dM[ii] + contract("ik", mttkrp_map[key], "ijk", dW[i], "jk")
            for ii in range(i + 1, N):  # :1046-1053
                key = "".join(c for k, c in enumerate(seq) if k not in (ii, i))
                dM[ii] = dM[ii] + contract("jk", mttkrp_map[key], "ijk", dW[i], "ik")
        for i in range(N):  # :1060-1064  (0/0 -> nan exactly as the reference does before M is ever formed)
            with np.errstate(divide="ignore", invalid="ignore"):
                rel[i] = float(np.float64(np.linalg.norm(dM[i])) / np.float64(np.linalg.norm(M[i])))
        normalize(W)
        trace.sweeps.append(("PP", it))
        st.iter += 1
    if bench:
        st.iter += 1


def alsCP_PP_partupdate(V, W, grad_W, tol, tol_init, maxiter, lam=0.0, ratio_step=1.0, update_percentage=1.0,
                        resprint=10, bench=False, want_residual=True):
    """als_CP.cxx:1146-1207."""
    trace = Trace()
    st = _State()
    st.iter = 0
    st.projnorm = 10.0
    dW = [np.zeros_like(w) for w in W]
    while st.projnorm > tol and st.iter <= maxiter:
        if not bench:
            trace.events.append(("DT", st.iter))
            alsCP_DT_sub(V, W, grad_W, dW, st, tol, tol_init, maxiter, lam, resprint, trace, want_residual)
        trace.events.append(("PP", st.iter))
        alsCP_PP_partupdate_sub(V, W, grad_W, dW, st, tol, tol_init, maxiter, update_percentage, lam, ratio_step,
                                resprint, bench, trace, want_residual)
    return (st.iter != maxiter + 1), trace


# ----------------------------------------------------------------------------------------------------------------
# src/optimizer/*  and  src/CP.cxx
# ----------------------------------------------------------------------------------------------------------------
class CPOptimizer:
    """cp_als_optimizer.{h,cxx}."""

    def __init__(self, order, r):
        self.order, self.rank = order, r
        self.V = self.W = self.grad_W = None
        self.lam = 0.0

    def configure(self, V, W, grad_W, lam):
        assert V.ndim == self.order
        for w in W:
            assert w.shape[1] == self.rank
        self.V, self.W, self.grad_W, self.lam = V, W, grad_W, lam

    def update_S(self, update_index):
        """cp_als_optimizer.cxx:20-38  (increasing-j order; lambda*I always added)."""
        idx = [j for j in range(self.order) if j != update_index]
        S = self.W[idx[0]].T @ self.W[idx[0]]
        for j in idx[1:]:
            S = S * (self.W[j].T @ self.W[j])
        return S + self.lam * np.eye(self.rank)

    def _solve_mode(self, i, M):
        S = self.update_S(i)
        self.grad_W[i] = -M + self.W[i] @ S
        self.W[i] = cholesky_solve(M, S)


class CPSimpleOptimizer(CPOptimizer):
    """cp_simple_optimizer.cxx:23-56."""

    def step(self):
        N = self.order
        for i in range(N):
            seq = list(range(N))
            seq[i], seq[N - 1] = seq[N - 1], seq[i]
            M = KhatriRao_contract(self.V, self.W, seq)
            self._solve_mode(i, M)
        return 1.0


def vec2str(vec):
    """common.cxx:10-18."""
    return "".join(chr(97 + v) for v in vec) + "*"


class CPDTOptimizer(CPOptimizer):
    """cp_dt_optimizer.{h,cxx}: caterpillar tree over the local indices 0..N-2 of the rotated mode list."""

    def __init__(self, order, r):
        super().__init__(order, r)
        self.parent, self.contract_index = {}, {}
        self._construct_subtree(list(range(order - 1)))  # :68-76
        self.indexes1 = list(range(order - 1))
        self.left_index1 = order - 1
        self.left_index2 = (self.left_index1 + order - 1) % order
        self.indexes2 = self._rot(self.left_index2)
        self.special_index = 0
        self.first_subtree = True
        self.mttkrp_map = {}

    def _rot(self, left_index):
        """update_indexes :52-66."""
        return list(range(left_index + 1, self.order)) + list(range(0, left_index))

    def _construct_subtree(self, top):
        """:78-100."""
        self._right_subtree(top)
        child = top[:-1]
        self.parent[vec2str(child)] = vec2str(top)
        self.contract_index[vec2str(child)] = vec2str([top[-1]])
        if len(child) > 1:
            self._construct_subtree(child)

    def _right_subtree(self, top):
        """:102-124."""
        child = top[:-1]
        child[-1] = top[-1]
        self.parent[vec2str(child)] = vec2str(top)
        self.contract_index[vec2str(child)] = vec2str([top[-2]])
        if len(child) > 1:
            self._right_subtree(child)

    def mttkrp_map_init(self, left_index):
        """:127-160  root = V x W[left_index]; kept modes rotated: left+1..N-1, 0..left-1."""
        N = self.order
        out = "".join(chr(97 + i) for i in self._rot(left_index)) + "*"
        self.mttkrp_map[vec2str(list(range(N - 1)))] = contract(out, self.V, letters(N), self.W[left_index],
                                                                chr(97 + left_index) + "*")

    def mttkrp_map_DT(self, index):
        """:163-186."""
        par = self.parent[index]
        if par not in self.mttkrp_map:
            self.mttkrp_map_DT(par)
        mat = self.contract_index[index]
        w = self.indexes[ord(mat[0]) - 97]
        self.mttkrp_map[index] = contract(index, self.mttkrp_map[par], par, self.W[w], mat)

    def step(self):
        """:188-238."""
        if self.first_subtree:
            self.indexes, self.left_index = self.indexes1, self.left_index1
        else:
            self.indexes, self.left_index = self.indexes2, self.left_index2
        self.mttkrp_map = {}
        self.mttkrp_map_init(self.left_index)
        for i in range(len(self.indexes)):
            if self.first_subtree and i < self.special_index:
                continue
            if (not self.first_subtree) and i > self.special_index:
                break
            key = vec2str([i])
            if key not in self.mttkrp_map:
                self.mttkrp_map_DT(key)
            self._solve_mode(self.indexes[i], self.mttkrp_map[key])
        self.first_subtree = not self.first_subtree
        return 0.5


class CPMSDTOptimizer(CPDTOptimizer):
    """cp_msdt_optimizer.{h,cxx}: the root mode rotates N-1, N-2, ..., every step updates the other N-1 modes."""

    def __init__(self, order, r):
        super().__init__(order, r)
        self.left_index = order
        self.indexes = list(range(order - 1))

    def step(self):
        """:173-208."""
        N = self.order
        self.mttkrp_map = {}
        self.left_index = (self.left_index + N - 1) % N  # update_indexes :36-49
        self.indexes = self._rot(self.left_index)
        self.mttkrp_map_init(self.left_index)
        for i in range(len(self.indexes)):
            key = vec2str([i])
            if key not in self.mttkrp_map:
                self.mttkrp_map_DT(key)
            self._solve_mode(self.indexes[i], self.mttkrp_map[key])
        return 1.0 * (N - 1) / N



# ----------------------------------------------------------------------------------------------------------------
# low-rank-update optimizers (src/optimizer/cp_dt_lr_optimizer.cxx, cp_msdt_lr_optimizer.cxx; run.cxx -pp 2 / 3)
# ----------------------------------------------------------------------------------------------------------------
def randomized_range(X, r, seed, draw_id):
    """randomized_svd (common.cxx:691-708) with iter = 1, reduced to what its callers use: U s VT2 = B Q^T = X Q Q^T
    with Q = orth(X^T X orth(Omega)).  span(X^T X orth(Omega)) = span(X^T X Omega), so Q = orth(X^T X Omega);
    Omega = u(seed, draw_id) in [0,1), n x r."""
    n = X.shape[1]
    Omega = fill_uniform((n, r), seed, draw_id)
    Q, _ = np.linalg.qr((X.T @ X) @ Omega)
    return Q


def get_rankR_update_cholesky(r, M, A, gamma, random=False, seed=1, draw_id=5000):
    """common.cxx:768-786.  Returns (U s, VT): the rank-r update of A towards M gamma^-1 is (U s) @ VT.
    (The reference returns U, s, VT separately; every caller only ever uses U diag(s) and VT.)"""
    L = np.linalg.cholesky(gamma)
    Z = np.linalg.inv(L)                  # lower triangular
    rhs = M - A @ gamma
    X = rhs @ Z.T                         # solve_tri(L, X, lower, from the right, transposed): X L^T = rhs
    if random:
        Q = randomized_range(X, r, seed, draw_id)
    else:
        _, _, VTfull = np.linalg.svd(X, full_matrices=False)
        Q = VTfull[:r].T                  # leading r right singular vectors
    Us = X @ Q                            # = U diag(s)
    VT = Q.T @ Z                          # xVT.solve_tri(L, xVT, lower, from the right): xVT_new L = xVT
    return Us, VT


class CPDTLROptimizer(CPDTOptimizer):
    """cp_dt_lr_optimizer.cxx: dimension tree whose two root contractions are refreshed by rank-`update_rank`
    patches  cached += V x (U s) x VT  for num_subiteration - 2 of every num_subiteration sweeps."""

    def __init__(self, order, r, update_rank, randomsvd=0, seed=1):
        super().__init__(order, r)
        self.randomsvd = randomsvd > 0
        self.num_subiteration = 5
        self.lr_rank = update_rank
        self.cached = {True: None, False: None}  # first_subtree -> cached root tensor
        self.seed, self.draw = seed, 5000
        self.initialize_low_rank_param()

    def initialize_low_rank_param(self):
        self.count_subiteration = 0
        self.low_rank_decomp = False

    def _root_modes(self, left_index):
        return "".join(chr(97 + i) for i in self._rot(left_index)) + "*"

    def mttkrp_map_init(self, left_index):
        """:39-99."""
        N = self.order
        top = vec2str(list(range(N - 1)))
        if self.low_rank_decomp and self.count_subiteration > 1:
            self.update_cached_tensor(left_index)
            self.mttkrp_map[top] = self.cached[self.first_subtree].copy()
        else:
            self.mttkrp_map[top] = contract(self._root_modes(left_index), self.V, letters(N), self.W[left_index],
                                            chr(97 + left_index) + "*")
            self.cached[self.first_subtree] = self.mttkrp_map[top].copy()

    def update_cached_tensor(self, left_index):
        """:127-159  cached += V x_left (U s) x VT."""
        N = self.order
        patch = contract(self._root_modes(left_index), self.V, letters(N), self.Us, chr(97 + left_index) + "&",
                         self.VT, "&*")
        self.cached[self.first_subtree] = self.cached[self.first_subtree] + patch

    def step(self):
        """:161-236."""
        N = self.order
        if self.first_subtree:
            self.indexes, self.left_index = self.indexes1, self.left_index1
        else:
            self.indexes, self.left_index = self.indexes2, self.left_index2
        self.mttkrp_map = {}
        self.mttkrp_map_init(self.left_index)
        n = len(self.indexes)
        for i in range(n):
            if self.first_subtree and i < self.special_index:
                continue
            if (not self.first_subtree) and i > self.special_index:
                break
            key = vec2str([i])
            if key not in self.mttkrp_map:
                self.mttkrp_map_DT(key)
            M = self.mttkrp_map[key]
            m = self.indexes[i]
            S = self.update_S(m)
            self.grad_W[m] = -M + self.W[m] @ S
            lr_slot = (self.first_subtree and i == n - 1) or ((not self.first_subtree) and i == 0)
            if lr_slot and self.count_subiteration >= 1:
                self.Us, self.VT = get_rankR_update_cholesky(self.lr_rank, M, self.W[m], S, self.randomsvd, self.seed,
                                                             self.draw)
                self.draw += 1
                self.W[m] = self.W[m] + self.Us @ self.VT
                self.low_rank_decomp = True
            else:
                self.W[m] = cholesky_solve(M, S)
        if not self.first_subtree:
            self.count_subiteration += 1
        if self.count_subiteration == self.num_subiteration and not self.first_subtree:
            self.special_index = (self.special_index + 1) % (N - 1)
            self.initialize_low_rank_param()
            if self.special_index != 0:
                self.left_index1 = (self.left_index1 + N - 1) % N
                self.left_index2 = (self.left_index2 + N - 1) % N
            else:
                self.left_index = N - 1
                self.left_index1 = self.left_index
                self.left_index2 = (self.left_index + N - 1) % N
            self.indexes1 = self._rot(self.left_index1)
            self.indexes2 = self._rot(self.left_index2)
        self.first_subtree = not self.first_subtree
        return 0.5


class CPMSDTLROptimizer(CPMSDTOptimizer):
    """cp_msdt_lr_optimizer.cxx: multi-sweep dimension tree with one cached root tensor per mode, refreshed by a
    rank-`update_rank` patch when that mode comes round as the root again."""

    def __init__(self, order, r, update_rank, randomsvd=0, seed=1):
        super().__init__(order, r)
        self.randomsvd = randomsvd > 0
        self.lr_rank = update_rank
        self.low_rank_decomp = False
        self.is_cached = [False] * order
        self.cached_tensors = [None] * order
        self.old_W = [None] * order
        self.seed, self.draw = seed, 5000

    def _root_modes(self, left_index):
        return "".join(chr(97 + i) for i in self._rot(left_index)) + "*"

    def mttkrp_map_init(self, left_index):
        """:35-80."""
        N = self.order
        top = vec2str(list(range(N - 1)))
        if self.low_rank_decomp and self.is_cached[left_index]:
            self.update_cached_tensor(left_index)
            self.mttkrp_map[top] = self.cached_tensors[left_index].copy()
        else:
            self.mttkrp_map[top] = contract(self._root_modes(left_index), self.V, letters(N), self.W[left_index],
                                            chr(97 + left_index) + "*")
            self.cached_tensors[left_index] = self.mttkrp_map[top].copy()
            self.old_W[left_index] = self.W[left_index].copy()
            self.is_cached[left_index] = True

    def update_cached_tensor(self, left_index):
        """:108-151."""
        N = self.order
        patch = contract(self._root_modes(left_index), self.V, letters(N), self.Us, chr(97 + left_index) + "&",
                         self.VT, "&*")
        self.cached_tensors[left_index] = self.cached_tensors[left_index] + patch
        self.old_W[left_index] = self.W[left_index].copy()
        self.is_cached[left_index] = True

    def step(self):
        """:153-205."""
        N = self.order
        self.mttkrp_map = {}
        self.left_index = (self.left_index + N - 1) % N
        self.indexes = self._rot(self.left_index)
        self.mttkrp_map_init(self.left_index)
        n = len(self.indexes)
        for i in range(n):
            key = vec2str([i])
            if key not in self.mttkrp_map:
                self.mttkrp_map_DT(key)
            M = self.mttkrp_map[key]
            m = self.indexes[i]
            S = self.update_S(m)
            self.grad_W[m] = -M + self.W[m] @ S
            if (not self.is_cached[m]) or i != n - 1:
                self.W[m] = cholesky_solve(M, S)
            else:
                self.Us, self.VT = get_rankR_update_cholesky(self.lr_rank, M, self.old_W[m], S, self.randomsvd,
                                                             self.seed, self.draw)
                self.draw += 1
                self.W[m] = self.old_W[m] + self.Us @ self.VT
                self.low_rank_decomp = True
        return 1.0 * (N - 1) / N

class CPD:
    """src/decomposition.{h,cxx} + src/CP.{h,cxx}."""

    def __init__(self, order, size, r, optimizer_cls, *opt_args):
        self.order = order
        self.size = [size] * order if np.isscalar(size) else list(size)
        self.rank = [r] * order if np.isscalar(r) else list(r)
        self.optimizer = optimizer_cls(order, self.rank[0], *opt_args)  # (update_rank, randomsvd) for the LR ones
        self.V = self.W = self.grad_W = None
        self.gradnorm = 0.0

    def Init(self, V, W, lam=0.0, grad_W=None):
        """src/CP.cxx:67-84 (grad_W is random in the reference; here zeros unless given)."""
        assert V.ndim == self.order
        for i in range(self.order):
            assert V.shape[i] == self.size[i] and W[i].shape[1] == self.rank[i]
        self.V, self.W = V, W
        self.grad_W = grad_W if grad_W is not None else [np.zeros_like(w) for w in W]
        self.optimizer.configure(V, W, self.grad_W, lam)

    def update_gradnorm(self):
        """src/CP.cxx:101-108."""
        self.gradnorm = math.sqrt(sum(np.linalg.norm(g) ** 2 for g in self.grad_W))

    def als(self, tol, maxsweep, resprint, want_residual=True):
        """src/CP.cxx:110-187.  Returns (stopped_early, rows[(sweeps, gradnorm, residual)])."""
        iters, sweeps = 0, 0.0
        rows = []
        while int(sweeps) <= maxsweep:
            if iters % resprint == 0 or sweeps >= maxsweep or sweeps == 0:
                self.update_gradnorm()
                diff = cp_residual(self.V, self.W) if want_residual else float("nan")
                rows.append((sweeps, self.gradnorm, diff))
                if self.gradnorm < tol:
                    break
            sweeps += self.optimizer.step()
            iters += 1
        return (sweeps != maxsweep + 1), rows


# ----------------------------------------------------------------------------------------------------------------
# als_Tucker.cxx
# ----------------------------------------------------------------------------------------------------------------
def top_left_singular(MTM, r):
    """MTM.svd(U,S,VT,r) on a symmetric PSD matrix: leading r left singular vectors (als_Tucker.cxx:20,402)."""
    U, s, VT = np.linalg.svd(MTM)
    return U[:, :r].copy()


def get_factor_matrices(T, ranks):
    """als_Tucker.cxx:12-23."""
    return [top_left_singular(unroll_tensor_contraction(T, i), ranks[i]) for i in range(T.ndim)]


def ttm(T, x, Wx):
    """Y[..k..] = sum_x T[..x..] W[x,k]  (rank replaces mode x in place; als_Tucker.cxx:102,224,465)."""
    return np.moveaxis(np.tensordot(T, Wx, axes=([x], [0])), -1, x)


def TTMc(V, W, i):
    """als_Tucker.cxx:76-110  contract every mode except i, in increasing mode order; i=-1 -> all."""
    Y = V
    for index in range(V.ndim):
        if index != i:
            Y = ttm(Y, index, W[index])
    return Y


def get_core_tensor(T, W):
    """als_Tucker.cxx:25-64  core = T x_0 W_0^T ... (transpose[ai]*T[i...] == contraction with W[i,a])."""
    return TTMc(T, W, -1)


def hosvd(T, ranks):
    """als_Tucker.cxx:66-70."""
    W = get_factor_matrices(T, ranks)
    return get_core_tensor(T, W), W


def tucker_residual(V, core, W):
    """als_Tucker.cxx:295-310  V_check = core x_j W_j^T ; ||V_check - V||."""
    Vc = core
    for j in range(V.ndim):
        Vc = ttm(Vc, j, W[j].T)
    return float(np.linalg.norm((Vc - V).ravel()))


def ttmc_map_DT(ttmc_map, parent, sibling, V, W, args):
    """als_Tucker.cxx:178-230."""
    if args in ttmc_map:
        return
    N = V.ndim
    if len(parent[args]) == N:
        cur = V
    else:
        if parent[args] not in ttmc_map:
            ttmc_map_DT(ttmc_map, parent, sibling, V, W, parent[args])
        cur = ttmc_map[parent[args]]
    for x in sibling[args]:
        cur = ttm(cur, ord(x) - 97, W[ord(x) - 97])
    ttmc_map[args] = cur


def _tucker_leaf_Y(ttmc_map, parent, sibling, V, W, i):
    """als_Tucker.cxx:356-394 (and :581-619).  Order-3 extension as in _leaf_M."""
    N = V.ndim
    a = chr(97 + i)
    par = parent[a]
    if len(par) == N:
        cur = V
    else:
        if par not in ttmc_map:
            ttmc_map_DT(ttmc_map, parent, sibling, V, W, par)
        cur = ttmc_map[par]
    for c in par:
        if c != a:
            cur = ttm(cur, ord(c) - 97, W[ord(c) - 97])
    return cur


def sign_align(U, Wref):
    """als_Tucker.cxx:632-643  U <- U diag(sign(diag(U^T Wref))), sign(b)= +1 if b>0 else -1."""
    d = np.einsum("ji,ji->i", U, Wref)
    return U * np.where(d > 0, 1.0, -1.0)[None, :]


def alsTucker_DT(V, core, W, tol, maxiter, resprint=10, want_residual=True):
    """als_Tucker.cxx:240-424.  Returns (stopped_early, rows[(iter, diffnorm, diffV)], core)."""
    N = V.ndim
    parent, sibling = {}, {}
    construct_dimension_tree(parent, sibling, 0, N - 1)
    core_prev = core.copy()
    rows = []
    it = 0
    for it in range(0, maxiter + 2):
        if it > maxiter:
            break
        if (it % resprint == 0 and it != 0) or it == 1 or it == maxiter:
            core = TTMc(V, W, -1)
            diffnorm = abs(np.linalg.norm(core.ravel()) - np.linalg.norm(core_prev.ravel()))
            diffV = tucker_residual(V, core, W) if want_residual else float("nan")
            rows.append((it, diffnorm, diffV))
            if diffnorm < tol:
                break
            core_prev = core.copy()
        ttmc_map = {}
        for i in range(N):
            Y = _tucker_leaf_Y(ttmc_map, parent, sibling, V, W, i)
            if i == N - 1:
                Y_end = Y
            W[i] = top_left_singular(unroll_tensor_contraction(Y, i), core.shape[i])
        core = ttm(Y_end, N - 1, W[N - 1])  # :408
    return (it != maxiter + 1), rows, core


def Build_ttmc_map(ttmc_map, V, W, args):
    """als_Tucker.cxx:426-466."""
    if len(args) == 1:
        M = V
    else:
        a2 = args[:-1]
        if a2 not in ttmc_map:
            Build_ttmc_map(ttmc_map, V, W, a2)
        M = ttmc_map[a2]
    x = ord(args[-1]) - 97
    ttmc_map[args] = ttm(M, x, W[x])


def build_tucker_pp_operators(V, W):
    """als_Tucker.cxx:742-760."""
    N = V.ndim
    seq = letters(N)
    m = {}
    for ii in range(N):
        for jj in range(ii + 1, N):
            Build_ttmc_map(m, V, W, "".join(c for k, c in enumerate(seq) if k not in (ii, jj)))
    for ii in range(N):
        Build_ttmc_map(m, V, W, "".join(c for k, c in enumerate(seq) if k != ii))
    return m


def alsTucker_DT_sub(V, cs, W, dW, st, tol, tol_init, maxiter, resprint, rows, sweeps, want_residual=True):
    """als_Tucker.cxx:476-669.  cs.core / cs.core_prev are the by-reference tensors."""
    N = V.ndim
    W_prev = [np.zeros_like(w) for w in W]
    parent, sibling = {}, {}
    construct_dimension_tree(parent, sibling, 0, N - 1)
    while st.iter <= maxiter:
        it = st.iter
        if (it % resprint == 0 and it != 0) or it == 1 or it == maxiter:
            cs.core = TTMc(V, W, -1)
            st.diffnorm = abs(np.linalg.norm(cs.core.ravel()) - np.linalg.norm(cs.core_prev.ravel()))
            diffV = tucker_residual(V, cs.core, W) if want_residual else float("nan")
            rows.append((it, st.diffnorm, 0, diffV))
            if st.diffnorm < tol:
                break
            cs.core_prev = cs.core.copy()
        ttmc_map = {}
        for i in range(N):
            Y = _tucker_leaf_Y(ttmc_map, parent, sibling, V, W, i)
            if i == N - 1:
                Y_end = Y
            U = top_left_singular(unroll_tensor_contraction(Y, i), cs.core.shape[i])
            W[i] = sign_align(U, W_prev[i])
        cs.core = ttm(Y_end, N - 1, W[N - 1])
        sweeps.append(("DT", it))
        num_dw_break = 0
        for i in range(N):
            dW[i] = W[i] - W_prev[i]
            W_prev[i] = W[i].copy()
            if abs(np.linalg.norm(dW[i]) / np.linalg.norm(W[i])) < tol_init:
                num_dw_break += 1
        if num_dw_break == N:
            return
        st.iter += 1


def tucker_pp_corrected_Y(ttmc_map, dW, i, N):
    """als_Tucker.cxx:828-860."""
    seq = letters(N)
    Y = ttmc_map["".join(c for k, c in enumerate(seq) if k != i)].copy()
    for ii in range(N):
        if ii == i:
            continue
        key = "".join(c for k, c in enumerate(seq) if k not in (ii, i))
        Y = Y + ttm(ttmc_map[key], ii, dW[ii])
    return Y


def alsTucker_PP_sub(V, cs, W, dW, st, tol, tol_init, maxiter, resprint, bench, rows, sweeps, want_residual=True):
    """als_Tucker.cxx:679-896."""
    N = V.ndim
    init_iter = st.iter
    W_init = [None] * N
    ttmc_map = {}
    while st.iter <= maxiter:
        it = st.iter
        num_dw_break = 0
        if not bench:
            for i in range(N):
                if abs(np.linalg.norm(dW[i]) / np.linalg.norm(W[i])) > tol_init:
                    num_dw_break += 1
        if it == init_iter or num_dw_break > 0:
            if num_dw_break > 0:
                return
            for j in range(N):
                W_init[j] = W[j].copy()
                dW[j] = np.zeros_like(W[j])
            ttmc_map = build_tucker_pp_operators(V, W)
            sweeps.append(("PPinit", it))
        if (it % resprint == 0 and it != 0) or it == 1 or it == maxiter or it == init_iter:
            cs.core = TTMc(V, W, -1)
            st.diffnorm = abs(np.linalg.norm(cs.core.ravel()) - np.linalg.norm(cs.core_prev.ravel()))
            diffV = tucker_residual(V, cs.core, W) if want_residual else float("nan")
            rows.append((it, st.diffnorm, 1, diffV))
            if st.diffnorm < tol or it == maxiter:
                break
            cs.core_prev = cs.core.copy()
        for i in range(N):
            Y = tucker_pp_corrected_Y(ttmc_map, dW, i, N)
            if i == N - 1:
                Y_end = Y
            U = top_left_singular(unroll_tensor_contraction(Y, i), cs.core.shape[i])
            W[i] = sign_align(U, W_init[i])
            dW[i] = W[i] - W_init[i]
        cs.core = ttm(Y_end, N - 1, W[N - 1])
        sweeps.append(("PP", it))
        st.iter += 1
    if bench:
        st.iter += 1


def alsTucker_PP(V, core, W, tol, tol_init, maxiter, resprint=10, bench=False, want_residual=True):
    """als_Tucker.cxx:906-962.  Returns (stopped_early, rows, events, sweeps, core)."""
    st = _State()
    st.iter = 0
    st.diffnorm = 10.0
    cs = _State()
    cs.core = core
    cs.core_prev = core.copy()
    dW = [np.zeros_like(w) for w in W]
    rows, events, sweeps = [], [], []
    while st.diffnorm > tol and st.iter <= maxiter:
        if not bench:
            events.append(("DT", st.iter))
            alsTucker_DT_sub(V, cs, W, dW, st, tol, tol_init, maxiter, resprint, rows, sweeps, want_residual)
        events.append(("PP", st.iter))
        alsTucker_PP_sub(V, cs, W, dW, st, tol, tol_init, maxiter, resprint, bench, rows, sweeps, want_residual)
        if tol_init > 5e-3:
            tol_init *= 0.9
    return (st.iter != maxiter + 1), rows, events, sweeps, cs.core


# ----------------------------------------------------------------------------------------------------------------
# synthetic inputs (test_ALS.cxx:262-286, 332-345)
# ----------------------------------------------------------------------------------------------------------------
def make_tensor_r(lens, R, seed=1):
    """tensor 'r': V = [[W_true]] with W_true[i] = u(seed, id=i) in [0,1)."""
    Wt = [fill_uniform((lens[i], R), seed, i) for i in range(len(lens))]
    return np.asfortranarray(build_V(Wt)), Wt


def make_tensor_r2(lens, seed=1, lo=0.5, hi=1.0):
    """tensor 'r2' (test_ALS.cxx:272): uniform in [0.5,1)."""
    return fill_uniform(tuple(lens), seed, 100, lo, hi)


def identity_tensor(N, s):
    """common.cxx:462-498  I[a,b,c,d,..] = delta(a,b) delta(c,d) ... (order N even), built pair by pair."""
    ident = np.eye(s)
    I = ident
    for _ in range(1, N // 2):
        I = np.multiply.outer(ident, I)  # I_temp[ab c..] = ident[ab] * I_temp2[c..]
    return I


def laplacian_tensor(N, s):
    """common.cxx:575-642  V = D x I x .. x I + I x D x I .. + ... over the d = N/2 index pairs, D = tridiag(-1,2,-1)."""
    d = N // 2
    D = 2.0 * np.eye(s) - np.eye(s, k=1) - np.eye(s, k=-1)
    V = np.zeros((s,) * N)
    letters_ = "abcdefghijklmnop"[:N]
    for k in range(d):
        ops, subs = [], []
        for m in range(d):
            ops.append(D if m == k else np.eye(s))
            subs.append(letters_[2 * m:2 * m + 2])
        V += np.einsum(",".join(subs) + "->" + letters_, *ops)
    return np.asfortranarray(V)


def make_tensor_p(dim, s, folded=True):
    """tensor 'p2' (order dim) / 'p' (the same entries as dim/2 modes of size s*s; fold_unfold, common.cxx:870-882)."""
    V = laplacian_tensor(dim, s)
    if not folded:
        return V
    return np.asfortranarray(V.reshape((s * s,) * (dim // 2), order="F"))


def collinearity(v1, v2):
    """common.cxx:297-302"""
    ip = n1 = n2 = 0.0
    for a, b in zip(v1.tolist(), v2.tolist()):  # same summation order as the host loop
        ip += a * b
        n1 += a * a
        n2 += b * b
    return ip / (np.sqrt(n1) * np.sqrt(n2))


def gen_collinearity(lens, R, col_min, col_max, seed=1):
    """common.cxx:361-423.  Draw k of the run is u(seed, id = 1000 + k, .); vectors are redrawn until their
    collinearity with every earlier vector of the mode lies in [col_min, col_max]; lambda_i = 0.2 + 0.6 (i+1)/R."""
    dim = len(lens)
    draw = [1000]

    def fill(n):
        v = u01(seed, draw[0], n)
        draw[0] += 1
        return v

    vec = [[fill(lens[j]) for j in range(dim)] for _ in range(R)]
    for j in range(dim):
        for i in range(1, R):
            while True:
                ok = True
                for k in range(i):
                    col = collinearity(vec[i][j], vec[k][j])
                    if col < col_min or col > col_max:
                        ok = False
                        break
                if ok:
                    break
                vec[i][j] = fill(lens[j])
    W = []
    for j in range(dim):
        Wj = np.zeros((lens[j], R))
        for i in range(R):
            lam = 0.2 + 0.6 / R * (i + 1) if j == 0 else 1.0
            Wj[:, i] = lam * vec[i][j]
        W.append(Wj)
    return np.asfortranarray(build_V(W)), vec


def make_tensor_c(lens, R, col_min=0.5, col_max=0.9, ratio_noise=0.01, seed=1):
    """tensor 'c' (test_ALS.cxx:245-261): collinearity-constrained rank-R tensor + uniform(-1,1) noise of norm
    ratio_noise * ||V||."""
    V, _ = gen_collinearity(lens, R, col_min, col_max, seed)
    noise = fill_uniform(tuple(lens), seed, 101, -1.0, 1.0)
    return np.asfortranarray(V + ratio_noise * np.linalg.norm(V) / np.linalg.norm(noise) * noise)


def init_factors(lens, R, seed=2):
    """W0[i] = u(seed, id=i) in [0,1) (test_ALS.cxx:337)."""
    return [fill_uniform((lens[i], R), seed, i) for i in range(len(lens))]


def init_grad(lens, R, seed=3):
    """grad_W[i] = u(seed, id=i) in [0,1) (test_ALS.cxx:338: the reference fills grad_W with fill_random(0,1); it
    must be non-zero or the iteration-0 check `gradnorm < tol` (als_CP.cxx:211) stops the run immediately)."""
    return [fill_uniform((lens[i], R), seed, i) for i in range(len(lens))]
