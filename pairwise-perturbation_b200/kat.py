"""Closed-form known answers for a tensor that is exactly a CP model, V = [[A_0 .. A_{N-1}]] (the synthetic tensor 'r' of
test_ALS.cxx:275-286).  Everything the dimension tree and the PP operator build compute from such a V has a closed
form in the small matrices A_j^T W_j, so the CUDA contractions can be checked at ANY size -- including BASELINE
configs[1] (s=300, R=50, 8.1e9 elements, beyond 2^31) -- at O(s R^2) host cost, on one GPU and on every shard:

    MTTKRP_i(W)[x, r]        = sum_k A_i[x, k] * prod_{m != i}      (A_m^T W_m)[k, r]            (als_CP.cxx:236-284)
    P^(i,j)(W)[x_i, x_j, r]  = sum_k A_i[x_i, k] A_j[x_j, k] * prod_{m != i, j} (A_m^T W_m)[k, r]   (als_CP.cxx:352-409)

Host-side NumPy on s x R matrices only; used by tests/ and by bench.py's parity_probe as the checker.  Not a compute
path: nothing here touches the tensor.
"""
import numpy as np


def cross_grams(A, W):
    """C[m] = A_m^T W_m  (R_true x R) from the GLOBAL factors."""
    return [a.T @ w for a, w in zip(A, W)]


def mttkrp(A, C, i, rows=None):
    """MTTKRP of mode i; rows = (begin, end) restricts to the local rows of a sharded mode i."""
    H = np.ones_like(C[0])
    for m in range(len(A)):
        if m != i:
            H = H * C[m]
    Ai = A[i] if rows is None else A[i][rows[0]:rows[1]]
    return Ai @ H


def pair_operator(A, C, i, j, rows_i=None):
    """P^(i,j), i < j: s_i x s_j x R (rank last), the value of mttkrp_map[all modes but i, j] (als_CP.cxx:385-408)."""
    assert i < j
    H = np.ones_like(C[0])
    for m in range(len(A)):
        if m != i and m != j:
            H = H * C[m]
    Ai = A[i] if rows_i is None else A[i][rows_i[0]:rows_i[1]]
    return np.einsum("ak,bk,kr->abr", Ai, A[j], H, optimize=True)


def max_rel_err(got, ref):
    scale = float(np.abs(ref).max())
    return float(np.abs(got - ref).max() / (scale if scale > 0 else 1.0))
