"""Design prototype (NumPy, not product code) for the next step on DESIGN.md section 6b, gap 1: a BLOCKED inverse of
the R x R Gram-Hadamard matrix S, to replace the 50 dependent elimination steps of spd_inverse_ldl_kernel (37 us at
R = 50, the critical path of the PP approximate sweep) by R/8 block steps.

Block LDL^T with an 8 x 8 block-diagonal D (SPD blocks, NOT factored further) and a unit block-lower L:

    for k in blocks:                                     # R/8 sequential steps (7 at R = 50, padded to 56)
        Dinv_k = inverse(S_kk)                           # 8 x 8 SPD: Gauss-Jordan inside ONE warp, 8 shuffle rounds
        P_ik   = S_ik @ Dinv_k            (i > k)        # panel: every row block independently (DMMA 8x8x4 tiles)
        S_ij  -= P_ik @ S_jk^T            (i >= j > k)   # trailing update, independent tiles, one CTA barrier
        L_ik   = P_ik
    Y = L^-1 by block forward substitution (unit diagonal: Y_ik = -sum_{j<i, j>=k} L_ij Y_jk, done per block row)
    S^-1 = Y^T Dinv Y                                    # one small product, all tiles independent

Dependent depth per block step: 8 shuffle rounds (~30 cycles each) + panel product + barrier + update + barrier
~ 1.5 k cycles, against 16 x 1080 cycles for the same 16 columns today; the forward substitution can be folded into
the same loop the way the current kernel accumulates S^-1 inside its column loop.

Run: python tools/inv_blocked_proto.py   (checks the scheme against numpy.linalg.inv on Gram-Hadamard matrices)."""
import numpy as np

B = 8


def blocked_ldl_inverse(S):
    R = S.shape[0]
    Rp = ((R + B - 1) // B) * B
    A = np.eye(Rp)
    A[:R, :R] = S  # identity padding keeps the factorisation of the padded matrix trivial
    nb = Rp // B
    L = np.eye(Rp)
    Dinv = np.zeros((Rp, Rp))
    steps = 0
    for k in range(nb):
        ks = slice(k * B, (k + 1) * B)
        Dinv[ks, ks] = np.linalg.inv(A[ks, ks])  # one warp, 8 Gauss-Jordan rounds
        steps += 1
        for i in range(k + 1, nb):
            i_s = slice(i * B, (i + 1) * B)
            L[i_s, ks] = A[i_s, ks] @ Dinv[ks, ks]  # panel (independent row blocks)
        for i in range(k + 1, nb):
            i_s = slice(i * B, (i + 1) * B)
            for j in range(k + 1, i + 1):
                j_s = slice(j * B, (j + 1) * B)
                A[i_s, j_s] -= L[i_s, ks] @ A[j_s, ks].T  # trailing update (independent tiles)
                if i != j:
                    A[j_s, i_s] = A[i_s, j_s].T
    # Y = L^-1, block forward substitution on the identity
    Y = np.eye(Rp)
    for i in range(nb):
        i_s = slice(i * B, (i + 1) * B)
        for k in range(i):
            ks = slice(k * B, (k + 1) * B)
            acc = L[i_s, ks].copy()
            for j in range(k + 1, i):
                j_s = slice(j * B, (j + 1) * B)
                acc += L[i_s, j_s] @ Y[j_s, ks]
            Y[i_s, ks] = -acc
    Sinv = Y.T @ Dinv @ Y
    return Sinv[:R, :R], steps


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for R, s, N in [(50, 300, 4), (10, 40, 6), (64, 200, 4), (100, 300, 3), (7, 30, 4)]:
        W = [rng.random((s, R)) for _ in range(N)]
        S = np.ones((R, R))
        for w in W[1:]:
            S *= w.T @ w
        ref = np.linalg.inv(S)
        got, steps = blocked_ldl_inverse(S)
        err = np.abs(got - ref).max() / np.abs(ref).max()
        res = np.abs(got @ S - np.eye(R)).max()
        print("R=%3d cond=%.2e block steps=%2d  rel.err vs inv %.2e  |S^-1 S - I| %.2e" % (R, np.linalg.cond(S), steps, err, res))
        assert res < 1e-6 * max(1.0, np.linalg.cond(S) * 1e-10)
