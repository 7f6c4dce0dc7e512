"""NumPy model of the register-resident symmetric sweep inverse (csrc/k45_solve.cu, spd_inverse_sweep_kernel):
every "thread" j holds column j; step k reads the PUBLISHED column k only (row k is never read, it is overwritten
from the column by symmetry); after R steps the matrix is -S^-1.   python tools/inv_sweep_proto.py"""
import numpy as np


def sweep_inverse(S):
    R = S.shape[0]
    C = S.copy()  # C[:, j] = registers of thread j
    for k in range(R):
        col = C[:, k].copy()  # published
        rp = 1.0 / col[k]
        for j in range(R):
            if j == k:
                u = -rp
                C[:, j] = 0.0
            else:
                u = col[j] * rp
            keep = C[k, j]
            C[:, j] -= col * u
            C[k, j] = u
            del keep
    X = -C
    return 0.5 * (X + X.T)


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for R in (1, 2, 7, 10, 50, 64):
        for cond_pow in (0, 3):
            G = np.ones((R, R))
            for _ in range(3):
                w = rng.random((300, R)) ** (1 + cond_pow)
                G *= w.T @ w
            X = sweep_inverse(G)
            ref = np.linalg.inv(G)
            print(R, cond_pow, "cond %.2e" % np.linalg.cond(G), "rel err %.2e" % (np.abs(X - ref).max() / np.abs(ref).max()),
                  "resid %.2e" % np.abs(X @ G - np.eye(R)).max())
