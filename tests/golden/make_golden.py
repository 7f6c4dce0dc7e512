"""Generates the golden fixtures in this directory from oracle/pp_oracle.py (the reference cannot be run here, so
these pin the ORACLE: seeded inputs -> per-print-point gradient norm / residual, DT<->PP switching iterations, final
factors).  Run from the repo root:  python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pp_oracle as o  # noqa: E402

CASES = {
    "cp_pp_n4_s13_r5": dict(lens=(13, 13, 13, 13), R=5, tol_init=0.1, maxiter=40, resprint=5),
    "cp_pp_n3_s20_r3": dict(lens=(20, 20, 20), R=3, tol_init=0.1, maxiter=40, resprint=5),
    "cp_pp_n6_s8_r3": dict(lens=(8, 8, 8, 8, 8, 8), R=3, tol_init=0.1, maxiter=30, resprint=5),
    "cp_pp_n4_ragged_r4": dict(lens=(9, 14, 5, 11), R=4, tol_init=0.05, maxiter=40, resprint=4),
}

for name, c in CASES.items():
    V, _ = o.make_tensor_r(c["lens"], c["R"])
    W, G = o.init_factors(c["lens"], c["R"]), o.init_grad(c["lens"], c["R"])
    vnorm = np.linalg.norm(V)
    _, tr = o.alsCP_PP(V, W, G, 1e-10 * vnorm, c["tol_init"], c["maxiter"], resprint=c["resprint"])
    out = dict(lens=np.array(c["lens"]), R=c["R"], tol_init=c["tol_init"], maxiter=c["maxiter"], resprint=c["resprint"],
               vnorm=vnorm, events=np.array([(0 if k == "DT" else 1, it) for k, it in tr.events]),
               rows=np.array([(r[0], r[1], r[2], r[3]) for r in tr.rows]))
    for i, w in enumerate(W):
        out["W%d" % i] = w
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), name + ".npz"), **out)
    print(name, "events", tr.events, "final residual %.3e" % tr.rows[-1][3])
