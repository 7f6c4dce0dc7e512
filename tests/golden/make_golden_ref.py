"""Generates tests/golden/ref/*.npz by running THE REFERENCE'S OWN CODE (oracle/_ref/ref_driver: the unmodified
als_CP.cxx / als_Tucker.cxx / common.cxx of /root/reference compiled against oracle/ctf_standin/ctf.hpp) on seeded
inputs.  The fixtures travel; /root/reference and oracle/_ref are only needed to regenerate them:

    make -C oracle ref && python tests/golden/make_golden_ref.py

Inputs are not stored: they are regenerated from the seeds with the counter-based generator (pp_oracle.make_tensor_r /
make_tensor_r2 / init_factors / init_grad), except the Tucker starting factors (HOSVD of V, stored as W_init so that
every implementation starts from the same column signs).  Stored: what the reference printed at each print point
(iteration, gradient norm or core-norm difference, pp_update flag, residual) at 17 digits, its DT<->PP switching
markers, and its final factors."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import pp_oracle as o  # noqa: E402
from oracle import ref_harness as rh  # noqa: E402

CASES = {
    "cp_dt_n4": dict(op="alsCP_DT", lens=(12, 13, 14, 15), R=4, maxiter=20, resprint=5),
    "cp_dt_n6_lambda": dict(op="alsCP_DT", lens=(6, 7, 6, 5, 6, 7), R=3, maxiter=12, resprint=4, lambda_=1e-3),
    "cp_pp_n4_s13_r5": dict(op="alsCP_PP", lens=(13, 13, 13, 13), R=5, tol_init=0.1, maxiter=40, resprint=5),
    "cp_pp_n6_s8_r3": dict(op="alsCP_PP", lens=(8,) * 6, R=3, tol_init=0.1, maxiter=30, resprint=5),
    "cp_pp_n4_ragged_r4": dict(op="alsCP_PP", lens=(9, 14, 5, 11), R=4, tol_init=0.05, maxiter=40, resprint=4),
    "cp_pp_n4_ratio_lambda": dict(op="alsCP_PP", lens=(10, 9, 8, 11), R=3, tol_init=0.2, maxiter=40, resprint=5,
                                  ratio_step=0.8, lambda_=1e-4),
    "cp_pp_n7": dict(op="alsCP_PP", lens=(5, 4, 5, 4, 5, 4, 5), R=2, tol_init=0.1, maxiter=20, resprint=5),
    "cp_part_n4_full": dict(op="alsCP_PP_partupdate", lens=(12, 13, 14, 15), R=4, tol_init=0.1, maxiter=30,
                            resprint=5, update_pct=1.0),
    "cp_part_n4_half": dict(op="alsCP_PP_partupdate", lens=(12, 13, 14, 15), R=4, tol_init=0.1, maxiter=30,
                            resprint=5, update_pct=0.5),
    # sizes at which the streaming first contraction (X <= 64, R <= 16), the one-thread-per-output children and the TMA
    # tiles are the kernels that run on the CUDA side
    "cp_pp_n4_s40_r10": dict(op="alsCP_PP", lens=(40, 40, 40, 40), R=10, tol_init=0.1, maxiter=30, resprint=5),
    "cp_pp_n6_s12_r4": dict(op="alsCP_PP", lens=(12,) * 6, R=4, tol_init=0.1, maxiter=20, resprint=5),
    "cp_dt_n4_ragged_r12": dict(op="alsCP_DT", lens=(48, 20, 36, 30), R=12, maxiter=10, resprint=5),
    # HOSVD with one long mode: on the CUDA side its Gram is solved by the subspace iteration (n >= 384)
    "hosvd_n3_long_mode": dict(op="hosvd", lens=(400, 12, 10), R=4, ranks=(8, 4, 4)),
    "tucker_dt_n4": dict(op="alsTucker_DT", lens=(9, 10, 8, 7), R=3, maxiter=12, resprint=4),
    "tucker_pp_n4": dict(op="alsTucker_PP", lens=(9, 10, 8, 7), R=3, tol_init=0.3, maxiter=30, resprint=5),
}


def main():
    assert rh.available(), "build oracle/_ref first: make -C oracle ref"
    os.makedirs(os.path.join(HERE, "ref"), exist_ok=True)
    for name, c in CASES.items():
        c = dict(c)
        op, lens, R = c.pop("op"), c.pop("lens"), c.pop("R")
        N = len(lens)
        out = dict(op=op, lens=np.array(lens), R=R, source="reference (oracle/_ref/ref_driver)")
        for k, v in c.items():
            out[k] = v
        if op == "hosvd":
            ranks = c.pop("ranks")
            V = o.make_tensor_r2(lens)
            vnorm = np.linalg.norm(V)
            ref = rh.run_driver(op, V, ranks=list(ranks))
            out["ranks"] = np.array(ranks)
            out["core_norm"] = float(np.linalg.norm(ref["core"]))
            ref["rows"], ref["events"] = [(0.0, 0.0, 0, 0.0)], []
        elif op.startswith("alsTucker"):
            V = o.make_tensor_r2(lens)
            vnorm = np.linalg.norm(V)
            _, W0 = o.hosvd(V, [R] * N)
            ref = rh.run_driver(op, V, W0, ranks=[R] * N, tol=1e-10 * vnorm, **c)
            for i, w in enumerate(W0):
                out["W_init%d" % i] = w
            out["core_norm"] = float(np.linalg.norm(ref["core"]))
        else:
            V, _ = o.make_tensor_r(lens, R)
            vnorm = np.linalg.norm(V)
            ref = rh.run_driver(op, V, o.init_factors(lens, R), o.init_grad(lens, R), tol=1e-10 * vnorm, **c)
            for i, g in enumerate(ref["grad"]):
                out["grad%d" % i] = g
        out["vnorm"] = vnorm
        out["rows"] = np.array(ref["rows"], dtype=float)
        out["events"] = np.array([(0 if k == "DT" else 1, it) for k, it in ref["events"]], dtype=np.int64).reshape(-1, 2)
        for i, w in enumerate(ref["W"]):
            out["W%d" % i] = w
        np.savez_compressed(os.path.join(HERE, "ref", name + ".npz"), **out)
        print(name, "rows", len(ref["rows"]), "events", ref["events"], "final residual %.3e" % ref["rows"][-1][3])


if __name__ == "__main__":
    main()
