"""Times the residual evaluation (K7, ppx_cp_residual) at BASELINE configs[1] (order 4, s = 300, R = 50) and at the
small-rank shapes: wall clock around H.cp_residual (one stream synchronise inside), best of 5.
    python tools/time_k7.py            (PPX_K7_DFMA=1: the pre-round-2 DFMA kernel, for A/B)"""
import importlib
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
H = importlib.import_module("pairwise-perturbation_b200.host_api")
world = H.World(0, workspace_bytes=1 << 30)
out = []
for lens, R in [((300,) * 4, 50), ((40,) * 6, 10), ((3, 128, 128, 7200), 10), ((200,) * 3, 10)]:
    N = len(lens)
    At = [H.Matrix(world, lens[i], R) for i in range(N)]
    W = [H.Matrix(world, lens[i], R) for i in range(N)]
    for i in range(N):
        At[i].fill(1, i)
        W[i].fill(2, i)
    V = H.Tensor(world, lens)
    t0 = time.perf_counter()
    H.build_V(world, V, At)
    world.sync()
    t_build = time.perf_counter() - t0
    H.cp_residual(world, V, W)
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        r = H.cp_residual(world, V, W)
        ts.append(time.perf_counter() - t0)
    exact = H.cp_residual(world, V, At)
    P = 1
    for x in lens:
        P *= x
    out.append({"lens": lens, "R": R, "residual_ms": 1e3 * min(ts), "tflops": 2.0 * P * R / min(ts) / 1e12,
                "gbs": 8.0 * P / min(ts) / 1e9, "build_V_ms_first_call": 1e3 * t_build, "residual": r,
                "residual_at_truth_over_norm": exact / V.norm2()})
    for t in [V] + At + W:
        t.free()
    world.trim()
print(json.dumps(out))
world.close()
