"""Times ppx_mttv3 (the Hadamard contractions of a level-1 tensor in one pass) against one ppx_mttv per output at the
PP-build shape of BASELINE configs[1]: T = 300 x 300 x 300 x 50 doubles (10.8 GB).  CUDA events, best of 5.
    python tools/time_mttv3.py [s] [R]"""
import ctypes as C
import importlib
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ppx = importlib.import_module("pairwise-perturbation_b200")
s = int(sys.argv[1]) if len(sys.argv) > 1 else 300
R = int(sys.argv[2]) if len(sys.argv) > 2 else 50
ctx = ppx.Ctx(0, workspace_bytes=1 << 30)
lib = ppx.load_library()
lens = (s, s, s)
T = ctx.empty(s * s * s * R)
ctx.fill_uniform(T, 1, 0)
Ws = []
for i in range(3):
    w = ctx.empty(s * R)
    ctx.fill_uniform(w, 2, i)
    Ws.append(w)
outs = [ctx.empty(s * s * R) for _ in range(3)]
refs = [ctx.empty(s * s * R) for _ in range(3)]
e0, e1 = C.c_void_p(), C.c_void_p()
lib.ppx_event_create(ctx.h, C.byref(e0))
lib.ppx_event_create(ctx.h, C.byref(e1))


def timed(fn, reps=5):
    fn()
    best = 1e9
    for _ in range(reps):
        lib.ppx_event_record(ctx.h, e0)
        fn()
        lib.ppx_event_record(ctx.h, e1)
        ms = C.c_float(0)
        lib.ppx_event_elapsed_ms(ctx.h, e0, e1, C.byref(ms))
        best = min(best, ms.value)
    return best


res = {"shape": [s, s, s, R], "bytes": 8 * s * s * s * R}
res["separate_ms"] = [timed(lambda i=i: ctx.mttv(T, lens, i, Ws[i], R, refs[i])) for i in range(3)]
res["fused3_ms"] = timed(lambda: ctx.mttv3(T, lens, Ws[0], Ws[1], Ws[2], R, outs[0], outs[1], outs[2]))
res["fused_lx_ms"] = timed(lambda: ctx.mttv3(T, lens, Ws[0], Ws[1], None, R, outs[0], outs[1], None))
res["fused_xt_ms"] = timed(lambda: ctx.mttv3(T, lens, None, Ws[1], Ws[2], R, None, outs[1], outs[2]))
res["fused_lt_ms"] = timed(lambda: ctx.mttv3(T, lens, Ws[0], None, Ws[2], R, outs[0], None, outs[2]))
ctx.mttv3(T, lens, Ws[0], Ws[1], Ws[2], R, outs[0], outs[1], outs[2])
import numpy as np
res["max_rel_diff_vs_separate"] = [float(np.abs(ctx.to_host(outs[i], (s * s * R,)) - ctx.to_host(refs[i], (s * s * R,))).max()
                                         / np.abs(ctx.to_host(refs[i], (s * s * R,))).max()) for i in range(3)]
res["gbs_fused3"] = res["bytes"] / res["fused3_ms"] / 1e6
print(json.dumps(res))
