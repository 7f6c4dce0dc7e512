"""Times the R x R SPD inverse (ppx_spd_inverse_g: Hadamard of three Grams + the register-resident sweep kernel for
R <= 64, the column LDL^T kernel above) alone, CUDA events around 200 back-to-back launches.
PPX_INV_COLUMN=1: the column kernel for every R, for A/B."""
import ctypes as C
import importlib
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ppx = importlib.import_module("pairwise-perturbation_b200")
ctx = ppx.Ctx(0)
lib = ppx.load_library()
out = []
for R in (10, 40, 50, 64, 100):
    s = 300
    rng = np.random.default_rng(R)
    Gs = []
    for j in range(4):
        w = rng.random((s, R))
        G = ctx.empty(R * R)
        ctx.gram(ctx.to_device(w), s, R, G)
        Gs.append(G)
    S, Si = ctx.empty(R * R), ctx.empty(R * R)
    for _ in range(10):
        ctx.spd_inverse_g(Gs, 1, 0.0, R, 0, S, Si)
    e0, e1 = C.c_void_p(), C.c_void_p()
    lib.ppx_event_create(ctx.h, C.byref(e0))
    lib.ppx_event_create(ctx.h, C.byref(e1))
    lib.ppx_event_record(ctx.h, e0)
    for _ in range(200):
        ctx.spd_inverse_g(Gs, 1, 0.0, R, 0, S, Si)
    lib.ppx_event_record(ctx.h, e1)
    ms = C.c_float(0)
    lib.ppx_event_elapsed_ms(ctx.h, e0, e1, C.byref(ms))
    out.append({"R": R, "us_per_inverse": 1e3 * ms.value / 200})
print(json.dumps(out))
