"""First-contraction (K1) and Hadamard-batched (K2) kernels at the shapes of the PP operator build of BASELINE
configs[3] (order 6, s = 40, R = 10) and configs[4] (3 x 128 x 128 x 7200, R = 10): achieved HBM GB/s against the
measured copy bandwidth.  CUDA events on the context stream, after warm-up; algorithmic bytes as in SURVEY.md 8(d).

Usage: python tools/bench_shapes.py [--cfg 4|5|all]      A/B: PPX_NO_STREAM=1 / PPX_NO_FLAT=1 select the tile kernels
"""
import argparse
import importlib
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ppx = importlib.import_module("pairwise-perturbation_b200")
import torch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cfg", default="all")
ap.add_argument("--reps", type=int, default=5)
args = ap.parse_args()
PEAK = 6543.1
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
ctx = ppx.Ctx(0, workspace_bytes=1 << 30)


def timeit(fn, reps=args.reps, warm=2):
    for _ in range(warm):
        fn()
    ctx.sync()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(ctx.stream):
            e0.record()
            fn()
            e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


def prod(v):
    p = 1
    for x in v:
        p *= x
    return p


res = []


def k1(name, lens, x, R):
    P = prod(lens)
    V = ctx.empty(P)
    ctx.fill_uniform(V, 1, 0)
    W = ctx.empty(lens[x] * R)
    ctx.fill_uniform(W, 2, 0)
    out = ctx.empty(P // lens[x] * R)
    ms = timeit(lambda: ctx.ttm_first(V, list(lens), x, W, R, out))
    by = 8.0 * (P + P // lens[x] * R + lens[x] * R)
    res.append(dict(kernel="k1 " + name, lens=lens, x=x, R=R, ms=ms, gbs=by / ms / 1e6, frac=by / ms / 1e6 / PEAK))
    print(json.dumps(res[-1]), flush=True)
    del V, out


def k2(name, lens, x, R):
    P = prod(lens) * R
    T = ctx.empty(P)
    ctx.fill_uniform(T, 3, 0)
    W = ctx.empty(lens[x] * R)
    ctx.fill_uniform(W, 2, 0)
    out = ctx.empty(P // lens[x])
    ms = timeit(lambda: ctx.mttv(T, list(lens), x, W, R, out))
    by = 8.0 * (P + P // lens[x] + lens[x] * R)
    res.append(dict(kernel="k2 " + name, lens=lens, x=x, R=R, ms=ms, gbs=by / ms / 1e6, frac=by / ms / 1e6 / PEAK))
    print(json.dumps(res[-1]), flush=True)
    del T, out


if args.cfg in ("4", "all"):
    for x in (2, 1, 0):
        k1("cfg4 level-1", (40,) * 6, x, 10)
    for x in (0, 1, 2, 4):
        k2("cfg4 level-2", (40,) * 5, x, 10)
    for x in (0, 1, 3):
        k2("cfg4 level-3", (40,) * 4, x, 10)
if args.cfg in ("5", "all"):
    for x in (2, 1, 0):
        k1("cfg5 level-1", (3, 128, 128, 7200), x, 10)
    for x in (0, 1, 2):
        k2("cfg5 level-2 of T_a", (128, 128, 7200), x, 10)
    k2("cfg5 leaf (3 x 7200)", (3, 7200), 1, 10)
    k2("cfg5 leaf (128 x 7200)", (128, 7200), 1, 10)
    k2("cfg5 leaf (128 x 7200) x=0", (128, 7200), 0, 10)
    # K3 at the coil shape: the PP correction of every mode
    lens, R = (3, 128, 128, 7200), 10
    for i in range(4):
        ops, which, dws, so = [], [], [], []
        for j in range(4):
            if j == i:
                continue
            t = ctx.empty(lens[i] * lens[j] * R)
            ctx.fill_uniform(t, 4, j)
            ops.append(t)
            which.append(0 if j < i else 1)
            d = ctx.empty(lens[j] * R)
            ctx.fill_uniform(d, 5, j)
            dws.append(d)
            so.append(lens[j])
        M0, Mo = ctx.empty(lens[i] * R), ctx.empty(lens[i] * R)
        ms = timeit(lambda: ctx.pp_correct(M0, ops, which, dws, so, lens[i], R, Mo), reps=10)
        by = 8.0 * (sum(lens[i] * s * R + s * R for s in so) + 2 * lens[i] * R)
        res.append(dict(kernel="k3 cfg5 mode %d" % i, ms=ms, gbs=by / ms / 1e6, frac=by / ms / 1e6 / PEAK))
        print(json.dumps(res[-1]), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
tag = ("_nostream" if os.environ.get("PPX_NO_STREAM") else "") + ("_noflat" if os.environ.get("PPX_NO_FLAT") else "")
json.dump(res, open("gpurun_out/shapes%s.json" % tag, "w"), indent=1)
