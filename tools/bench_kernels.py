"""Per-kernel timings at the BASELINE configurations (CUDA events on the context stream, after warm-up).

Usage: python tools/bench_kernels.py [--s 300] [--R 50] [--N 4] [--only k1]   (writes gpurun_out/kernels.json)
Algorithmic flops/bytes follow SURVEY.md 8(d): first contraction flops = 2*P*R, bytes = 8*(P + P*R/s_x + s_x*R).
"""
import argparse
import importlib
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ppx = importlib.import_module("pairwise-perturbation_b200")

ap = argparse.ArgumentParser()
ap.add_argument("--s", type=int, default=300)
ap.add_argument("--R", type=int, default=50)
ap.add_argument("--N", type=int, default=4)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--only", default="")
args = ap.parse_args()

import torch

ctx = ppx.Ctx(0, workspace_bytes=1 << 30)
s, R, N = args.s, args.R, args.N
lens = [s] * N
P = s**N
PEAK_TF, PEAK_GBS = 35.46, 6543.1
V = ctx.empty(P)
ctx.fill_uniform(V, 1, 0)
W = []
for i in range(N):
    w = ctx.empty(s * R)
    ctx.fill_uniform(w, 2, i)
    W.append(w)
ctx.sync()
res = {}


def timeit(fn, reps=args.reps, warm=2):
    for _ in range(warm):
        fn()
    ctx.sync()
    ts = []
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(ctx.stream):
            e0.record()
            fn()
            e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[0], ts[len(ts) // 2]


def report(name, best, med, flops, bytes_):
    tf = flops / best / 1e9
    gbs = bytes_ / best / 1e6
    res[name] = {"ms_best": best, "ms_median": med, "tflops": tf, "gbs": gbs, "frac_fp64": tf / PEAK_TF,
                 "frac_hbm": gbs / PEAK_GBS}
    print(f"{name:34s} best {best:9.3f} ms  med {med:9.3f} ms  {tf:7.2f} TF/s ({tf / PEAK_TF:5.1%})  "
          f"{gbs:8.1f} GB/s ({gbs / PEAK_GBS:5.1%})", flush=True)


if not args.only or "k1" in args.only:
    out = ctx.empty(P // s * R)
    for x in ([2, 0, 1] if N >= 3 else [0, 1]):
        b, m = timeit(lambda: ctx.ttm_first(V, lens, x, W[x], R, out))
        report(f"k1 ttm_first x={x} R={R}", b, m, 2.0 * P * R, 8.0 * (P + P // s * R + s * R))
    if R != 10:
        W10 = ctx.empty(s * 10)
        ctx.fill_uniform(W10, 2, 9)
        for x in [2, 0]:
            b, m = timeit(lambda: ctx.ttm_first(V, lens, x, W10, 10, out))
            report(f"k1 ttm_first x={x} R=10", b, m, 2.0 * P * 10, 8.0 * (P + P // s * 10 + s * 10))

if (not args.only or "k1f" in args.only) and N == 4:
    outf = ctx.empty(s * s * R)
    for xf in [2, 0]:
        b, m = timeit(lambda: ctx.ttm_multi(V, lens, xf, [W[xf], W[xf + 1]], R, outf))
        report(f"k1 fused ttm_multi x={xf},{xf+1} R={R}", b, m, 2.0 * P * R, 8.0 * (P + s * s * R + 2 * s * R))

if (not args.only or "k2" in args.only) and N == 4:
    T = ctx.empty(s**3 * R)
    ctx.fill_uniform(T, 3, 0)
    out2 = ctx.empty(s**2 * R)
    for x in [2, 1, 0]:
        b, m = timeit(lambda: ctx.mttv(T, [s, s, s], x, W[x], R, out2))
        report(f"k2 mttv level-2 x={x}", b, m, 2.0 * s**3 * R, 8.0 * (s**3 * R + s**2 * R))
    out3 = ctx.empty(s * R)
    for x in [1, 0]:
        b, m = timeit(lambda: ctx.mttv(T, [s, s], x, W[x], R, out3), reps=20)
        report(f"k2 mttv leaf x={x}", b, m, 2.0 * s**2 * R, 8.0 * (s**2 * R + s * R))
    del T

if (not args.only or "k3" in args.only):
    ops = []
    for j in range(N - 1):
        t = ctx.empty(s * s * R)
        ctx.fill_uniform(t, 4, j)
        ops.append(t)
    M0 = ctx.empty(s * R)
    Mo = ctx.empty(s * R)
    for i in [0, N - 1, 1]:
        which = [0] * i + [1] * (N - 1 - i)
        b, m = timeit(lambda: ctx.pp_correct(M0, ops, which, [W[j] for j in range(N) if j != i], [s] * (N - 1), s, R, Mo),
                      reps=20)
        report(f"k3 pp_correct mode {i}", b, m, 2.0 * (N - 1) * s * s * R, 8.0 * ((N - 1) * (s * s * R + s * R) + 2 * s * R))

if (not args.only or "k5" in args.only):
    G = [ctx.empty(R * R) for _ in range(N)]
    b, m = timeit(lambda: ctx.gram(W[0], s, R, G[0]), reps=20)
    report("k4 gram", b, m, 2.0 * s * R * R, 8.0 * (s * R + R * R))
    for i in range(N):
        ctx.gram(W[i], s, R, G[i])
    S = ctx.empty(R * R)
    b, m = timeit(lambda: ctx.hadamard_grams(G, 0, R, 0.0, S), reps=20)
    report("k4 hadamard", b, m, R * R * (N - 2), 8.0 * R * R * N)
    M = ctx.empty(s * R)
    ctx.fill_uniform(M, 5, 0)
    Wc = W[0].clone()
    gr, dw = ctx.empty(s * R), ctx.empty(s * R)
    for mode in (0, 1):
        b, m = timeit(lambda: ctx.solve_update(M, S, Wc, s, R, W_init=W[1], mode=mode, grad=gr, dW=dw), reps=20)
        report(f"k5 solve_update mode={mode}", b, m, 4.0 * s * R * R, 8.0 * (5 * s * R + R * R))
    b, m = timeit(lambda: ctx.normalize([w for w in W], [s] * N, R, G), reps=20)
    report("k6 normalize", b, m, 2.0 * N * s * R, 8.0 * 3 * N * s * R)

if (not args.only or "k7" in args.only):
    sq = ctx.empty(1)
    b, m = timeit(lambda: ctx.cp_residual(V, lens, W, R, sq), reps=3, warm=1)
    report("k7 cp_residual", b, m, 2.0 * P * R, 8.0 * P)

os.makedirs("gpurun_out", exist_ok=True)
json.dump({"config": vars(args), "results": res}, open("gpurun_out/kernels.json", "w"), indent=1)
