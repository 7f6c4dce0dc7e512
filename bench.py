#!/usr/bin/env python
"""bench.py -- ALS / pairwise-perturbation sweeps per second on the BASELINE.json headline configuration:
CP order-4, s=300, R=50, FP64 (configs[1]), synthetic tensor 'r' built on the device from seeded random factors.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched under torch.distributed.run)
    python bench.py --impl reference --gpus N --steps K --warmup W

A "step" is ONE exact ALS sweep with the dimension tree over all four modes (2 first contractions, 2 level-2 and
4 leaf Hadamard contractions, 4 Gram-Hadamard solves, Normalize) -- als_CP.cxx:215-303 of the reference.
`value`   : sweeps/s with every input resident in HBM, device-timed (CUDA events on the engine's stream), max over
            ranks.  The tensor (64.8 GB) is larger than L2, so nothing is flushed between steps.
`e2e`     : the same sweep through the reference-facing driver call alsCP_DT(V, W, grad_W, F, ..., maxiter=0) with the
            factor and gradient matrices in pinned HOST memory: per step they are copied host->device, the sweep
            runs, and they are copied back; wall clock around the whole thing.  The tensor V is the data set of the
            iteration (it never changes between sweeps, exactly like CTF keeps it distributed in memory) and stays
            resident; h2d/d2h bytes are counted from the matrices copied.
`pp`      : the PP phase with the reference's pp_bench protocol: operator build, then the approximate sweep.
`roofline`: the first dimension-tree contraction (K1), timed alone with CUDA events; FP64 tensor pipe roofline.
`tucker`  : side measurement, not part of `value`: hosvd + alsTucker_DT sweeps at BASELINE configs[2] (order-3 s=800
            ranks 40, tensor 'r2'), ms per HOOI sweep through the same C++ drivers.
`cpu_baseline`: oracle/pp_oracle.py (NumPy/OpenBLAS restatement of the reference, all host cores) on a bounded
            mode-0 slab of the same tensor, scaled to the full tensor.  Cyclops CTF + MPI cannot be built in this
            image; oracle/_ref (the reference's sources on a loop-based CTF stand-in) is a checker whose contraction
            engine is ours and slow, so timing it would flatter the GPU: the OpenBLAS-backed port is the baseline.
With N>1 the tensor is sharded along mode 0 (strong scaling: the problem is fixed); the only collectives are the
NCCL all-reduces of the s x R partial MTTKRPs and the R x R Gram of the sharded factor.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FP64_PEAK_FALLBACK_TF = 35.46  # cuBLAS DGEMM 8192^3 measured on this pool's B200 (profiles/r01_fp64_peak.json)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=300)
    ap.add_argument("--rank", type=int, default=50)
    ap.add_argument("--order", type=int, default=4)
    ap.add_argument("--pp-sweeps", type=int, default=10)
    ap.add_argument("--cpu-slab", type=int, default=8,
                    help="mode-0 rows of the smaller of the two slabs (the other has twice as many) timed for the CPU baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-tucker", action="store_true", help="skip the Tucker HOOI side measurement (BASELINE configs[2])")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------------------
# CPU side: the oracle as baseline (the only place bench.py touches oracle/)
# ---------------------------------------------------------------------------------------------------------------
def cpu_sweep_time(s, R, N, slab, reps):
    """Seconds per exact ALS-DT sweep of the oracle on a mode-0 slab of `slab` rows; returns (sec, cores)."""
    import numpy as np

    from oracle import pp_oracle as o

    lens = (slab,) + (s,) * (N - 1)
    Wt = [o.fill_uniform((l, R), 1, i) for i, l in enumerate(lens)]
    V = np.asfortranarray(o.build_V(Wt))
    W = [o.fill_uniform((l, R), 2, i) for i, l in enumerate(lens)]
    G = [np.zeros_like(w) for w in W]
    parent, sibling = {}, {}
    o.construct_dimension_tree(parent, sibling, 0, N - 1)

    def sweep():
        mm = {}
        for i in range(N):
            M = o._leaf_M(mm, parent, sibling, V, W, i)
            S = o.gram_hadamard(W, i, 0.0, True)
            G[i] = -M + W[i] @ S
            W[i] = o.SVD_solve(M, S)
        o.normalize(W)

    sweep()  # warm-up (BLAS thread pool, page faults)
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        sweep()
        ts.append(time.perf_counter() - t0)
    return min(ts), os.cpu_count() or 1


def cpu_full_sweep_estimate(s, R, N, slab, reps):
    """Seconds per sweep of the full tensor from two mode-0 slabs (slab and 2*slab rows): the contractions that read V
    scale with the slab, the ones that start from an intermediate that has lost mode 0 do not, so the time is
    a + b*rows and is extrapolated linearly to s rows.  Returns (sec_full, sec_slab, sec_2slab, cores)."""
    slab2 = min(2 * slab, s)
    t1, cores = cpu_sweep_time(s, R, N, slab, reps)
    if slab2 == slab:
        return t1, t1, t1, cores
    t2, _ = cpu_sweep_time(s, R, N, slab2, reps)
    per_row = max((t2 - t1) / (slab2 - slab), 0.0)
    return t1 + per_row * (s - slab), t1, t2, cores


def ref_standin_sweep(s, R, N, small=28):
    """For the record only: the reference's own alsCP_DT (oracle/_ref/pp_bench, the unmodified sources on the loop-based
    CTF stand-in) on a size-`small` cube, scaled by (s/small)^N.  One scalar thread; never used as the baseline."""
    exe = os.path.join(ROOT, "oracle", "_ref", "pp_bench")
    if not os.path.exists(exe):
        return None
    import re
    import tempfile
    try:
        with tempfile.TemporaryDirectory() as td:
            out = subprocess.run([exe, "-model", "CP", "-tensor", "r", "-dim", str(N), "-size", str(small), "-rank", str(R),
                                  "-maxiter", "1", "-filename", os.path.join(td, "x.csv")], capture_output=True, text=True,
                                 timeout=300, cwd=td).stdout
        ts = [float(x) for x in re.findall(r"\[dimension tree step time\]\s+(\S+)", out)]
        if not ts:
            return None
        sec = min(ts) * (s / small) ** N
        return {"value": 1.0 / sec, "unit": "sweeps/s", "cores": 1, "kind": "reference",
                "sample": "oracle/_ref/pp_bench (reference sources, loop-based CTF stand-in) at size %d, one DT sweep "
                          "%.3f s, scaled x%.0f; NOT the baseline: the stand-in's contraction engine is ours and scalar"
                          % (small, min(ts), (s / small) ** N)}
    except Exception as exc:  # noqa: BLE001
        return {"error": "%s: %s" % (type(exc).__name__, exc)}


def run_reference(args):
    """The reference arm: the reference's own algorithm on the host cores (the oracle port, NumPy + OpenBLAS on all
    cores -- CTF is not buildable here and oracle/_ref's loop-based stand-in engine would be an unfairly slow
    baseline), each step a bounded mode-0 slab of the same workload, scaled to the full tensor."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    s, R, N, slab = args.size, args.rank, args.order, min(args.cpu_slab, args.size)
    t0 = time.perf_counter()
    full, sec, sec2, cores = cpu_full_sweep_estimate(s, R, N, slab, max(1, min(args.steps, 3)))
    val = 1.0 / full
    line = {
        "impl": "reference", "metric": "ALS-DT sweeps/s (CP order-%d s=%d R=%d FP64)" % (N, s, R), "value": val,
        "unit": "sweeps/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": full * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "CP-ALS dimension-tree sweep, order-%d s=%d R=%d, tensor 'r'" % (N, s, R),
                   "sample": "mode-0 slabs of %d and %d of %d rows, linear extrapolation" % (slab, min(2 * slab, s), s)},
        "cpu_baseline": {"value": val, "unit": "sweeps/s", "cores": cores, "kind": "port",
                         "sample": "oracle/pp_oracle.py (NumPy + OpenBLAS, %d threads) on mode-0 slabs of %d and %d of %d "
                                   "rows: %.2f s and %.2f s per sweep, extrapolated linearly to %.1f s"
                                   % (cores, slab, min(2 * slab, s), s, sec, sec2, full)},
        "e2e": {"value": val, "unit": "sweeps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "reference_sources_on_standin": ref_standin_sweep(s, R, N), "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi polled every 20 ms from before the warm-up; stop() keeps the samples taken inside the timed window
    (mark_begin .. mark_end) -- nvidia-smi needs a few hundred ms to start, longer than a short timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.lines, self.proc = device, [], None
        self.t_begin = self.t_end = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def wait_first_sample(self, timeout=5.0):
        t0 = time.perf_counter()
        while self.proc and not self.lines and time.perf_counter() - t0 < timeout:
            time.sleep(0.01)

    def mark_begin(self):
        self.t_begin = time.perf_counter()

    def mark_end(self):
        self.t_end = time.perf_counter()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()

        def parse(rows):
            sm, mx, reasons = [], [], set()
            for ln in rows:
                f = [x.strip() for x in ln.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"],
                                   f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            return sm, mx, reasons

        inside = [ln for t, ln in self.lines if self.t_begin is not None and self.t_begin <= t <= (self.t_end or 1e300)]
        sm, mx, reasons = parse(inside)
        window = "timed region"
        if not sm:  # a very short timed region: fall back to every sample since the warm-up began
            sm, mx, reasons = parse([ln for _, ln in self.lines])
            window = "warm-up + timed region"
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx[0] if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


# ---------------------------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    import numpy as np
    import torch

    rank = int(os.environ.get("RANK", "0"))
    nranks = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if nranks > 1:
        import torch.distributed as dist

        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    ppx = importlib.import_module("pairwise-perturbation_b200")
    H = importlib.import_module("pairwise-perturbation_b200.host_api")
    lib = ppx.load_library()

    s, R, N = args.size, args.rank, args.order
    world = H.World(local_rank, solver=0, use_graph=True, workspace_bytes=1 << 30)
    b, e = ppx.shard_range(s, nranks, rank)
    if nranks > 1:
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt = torch.tensor(list(ppx.comm_unique_id()), dtype=torch.uint8, device="cuda")
        dist.broadcast(idt, 0)
        world.comm_init(bytes(idt.cpu().tolist()), nranks, rank, 0, s, b, e)
    rows0 = e - b

    def barrier():
        world.sync()
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def seeded_matrix(seed, mode):
        """s x R uniform [0,1) factor (full), sliced to the local rows for the sharded mode 0."""
        m = H.Matrix(world, s, R)
        m.fill(seed, mode)
        if mode != 0 or nranks == 1:
            return m
        full = m.numpy()
        m.free()
        return H.Tensor.from_numpy(world, np.ascontiguousarray(full[b:e]), matrix=True)

    # synthetic tensor 'r' (test_ALS.cxx:275-286): V = [[W_true]], local slab of mode 0
    Wt = [seeded_matrix(1, i) for i in range(N)]
    lens_local = (rows0,) + (s,) * (N - 1)
    V = H.Tensor(world, lens_local)
    H.build_V(world, V, Wt)
    for w in Wt:
        w.free()
    W0_host = [seeded_matrix(2, i) for i in range(N)]
    W = W0_host
    G = [seeded_matrix(3, i) for i in range(N)]
    F = [H.Matrix(world, w.lens[0], R) for w in W]
    W_start = [w.numpy() for w in W]
    world.sync()

    def reset_W():
        for w, h in zip(W, W_start):
            w.write(h)

    evs = [C.c_void_p() for _ in range(2)]
    for ev in evs:
        lib.ppx_event_create(world.ctx_handle(), C.byref(ev))

    # ---- device-resident sweeps: value ---------------------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        sampler.wait_first_sample()
    H.cp_dt_sweeps(world, V, W, G, max(args.warmup, 3))
    barrier()
    sampler.mark_begin()
    launches0 = world.launch_count()
    lib.ppx_event_record(world.ctx_handle(), evs[0])
    H.cp_dt_sweeps(world, V, W, G, args.steps)
    lib.ppx_event_record(world.ctx_handle(), evs[1])
    ms = C.c_float(0)
    lib.ppx_event_elapsed_ms(world.ctx_handle(), evs[0], evs[1], C.byref(ms))
    barrier()
    sampler.mark_end()
    launches = world.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms_total = max_over_ranks(float(ms.value))
    ms_per_step = ms_total / args.steps
    per_rank_ms = [float(ms.value) / args.steps]
    if dist is not None:  # every rank's own device time per step (the collectives make them wait for the slowest)
        t_all = [torch.zeros(1, dtype=torch.float64, device="cuda") for _ in range(nranks)]
        dist.all_gather(t_all, torch.tensor([per_rank_ms[0]], dtype=torch.float64, device="cuda"))
        per_rank_ms = [float(t.item()) for t in t_all]
    value = 1e3 / ms_per_step

    # ---- PP phase (pp_bench protocol): operator build + approximate sweeps ----------------------------------------
    reset_W()
    H.cp_dt_sweeps(world, V, W, G, 2)  # PP starts from an ALS iterate, as alsCP_PP does
    barrier()
    H.cp_pp_phase_timed(world, V, W, G, 1)  # warm-up: first-use kernel loading and the 3 x 10.8 GB allocations
    reset_W()
    H.cp_dt_sweeps(world, V, W, G, 2)
    barrier()
    ms_build, ms_pp = H.cp_pp_phase_timed(world, V, W, G, args.pp_sweeps)
    ms_build, ms_pp = max_over_ranks(ms_build), max_over_ranks(ms_pp)
    world.trim()  # the level-1 tensors of the build go back to the driver
    pp = {"operator_build_ms": ms_build, "approx_sweep_ms": ms_pp / args.pp_sweeps,
          "approx_sweeps_per_s": 1e3 * args.pp_sweeps / ms_pp, "sweeps_timed": args.pp_sweeps,
          "note": "one CUDA graph per sweep (N x [correct, Gram-Hadamard, LDL^T inverse + solve+grad+dW, Gram] + Normalize + norms)"
                  if nranks == 1 else "eager launches + NCCL all-reduce per mode (latency bound, does not scale)"}

    # ---- K3 (PP correction of one mode) and the solve kernels alone, as HBM GB/s ----------------------
    if nranks == 1:
        try:
            i_mode = 1
            s_i = lens_local[i_mode]
            others = [j for j in range(N) if j != i_mode]
            ops = [H.Tensor(world, (s_i, lens_local[j], R)) for j in others]
            for k_, t_ in enumerate(ops):
                t_.fill(4, k_)
            M0, Mo = H.Matrix(world, s_i, R), H.Matrix(world, s_i, R)
            n_ops = len(ops)
            opp = (C.c_void_p * n_ops)(*[t_.data_ptr() for t_ in ops])
            dwp = (C.c_void_p * n_ops)(*[W[j].data_ptr() for j in others])
            which = (C.c_int * n_ops)(*[0 if j < i_mode else 1 for j in others])
            so = (C.c_int64 * n_ops)(*[lens_local[j] for j in others])

            def k3():
                rc_ = lib.ppx_pp_correct(world.ctx_handle(), C.c_void_p(M0.data_ptr()), opp, which, dwp, so, n_ops, s_i, R,
                                         C.c_void_p(Mo.data_ptr()))
                assert rc_ == 0, lib.ppx_last_error(world.ctx_handle())

            for _ in range(5):
                k3()
            lib.ppx_event_record(world.ctx_handle(), evs[0])
            for _ in range(20):
                k3()
            lib.ppx_event_record(world.ctx_handle(), evs[1])
            lib.ppx_event_elapsed_ms(world.ctx_handle(), evs[0], evs[1], C.byref(ms))
            k3_ms = float(ms.value) / 20
            k3_bytes = 8.0 * (sum(s_i * lens_local[j] * R + lens_local[j] * R for j in others) + 2 * s_i * R)
            hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] \
                if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6543.1
            pp["k3_pp_correct"] = {"us": 1e3 * k3_ms, "algorithmic_bytes": k3_bytes, "gbs": k3_bytes / k3_ms / 1e6,
                                   "frac_hbm": k3_bytes / k3_ms / 1e6 / hbm,
                                   "timing": "20 back-to-back launches, CUDA events (launch ramp included)"}
            sweep_bytes = 8.0 * sum(lens_local[i] * lens_local[j] * R + lens_local[j] * R
                                    for i in range(N) for j in range(N) if i != j) + 24.0 * sum(lens_local) * R
            pp["approx_sweep_algorithmic_bytes"] = sweep_bytes
            pp["approx_sweep_frac_hbm"] = sweep_bytes / (ms_pp / args.pp_sweeps) / 1e6 / hbm
            for t_ in ops + [M0, Mo]:
                t_.free()
        except Exception as exc:  # noqa: BLE001
            pp["k3_pp_correct"] = {"error": "%s: %s" % (type(exc).__name__, exc)}

    # ---- roofline of the dominant kernel: the first dimension-tree contraction (K1) ------------------------------
    # The sweep contracts the sibling modes of the first-level tree node in ONE launch (ppx_ttm_multi: DMMA GEMM against
    # the Khatri-Rao rows of their factors); that call is what is timed here.
    x_first = (N - 1) // 2 + 1
    n_modes = N - x_first
    keep = lens_local[:x_first]
    out1 = H.Tensor(world, tuple(keep) + (R,))
    lens_c = (C.c_int64 * N)(*lens_local)
    wptrs = (C.c_void_p * n_modes)(*[W[x_first + j].data_ptr() for j in range(n_modes)])
    wld = (C.c_int64 * n_modes)(*[lens_local[x_first + j] for j in range(n_modes)])

    def k1():
        rc = lib.ppx_ttm_multi(world.ctx_handle(), C.c_void_p(V.data_ptr()), lens_c, N, x_first, n_modes, wptrs, wld, R,
                               C.c_void_p(out1.data_ptr()))
        assert rc == 0, lib.ppx_last_error(world.ctx_handle())

    for _ in range(3):
        k1()
    barrier()
    reps = 5
    lib.ppx_event_record(world.ctx_handle(), evs[0])
    for _ in range(reps):
        k1()
    lib.ppx_event_record(world.ctx_handle(), evs[1])
    lib.ppx_event_elapsed_ms(world.ctx_handle(), evs[0], evs[1], C.byref(ms))
    k1_ms = float(ms.value) / reps
    k1_per_rank = [k1_ms]
    if dist is not None:
        t_all = [torch.zeros(1, dtype=torch.float64, device="cuda") for _ in range(nranks)]
        dist.all_gather(t_all, torch.tensor([k1_ms], dtype=torch.float64, device="cuda"))
        k1_per_rank = [float(t.item()) for t in t_all]
        k1_ms = max(k1_per_rank)
    out1.free()
    P_local = float(np.prod(lens_local))
    Kc = float(np.prod(lens_local[x_first:]))
    k1_flops = 2.0 * P_local * R
    k1_bytes = 8.0 * (P_local + P_local / Kc * R + sum(lens_local[x_first:]) * R)
    peak_tf, peak_src = FP64_PEAK_FALLBACK_TF, "profiles/r01_fp64_peak.json (cuBLAS DGEMM 8192^3 on this pool's B200)"
    try:
        peak_tf = json.load(open(os.path.join(ROOT, "profiles", "r01_fp64_peak.json")))["fp64_tflops_burst"]
    except Exception:
        pass
    traffic = None
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "k1_ncu_summary.json")))
        if nranks == 1 and [s, R, N] == prof.get("size_rank_order"):
            traffic = prof.get("dram_bytes_per_launch")
    except Exception:
        pass
    achieved = k1_flops / (k1_ms * 1e-3) / 1e12
    roofline = {"kernel": "ttm_tma_kernel via ppx_ttm_multi (K1, first dimension-tree contraction, TMA-staged, modes %d..%d at once, "
                          "R=%d)" % (x_first, N - 1, R),
                "bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                "traffic": traffic, "algorithmic_flops": k1_flops, "algorithmic_bytes": k1_bytes, "ms": k1_ms,
                "hbm_gbs": k1_bytes / (k1_ms * 1e-3) / 1e9,
                "peak_source": "FP64 peak is not in MEASURED_PEAKS.json (it holds HBM and bf16); " + peak_src,
                "share_of_step": 2 * k1_ms / ms_per_step, "ms_per_rank": k1_per_rank}

    # ---- e2e: the driver call with host buffers --------------------------------------------------------------------
    pin = [torch.from_numpy(h.ravel(order="F").copy()).pin_memory() for h in W_start]
    pin_g = [torch.zeros(h.size, dtype=torch.float64).pin_memory() for h in W_start]
    pin_out = [torch.empty(h.size, dtype=torch.float64).pin_memory() for h in W_start]
    hlib = H.load_host_library()

    def e2e_step():
        for t_, w in zip(pin, W):
            hlib.ppxh_tensor_write(w.h, C.c_void_p(t_.data_ptr()))
        for t_, g in zip(pin_g, G):
            hlib.ppxh_tensor_write(g.h, C.c_void_p(t_.data_ptr()))
        H.alsCP_DT(world, V, W, G, F, 0.0, 0, lam=0.0, resprint=1 << 30, bench=True)  # exactly one sweep
        for t_, w in zip(pin_out, W):
            hlib.ppxh_tensor_read(w.h, C.c_void_p(t_.data_ptr()))
        for t_, g in zip(pin_g, G):
            hlib.ppxh_tensor_read(g.h, C.c_void_p(t_.data_ptr()))

    with H.Trace(quiet=True, skip_residual=True):
        for _ in range(max(1, min(args.warmup, 2))):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        barrier()
        e2e_s = max_over_ranks(time.perf_counter() - t0)
    bytes_mats = sum(h.size for h in W_start) * 8
    e2e = {"value": args.steps / e2e_s, "unit": "sweeps/s", "h2d_bytes_per_step": 2 * bytes_mats,
           "d2h_bytes_per_step": 2 * bytes_mats, "ms_per_step": 1e3 * e2e_s / args.steps,
           "call": "alsCP_DT(V, W, grad_W, F, tol, timelimit, maxiter=0, lambda, csv, resprint, bench=true, dw) with W/grad_W "
                   "copied from/to pinned host memory every step; V (the data set) resident"}

    # ---- CPU baseline ---------------------------------------------------------------------------------------------
    cpu = None
    if rank == 0 and nranks == 1 and not args.no_cpu_baseline:
        slab = min(args.cpu_slab, s)
        full, sec, sec2, cores = cpu_full_sweep_estimate(s, R, N, slab, 2)
        cpu = {"value": 1.0 / full, "unit": "sweeps/s", "cores": cores, "kind": "port",
               "sample": "oracle/pp_oracle.py (NumPy + OpenBLAS restatement of als_CP.cxx, %d threads) on mode-0 slabs "
                         "of %d and %d of %d rows: %.2f s and %.2f s per sweep, extrapolated linearly to %.1f s; Cyclops "
                         "CTF + MPI are not buildable in this image and oracle/_ref runs on a loop-based stand-in (a "
                         "checker, not a fair baseline)" % (cores, slab, min(2 * slab, s), s, sec, sec2, full),
               "reference_sources_on_standin": ref_standin_sweep(s, R, N)}

    # ---- side measurement: Tucker HOOI at BASELINE configs[2] (order-3 s=800 ranks 40, tensor 'r2') ---------------
    # Not part of `value`; reported so that the Tucker half of the path has a number in the same file.  Never allowed
    # to take the headline down: any failure is recorded in the object instead.
    tucker = None
    if not args.no_tucker:
        try:
            for t_ in [V] + W + G + F:
                t_.free()
            world.trim()
            ts, tr, tn = 800, 40, 3
            tb, te = ppx.shard_range(ts, nranks, rank)
            if nranks > 1:
                world.set_shard(0, ts, tb, te)
            full = H.Tensor(world, (ts,) * tn)
            full.fill(1, 100, 0.5, 1.0)
            if nranks > 1:  # local rows of mode 0 (the fastest index): a strided device copy
                Vt = H.Tensor(world, (te - tb,) + (ts,) * (tn - 1))
                rc = lib.ppx_memcpy2d_d2d(world.ctx_handle(), C.c_void_p(Vt.data_ptr()), C.c_size_t(8 * (te - tb)),
                                          C.c_void_p(full.data_ptr() + 8 * tb), C.c_size_t(8 * ts),
                                          C.c_size_t(8 * (te - tb)), C.c_size_t(ts ** (tn - 1)))
                assert rc == 0
                world.sync()
                full.free()
            else:
                Vt = full
            Wt_ = [H.Matrix(world, ts, tr) for _ in range(tn)]
            core = H.Tensor(world, (tr,) * tn)
            H.hosvd(world, Vt, core, Wt_, [tr] * tn)  # warm-up: first-use kernel loading and pool allocations
            barrier()
            t0 = time.perf_counter()
            H.hosvd(world, Vt, core, Wt_, [tr] * tn)
            barrier()
            t_hosvd = max_over_ranks(time.perf_counter() - t0)
            nsw = 6
            with H.Trace(quiet=True, skip_residual=True):
                H.alsTucker_DT(world, Vt, core, Wt_, 0.0, 1, resprint=1 << 30)  # warm-up: 2 sweeps
                barrier()
                t0 = time.perf_counter()
                H.alsTucker_DT(world, Vt, core, Wt_, 0.0, nsw - 1, resprint=1 << 30)
                barrier()
                t_sw = max_over_ranks(time.perf_counter() - t0)
            tucker = {"workload": "Tucker HOOI (alsTucker_DT) order-3 s=800 ranks 40, tensor 'r2' (BASELINE configs[2])",
                      "hosvd_ms": 1e3 * t_hosvd, "ms_per_sweep": 1e3 * t_sw / nsw, "sweeps_per_s": nsw / t_sw,
                      "sweeps_timed": nsw, "timing": "wall clock around the driver call, synchronised on both sides"}
        except Exception as exc:  # noqa: BLE001
            tucker = {"error": "%s: %s" % (type(exc).__name__, exc)}

    if rank == 0:
        line = {
            "metric": "ALS-DT sweeps/s (CP order-%d s=%d R=%d FP64)" % (N, s, R), "value": value, "unit": "sweeps/s",
            "n_gpus": nranks, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "CP-ALS dimension-tree sweep, order-%d s=%d R=%d, tensor 'r' (BASELINE configs[1])"
                                   % (N, s, R),
                       "tensor_bytes_per_gpu": 8 * P_local, "l2": "inputs (%.1f GB) larger than L2; no flush" % (8 * P_local / 1e9),
                       "parallelism": "mode-0 shards x%d, NCCL all-reduce of s x R partial MTTKRPs" % nranks if nranks > 1
                       else "single GPU", "solver": "cholesky"},
            "ms_per_step_per_rank": per_rank_ms, "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu, "pp": pp,
            "tucker": tucker,
        }
        print(json.dumps(line), flush=True)
    world.close()
    if dist is not None:
        dist.destroy_process_group()


import ctypes as C  # noqa: E402

if __name__ == "__main__":
    main()
