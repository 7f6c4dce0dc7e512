#!/usr/bin/env python
"""bench.py -- ALS / pairwise-perturbation sweeps per second on the BASELINE.json configurations, headline =
configs[1]: CP order-4, s=300, R=50, FP64, synthetic tensor 'r' built on the device from seeded random factors.

    python bench.py --gpus N --steps K --warmup W [--workload cfg2|cfg1|cfg4|cfg5]   (N>1: under torch.distributed.run)
    python bench.py --impl reference --gpus N --steps K --warmup W                   (the CPU arm, rank 0 only)

A "step" is ONE exact ALS sweep with the dimension tree over all modes (for order 4: 2 fused first contractions, 4 leaf
Hadamard contractions, 4 Gram-Hadamard solves, Normalize) -- als_CP.cxx:215-303 of the reference.
`value`    : sweeps/s with every input resident in HBM, device-timed (CUDA events on the engine's stream), max over
             ranks.  The tensor (64.8 GB at cfg2) is larger than L2, so nothing is flushed between steps.
`e2e`      : the same sweep through the reference-facing driver call alsCP_DT(V, W, grad_W, F, ..., maxiter=0) with the
             factor and gradient matrices in pinned HOST memory: per step they are copied host->device, the sweep
             runs, and they are copied back; wall clock around the whole thing.  The tensor V is the data set of the
             iteration (it never changes between sweeps, exactly like CTF keeps it distributed in memory) and stays
             resident; h2d/d2h bytes are counted from the matrices copied.
`roofline` : the first dimension-tree contraction (K1), timed alone with CUDA events.  The bound is chosen from the
             arithmetic intensity against the ridge of the two MEASURED peaks: FP64 = cuBLAS DGEMM 8192^3 timed in this
             very process (MEASURED_PEAKS.json has no FP64 entry), HBM = MEASURED_PEAKS.json.
`parity_probe`: closed-form known answers at the FULL size of the run (pairwise-perturbation_b200/kat.py): the tensor is
             exactly [[A]], so the MTTKRP of every mode and every PP operator have closed forms in s x R matrices; max
             relative error over all of them on this rank's shard, max over ranks.  At N=1 also the factors after 3 sweeps
             against the CPU port's.
`pp`       : the PP phase with the reference's pp_bench protocol (operator build, then the approximate sweep), K3 alone,
             and `mixed_run`: alsCP_PP (-pp 1) for --pp-maxiter iterations -- #DT sweeps, #PP sweeps, #operator builds,
             total time, sweeps/s (residual evaluations off the clock as in als_CP.cxx:189).
`tucker`   : side measurement, not part of `value`: hosvd + alsTucker_DT sweeps at BASELINE configs[2] (order-3 s=800
             ranks 40, tensor 'r2') with roofline objects for the first TTM and the HOSVD unfolding Gram.
`cpu_baseline` / `--impl reference`: oracle/pp_oracle.py (NumPy + OpenBLAS restatement of the reference's sweep; Cyclops
             CTF + MPI + ScaLAPACK cannot be built in this image) on the host cores, on the SAME full tensor when the
             host has the memory (it needs ~80 GB at cfg2), BLAS threads set explicitly in a clean subprocess (torchrun
             exports OMP_NUM_THREADS=1) and the count actually used read back with threadpoolctl.
With N>1 the tensor is sharded along mode 0 (strong scaling: the problem is fixed); the only collectives are the
NCCL all-reduces of the s x R partial MTTKRPs and the R x R Gram of the sharded factor.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (lens, R, pp_res_tol, which BASELINE.json config)
    "cfg2": ((300, 300, 300, 300), 50, 1e-2, "BASELINE configs[1]"),
    "cfg1": ((200, 200, 200), 10, 1e-2, "BASELINE configs[0]"),
    "cfg4": ((40,) * 6, 10, 1e-2, "BASELINE configs[3]"),
    "cfg5": ((3, 128, 128, 7200), 10, 5e-2, "BASELINE configs[4], coil-100 shape, synthetic"),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--size", type=int, default=None, help="development: cubic tensor of this size instead of --workload")
    ap.add_argument("--rank", type=int, default=None)
    ap.add_argument("--order", type=int, default=None)
    ap.add_argument("--pp-sweeps", type=int, default=10)
    ap.add_argument("--pp-maxiter", type=int, default=50, help="iterations of the -pp 1 mixed run (alsCP_PP)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-tucker", action="store_true", help="skip the Tucker HOOI side measurement (BASELINE configs[2])")
    ap.add_argument("--no-pp", action="store_true", help="skip the PP phase / mixed-run measurements")
    ap.add_argument("--cpu-steps", type=int, default=2, help="timed sweeps of the cpu_baseline leg of the GPU arm")
    ap.add_argument("--dump-factors", default=None, help="(reference arm) write the factors after the last sweep here")
    return ap.parse_args()


def workload_of(args):
    lens, R, tol_init, tag = WORKLOADS[args.workload]
    if args.size is not None or args.order is not None:
        n = args.order or len(lens)
        lens, tag = (args.size or lens[0],) * n, "custom"
    if args.rank is not None:
        R = args.rank
        tag = tag if args.rank == WORKLOADS[args.workload][1] else "custom"
    return tuple(lens), R, tol_init, tag


def workload_name(lens, R, tag):
    shape = "s=%d" % lens[0] if len(set(lens)) == 1 else "x".join(str(x) for x in lens)
    return "CP-ALS dimension-tree sweep, order-%d %s R=%d, tensor 'r' (%s)" % (len(lens), shape, R, tag)


def config_of(lens, R, tag, nranks):
    P = 1
    for x in lens:
        P *= x
    return {"workload": workload_name(lens, R, tag), "lens": list(lens), "rank": R,
            "tensor_bytes_total": 8 * P, "l2": "inputs (%.1f GB per GPU) larger than L2; no flush" % (8 * P / nranks / 1e9)
            if 8 * P / nranks > 2e8 else "inputs smaller than L2 (%.0f MB per GPU); no flush: a sweep rewrites every factor "
            "and intermediate it reads" % (8 * P / nranks / 1e6),
            "parallelism": "mode-0 shards x%d, NCCL all-reduce of s x R partial MTTKRPs" % nranks if nranks > 1
            else "single GPU", "solver": "cholesky"}


# ---------------------------------------------------------------------------------------------------------------
# CPU side: the oracle as baseline (the only place bench.py touches oracle/)
# ---------------------------------------------------------------------------------------------------------------
def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def clean_cpu_env():
    """Environment of the CPU arm: BLAS threads = the host cores this process may use, whatever the launcher exported
    (torch.distributed.run sets OMP_NUM_THREADS=1 for its workers)."""
    env = dict(os.environ)
    n = str(host_threads())
    for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        env[k] = n
    env["PPX_CPU_ARM_CHILD"] = "1"
    return env


def mem_available_bytes():
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable:"):
                return int(line.split()[1]) * 1024
    except Exception:
        pass
    return None


def cpu_build_V(np, o, Wt, rows0):
    """[[Wt]] restricted to the first rows0 rows of mode 0, as a first-index-fastest array, without a transposing copy:
    the memory of V in that order is the C-order matrix V2[(i_{N-1},..,i_1), i_0] = KRP(Wt_{N-1},..,Wt_1) Wt_0^T, filled
    slab by slab of the slowest mode with one GEMM each."""
    N = len(Wt)
    lens = [w.shape[0] for w in Wt]
    R = Wt[0].shape[1]
    K = np.ascontiguousarray(Wt[1]) if N > 2 else None
    for m in range(2, N - 1):
        K = (np.ascontiguousarray(Wt[m])[:, None, :] * K[None, :, :]).reshape(-1, R)
    W0T = np.ascontiguousarray(Wt[0][:rows0].T)
    if N == 2:
        return np.asfortranarray(Wt[0][:rows0] @ Wt[1].T)
    P = K.shape[0]
    V2 = np.empty((lens[N - 1] * P, rows0))
    last = np.ascontiguousarray(Wt[N - 1])
    for d in range(lens[N - 1]):
        np.matmul(K * last[d][None, :], W0T, out=V2[d * P:(d + 1) * P])
    return V2.reshape(tuple(lens[:0:-1]) + (rows0,)).T  # F-contiguous view with shape (rows0, s_1, .., s_{N-1})


def openblas_path():
    """The OpenBLAS shared library NumPy ships (numpy.libs/libscipy_openblas64_*.so), for the stand-in's DGEMM path."""
    try:
        import glob

        import numpy

        hits = glob.glob(os.path.join(os.path.dirname(numpy.__file__), "..", "numpy.libs", "libscipy_openblas*.so"))
        return os.path.abspath(hits[0]) if hits else None
    except Exception:
        return None


def reference_sources_crosscheck(R, N, small=100, sweeps=2):
    """For the record (not the baseline): the reference's OWN alsCP_DT -- oracle/_ref/test_ALS, the unmodified sources on
    the CTF stand-in with its DGEMM path switched on -- against the NumPy port on the same small cube, both on the same
    BLAS and thread count.  It answers "is the port a strawman?": the reference's code copies the tensor into V_front for
    every first-level tree node (common.cxx:31-33) and materialises every level-1 intermediate, so it is the slower one."""
    exe = os.path.join(ROOT, "oracle", "_ref", "test_ALS")
    blas = openblas_path()
    if not os.path.exists(exe) or not blas:
        return None
    import re

    import numpy as np
    from oracle import pp_oracle as o
    try:
        env = dict(os.environ, CTF_STANDIN_BLAS=blas)
        with tempfile.TemporaryDirectory() as td:
            out = subprocess.run([exe, "-model", "CP", "-tensor", "r", "-dim", str(N), "-size", str(small), "-rank", str(R),
                                  "-pp", "0", "-maxiter", str(sweeps), "-resprint", "1000000", "-filename",
                                  os.path.join(td, "x.csv")], capture_output=True, text=True, timeout=600, cwd=td, env=env).stdout
        m = re.findall(r"\[iter\]=\s+%d\s.*\[dtime\]\s+(\S+)" % sweeps, out)
        if not m:
            return {"error": "no [dtime] line"}
        ref_sec = float(m[-1]) / sweeps
        lens = (small,) * N
        Wt = [o.fill_uniform((l, R), 1, i) for i, l in enumerate(lens)]
        V = cpu_build_V(np, o, Wt, small)
        W = [o.fill_uniform((l, R), 2, i) for i, l in enumerate(lens)]
        parent, sibling = {}, {}
        o.construct_dimension_tree(parent, sibling, 0, N - 1)

        def sweep():
            mm = {}
            for i in range(N):
                M = o._leaf_M(mm, parent, sibling, V, W, i)
                S = o.gram_hadamard(W, i, 0.0, True)
                W[i] = o.SVD_solve(M, S)
            o.normalize(W)

        sweep()
        t0 = time.perf_counter()
        for _ in range(sweeps):
            sweep()
        port_sec = (time.perf_counter() - t0) / sweeps
        return {"size": small, "reference_sources_s_per_sweep": ref_sec, "numpy_port_s_per_sweep": port_sec,
                "what": "oracle/_ref/test_ALS (the reference's unmodified sources; CTF stand-in with its DGEMM path, same "
                        "OpenBLAS and threads) vs oracle/pp_oracle.py on an order-%d cube of size %d, R=%d; residual "
                        "evaluations off the clock in both" % (N, small, R)}
    except Exception as exc:  # noqa: BLE001
        return {"error": "%s: %s" % (type(exc).__name__, exc)}


def run_reference_child(args):
    """The reference arm proper (runs with the BLAS thread count fixed by the parent): the reference's sweep on the host
    cores, NumPy + OpenBLAS port (oracle/pp_oracle.py), full tensor when the host memory allows."""
    import numpy as np
    from oracle import pp_oracle as o

    try:
        from threadpoolctl import threadpool_info
        pools = [(p.get("internal_api"), p.get("num_threads")) for p in threadpool_info()]
        blas_threads = max([n for api, n in pools if api in ("openblas", "mkl", "blis")] or [0]) or None
    except Exception:
        pools, blas_threads = [], None
    t_start = time.perf_counter()
    lens, R, _tol_init, tag = workload_of(args)
    N = len(lens)
    P = 1
    for x in lens:
        P *= x
    avail = mem_available_bytes()
    # V + the largest level-1 intermediate (P/min(len) * R) + the KRP slab, with 25% head room
    need = lambda rows: 1.25 * 8 * (P // lens[0] * rows * (1 + R / min(lens[1:])) + P // lens[0] // lens[-1] * R * 3)
    rows0 = lens[0]
    if avail is not None:
        while rows0 > 1 and need(rows0) > 0.9 * avail:
            rows0 = max(1, rows0 // 2)
    full = rows0 == lens[0]
    Wt = [o.fill_uniform((l, R), 1, i) for i, l in enumerate(lens)]
    V = cpu_build_V(np, o, Wt, rows0)
    del Wt
    W = [o.fill_uniform((l, R), 2, i) for i, l in enumerate(lens)]
    W[0] = np.asfortranarray(W[0][:rows0])
    G = [np.zeros_like(w) for w in W]
    parent, sibling = {}, {}
    o.construct_dimension_tree(parent, sibling, 0, N - 1)
    t_built = time.perf_counter()

    def sweep():  # als_CP.cxx:215-303 as restated by oracle.alsCP_DT
        mm = {}
        for i in range(N):
            M = o._leaf_M(mm, parent, sibling, V, W, i)
            S = o.gram_hadamard(W, i, 0.0, True)
            G[i] = -M + W[i] @ S
            W[i] = o.SVD_solve(M, S)
        o.normalize(W)

    # bounded run: honour --steps/--warmup while the whole thing stays within a few minutes, else fewer timed sweeps
    budget_s = 200.0
    ts = []
    t0 = time.perf_counter()
    sweep()
    first = time.perf_counter() - t0
    warm = 1
    while warm < args.warmup and (warm + 1 + args.steps) * first < budget_s:
        sweep()
        warm += 1
    while len(ts) < args.steps and (len(ts) == 0 or time.perf_counter() - t0 + first < budget_s):
        t1 = time.perf_counter()
        sweep()
        ts.append(time.perf_counter() - t1)
    sec = sum(ts) / len(ts)
    scale = lens[0] / rows0
    sec_full = sec * scale
    if args.dump_factors:
        np.savez(args.dump_factors, *W, sweeps=warm + len(ts), rows0=rows0)
    cores = blas_threads or host_threads()
    sample = ("oracle/pp_oracle.py (NumPy + OpenBLAS restatement of als_CP.cxx:215-303; %s BLAS threads in use, read back with "
              "threadpoolctl) on %s: %d warm-up + %d timed sweeps, %.2f s per sweep (min %.2f, max %.2f)%s; tensor built on "
              "the host in %.1f s (off the clock); host MemAvailable %.0f GB"
              % (cores, "the FULL tensor (same configuration as the GPU arm)" if full else
                 "a mode-0 slab of %d of %d rows (the host cannot hold the tensor)" % (rows0, lens[0]), warm, len(ts), sec,
                 min(ts), max(ts), "" if full else ", scaled x%.2f to the full tensor" % scale, t_built - t_start,
                 (avail or 0) / 1e9))
    val = 1.0 / sec_full
    line = {
        "impl": "reference", "metric": "ALS-DT sweeps/s (CP order-%d FP64)" % N, "value": val, "unit": "sweeps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "steps_timed": len(ts), "warmup_done": warm,
        "ms_per_step": sec_full * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": config_of(lens, R, tag, max(args.gpus, 1)),
        "cpu_baseline": {"value": val, "unit": "sweeps/s", "cores": cores, "kind": "port", "sample": sample,
                         "full_tensor": full, "blas_pools": pools, "host_threads_available": host_threads()},
        "e2e": {"value": val, "unit": "sweeps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "reference_sources_on_blas_standin": reference_sources_crosscheck(R, N) if (args.dump_factors is None and len(set(lens)) == 1) else None,
        "wall_s": time.perf_counter() - t_start,
        "reference_note": "Cyclops CTF / MPI / ScaLAPACK are absent from this image and cannot be installed (no network); "
                          "oracle/_ref (the reference's own sources on a loop-based CTF stand-in) is a correctness checker "
                          "and would be an unfairly slow baseline, so the BLAS-backed port is what is timed",
    }
    print(json.dumps(line), flush=True)


def run_reference(args):
    """--impl reference: rank 0 alone, in a clean subprocess (BLAS threads fixed before NumPy is imported)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    if os.environ.get("PPX_CPU_ARM_CHILD") == "1":
        run_reference_child(args)
        return
    sys.exit(subprocess.call([sys.executable, os.path.abspath(__file__)] + sys.argv[1:], env=clean_cpu_env()))


def cpu_baseline_leg(args, dump_path):
    """cpu_baseline of the GPU arm: the reference arm with a bounded number of sweeps (1 warm-up + --cpu-steps timed)."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--gpus", "1", "--steps", str(args.cpu_steps),
           "--warmup", "1", "--workload", args.workload, "--dump-factors", dump_path]
    for k in ("size", "rank", "order"):
        if getattr(args, k) is not None:
            cmd += ["--" + k, str(getattr(args, k))]
    env = clean_cpu_env()
    env.pop("RANK", None)
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    for ln in reversed(out.stdout.strip().splitlines()):
        if ln.startswith("{"):
            return json.loads(ln)
    raise RuntimeError("cpu baseline produced no line: %s" % out.stderr[-500:])


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


def measure_fp64_peak(torch):
    """cuBLAS DGEMM 8192^3 in this process, on this GPU, at this moment: the FP64 roofline denominator (the same protocol
    as MEASURED_PEAKS.json's bf16 entry: best of 10, CUDA events).  B200 has no separate FP64 tensor rate: DMMA and DFMA
    share one pipe (profiles/r01_dmma_microbench.log), so the library's DGEMM is the attainable peak."""
    n = 8192
    a = torch.rand(n, n, dtype=torch.float64, device="cuda")
    b = torch.rand(n, n, dtype=torch.float64, device="cuda")
    c = torch.empty(n, n, dtype=torch.float64, device="cuda")
    for _ in range(2):
        torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b, out=c)
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    del a, b, c
    torch.cuda.empty_cache()
    return 2.0 * n ** 3 / (best * 1e-3) / 1e12


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
    torch.cuda.set_device(local_rank)
    if nranks > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    ppx = importlib.import_module("pairwise-perturbation_b200")
    H = importlib.import_module("pairwise-perturbation_b200.host_api")
    kat = importlib.import_module("pairwise-perturbation_b200.kat")
    lib = ppx.load_library()

    lens, R, tol_init, tag = workload_of(args)
    N = len(lens)
    if nranks > lens[0]:
        raise SystemExit("workload %s: the leading mode (%d) is shorter than the number of GPUs" % (args.workload, lens[0]))
    s0 = lens[0]
    hbm_peak = 6543.1
    try:
        hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        hbm_src = "MEASURED_PEAKS.json"
    except Exception:
        hbm_src = "fallback 6543.1 GB/s (MEASURED_PEAKS.json absent)"
    fp64_peak = measure_fp64_peak(torch)  # before the big allocations; every rank measures its own GPU

    world = H.World(local_rank, solver=0, use_graph=True, workspace_bytes=1 << 30)
    b, e = ppx.shard_range(s0, nranks, rank)
    if nranks > 1:
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt = torch.tensor(list(ppx.comm_unique_id()), dtype=torch.uint8, device="cuda")
        dist.broadcast(idt, 0)
        world.comm_init(bytes(idt.cpu().tolist()), nranks, rank, 0, s0, b, e)
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

    def gather_ranks(x):
        if dist is None:
            return [x]
        t_all = [torch.zeros(1, dtype=torch.float64, device="cuda") for _ in range(nranks)]
        dist.all_gather(t_all, torch.tensor([x], dtype=torch.float64, device="cuda"))
        return [float(t.item()) for t in t_all]

    def seeded_matrix(seed, mode):
        """s x R uniform [0,1) factor (full), sliced to the local rows for the sharded mode 0; returns (device, host-full)."""
        m = H.Matrix(world, lens[mode], R)
        m.fill(seed, mode)
        full = m.numpy()
        if mode != 0 or nranks == 1:
            return m, full
        m.free()
        return H.Tensor.from_numpy(world, np.ascontiguousarray(full[b:e]), matrix=True), full

    # synthetic tensor 'r' (test_ALS.cxx:275-286): V = [[W_true]], local slab of mode 0
    Wt, A_host = zip(*[seeded_matrix(1, i) for i in range(N)])
    lens_local = (rows0,) + tuple(lens[1:])
    V = H.Tensor(world, lens_local)
    H.build_V(world, V, list(Wt))
    for w in Wt:
        w.free()
    W, W_host = zip(*[seeded_matrix(2, i) for i in range(N)])
    W, G = list(W), [seeded_matrix(3, i)[0] for i in range(N)]
    F = [H.Matrix(world, w.lens[0], R) for w in W]
    W_start = [w.numpy() for w in W]
    world.sync()
    vsq = V.norm2() ** 2
    if dist is not None:
        t = torch.tensor([vsq], dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        vsq = float(t.item())
    vnorm = vsq ** 0.5

    def reset_W():
        for w, h in zip(W, W_start):
            w.write(h)

    evs = [C.c_void_p() for _ in range(2)]
    for ev in evs:
        lib.ppx_event_create(world.ctx_handle(), C.byref(ev))
    ms = C.c_float(0)

    def timed(fn, reps):
        lib.ppx_event_record(world.ctx_handle(), evs[0])
        for _ in range(reps):
            fn()
        lib.ppx_event_record(world.ctx_handle(), evs[1])
        lib.ppx_event_elapsed_ms(world.ctx_handle(), evs[0], evs[1], C.byref(ms))
        return float(ms.value) / reps

    # ---- device-resident sweeps: value ---------------------------------------------------------------------------
    warm = max(args.warmup, 3)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        sampler.wait_first_sample()
    H.cp_dt_sweeps(world, V, W, G, warm)
    barrier()
    sampler.mark_begin()
    launches0 = world.launch_count()
    lib.ppx_event_record(world.ctx_handle(), evs[0])
    H.cp_dt_sweeps(world, V, W, G, args.steps)
    lib.ppx_event_record(world.ctx_handle(), evs[1])
    lib.ppx_event_elapsed_ms(world.ctx_handle(), evs[0], evs[1], C.byref(ms))
    barrier()
    sampler.mark_end()
    launches = world.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms_total = max_over_ranks(float(ms.value))
    ms_per_step = ms_total / args.steps
    per_rank_ms = gather_ranks(float(ms.value) / args.steps)
    value = 1e3 / ms_per_step

    # ---- parity probe at the full size of this run (closed forms, kat.py) -----------------------------------------
    parity = {}
    try:
        reset_W()
        Cg = kat.cross_grams(list(A_host), list(W_host))
        rows = (b, e) if nranks > 1 else None
        Ms = H.cp_dt_mttkrps(world, V, W)
        per_mode = []
        for i in range(N):
            per_mode.append(kat.max_rel_err(Ms[i].numpy(), kat.mttkrp(A_host, Cg, i, rows if i == 0 else None)))
        for m in Ms:
            m.free()
        parity["dt_tree_mttkrp_max_rel_err_per_mode"] = [max_over_ranks(x) for x in per_mode]
        parity["what"] = ("V = [[A]] exactly, so MTTKRP_i(W) = A_i . hadamard_{m != i}(A_m^T W_m) and P^(i,j) have closed "
                          "forms (pairwise-perturbation_b200/kat.py); max |gpu - closed form| / max |closed form| over this "
                          "rank's rows, max over the %d ranks, at the full size of the run" % nranks)
        if not args.no_pp:
            ops = H.PPOperators(world, V, W)
            seq = "".join(chr(97 + i) for i in range(N))
            worst_pair = worst_single = 0.0
            for i in range(N):
                ref = kat.mttkrp(A_host, Cg, i, rows if i == 0 else None)
                worst_single = max(worst_single, kat.max_rel_err(ops.get(seq.replace(seq[i], ""), ref.shape), ref))
                for j in range(i + 1, N):
                    ref = kat.pair_operator(A_host, Cg, i, j, rows if i == 0 else None)
                    key = seq.replace(seq[i], "").replace(seq[j], "")
                    worst_pair = max(worst_pair, kat.max_rel_err(ops.get(key, ref.shape), ref))
            ops.free()
            world.trim()
            parity["pp_pair_operators_max_rel_err"] = max_over_ranks(worst_pair)
            parity["pp_single_operators_max_rel_err"] = max_over_ranks(worst_single)
        parity["max_rel_err"] = max([v for k, v in parity.items() if k.endswith("max_rel_err")]
                                    + parity["dt_tree_mttkrp_max_rel_err_per_mode"])
        parity["tolerance"] = 1e-12
        parity["ok"] = parity["max_rel_err"] < 1e-12
    except Exception as exc:  # noqa: BLE001
        parity["error"] = "%s: %s" % (type(exc).__name__, exc)

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
    k1_ms_local = timed(k1, 5 if 8 * float(np.prod(lens_local)) > 1e9 else 50)
    k1_per_rank = gather_ranks(k1_ms_local)
    k1_ms = max(k1_per_rank)
    out1.free()
    P_local = float(np.prod(lens_local))
    Kc = float(np.prod(lens_local[x_first:]))
    k1_flops = 2.0 * P_local * R
    k1_bytes = 8.0 * (P_local + P_local / Kc * R + sum(lens_local[x_first:]) * R)
    fp64_all = gather_ranks(fp64_peak)
    ridge = fp64_peak * 1e3 / hbm_peak  # flop per byte
    ai = k1_flops / k1_bytes
    tensor_bound = ai >= ridge
    traffic, traffic_src = None, None
    try:  # DRAM bytes per launch of the same kernel from an `ncu --set full` capture committed under profiles/
        prof = json.load(open(os.path.join(ROOT, "profiles", "k1_ncu_summary.json")))
        if nranks == 1 and list(lens) + [R] == prof.get("lens_rank"):
            traffic, traffic_src = prof.get("dram_bytes_per_launch"), prof.get("source")
    except Exception:
        pass
    ach_tf = k1_flops / (k1_ms * 1e-3) / 1e12
    ach_gbs = k1_bytes / (k1_ms * 1e-3) / 1e9
    roofline = {"kernel": "first dimension-tree contraction via ppx_ttm_multi (K1: ttm_tma_kernel / ttm_first_kernel / "
                          "ttm_stream_*_kernel by shape; modes %d..%d at once, R=%d)" % (x_first, N - 1, R),
                "bound": "tensor" if tensor_bound else "hbm",
                "achieved": ach_tf if tensor_bound else ach_gbs, "peak": fp64_peak if tensor_bound else hbm_peak,
                "unit": "TFLOP/s" if tensor_bound else "GB/s",
                "frac": (ach_tf / fp64_peak) if tensor_bound else (ach_gbs / hbm_peak),
                "traffic": traffic, "traffic_source": traffic_src,
                "algorithmic_flops": k1_flops, "algorithmic_bytes": k1_bytes, "arithmetic_intensity": ai,
                "ridge_flop_per_byte": ridge, "ms": k1_ms, "tflops": ach_tf, "hbm_gbs": ach_gbs,
                "peak_source": ("FP64: cuBLAS DGEMM 8192^3 (torch.matmul, best of 10, CUDA events) measured in this process "
                                "before the run = %.2f TFLOP/s on rank 0%s; MEASURED_PEAKS.json carries no FP64 entry. HBM: %s"
                                % (fp64_all[0], "" if nranks == 1 else " (all ranks: %s)" % ", ".join("%.2f" % x for x in fp64_all),
                                   hbm_src)),
                "fp64_peak_tflops_measured": fp64_all, "hbm_peak_gbs": hbm_peak,
                "share_of_step": (2 if N >= 4 else 1) * k1_ms / ms_per_step, "ms_per_rank": k1_per_rank}

    # ---- the collective alone: one s x R all-reduce (N-1 of them + one R x R per sweep) ----------------------------
    comm = None
    if nranks > 1:
        buf = H.Matrix(world, max(lens[1:]), R)
        bp = (C.c_void_p * 1)(buf.data_ptr())
        bn = (C.c_int64 * 1)(max(lens[1:]) * R)

        def ar():
            rc = lib.ppx_allreduce_packed(world.ctx_handle(), bp, bn, 1)
            assert rc == 0

        for _ in range(5):
            ar()
        barrier()
        ar_ms = max_over_ranks(timed(ar, 50))
        buf.free()
        comm = {"allreduce_sxR_us": 1e3 * ar_ms, "bytes": 8 * max(lens[1:]) * R, "per_sweep": "%d of s x R + 1 of R x R" % (N - 1),
                "path": ("one kernel over NVLink peer memory (stage, flag every peer, sum the staged copies in rank order)"
                         if lib.ppx_comm_p2p(world.ctx_handle()) else "NCCL"),
                "timing": "50 back-to-back ppx_allreduce_packed calls on the engine's stream, CUDA events, max over ranks"}

    # ---- PP phase (pp_bench protocol): operator build + approximate sweeps; the -pp 1 mixed run -------------------
    pp = None
    if not args.no_pp:
        reset_W()
        H.cp_dt_sweeps(world, V, W, G, 2)  # PP starts from an ALS iterate, as alsCP_PP does
        barrier()
        H.cp_pp_phase_timed(world, V, W, G, 1)  # warm-up: first-use kernel loading and the level-1 allocations
        reset_W()
        H.cp_dt_sweeps(world, V, W, G, 2)
        barrier()
        ms_build, ms_pp = H.cp_pp_phase_timed(world, V, W, G, args.pp_sweeps)
        ms_build, ms_pp = max_over_ranks(ms_build), max_over_ranks(ms_pp)
        world.trim()  # the level-1 tensors of the build go back to the driver
        sweep_bytes = 8.0 * sum(lens_local[i] * lens_local[j] * R + lens_local[j] * R
                                for i in range(N) for j in range(N) if i != j) + 24.0 * sum(lens_local) * R
        pp = {"operator_build_ms": ms_build, "approx_sweep_ms": ms_pp / args.pp_sweeps,
              "approx_sweeps_per_s": 1e3 * args.pp_sweeps / ms_pp, "sweeps_timed": args.pp_sweeps,
              "approx_sweep_algorithmic_bytes": sweep_bytes,
              "approx_sweep_frac_hbm": sweep_bytes / (ms_pp / args.pp_sweeps) / 1e6 / hbm_peak}
        # K3 (PP correction of one mode) alone, as HBM GB/s
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
            k3_ms = timed(k3, 20)
            k3_bytes = 8.0 * (sum(s_i * lens_local[j] * R + lens_local[j] * R for j in others) + 2 * s_i * R)
            pp["k3_pp_correct"] = {"us": 1e3 * k3_ms, "algorithmic_bytes": k3_bytes, "gbs": k3_bytes / k3_ms / 1e6,
                                   "frac_hbm": k3_bytes / k3_ms / 1e6 / hbm_peak,
                                   "timing": "20 back-to-back launches, CUDA events (launch ramp included)"}
            # the solve kernels of one mode update alone ("solve kernels as HBM GB/s"): Hadamard of the Grams
            # + R x R inverse (one CTA), then W = M S^-1 with gradient and dW (row tiles); both are latency-size work --
            # 8 (5 s R + 2 R^2) bytes -- and are reported as what they are
            Gs = [H.Matrix(world, R, R) for _ in range(N)]
            for j_, g_ in enumerate(Gs):
                rc_ = lib.ppx_gram(world.ctx_handle(), C.c_void_p(W[j_].data_ptr()), lens_local[j_], lens_local[j_], R,
                                   C.c_void_p(g_.data_ptr()))
                assert rc_ == 0
            gp = (C.c_void_p * N)(*[g_.data_ptr() for g_ in Gs])
            S_, Si_ = H.Matrix(world, R, R), H.Matrix(world, R, R)
            Wc, Gc, Dc = H.Matrix(world, s_i, R), H.Matrix(world, s_i, R), H.Matrix(world, s_i, R)

            def inv():
                rc_ = lib.ppx_spd_inverse_g(world.ctx_handle(), gp, N, i_mode, 0.0, R, 0, C.c_void_p(S_.data_ptr()),
                                            C.c_void_p(Si_.data_ptr()))
                assert rc_ == 0, lib.ppx_last_error(world.ctx_handle())

            def app():
                rc_ = lib.ppx_solve_apply(world.ctx_handle(), C.c_void_p(Mo.data_ptr()), C.c_void_p(S_.data_ptr()),
                                          C.c_void_p(Si_.data_ptr()), C.c_void_p(Wc.data_ptr()), s_i, R,
                                          C.c_void_p(W[i_mode].data_ptr()), 1.0, C.c_void_p(Gc.data_ptr()),
                                          C.c_void_p(Dc.data_ptr()))
                assert rc_ == 0, lib.ppx_last_error(world.ctx_handle())

            for _ in range(5):
                inv()
                app()
            inv_ms, app_ms = timed(inv, 50), timed(app, 50)
            solve_bytes = 8.0 * (5 * s_i * R + 2 * R * R)
            pp["solve"] = {"inverse_us": 1e3 * inv_ms, "apply_us": 1e3 * app_ms, "algorithmic_bytes": solve_bytes,
                           "gbs": solve_bytes / (inv_ms + app_ms) / 1e6, "frac_hbm": solve_bytes / (inv_ms + app_ms) / 1e6 / hbm_peak,
                           "note": "latency bound: one CTA factorises the R x R matrix (R dependent elimination steps), the "
                                   "apply kernel moves 0.6 MB; 50 back-to-back launches each, CUDA events"}
            for t_ in ops + [M0, Mo, S_, Si_, Wc, Gc, Dc] + Gs:
                t_.free()
        except Exception as exc:  # noqa: BLE001
            pp.setdefault("k3_pp_correct", {"error": "%s: %s" % (type(exc).__name__, exc)})
            pp["solve"] = pp.get("solve") or {"error": "%s: %s" % (type(exc).__name__, exc)}
        # the mixed run: alsCP_PP exactly as `test_ALS -pp 1 -maxiter M` drives it (als_CP.cxx:1082-1137), residual
        # evaluations skipped (the reference takes them off its clock, :189); wall clock, synchronised on both sides
        try:
            def mixed(maxiter, tol_sw):
                reset_W()
                for i_, g_ in enumerate(G):
                    g_.fill(3, i_)
                barrier()
                with H.Trace(quiet=True, skip_residual=True) as t_:
                    t0_ = time.perf_counter()
                    H.alsCP_PP(world, V, W, G, F, 1e-10 * vnorm, tol_sw, maxiter, resprint=10)
                    barrier()
                    sec_ = max_over_ranks(time.perf_counter() - t0_)
                n_dt = sum(1 for k, _ in t_.sweeps if k == 0)
                n_pp = sum(1 for k, _ in t_.sweeps if k == 1)
                n_build = sum(1 for k, _ in t_.sweeps if k == 2)
                world.trim()
                return {"maxiter": maxiter, "pp_res_tol": tol_sw, "dt_sweeps": n_dt, "pp_sweeps": n_pp,
                        "operator_builds": n_build, "seconds": sec_, "sweeps_per_s": (n_dt + n_pp) / sec_,
                        "switches": [[int(k), int(it)] for k, it in t_.events]}

            mixed(min(args.pp_maxiter, 12), 0.5)  # warm-up of both phases
            pp["mixed_run"] = mixed(args.pp_maxiter, tol_init)
            pp["mixed_run"]["call"] = "alsCP_PP(V, W, grad_W, F, 1e-10*||V||, pp_res_tol, timelimit, maxiter, 0, 1, csv, resprint=10, false, dw)"
            # the default switching tolerance may never be met inside maxiter sweeps from a random start (then the run is
            # all exact sweeps); a second run with a loose tolerance shows the DT/PP mix at work
            pp["mixed_run_loose_tol"] = mixed(args.pp_maxiter, 0.2)
        except Exception as exc:  # noqa: BLE001
            pp["mixed_run"] = {"error": "%s: %s" % (type(exc).__name__, exc)}
        pp["note"] = ("approximate sweep = one CUDA graph (N x [correct, inverse || , solve+grad+dW, Gram] + Normalize + norms)"
                      if nranks == 1 else "approximate sweep on %d GPUs: see pp.graph" % nranks)

    # ---- e2e: the driver call with host buffers --------------------------------------------------------------------
    reset_W()
    pin = [torch.from_numpy(h.ravel(order="F").copy()).pin_memory() for h in W_start]
    pin_g = [torch.zeros(h.size, dtype=torch.float64).pin_memory() for h in W_start]
    pin_out = [torch.empty(h.size, dtype=torch.float64).pin_memory() for h in W_start]
    hlib = H.load_host_library()

    def e2e_step():
        # pinned host -> device on the engine's stream (stream order makes them visible to the sweep), the driver call,
        # device -> pinned host, ONE synchronise at the end of the step: the results are on the host when it returns
        for t_, w in zip(pin, W):
            hlib.ppxh_tensor_write_async(w.h, C.c_void_p(t_.data_ptr()))
        for t_, g in zip(pin_g, G):
            hlib.ppxh_tensor_write_async(g.h, C.c_void_p(t_.data_ptr()))
        H.alsCP_DT(world, V, W, G, F, 0.0, 0, lam=0.0, resprint=1 << 30, bench=True)  # exactly one sweep
        for t_, w in zip(pin_out, W):
            hlib.ppxh_tensor_read_async(w.h, C.c_void_p(t_.data_ptr()))
        for t_, g in zip(pin_g, G):
            hlib.ppxh_tensor_read_async(g.h, C.c_void_p(t_.data_ptr()))
        hlib.ppxh_world_sync(world.h)

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

    # ---- CPU baseline (rank 0, one GPU): the reference arm, bounded; also a full-size factor comparison ---------------
    cpu = None
    if rank == 0 and nranks == 1 and not args.no_cpu_baseline:
        try:
            with tempfile.TemporaryDirectory() as td:
                dump = os.path.join(td, "cpu_factors.npz")
                ref_line = cpu_baseline_leg(args, dump)
                cpu = ref_line["cpu_baseline"]
                cpu["ms_per_sweep"] = ref_line["ms_per_step"]
                z = np.load(dump)
                k_sweeps = int(z["sweeps"])
                if int(z["rows0"]) == lens[0]:  # the port swept the same full tensor: compare the factors sweep for sweep
                    reset_W()
                    H.cp_dt_sweeps(world, V, W, G, k_sweeps)
                    errs = [kat.max_rel_err(W[i].numpy(), z["arr_%d" % i]) for i in range(N)]
                    parity["factors_vs_cpu_port"] = {"sweeps": k_sweeps, "max_rel_err_per_mode": errs, "tolerance": 1e-8,
                                                     "ok": max(errs) < 1e-8,
                                                     "what": "factors after %d exact sweeps from the same seeded start, CUDA "
                                                             "path vs the NumPy port, at the full size" % k_sweeps}
        except Exception as exc:  # noqa: BLE001
            cpu = {"error": "%s: %s" % (type(exc).__name__, exc)}

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
            # roofline objects: the first TTM of the HOOI tree (V x_2 W_2: 800^3 -> 800 x 800 x 40) and the unfolding Gram
            # of the HOSVD (800 x 640000 SYRK), each timed alone with CUDA events on this rank's slab
            lt = (te - tb,) + (ts,) * (tn - 1)
            ltc = (C.c_int64 * tn)(*lt)
            out_t = H.Tensor(world, lt[:2] + (tr,))

            def ttm1():
                rc_ = lib.ppx_ttm(world.ctx_handle(), C.c_void_p(Vt.data_ptr()), ltc, tn, 2, C.c_void_p(Wt_[2].data_ptr()),
                                  ts, tr, C.c_void_p(out_t.data_ptr()))
                assert rc_ == 0, lib.ppx_last_error(world.ctx_handle())

            for _ in range(3):
                ttm1()
            t_ttm = max_over_ranks(timed(ttm1, 10))
            Pt = float(np.prod(lt))
            fl, by = 2.0 * Pt * tr, 8.0 * (Pt + Pt / ts * tr + ts * tr)
            tucker["roofline_first_ttm"] = {"kernel": "K1 (ttm_tma_kernel) as Tucker TTM, mode 2 of %s, Q=%d" % (list(lt), tr),
                                            "bound": "tensor", "ms": t_ttm, "achieved": fl / t_ttm / 1e9, "peak": fp64_peak,
                                            "unit": "TFLOP/s", "frac": fl / t_ttm / 1e9 / fp64_peak,
                                            "algorithmic_flops": fl, "algorithmic_bytes": by, "hbm_gbs": by / t_ttm / 1e6}
            out_t.free()
            mtm = H.Matrix(world, ts, ts)

            def syrk():
                rc_ = lib.ppx_unfold_gram(world.ctx_handle(), C.c_void_p(Vt.data_ptr()), ltc, tn, 1, C.c_void_p(mtm.data_ptr()))
                assert rc_ == 0, lib.ppx_last_error(world.ctx_handle())

            for _ in range(2):
                syrk()
            t_syrk = max_over_ranks(timed(syrk, 5))
            fl_useful = Pt * (ts + 1)  # lower triangle incl. diagonal: ts(ts+1)/2 entries x 2 flop x (P/ts) terms
            tucker["roofline_unfold_gram"] = {"kernel": "gram_dmma_kernel (DMMA SYRK of the mode-1 unfolding, %d x %d)" % (ts, int(Pt / ts)),
                                              "bound": "tensor", "ms": t_syrk, "achieved": fl_useful / t_syrk / 1e9,
                                              "peak": fp64_peak, "unit": "TFLOP/s", "frac": fl_useful / t_syrk / 1e9 / fp64_peak,
                                              "algorithmic_flops": fl_useful, "algorithmic_bytes": 8.0 * (Pt + ts * ts),
                                              "note": "algorithmic flops = the symmetric half; the kernel also executes the "
                                                      "upper parts of its diagonal 128-blocks"}
            mtm.free()
        except Exception as exc:  # noqa: BLE001
            tucker = (tucker or {})
            tucker["error"] = "%s: %s" % (type(exc).__name__, exc)

    if rank == 0:
        line = {
            "metric": "ALS-DT sweeps/s (CP order-%d FP64)" % N, "value": value, "unit": "sweeps/s",
            "n_gpus": nranks, "steps": args.steps, "warmup": warm, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_of(lens, R, tag, nranks),
            "ms_per_step_per_rank": per_rank_ms, "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
            "roofline": roofline, "parity_probe": parity, "cpu_baseline": cpu, "comm": comm, "pp": pp, "tucker": tucker,
        }
        print(json.dumps(line), flush=True)
    world.close()
    if dist is not None:
        dist.destroy_process_group()


import ctypes as C  # noqa: E402

if __name__ == "__main__":
    main()
