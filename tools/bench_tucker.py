"""Tucker HOOI at BASELINE configs[2] (order-3, s=800, ranks 40, tensor 'r2') on N GPUs: HOSVD time and seconds per
HOOI sweep through hosvd / alsTucker_DT of the C++ host layer (mode-0 sharded tensor, replicated factors).

    python tools/bench_tucker.py [--size 800 --rank 40 --sweeps 6]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_tucker.py
"""
import argparse
import importlib
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=800)
ap.add_argument("--rank", type=int, default=40)
ap.add_argument("--order", type=int, default=3)
ap.add_argument("--sweeps", type=int, default=6)
args = ap.parse_args()
rank, nranks, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
ppx = importlib.import_module("pairwise-perturbation_b200")
H = importlib.import_module("pairwise-perturbation_b200.host_api")
dist = None
if nranks > 1:
    import torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
world = H.World(local, workspace_bytes=2 << 30)
s, R, N = args.size, args.rank, args.order
b, e = ppx.shard_range(s, nranks, rank)
if nranks > 1:
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt = torch.tensor(list(ppx.comm_unique_id()), dtype=torch.uint8, device="cuda")
    dist.broadcast(idt, 0)
    world.comm_init(bytes(idt.cpu().tolist()), nranks, rank, 0, s, b, e)
# tensor 'r2': uniform [0.5,1); the local slab is rows b..e of mode 0 of the global tensor (index = i0 + s*(...)):
# generate the full tensor once per rank and keep the slab (4 GB at s=800; a one-off outside the timed region)
full = H.Tensor(world, (s,) * N)
full.fill(1, 100, 0.5, 1.0)
if nranks > 1:
    V = H.Tensor.from_numpy(world, np.ascontiguousarray(full.numpy()[b:e]))
    full.free()
else:
    V = full
W = [H.Matrix(world, s, R) for _ in range(N)]
core = H.Tensor(world, (R,) * N)


def sync():
    world.sync()
    if dist is not None:
        dist.barrier()


H.hosvd(world, V, core, W, [R] * N)  # warm-up: first-use kernel loading and pool allocations
sync()
t0 = time.perf_counter()
H.hosvd(world, V, core, W, [R] * N)
sync()
t_hosvd = time.perf_counter() - t0
with H.Trace(quiet=True, skip_residual=True):
    H.alsTucker_DT(world, V, core, W, 0.0, 1, resprint=1 << 30, bench=False)  # warm-up: 2 sweeps
    sync()
    t0 = time.perf_counter()
    H.alsTucker_DT(world, V, core, W, 0.0, args.sweeps - 1, resprint=1 << 30, bench=False)
    sync()
    t_sweeps = time.perf_counter() - t0
if rank == 0:
    print(json.dumps({"workload": "Tucker HOOI order-%d s=%d ranks %d, tensor r2" % (N, s, R), "n_gpus": nranks,
                      "hosvd_s": t_hosvd, "sweeps_timed": args.sweeps, "ms_per_sweep": 1e3 * t_sweeps / args.sweeps}))
world.close()
if dist is not None:
    dist.destroy_process_group()
