#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02p}
N=$(nvidia-smi -L | wc -l)
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/check_multi_gpu.py > gpurun_out/${T}_multigpu_parity_${N}.log 2>&1; echo "check_multi_gpu rc=$?"
grep -E "MISMATCH|PARITY|differs|hosvd proj" gpurun_out/${T}_multigpu_parity_${N}.log | cut -c1-1500 | head -20
