#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02t}
N=$(nvidia-smi -L | wc -l)
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/check_multi_gpu.py > gpurun_out/${T}_multigpu_parity_${N}.log 2>&1; echo "check_multi_gpu rc=$?"
grep -E "MISMATCH|PARITY|plain" gpurun_out/${T}_multigpu_parity_${N}.log | cut -c1-300 | head -12
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 --no-tucker > gpurun_out/${T}_bench_${N}gpu.log 2> gpurun_out/${T}_bench_${N}gpu.err; echo "bench rc=$?"
tail -c 300 gpurun_out/${T}_bench_${N}gpu.err
python - <<PY
import json
d=json.loads(open('gpurun_out/${T}_bench_${N}gpu.log').read().strip().splitlines()[-1])
print(d['value'], d['pp'].get('solve'), d['pp'].get('k3_pp_correct'))
PY
