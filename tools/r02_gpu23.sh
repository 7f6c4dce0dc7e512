#!/bin/bash
# peer-memory all-reduce on N GPUs: parity (incl. the direct all-reduce check), bench with and without it
mkdir -p gpurun_out
T=${1:-r02z}
N=$(nvidia-smi -L | wc -l)
export PPX_COMM_VERBOSE=1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/check_multi_gpu.py > gpurun_out/${T}_multigpu_parity_${N}.log 2>&1; echo "check_multi_gpu rc=$?"
grep -E "MISMATCH|PARITY|all-reduce|ppx:" gpurun_out/${T}_multigpu_parity_${N}.log | cut -c1-300 | head -12
for mode in ${MODES:-p2p nccl}; do
  if [ $mode = nccl ]; then export PPX_NO_P2P=1; fi
  timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --no-tucker > gpurun_out/${T}_bench_${N}gpu_$mode.log 2> gpurun_out/${T}_bench_${N}gpu_$mode.err; echo "bench $mode rc=$?"
  grep "ppx:" gpurun_out/${T}_bench_${N}gpu_$mode.err | head -3
  python - <<PY
import json
d=json.loads(open('gpurun_out/${T}_bench_${N}gpu_$mode.log').read().strip().splitlines()[-1])
print('$mode', 'value', d['value'], 'e2e', d['e2e']['value'], 'comm', d['comm']['allreduce_sxR_us'], d['comm'].get('path'), 'pp sweep', d['pp']['approx_sweep_ms'], 'build', d['pp']['operator_build_ms'], 'probe', d['parity_probe']['max_rel_err'], 'mixed', d['pp']['mixed_run']['sweeps_per_s'], d['pp']['mixed_run_loose_tol']['sweeps_per_s'])
PY
done
