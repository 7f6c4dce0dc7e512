#!/bin/bash
# N-GPU pass: GPU tests on this box, sharded parity against the oracle, K7 timing, bench line at N GPUs
mkdir -p gpurun_out
T=${1:-r02f}
N=$(nvidia-smi -L | wc -l)
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${T}_pytest.log
tail -6 gpurun_out/${T}_pytest.log
timeout 300 python tools/time_k7.py > gpurun_out/${T}_k7_dmma.json 2> gpurun_out/${T}_k7.err; echo "k7 rc=$?"; cat gpurun_out/${T}_k7_dmma.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/check_multi_gpu.py > gpurun_out/${T}_multigpu_parity.log 2>&1; echo "check_multi_gpu rc=$?"
tail -3 gpurun_out/${T}_multigpu_parity.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${T}_bench_${N}gpu.log 2> gpurun_out/${T}_bench_${N}gpu.err; echo "bench rc=$?"
tail -c 300 gpurun_out/${T}_bench_${N}gpu.err
