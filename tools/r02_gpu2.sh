#!/bin/bash
# round-2 multi-GPU pass (N = number of visible GPUs): GPU test suite (incl. the SPMD command-line tests), parity of the
# sharded drivers against the oracle, bench line at N GPUs
mkdir -p gpurun_out
T=${1:-r02b}
N=$(nvidia-smi -L | wc -l)
echo "GPUs: $N"
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${T}_pytest.log
tail -8 gpurun_out/${T}_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/check_multi_gpu.py > gpurun_out/${T}_multigpu_parity.log 2>&1; echo "check_multi_gpu rc=$?"
tail -4 gpurun_out/${T}_multigpu_parity.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${T}_bench_${N}gpu.log 2> gpurun_out/${T}_bench_${N}gpu.err; echo "bench rc=$?"
tail -c 400 gpurun_out/${T}_bench_${N}gpu.err
PPX_BOOT_PORT=29620 timeout 600 python -m torch.distributed.run --no-python --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 pairwise-perturbation_b200/pp_bench -model CP -tensor r -dim 4 -size 300 -rank 50 -maxiter 3 -filename gpurun_out/${T}_pp_bench_cfg2_${N}gpu.csv > gpurun_out/${T}_pp_bench_cfg2_${N}gpu.log 2>&1; echo "pp_bench rc=$?"
tail -12 gpurun_out/${T}_pp_bench_cfg2_${N}gpu.log
