#!/bin/bash
# 8-GPU pass (one box): sharded parity against the oracle on every rank, bench line, the SPMD command lines at the
# BASELINE configurations (pp_bench cfg2; test_ALS Tucker cfg3 = configs[2] "on 8 x B200")
mkdir -p gpurun_out
T=${1:-r02g}
N=$(nvidia-smi -L | wc -l)
echo "GPUs: $N"; free -g | head -2
PPX_COMM_VERBOSE=1 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/check_multi_gpu.py > gpurun_out/${T}_multigpu_parity.log 2>&1; echo "check_multi_gpu rc=$?"
tail -2 gpurun_out/${T}_multigpu_parity.log
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${T}_bench_${N}gpu.log 2> gpurun_out/${T}_bench_${N}gpu.err; echo "bench rc=$?"
tail -c 300 gpurun_out/${T}_bench_${N}gpu.err
PPX_BOOT_PORT=29620 timeout 200 python -m torch.distributed.run --no-python --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 pairwise-perturbation_b200/pp_bench -model CP -tensor r -dim 4 -size 300 -rank 50 -maxiter 5 -filename gpurun_out/${T}_pp_bench_cfg2_${N}gpu.csv > gpurun_out/${T}_pp_bench_cfg2_${N}gpu.log 2>&1; echo "pp_bench rc=$?"
grep -E "step time|PP first|PP second|experiment" gpurun_out/${T}_pp_bench_cfg2_${N}gpu.log | tail -16
PPX_BOOT_PORT=29621 timeout 200 python -m torch.distributed.run --no-python --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 pairwise-perturbation_b200/test_ALS -model Tucker -tensor r2 -dim 3 -size 800 -rank 40 -pp 1 -maxiter 30 -filename gpurun_out/${T}_test_ALS_tucker_cfg3_${N}gpu.csv > gpurun_out/${T}_test_ALS_tucker_cfg3_${N}gpu.log 2>&1; echo "test_ALS tucker rc=$?"
tail -8 gpurun_out/${T}_test_ALS_tucker_cfg3_${N}gpu.log
PPX_BOOT_PORT=29622 timeout 200 python -m torch.distributed.run --no-python --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515 pairwise-perturbation_b200/test_ALS -model CP -tensor r -dim 4 -size 300 -rank 50 -pp 1 -maxiter 50 -filename gpurun_out/${T}_test_ALS_cp_cfg2_${N}gpu.csv > gpurun_out/${T}_test_ALS_cp_cfg2_${N}gpu.log 2>&1; echo "test_ALS cp rc=$?"
tail -6 gpurun_out/${T}_test_ALS_cp_cfg2_${N}gpu.log
