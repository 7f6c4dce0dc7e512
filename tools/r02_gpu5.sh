#!/bin/bash
# one GPU: tests, K7 timing + ncu of the residual kernel, bench lines for the other BASELINE workloads
mkdir -p gpurun_out
T=${1:-r02e}
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${T}_pytest.log
tail -6 gpurun_out/${T}_pytest.log
timeout 300 python tools/time_k7.py > gpurun_out/${T}_k7_dmma.json 2> gpurun_out/${T}_k7.err; echo "k7 rc=$?"; cat gpurun_out/${T}_k7_dmma.json
for W in cfg1 cfg4 cfg5; do
  timeout 600 python bench.py --workload $W --steps 20 --warmup 5 --no-tucker > gpurun_out/${T}_bench_$W.log 2> gpurun_out/${T}_bench_$W.err; echo "bench $W rc=$?"; tail -c 300 gpurun_out/${T}_bench_$W.err
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cp_reconstruct_dmma_kernel --launch-skip 2 -c 1 -o gpurun_out/${T}_k7_full -f python tools/time_k7.py > gpurun_out/${T}_ncu_k7.log 2>&1; echo "ncu k7 rc=$?"
