#!/bin/bash
# round-2 one-GPU pass after the kernel work: GPU tests, K7 timing (DMMA vs DFMA), Tucker launch list, bench line
mkdir -p gpurun_out
T=${1:-r02c}
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${T}_pytest.log
tail -8 gpurun_out/${T}_pytest.log
timeout 300 python tools/time_k7.py > gpurun_out/${T}_k7_dmma.json 2> gpurun_out/${T}_k7.err; echo "k7 rc=$?"; cat gpurun_out/${T}_k7_dmma.json
PPX_K7_DFMA=1 timeout 300 python tools/time_k7.py > gpurun_out/${T}_k7_dfma.json 2>> gpurun_out/${T}_k7.err; echo "k7 dfma rc=$?"; cat gpurun_out/${T}_k7_dfma.json
timeout 300 python tools/bench_tucker.py > gpurun_out/${T}_tucker.log 2>&1; echo "tucker rc=$?"; tail -3 gpurun_out/${T}_tucker.log
PPX_EIG_RR_ONESIDED=1 timeout 300 python tools/bench_tucker.py > gpurun_out/${T}_tucker_onesided.log 2>&1; tail -3 gpurun_out/${T}_tucker_onesided.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${T}_launches_tucker.csv python tools/bench_tucker.py --sweeps 3 > gpurun_out/${T}_ncu_tucker.log 2>&1; echo "ncu tucker rc=$?"
timeout 900 python bench.py --steps 10 --warmup 3 --cpu-steps 1 > gpurun_out/${T}_bench.log 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
tail -c 300 gpurun_out/${T}_bench.err
