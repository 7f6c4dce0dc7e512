#!/bin/bash
# round-2 GPU pass on one B200: GPU test suite, smoke, bench line (+ reference arm), ncu launch list and one full capture of K1
mkdir -p gpurun_out
T=${1:-r02a}
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${T}_pytest.log
tail -5 gpurun_out/${T}_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench.log 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/${T}_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_ref.log 2>&1; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-tucker > gpurun_out/${T}_ncu_bench.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ttm_tma_kernel -c 2 -o gpurun_out/${T}_k1_full -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-tucker --no-pp > gpurun_out/${T}_ncu_k1.log 2>&1; echo "ncu k1 rc=$?"
ls -la gpurun_out | tail -12
