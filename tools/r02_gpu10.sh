#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02i}
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "residual or known" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${T}_pytest.log
timeout 600 python -m pytest tests/test_drivers_gpu.py -m gpu -q -x -k "known_answer or full_size" >> gpurun_out/${T}_pytest.log 2>&1; echo "pytest2 rc=$?"; tail -3 gpurun_out/${T}_pytest.log
timeout 300 python tools/time_k7.py > gpurun_out/${T}_k7_dmma.json 2> gpurun_out/${T}_k7.err; echo "k7 rc=$?"; cat gpurun_out/${T}_k7_dmma.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-tucker --pp-maxiter 12 > gpurun_out/${T}_ncu_bench.log 2>&1; echo "ncu list rc=$?"
