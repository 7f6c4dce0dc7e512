#!/bin/bash
# one GPU: tests, K7 timing + ncu capture of the residual kernel, full bench line
mkdir -p gpurun_out
T=${1:-r02h}
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${T}_pytest.log
tail -5 gpurun_out/${T}_pytest.log
timeout 300 python tools/time_k7.py > gpurun_out/${T}_k7_dmma.json 2> gpurun_out/${T}_k7.err; echo "k7 rc=$?"; cat gpurun_out/${T}_k7_dmma.json
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench.log 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
tail -c 300 gpurun_out/${T}_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cp_reconstruct_dmma_kernel --launch-skip 2 -c 1 -o gpurun_out/${T}_k7_full -f python tools/time_k7.py > gpurun_out/${T}_ncu_k7.log 2>&1; echo "ncu k7 rc=$?"
