#!/bin/bash
# one GPU: tests, K7 timing + ncu, inverse timing A/B, PP sweep timing via bench (no cpu baseline, no tucker)
mkdir -p gpurun_out
T=${1:-r02d}
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${T}_pytest.log
tail -8 gpurun_out/${T}_pytest.log
timeout 300 python tools/time_k7.py > gpurun_out/${T}_k7_dmma.json 2> gpurun_out/${T}_k7.err; echo "k7 rc=$?"; cat gpurun_out/${T}_k7_dmma.json
timeout 300 python tools/time_inverse.py > gpurun_out/${T}_inverse_blocked.json 2> gpurun_out/${T}_inv.err; cat gpurun_out/${T}_inverse_blocked.json
PPX_INV_COLUMN=1 timeout 300 python tools/time_inverse.py > gpurun_out/${T}_inverse_column.json 2>> gpurun_out/${T}_inv.err; cat gpurun_out/${T}_inverse_column.json
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-tucker > gpurun_out/${T}_bench.log 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
tail -c 300 gpurun_out/${T}_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cp_reconstruct_dmma_kernel -c 2 -o gpurun_out/${T}_k7_full -f python tools/time_k7.py > gpurun_out/${T}_ncu_k7.log 2>&1; echo "ncu k7 rc=$?"
