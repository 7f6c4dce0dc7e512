#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02m}
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "residual or known" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${T}_pytest.log
timeout 600 python -m pytest tests/test_drivers_gpu.py -m gpu -q -x -k "known_answer or full_size" >> gpurun_out/${T}_pytest.log 2>&1; echo "pytest2 rc=$?"; tail -3 gpurun_out/${T}_pytest.log
for D in 0 1 2 3; do echo "dbg=$D"; PPX_K7_DBG=$D timeout 200 python tools/time_k7.py 2>/dev/null | python -c "import sys,json; d=json.load(sys.stdin); print([ (x['lens'][0], round(x['residual_ms'],2), round(x['build_V_ms_first_call'],1)) for x in d])"; done | tee gpurun_out/${T}_k7_dbg.txt
