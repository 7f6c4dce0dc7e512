#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02r}
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/${T}_pytest.log
timeout 300 python tools/time_k7.py > gpurun_out/${T}_k7.json 2>/dev/null; cat gpurun_out/${T}_k7.json
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench.log 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/${T}_bench.err
