#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02q}
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "shards_sums" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; grep -E "^FAILED|passed|failed" gpurun_out/${T}_pytest.log | head -40
