#!/bin/bash
# sweep inverse: parity + A/B timing against the column kernel; then the K7 warps-per-CTA A/B
mkdir -p gpurun_out
T=${1:-r02v}
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x -m gpu -k "spd_inverse or solve or normalize" 2>&1 | tail -5
timeout 120 python tools/time_inverse.py > gpurun_out/${T}_inverse_sweep_us.json 2> gpurun_out/${T}_inv.err; echo rc=$?; cat gpurun_out/${T}_inverse_sweep_us.json
PPX_INV_COLUMN=1 timeout 120 python tools/time_inverse.py > gpurun_out/${T}_inverse_column_us.json 2>> gpurun_out/${T}_inv.err; echo rc=$?; cat gpurun_out/${T}_inverse_column_us.json
bash tools/r02_gpu18.sh $T
