#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02x}
for R in 10 40 50 64; do echo R=$R; ./tools/inv_bench_bin $R | grep -E "sweep rep [4]|max"; ./tools/inv_bench_np $R | grep -E "rep [4]" | cut -c1-30; done
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x -m gpu -k "spd_inverse or solve or normalize or residual" 2>&1 | tail -5
timeout 120 python tools/time_inverse.py > gpurun_out/${T}_inverse_sweep_us.json 2> gpurun_out/${T}_inv.err; echo rc=$?; cat gpurun_out/${T}_inverse_sweep_us.json
timeout 200 python tools/time_k7.py > gpurun_out/${T}_k7.json 2>gpurun_out/${T}_k7.err; echo rc=$?; cat gpurun_out/${T}_k7.json
