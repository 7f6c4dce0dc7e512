#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02l}
for D in 0 1 2 3; do echo "dbg=$D"; PPX_K7_DBG=$D timeout 200 python tools/time_k7.py 2>/dev/null | python -c "import sys,json; d=json.load(sys.stdin); print([ (x['lens'][0], round(x['residual_ms'],2)) for x in d])"; done | tee gpurun_out/${T}_k7_dbg.txt
