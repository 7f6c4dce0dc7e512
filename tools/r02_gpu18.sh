#!/bin/bash
# K7 A/B: warps per CTA x column blocks per pass (libraries built with -DPPX_RD_THREADS / -DPPX_RD_NB under tools/k7_variants)
mkdir -p gpurun_out
T=${1:-r02v}
echo "base"; timeout 200 python tools/time_k7.py > gpurun_out/${T}_k7_base.json 2>gpurun_out/${T}_k7_base.err; echo rc=$?
for v in $(ls tools/k7_variants | sed "s/libppx_k7_//; s/.so//"); do
  echo "$v"; PPX_LIB=$PWD/tools/k7_variants/libppx_k7_$v.so timeout 200 python tools/time_k7.py > gpurun_out/${T}_k7_$v.json 2>gpurun_out/${T}_k7_$v.err; echo rc=$?
done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${T}_k7_*.json")):
    try:
        d=json.load(open(f)); print(f, [round(x["residual_ms"],2) for x in d], d[0]["residual"], d[0]["residual_at_truth_over_norm"])
    except Exception as e: print(f, "bad", e)
PY
