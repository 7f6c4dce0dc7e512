#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02y}
timeout 900 python -m pytest tests -q -x -m gpu 2>&1 | tail -5
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/${T}_bench_1gpu.log 2> gpurun_out/${T}_bench_1gpu.err; echo bench rc=$?
python - <<PY
import json
for l in open("gpurun_out/${T}_bench_1gpu.log"):
    if l.startswith("{"):
        d=json.loads(l); pp=d["pp"]
        print(d["value"], d["e2e"]["value"], pp["operator_build_ms"], pp["approx_sweep_ms"], pp["solve"]["inverse_us"], pp["solve"]["apply_us"], pp["k3_pp_correct"]["us"], pp["mixed_run"]["sweeps_per_s"], pp["mixed_run_loose_tol"]["sweeps_per_s"], d["parity_probe"]["max_rel_err"])
PY
PPX_PP_NO_SPLIT=1 timeout 300 python bench.py --no-cpu-baseline --no-tucker --steps 3 --warmup 3 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('no split: approx_sweep_ms', d['pp']['approx_sweep_ms'])"
PPX_K7_FORCE_DMMA=1 timeout 200 python tools/time_k7.py > gpurun_out/${T}_k7_force_dmma.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/${T}_k7_force_dmma.json')); print('forced DMMA', [(x['lens'],x['R'],round(x['residual_ms'],2)) for x in d])"
