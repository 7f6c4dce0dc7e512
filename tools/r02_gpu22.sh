#!/bin/bash
# PP sweep A/B: split correction x inverse isolation
run() { env "$@" timeout 300 python bench.py --no-cpu-baseline --no-tucker --steps 3 --warmup 3 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$*', 'approx_sweep_ms', d['pp']['approx_sweep_ms'], 'value', d['value'], 'inv', d['pp']['solve']['inverse_us'])"; }
run A=1
run PPX_INV_NO_ISOLATE=1
run PPX_PP_SPLIT=1
run PPX_PP_SPLIT=1 PPX_INV_NO_ISOLATE=1
