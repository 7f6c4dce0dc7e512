#!/bin/bash
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -m gpu -k "mttv3" 2>&1 | tail -3
for e in A=1 PPX_NO_MTTV3=1; do env $e timeout 300 python bench.py --no-cpu-baseline --no-tucker --steps 3 --warmup 3 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$e', 'build', d['pp']['operator_build_ms'], 'sweep', d['pp']['approx_sweep_ms'], 'probe', d['parity_probe']['max_rel_err'])"; done
