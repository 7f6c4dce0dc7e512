#!/bin/bash
# final 1-GPU record: full GPU suite, bench at the headline workload (with the CPU arm), the other workloads, reference arm
mkdir -p gpurun_out
T=${1:-r02f}
timeout 900 python -m pytest tests -q -x -m gpu 2>&1 | tail -3 | tee gpurun_out/${T}_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/${T}_bench_1gpu.log 2> gpurun_out/${T}_bench_1gpu.err; echo "bench rc=$?"
for w in cfg1 cfg4 cfg5; do timeout 400 python bench.py --workload $w --no-cpu-baseline --no-tucker > gpurun_out/${T}_bench_1gpu_$w.log 2> gpurun_out/${T}_bench_1gpu_$w.err; echo "bench $w rc=$?"; done
timeout 200 python tools/time_k7.py > gpurun_out/${T}_k7.json 2>/dev/null
timeout 100 python tools/time_inverse.py > gpurun_out/${T}_inverse_us.json 2>/dev/null
timeout 100 python tools/time_mttv3.py > gpurun_out/${T}_mttv3.json 2>/dev/null
python - <<PY
import json
for w in ["", "_cfg1", "_cfg4", "_cfg5"]:
    for l in open("gpurun_out/${T}_bench_1gpu%s.log" % w):
        if l.startswith("{"):
            d=json.loads(l); pp=d.get("pp") or {}
            print(w or "cfg2", d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["bound"], pp.get("operator_build_ms"), pp.get("approx_sweep_ms"), (pp.get("mixed_run") or {}).get("sweeps_per_s"), (pp.get("mixed_run_loose_tol") or {}).get("sweeps_per_s"), (d.get("cpu_baseline") or {}).get("value"), d["parity_probe"]["max_rel_err"], (d.get("tucker") or {}).get("ms_per_sweep"))
PY
