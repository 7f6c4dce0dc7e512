#!/bin/bash
# Runs the BASELINE.json configurations through the reference-facing command line (test_ALS) on one GPU and keeps
# the tail of each log: sweeps, switching markers, residuals and wall time.   usage: tools/run_configs.sh [outdir]
out=${1:-gpurun_out/configs}
mkdir -p "$out"
cd "$(dirname "$0")/../pairwise-perturbation_b200" || exit 1
run() {  # name, args...
  name=$1; shift
  echo "=== $name: test_ALS $*"
  t0=$(date +%s%N)
  ./test_ALS "$@" -filename "../$out/$name.csv" > "../$out/$name.log" 2>&1
  rc=$?
  echo "wall $(( ($(date +%s%N) - t0) / 1000000 )) ms (rc=$rc)" >> "../$out/$name.log"
  grep -E "starts from|experiment took|wall|error|rror" "../$out/$name.log" | tail -8
  tail -3 "../$out/$name.csv"
}
run cfg1_cp_n3_s200_r10_dt   -model CP -tensor r -dim 3 -size 200 -rank 10 -pp 0 -maxiter 50
run cfg2_cp_n4_s300_r50_dt   -model CP -tensor r -dim 4 -size 300 -rank 50 -pp 0 -maxiter 30
run cfg2_cp_n4_s300_r50_pp   -model CP -tensor r -dim 4 -size 300 -rank 50 -pp 1 -maxiter 30
run cfg3_tucker_n3_s800_r40_dt -model Tucker -tensor r2 -dim 3 -size 800 -rank 40 -pp 0 -maxiter 10
run cfg3_tucker_n3_s800_r40_pp -model Tucker -tensor r2 -dim 3 -size 800 -rank 40 -pp 1 -maxiter 10
run cfg4_cp_n6_s40_r10_pp    -model CP -tensor r -dim 6 -size 40 -rank 10 -pp 1 -maxiter 50
run cfg5_coil_shape_r10_pp   -model CP -tensor r -lens 3,128,128,7200 -rank 10 -pp 1 -pp_res_tol 0.05 -maxiter 250
