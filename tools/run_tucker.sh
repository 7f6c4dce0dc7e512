#!/bin/bash
cd "$(dirname "$0")/../pairwise-perturbation_b200" || exit 1
./test_ALS -model Tucker -tensor r2 -dim 3 -size 800 -rank 40 -pp 0 -maxiter 10 -filename /tmp/t.csv | grep -E "iter|took" | tail -6
./test_ALS -model Tucker -tensor r2 -dim 3 -size 800 -rank 40 -pp 1 -maxiter 30 -filename /tmp/t.csv | grep -E "starts|took" | tail -8
