#!/bin/bash
# A/B of library builds on the step time: scripts/r2_ab.sh <variant ...>  (the in-tree build is always measured as "head");
# the parity tests run on the in-tree build and on the first variant
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r2_tests.log
if [ -n "$1" ]; then HGSF_LIB=hgsfusion_b200/variants/$1.so timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests_$1.log 2>&1; echo "tests $1 rc=$?"; tail -2 gpurun_out/r2_tests_$1.log; fi
for rep in 1 2; do
for w in "vod clustered 16 30000" "vod uniform 16 30000" "tj4d clustered 16 30000" "stress clustered 16 200000" "vod clustered 16 2000"; do
  echo "head $w: $(timeout 300 python scripts/r2_step.py $w 2>&1 | tail -1 | cut -c1-140)"
  for v in "$@"; do
    echo "$v $w: $(HGSF_LIB=hgsfusion_b200/variants/$v.so timeout 300 python scripts/r2_step.py $w 2>&1 | tail -1 | cut -c1-140)"
  done
done
done | tee gpurun_out/r2_ab.log
