#!/bin/bash
# one GPU round trip of round 2: parity tests, step / kernel times on the four workloads, k_front phase stamps
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2_tests.log
for w in "vod clustered 16 30000" "vod uniform 16 30000" "tj4d clustered 16 30000" "stress clustered 16 200000" "vod clustered 16 2000"; do
  timeout 300 python scripts/r2_step.py $w 2>&1 | tail -1
done | tee gpurun_out/r2_steps.log
if [ -f hgsfusion_b200/variants/phase.so ]; then HGSF_LIB=hgsfusion_b200/variants/phase.so timeout 300 python scripts/phase_times.py 2>&1 | tail -2 | tee gpurun_out/r2_phases.log; fi
