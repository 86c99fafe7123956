#!/bin/bash
# ncu --set full of one k_front + one k_emit launch (after warm-up launches) on a workload: scripts/r2_ncu.sh <tag> <config> <mode> <B> <n>
tag=$1; shift
mkdir -p gpurun_out
python scripts/prof_step.py "$@" 3 > gpurun_out/ncu_${tag}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/ncu_${tag}_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on --launch-skip 4 --launch-count 2 -f -o gpurun_out/ncu_${tag} \
    python scripts/prof_step.py "$@" 3 > gpurun_out/ncu_${tag}.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/ncu_${tag}.ncu-rep
