#!/bin/bash
# round-2 final measurement pass on one B200: parity tests, bench lines of the four workloads, the reference arm, a launch list and
# ncu --set full captures (after the same commands ran clean without ncu)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02_gpu_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r02_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 200 --warmup 10 > gpurun_out/r02_bench_1gpu_vod_clustered.json 2> gpurun_out/b.err || tail -3 gpurun_out/b.err
for c in "vod uniform 30000" "tj4d clustered 30000" "stress clustered 200000"; do set -- $c
  python bench.py --config $1 --mode $2 --points $3 --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_1gpu_$1_$2.json 2> gpurun_out/b.err || tail -3 gpurun_out/b.err
done
python bench.py --impl reference --steps 50 --warmup 3 > gpurun_out/r02_bench_reference_arm.json 2>> gpurun_out/b.err
python scripts/prof_step.py vod clustered 16 30000 8 > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches_vod_clustered.csv python scripts/prof_step.py vod clustered 16 30000 8 > /dev/null 2>&1
# ncu --set full captures: scripts/r2_ncu.sh <tag> <config> <mode> <B> <n>, two per gpurun call (the reports are ~15 MB each and
# gpurun_out/ brings back 64 MiB at most)
python scripts/bench_train.py vod clustered 16 30000 > gpurun_out/r02_train_step_vod_clustered.json 2>> gpurun_out/b.err; cat gpurun_out/r02_train_step_vod_clustered.json
