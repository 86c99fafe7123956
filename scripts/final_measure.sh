#!/bin/bash
# round-end measurement of the final build: bench lines (4 workloads + reference arm), ncu launch lists and full captures
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_clustered.json 2> gpurun_out/bench_clustered.err
python bench.py --mode uniform --no-cpu-baseline > gpurun_out/bench_uniform.json 2> gpurun_out/bench_uniform.err
python bench.py --config tj4d --no-cpu-baseline > gpurun_out/bench_tj4d.json 2> gpurun_out/bench_tj4d.err
python bench.py --config stress --points 200000 --ring 4 --steps 50 --no-cpu-baseline > gpurun_out/bench_stress200k.json 2> gpurun_out/bench_stress.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
for mode in clustered uniform; do
  python scripts/prof_step.py vod $mode 16 30000 2 > gpurun_out/plain_$mode.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$mode.csv python scripts/prof_step.py vod $mode 16 30000 2 > /dev/null 2>&1
  ncu --set full --clock-control none --import-source on -k regex:'k_front|k_emit' -s 2 -c 2 -f -o gpurun_out/full_$mode python scripts/prof_step.py vod $mode 16 30000 2 > gpurun_out/ncu_full_$mode.log 2>&1
done
grep -h '"value"' gpurun_out/bench_*.json | cut -c1-200
