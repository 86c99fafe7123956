#!/bin/bash
# one GPU round trip: parity tests, quick timing, per-kernel launch times (uniform + clustered)
timeout 800 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python scripts/canvas_only.py 2>&1 | tail -3
for mode in uniform clustered; do
  python scripts/prof_step.py vod $mode 16 30000 2 > gpurun_out/plain.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$mode.csv python scripts/prof_step.py vod $mode 16 30000 2 > /dev/null 2>&1
  echo "== $mode"; grep "k_" gpurun_out/launches_$mode.csv | awk -F'","' '{print $5, $NF}' | tail -2
done
