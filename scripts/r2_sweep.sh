#!/bin/bash
# sweep of the heavy-tile threshold / spread of the tile-major consumer
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r2_tests.log
for w in "vod clustered 16 30000" "tj4d clustered 16 30000" "stress clustered 16 200000" "vod uniform 16 30000"; do
  for sp in 0 85 100; do
    for hp in def 40 28 20 12; do
      if [ $hp = def ]; then unset HGSF_HEAVY_PTS; else export HGSF_HEAVY_PTS=$hp; fi
      echo "spread=$sp pts=$hp $w: $(HGSF_HEAVY_SPREAD=$sp timeout 300 python scripts/r2_step.py $w 2>&1 | tail -1 | cut -c40-110)"
    done
  done
done | tee gpurun_out/r2_sweep.log
