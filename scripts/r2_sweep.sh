#!/bin/bash
# sweep of the heavy-tile threshold of the tile-major consumer (HGSF_HEAVY_PTS overrides max(56, 3 x average))
mkdir -p gpurun_out
for w in "vod clustered 16 30000" "tj4d clustered 16 30000" "stress clustered 16 200000" "vod uniform 16 30000"; do
  for hp in def 40 28 20 12; do
    if [ $hp = def ]; then unset HGSF_HEAVY_PTS; else export HGSF_HEAVY_PTS=$hp; fi
    echo "pts=$hp $w: $(timeout 300 python scripts/r2_step.py $w 2>&1 | tail -1 | cut -c40-110)"
  done
done | tee gpurun_out/r2_sweep.log
