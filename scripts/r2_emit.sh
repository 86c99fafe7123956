#!/bin/bash
# tile-major (k_emit) against pillar-major (k_pillars) consumer: parity tests, then step times on the workloads
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2_tests.log
for w in "vod clustered 16 30000" "vod uniform 16 30000" "tj4d clustered 16 30000" "stress clustered 16 200000" "vod clustered 16 2000"; do
  echo "tile $w: $(timeout 300 python scripts/r2_step.py $w 2>&1 | tail -1)"
  echo "tile chunk2 $w: $(HGSF_TILE_CHUNK=2 timeout 300 python scripts/r2_step.py $w 2>&1 | tail -1)"
  echo "tile chunk1 $w: $(HGSF_TILE_CHUNK=1 timeout 300 python scripts/r2_step.py $w 2>&1 | tail -1)"
  echo "pillar $w: $(HGSF_CONSUMER=p timeout 300 python scripts/r2_step.py $w 2>&1 | tail -1)"
done | tee gpurun_out/r2_emit_steps.log
