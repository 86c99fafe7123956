#!/bin/bash
# role ablations of k_pillars (library built with -DHGSF_EXPERIMENT -> variants/exp.so; WRONG results for dbg != 0, timing only):
#   HGSF_DBG bit 0 = no pillar arithmetic (canvas role alone), bit 1 = no tile writes (pillar role alone), bits 8-10 = canvas-preferring warps per CTA
export HGSF_LIB=hgsfusion_b200/variants/exp.so
for w in "vod clustered 16 30000" "vod uniform 16 30000"; do
  for d in ${@:-0 1 2 3}; do echo -n "DBG=$d "; HGSF_DBG=$d timeout 120 python scripts/r2_step.py $w 10 50 2>&1 | tail -1; done
done
