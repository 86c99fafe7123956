#!/bin/bash
# profiles/r02_ncu_k_front_k_emit_<workload>.txt from gpurun_out/ncu_<tag>.ncu-rep: scripts/r2_ncu_summaries.sh tag:workload ...
SHA=$(git rev-parse --short HEAD)
for t in "$@"; do tag=${t%%:*}; name=${t#*:}
  { echo "# ncu --set full --clock-control none --import-source on, one launch each of k_front and k_emit after 4 warm-up launches"
    echo "# command: python scripts/prof_step.py ${name/_/ } 16 ... (scripts/r2_ncu.sh), commit $SHA, read with scripts/ncu_summary.py + scripts/ncu_lines.py"
    python scripts/ncu_summary.py gpurun_out/ncu_$tag.ncu-rep 30; echo; echo "## k_emit: stall samples by source line"
    python scripts/ncu_lines.py gpurun_out/ncu_$tag.ncu-rep k_emit 30; echo; echo "## k_front: stall samples by source line"
    python scripts/ncu_lines.py gpurun_out/ncu_$tag.ncu-rep k_front 20; } > profiles/r02_ncu_k_front_k_emit_$name.txt 2>&1
done
