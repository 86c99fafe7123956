#!/bin/bash
# times the default library and the experiment builds named on the command line (hgsfusion_b200/variants/<name>.so)
for v in "" "$@"; do
  echo "== variant ${v:-default}"
  if [ -n "$v" ]; then export HGSF_LIB=hgsfusion_b200/variants/$v.so; fi
  timeout 200 python scripts/quick_gpu.py 2>&1 | tail -2
  timeout 200 python scripts/quick_gpu.py 2>&1 | tail -2
done
