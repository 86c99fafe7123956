#!/bin/bash
# builds experiment variants of the library (different -D switches) next to the default one; run on the GPU box with
#   HGSF_LIB=hgsfusion_b200/variants/<name>.so python scripts/quick_gpu.py
# usage: scripts/variants.sh name1:"-DX=1" name2:"-DX=2" ...
mkdir -p hgsfusion_b200/variants
cp hgsfusion_b200/libhgsfusion_b200.so /tmp/hgsf_default.so
for spec in "$@"; do
  name="${spec%%:*}"; flags="${spec#*:}"
  HGSF_NVCC_EXTRA="$flags" python -c "from hgsfusion_b200 import build; build.build(force=True)" && cp hgsfusion_b200/libhgsfusion_b200.so hgsfusion_b200/variants/$name.so && echo "built $name ($flags)"
done
cp /tmp/hgsf_default.so hgsfusion_b200/libhgsfusion_b200.so
