#!/bin/bash
# Builds oracle/_ref/libref_pillar_ops.so: the REFERENCE's own Path B CUDA kernels (pcdet/ops/pillar_ops/src/*_gpu.cu),
# compiled unmodified from where they lie under /root/reference, for sm_100a, plus our extern "C" shim.
# Test infrastructure only (GPU parity tests compare our kernels with these on the same B200).
# The sources include <torch/extension.h> through cuda_utils.h, so torch's headers are on the include path; nothing
# from libtorch is linked (the kernels and launchers use none of it).
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF=${REF:-/root/reference/pcdet/ops/pillar_ops/src}
OUT="$HERE/_ref"
[ -d "$REF" ] || { echo "reference sources not found at $REF (fine on the GPU box: the prebuilt .so travels)"; exit 0; }
mkdir -p "$OUT"
TORCH=$(python -c "import torch, os; print(os.path.dirname(torch.__file__))")
PYINC=$(python -c "import sysconfig; print(sysconfig.get_paths()['include'])")
FLAGS="-gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -Xcompiler -fPIC -I$TORCH/include -I$TORCH/include/torch/csrc/api/include -I$PYINC -I$REF -D_GLIBCXX_USE_CXX11_ABI=1 -w"
for f in pillar_ops_gpu group_ops_gpu scatter_ops_gpu; do
  nvcc $FLAGS -c "$REF/$f.cu" -o "$OUT/$f.o" &
done
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -Xcompiler -fPIC -c "$HERE/ref_pillar_ops_shim.cu" -o "$OUT/shim.o" &
wait
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o "$OUT/libref_pillar_ops.so" "$OUT"/pillar_ops_gpu.o "$OUT"/group_ops_gpu.o "$OUT"/scatter_ops_gpu.o "$OUT"/shim.o
rm -f "$OUT"/*.o
echo "built $OUT/libref_pillar_ops.so"
