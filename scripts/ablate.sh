#!/bin/bash
# k_emit ablations (library built with HGSF_NVCC_EXTRA=-DHGSF_EXPERIMENT -> hgsfusion_b200/variants/exp.so): HGSF_DBG bit 0 =
# skip the unit phase (decorate + Linear + BN + max), bit 1 = skip the per-tile TMA store of occupied tiles, bit 5 (32) = rows
# taken as already ordered and the mean as known (WRONG results, timing only), bit 6 (64) = additionally no permutation table.
# Step time = k_front + k_emit.
export HGSF_LIB=hgsfusion_b200/variants/exp.so
for d in ${@:-0 1 2 3 32 96}; do echo "== HGSF_DBG=$d"; HGSF_DBG=$d timeout 120 python scripts/quick_gpu.py 2>&1 | tail -2; done
