#!/bin/bash
# k_emit ablations (library built with HGSF_NVCC_EXTRA=-DHGSF_EXPERIMENT): HGSF_DBG bit 0 = skip the unit phase
# (decorate + Linear + BN + max), bit 1 = skip the per-tile TMA store of occupied tiles.  Step time = k_front + k_emit.
for d in 0 1 2 3; do echo "== HGSF_DBG=$d"; HGSF_DBG=$d timeout 120 python scripts/quick_gpu.py 2>&1 | tail -2; done
