// hybrid_points.cuh -- parameters of the hybrid point assembly (hybrid_points.cu)
#pragma once
#include "common.cuh"

namespace hgsf {

struct HybridParams {
    const float *real, *gt, *virt;                 // [sum Nr, Fr], [sum Ng, W], [sum Nv, W]
    const int32_t *real_off, *gt_off, *virt_off;   // [B+1] each (device)
    const float *calib;                            // [B, 26] or nullptr (no FOV filter)
    int Fr, W, B;                                  // W = 0: sweep only (USE_VIRTUAL_POINTS False)
    long long n;                                   // candidate rows = sum Nr + sum Ng + sum Nv
    int no_dup, mask_range;
    double dup_threshold;
    double range_xy[4];                            // xmin, ymin, xmax, ymax (inclusive)
    unsigned char *flags;                          // workspace
    int *chunk_counts;                             // workspace
    float *out;                                    // [cap, 1 + (W ? W + 2 : Fr)]
    int32_t *frame_offsets_out;                    // [B+1]
};

size_t hybrid_workspace_bytes(long long n);
int launch_hybrid(HybridParams q, void *ws, size_t ws_bytes, cudaStream_t s, int *launches);

}  // namespace hgsf
