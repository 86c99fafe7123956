// subm_conv.cuh -- parameter blocks and launchers of the pillar-list consumer: the submanifold
// 3x3 convolutions of SpMiddlePillarEncoder18.conv1 (pillarnet_modules/pcnres18.py:82-106,108-188,212-215) evaluated on
// the pillar list the reader produced instead of on a dense canvas.
#pragma once
#include "common.cuh"

namespace hgsf {

struct SubmNeighborParams {
    const int *bev;          // [B, H, W] pillar id per cell, -1 none (hgsf_pillarnet_indices)
    const int *pillars;      // [M, 3] (b, y, x)
    const int *m_dev;        // optional device row count (counts[0] of hgsf_pillarnet_indices); null = M
    long long M;
    int B, H, W;             // of the INPUT table
    int stride;              // 1: submanifold (pillars = the input's own); 2: pillars = the output cells of a stride-2 convolution
    int *nbr;                // [M, 9] input pillar id at (y * stride + ky - 1, x * stride + kx - 1), tap = ky * 3 + kx; -1 none
};

struct Stride2CandidateParams {
    const int *pillars;      // [M, 3] (b, y, x) of the input, frames contiguous
    const int *m_dev;
    long long M;
    int B, Ho, Wo;
    float *cand;             // [4 M, 3] centres of the output cells each input pillar reaches, (-10, -10) for unused slots
    int *cnt;                // [B] zeroed; receives 4 x pillars per frame (xyz_batch_cnt of the candidates)
};

struct SubmConvParams {
    const float *in;         // [M, Cin]
    const int *nbr;          // [M, 9]
    const int *m_dev;        // optional device row count
    long long M;
    int Cin, Cout;
    const float *W;          // see layout
    int layout;              // 0: [Cout, 3, 3, Cin] (spconv 2.x), 1: [3, 3, Cin, Cout] (spconv 1.x)
    const float *bias;       // [Cout] or null
    const float *bn_w, *bn_b, *bn_m, *bn_v;   // BatchNorm1d eval or null
    float eps;
    const float *residual;   // [M, Cout] or null, added after the BatchNorm
    int relu;
    float *out;              // [M, Cout]
};

int launch_subm_neighbors(const SubmNeighborParams &q, cudaStream_t stream);
int launch_subm_conv(const SubmConvParams &q, cudaStream_t stream);
int launch_stride2_candidates(const Stride2CandidateParams &q, cudaStream_t stream);

}  // namespace hgsf
