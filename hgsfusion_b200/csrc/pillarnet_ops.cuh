// pillarnet_ops.cuh -- parameter block and launchers of the Path B (PillarNet reader) ops.
#pragma once
#include "common.cuh"

namespace hgsf {

struct PillarNetParams {
    const float *xyz;        // [N, 3] coordinates relative to the range minimum (dynamic_pillar_encoder.py:83-85)
    const int *cnt;          // [B] points per frame
    long long N;
    int B, H, W;             // H = Ny, W = Nx (pillar_utils.py:102-103)
    float bev_size;
    int *bev;                // [B, H, W] pillar id per cell, -1 = none  (pillar_bev_indices)
    int *pillars;            // [cap, 3] (b, y, x) in raster order
    int *pairs;              // [N] pillar id of each point or -1 (indice_pairs, K = 1); may be null
    int *point_idx;          // [N] compacted: points that have a pillar, input order
    int *pillar_idx;         // [N] ... and their pillar ids
    int *counts;             // [2] M, L
    int *key;                // workspace [N]
    uint32_t *partial;       // workspace [8192] per-CTA slice totals (cells, points)
};

struct SplitEncodeParams {
    const float *points;     // [L, 1 + Fin] collated points, column 0 = frame index (dataset.py:237-244)
    const int *order;        // optional [L] row order (stable by frame); null = input order
    long long L;
    int Fin, Fout, n_split;  // split: Fout >= 3 + 2 n_split + 2 (VoD 17 -> 29, n = 12; TJ4D 18 -> 31, n = 13)
    int B, mode;
    float pc_min[3];
    float *xyz;              // [L, 3] relative
    float *feat;             // [L, Fout]
    int *cnt;                // [B] xyz_batch_cnt
    int *info;               // [2] {flags, rows kept}
};

struct ReaderParams {
    const float *xyz;        // [N, 3] relative coordinates
    const float *feat;       // [N, Cf]
    int Cf;
    const int *point_idx, *pillar_idx;   // [L] grouped points and their pillars (hgsf_pillarnet_indices)
    long long L, M;
    const int *pillars;      // [M, 3] (b, y, x)
    float bev_size, z_center;
    const float *W, *bn_w, *bn_b, *bn_m, *bn_v;   // Linear [32, Cf+6] (no bias), BatchNorm1d eval
    float eps;
    float *out;              // [M, 32]
};

int launch_reader_fused(const ReaderParams &q, cudaStream_t stream, int *launches);
int launch_split_encode(const SplitEncodeParams &q, cudaStream_t stream);
int launch_pillarnet_indices(const PillarNetParams &q, cudaStream_t stream);
int launch_gather(long long L, int C, const int *idx, const float *f, float *out, cudaStream_t s);
int launch_gather_grad(long long L, int C, const int *idx, const float *gout, float *gin, cudaStream_t s);
int launch_scatter_max(int C, long long L, long long M, const int *index, const float *src, int *arg, float *out, cudaStream_t s,
                       int *launches);
int launch_scatter_max_grad(int C, long long M, const int *arg, const float *gout, float *gsrc, cudaStream_t s);

}  // namespace hgsf
