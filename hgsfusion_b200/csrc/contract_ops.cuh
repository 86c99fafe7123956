// contract_ops.cuh -- parameter blocks of the stand-alone PillarVFE / PointPillarScatter ops.
#pragma once
#include "common.cuh"
#include "pfn.cuh"

namespace hgsf {

struct VfeParams {
    const float *voxels;      // [M, P, F]
    const void *coords;       // [M, 4] (b, z, y, x) fp32 or int32
    const void *num;          // [M] fp32 or int32
    int coords_float, num_float;
    long long M;
    int P, F, C;
    float vsize[3], voff[3];
    PfnArgs pfn;
    float *out;               // [M, C]
    // stacked PFN (NUM_FILTERS with two entries, pillar_vfe.py:63-74): `pfn` is then the first layer (C = its out_channels =
    // NUM_FILTERS[0] / 2) and pfn1 the last one, in_channels 2*C -> C1
    PfnArgs pfn1;
    int C1;
    // train_ops.cu only: when set, the number of pillars is read on the device (min(M, *M_dev)): M is then the capacity of the
    // buffers and nothing has to come back to the host between the forward and the backward
    const int32_t *M_dev;
};

struct ScatterParams {
    const float *feats;       // [M, C]
    const void *coords;       // [M, 4]
    int coords_float;
    int coord_cols;           // 4 (b, z, y, x) or 3 (b, y, x; int32 only)
    long long M;
    int C, B, ny, nx;
    long long plane;          // nz*ny*nx (nz == 1)
    unsigned *map;            // [B*plane] workspace
    float *canvas;            // [B, C, ny, nx]
};

int launch_vfe(const VfeParams &q, bool abs_xyz, bool dist, cudaStream_t stream);
int launch_vfe_stacked(const VfeParams &q, bool abs_xyz, bool dist, cudaStream_t stream);
int launch_scatter(const ScatterParams &q, cudaStream_t stream, int *launches);

// shared host helpers (pillar_path.cu)
int make_canvas_map(CUtensorMap *map, float *canvas, int B, int C, int ny, int nx, int box_c);
int sm_count();

}  // namespace hgsf
