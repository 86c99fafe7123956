// pillar_path.cuh -- parameter block shared by the pillar-path kernels and their launchers.
#pragma once
#include "common.cuh"

namespace hgsf {

// Division by a run-time constant without the hardware's division routine: q = (n + mulhi(n, m)) >> s, exact for
// 0 <= n < 2^31 and 1 <= d < 2^31 (Granlund-Montgomery round-up method with a 33-bit multiplier whose top bit is the `n +`).
struct FastDiv {
    uint32_t m;
    int s;
};
inline FastDiv make_fastdiv(uint32_t d) {
    FastDiv f;
    int l = 0;
    while ((1ull << l) < d) ++l;
    f.m = (uint32_t)((((1ull << l) - d) << 32) / d + 1ull);
    f.s = l;
    return f;
}
__host__ __device__ __forceinline__ uint32_t fastdiv(uint32_t n, const FastDiv f) {
#ifdef __CUDA_ARCH__
    return (n + __umulhi(n, f.m)) >> f.s;
#else
    return (uint32_t)(((uint64_t)n + (((uint64_t)n * f.m) >> 32)) >> f.s);
#endif
}

struct PathParams {
    // ---- points (hgsf_points) ----
    const float *pts;
    int n, stride, xyz_col, F, batch_col;
    const int32_t *frame_offsets_in;   // caller's, or nullptr
    int B;
    // ---- geometry ----
    float rmin[3], vsize[3], voff[3];
    int nx, ny, nz;
    int nxp;                           // row pitch of the cell table: nx rounded up to 32, so that a canvas tile (32 cells of
                                       // one BEV row) is one aligned 32-entry group of the table and tile t owns cells 32t..32t+31
    int cells;                         // nz * ny * nxp: table entries per frame
    int tiles_per_row;                 // nxp / 32
    FastDiv div_cells, div_plane, div_nxp, div_tpr, div_ny;   // by cells, ny*nxp, nxp, tiles_per_row, ny
    // ---- limits ----
    int P, max_voxels;
    // ---- workspace ----
    uint32_t *ticket;                  // [128]: word 0 = k_pillars' run ticket, word 32 = heavy tiles listed by k_front, word 64 = CTAs of
                                       // k_pillars that have finished (each on its own 128-byte line); reset by k_front
    uint64_t *state;                   // == magic: the cell table is known to be all zero (the consumer kernel cleaned up after itself)
    uint64_t magic;                    // geometry / layout dependent
    uint32_t *scan_desc;               // [3 * 2048] per-CTA slice totals of k_front's scans (first points; points, occupied cells)
    int32_t *frame_raw_base;           // [B+1] raw (pre max_voxels) pillar id at each frame start
    uint32_t *cell_tag, *cell_cnt, *cell_start;   // [B*cells] each (padded), back to back
    size_t table_bytes;                // bytes of cell_tag + cell_cnt (the part that must be zero when k_front starts)
    uint4 *tile_rec;                   // [B*cells/32] per 32-cell tile: {CSR row of its first point, points in the tile,
                                       //   pillars (in cell order) before the tile, occupancy mask of its 32 cells}
    uint32_t *heavy_list;              // [B*cells/32] tiles with more than heavy_pts points (ticket[32] of them), in no particular order
    int heavy_pts;
    int32_t *frame_offsets;            // [B+1] (aliases frame_offsets_in when given)
    int32_t *key;                      // [n] cell key of each point, -1 = outside the grid
    uint32_t *arrival;                 // [n] unordered arrival rank of the point inside its cell
    float *sorted_rows;                // [n, RW] F features + point index of each point, grouped by cell (CSR in table order)
    int4 *pil;                         // [n] the pillars in CELL order (frames concatenated): {cell key, raw pillar id in first-seen
                                       //   order, arrivals, CSR start}
    int RW;
    int pslice, cslice;                // points / cells per k_front CTA (multiples of 32)
    int canvas_vec;                    // the canvas is written with 16-byte stores (nx % 4 == 0, aligned base)
    int flags;                         // hgsf_points.flags (HGSF_POINTS_*)
    int32_t *cutoff;                   // [B] HGSF_POINTS_SPCONV1_BREAK: index of the frame's first dropped point (INT_MAX: none)
    int dbg;                           // experiment switches (HGSF_DBG), 0 in normal use
    // ---- PFN ----
    const float *W, *bias, *bn_w, *bn_b, *bn_m, *bn_v;
    float eps;
    int Cin, C;
    // ---- train mode (launch_pillar_path_train): BatchNorm1d on the statistics of the batch ----
    double *stats;                     // out: train_ops.cu's statistics layout (Sx [C], Sxx [C], T [C][Cin], s [Cin]), read by the backward
    double *stats_S;                   // [16 * Cin] scratch behind it: second moments S = sum f f^T and s = sum f (zeroed by k_front)
    float *batch_mean, *batch_var;     // [64] out: what the forward normalises with (biased variance)
    float *run_mean, *run_var;         // [64] updated in place as torch.nn.BatchNorm1d does (may be null)
    float momentum;
    // ---- outputs ----
    int32_t *coords, *num, *num_pillars;
    float *voxels, *feats, *canvas;
};

struct WorkspaceLayout {
    size_t off_ticket, off_state, off_desc, off_raw_base, off_cutoff, off_table, off_tile_rec, off_heavy;
    size_t off_frame_offsets, off_key, off_arrival, off_sorted_rows, off_pil;
    size_t total, cell_array_bytes;
    int RW, nxp;
    int64_t cells;         // per frame, padded pitch
};

constexpr int SCAN_TILE = 1024;
constexpr int MAX_FRONT_CTAS = 2048;

inline WorkspaceLayout workspace_layout(int64_t n, int B, int nx, int ny, int nz, int F) {
    WorkspaceLayout w{};
    size_t o = 0;
    w.nxp = (nx + 31) / 32 * 32;
    w.cells = (int64_t)nz * ny * w.nxp;
    w.RW = (F + 1 + 3) / 4 * 4;   // F features + the point index
    w.off_ticket = o;      o = align_up(o + 512, 256);                         // three counters, one 128-byte line each
    w.off_state = o;       o = align_up(o + 256, 256);
    w.off_desc = o;        o = align_up(o + sizeof(uint32_t) * 3 * MAX_FRONT_CTAS, 256);
    w.off_raw_base = o;    o = align_up(o + sizeof(int32_t) * (size_t)(B + 1), 256);
    w.off_cutoff = o;      o = align_up(o + sizeof(int32_t) * (size_t)(B + 1), 256);
    w.cell_array_bytes = align_up(sizeof(uint32_t) * ((size_t)B * (size_t)w.cells + 32), 256);
    w.off_table = o;       o = o + 3 * w.cell_array_bytes;
    w.off_tile_rec = o;    o = align_up(o + sizeof(uint4) * ((size_t)B * (size_t)w.cells / 32 + 1), 256);
    w.off_heavy = o;       o = align_up(o + sizeof(uint32_t) * ((size_t)B * (size_t)w.cells / 32 + 1), 256);
    w.off_frame_offsets = o; o = align_up(o + sizeof(int32_t) * (size_t)(B + 1), 256);
    w.off_key = o;         o = align_up(o + sizeof(int32_t) * (size_t)n, 256);
    w.off_arrival = o;     o = align_up(o + sizeof(uint32_t) * (size_t)n, 256);
    w.off_sorted_rows = o; o = align_up(o + sizeof(float) * (size_t)n * (size_t)w.RW, 256);
    w.off_pil = o;         o = align_up(o + sizeof(int4) * (size_t)n, 256);
    w.total = o;
    return w;
}

// launchers (pillar_path.cu); return a cudaError_t-or-HGSF status and count launches
int launch_pillar_path(const PathParams &p, bool with_pfn, bool abs_xyz, bool dist, cudaStream_t stream, int *launches);
// train-mode forward: k_front, the statistics pass, the fused pass on the batch statistics (3 launches, no host round trip).
// HGSF_ERR_UNSUPPORTED outside the tile-major kernel's domain (BatchNorm, canvas, P <= 32, spconv-2 rule, nx % 4 == 0).
int launch_pillar_path_train(const PathParams &p, bool abs_xyz, bool dist, cudaStream_t stream, int *launches);

int emit_timing_begin(int capacity);
int emit_timing_collect(float *ms, int n);

}  // namespace hgsf
