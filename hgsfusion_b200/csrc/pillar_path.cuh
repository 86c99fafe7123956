// pillar_path.cuh -- parameter block shared by the pillar-path kernels and their launchers.
#pragma once
#include "common.cuh"

namespace hgsf {

struct PathParams {
    // ---- points (hgsf_points) ----
    const float *pts;
    int n, stride, xyz_col, F, batch_col;
    const int32_t *frame_offsets_in;   // caller's, or nullptr
    int B;
    // ---- geometry ----
    float rmin[3], vsize[3], voff[3];
    int nx, ny, nz, cells;             // cells = nx*ny*nz
    // ---- limits ----
    int P, max_voxels;
    // ---- workspace ----
    uint32_t *ticket;                  // [64] zeroed by k_front: word 0 = k_pfn's chunk ticket, word 32 = k_emit's tile ticket
    uint64_t *scan_desc;               // [tiles]  zeroed
    int32_t *frame_raw_base;           // [B+1]    zeroed; raw (pre max_voxels) pillar id at each frame start
    uint32_t *cell_tag, *cell_cnt, *cell_start;   // [B*cells] each (padded to 16 entries), back to back; zeroed by k_front
    size_t table_bytes;                // bytes of the three arrays together
    int32_t *frame_offsets;            // [B+1] (aliases frame_offsets_in when given)
    int32_t *key;                      // [n] cell key of each point, -1 = outside the grid
    uint32_t *arrival;                 // [n] unordered arrival rank of the point inside its cell
    float *sorted_rows;                // [n, RW] F features + point index of each point, grouped by pillar (CSR order)
    int4 *prec;                        // [n] per raw pillar id: {CSR start, arrivals, b<<16|z, y<<16|x}
    int RW;
    int dbg;                           // experiment switches (HGSF_DBG), 0 in normal use
    // ---- PFN ----
    const float *W, *bias, *bn_w, *bn_b, *bn_m, *bn_v;
    float eps;
    int Cin, C;
    // ---- outputs ----
    int32_t *coords, *num, *num_pillars;
    float *voxels, *feats, *canvas;
};

struct WorkspaceLayout {
    size_t zero_bytes;     // leading region that must be zero at the start of every call
    size_t off_ticket, off_desc, off_raw_base, off_table;
    size_t off_frame_offsets, off_key, off_arrival, off_sorted_rows, off_prec;
    size_t total, cell_array_bytes;
    int scan_tiles, RW;
};

constexpr int SCAN_TILE = 1024;

inline WorkspaceLayout workspace_layout(int64_t n, int B, int64_t cells, int F) {
    WorkspaceLayout w{};
    size_t o = 0;
    w.scan_tiles = (int)((n + SCAN_TILE - 1) / SCAN_TILE);
    w.RW = (F + 1 + 3) / 4 * 4;   // F features + the point index
    w.off_ticket = o;      o = align_up(o + 256, 256);                        // two counters, one 128-byte line each
    w.off_desc = o;        o = align_up(o + sizeof(uint64_t) * 4096, 256);          // per-CTA slice totals of k_front's scan
    w.off_raw_base = o;    o = align_up(o + sizeof(int32_t) * (size_t)(B + 1), 256);
    w.cell_array_bytes = align_up(sizeof(uint32_t) * ((size_t)B * (size_t)cells + 16), 256);
    w.off_table = o;       o = o + 3 * w.cell_array_bytes;
    w.zero_bytes = o;
    w.off_frame_offsets = o; o = align_up(o + sizeof(int32_t) * (size_t)(B + 1), 256);
    w.off_key = o;         o = align_up(o + sizeof(int32_t) * (size_t)n, 256);
    w.off_arrival = o;     o = align_up(o + sizeof(uint32_t) * (size_t)n, 256);
    w.off_sorted_rows = o; o = align_up(o + sizeof(float) * (size_t)n * (size_t)w.RW, 256);
    w.off_prec = o;        o = align_up(o + sizeof(int4) * (size_t)n, 256);
    w.total = o;
    return w;
}

// launchers (pillar_path.cu); return a cudaError_t-or-HGSF status and count launches
int launch_pillar_path(const PathParams &p, bool with_pfn, bool abs_xyz, bool dist, size_t zero_bytes, void *zero_base,
                       cudaStream_t stream, int *launches);

int emit_timing_begin(int capacity);
int emit_timing_collect(float *ms, int n);

}  // namespace hgsf
