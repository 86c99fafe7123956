// abi.cu -- the extern "C" surface declared in include/hgsfusion_b200.h.
// Validation, parameter packing, workspace carving; no allocation, no host synchronisation.
#include "../../include/hgsfusion_b200_debug.h"
#include "contract_ops.cuh"
#include "hybrid_points.cuh"
#include "pillar_path.cuh"
#include "pillarnet_ops.cuh"
#include "subm_conv.cuh"
#include "train_ops.cuh"

#include <climits>

using namespace hgsf;

static thread_local int g_last_launches = 0;

extern "C" {

int hgsf_abi_version(void) { return HGSF_ABI_VERSION; }

const char *hgsf_status_string(int status) {
    switch (status) {
        case HGSF_OK: return "ok";
        case HGSF_ERR_INVALID_ARG: return "invalid argument";
        case HGSF_ERR_UNSUPPORTED: return "configuration outside the compiled kernel set";
        case HGSF_ERR_WORKSPACE: return "workspace too small or misaligned";
        case HGSF_ERR_DRIVER: return "cuTensorMapEncodeTiled unavailable or failed";
        default: return status > 0 ? cudaGetErrorString((cudaError_t)status) : "unknown status";
    }
}

int hgsf_last_launch_count(void) { return g_last_launches; }

int hgsf_emit_timing_begin(int capacity) { return emit_timing_begin(capacity); }

int hgsf_emit_timing_collect(float *ms, int n) { return emit_timing_collect(ms, n); }

static bool geom_ok(const hgsf_geometry *g) {
    if (!g) return false;
    for (int j = 0; j < 3; ++j)
        if (g->grid[j] <= 0 || !(g->voxel_size[j] > 0.f)) return false;
    return true;
}

int64_t hgsf_pillar_capacity(const hgsf_geometry *g, int64_t n, int32_t B, int32_t max_voxels) {
    if (!geom_ok(g) || n < 0 || B <= 0 || max_voxels < 0) return -1;
    const int64_t cells = (int64_t)g->grid[0] * g->grid[1] * g->grid[2];
    const int64_t per_frame = cells < max_voxels ? cells : max_voxels;
    const int64_t cap = per_frame * B;
    return n < cap ? n : cap;
}

int hgsf_workspace_size(const hgsf_geometry *g, int64_t n, int32_t B, int32_t F, size_t *bytes) {
    if (!geom_ok(g) || !bytes || n < 0 || B <= 0 || F < 3) return HGSF_ERR_INVALID_ARG;
    const WorkspaceLayout w = workspace_layout(n, B, g->grid[0], g->grid[1], g->grid[2], F);
    if (w.cells * B > INT_MAX || n > INT_MAX) return HGSF_ERR_UNSUPPORTED;
    *bytes = w.total;
    return HGSF_OK;
}

static int path_params(const hgsf_geometry *g, const hgsf_points *pt, const hgsf_pfn *pfn, int32_t P, int32_t max_voxels,
                       void *ws, size_t ws_bytes, const hgsf_pillar_outputs *out, PathParams &p, bool &abs_xyz, bool &dist) {
    if (!geom_ok(g) || !pt || !out || !ws) return HGSF_ERR_INVALID_ARG;
    if (pt->n < 0 || pt->batch_size <= 0 || pt->num_features < 3 || pt->xyz_col < 0 ||
        pt->xyz_col + pt->num_features > pt->stride || (pt->n > 0 && !pt->data))
        return HGSF_ERR_INVALID_ARG;
    if (!pt->frame_offsets && (pt->batch_col < 0 || pt->batch_col >= pt->stride)) return HGSF_ERR_INVALID_ARG;
    if (P <= 0 || max_voxels < 0) return HGSF_ERR_INVALID_ARG;
    if (!out->voxel_coords || !out->voxel_num_points || !out->num_pillars) return HGSF_ERR_INVALID_ARG;
    if (P > 128) return HGSF_ERR_UNSUPPORTED;          // the rank list of a pillar with more than 32 points is held in shared memory
    const WorkspaceLayout w = workspace_layout(pt->n, pt->batch_size, g->grid[0], g->grid[1], g->grid[2], pt->num_features);
    const int64_t cells = w.cells;                     // table entries per frame (row pitch padded to 32)
    if (cells * pt->batch_size > INT_MAX || pt->n > INT_MAX) return HGSF_ERR_UNSUPPORTED;
    if (pt->batch_size > 32767 || g->grid[0] > 65535 || g->grid[1] > 65535 || g->grid[2] > 65535) return HGSF_ERR_UNSUPPORTED;
    if (out->pillar_capacity < hgsf_pillar_capacity(g, pt->n, pt->batch_size, max_voxels)) return HGSF_ERR_INVALID_ARG;
    if (ws_bytes < w.total || (reinterpret_cast<uintptr_t>(ws) & 255)) return HGSF_ERR_WORKSPACE;

    p = PathParams{};
    p.pts = pt->data; p.n = (int)pt->n; p.stride = pt->stride; p.xyz_col = pt->xyz_col; p.F = pt->num_features;
    p.batch_col = pt->batch_col; p.frame_offsets_in = pt->frame_offsets; p.B = pt->batch_size;
    for (int j = 0; j < 3; ++j) { p.rmin[j] = g->pc_range[j]; p.vsize[j] = g->voxel_size[j]; p.voff[j] = g->centre_off[j]; }
    p.nx = g->grid[0]; p.ny = g->grid[1]; p.nz = g->grid[2]; p.nxp = w.nxp; p.cells = (int)cells;
    p.tiles_per_row = w.nxp / 32;
    p.div_cells = make_fastdiv((uint32_t)cells); p.div_plane = make_fastdiv((uint32_t)(p.ny * p.nxp));
    p.div_nxp = make_fastdiv((uint32_t)p.nxp); p.div_tpr = make_fastdiv((uint32_t)p.tiles_per_row);
    p.div_ny = make_fastdiv((uint32_t)p.ny);
    p.P = P; p.max_voxels = max_voxels;
    uint8_t *base = static_cast<uint8_t *>(ws);
    p.ticket = reinterpret_cast<uint32_t *>(base + w.off_ticket);
    p.state = reinterpret_cast<uint64_t *>(base + w.off_state);
    // "the table is clean" is only meaningful for the same table: tie the mark to its place and size
    p.magic = 0x48475346c1ea9e5bull ^ ((uint64_t)w.cell_array_bytes * 0x9E3779B97F4A7C15ull) ^ (uint64_t)w.off_table;
    p.scan_desc = reinterpret_cast<uint32_t *>(base + w.off_desc);
    p.frame_raw_base = reinterpret_cast<int32_t *>(base + w.off_raw_base);
    p.cutoff = reinterpret_cast<int32_t *>(base + w.off_cutoff);
    p.flags = pt->flags;
    p.cell_tag = reinterpret_cast<uint32_t *>(base + w.off_table);
    p.cell_cnt = reinterpret_cast<uint32_t *>(base + w.off_table + w.cell_array_bytes);
    p.cell_start = reinterpret_cast<uint32_t *>(base + w.off_table + 2 * w.cell_array_bytes);
    p.table_bytes = 2 * w.cell_array_bytes;
    p.tile_rec = reinterpret_cast<uint4 *>(base + w.off_tile_rec);
    p.heavy_list = reinterpret_cast<uint32_t *>(base + w.off_heavy);
    p.frame_offsets = pt->frame_offsets ? const_cast<int32_t *>(pt->frame_offsets)
                                        : reinterpret_cast<int32_t *>(base + w.off_frame_offsets);
    p.key = reinterpret_cast<int32_t *>(base + w.off_key);
    p.arrival = reinterpret_cast<uint32_t *>(base + w.off_arrival);
    p.sorted_rows = reinterpret_cast<float *>(base + w.off_sorted_rows);
    p.pil = reinterpret_cast<int4 *>(base + w.off_pil);
    p.RW = w.RW;
    p.coords = out->voxel_coords; p.num = out->voxel_num_points; p.num_pillars = out->num_pillars;
    p.voxels = out->voxels;
    abs_xyz = true; dist = false;
    if (pfn) {
        if (!pfn->weight || pfn->out_channels <= 0) return HGSF_ERR_INVALID_ARG;
        const bool bn = pfn->bn_weight || pfn->bn_bias || pfn->bn_mean || pfn->bn_var;
        if (bn && !(pfn->bn_weight && pfn->bn_bias && pfn->bn_mean && pfn->bn_var)) return HGSF_ERR_INVALID_ARG;
        if (!bn && !pfn->bias) return HGSF_ERR_INVALID_ARG;
        abs_xyz = pfn->use_absolute_xyz != 0; dist = pfn->with_distance != 0;
        const int cin = (abs_xyz ? p.F : p.F - 3) + 6 + (dist ? 1 : 0);
        if (cin != pfn->in_channels) return HGSF_ERR_INVALID_ARG;
        if (out->spatial_features && g->grid[2] != 1) return HGSF_ERR_INVALID_ARG;   // PointPillarScatter asserts nz == 1
        p.W = pfn->weight; p.bias = pfn->bias; p.bn_w = pfn->bn_weight; p.bn_b = pfn->bn_bias;
        p.bn_m = pfn->bn_mean; p.bn_v = pfn->bn_var; p.eps = pfn->bn_eps;
        p.Cin = cin; p.C = pfn->out_channels;
        p.feats = out->pillar_features; p.canvas = out->spatial_features;
    }
    return HGSF_OK;
}

static int run_path(const hgsf_geometry *g, const hgsf_points *pt, const hgsf_pfn *pfn, int32_t P, int32_t max_voxels,
                    void *ws, size_t ws_bytes, const hgsf_pillar_outputs *out, hgsf_stream_t stream) {
    g_last_launches = 0;
    PathParams p{};
    bool abs_xyz = true, dist = false;
    const int st = path_params(g, pt, pfn, P, max_voxels, ws, ws_bytes, out, p, abs_xyz, dist);
    if (st != HGSF_OK) return st;
    return launch_pillar_path(p, pfn != nullptr, abs_xyz, dist, static_cast<cudaStream_t>(stream), &g_last_launches);
}

int hgsf_pillarize(const hgsf_geometry *g, const hgsf_points *pt, int32_t P, int32_t max_voxels, void *ws,
                   size_t ws_bytes, const hgsf_pillar_outputs *out, hgsf_stream_t stream) {
    return run_path(g, pt, nullptr, P, max_voxels, ws, ws_bytes, out, stream);
}

int hgsf_points_to_bev(const hgsf_geometry *g, const hgsf_points *pt, const hgsf_pfn *pfn, int32_t P,
                       int32_t max_voxels, void *ws, size_t ws_bytes, const hgsf_pillar_outputs *out,
                       hgsf_stream_t stream) {
    if (!pfn) return HGSF_ERR_INVALID_ARG;
    return run_path(g, pt, pfn, P, max_voxels, ws, ws_bytes, out, stream);
}

int hgsf_pillar_vfe(const hgsf_geometry *g, const hgsf_pfn *pfn, const float *voxels, const void *coords,
                    const void *num, int32_t coords_are_float, int32_t num_are_float, int64_t M, int32_t P,
                    int32_t F, float *pillar_features, hgsf_stream_t stream) {
    g_last_launches = 0;
    if (!geom_ok(g) || !pfn || M < 0 || P <= 0 || F < 3) return HGSF_ERR_INVALID_ARG;
    if (M > 0 && (!voxels || !coords || !num || !pillar_features)) return HGSF_ERR_INVALID_ARG;
    if (!pfn->weight) return HGSF_ERR_INVALID_ARG;
    const bool bn = pfn->bn_weight != nullptr;
    if (bn && !(pfn->bn_bias && pfn->bn_mean && pfn->bn_var)) return HGSF_ERR_INVALID_ARG;
    if (!bn && !pfn->bias) return HGSF_ERR_INVALID_ARG;
    const bool abs_xyz = pfn->use_absolute_xyz != 0, dist = pfn->with_distance != 0;
    if ((abs_xyz ? F : F - 3) + 6 + (dist ? 1 : 0) != pfn->in_channels) return HGSF_ERR_INVALID_ARG;
    VfeParams q{};
    q.voxels = voxels; q.coords = coords; q.num = num; q.coords_float = coords_are_float; q.num_float = num_are_float;
    q.M = M; q.P = P; q.F = F; q.C = pfn->out_channels;
    for (int j = 0; j < 3; ++j) { q.vsize[j] = g->voxel_size[j]; q.voff[j] = g->centre_off[j]; }
    q.pfn = PfnArgs{pfn->weight, pfn->bias, pfn->bn_weight, pfn->bn_bias, pfn->bn_mean, pfn->bn_var, pfn->bn_eps};
    q.out = pillar_features;
    const int st = launch_vfe(q, abs_xyz, dist, static_cast<cudaStream_t>(stream));
    if (st == HGSF_OK) g_last_launches = M > 0 ? 1 : 0;
    return st;
}

int hgsf_pillar_vfe_stacked(const hgsf_geometry *g, const hgsf_pfn *pfn0, const hgsf_pfn *pfn1, const float *voxels,
                            const void *coords, const void *num, int32_t coords_are_float, int32_t num_are_float, int64_t M,
                            int32_t P, int32_t F, float *pillar_features, hgsf_stream_t stream) {
    g_last_launches = 0;
    if (!geom_ok(g) || !pfn0 || !pfn1 || M < 0 || P <= 0 || F < 3) return HGSF_ERR_INVALID_ARG;
    if (M > 0 && (!voxels || !coords || !num || !pillar_features)) return HGSF_ERR_INVALID_ARG;
    for (const hgsf_pfn *pf : {pfn0, pfn1}) {
        if (!pf->weight) return HGSF_ERR_INVALID_ARG;
        const bool bn = pf->bn_weight != nullptr;
        if (bn && !(pf->bn_bias && pf->bn_mean && pf->bn_var)) return HGSF_ERR_INVALID_ARG;
        if (!bn && !pf->bias) return HGSF_ERR_INVALID_ARG;
    }
    const bool abs_xyz = pfn0->use_absolute_xyz != 0, dist = pfn0->with_distance != 0;
    if ((abs_xyz ? F : F - 3) + 6 + (dist ? 1 : 0) != pfn0->in_channels) return HGSF_ERR_INVALID_ARG;
    if (pfn1->in_channels != 2 * pfn0->out_channels) return HGSF_ERR_INVALID_ARG;
    VfeParams q{};
    q.voxels = voxels; q.coords = coords; q.num = num; q.coords_float = coords_are_float; q.num_float = num_are_float;
    q.M = M; q.P = P; q.F = F; q.C = pfn0->out_channels; q.C1 = pfn1->out_channels;
    for (int j = 0; j < 3; ++j) { q.vsize[j] = g->voxel_size[j]; q.voff[j] = g->centre_off[j]; }
    q.pfn = PfnArgs{pfn0->weight, pfn0->bias, pfn0->bn_weight, pfn0->bn_bias, pfn0->bn_mean, pfn0->bn_var, pfn0->bn_eps};
    q.pfn1 = PfnArgs{pfn1->weight, pfn1->bias, pfn1->bn_weight, pfn1->bn_bias, pfn1->bn_mean, pfn1->bn_var, pfn1->bn_eps};
    q.out = pillar_features;
    const int st = launch_vfe_stacked(q, abs_xyz, dist, static_cast<cudaStream_t>(stream));
    if (st == HGSF_OK) g_last_launches = M > 0 ? 1 : 0;
    return st;
}

int hgsf_scatter_workspace_size(const hgsf_geometry *g, int32_t B, size_t *bytes) {
    if (!geom_ok(g) || !bytes || B <= 0) return HGSF_ERR_INVALID_ARG;
    *bytes = align_up(sizeof(unsigned) * (size_t)B * g->grid[0] * g->grid[1] * g->grid[2], 256);
    return HGSF_OK;
}

int hgsf_pointpillar_scatter(const hgsf_geometry *g, const float *feats, const void *coords, int32_t coords_are_float,
                             int64_t M, int32_t C, int32_t B, void *ws, size_t ws_bytes, float *canvas,
                             hgsf_stream_t stream) {
    g_last_launches = 0;
    if (!geom_ok(g) || M < 0 || B <= 0 || C <= 0 || !canvas || !ws) return HGSF_ERR_INVALID_ARG;
    if (M > 0 && (!feats || !coords)) return HGSF_ERR_INVALID_ARG;
    if (g->grid[2] != 1) return HGSF_ERR_INVALID_ARG;          // pointpillar_scatter.py:12 asserts nz == 1
    if (C % 32 != 0 || C > 256) return HGSF_ERR_UNSUPPORTED;
    size_t need = 0;
    hgsf_scatter_workspace_size(g, B, &need);
    if (ws_bytes < need || (reinterpret_cast<uintptr_t>(ws) & 255)) return HGSF_ERR_WORKSPACE;
    if ((int64_t)g->grid[0] * g->grid[1] * B > INT_MAX || M >= UINT_MAX) return HGSF_ERR_UNSUPPORTED;
    ScatterParams q{};
    q.feats = feats; q.coords = coords; q.coords_float = coords_are_float; q.coord_cols = 4; q.M = M; q.C = C; q.B = B;
    q.ny = g->grid[1]; q.nx = g->grid[0]; q.plane = (long long)g->grid[0] * g->grid[1];
    q.map = static_cast<unsigned *>(ws); q.canvas = canvas;
    return launch_scatter(q, static_cast<cudaStream_t>(stream), &g_last_launches);
}

// ---- Path B (PillarNet reader): replaces the pybind module pillar_cuda (pillar_api.cpp:10-22) ----------------------
int hgsf_pillarnet_workspace_size(int64_t n_points, size_t *bytes) {
    if (!bytes || n_points < 0) return HGSF_ERR_INVALID_ARG;
    *bytes = align_up(sizeof(int) * (size_t)n_points, 256) + sizeof(uint32_t) * 8192;
    return HGSF_OK;
}

int hgsf_pillarnet_indices(float bev_size, const float *xyz, const int32_t *xyz_batch_cnt, int64_t N, int32_t B, int32_t H,
                           int32_t W, int32_t *pillar_bev_indices, int32_t *pillars, int32_t *indice_pairs,
                           int32_t *point_idx, int32_t *pillar_idx, int32_t *counts, void *ws, size_t ws_bytes,
                           hgsf_stream_t stream) {
    g_last_launches = 0;
    if (!(bev_size > 0.f) || N < 0 || B <= 0 || H <= 0 || W <= 0 || !xyz_batch_cnt || !pillar_bev_indices || !pillars ||
        !point_idx || !pillar_idx || !counts || !ws || (N > 0 && !xyz))
        return HGSF_ERR_INVALID_ARG;
    if ((int64_t)B * H * W > INT_MAX || N > INT_MAX || B > 1024) return HGSF_ERR_UNSUPPORTED;
    size_t need = 0;
    hgsf_pillarnet_workspace_size(N, &need);
    if (ws_bytes < need || (reinterpret_cast<uintptr_t>(ws) & 255)) return HGSF_ERR_WORKSPACE;
    PillarNetParams q{};
    q.xyz = xyz; q.cnt = xyz_batch_cnt; q.N = N; q.B = B; q.H = H; q.W = W; q.bev_size = bev_size;
    q.bev = pillar_bev_indices; q.pillars = pillars; q.pairs = indice_pairs; q.point_idx = point_idx; q.pillar_idx = pillar_idx;
    q.counts = counts;
    q.key = static_cast<int *>(ws);
    q.partial = reinterpret_cast<uint32_t *>(static_cast<uint8_t *>(ws) + align_up(sizeof(int) * (size_t)N, 256));
    const int st = launch_pillarnet_indices(q, static_cast<cudaStream_t>(stream));
    if (st == HGSF_OK) g_last_launches = 1;
    return st;
}

int hgsf_gather_feature(const int32_t *set_indices, const float *features, int64_t L, int32_t C, float *out,
                        hgsf_stream_t stream) {
    g_last_launches = 0;
    if (L < 0 || C <= 0 || (L > 0 && (!set_indices || !features || !out))) return HGSF_ERR_INVALID_ARG;
    const int st = launch_gather(L, C, set_indices, features, out, static_cast<cudaStream_t>(stream));
    if (st == HGSF_OK) g_last_launches = L > 0;
    return st;
}

int hgsf_gather_feature_grad(const int32_t *set_indices, const float *grad_out, int64_t L, int32_t C, float *grad_features,
                             hgsf_stream_t stream) {
    g_last_launches = 0;
    if (L < 0 || C <= 0 || (L > 0 && (!set_indices || !grad_out || !grad_features))) return HGSF_ERR_INVALID_ARG;
    const int st = launch_gather_grad(L, C, set_indices, grad_out, grad_features, static_cast<cudaStream_t>(stream));
    if (st == HGSF_OK) g_last_launches = L > 0;
    return st;
}

int hgsf_scatter_max(const int32_t *index, const float *src, int32_t C, int64_t L, int64_t M, int32_t *arg, float *out,
                     hgsf_stream_t stream) {
    g_last_launches = 0;
    if (C <= 0 || L < 0 || M < 0 || (M > 0 && !out) || (L > 0 && (!index || !src))) return HGSF_ERR_INVALID_ARG;
    if ((int64_t)C * L > INT_MAX || (int64_t)C * M > INT_MAX) return HGSF_ERR_UNSUPPORTED;   // arg holds flat ids as int32
    if (M == 0) return HGSF_OK;
    return launch_scatter_max(C, L, M, index, src, arg, out, static_cast<cudaStream_t>(stream), &g_last_launches);
}

int hgsf_scatter_max_grad(const int32_t *arg, const float *grad_out, int32_t C, int64_t M, float *grad_src,
                          hgsf_stream_t stream) {
    g_last_launches = 0;
    if (C <= 0 || M < 0 || (M > 0 && (!arg || !grad_out || !grad_src))) return HGSF_ERR_INVALID_ARG;
    const int st = launch_scatter_max_grad(C, M, arg, grad_out, grad_src, static_cast<cudaStream_t>(stream));
    if (st == HGSF_OK) g_last_launches = M > 0;
    return st;
}

int hgsf_split_encode(const float *points, int64_t n_rows, int32_t Fin, int32_t Fout, int32_t n_split, int32_t batch_size,
                      int32_t encoding, const float *pc_min, const int32_t *order, float *xyz, float *pt_features,
                      int32_t *xyz_batch_cnt, int32_t *info, hgsf_stream_t stream) {
    g_last_launches = 0;
    if (n_rows < 0 || Fin < 3 || batch_size <= 0 || !pc_min || !xyz_batch_cnt || !info) return HGSF_ERR_INVALID_ARG;
    if (n_rows > 0 && (!points || !xyz || !pt_features)) return HGSF_ERR_INVALID_ARG;
    if (Fin + 1 > 40 || batch_size > 32767) return HGSF_ERR_UNSUPPORTED;
    switch (encoding) {
        case HGSF_ENCODE_SPLIT:
            if (Fin < 5 || n_split < 0 || 3 + n_split > Fin - 2 || Fout < 3 + 2 * n_split + 2) return HGSF_ERR_INVALID_ARG;
            break;
        case HGSF_ENCODE_COPY: if (Fout != Fin) return HGSF_ERR_INVALID_ARG; break;
        case HGSF_ENCODE_DIRECT: if (Fin < 5 || Fout != Fin - 2) return HGSF_ERR_INVALID_ARG; break;
        default: return HGSF_ERR_INVALID_ARG;
    }
    SplitEncodeParams q{};
    q.points = points; q.order = order; q.L = n_rows; q.Fin = Fin; q.Fout = Fout; q.n_split = n_split;
    q.B = batch_size; q.mode = encoding;
    for (int j = 0; j < 3; ++j) q.pc_min[j] = pc_min[j];
    q.xyz = xyz; q.feat = pt_features; q.cnt = xyz_batch_cnt; q.info = info;
    const int st = launch_split_encode(q, static_cast<cudaStream_t>(stream));
    if (st == HGSF_OK) g_last_launches = n_rows > 0;
    return st;
}

int hgsf_pillarnet_reader(const float *xyz, const float *pt_features, int32_t Cf, const int32_t *point_idx,
                          const int32_t *pillar_idx, int64_t L, const int32_t *pillars, int64_t M, float bev_size, float z_center,
                          const hgsf_pfn *pfn, float *pillar_features, hgsf_stream_t stream) {
    g_last_launches = 0;
    if (L < 0 || M < 0 || Cf <= 0 || !pfn || !pfn->weight || !(bev_size > 0.f)) return HGSF_ERR_INVALID_ARG;
    if (!(pfn->bn_weight && pfn->bn_bias && pfn->bn_mean && pfn->bn_var)) return HGSF_ERR_INVALID_ARG;
    if (pfn->in_channels != Cf + 6) return HGSF_ERR_INVALID_ARG;
    if (M > 0 && !pillar_features) return HGSF_ERR_INVALID_ARG;
    if (L > 0 && (!xyz || !pt_features || !point_idx || !pillar_idx || !pillars || M == 0)) return HGSF_ERR_INVALID_ARG;
    if (pfn->out_channels != 32 || Cf + 6 > 40) return HGSF_ERR_UNSUPPORTED;
    ReaderParams q{};
    q.xyz = xyz; q.feat = pt_features; q.Cf = Cf; q.point_idx = point_idx; q.pillar_idx = pillar_idx; q.L = L; q.M = M;
    q.pillars = pillars; q.bev_size = bev_size; q.z_center = z_center;
    q.W = pfn->weight; q.bn_w = pfn->bn_weight; q.bn_b = pfn->bn_bias; q.bn_m = pfn->bn_mean; q.bn_v = pfn->bn_var;
    q.eps = pfn->bn_eps; q.out = pillar_features;
    return launch_reader_fused(q, static_cast<cudaStream_t>(stream), &g_last_launches);
}

// ---- training (train_ops.cu) ------------------------------------------------------------------------------------------
// the statistics layout of train_ops.cu + the second-moment scratch of hgsf_points_to_bev_train behind it (16 rows of cin)
int64_t hgsf_train_stats_doubles(int32_t C, int32_t cin) { return (C > 0 && cin > 0) ? (int64_t)(train_stats_len(C, cin) + 16 * (size_t)cin) : 0; }
int64_t hgsf_train_scratch_doubles(int32_t C, int32_t cin) { return (C > 0 && cin > 0) ? (int64_t)train_acc_len(C, cin) : 0; }

static int vfe_params(const hgsf_geometry *g, const hgsf_pfn *pfn, const float *voxels, const void *coords, const void *num,
                      int32_t coords_are_float, int32_t num_are_float, int64_t M, int32_t P, int32_t F, VfeParams &q,
                      bool &abs_xyz, bool &dist) {
    if (!geom_ok(g) || !pfn || M < 0 || P <= 0 || F < 3 || !pfn->weight) return HGSF_ERR_INVALID_ARG;
    if (M > 0 && (!voxels || !coords || !num)) return HGSF_ERR_INVALID_ARG;
    abs_xyz = pfn->use_absolute_xyz != 0; dist = pfn->with_distance != 0;
    if ((abs_xyz ? F : F - 3) + 6 + (dist ? 1 : 0) != pfn->in_channels) return HGSF_ERR_INVALID_ARG;
    q = VfeParams{};
    q.voxels = voxels; q.coords = coords; q.num = num; q.coords_float = coords_are_float; q.num_float = num_are_float;
    q.M = M; q.P = P; q.F = F; q.C = pfn->out_channels;
    for (int j = 0; j < 3; ++j) { q.vsize[j] = g->voxel_size[j]; q.voff[j] = g->centre_off[j]; }
    q.pfn = PfnArgs{pfn->weight, pfn->bias, pfn->bn_weight, pfn->bn_bias, pfn->bn_mean, pfn->bn_var, pfn->bn_eps};
    return HGSF_OK;
}

int hgsf_pillar_vfe_batch_stats(const hgsf_geometry *g, const hgsf_pfn *pfn, const float *voxels, const void *coords,
                                const void *num, int32_t coords_are_float, int32_t num_are_float, int64_t M, int32_t P,
                                int32_t F, float momentum, float *running_mean, float *running_var, float *batch_mean,
                                float *batch_var, double *stats, hgsf_stream_t stream) {
    g_last_launches = 0;
    VfeParams q;
    bool abs_xyz, dist;
    int st = vfe_params(g, pfn, voxels, coords, num, coords_are_float, num_are_float, M, P, F, q, abs_xyz, dist);
    if (st != HGSF_OK) return st;
    if (!batch_mean || !batch_var || !stats || M == 0) return HGSF_ERR_INVALID_ARG;   // BatchNorm needs at least one row
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    st = launch_vfe_stats(q, abs_xyz, dist, stats, s);
    if (st != HGSF_OK) return st;
    st = launch_bn_finalize(stats, (double)M * P, q.C, momentum, running_mean, running_var, batch_mean, batch_var, s);
    if (st == HGSF_OK) g_last_launches = 2;
    return st;
}

int hgsf_pillar_vfe_backward(const hgsf_geometry *g, const hgsf_pfn *pfn, const float *voxels, const void *coords,
                             const void *num, int32_t coords_are_float, int32_t num_are_float, int64_t M, int32_t P,
                             int32_t F, const float *grad_out, const double *stats, double *scratch, float *grad_weight,
                             float *grad_bn_weight, float *grad_bn_bias, hgsf_stream_t stream) {
    g_last_launches = 0;
    VfeParams q;
    bool abs_xyz, dist;
    int st = vfe_params(g, pfn, voxels, coords, num, coords_are_float, num_are_float, M, P, F, q, abs_xyz, dist);
    if (st != HGSF_OK) return st;
    if (!scratch || !grad_weight || (M > 0 && !grad_out)) return HGSF_ERR_INVALID_ARG;
    const bool bn = pfn->bn_weight != nullptr;
    if (bn && !(pfn->bn_bias && pfn->bn_mean && pfn->bn_var && grad_bn_weight && grad_bn_bias)) return HGSF_ERR_INVALID_ARG;
    if (!bn && (!pfn->bias || stats)) return HGSF_ERR_INVALID_ARG;
    const int mode = !bn ? 2 : (stats ? 0 : 1);
    return launch_vfe_backward(q, abs_xyz, dist, grad_out, stats, mode, scratch, grad_weight, grad_bn_weight, grad_bn_bias,
                               static_cast<cudaStream_t>(stream), &g_last_launches);
}

int hgsf_pointpillar_scatter_backward(const hgsf_geometry *g, const float *grad_canvas, const void *coords,
                                      int32_t coords_are_float, int64_t M, int32_t C, int32_t B, float *grad_feats,
                                      hgsf_stream_t stream) {
    g_last_launches = 0;
    if (!geom_ok(g) || M < 0 || B <= 0 || C <= 0 || g->grid[2] != 1) return HGSF_ERR_INVALID_ARG;
    if (M > 0 && (!grad_canvas || !coords || !grad_feats)) return HGSF_ERR_INVALID_ARG;
    const int st = launch_scatter_grad(grad_canvas, coords, coords_are_float, M, C, B, g->grid[1], g->grid[0], grad_feats,
                                       static_cast<cudaStream_t>(stream));
    if (st == HGSF_OK) g_last_launches = M > 0;
    return st;
}

int hgsf_points_to_bev_train(const hgsf_geometry *g, const hgsf_points *pt, const hgsf_pfn *pfn, int32_t P, int32_t max_voxels,
                             void *ws, size_t ws_bytes, const hgsf_pillar_outputs *out, float momentum, float *running_mean,
                             float *running_var, float *batch_mean, float *batch_var, double *stats, hgsf_stream_t stream) {
    g_last_launches = 0;
    if (!pfn || !out || !pfn->bn_weight || !pfn->bn_bias || !batch_mean || !batch_var || !stats) return HGSF_ERR_INVALID_ARG;
    if (!out->spatial_features || !out->pillar_features) return HGSF_ERR_INVALID_ARG;
    hgsf_pfn train = *pfn;                       // the statistics the forward normalises with are this call's own
    train.bn_mean = batch_mean; train.bn_var = batch_var;
    PathParams p{};
    bool abs_xyz = true, dist = false;
    const int st = path_params(g, pt, &train, P, max_voxels, ws, ws_bytes, out, p, abs_xyz, dist);
    if (st != HGSF_OK) return st;
    p.stats = stats; p.stats_S = stats + train_stats_len(p.C, p.Cin); p.batch_mean = batch_mean; p.batch_var = batch_var; p.run_mean = running_mean; p.run_var = running_var;
    p.momentum = momentum;
    return launch_pillar_path_train(p, abs_xyz, dist, static_cast<cudaStream_t>(stream), &g_last_launches);
}

int hgsf_points_to_bev_train_backward(const hgsf_geometry *g, const hgsf_pfn *pfn, const float *voxels, const int32_t *coords,
                                      const int32_t *num, int64_t capacity, const int32_t *num_pillars, int32_t P, int32_t F,
                                      int32_t B, const float *grad_canvas, const float *grad_feats_in, float *grad_rows,
                                      const double *stats, double *scratch, float *grad_weight, float *grad_bn_weight,
                                      float *grad_bn_bias, hgsf_stream_t stream) {
    g_last_launches = 0;
    VfeParams q;
    bool abs_xyz, dist;
    int st = vfe_params(g, pfn, voxels, coords, num, 0, 0, capacity, P, F, q, abs_xyz, dist);
    if (st != HGSF_OK) return st;
    if (!num_pillars || !grad_rows || !stats || !scratch || !grad_weight || !grad_bn_weight || !grad_bn_bias || B <= 0 ||
        g->grid[2] != 1)
        return HGSF_ERR_INVALID_ARG;
    if (!(pfn->bn_weight && pfn->bn_bias && pfn->bn_mean && pfn->bn_var)) return HGSF_ERR_INVALID_ARG;
    if (capacity == 0) return HGSF_ERR_INVALID_ARG;
    q.M_dev = num_pillars;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    st = launch_scatter_grad(grad_canvas, coords, 0, capacity, q.C, B, g->grid[1], g->grid[0], grad_rows, s, num_pillars,
                             grad_feats_in);
    if (st != HGSF_OK) return st;
    int nl = 0;
    st = launch_vfe_backward(q, abs_xyz, dist, grad_rows, stats, 0, scratch, grad_weight, grad_bn_weight, grad_bn_bias, s, &nl);
    if (st == HGSF_OK) g_last_launches = 1 + nl;
    return st;
}

int hgsf_hybrid_workspace_size(int64_t n_candidates, size_t *bytes) {
    if (n_candidates < 0 || !bytes) return HGSF_ERR_INVALID_ARG;
    *bytes = hybrid_workspace_bytes(n_candidates);
    return HGSF_OK;
}

int hgsf_assemble_hybrid_points(const hgsf_hybrid_inputs *in, const float *calib, const double *range_xy, void *ws,
                                size_t ws_bytes, float *points_out, int32_t *frame_offsets_out, hgsf_stream_t stream) {
    g_last_launches = 0;
    if (!in || !frame_offsets_out || in->n_candidates < 0 || in->batch_size <= 0 || in->real_features < 3 ||
        !in->real_offsets)
        return HGSF_ERR_INVALID_ARG;
    const bool hybrid = in->hybrid_features != 0;
    if (hybrid && (in->hybrid_features < in->real_features || !in->gt_offsets || !in->virt_offsets)) return HGSF_ERR_INVALID_ARG;
    if (in->n_candidates > 0 && (!points_out || !ws)) return HGSF_ERR_INVALID_ARG;
    if (in->n_candidates >= (int64_t)INT_MAX || in->batch_size > 32767) return HGSF_ERR_UNSUPPORTED;
    HybridParams q{};
    q.real = in->real; q.gt = hybrid ? in->gt_real : nullptr; q.virt = hybrid ? in->virt : nullptr;
    q.real_off = in->real_offsets; q.gt_off = in->gt_offsets; q.virt_off = in->virt_offsets;
    q.calib = calib;
    q.Fr = in->real_features; q.W = in->hybrid_features; q.B = in->batch_size; q.n = in->n_candidates;
    q.no_dup = hybrid ? in->no_dup : 0; q.dup_threshold = in->dup_threshold;
    q.mask_range = range_xy != nullptr;
    for (int j = 0; j < 4; ++j) q.range_xy[j] = range_xy ? range_xy[j] : 0.0;
    q.out = points_out; q.frame_offsets_out = frame_offsets_out;
    return launch_hybrid(q, ws, ws_bytes, static_cast<cudaStream_t>(stream), &g_last_launches);
}

int hgsf_sparse_to_dense_workspace_size(int32_t B, int32_t ny, int32_t nx, size_t *bytes) {
    if (!bytes || B <= 0 || ny <= 0 || nx <= 0) return HGSF_ERR_INVALID_ARG;
    *bytes = align_up(sizeof(unsigned) * (size_t)B * ny * nx, 256);
    return HGSF_OK;
}

int hgsf_sparse_to_dense(const float *features, const int32_t *indices, int64_t M, int32_t C, int32_t B, int32_t ny,
                         int32_t nx, void *ws, size_t ws_bytes, float *dense, hgsf_stream_t stream) {
    g_last_launches = 0;
    if (M < 0 || B <= 0 || C <= 0 || ny <= 0 || nx <= 0 || !dense || !ws) return HGSF_ERR_INVALID_ARG;
    if (M > 0 && (!features || !indices)) return HGSF_ERR_INVALID_ARG;
    if (C % 32 != 0 || C > 256) return HGSF_ERR_UNSUPPORTED;
    if ((int64_t)nx * ny * B > INT_MAX || M >= UINT_MAX || ny > 65535 || nx > 65535) return HGSF_ERR_UNSUPPORTED;
    size_t need = 0;
    hgsf_sparse_to_dense_workspace_size(B, ny, nx, &need);
    if (ws_bytes < need || (reinterpret_cast<uintptr_t>(ws) & 255)) return HGSF_ERR_WORKSPACE;
    ScatterParams q{};
    q.feats = features; q.coords = indices; q.coords_float = 0; q.coord_cols = 3; q.M = M; q.C = C; q.B = B;
    q.ny = ny; q.nx = nx; q.plane = (long long)nx * ny;
    q.map = static_cast<unsigned *>(ws); q.canvas = dense;
    return launch_scatter(q, static_cast<cudaStream_t>(stream), &g_last_launches);
}

int hgsf_subm_neighbors(const int32_t *bev, const int32_t *pillars, int64_t M, const int32_t *m_dev, int32_t B, int32_t H,
                        int32_t W, int32_t *nbr, hgsf_stream_t stream) {
    g_last_launches = 0;
    if (M < 0 || B <= 0 || H <= 0 || W <= 0) return HGSF_ERR_INVALID_ARG;
    if (M > 0 && (!bev || !pillars || !nbr)) return HGSF_ERR_INVALID_ARG;
    if ((int64_t)B * H * W > INT_MAX || M > INT_MAX / 9) return HGSF_ERR_UNSUPPORTED;
    SubmNeighborParams q{};
    q.bev = bev; q.pillars = pillars; q.m_dev = m_dev; q.M = M; q.B = B; q.H = H; q.W = W; q.stride = 1; q.nbr = nbr;
    const int st = launch_subm_neighbors(q, static_cast<cudaStream_t>(stream));
    if (st == HGSF_OK && M > 0) g_last_launches = 1;
    return st;
}

int hgsf_subm_conv3x3(const hgsf_subm_conv *conv, const float *features, const int32_t *nbr, int64_t M, const int32_t *m_dev,
                      const float *residual, float *out, hgsf_stream_t stream) {
    g_last_launches = 0;
    if (!conv || !conv->weight || M < 0 || conv->in_channels <= 0 || conv->out_channels <= 0) return HGSF_ERR_INVALID_ARG;
    if (M > 0 && (!features || !nbr || !out)) return HGSF_ERR_INVALID_ARG;
    if (conv->weight_layout != HGSF_WEIGHT_KRSC && conv->weight_layout != HGSF_WEIGHT_RSCK) return HGSF_ERR_INVALID_ARG;
    const int n_bn = (conv->bn_weight != nullptr) + (conv->bn_bias != nullptr) + (conv->bn_mean != nullptr) + (conv->bn_var != nullptr);
    if (n_bn != 0 && n_bn != 4) return HGSF_ERR_INVALID_ARG;
    if (out == features && M > 0) return HGSF_ERR_INVALID_ARG;            // neighbours' rows are read by other CTAs
    if (M > INT_MAX / 9) return HGSF_ERR_UNSUPPORTED;
    SubmConvParams q{};
    q.in = features; q.nbr = nbr; q.m_dev = m_dev; q.M = M; q.Cin = conv->in_channels; q.Cout = conv->out_channels;
    q.W = conv->weight; q.layout = conv->weight_layout; q.bias = conv->bias;
    q.bn_w = conv->bn_weight; q.bn_b = conv->bn_bias; q.bn_m = conv->bn_mean; q.bn_v = conv->bn_var; q.eps = conv->bn_eps;
    q.residual = residual; q.relu = conv->relu; q.out = out;
    const int st = launch_subm_conv(q, static_cast<cudaStream_t>(stream));
    if (st == HGSF_OK && M > 0) g_last_launches = 1;
    return st;
}

int hgsf_sparse_conv_s2_workspace_size(int64_t M, int32_t B, size_t *bytes) {
    if (!bytes || M < 0 || B <= 0) return HGSF_ERR_INVALID_ARG;
    size_t pn = 0;
    hgsf_pillarnet_workspace_size(4 * M, &pn);
    *bytes = align_up(sizeof(float) * 12 * (size_t)M, 256) + align_up(sizeof(int) * (size_t)B, 256) +
             2 * align_up(sizeof(int) * 4 * (size_t)M, 256) + align_up(pn, 256);
    return HGSF_OK;
}

int hgsf_sparse_conv_s2_indices(const int32_t *bev, const int32_t *pillars, int64_t M, const int32_t *m_dev, int32_t B, int32_t H,
                                int32_t W, int32_t *out_bev, int32_t *out_pillars, int32_t *out_counts, int32_t *nbr,
                                int64_t out_capacity, void *ws, size_t ws_bytes, hgsf_stream_t stream) {
    g_last_launches = 0;
    if (M < 0 || B <= 0 || H <= 0 || W <= 0 || !out_bev || !out_counts || !ws || out_capacity < 0) return HGSF_ERR_INVALID_ARG;
    if (M > 0 && (!bev || !pillars || !out_pillars || !nbr)) return HGSF_ERR_INVALID_ARG;
    const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
    const int64_t cells_o = (int64_t)B * Ho * Wo;
    if ((int64_t)B * H * W > INT_MAX || 4 * M > INT_MAX / 9 || B > 1024) return HGSF_ERR_UNSUPPORTED;
    if (out_capacity < (4 * M < cells_o ? 4 * M : cells_o)) return HGSF_ERR_INVALID_ARG;
    size_t need = 0;
    hgsf_sparse_conv_s2_workspace_size(M, B, &need);
    if (ws_bytes < need || (reinterpret_cast<uintptr_t>(ws) & 255)) return HGSF_ERR_WORKSPACE;
    uint8_t *w = static_cast<uint8_t *>(ws);
    float *cand = reinterpret_cast<float *>(w);            w += align_up(sizeof(float) * 12 * (size_t)M, 256);
    int *cnt = reinterpret_cast<int *>(w);                 w += align_up(sizeof(int) * (size_t)B, 256);
    int *point_idx = reinterpret_cast<int *>(w);           w += align_up(sizeof(int) * 4 * (size_t)M, 256);
    int *pillar_idx = reinterpret_cast<int *>(w);          w += align_up(sizeof(int) * 4 * (size_t)M, 256);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    cudaError_t e = cudaMemsetAsync(cnt, 0, sizeof(int) * (size_t)B, s);
    if (e != cudaSuccess) return (int)e;
    Stride2CandidateParams c{};
    c.pillars = pillars; c.m_dev = m_dev; c.M = M; c.B = B; c.Ho = Ho; c.Wo = Wo; c.cand = cand; c.cnt = cnt;
    int st = launch_stride2_candidates(c, s);
    if (st != HGSF_OK) return st;
    PillarNetParams q{};
    q.xyz = cand; q.cnt = cnt; q.N = 4 * M; q.B = B; q.H = Ho; q.W = Wo; q.bev_size = 1.f;
    q.bev = out_bev; q.pillars = out_pillars; q.pairs = nullptr; q.point_idx = point_idx; q.pillar_idx = pillar_idx;
    q.counts = out_counts;
    q.key = reinterpret_cast<int *>(w);
    q.partial = reinterpret_cast<uint32_t *>(w + align_up(sizeof(int) * 4 * (size_t)M, 256));
    st = launch_pillarnet_indices(q, s);
    if (st != HGSF_OK) return st;
    g_last_launches = M > 0 ? 2 : 1;
    if (out_capacity == 0) return HGSF_OK;
    SubmNeighborParams n{};
    n.bev = bev; n.pillars = out_pillars; n.m_dev = out_counts; n.M = out_capacity; n.B = B; n.H = H; n.W = W; n.stride = 2; n.nbr = nbr;
    st = launch_subm_neighbors(n, s);
    if (st == HGSF_OK) g_last_launches += 1;
    return st;
}

}  // extern "C"
