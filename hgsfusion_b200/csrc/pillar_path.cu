// pillar_path.cu -- points -> pillars (first-seen order) -> decorate + PFN + max -> BEV canvas.
//
// Four kernels on one stream, no host sync, no allocation:
//   k_count  (1 thread / point)  cell key; per cell: min point index + count (warp-aggregated atomics
//                                into the direct-address cell table)
//   k_scan   (1024 points / CTA) single-pass decoupled look-back scan over points: a point that is the
//                                first of its cell gets (raw pillar id, CSR start) = exclusive prefix of
//                                (first-flags, cell counts) -> pillar ids come out in first-seen order
//   k_fill   (1 thread / point)  copies each point's features (+ its index) to its pillar's CSR segment
//   k_emit   (persistent CTAs, one 32-cell x C-channel canvas tile at a time) orders each pillar's
//                                points by index, keeps the first P, decorates, runs the PFN with the
//                                weights in registers, takes the max, writes pillar_features /
//                                voxel_coords / voxel_num_points rows and the canvas tile (zeros
//                                included) with one TMA tensor store per tile.
//
// What it reproduces (file:line under the reference):
//   spconv Point2VoxelCPU3d.point_to_voxel as called by pcdet/datasets/processor/data_processor.py:55
//   collate_batch voxel keys                      pcdet/datasets/dataset.py:232-244
//   PillarVFE.forward + PFNLayer.forward          pcdet/models/backbones_3d/vfe/pillar_vfe.py:29-49,94-123
//   PointPillarScatter.forward                    pcdet/models/backbones_2d/map_to_bev/pointpillar_scatter.py:14-41
// The fp32 operation order is the one the CPU reference was measured to use (oracle/pillar_oracle.c).
#include "pillar_path.cuh"
#include "pfn.cuh"
#include "contract_ops.cuh"

#include <algorithm>
#include <cstring>
#include <vector>

namespace hgsf {

// ------------------------------------------------------------------------------------------------
// k_count
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int find_frame(const int32_t *__restrict__ off, int B, int i) {
    // largest b in [0, B) with off[b] <= i  (frames are contiguous; empty frames are skipped)
    int lo = 0, hi = B;   // invariant: off[lo] <= i < off[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(off + mid) <= i) lo = mid; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(256) k_count(const PathParams p) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    const int lane = threadIdx.x & 31;
    int key = -1;
    if (i < p.n) {
        const float *row = p.pts + (size_t)i * p.stride;
        const float x = __ldg(row + p.xyz_col), y = __ldg(row + p.xyz_col + 1), z = __ldg(row + p.xyz_col + 2);
        int b;
        if (p.frame_offsets_in) {
            b = find_frame(p.frame_offsets_in, p.B, i);
        } else {
            // batch index column (collate_batch, dataset.py:237-244); rows are frame-contiguous, so a
            // change of value marks a frame start: derive frame_offsets on the fly.
            const float fb = __ldg(row + p.batch_col);
            b = (fb >= 0.f && fb < (float)p.B) ? (int)fb : -1;
            const int bc = b < 0 ? 0 : b;
            int bprev = -1;
            if (i > 0) {
                const float fp = __ldg(row - p.stride + p.batch_col);
                bprev = (fp >= 0.f && fp < (float)p.B) ? (int)fp : 0;
            }
            for (int bb = bprev + 1; bb <= bc; ++bb) p.frame_offsets[bb] = i;
            if (i == p.n - 1)
                for (int bb = bc + 1; bb <= p.B; ++bb) p.frame_offsets[bb] = p.n;
        }
        // c = floor((pt - range_min) / voxel_size): IEEE fp32 subtract and divide, upper bound exclusive
        const float qx = floorf(__fdiv_rn(__fsub_rn(x, p.rmin[0]), p.vsize[0]));
        const float qy = floorf(__fdiv_rn(__fsub_rn(y, p.rmin[1]), p.vsize[1]));
        const float qz = floorf(__fdiv_rn(__fsub_rn(z, p.rmin[2]), p.vsize[2]));
        const bool ok = (b >= 0) && (qx >= 0.f) && (qx < (float)p.nx) && (qy >= 0.f) && (qy < (float)p.ny) &&
                        (qz >= 0.f) && (qz < (float)p.nz);
        if (ok) key = b * p.cells + (__float2int_rz(qz) * p.ny + __float2int_rz(qy)) * p.nx + __float2int_rz(qx);
    }
    // warp-aggregated atomics: lanes of the same cell elect the lowest lane (= lowest point index)
    const unsigned peers = __match_any_sync(FULL, key);
    const int leader = __ffs(peers) - 1;
    const int rank = __popc(peers & ((1u << lane) - 1u));
    unsigned base = 0;
    if (lane == leader && key >= 0) {
        CellEntry *e = p.table + key;
        base = atomicAdd(&e->cnt, (unsigned)__popc(peers));
        atomicMax(&e->tag, 0xFFFFFFFFu - (unsigned)i);
    }
    base = __shfl_sync(FULL, base, leader);
    if (i < p.n) {
        p.key[i] = key;
        p.arrival[i] = base + (unsigned)rank;
    }
}

// ------------------------------------------------------------------------------------------------
// k_scan : exclusive prefix over points of (is-first-of-its-cell, that cell's count)
// descriptor word: [63:62] status (0 none, 1 aggregate, 2 inclusive prefix) [61:31] pillars [30:0] points
// ------------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = SCAN_TILE / SCAN_THREADS;   // 4
__device__ __forceinline__ uint64_t pack2(uint32_t pillars, uint32_t points) { return ((uint64_t)pillars << 31) | points; }
constexpr uint64_t VAL_MASK = (1ull << 62) - 1;

__global__ void __launch_bounds__(SCAN_THREADS) k_scan(const PathParams p) {
    __shared__ uint32_t s_tile;
    __shared__ uint64_t s_warp[SCAN_THREADS / 32];
    __shared__ uint64_t s_prefix;
    __shared__ uint32_t s_excl[SCAN_TILE];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(p.ticket, 1u);   // dynamic tile id: look-back never waits on an unscheduled CTA
    __syncthreads();
    const int tile = (int)s_tile;
    const int base = tile * SCAN_TILE + tid * SCAN_ITEMS;

    int keys[SCAN_ITEMS];
    uint32_t flag[SCAN_ITEMS], cnt[SCAN_ITEMS];
    uint64_t local = 0;
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j) {
        const int i = base + j;
        keys[j] = (i < p.n) ? p.key[i] : -1;
    }
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j) {
        flag[j] = 0; cnt[j] = 0;
        if (keys[j] >= 0) {
            const uint2 e = *reinterpret_cast<const uint2 *>(p.table + keys[j]);   // tag, cnt
            flag[j] = (e.x == 0xFFFFFFFFu - (uint32_t)(base + j)) ? 1u : 0u;
            cnt[j] = flag[j] ? e.y : 0u;
        }
        local += pack2(flag[j], cnt[j]);
    }
    // block-wide exclusive scan of `local`
    uint64_t incl = local;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint64_t o = __shfl_up_sync(FULL, incl, d);
        if (lane >= d) incl += o;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint64_t warp_off = 0, block_total = 0;
#pragma unroll
    for (int w = 0; w < SCAN_THREADS / 32; ++w) {
        const uint64_t v = s_warp[w];
        if (w < warp) warp_off += v;
        block_total += v;
    }
    // decoupled look-back (warp 0)
    if (warp == 0) {
        if (lane == 0) st_volatile_u64(p.scan_desc + tile, (tile == 0 ? (2ull << 62) : (1ull << 62)) | block_total);
        uint64_t excl = 0;
        int look = tile - 1;
        while (look >= 0) {
            const int idx = look - lane;
            uint64_t d;
            do {
                d = (idx >= 0) ? ld_volatile_u64(p.scan_desc + idx) : (2ull << 62);
            } while (__any_sync(FULL, (d >> 62) == 0));
            const unsigned pm = __ballot_sync(FULL, (d >> 62) == 2);
            const int first = pm ? (__ffs(pm) - 1) : 32;
            uint64_t v = (lane <= first) ? (d & VAL_MASK) : 0ull;
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(FULL, v, s);
            excl += v;
            if (pm) break;
            look -= 32;
        }
        if (lane == 0) {
            if (tile > 0) st_volatile_u64(p.scan_desc + tile, (2ull << 62) | (excl + block_total));
            s_prefix = excl;
        }
    }
    __syncthreads();
    uint64_t run = s_prefix + warp_off + (incl - local);
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j) {
        const uint32_t pillars = (uint32_t)(run >> 31), points = (uint32_t)(run & 0x7FFFFFFFu);
        s_excl[tid * SCAN_ITEMS + j] = pillars;
        if (flag[j]) {
            CellEntry *e = p.table + keys[j];
            e->tag = pillars + 1u;     // raw pillar id + 1 (disjoint from the 0xFFFFFFFF-i range other threads compare against)
            e->start = points;
        }
        run += pack2(flag[j], cnt[j]);
    }
    __syncthreads();
    // raw pillar id at each frame start
    const int lo = tile * SCAN_TILE, hi = lo + SCAN_TILE;
    const bool last = (hi >= p.n);
    const uint32_t total = (uint32_t)((s_prefix + block_total) >> 31);
    for (int b = tid; b <= p.B; b += SCAN_THREADS) {
        const int o = p.frame_offsets[b];
        if (o >= lo && o < hi && o < p.n) p.frame_raw_base[b] = (int32_t)s_excl[o - lo];
        else if (last && o >= p.n) p.frame_raw_base[b] = (int32_t)total;
    }
}

// ------------------------------------------------------------------------------------------------
// k_fill : point features -> CSR segment of the point's pillar
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_fill(const PathParams p) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= p.n) return;
    const int key = p.key[i];
    if (key < 0) return;
    const uint4 e = __ldg(reinterpret_cast<const uint4 *>(p.table + key));
    const int b = key / p.cells;
    const int local = (int)(e.x - 1u) - __ldg(p.frame_raw_base + b);
    if (local >= p.max_voxels) return;            // pillar beyond max_voxels: never created
    const size_t pos = (size_t)e.z + p.arrival[i];
    const float *src = p.pts + (size_t)i * p.stride + p.xyz_col;
    float4 *dst = reinterpret_cast<float4 *>(p.sorted_rows + pos * p.RW);
    for (int k = 0; k < p.RW; k += 4) {
        float4 v;
        // F features, then the point index (slot F) that k_emit orders the pillar by
        const float fi = __int_as_float(i);
        v.x = (k + 0 < p.F) ? __ldg(src + k + 0) : (k + 0 == p.F ? fi : 0.f);
        v.y = (k + 1 < p.F) ? __ldg(src + k + 1) : (k + 1 == p.F ? fi : 0.f);
        v.z = (k + 2 < p.F) ? __ldg(src + k + 2) : (k + 2 == p.F ? fi : 0.f);
        v.w = (k + 3 < p.F) ? __ldg(src + k + 3) : (k + 3 == p.F ? fi : 0.f);
        dst[k >> 2] = v;
    }
}

// ------------------------------------------------------------------------------------------------
// k_emit
// ------------------------------------------------------------------------------------------------
// ascending bitonic sort of (key, val) across the 32 lanes of a warp
__device__ __forceinline__ void warp_bitonic(uint32_t &key, int &val, int lane, int k_begin) {
    for (int k = k_begin; k <= 32; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            const uint32_t ok = __shfl_xor_sync(FULL, key, j);
            const int ov = __shfl_xor_sync(FULL, val, j);
            const bool up = (k == 32) ? true : ((lane & k) == 0);
            const bool lower = (lane & j) == 0;
            const bool take_min = (lower == up);
            const bool swap = take_min ? (ok < key) : (ok > key);
            if (swap) { key = ok; val = ov; }
        }
    }
}

// cnt > 32 arrivals in a cell: positions of the 32 smallest point indices, ascending, one per lane
__device__ __noinline__ int select_first32(const float *__restrict__ idx0, int stride, int cnt, int lane) {
    uint32_t best = 0xFFFFFFFFu;
    int bestv = 0;
    for (int base = 0; base < cnt; base += 32) {
        const int j = base + lane;
        uint32_t k = (j < cnt) ? __float_as_uint(__ldg(idx0 + (size_t)j * stride)) : 0xFFFFFFFFu;
        int v = j;
        if (base > 0) {
            const uint32_t worst = __shfl_sync(FULL, best, 31);
            if (!__any_sync(FULL, k < worst)) continue;
        }
        warp_bitonic(k, v, lane, 2);
        if (base == 0) { best = k; bestv = v; continue; }
        // the 32 smallest of two ascending runs: min(best[l], chunk[31-l]) is bitonic; one merge pass sorts it
        const uint32_t rk = __shfl_sync(FULL, k, 31 - lane);
        const int rv = __shfl_sync(FULL, v, 31 - lane);
        if (rk < best) { best = rk; bestv = rv; }
        warp_bitonic(best, bestv, lane, 32);
    }
    return bestv;
}

// Position of a canvas tile, advanced by gridDim.x tiles at a time without any division:
//   r  = BEV row index (b*nz + z)*ny + y,  xt = tile within the row,  b = frame,  zy = row within the frame
struct TilePos {
    int r, xt, b, zy;
};
struct TileStep {
    int dr, dxt, tiles_per_row, rows_per_frame;
    __device__ __forceinline__ void advance(TilePos &t) const {
        t.xt += dxt;
        int dr2 = dr;
        if (t.xt >= tiles_per_row) { t.xt -= tiles_per_row; ++dr2; }
        t.r += dr2;
        t.zy += dr2;
        while (t.zy >= rows_per_frame) { t.zy -= rows_per_frame; ++t.b; }
    }
};

// ---- k_emit -------------------------------------------------------------------------------------
// Every WARP is an autonomous worker: it walks its own sequence of canvas tiles (a tile = 32 cells of
// one BEV row x all C channels = C rows of 128 B) with its own tile buffer, staging buffers and
// software pipeline; there is no CTA barrier inside the loop, so a warp that waits (gather, TMA
// read-out) never holds up another.  The CTA only shares the PFN weights (k-major in shared memory,
// read as broadcast float4s) and a small zero tile.
//
// Lane l owns cell l of the tile for the bookkeeping (table entry, ordering, mean: all in registers).
// The arithmetic is cut into UNITS of (pillar, 4 output channels) dealt round-robin to the 32 lanes, so
// lanes stay busy whatever the number of pillars in the tile.
//
// Pipeline of one warp, iteration i:  prefetch table entries of tile i+2  |  cp.async gather of tile
// i+1's point rows  |  order + mean + units of tile i  |  one TMA tensor store of tile i.
constexpr int EMIT_WARPS = 4;
constexpr int EMIT_THREADS = EMIT_WARPS * 32;
constexpr int STAGE_W = 64;           // staged point rows per tile (a typical tile holds ~10; the rest is read from L2)

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// what a lane knows about its cell of a tile
struct CellState {
    int m, cnt, start, off;     // raw pillar id, arrivals, CSR start, staging offset
    bool occ, staged;           // holds a kept pillar / its rows are (being) staged
};

template <int F, bool ABS, bool DIST, bool BN, int C, bool PFN, bool TMA>
__global__ void __launch_bounds__(EMIT_THREADS, 4)
k_emit(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap zmap, const PathParams p) {
    constexpr int CIN = PFN ? ((ABS ? F : F - 3) + 6 + (DIST ? 1 : 0)) : 1;
    constexpr int RWc = (F + 1 + 3) / 4 * 4;   // F features + the point index, padded to float4
    constexpr int TILE = C * 32;               // floats per tile
    constexpr int ZC = C / 4;                  // channels of the shared zero tile (an empty tile = 4 stores of it)
    constexpr int NT = EMIT_THREADS;
    static_assert(C == 64, "units are 4 of 64 channels; the fallback path maps 2 channels per lane");
    const int Fr = PFN ? F : p.F, RW = PFN ? RWc : p.RW, NV = RW >> 2;

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    float *tiles = reinterpret_cast<float *>(smem_raw);                    // [EMIT_WARPS][TILE]
    float *zerobuf = tiles + EMIT_WARPS * TILE;                            // [ZC*32]
    float *s_W = zerobuf + ZC * 32;                                        // [CIN][C]  k-major Linear weight
    float *s_bn = s_W + CIN * C;                                           // [5][C]    mean, invstd, gamma, beta(bias), pad value
    float *stage_all = s_bn + 5 * C;                                       // [EMIT_WARPS][2][STAGE_W * RW]
    float4 *rec_all = reinterpret_cast<float4 *>(stage_all + EMIT_WARPS * 2 * STAGE_W * RW);   // [EMIT_WARPS][32][2]
    int *s_R = reinterpret_cast<int *>(rec_all + EMIT_WARPS * 64);         // [B+1] raw pillar base per frame
    int *s_K = s_R + (p.B + 1);                                            // [B+1] kept (final) pillar base per frame
    __shared__ unsigned char s_perm_all[EMIT_WARPS][STAGE_W];              // per pillar: arrival position of its rank-th point
    __shared__ int s_bperm_all[EMIT_WARPS][32];                            // warp-per-pillar path

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool canvas_on = PFN && (p.canvas != nullptr);
    float *tile = tiles + warp * TILE;
    float *stage = stage_all + (size_t)warp * 2 * STAGE_W * RW;
    float4 *rec = rec_all + warp * 64;
    unsigned char *s_perm = s_perm_all[warp];
    int *s_bperm = s_bperm_all[warp];

    // ---- one-time setup (the only CTA barriers) ----
    for (int b = tid; b <= p.B; b += NT) s_R[b] = p.frame_raw_base[b];
    if (canvas_on) {
        for (int t = tid; t < ZC * 32; t += NT) zerobuf[t] = 0.f;
        for (int t = tid; t < EMIT_WARPS * TILE; t += NT) tiles[t] = 0.f;
    }
    if (PFN) {
        for (int t = tid; t < CIN * C; t += NT) { const int c = t / CIN, k = t - c * CIN; s_W[k * C + c] = __ldg(p.W + t); }
        for (int c = tid; c < C; c += NT) {
            float y;
            if (BN) {
                const float mu = __ldg(p.bn_m + c), g = __ldg(p.bn_w + c), be = __ldg(p.bn_b + c);
                const float inv = __frcp_rn(__fsqrt_rn(__fadd_rn(__ldg(p.bn_v + c), p.eps)));
                s_bn[c] = mu; s_bn[C + c] = inv; s_bn[2 * C + c] = g; s_bn[3 * C + c] = be;
                // a zero (padded) row still goes through BN + ReLU and joins the max (pillar_vfe.py:37-42)
                y = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(0.f, mu), inv), g), be);
            } else {
                const float be = __ldg(p.bias + c);
                s_bn[c] = 0.f; s_bn[C + c] = 0.f; s_bn[2 * C + c] = 0.f; s_bn[3 * C + c] = be;
                y = __fadd_rn(0.f, be);
            }
            s_bn[4 * C + c] = (y > 0.f || y != y) ? y : 0.f;
        }
    }
    __syncthreads();
    if (tid == 0) {
        int acc = 0;
        for (int b = 0; b < p.B; ++b) {
            s_K[b] = acc;
            const int m = min(s_R[b + 1] - s_R[b], p.max_voxels);
            if (blockIdx.x == 0) p.num_pillars[1 + b] = m;
            acc += m;
        }
        s_K[p.B] = acc;
        if (blockIdx.x == 0) p.num_pillars[0] = acc;
    }
    if (canvas_on && TMA) fence_proxy_async_smem();
    __syncthreads();

    const int tiles_per_row = (p.nx + 31) >> 5;
    const int rows_per_frame = p.nz * p.ny;
    const int n_rows = p.B * rows_per_frame;
    const int P4 = (p.P >> 2) << 2;
    const int maxv = p.max_voxels, Pmax = p.P;
    const float vsx = p.vsize[0], vsy = p.vsize[1], vsz = p.vsize[2], vox = p.voff[0], voy = p.voff[1], voz = p.voff[2];
    const CellEntry *__restrict__ table = p.table;
    const float *__restrict__ grows = p.sorted_rows;
    const unsigned lt = (1u << lane) - 1u;

    // this warp's tile sequence: first tile blockIdx*W + warp, stride gridDim*W; stepping is division free
    TileStep step;
    step.tiles_per_row = tiles_per_row; step.rows_per_frame = rows_per_frame;
    {
        const int stride_tiles = (int)gridDim.x * EMIT_WARPS;
        step.dr = stride_tiles / tiles_per_row; step.dxt = stride_tiles - step.dr * tiles_per_row;
    }
    TilePos cur;
    {
        const int t0 = (int)blockIdx.x * EMIT_WARPS + warp;      // the only divisions: where this warp starts
        cur.r = t0 / tiles_per_row; cur.xt = t0 - cur.r * tiles_per_row;
        cur.b = cur.r / rows_per_frame; cur.zy = cur.r - cur.b * rows_per_frame;
    }
    TilePos nxt = cur;
    step.advance(nxt);
    TilePos nxt2 = nxt;
    step.advance(nxt2);

    auto load_entry = [&](const TilePos &t) -> uint4 {
        const int x = t.xt * 32 + lane;
        // row r = (b*nz + z)*ny + y and the table is [b][z][y][x]: the cell index is r*nx + x
        return (t.r < n_rows && x < p.nx) ? __ldg(reinterpret_cast<const uint4 *>(table + (size_t)t.r * p.nx + x))
                                          : make_uint4(0, 0, 0, 0);
    };
    auto make_state = [&](const uint4 e, const TilePos &t) -> CellState {
        CellState c;
        const int b = (t.r < n_rows) ? t.b : 0;
        c.m = (int)(e.x - 1u); c.cnt = (int)e.y; c.start = (int)e.z;
        c.occ = (e.x != 0u) && (c.m - s_R[b] < maxv);
        const int need = (c.occ && c.cnt <= 32) ? c.cnt : 0;
        int incl = need;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(FULL, incl, d);
            if (lane >= d) incl += o;
        }
        c.off = incl - need;
        c.staged = need > 0 && incl <= STAGE_W;
        return c;
    };
    auto issue_gather = [&](const CellState &c, float *stg) {
        if (c.staged) {
            const float *src = grows + (size_t)c.start * RW;
            float *dst = stg + (size_t)c.off * RW;
            for (int j = 0; j < c.cnt; ++j)
                for (int v = 0; v < NV; ++v) cp_async16(dst + j * RW + 4 * v, src + (size_t)j * RW + 4 * v);
        }
        cp_async_commit();
    };
    // order one unstaged pillar cooperatively: s_bperm[rank] = arrival position of the rank-th smallest index
    auto coop_order = [&](const float *grow, int cnt) {
        if (cnt <= 32) {
            const uint32_t mine = (lane < cnt) ? __float_as_uint(__ldg(grow + (size_t)lane * RW + Fr)) : 0xFFFFFFFFu;
            int rank = 0;
            for (int qq = 0; qq < cnt; ++qq) rank += (__shfl_sync(FULL, mine, qq) < mine) ? 1 : 0;
            if (lane < cnt) s_bperm[rank] = lane;
        } else {
            s_bperm[lane] = select_first32(grow + Fr, RW, cnt, lane);
        }
        __syncwarp();
    };

    CellState st_cur = make_state(load_entry(cur), cur);
    issue_gather(st_cur, stage);
    uint4 e_next = load_entry(nxt);
    unsigned dirty = 0;                  // cells of the tile buffer that hold non-zero columns
    bool store_pending = false;          // a TMA store from the tile buffer may still be reading it

    for (int it = 0; cur.r < n_rows; ++it) {
        const int slot = it & 1;
        const float *stg = stage + (size_t)slot * STAGE_W * RW;
        // ---- next tile: cell states, gather in flight; entries of the one after it ----
        const CellState st_nxt = make_state(e_next, nxt);
        issue_gather(st_nxt, stage + (size_t)(slot ^ 1) * STAGE_W * RW);
        e_next = load_entry(nxt2);

        const CellState c = st_cur;
        const int b = cur.b, zy = cur.zy, x0 = cur.xt * 32;
        const int z = (p.nz == 1) ? 0 : zy / p.ny;
        const int y = zy - z * p.ny;
        const unsigned bal_occ = __ballot_sync(FULL, c.occ);
        if (bal_occ == 0u) {
            // empty tile: four stores of the shared zero tile
            if (canvas_on) {
                if (TMA) {
                    if (lane == 0) {
#pragma unroll
                        for (int q4 = 0; q4 < 4; ++q4) tma_store_3d(&zmap, zerobuf, x0, zy, b * C + q4 * ZC);
                        tma_commit();
                    }
                } else if (x0 + lane < p.nx) {
                    for (int ch = 0; ch < C; ++ch) p.canvas[(((size_t)b * C + ch) * p.ny + y) * p.nx + x0 + lane] = 0.f;
                }
            }
        } else {
            const unsigned bal_st = __ballot_sync(FULL, c.staged);
            const int n_keep = min(c.cnt, Pmax);
            const int f = c.occ ? s_K[b] + (c.m - s_R[b]) : 0;     // final pillar id (first-seen order, frames concatenated)
            if (c.occ) {
                p.num[f] = n_keep;
                *reinterpret_cast<int4 *>(p.coords + 4 * (size_t)f) = make_int4(b, z, y, x0 + lane);
            }
            cp_async_wait<1>();          // this tile's rows have landed (this lane's copies) ...
            __syncwarp();                // ... and every other lane's
            // ---- order the points of multi-point pillars by input index ----
            if (c.staged && c.cnt > 1 && c.cnt <= 4) {            // small: the owning lane ranks them itself
                const float *ib = stg + (size_t)c.off * RW + Fr;
                for (int j = 0; j < c.cnt; ++j) {
                    const uint32_t mine = __float_as_uint(ib[j * RW]);
                    int rank = 0;
                    for (int qq = 0; qq < c.cnt; ++qq) rank += (__float_as_uint(ib[qq * RW]) < mine) ? 1 : 0;
                    s_perm[c.off + rank] = (unsigned char)j;
                }
            }
            unsigned coop = __ballot_sync(FULL, c.staged && c.cnt > 4);   // larger: the warp ranks one pillar at a time
            while (coop) {
                const int o = __ffs(coop) - 1;
                coop &= coop - 1;
                const int cnt_o = __shfl_sync(FULL, c.cnt, o), off_o = __shfl_sync(FULL, c.off, o);
                const uint32_t mine = (lane < cnt_o) ? __float_as_uint(stg[(size_t)(off_o + lane) * RW + Fr]) : 0xFFFFFFFFu;
                int rank = 0;
                for (int qq = 0; qq < cnt_o; ++qq) rank += (__shfl_sync(FULL, mine, qq) < mine) ? 1 : 0;
                if (lane < cnt_o) s_perm[off_o + rank] = (unsigned char)lane;
            }
            __syncwarp();
            if (PFN) {
                // ---- per-pillar record: mean (torch CPU sum order) and bookkeeping, by the owning lane ----
                if (c.staged) {
                    const float *srow = stg + (size_t)c.off * RW;
                    float mx, my, mz;
                    if (c.cnt == 1) {
                        mx = srow[0]; my = srow[1]; mz = srow[2];   // mean of one point is the point (x/1 is exact)
                    } else {
                        SlotSum sum;
                        for (int s2 = 0; s2 < n_keep; ++s2) {
                            const float4 v = *reinterpret_cast<const float4 *>(srow + s_perm[c.off + s2] * RW);
                            sum.add(s2, P4, v.x, v.y, v.z);
                        }
                        const float fn = (float)n_keep;
                        mx = __fdiv_rn(sum.sx(), fn); my = __fdiv_rn(sum.sy(), fn); mz = __fdiv_rn(sum.sz(), fn);
                    }
                    const int k = __popc(bal_st & lt);
                    rec[2 * k] = make_float4(mx, my, mz, __int_as_float(lane | (n_keep << 8) | (c.cnt == 1 ? 0x10000 : 0)));
                    rec[2 * k + 1] = make_float4(__int_as_float(c.off), __int_as_float(f), 0.f, 0.f);
                }
                // the tile buffer: wait until the previous store has read it, then clear the columns it dirtied
                if (canvas_on) {
                    if (TMA && store_pending) {
                        if (lane == 0) tma_wait_read<0>();
                        store_pending = false;
                    }
                    __syncwarp();
                    while (dirty) {
                        const int dc = __ffs(dirty) - 1;
                        dirty &= dirty - 1;
                        tile[swz128(lane, dc)] = 0.f; tile[swz128(lane + 32, dc)] = 0.f;
                    }
                }
                __syncwarp();
                // ---- units: (pillar, 4 channels), dealt round-robin to the lanes ----
                const float cy = __fadd_rn(__fmul_rn((float)y, vsy), voy);   // pillar centre: fl(fl(c*v)+off), two roundings,
                const float cz = __fadd_rn(__fmul_rn((float)z, vsz), voz);   // no FMA (pillar_vfe.py:101-103)
                const int n_units = __popc(bal_st) * 16;
                for (int u = lane; u < n_units; u += 32) {
                    const int k = u >> 4, c0 = (u & 15) * 4;
                    const float4 r0 = rec[2 * k], r1 = rec[2 * k + 1];
                    const int meta = __float_as_int(r0.w);
                    const int cell = meta & 31, nk = (meta >> 8) & 63;
                    const bool single = (meta & 0x10000) != 0;
                    const int off = __float_as_int(r1.x), fid = __float_as_int(r1.y);
                    const float cx = __fadd_rn(__fmul_rn((float)(x0 + cell), vsx), vox);
                    const float *srow = stg + (size_t)off * RWc;
                    float4 w4[CIN];
#pragma unroll
                    for (int kk = 0; kk < CIN; ++kk) w4[kk] = *reinterpret_cast<const float4 *>(s_W + kk * C + c0);
                    const float4 be = *reinterpret_cast<const float4 *>(s_bn + 3 * C + c0);
                    float4 mu, iv, ga;
                    if (BN) {
                        mu = *reinterpret_cast<const float4 *>(s_bn + c0);
                        iv = *reinterpret_cast<const float4 *>(s_bn + C + c0);
                        ga = *reinterpret_cast<const float4 *>(s_bn + 2 * C + c0);
                    }
                    // max over slots as an integer max on the float bits: exact for the non-negative post-ReLU values,
                    // drops negatives and -0 (the ReLU), and lets a NaN (0x7fffffff) win as torch.max does
                    int v0 = 0, v1 = 0, v2 = 0, v3 = 0;
                    if (nk < Pmax) {
                        const float4 pv = *reinterpret_cast<const float4 *>(s_bn + 4 * C + c0);
                        v0 = __float_as_int(pv.x); v1 = __float_as_int(pv.y); v2 = __float_as_int(pv.z); v3 = __float_as_int(pv.w);
                    }
#pragma unroll 1
                    for (int s2 = 0; s2 < nk; ++s2) {
                        const float *rp = srow + (single ? 0 : s_perm[off + s2] * RWc);
                        float row[RWc];
#pragma unroll
                        for (int v = 0; v < RWc / 4; ++v) {
                            const float4 t4 = *reinterpret_cast<const float4 *>(rp + 4 * v);
                            row[4 * v] = t4.x; row[4 * v + 1] = t4.y; row[4 * v + 2] = t4.z; row[4 * v + 3] = t4.w;
                        }
                        float feat[CIN];
                        {
                            int kf = 0;
#pragma unroll
                            for (int qq = ABS ? 0 : 3; qq < F; ++qq) feat[kf++] = row[qq];
                            feat[kf++] = __fsub_rn(row[0], r0.x); feat[kf++] = __fsub_rn(row[1], r0.y); feat[kf++] = __fsub_rn(row[2], r0.z);
                            feat[kf++] = __fsub_rn(row[0], cx); feat[kf++] = __fsub_rn(row[1], cy); feat[kf++] = __fsub_rn(row[2], cz);
                            // torch.norm(xyz, 2, 2) on the CPU: sqrt(fma(z,z, fma(y,y, x*x)))  (pillar_vfe.py:110-112)
                            if (DIST) feat[kf++] = __fsqrt_rn(fmaf(row[2], row[2], fmaf(row[1], row[1], __fmul_rn(row[0], row[0]))));
                        }
                        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
                        for (int kk = 0; kk < CIN; ++kk) {     // Linear: sequential FMA in k order (pillar_vfe.py:37)
                            a0 = fmaf(feat[kk], w4[kk].x, a0); a1 = fmaf(feat[kk], w4[kk].y, a1);
                            a2 = fmaf(feat[kk], w4[kk].z, a2); a3 = fmaf(feat[kk], w4[kk].w, a3);
                        }
                        float y0, y1, y2, y3;
                        if (BN) {                              // BN eval: (((x-mean)*invstd)*gamma)+beta, 4 roundings (:39)
                            y0 = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(a0, mu.x), iv.x), ga.x), be.x);
                            y1 = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(a1, mu.y), iv.y), ga.y), be.y);
                            y2 = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(a2, mu.z), iv.z), ga.z), be.z);
                            y3 = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(a3, mu.w), iv.w), ga.w), be.w);
                        } else {
                            y0 = __fadd_rn(a0, be.x); y1 = __fadd_rn(a1, be.y); y2 = __fadd_rn(a2, be.z); y3 = __fadd_rn(a3, be.w);
                        }
                        v0 = max(v0, __float_as_int(y0)); v1 = max(v1, __float_as_int(y1));
                        v2 = max(v2, __float_as_int(y2)); v3 = max(v3, __float_as_int(y3));
                    }
                    const float4 o = make_float4(__int_as_float(v0), __int_as_float(v1), __int_as_float(v2), __int_as_float(v3));
                    if (p.feats) *reinterpret_cast<float4 *>(p.feats + (size_t)fid * C + c0) = o;
                    if (canvas_on) {
                        tile[swz128(c0, cell)] = o.x; tile[swz128(c0 + 1, cell)] = o.y;
                        tile[swz128(c0 + 2, cell)] = o.z; tile[swz128(c0 + 3, cell)] = o.w;
                    }
                }
            }
            // ---- unstaged pillars (> 32 arrivals, or staging full): the warp takes them one at a time, lanes = channels ----
            unsigned big = bal_occ & ~bal_st;
            while (big) {
                const int o = __ffs(big) - 1;
                big &= big - 1;
                const int cnt_o = __shfl_sync(FULL, c.cnt, o), start_o = __shfl_sync(FULL, c.start, o), f_o = __shfl_sync(FULL, f, o);
                const int nk = min(cnt_o, Pmax);
                const float *grow = grows + (size_t)start_o * RW;
                coop_order(grow, cnt_o);
                if (p.voxels) {
                    float *vo = p.voxels + (size_t)f_o * Pmax * Fr;
                    for (int t = lane; t < Pmax * Fr; t += 32) {
                        const int s2 = t / Fr, kk = t - s2 * Fr;
                        vo[t] = (s2 < nk) ? __ldg(grow + (size_t)s_bperm[s2] * RW + kk) : 0.f;
                    }
                }
                if (PFN) {
                    const float cx = __fadd_rn(__fmul_rn((float)(x0 + o), vsx), vox);
                    const float cy = __fadd_rn(__fmul_rn((float)y, vsy), voy);
                    const float cz = __fadd_rn(__fmul_rn((float)z, vsz), voz);
                    SlotSum sum;
                    for (int s2 = 0; s2 < nk; ++s2) {
                        const float4 v = __ldg(reinterpret_cast<const float4 *>(grow + (size_t)s_bperm[s2] * RWc));
                        sum.add(s2, P4, v.x, v.y, v.z);
                    }
                    const float fn = (float)nk;
                    const float mx = __fdiv_rn(sum.sx(), fn), my = __fdiv_rn(sum.sy(), fn), mz = __fdiv_rn(sum.sz(), fn);
                    int v0 = (nk < Pmax) ? __float_as_int(s_bn[4 * C + lane]) : 0;
                    int v1 = (nk < Pmax) ? __float_as_int(s_bn[4 * C + lane + 32]) : 0;
                    for (int s2 = 0; s2 < nk; ++s2) {
                        const float4 *r4 = reinterpret_cast<const float4 *>(grow + (size_t)s_bperm[s2] * RWc);
                        float row[RWc];
#pragma unroll
                        for (int v = 0; v < RWc / 4; ++v) {
                            const float4 t4 = __ldg(r4 + v);
                            row[4 * v] = t4.x; row[4 * v + 1] = t4.y; row[4 * v + 2] = t4.z; row[4 * v + 3] = t4.w;
                        }
                        float feat[CIN];
                        {
                            int kf = 0;
#pragma unroll
                            for (int qq = ABS ? 0 : 3; qq < F; ++qq) feat[kf++] = row[qq];
                            feat[kf++] = __fsub_rn(row[0], mx); feat[kf++] = __fsub_rn(row[1], my); feat[kf++] = __fsub_rn(row[2], mz);
                            feat[kf++] = __fsub_rn(row[0], cx); feat[kf++] = __fsub_rn(row[1], cy); feat[kf++] = __fsub_rn(row[2], cz);
                            if (DIST) feat[kf++] = __fsqrt_rn(fmaf(row[2], row[2], fmaf(row[1], row[1], __fmul_rn(row[0], row[0]))));
                        }
                        float a0 = 0.f, a1 = 0.f;
#pragma unroll
                        for (int kk = 0; kk < CIN; ++kk) {
                            a0 = fmaf(feat[kk], s_W[kk * C + lane], a0);
                            a1 = fmaf(feat[kk], s_W[kk * C + lane + 32], a1);
                        }
                        float y0, y1;
                        if (BN) {
                            y0 = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(a0, s_bn[lane]), s_bn[C + lane]), s_bn[2 * C + lane]), s_bn[3 * C + lane]);
                            y1 = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(a1, s_bn[lane + 32]), s_bn[C + lane + 32]), s_bn[2 * C + lane + 32]), s_bn[3 * C + lane + 32]);
                        } else {
                            y0 = __fadd_rn(a0, s_bn[3 * C + lane]); y1 = __fadd_rn(a1, s_bn[3 * C + lane + 32]);
                        }
                        v0 = max(v0, __float_as_int(y0)); v1 = max(v1, __float_as_int(y1));
                    }
                    if (p.feats) { p.feats[(size_t)f_o * C + lane] = __int_as_float(v0); p.feats[(size_t)f_o * C + lane + 32] = __int_as_float(v1); }
                    if (canvas_on) { tile[swz128(lane, o)] = __int_as_float(v0); tile[swz128(lane + 32, o)] = __int_as_float(v1); }
                }
                __syncwarp();
            }
            // ---- optional contract output for the staged pillars: padded voxels [M, P, F], coalesced ----
            if (p.voxels) {
                unsigned todo = bal_st;
                while (todo) {
                    const int o = __ffs(todo) - 1;
                    todo &= todo - 1;
                    const int cnt_o = __shfl_sync(FULL, c.cnt, o), off_o = __shfl_sync(FULL, c.off, o), f_o = __shfl_sync(FULL, f, o);
                    const int nk = min(cnt_o, Pmax);
                    float *vo = p.voxels + (size_t)f_o * Pmax * Fr;
                    const float *srow = stg + (size_t)off_o * RW;
                    for (int t = lane; t < Pmax * Fr; t += 32) {
                        const int s2 = t / Fr, kk = t - s2 * Fr;
                        vo[t] = (s2 < nk) ? srow[((cnt_o == 1) ? 0 : (int)s_perm[off_o + s2]) * RW + kk] : 0.f;
                    }
                }
            }
            // ---- the tile goes out in one piece ----
            if (canvas_on) {
                dirty = bal_occ;
                if (TMA) {
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) { tma_store_3d(&tmap, tile, x0, zy, b * C); tma_commit(); }
                    store_pending = true;
                } else {
                    __syncwarp();
                    if (x0 + lane < p.nx)
                        for (int ch = 0; ch < C; ++ch) p.canvas[(((size_t)b * C + ch) * p.ny + y) * p.nx + x0 + lane] = tile[swz128(ch, lane)];
                }
            }
            __syncwarp();
        }
        st_cur = st_nxt;
        cur = nxt; nxt = nxt2; step.advance(nxt2);
    }
    cp_async_wait<0>();
    if (canvas_on && TMA && lane == 0) tma_wait_read<0>();
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = []() -> EncodeTiledFn {
        void *f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return nullptr;
        return reinterpret_cast<EncodeTiledFn>(f);
    }();
    return fn;
}

// canvas [B*C, ny, nx] fp32, box = 32 cells x 1 row x C channels, 128-byte swizzle
int make_canvas_map(CUtensorMap *map, float *canvas, int B, int C, int ny, int nx, int box_c) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return HGSF_ERR_DRIVER;
    const cuuint64_t gdim[3] = {(cuuint64_t)nx, (cuuint64_t)ny, (cuuint64_t)B * C};
    const cuuint64_t gstr[2] = {(cuuint64_t)nx * 4, (cuuint64_t)nx * ny * 4};
    const cuuint32_t box[3] = {32, 1, (cuuint32_t)box_c};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, canvas, gdim, gstr, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? HGSF_OK : HGSF_ERR_DRIVER;
}

int sm_count() {
    static int n = []() {
        int dev = 0, v = 148;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        return v;
    }();
    return n;
}

template <int F, bool ABS, bool DIST, int C, bool PFN>
static int launch_emit_t(const PathParams &p, cudaStream_t stream) {
    const bool canvas_on = PFN && p.canvas;
    const bool tma = canvas_on && (p.nx % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.canvas) & 15) == 0) && C <= 256;
    CUtensorMap map, zmap;
    memset(&map, 0, sizeof(map));
    memset(&zmap, 0, sizeof(zmap));
    if (tma) {
        int st = make_canvas_map(&map, p.canvas, p.B, C, p.ny, p.nx, C);
        if (st == HGSF_OK) st = make_canvas_map(&zmap, p.canvas, p.B, C, p.ny, p.nx, C / 4);
        if (st != HGSF_OK) return st;
    }
    const int cin = PFN ? p.Cin : 1;
    const size_t smem = sizeof(float) * (EMIT_WARPS * C * 32 + (C / 4) * 32 + (size_t)cin * C + 5 * C +
                                         EMIT_WARPS * 2 * STAGE_W * (size_t)p.RW + EMIT_WARPS * 64 * 4) +
                        sizeof(int) * 2 * (size_t)(p.B + 1);
    const long long n_tiles = ((long long)p.B * p.nz * p.ny * ((p.nx + 31) / 32) + EMIT_WARPS - 1) / EMIT_WARPS;   // CTA-loads of tiles
    if (n_tiles == 0) return HGSF_OK;
    const bool bn = p.bn_w != nullptr;
    auto go = [&](auto kern) -> int {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        int per_sm = 1;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, EMIT_THREADS, smem);
        if (e != cudaSuccess) return (int)e;
        if (per_sm < 1) per_sm = 1;
        const long long grid = std::min<long long>(n_tiles, (long long)sm_count() * per_sm);
        kern<<<(unsigned)grid, EMIT_THREADS, smem, stream>>>(map, zmap, p);
        return (int)cudaGetLastError();
    };
    if constexpr (PFN) {
        if (tma) return bn ? go(k_emit<F, ABS, DIST, true, C, true, true>) : go(k_emit<F, ABS, DIST, false, C, true, true>);
        return bn ? go(k_emit<F, ABS, DIST, true, C, true, false>) : go(k_emit<F, ABS, DIST, false, C, true, false>);
    }
    return go(k_emit<F, ABS, DIST, true, C, false, false>);
}

template <int C>
static int launch_emit_pfn(const PathParams &p, bool abs_xyz, bool dist, cudaStream_t s) {
#define HGSF_CASE(FV, A, D) if (p.F == FV && abs_xyz == A && dist == D) return launch_emit_t<FV, A, D, C, true>(p, s);
    HGSF_CASE(4, true, false) HGSF_CASE(5, true, false) HGSF_CASE(6, true, false) HGSF_CASE(7, true, false)
    HGSF_CASE(8, true, false) HGSF_CASE(7, false, false) HGSF_CASE(8, false, false)
    HGSF_CASE(7, true, true) HGSF_CASE(8, true, true)
#undef HGSF_CASE
    return HGSF_ERR_UNSUPPORTED;
}

static int launch_emit_plain(const PathParams &p, cudaStream_t s) {
    return launch_emit_t<4, true, false, 64, false>(p, s);   // F / RW are read from the params when PFN is off
}

// ---- optional per-launch timing of k_emit (bench.py's roofline leg) ---------------------------------
// A ring of CUDA event pairs recorded on the launching stream around the k_emit launch.  Off by default.
struct EmitTiming {
    std::vector<cudaEvent_t> ev;   // 2 * capacity
    int capacity = 0, count = 0;
};
static thread_local EmitTiming g_timing;

int emit_timing_begin(int capacity) {
    for (cudaEvent_t e : g_timing.ev) cudaEventDestroy(e);
    g_timing.ev.clear();
    g_timing.capacity = g_timing.count = 0;
    if (capacity <= 0) return HGSF_OK;
    g_timing.ev.resize(2 * (size_t)capacity);
    for (auto &e : g_timing.ev) {
        const cudaError_t st = cudaEventCreate(&e);
        if (st != cudaSuccess) return (int)st;
    }
    g_timing.capacity = capacity;
    return HGSF_OK;
}

int emit_timing_collect(float *ms, int n) {
    const int have = g_timing.count < g_timing.capacity ? g_timing.count : g_timing.capacity;
    int out = 0;
    for (int i = 0; i < have && out < n; ++i, ++out) {
        cudaError_t st = cudaEventSynchronize(g_timing.ev[2 * i + 1]);
        if (st == cudaSuccess) st = cudaEventElapsedTime(ms + out, g_timing.ev[2 * i], g_timing.ev[2 * i + 1]);
        if (st != cudaSuccess) return -(int)st;
    }
    g_timing.count = 0;
    return out;
}

int launch_pillar_path(const PathParams &p, bool with_pfn, bool abs_xyz, bool dist, size_t zero_bytes, void *zero_base,
                       cudaStream_t stream, int *launches) {
    int nl = 0;   // kernels only; the workspace memset below is not counted
    cudaError_t e = cudaMemsetAsync(zero_base, 0, zero_bytes, stream);
    if (e != cudaSuccess) return (int)e;
    if (p.n > 0) {
        const unsigned g = (unsigned)((p.n + 255) / 256);
        k_count<<<g, 256, 0, stream>>>(p);
        if ((e = cudaGetLastError()) != cudaSuccess) return (int)e;
        k_scan<<<(unsigned)((p.n + SCAN_TILE - 1) / SCAN_TILE), SCAN_THREADS, 0, stream>>>(p);
        if ((e = cudaGetLastError()) != cudaSuccess) return (int)e;
        k_fill<<<g, 256, 0, stream>>>(p);
        if ((e = cudaGetLastError()) != cudaSuccess) return (int)e;
        nl += 3;
    }
    const bool timed = g_timing.capacity > 0 && g_timing.count < g_timing.capacity;
    if (timed) cudaEventRecord(g_timing.ev[2 * g_timing.count], stream);
    int st;
    if (with_pfn) {
        if (p.C == 64) st = launch_emit_pfn<64>(p, abs_xyz, dist, stream);
        else st = HGSF_ERR_UNSUPPORTED;
    } else {
        st = launch_emit_plain(p, stream);
    }
    if (st != HGSF_OK) return st;
    if (timed) cudaEventRecord(g_timing.ev[2 * g_timing.count++ + 1], stream);
    ++nl;
    if (launches) *launches = nl;
    return HGSF_OK;
}

}  // namespace hgsf
