// pillar_path.cu -- points -> pillars (first-seen order) -> decorate + PFN + max -> BEV canvas.
//
// Five kernels on one stream, no host sync, no allocation:
//   k_count  (1 thread / point)  cell key; per cell: min point index + count (warp-aggregated atomics
//                                into the direct-address cell table)
//   k_scan   (1024 points / CTA) single-pass decoupled look-back scan over points: a point that is the
//                                first of its cell gets (raw pillar id, CSR start) = exclusive prefix of
//                                (first-flags, cell counts) -> pillar ids come out in first-seen order;
//                                it also writes the pillar's record {start, cnt, b, z, y, x}
//   k_fill   (1 thread / point)  copies each point's features (+ its index) to its pillar's CSR segment
//   k_pfn    (pillar major)      a warp takes 32 consecutive pillars (their CSR rows are contiguous): orders each
//                                pillar's points by index, keeps the first P, decorates, runs the PFN with the
//                                weights in registers, takes the max; writes voxel_coords / voxel_num_points /
//                                pillar_features rows (and the padded voxels tensor on request)
//   k_canvas (tile major)        warps walk the canvas in tiles of 32 cells x C channels: cell table -> pillar row
//                                -> tile in shared memory -> one TMA tensor store per tile, zeros included
//
// What it reproduces (file:line under the reference):
//   spconv Point2VoxelCPU3d.point_to_voxel as called by pcdet/datasets/processor/data_processor.py:55
//   collate_batch voxel keys                      pcdet/datasets/dataset.py:232-244
//   PillarVFE.forward + PFNLayer.forward          pcdet/models/backbones_3d/vfe/pillar_vfe.py:29-49,94-123
//   PointPillarScatter.forward                    pcdet/models/backbones_2d/map_to_bev/pointpillar_scatter.py:14-41
// The fp32 operation order is the one the CPU reference was measured to use (oracle/pillar_oracle.c).
#include <cooperative_groups.h>

#include "pillar_path.cuh"
#include "pfn.cuh"
#include "contract_ops.cuh"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace cg = cooperative_groups;

namespace hgsf {

// ------------------------------------------------------------------------------------------------
// k_front : ONE cooperative kernel for the whole front end (grid barriers instead of launches)
//   phase 0  zero the cell table
//   phase 1  count : 1 thread / point.  cell key; per cell min point index + count through warp-aggregated atomics
//   phase 2  two scans, each two level (every CTA reduces its contiguous slice, barrier, then scans it with the sum of
//            the slices before it as carry-in):
//            (a) over POINTS, of "is the first point of its cell": the exclusive prefix at a first point is the raw
//                pillar id -> pillar ids come out in first-seen order without a sort.  Also writes the pillar records
//                and the raw id at each frame start.
//            (b) over CELLS in table order (b, z, y, x), of the cell counts: the exclusive prefix is the cell's CSR
//                start -> the point rows of a 32-cell canvas tile are CONTIGUOUS in sorted_rows.
//   phase 3  fill  : 1 thread / point: copies the point's features (+ its index) to its cell's CSR segment
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

// phase 1, split in three so that two points per thread can be in flight at once:
//   point_key     loads the point, derives frame_offsets on the fly, returns the cell key (-1 = outside the grid)
//   count_issue   warp-aggregated atomics: lanes of the same cell elect the lowest lane (= lowest point index), which
//                 adds the group size to the cell count and maxes the inverted index into the tag
//   count_finish  broadcasts the leader's base and stores key + arrival rank
// every lane of the warp calls them (i >= n for the padding lanes)
__device__ __forceinline__ int point_key(const PathParams &p, int i) {
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
    return key;
}
__device__ __forceinline__ unsigned count_issue(const PathParams &p, int i, int key, int lane, int &leader, int &rank) {
    const unsigned peers = __match_any_sync(FULL, key);
    leader = __ffs(peers) - 1;
    rank = __popc(peers & ((1u << lane) - 1u));
    unsigned base = 0;
    if (lane == leader && key >= 0) {
        base = atomicAdd(p.cell_cnt + key, (unsigned)__popc(peers));
        atomicMax(p.cell_tag + key, 0xFFFFFFFFu - (unsigned)i);
    }
    return base;
}
__device__ __forceinline__ void count_finish(const PathParams &p, int i, int key, unsigned base, int leader, int rank) {
    base = __shfl_sync(FULL, base, leader);
    if (i < p.n) {
        p.key[i] = key;
        p.arrival[i] = base + (unsigned)rank;
    }
}

// phase 3 body: U independent points per thread, staged by hand (keys -> table entries -> source rows -> stores) so that
// the dependent loads of the U points overlap; the compiler cannot hoist them itself across the early exits and stores
template <int U>
__device__ __forceinline__ void fill_points(const PathParams &p, long long i0, long long step) {
    int idx[U], key[U];
    bool on[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const long long i = i0 + u * step;
        on[u] = i < p.n;
        idx[u] = (int)i;
        key[u] = on[u] ? p.key[i] : -1;
    }
    uint32_t tag[U], start[U], arr[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        on[u] = on[u] && key[u] >= 0;
        tag[u] = start[u] = arr[u] = 0u;
        if (on[u]) { tag[u] = p.cell_tag[key[u]]; start[u] = p.cell_start[key[u]]; arr[u] = p.arrival[idx[u]]; }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
        if (on[u]) {
            const int b = key[u] / p.cells;
            const int local = (int)(tag[u] - 1u) - p.frame_raw_base[b];
            on[u] = local < p.max_voxels;             // pillar beyond max_voxels: never created
        }
    }
    for (int k = 0; k < p.RW; k += 4) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (!on[u]) continue;
            const float *src = p.pts + (size_t)idx[u] * p.stride + p.xyz_col;
            // F features, then the point index (slot F) that k_emit / k_pfn order the pillar by
            const float fi = __int_as_float(idx[u]);
            v[u].x = (k + 0 < p.F) ? __ldg(src + k + 0) : (k + 0 == p.F ? fi : 0.f);
            v[u].y = (k + 1 < p.F) ? __ldg(src + k + 1) : (k + 1 == p.F ? fi : 0.f);
            v[u].z = (k + 2 < p.F) ? __ldg(src + k + 2) : (k + 2 == p.F ? fi : 0.f);
            v[u].w = (k + 3 < p.F) ? __ldg(src + k + 3) : (k + 3 == p.F ? fi : 0.f);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (!on[u]) continue;
            const size_t pos = (size_t)start[u] + arr[u];
            reinterpret_cast<float4 *>(p.sorted_rows + pos * p.RW)[k >> 2] = v[u];
        }
    }
}

constexpr int FRONT_THREADS = 256;
#ifndef HGSF_FRONT_ILP
#define HGSF_FRONT_ILP 2
#endif
constexpr int FRONT_ILP = HGSF_FRONT_ILP;         // independent points per thread in the count and fill phases
constexpr int SCAN_ITEMS = SCAN_TILE / FRONT_THREADS;   // 4
__device__ __forceinline__ uint64_t pack2(uint32_t pillars, uint32_t points) { return ((uint64_t)pillars << 32) | points; }

// block-wide sum of a 64-bit value; every thread gets the total
__device__ __forceinline__ uint64_t block_sum(uint64_t v, uint64_t *s_warp, int lane, int warp) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(FULL, v, d);
    __syncthreads();
    if (lane == 0) s_warp[warp] = v;
    __syncthreads();
    uint64_t t = 0;
#pragma unroll
    for (int w = 0; w < FRONT_THREADS / 32; ++w) t += s_warp[w];
    return t;
}

__global__ void __launch_bounds__(FRONT_THREADS) k_front(const PathParams p) {
    cg::grid_group grid = cg::this_grid();
#ifdef HGSF_PHASE_TIMES
    auto stamp = [&](int k) { if (blockIdx.x == 0 && threadIdx.x == 0) { uint64_t t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); p.scan_desc[4000 + k] = t; } };
#else
    auto stamp = [&](int) {};
#endif
    stamp(0);
    __shared__ uint64_t s_warp[FRONT_THREADS / 32];
    __shared__ uint32_t s_excl[SCAN_TILE];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long gtid = (long long)blockIdx.x * FRONT_THREADS + tid, nthr = (long long)gridDim.x * FRONT_THREADS;

    // ---- phase 0: zero the cell table (and k_pfn's chunk ticket) ----
    if (gtid == 0) { p.ticket[0] = 0u; p.ticket[32] = 0u; }     // k_pfn's chunk ticket, k_emit's tile ticket (own 128-byte line)
    {
        uint4 *t4 = reinterpret_cast<uint4 *>(p.cell_tag);  // tag, cnt, start: three arrays back to back
        const long long n4 = (long long)(p.table_bytes >> 4);
        for (long long i = gtid; i < n4; i += nthr) t4[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    grid.sync();
    stamp(1);
    // ---- phase 1: count ----
    {
        const long long n_pad = ((long long)p.n + 31) & ~31ll;
        // FRONT_ILP independent points per thread: their loads, divides and atomics overlap (the phase is a chain of L2 round
        // trips).  `on` is warp-uniform (n_pad and nthr are multiples of 32), so the warp-wide match stays convergent.
        for (long long i = gtid; i < n_pad; i += FRONT_ILP * nthr) {
            int k[FRONT_ILP], l[FRONT_ILP], r[FRONT_ILP];
            unsigned bs[FRONT_ILP];
            bool on[FRONT_ILP];
#pragma unroll
            for (int u = 0; u < FRONT_ILP; ++u) {
                on[u] = i + u * nthr < n_pad;
                k[u] = on[u] ? point_key(p, (int)(i + u * nthr)) : -1;
            }
#pragma unroll
            for (int u = 0; u < FRONT_ILP; ++u) {
                l[u] = 0; r[u] = 0; bs[u] = 0;
                if (on[u]) bs[u] = count_issue(p, (int)(i + u * nthr), k[u], lane, l[u], r[u]);
            }
#pragma unroll
            for (int u = 0; u < FRONT_ILP; ++u)
                if (on[u]) count_finish(p, (int)(i + u * nthr), k[u], bs[u], l[u], r[u]);
        }
    }
    grid.sync();
    stamp(2);
    // ---- phase 2: scan.  CTA c owns points [lo, hi), a whole number of 1024-point tiles ----
    const int tiles_total = (p.n + SCAN_TILE - 1) / SCAN_TILE;
    const int tiles_per_cta = (tiles_total + (int)gridDim.x - 1) / (int)gridDim.x;
    const int lo = min((int)blockIdx.x * tiles_per_cta, tiles_total) * SCAN_TILE;
    const int hi = min(((int)blockIdx.x + 1) * tiles_per_cta, tiles_total) * SCAN_TILE;
    auto flags_of = [&](int base, int (&keys)[SCAN_ITEMS], uint32_t (&flag)[SCAN_ITEMS], uint32_t (&cnt)[SCAN_ITEMS]) -> uint64_t {
        uint64_t local = 0;
#pragma unroll
        for (int j = 0; j < SCAN_ITEMS; ++j) keys[j] = (base + j < p.n) ? p.key[base + j] : -1;
#pragma unroll
        for (int j = 0; j < SCAN_ITEMS; ++j) {
            flag[j] = 0; cnt[j] = 0;
            if (keys[j] >= 0) {
                flag[j] = (p.cell_tag[keys[j]] == 0xFFFFFFFFu - (uint32_t)(base + j)) ? 1u : 0u;
                cnt[j] = flag[j] ? p.cell_cnt[keys[j]] : 0u;
            }
            local += pack2(flag[j], cnt[j]);
        }
        return local;
    };
    {
        uint64_t mine = 0;
        for (int t0 = lo; t0 < hi; t0 += SCAN_TILE) {
            int keys[SCAN_ITEMS];
            uint32_t flag[SCAN_ITEMS], cnt[SCAN_ITEMS];
            mine += flags_of(t0 + tid * SCAN_ITEMS, keys, flag, cnt);
        }
        const uint64_t total = block_sum(mine, s_warp, lane, warp);
        if (tid == 0) p.scan_desc[blockIdx.x] = total;
    }
    stamp(7);
    // (b) cells: CTA c owns cells [clo, chi), a whole number of 1024-cell tiles
    const long long n_cells = (long long)p.B * p.cells;
    const long long ctiles_total = (n_cells + SCAN_TILE - 1) / SCAN_TILE;
    const long long ctiles_per_cta = (ctiles_total + gridDim.x - 1) / gridDim.x;
    const long long clo = min((long long)blockIdx.x * ctiles_per_cta, ctiles_total) * SCAN_TILE;
    const long long chi = min(((long long)blockIdx.x + 1) * ctiles_per_cta, ctiles_total) * SCAN_TILE;
    {
        uint64_t mine = 0;
        for (long long t0 = clo; t0 < chi; t0 += SCAN_TILE) {
            const long long c = t0 + tid * SCAN_ITEMS;          // 4 consecutive counts = one 16-byte load (arrays are padded)
            if (c < n_cells) {
                const uint4 v = *reinterpret_cast<const uint4 *>(p.cell_cnt + c);
                mine += v.x + v.y + v.z + v.w;
            }
        }
        const uint64_t total = block_sum(mine, s_warp, lane, warp);
        if (tid == 0) p.scan_desc[2048 + blockIdx.x] = total;
    }
    grid.sync();
    stamp(3);
    {
        uint64_t before = 0;
        for (int c = tid; c < (int)blockIdx.x; c += FRONT_THREADS) before += p.scan_desc[c];
        uint64_t carry = block_sum(before, s_warp, lane, warp);     // (pillars, points) of all slices before this one
        for (int t0 = lo; t0 < hi; t0 += SCAN_TILE) {
            const int base = t0 + tid * SCAN_ITEMS;
            int keys[SCAN_ITEMS];
            uint32_t flag[SCAN_ITEMS], cnt[SCAN_ITEMS];
            const uint64_t local = flags_of(base, keys, flag, cnt);
            // block-wide exclusive scan of `local`
            uint64_t incl = local;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint64_t o = __shfl_up_sync(FULL, incl, d);
                if (lane >= d) incl += o;
            }
            __syncthreads();
            if (lane == 31) s_warp[warp] = incl;
            __syncthreads();
            uint64_t warp_off = 0, tile_total = 0;
#pragma unroll
            for (int w = 0; w < FRONT_THREADS / 32; ++w) {
                const uint64_t v = s_warp[w];
                if (w < warp) warp_off += v;
                tile_total += v;
            }
            uint64_t run = carry + warp_off + (incl - local);
#pragma unroll
            for (int j = 0; j < SCAN_ITEMS; ++j) {
                const uint32_t pillars = (uint32_t)(run >> 32);
                s_excl[tid * SCAN_ITEMS + j] = pillars;
                if (flag[j]) {
                    p.cell_tag[keys[j]] = pillars + 1u;   // raw pillar id + 1 (disjoint from the 0xFFFFFFFF-i values phase 1 left)
                    // the pillar's record, in first-seen order (one thread per pillar pays the divisions)
                    const int key = keys[j];
                    const int b = key / p.cells, rem = key - b * p.cells;
                    const int plane = p.ny * p.nx;
                    const int z = rem / plane, rem2 = rem - z * plane;
                    const int y = rem2 / p.nx, x = rem2 - y * p.nx;
                    p.prec[pillars] = make_int4(key, (int)cnt[j], (b << 16) | z, (y << 16) | x);
                }
                run += pack2(flag[j], cnt[j]);
            }
            __syncthreads();
            // raw pillar id at each frame start
            const bool last = (t0 + SCAN_TILE >= p.n);
            const uint32_t total = (uint32_t)((carry + tile_total) >> 32);
            for (int b = tid; b <= p.B; b += FRONT_THREADS) {
                const int o = p.frame_offsets[b];
                if (o >= t0 && o < t0 + SCAN_TILE && o < p.n) p.frame_raw_base[b] = (int32_t)s_excl[o - t0];
                else if (last && o >= p.n) p.frame_raw_base[b] = (int32_t)total;
            }
            carry += tile_total;
        }
        if (p.n == 0 && blockIdx.x == 0)
            for (int b = tid; b <= p.B; b += FRONT_THREADS) p.frame_raw_base[b] = 0;
    }
    stamp(6);
    {
        // (b) cells: CSR start of every cell (empty ones too: a tile's row span is start[first cell] .. start[last]+cnt)
        uint64_t before = 0;
        for (int c = tid; c < (int)blockIdx.x; c += FRONT_THREADS) before += p.scan_desc[2048 + c];
        uint32_t carry = (uint32_t)block_sum(before, s_warp, lane, warp);
        for (long long t0 = clo; t0 < chi; t0 += SCAN_TILE) {
            const long long c0 = t0 + tid * SCAN_ITEMS;
            uint32_t cn[SCAN_ITEMS], local = 0;
#pragma unroll
            for (int j = 0; j < SCAN_ITEMS; ++j) cn[j] = 0u;
            if (c0 < n_cells) {
                const uint4 v = *reinterpret_cast<const uint4 *>(p.cell_cnt + c0);
                cn[0] = v.x; cn[1] = v.y; cn[2] = v.z; cn[3] = v.w;
            }
#pragma unroll
            for (int j = 0; j < SCAN_ITEMS; ++j) local += cn[j];
            uint32_t incl = local;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t o = __shfl_up_sync(FULL, incl, d);
                if (lane >= d) incl += o;
            }
            __syncthreads();
            if (lane == 31) s_warp[warp] = incl;
            __syncthreads();
            uint32_t warp_off = 0, tile_total = 0;
#pragma unroll
            for (int w = 0; w < FRONT_THREADS / 32; ++w) {
                const uint32_t v = (uint32_t)s_warp[w];
                if (w < warp) warp_off += v;
                tile_total += v;
            }
            const uint32_t run = carry + warp_off + (incl - local);
            if (c0 < n_cells)
                *reinterpret_cast<uint4 *>(p.cell_start + c0) = make_uint4(run, run + cn[0], run + cn[0] + cn[1], run + cn[0] + cn[1] + cn[2]);
            carry += tile_total;
        }
    }
    grid.sync();
    stamp(4);
    // ---- phase 3: fill ----
    for (long long i = gtid; i < p.n; i += FRONT_ILP * nthr) fill_points<FRONT_ILP>(p, i, nthr);
    stamp(5);
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

// ---- k_pfn --------------------------------------------------------------------------------------
// Pillar major.  A warp takes a chunk of 32 consecutive pillars (raw first-seen ids m0 .. m0+31); because the CSR
// start is the prefix sum over that same order, the chunk's point rows are one contiguous span of sorted_rows and
// stream through L1.  Lane j OWNS pillar m0+j for the bookkeeping (ordering by point index, first P kept, mean in
// torch's summation order, voxel_coords / voxel_num_points).  The arithmetic is cut into UNITS of (pillar, 4 output
// channels): lane l always computes channels 4*(l&15) .. +3 -- so its 13/14 Linear weight float4s and BatchNorm
// constants stay in REGISTERS for the whole kernel -- for pillars (l>>4) + 2*it, it = 0..15.  CUDA-core FMA: a
// 13x64 contraction is far below a tensor-core tile.
constexpr int PFN_WARPS = 4;
constexpr int PFN_THREADS = PFN_WARPS * 32;
constexpr int SMALL_CNT = 6;          // up to this many arrivals the owning lane ranks them itself

template <int F, bool ABS, bool DIST, bool BN, bool PFN>
__global__ void __launch_bounds__(PFN_THREADS, 4) k_pfn(const PathParams p) {
    constexpr int C = 64;
    constexpr int CIN = PFN ? ((ABS ? F : F - 3) + 6 + (DIST ? 1 : 0)) : 1;
    constexpr int RWc = (F + 1 + 3) / 4 * 4;   // F features + the point index, padded to float4
    const int Fr = PFN ? F : p.F, RW = PFN ? RWc : p.RW;

    extern __shared__ __align__(16) uint8_t smem_raw[];
    int *s_R = reinterpret_cast<int *>(smem_raw);                          // [B+1] raw pillar base per frame
    int *s_K = s_R + (p.B + 1);                                            // [B+1] kept (final) pillar base per frame
    __shared__ float4 s_rec_all[PFN_WARPS][32][2];                         // per pillar of the chunk: mean + bookkeeping
    __shared__ unsigned char s_perm_all[PFN_WARPS][32][32];                // arrival position of the pillar's rank-th point
    __shared__ int s_bperm_all[PFN_WARPS][32];                             // same for a pillar with > 32 arrivals

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float4(*rec)[2] = s_rec_all[warp];
    unsigned char(*perm)[32] = s_perm_all[warp];
    int *bperm = s_bperm_all[warp];

    for (int b = tid; b <= p.B; b += PFN_THREADS) s_R[b] = p.frame_raw_base[b];
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
    __syncthreads();

    // this lane's 4 channels: Linear rows and BatchNorm constants, in registers for the whole kernel
    const int c0 = 4 * (lane & 15);
    uint64_t w01[CIN], w23[CIN];           // channel pairs (c0, c0+1), (c0+2, c0+3): one FFMA2 each per input feature
    float4 mu = make_float4(0.f, 0.f, 0.f, 0.f), iv = mu, ga = mu, be = mu, pv = mu;
    if (PFN) {
#pragma unroll
        for (int k = 0; k < CIN; ++k) {
            w01[k] = pack_f2(__ldg(p.W + (c0 + 0) * CIN + k), __ldg(p.W + (c0 + 1) * CIN + k));
            w23[k] = pack_f2(__ldg(p.W + (c0 + 2) * CIN + k), __ldg(p.W + (c0 + 3) * CIN + k));
        }
        float bnv[5][4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = c0 + j;
            float y;
            if (BN) {
                bnv[0][j] = __ldg(p.bn_m + c);
                bnv[1][j] = __frcp_rn(__fsqrt_rn(__fadd_rn(__ldg(p.bn_v + c), p.eps)));
                bnv[2][j] = __ldg(p.bn_w + c);
                bnv[3][j] = __ldg(p.bn_b + c);
                // a zero (padded) row still goes through BN + ReLU and joins the max (pillar_vfe.py:37-42)
                y = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(0.f, bnv[0][j]), bnv[1][j]), bnv[2][j]), bnv[3][j]);
            } else {
                bnv[0][j] = bnv[1][j] = bnv[2][j] = 0.f;
                bnv[3][j] = __ldg(p.bias + c);
                y = __fadd_rn(0.f, bnv[3][j]);
            }
            bnv[4][j] = (y > 0.f || y != y) ? y : 0.f;
        }
        mu = make_float4(bnv[0][0], bnv[0][1], bnv[0][2], bnv[0][3]); iv = make_float4(bnv[1][0], bnv[1][1], bnv[1][2], bnv[1][3]);
        ga = make_float4(bnv[2][0], bnv[2][1], bnv[2][2], bnv[2][3]); be = make_float4(bnv[3][0], bnv[3][1], bnv[3][2], bnv[3][3]);
        pv = make_float4(bnv[4][0], bnv[4][1], bnv[4][2], bnv[4][3]);
    }
    const uint64_t mu01 = pack_f2(mu.x, mu.y), mu23 = pack_f2(mu.z, mu.w), iv01 = pack_f2(iv.x, iv.y), iv23 = pack_f2(iv.z, iv.w),
                   ga01 = pack_f2(ga.x, ga.y), ga23 = pack_f2(ga.z, ga.w);

    const int P4 = (p.P >> 2) << 2;
    const int maxv = p.max_voxels, Pmax = p.P;
    const float vsx = p.vsize[0], vsy = p.vsize[1], vsz = p.vsize[2], vox = p.voff[0], voy = p.voff[1], voz = p.voff[2];
    const float *__restrict__ grows = p.sorted_rows;
    const uint64_t keep_policy = l2_policy_evict_last();      // pillar rows: k_canvas re-reads them from L2
    const int m_raw = s_R[p.B];
    const int n_chunks = (m_raw + 31) >> 5;
    const int half = lane >> 4;
    const unsigned lt = (1u << lane) - 1u;

    // one point through decorate + Linear + BN, folded into the running max (integer max on the float bits: exact for
    // the non-negative post-ReLU values, drops negatives and -0 = the ReLU, lets a NaN 0x7fffffff win as torch.max does)
    auto eval_point = [&](const float *rowp, float mx, float my, float mz, float cx, float cy, float cz,
                          int &v0, int &v1, int &v2, int &v3) {
        const float4 *r4 = reinterpret_cast<const float4 *>(rowp);
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
            for (int q = ABS ? 0 : 3; q < F; ++q) feat[kf++] = row[q];
            feat[kf++] = __fsub_rn(row[0], mx); feat[kf++] = __fsub_rn(row[1], my); feat[kf++] = __fsub_rn(row[2], mz);
            feat[kf++] = __fsub_rn(row[0], cx); feat[kf++] = __fsub_rn(row[1], cy); feat[kf++] = __fsub_rn(row[2], cz);
            // torch.norm(xyz, 2, 2) on the CPU: sqrt(fma(z,z, fma(y,y, x*x)))  (pillar_vfe.py:110-112)
            if (DIST) feat[kf++] = __fsqrt_rn(fmaf(row[2], row[2], fmaf(row[1], row[1], __fmul_rn(row[0], row[0]))));
        }
        uint64_t a01 = 0ull, a23 = 0ull;       // (+0, +0)
#pragma unroll
        for (int kk = 0; kk < CIN; ++kk) {     // Linear: sequential FMA in k order (pillar_vfe.py:37), two channels per FFMA2
            const uint64_t ff = pack_f2(feat[kk], feat[kk]);
            a01 = fma2_rn(ff, w01[kk], a01);
            a23 = fma2_rn(ff, w23[kk], a23);
        }
        float y0, y1, y2, y3;
        if (BN) {                              // BN eval: (((x-mean)*invstd)*gamma)+beta, 4 roundings (:39); the last add scalar (common.cuh)
            unpack_f2(mul2_rn(mul2_rn(sub2_rn(a01, mu01), iv01), ga01), y0, y1);
            unpack_f2(mul2_rn(mul2_rn(sub2_rn(a23, mu23), iv23), ga23), y2, y3);
            y0 = __fadd_rn(y0, be.x); y1 = __fadd_rn(y1, be.y); y2 = __fadd_rn(y2, be.z); y3 = __fadd_rn(y3, be.w);
        } else {
            unpack_f2(a01, y0, y1); unpack_f2(a23, y2, y3);
            y0 = __fadd_rn(y0, be.x); y1 = __fadd_rn(y1, be.y); y2 = __fadd_rn(y2, be.z); y3 = __fadd_rn(y3, be.w);
        }
        v0 = max(v0, __float_as_int(y0)); v1 = max(v1, __float_as_int(y1));
        v2 = max(v2, __float_as_int(y2)); v3 = max(v3, __float_as_int(y3));
    };

    for (;;) {
        // chunks are handed out dynamically: dense chunks (many points per pillar) cost several times a sparse one
        int ch = 0;
        if (lane == 0) ch = (int)atomicAdd(p.ticket, 1u);
        ch = __shfl_sync(FULL, ch, 0);
        if (ch >= n_chunks) break;
        // ---- owner phase: lane j looks after pillar m0 + j ----
        const int m = ch * 32 + lane;
        int4 pr = make_int4(0, 0, 0, 0);
        if (m < m_raw) pr = __ldg(p.prec + m);
        const int cnt = pr.y, pb = pr.z >> 16, pz = pr.z & 0xFFFF, py = pr.w >> 16, px = pr.w & 0xFFFF;
        const int start = (m < m_raw) ? (int)__ldg(p.cell_start + pr.x) : 0;   // CSR start of the pillar's cell
        const int local = m - s_R[pb];
        const bool kept = (m < m_raw) && (local < maxv);      // pillars beyond max_voxels were never created
        const int f = s_K[pb] + local;                          // final pillar id (first-seen order, frames concatenated)
        const int n_keep = min(cnt, Pmax);
        const float *grow = grows + (size_t)start * RW;
        if (kept) {
            p.num[f] = n_keep;
            *reinterpret_cast<int4 *>(p.coords + 4 * (size_t)f) = make_int4(pb, pz, py, px);
        }
        // order the pillar's points by input index
        if (kept && cnt > 1 && cnt <= SMALL_CNT) {
            uint32_t idx[SMALL_CNT];
#pragma unroll
            for (int j = 0; j < SMALL_CNT; ++j) idx[j] = (j < cnt) ? __float_as_uint(__ldg(grow + (size_t)j * RW + Fr)) : 0xFFFFFFFFu;
#pragma unroll
            for (int j = 0; j < SMALL_CNT; ++j) {
                int rank = 0;
#pragma unroll
                for (int q = 0; q < SMALL_CNT; ++q) rank += (idx[q] < idx[j]) ? 1 : 0;
                if (j < cnt) perm[lane][rank] = (unsigned char)j;
            }
        }
        unsigned coop = __ballot_sync(FULL, kept && cnt > SMALL_CNT && cnt <= 32);   // the warp ranks these one at a time
        while (coop) {
            const int o = __ffs(coop) - 1;
            coop &= coop - 1;
            const int cnt_o = __shfl_sync(FULL, cnt, o), start_o = __shfl_sync(FULL, start, o);
            const uint32_t mine = (lane < cnt_o) ? __float_as_uint(__ldg(grows + (size_t)(start_o + lane) * RW + Fr)) : 0xFFFFFFFFu;
            int rank = 0;
            for (int q = 0; q < cnt_o; ++q) rank += (__shfl_sync(FULL, mine, q) < mine) ? 1 : 0;
            if (lane < cnt_o) perm[o][rank] = (unsigned char)lane;
        }
        __syncwarp();
        const unsigned huge = __ballot_sync(FULL, kept && cnt > 32);
        if (PFN) {
            // mean of the kept points (torch CPU sum order) and the record the unit lanes read
            const bool live = kept && cnt <= 32;
            float mx = 0.f, my = 0.f, mz = 0.f;
            if (live) {
                if (cnt == 1) {
                    const float4 v = __ldg(reinterpret_cast<const float4 *>(grow));
                    mx = v.x; my = v.y; mz = v.z;               // mean of one point is the point (x/1 is exact)
                } else {
                    SlotSum sum;
                    for (int s2 = 0; s2 < n_keep; ++s2) {
                        const float4 v = __ldg(reinterpret_cast<const float4 *>(grow + (size_t)perm[lane][s2] * RW));
                        sum.add(s2, P4, v.x, v.y, v.z);
                    }
                    const float fn = (float)n_keep;
                    mx = __fdiv_rn(sum.sx(), fn); my = __fdiv_rn(sum.sy(), fn); mz = __fdiv_rn(sum.sz(), fn);
                }
            }
            // two work lists: pillars with ONE point to evaluate (paired across the half-warps) and pillars with several
            // (taken one at a time, the half-warps splitting the points)
            const unsigned sbal = __ballot_sync(FULL, live && n_keep == 1);
            const unsigned mbal = __ballot_sync(FULL, live && n_keep > 1);
            const int n_s = __popc(sbal), n_m = __popc(mbal);
            if (live) {
                const int slot = (n_keep == 1) ? __popc(sbal & lt) : 31 - __popc(mbal & lt);   // singles from the front, multis from the back
                // position of the single evaluated point: 0, or the rank-0 arrival when P == 1 truncated a larger pillar
                const int pos0 = (cnt == 1) ? 0 : (int)perm[lane][0];
                rec[slot][0] = make_float4(mx, my, mz, __int_as_float(n_keep | (lane << 8) | (pos0 << 16)));
                rec[slot][1] = make_float4(__int_as_float(start), __int_as_float(f), __int_as_float(pr.z), __int_as_float(pr.w));
            }
            __syncwarp();
            // ---- unit phase.  lane l always computes channels c0..c0+3 ----
            // (a) single-point pillars: half-warp h takes list entries 2*it + h
#pragma unroll 1
            for (int it = 0; 2 * it < n_s; ++it) {
                const int e = 2 * it + half;
                if (e < n_s) {
                    const float4 r0 = rec[e][0], r1 = rec[e][1];
                    const int meta = __float_as_int(r0.w);
                    const int zz = __float_as_int(r1.z) & 0xFFFF, yx = __float_as_int(r1.w);
                    // pillar centre: fl(fl(c*v)+off), two roundings, no FMA (pillar_vfe.py:101-103)
                    const float cx = __fadd_rn(__fmul_rn((float)(yx & 0xFFFF), vsx), vox);
                    const float cy = __fadd_rn(__fmul_rn((float)(yx >> 16), vsy), voy);
                    const float cz = __fadd_rn(__fmul_rn((float)zz, vsz), voz);
                    int v0 = 0, v1 = 0, v2 = 0, v3 = 0;
                    if (1 < Pmax) { v0 = __float_as_int(pv.x); v1 = __float_as_int(pv.y); v2 = __float_as_int(pv.z); v3 = __float_as_int(pv.w); }
                    eval_point(grows + ((size_t)__float_as_int(r1.x) + (meta >> 16)) * RWc, r0.x, r0.y, r0.z, cx, cy, cz, v0, v1, v2, v3);
                    if (p.feats) st_f4_hint(p.feats + (size_t)__float_as_int(r1.y) * C + c0,
                               make_float4(__int_as_float(v0), __int_as_float(v1), __int_as_float(v2), __int_as_float(v3)), keep_policy);
                }
            }
            // (b) multi-point pillars: both half-warps on the same pillar, half h takes slots h, h+2, ...; max-combined
#pragma unroll 1
            for (int it = 0; it < n_m; ++it) {
                const float4 r0 = rec[31 - it][0], r1 = rec[31 - it][1];
                const int meta = __float_as_int(r0.w);
                const int nk = meta & 0xFF, owner = (meta >> 8) & 0xFF;
                const int zz = __float_as_int(r1.z) & 0xFFFF, yx = __float_as_int(r1.w);
                const float cx = __fadd_rn(__fmul_rn((float)(yx & 0xFFFF), vsx), vox);
                const float cy = __fadd_rn(__fmul_rn((float)(yx >> 16), vsy), voy);
                const float cz = __fadd_rn(__fmul_rn((float)zz, vsz), voz);
                const float *rowb = grows + (size_t)__float_as_int(r1.x) * RWc;
                int v0 = 0, v1 = 0, v2 = 0, v3 = 0;
                if (nk < Pmax) { v0 = __float_as_int(pv.x); v1 = __float_as_int(pv.y); v2 = __float_as_int(pv.z); v3 = __float_as_int(pv.w); }
#pragma unroll 1
                for (int s2 = half; s2 < nk; s2 += 2)
                    eval_point(rowb + (size_t)perm[owner][s2] * RWc, r0.x, r0.y, r0.z, cx, cy, cz, v0, v1, v2, v3);
                v0 = max(v0, __shfl_xor_sync(FULL, v0, 16)); v1 = max(v1, __shfl_xor_sync(FULL, v1, 16));
                v2 = max(v2, __shfl_xor_sync(FULL, v2, 16)); v3 = max(v3, __shfl_xor_sync(FULL, v3, 16));
                if (half == 0 && p.feats)
                    st_f4_hint(p.feats + (size_t)__float_as_int(r1.y) * C + c0,
                               make_float4(__int_as_float(v0), __int_as_float(v1), __int_as_float(v2), __int_as_float(v3)), keep_policy);
            }
        }
        // ---- pillars with more than 32 arrivals: the warp selects the 32 smallest point indices, then as (b) ----
        unsigned hm = huge;
        while (hm) {
            const int o = __ffs(hm) - 1;
            hm &= hm - 1;
            const int cnt_o = __shfl_sync(FULL, cnt, o), start_o = __shfl_sync(FULL, start, o), f_o = __shfl_sync(FULL, f, o);
            const int prz = __shfl_sync(FULL, pr.z, o), prw = __shfl_sync(FULL, pr.w, o);
            const int nk = min(cnt_o, Pmax);
            const float *grow_o = grows + (size_t)start_o * RW;
            bperm[lane] = select_first32(grow_o + Fr, RW, cnt_o, lane);
            __syncwarp();
            if (p.voxels) {
                float *vo = p.voxels + (size_t)f_o * Pmax * Fr;
                for (int t = lane; t < Pmax * Fr; t += 32) {
                    const int s2 = t / Fr, kk = t - s2 * Fr;
                    vo[t] = (s2 < nk) ? __ldg(grow_o + (size_t)bperm[s2] * RW + kk) : 0.f;
                }
            }
            if (PFN) {
                SlotSum sum;
                for (int s2 = 0; s2 < nk; ++s2) {
                    const float4 v = __ldg(reinterpret_cast<const float4 *>(grow_o + (size_t)bperm[s2] * RWc));
                    sum.add(s2, P4, v.x, v.y, v.z);
                }
                const float fn = (float)nk;
                const float mx = __fdiv_rn(sum.sx(), fn), my = __fdiv_rn(sum.sy(), fn), mz = __fdiv_rn(sum.sz(), fn);
                const float cx = __fadd_rn(__fmul_rn((float)(prw & 0xFFFF), vsx), vox);
                const float cy = __fadd_rn(__fmul_rn((float)(prw >> 16), vsy), voy);
                const float cz = __fadd_rn(__fmul_rn((float)(prz & 0xFFFF), vsz), voz);
                int v0 = 0, v1 = 0, v2 = 0, v3 = 0;
                if (nk < Pmax) { v0 = __float_as_int(pv.x); v1 = __float_as_int(pv.y); v2 = __float_as_int(pv.z); v3 = __float_as_int(pv.w); }
#pragma unroll 1
                for (int s2 = half; s2 < nk; s2 += 2)
                    eval_point(grow_o + (size_t)bperm[s2] * RWc, mx, my, mz, cx, cy, cz, v0, v1, v2, v3);
                v0 = max(v0, __shfl_xor_sync(FULL, v0, 16)); v1 = max(v1, __shfl_xor_sync(FULL, v1, 16));
                v2 = max(v2, __shfl_xor_sync(FULL, v2, 16)); v3 = max(v3, __shfl_xor_sync(FULL, v3, 16));
                if (half == 0 && p.feats)
                    st_f4_hint(p.feats + (size_t)f_o * C + c0,
                               make_float4(__int_as_float(v0), __int_as_float(v1), __int_as_float(v2), __int_as_float(v3)), keep_policy);
            }
            __syncwarp();
        }
        // ---- optional contract output: the padded voxels tensor [M, P, F], coalesced, one pillar at a time ----
        if (p.voxels) {
            unsigned todo = __ballot_sync(FULL, kept && cnt <= 32);
            while (todo) {
                const int o = __ffs(todo) - 1;
                todo &= todo - 1;
                const int cnt_o = __shfl_sync(FULL, cnt, o), start_o = __shfl_sync(FULL, start, o), f_o = __shfl_sync(FULL, f, o);
                const int nk = min(cnt_o, Pmax);
                float *vo = p.voxels + (size_t)f_o * Pmax * Fr;
                const float *grow_o = grows + (size_t)start_o * RW;
                for (int t = lane; t < Pmax * Fr; t += 32) {
                    const int s2 = t / Fr, kk = t - s2 * Fr;
                    vo[t] = (s2 < nk) ? __ldg(grow_o + (size_t)((cnt_o == 1) ? 0 : (int)perm[o][s2]) * RW + kk) : 0.f;
                }
            }
        }
        __syncwarp();   // rec / perm are rewritten by the next chunk
    }
}

// ---- k_emit -------------------------------------------------------------------------------------
// The fused kernel (used whenever the canvas is requested): order + decorate + PFN + max AND the canvas tile, one pass.
// Every WARP is an autonomous worker walking its own sequence of canvas tiles (a tile = 32 cells of one BEV row x 64
// channels = 64 rows of 128 B); no CTA barrier in the loop.  Because k_front laid the CSR out in cell order, the point
// rows of a tile are ONE contiguous span of sorted_rows: they are staged with a single cooperative cp.async copy issued
// a tile ahead, from table entries loaded two tiles ahead -- no dependent gathers anywhere.
//   lane l OWNS cell l of the tile for the bookkeeping (ordering by point index, first P kept, mean in torch's
//   summation order, voxel_coords / voxel_num_points);
//   the arithmetic is cut into UNITS of (pillar, 4 output channels): lane l always computes channels 4*(l&15)..+3, so
//   its Linear weight float4s and BatchNorm constants stay in REGISTERS for the whole kernel (CUDA-core FMA: a 13x64
//   contraction is far below a tensor-core tile).  Single-point pillars are paired across the half-warps; a
//   multi-point pillar is taken by both halves, which split its points and max-combine.
//   A finished tile leaves in ONE TMA tensor store (128-byte swizzle so the column writes spread over banks); an
//   empty tile is four stores of a shared 2 KB zero tile.  The canvas is written exactly once, zeros included.
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

#ifndef HGSF_EMIT_MINB
#define HGSF_EMIT_MINB 3
#endif
template <int F, bool ABS, bool DIST, bool BN, int STORE, int CHUNK>
__global__ void __launch_bounds__(EMIT_THREADS, HGSF_EMIT_MINB)
k_emit(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap zmap, const PathParams p) {
    constexpr int C = 64;
    constexpr int CIN = (ABS ? F : F - 3) + 6 + (DIST ? 1 : 0);
    constexpr int RWc = (F + 1 + 3) / 4 * 4;   // F features + the point index, padded to float4
    constexpr int NV = RWc / 4;
    constexpr int TILE = C * 32;
    constexpr int ZC = C / 4;
    constexpr bool TMA = (STORE == 0);
    constexpr int NT = EMIT_THREADS;

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // the TMA swizzle works on absolute shared-memory address bits: the tiles must start on a 1024-byte boundary
    // (static shared memory in front of the dynamic part can shift it; the launcher reserves the slack)
    uint8_t *smem_al = smem_raw + ((1024u - ((uint32_t)__cvta_generic_to_shared(smem_raw) & 1023u)) & 1023u);
    float *tiles = reinterpret_cast<float *>(smem_al);                     // [EMIT_WARPS][TILE]
    float *zerobuf = tiles + EMIT_WARPS * TILE;                            // [ZC*32]
    float *stage_all = zerobuf + ZC * 32;                                  // [EMIT_WARPS][2][STAGE_W * RWc]
    int *s_R = reinterpret_cast<int *>(stage_all + EMIT_WARPS * 2 * STAGE_W * RWc);   // [B+1] raw pillar base per frame
    int *s_K = s_R + (p.B + 1);                                            // [B+1] kept (final) pillar base per frame
    __shared__ float4 s_rec_all[EMIT_WARPS][32][2];                        // work lists: singles from the front, multis from the back
    __shared__ unsigned char s_perm_all[EMIT_WARPS][32][32];               // per cell: arrival position of its rank-th point
    __shared__ int s_bperm_all[EMIT_WARPS][32];                            // same for a pillar with > 32 arrivals

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float *tile = tiles + warp * TILE;
    float *stage = stage_all + (size_t)warp * 2 * STAGE_W * RWc;
    float4(*rec)[2] = s_rec_all[warp];
    unsigned char(*perm)[32] = s_perm_all[warp];
    int *bperm = s_bperm_all[warp];

    // ---- one-time setup (the only CTA barriers) ----
    for (int b = tid; b <= p.B; b += NT) s_R[b] = p.frame_raw_base[b];
    for (int t = tid; t < ZC * 32; t += NT) zerobuf[t] = 0.f;
    for (int t = tid; t < EMIT_WARPS * TILE; t += NT) tiles[t] = 0.f;
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
    if (TMA) fence_proxy_async_smem();
    __syncthreads();

    // this lane's 4 channels: Linear rows and BatchNorm constants, in registers for the whole kernel
    const int c0 = 4 * (lane & 15);
    const int half = lane >> 4;
    uint64_t w01[CIN], w23[CIN];           // channel pairs (c0, c0+1), (c0+2, c0+3): one FFMA2 each per input feature
    float4 mu = make_float4(0.f, 0.f, 0.f, 0.f), iv = mu, ga = mu, be = mu, pv = mu;
    {
#pragma unroll
        for (int k = 0; k < CIN; ++k) {
            w01[k] = pack_f2(__ldg(p.W + (c0 + 0) * CIN + k), __ldg(p.W + (c0 + 1) * CIN + k));
            w23[k] = pack_f2(__ldg(p.W + (c0 + 2) * CIN + k), __ldg(p.W + (c0 + 3) * CIN + k));
        }
        float bnv[5][4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = c0 + j;
            float y;
            if (BN) {
                bnv[0][j] = __ldg(p.bn_m + c);
                bnv[1][j] = __frcp_rn(__fsqrt_rn(__fadd_rn(__ldg(p.bn_v + c), p.eps)));
                bnv[2][j] = __ldg(p.bn_w + c);
                bnv[3][j] = __ldg(p.bn_b + c);
                // a zero (padded) row still goes through BN + ReLU and joins the max (pillar_vfe.py:37-42)
                y = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(0.f, bnv[0][j]), bnv[1][j]), bnv[2][j]), bnv[3][j]);
            } else {
                bnv[0][j] = bnv[1][j] = bnv[2][j] = 0.f;
                bnv[3][j] = __ldg(p.bias + c);
                y = __fadd_rn(0.f, bnv[3][j]);
            }
            bnv[4][j] = (y > 0.f || y != y) ? y : 0.f;
        }
        mu = make_float4(bnv[0][0], bnv[0][1], bnv[0][2], bnv[0][3]); iv = make_float4(bnv[1][0], bnv[1][1], bnv[1][2], bnv[1][3]);
        ga = make_float4(bnv[2][0], bnv[2][1], bnv[2][2], bnv[2][3]); be = make_float4(bnv[3][0], bnv[3][1], bnv[3][2], bnv[3][3]);
        pv = make_float4(bnv[4][0], bnv[4][1], bnv[4][2], bnv[4][3]);
    }
    const uint64_t mu01 = pack_f2(mu.x, mu.y), mu23 = pack_f2(mu.z, mu.w), iv01 = pack_f2(iv.x, iv.y), iv23 = pack_f2(iv.z, iv.w),
                   ga01 = pack_f2(ga.x, ga.y), ga23 = pack_f2(ga.z, ga.w);

    const int tiles_per_row = (p.nx + 31) >> 5;
    const int rows_per_frame = p.ny;                       // nz == 1 (PointPillarScatter asserts it)
    const int n_rows = p.B * rows_per_frame;
    const int P4 = (p.P >> 2) << 2;
    const int maxv = p.max_voxels, Pmax = p.P;
    const float vsx = p.vsize[0], vsy = p.vsize[1], vox = p.voff[0], voy = p.voff[1];
    const float cz = __fadd_rn(__fmul_rn(0.f, p.vsize[2]), p.voff[2]);     // z index 0: fl(fl(0*vz)+z_off)
    const float *__restrict__ grows = p.sorted_rows;
    const unsigned lt = (1u << lane) - 1u;
    const uint64_t stream_policy = l2_policy_evict_first();   // canvas: written once, never re-read here
#ifdef HGSF_EXPERIMENT
    const uint64_t feats_policy = (p.dbg & 8) ? l2_policy_evict_last() : ((p.dbg & 16) ? l2_policy_evict_normal() : stream_policy);
#else
    const uint64_t feats_policy = stream_policy;
#endif

    // one point through decorate + Linear + BN, folded into the running max (integer max on the float bits: exact for
    // the non-negative post-ReLU values, drops negatives and -0 = the ReLU, lets a NaN 0x7fffffff win as torch.max does)
    auto eval_row = [&](const float (&row)[RWc], float mx, float my, float mz, float cx, float cy,
                        int &v0, int &v1, int &v2, int &v3) {
        float feat[CIN];
        {
            int kf = 0;
#pragma unroll
            for (int q = ABS ? 0 : 3; q < F; ++q) feat[kf++] = row[q];
            feat[kf++] = __fsub_rn(row[0], mx); feat[kf++] = __fsub_rn(row[1], my); feat[kf++] = __fsub_rn(row[2], mz);
            feat[kf++] = __fsub_rn(row[0], cx); feat[kf++] = __fsub_rn(row[1], cy); feat[kf++] = __fsub_rn(row[2], cz);
            // torch.norm(xyz, 2, 2) on the CPU: sqrt(fma(z,z, fma(y,y, x*x)))  (pillar_vfe.py:110-112)
            if (DIST) feat[kf++] = __fsqrt_rn(fmaf(row[2], row[2], fmaf(row[1], row[1], __fmul_rn(row[0], row[0]))));
        }
        uint64_t a01 = 0ull, a23 = 0ull;       // (+0, +0)
#pragma unroll
        for (int kk = 0; kk < CIN; ++kk) {     // Linear: sequential FMA in k order (pillar_vfe.py:37), two channels per FFMA2
            const uint64_t ff = pack_f2(feat[kk], feat[kk]);
            a01 = fma2_rn(ff, w01[kk], a01);
            a23 = fma2_rn(ff, w23[kk], a23);
        }
        float y0, y1, y2, y3;
        if (BN) {                              // BN eval: (((x-mean)*invstd)*gamma)+beta, 4 roundings (:39); the last add scalar (common.cuh)
            unpack_f2(mul2_rn(mul2_rn(sub2_rn(a01, mu01), iv01), ga01), y0, y1);
            unpack_f2(mul2_rn(mul2_rn(sub2_rn(a23, mu23), iv23), ga23), y2, y3);
            y0 = __fadd_rn(y0, be.x); y1 = __fadd_rn(y1, be.y); y2 = __fadd_rn(y2, be.z); y3 = __fadd_rn(y3, be.w);
        } else {
            unpack_f2(a01, y0, y1); unpack_f2(a23, y2, y3);
            y0 = __fadd_rn(y0, be.x); y1 = __fadd_rn(y1, be.y); y2 = __fadd_rn(y2, be.z); y3 = __fadd_rn(y3, be.w);
        }
        v0 = max(v0, __float_as_int(y0)); v1 = max(v1, __float_as_int(y1));
        v2 = max(v2, __float_as_int(y2)); v3 = max(v3, __float_as_int(y3));
    };
    // a point row: from the staging buffer (rel >= 0: row index in it) or from global memory (rel < 0: -1 - CSR row),
    // read through ONE generic pointer so that the two sources do not become two divergent code paths
    auto load_row = [&](const float *stg, int rel, int pos, float (&row)[RWc]) {
        const float *base = (rel >= 0) ? stg + (size_t)rel * RWc : grows + (size_t)(-1 - rel) * RWc;
        const float4 *r4 = reinterpret_cast<const float4 *>(base + (size_t)pos * RWc);
#pragma unroll
        for (int v = 0; v < NV; ++v) { const float4 t4 = r4[v]; row[4 * v] = t4.x; row[4 * v + 1] = t4.y; row[4 * v + 2] = t4.z; row[4 * v + 3] = t4.w; }
    };
    // channel c0+i of cell `cell` sits at tile[(c0+i)*32 + (((cell>>2) ^ ((c0+i)&7)) << 2 | (cell&3))] (128-byte swizzle);
    // with c0 = 4*(lane&15): (c0+i)&7 = ((lane&1)<<2) ^ i, so the lane-constant part is folded once
    float *const tbase = tile + c0 * 32;
    const int swb = (lane & 1) << 2;
    auto put_tile = [&](int cell, int v0, int v1, int v2, int v3) {
        const int xs = (cell >> 2) ^ swb, xr = cell & 3;
        tbase[0 * 32 + (((xs ^ 0) << 2) | xr)] = __int_as_float(v0);
        tbase[1 * 32 + (((xs ^ 1) << 2) | xr)] = __int_as_float(v1);
        tbase[2 * 32 + (((xs ^ 2) << 2) | xr)] = __int_as_float(v2);
        tbase[3 * 32 + (((xs ^ 3) << 2) | xr)] = __int_as_float(v3);
    };

    // Tiles are handed out DYNAMICALLY, one ticket per tile, fetched a tile ahead: a dense tile costs ten times a sparse
    // one, and a static assignment leaves the unlucky warps running alone at the end.  One tile per ticket (rather than a
    // run of consecutive tiles per warp) also keeps x-adjacent tiles -- adjacent 128-byte pieces of the same canvas rows --
    // in flight at the same time on different warps, which the DRAM write stream rewards (measured: runs of 10 / 4 / 2 / 1
    // tiles -> 0.195 / 0.178 / 0.169 / 0.167 ms per step).
    // A ticket covers CHUNK consecutive tiles (chosen by the launcher, a compile-time constant: as a run-time value it cost
    // the dense workloads 3 %): 1 for the usual density, 2 for sparse
    // scenes.  With mostly empty tiles (two thirds of TJ4D's at 30 000 points per frame) an iteration is shorter than the
    // round trip of the same-address atomic under load and the warps wait for tickets (ncu: the atomic was the top stall,
    // k_emit 0.374 ms against 0.239 ms with two tiles per ticket).  Tried instead and measured worse: several ticket
    // counters over separate tile ranges (cure the stall but split the write stream: VoD uniform 0.169 -> 0.188 ms with 16
    // queues), more tickets in flight per warp (no effect on the stall, tiles leave further out of order), tile order
    // interleaved over the frames, a static round-robin assignment (loses the balancing).
    const int n_tiles = n_rows * tiles_per_row;
    constexpr int chunk = CHUNK;
    // the ticket stays in lane 0's register until the chunk is actually started: broadcasting it right away would
    // stall the whole warp on the atomic's round trip
    auto fetch_raw = [&]() -> int {
        int v = 0;
        if (lane == 0) v = (int)atomicAdd(p.ticket + 32, 1u);
        return v;
    };
    auto decode = [&](int t) -> TilePos {
        TilePos q;
        if (t >= n_tiles) { q.r = n_rows; q.xt = 0; q.b = p.B; q.zy = 0; return q; }
        q.r = t / tiles_per_row; q.xt = t - q.r * tiles_per_row;
        q.b = q.r / rows_per_frame; q.zy = q.r - q.b * rows_per_frame;
        return q;
    };
    int seq_left = chunk, next_raw = 0;           // tiles left in the current chunk; ticket of the prefetched next chunk (lane 0)
    TilePos seq = decode(__shfl_sync(FULL, fetch_raw(), 0) * chunk);   // the furthest tile handed to the pipeline so far
    next_raw = fetch_raw();
    auto next_tile = [&]() -> TilePos {
        if (seq_left > 1 && seq.r < n_rows) {
            --seq_left;
            if (++seq.xt == tiles_per_row) { seq.xt = 0; ++seq.r; if (++seq.zy == rows_per_frame) { seq.zy = 0; ++seq.b; } }
        } else {
            seq = decode(__shfl_sync(FULL, next_raw, 0) * chunk);
            seq_left = chunk;
            if (seq.r < n_rows) next_raw = fetch_raw();
        }
        return seq;
    };
    TilePos cur = seq;
    TilePos nxt = next_tile();
    TilePos nxt2 = next_tile();
    auto load_entry = [&](const TilePos &t) -> uint4 {
        const int x = t.xt * 32 + lane;
        // row r = b*ny + y and the table is [b][y][x]: the cell index is r*nx + x
        uint4 e = make_uint4(0, 0, 0, 0);
        if (t.r < n_rows && x < p.nx) {
            const size_t c = (size_t)t.r * p.nx + x;
            e.x = __ldg(p.cell_tag + c); e.y = __ldg(p.cell_cnt + c); e.z = __ldg(p.cell_start + c);
        }
        return e;
    };
    // the tile's rows are sorted_rows[row0, row0 + total): one cooperative async copy of (at most STAGE_W of) them
    auto issue_stage = [&](const uint4 e, float *stg) {
        const int row0 = __shfl_sync(FULL, (int)e.z, 0);
        const int total = (int)__reduce_add_sync(FULL, e.y);      // one REDUX instead of a shuffle tree
        const int chunks = min(total, STAGE_W) * NV;
        const float *src = grows + (size_t)row0 * RWc;
        for (int c = lane; c < chunks; c += 32) cp_async16(stg + 4 * c, src + 4 * c);
        cp_async_commit();
    };

    uint4 e_cur = load_entry(cur);
    issue_stage(e_cur, stage);
    uint4 e_nxt = load_entry(nxt);
    unsigned dirty = 0;                  // cells of the tile buffer that hold non-zero columns
    bool store_pending = false;          // a TMA store from the tile buffer may still be reading it

    for (int it = 0; cur.r < n_rows; ++it) {
        const float *stg = stage + (size_t)(it & 1) * STAGE_W * RWc;
        // ---- pipeline: entries of the tile after next, rows of the next tile ----
        const uint4 e_nn = load_entry(nxt2);
        issue_stage(e_nxt, stage + (size_t)((it + 1) & 1) * STAGE_W * RWc);

        const int b = cur.b, y = cur.zy, x0 = cur.xt * 32;
        const int m = (int)(e_cur.x - 1u), cnt = (int)e_cur.y, start = (int)e_cur.z;
        const int local = m - s_R[b];
        const bool occ = (e_cur.x != 0u) && (local < maxv);     // pillars beyond max_voxels were never created
        const unsigned bal_occ = __ballot_sync(FULL, occ);
        if (bal_occ == 0u) {
            // empty tile: four stores of the shared zero tile
            if (TMA) {
                // issued by lane 1: bulk async-groups are per thread, so lane 0's wait for its tile store to have read the
                // tile buffer (below) does not also wait for the zero stores of the empty tiles that came after it
                if (lane == 1) {
#pragma unroll
                    for (int q4 = 0; q4 < 4; ++q4) tma_store_3d_hint(&zmap, zerobuf, x0, y, b * C + q4 * ZC, stream_policy);
                    tma_commit();
                }
            } else if (STORE == 1) {
                const int xc = x0 + 4 * (lane & 7);      // lane l: 16-byte chunk (l & 7) of channel rows (l >> 3) + 4*i
                if (xc < p.nx) {
                    float *dst = p.canvas + (((size_t)b * C + (lane >> 3)) * p.ny + y) * p.nx + xc;
                    const size_t plane4 = (size_t)4 * p.ny * p.nx;
#pragma unroll
                    for (int i = 0; i < C / 4; ++i) __stcs(reinterpret_cast<float4 *>(dst + i * plane4), make_float4(0.f, 0.f, 0.f, 0.f));
                }
            } else if (x0 + lane < p.nx) {
                for (int ch = 0; ch < C; ++ch) p.canvas[(((size_t)b * C + ch) * p.ny + y) * p.nx + x0 + lane] = 0.f;
            }
        } else {
            const int row0 = __shfl_sync(FULL, start, 0);
            const int rel0 = start - row0;
            const bool staged = occ && cnt <= 32 && rel0 + cnt <= STAGE_W;
            const int rel = staged ? rel0 : (-1 - start);         // where the pillar's rows are (see load_row)
            const int n_keep = min(cnt, Pmax);
            const int f = s_K[b] + local;                          // final pillar id (first-seen order, frames concatenated)
            if (occ) {
                p.num[f] = n_keep;
                *reinterpret_cast<int4 *>(p.coords + 4 * (size_t)f) = make_int4(b, 0, y, x0 + lane);
            }
            cp_async_wait<1>();          // this tile's rows have landed (this lane's copies) ...
            __syncwarp();                // ... and every other lane's
            // ---- order the pillar's points by input index; mean of the kept points (torch CPU sum order) ----
            const bool live = occ && cnt <= 32;
            float mx = 0.f, my = 0.f, mz = 0.f;
            // the common tile has only 1- and 2-point pillars: a warp-uniform short cut for it (one compare instead of the
            // 6-way ranking, no summation loop, and x/2 as the exact x*0.5 instead of the IEEE division routine)
            const unsigned multi_bal = __ballot_sync(FULL, occ && cnt > 1);
            const unsigned pair_bal = __ballot_sync(FULL, occ && cnt == 2 && staged);
            const bool pairs_only = (multi_bal == pair_bal) && Pmax >= 2;
#ifdef HGSF_EXPERIMENT
            // ablation (WRONG results, timing only): as if the rows arrived ordered and the mean were precomputed
            if (p.dbg & 32) {
                if (live) {
                    if (!(p.dbg & 64)) for (int j = 0; j < n_keep; ++j) perm[lane][j] = (unsigned char)j;
                    const float4 a = *reinterpret_cast<const float4 *>((rel >= 0) ? stg + (size_t)rel * RWc : grows + (size_t)(-1 - rel) * RWc);
                    mx = a.x; my = a.y; mz = a.z;
                }
                __syncwarp();
            } else
#endif
            if (pairs_only) {
                if (occ && cnt == 2) {
                    const float *r0p = stg + (size_t)rel * RWc, *r1p = r0p + RWc;
                    const int first = (__float_as_uint(r1p[F]) < __float_as_uint(r0p[F])) ? 1 : 0;
                    perm[lane][0] = (unsigned char)first; perm[lane][1] = (unsigned char)(first ^ 1);
                    const float4 a = *reinterpret_cast<const float4 *>(first ? r1p : r0p);
                    const float4 c = *reinterpret_cast<const float4 *>(first ? r0p : r1p);
                    SlotSum sum;
                    sum.add(0, P4, a.x, a.y, a.z);
                    sum.add(1, P4, c.x, c.y, c.z);
                    mx = __fmul_rn(sum.sx(), 0.5f); my = __fmul_rn(sum.sy(), 0.5f); mz = __fmul_rn(sum.sz(), 0.5f);
                } else if (occ) {
                    float row[RWc];
                    load_row(stg, rel, 0, row);
                    mx = row[0]; my = row[1]; mz = row[2];       // mean of one point is the point (x/1 is exact)
                }
                __syncwarp();
            } else {
            if (occ && cnt > 1 && cnt <= SMALL_CNT) {
                uint32_t idx[SMALL_CNT];
#pragma unroll
                for (int j = 0; j < SMALL_CNT; ++j) {
                    idx[j] = 0xFFFFFFFFu;
                    if (j < cnt) idx[j] = staged ? __float_as_uint(stg[(size_t)(rel + j) * RWc + F])
                                                 : __float_as_uint(__ldg(grows + (size_t)(start + j) * RWc + F));
                }
#pragma unroll
                for (int j = 0; j < SMALL_CNT; ++j) {
                    int rank = 0;
#pragma unroll
                    for (int q = 0; q < SMALL_CNT; ++q) rank += (idx[q] < idx[j]) ? 1 : 0;
                    if (j < cnt) perm[lane][rank] = (unsigned char)j;
                }
            }
            unsigned coop = __ballot_sync(FULL, occ && cnt > SMALL_CNT && cnt <= 32);   // the warp ranks these one at a time
            while (coop) {
                const int o = __ffs(coop) - 1;
                coop &= coop - 1;
                const int cnt_o = __shfl_sync(FULL, cnt, o), rel_o = __shfl_sync(FULL, rel, o);
                uint32_t mine = 0xFFFFFFFFu;
                if (lane < cnt_o) mine = (rel_o >= 0) ? __float_as_uint(stg[(size_t)(rel_o + lane) * RWc + F])
                                                      : __float_as_uint(__ldg(grows + (size_t)(-1 - rel_o + lane) * RWc + F));
                int rank = 0;
                for (int q = 0; q < cnt_o; ++q) rank += (__shfl_sync(FULL, mine, q) < mine) ? 1 : 0;
                if (lane < cnt_o) perm[o][rank] = (unsigned char)lane;
            }
            __syncwarp();
            if (live) {
                float row[RWc];
                if (cnt == 1) {
                    load_row(stg, rel, 0, row);
                    mx = row[0]; my = row[1]; mz = row[2];       // mean of one point is the point (x/1 is exact)
                } else {
                    SlotSum sum;
                    for (int s2 = 0; s2 < n_keep; ++s2) {
                        load_row(stg, rel, perm[lane][s2], row);
                        sum.add(s2, P4, row[0], row[1], row[2]);
                    }
                    const float fn = (float)n_keep;
                    mx = __fdiv_rn(sum.sx(), fn); my = __fdiv_rn(sum.sy(), fn); mz = __fdiv_rn(sum.sz(), fn);
                }
            }
            }
            const unsigned sbal = __ballot_sync(FULL, live && n_keep == 1);
            const unsigned mbal = __ballot_sync(FULL, live && n_keep > 1);
            const int n_s = __popc(sbal), n_m = __popc(mbal);
            if (live) {
                const int slot = (n_keep == 1) ? __popc(sbal & lt) : 31 - __popc(mbal & lt);   // singles from the front, multis from the back
#ifdef HGSF_EXPERIMENT
                const int pos0 = (cnt == 1 || (p.dbg & 64)) ? 0 : (int)perm[lane][0];
#else
                const int pos0 = (cnt == 1) ? 0 : (int)perm[lane][0];   // the one evaluated point (rank 0 when P == 1 truncated)
#endif
                rec[slot][0] = make_float4(mx, my, mz, __int_as_float(n_keep | (lane << 8) | (pos0 << 16)));
                rec[slot][1] = make_float4(__int_as_float(rel), __int_as_float(f), 0.f, 0.f);
            }
            // the tile buffer: wait until the previous store has read it, then clear what that tile dirtied
            if (TMA && store_pending) {
                if (lane == 0) tma_wait_read<0>();
                store_pending = false;
            }
            __syncwarp();
            if (__popc(dirty) > 2) {
#pragma unroll
                for (int t = 0; t < TILE / 128; ++t) *reinterpret_cast<float4 *>(tile + t * 128 + lane * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
            } else {
                while (dirty) {
                    const int dc = __ffs(dirty) - 1;
                    dirty &= dirty - 1;
                    tile[swz128(lane, dc)] = 0.f; tile[swz128(lane + 32, dc)] = 0.f;
                }
            }
            dirty = bal_occ;
            __syncwarp();
            // ---- unit phase.  lane l always computes channels c0..c0+3 ----
            const float cy = __fadd_rn(__fmul_rn((float)y, vsy), voy);   // pillar centre: fl(fl(c*v)+off), two roundings,
                                                                         // no FMA (pillar_vfe.py:101-103)
            // (a) single-point pillars: half-warp h takes list entries 4*j + h and 4*j + 2 + h -- two independent points per
            //     iteration, so that their FMA chains interleave (the kernel is latency-bound at 12 warps per SM)
#ifdef HGSF_EXPERIMENT
            const int n_s_run = (p.dbg & 1) ? 0 : n_s, n_m_run = (p.dbg & 1) ? 0 : n_m;
#else
            const int n_s_run = n_s, n_m_run = n_m;
#endif
#pragma unroll 1
            for (int j = 0; 4 * j < n_s_run; ++j) {
                const int eA = 4 * j + half;
                if (eA < n_s) {
                    const bool okB = eA + 2 < n_s;
                    const int eB = okB ? eA + 2 : eA;
                    const float4 rA0 = rec[eA][0], rA1 = rec[eA][1], rB0 = rec[eB][0], rB1 = rec[eB][1];
                    const int metaA = __float_as_int(rA0.w), metaB = __float_as_int(rB0.w);
                    const int cellA = (metaA >> 8) & 0xFF, cellB = (metaB >> 8) & 0xFF;
                    const float cxA = __fadd_rn(__fmul_rn((float)(x0 + cellA), vsx), vox);
                    const float cxB = __fadd_rn(__fmul_rn((float)(x0 + cellB), vsx), vox);
                    int a0 = 0, a1 = 0, a2 = 0, a3 = 0;
                    if (1 < Pmax) { a0 = __float_as_int(pv.x); a1 = __float_as_int(pv.y); a2 = __float_as_int(pv.z); a3 = __float_as_int(pv.w); }
                    int b0 = a0, b1 = a1, b2 = a2, b3 = a3;
                    float rowA[RWc], rowB[RWc];
                    load_row(stg, __float_as_int(rA1.x), metaA >> 16, rowA);
                    load_row(stg, __float_as_int(rB1.x), metaB >> 16, rowB);
                    eval_row(rowA, rA0.x, rA0.y, rA0.z, cxA, cy, a0, a1, a2, a3);
                    eval_row(rowB, rB0.x, rB0.y, rB0.z, cxB, cy, b0, b1, b2, b3);
                    if (p.feats) {
                        st_f4_hint(p.feats + (size_t)__float_as_int(rA1.y) * C + c0,
                                   make_float4(__int_as_float(a0), __int_as_float(a1), __int_as_float(a2), __int_as_float(a3)), feats_policy);
                        if (okB)
                            st_f4_hint(p.feats + (size_t)__float_as_int(rB1.y) * C + c0,
                                       make_float4(__int_as_float(b0), __int_as_float(b1), __int_as_float(b2), __int_as_float(b3)), feats_policy);
                    }
                    put_tile(cellA, a0, a1, a2, a3);
                    if (okB) put_tile(cellB, b0, b1, b2, b3);
                }
            }
            // (b) multi-point pillars: both half-warps on the same pillar, half h takes slots h, h+2, ... (two per iteration);
            //     max-combined
#pragma unroll 1
            for (int j = 0; j < n_m_run; ++j) {
                const float4 r0 = rec[31 - j][0], r1 = rec[31 - j][1];
                const int meta = __float_as_int(r0.w);
                const int nk = meta & 0xFF, cell = (meta >> 8) & 0xFF;
                const int relp = __float_as_int(r1.x);
                const float cx = __fadd_rn(__fmul_rn((float)(x0 + cell), vsx), vox);
                int v0 = 0, v1 = 0, v2 = 0, v3 = 0;
                if (nk < Pmax) { v0 = __float_as_int(pv.x); v1 = __float_as_int(pv.y); v2 = __float_as_int(pv.z); v3 = __float_as_int(pv.w); }
                int u0 = v0, u1 = v1, u2 = v2, u3 = v3;
#pragma unroll 1
                for (int s2 = half; s2 < nk; s2 += 4) {
                    const int s3 = (s2 + 2 < nk) ? s2 + 2 : s2;        // the last odd one is evaluated twice: max is idempotent
                    float rowA[RWc], rowB[RWc];
#ifdef HGSF_EXPERIMENT
                    load_row(stg, relp, (p.dbg & 64) ? s2 : (int)perm[cell][s2], rowA);
                    load_row(stg, relp, (p.dbg & 64) ? s3 : (int)perm[cell][s3], rowB);
#else
                    load_row(stg, relp, perm[cell][s2], rowA);
                    load_row(stg, relp, perm[cell][s3], rowB);
#endif
                    eval_row(rowA, r0.x, r0.y, r0.z, cx, cy, v0, v1, v2, v3);
                    eval_row(rowB, r0.x, r0.y, r0.z, cx, cy, u0, u1, u2, u3);
                }
                v0 = max(v0, u0); v1 = max(v1, u1); v2 = max(v2, u2); v3 = max(v3, u3);
                v0 = max(v0, __shfl_xor_sync(FULL, v0, 16)); v1 = max(v1, __shfl_xor_sync(FULL, v1, 16));
                v2 = max(v2, __shfl_xor_sync(FULL, v2, 16)); v3 = max(v3, __shfl_xor_sync(FULL, v3, 16));
                if (half == 0) {
                    if (p.feats)
                        st_f4_hint(p.feats + (size_t)__float_as_int(r1.y) * C + c0,
                                   make_float4(__int_as_float(v0), __int_as_float(v1), __int_as_float(v2), __int_as_float(v3)), feats_policy);
                    put_tile(cell, v0, v1, v2, v3);
                }
            }
            // ---- pillars with more than 32 arrivals: the warp selects the 32 smallest point indices, then as (b) ----
            unsigned hm = __ballot_sync(FULL, occ && cnt > 32);
            while (hm) {
                const int o = __ffs(hm) - 1;
                hm &= hm - 1;
                const int cnt_o = __shfl_sync(FULL, cnt, o), start_o = __shfl_sync(FULL, start, o), f_o = __shfl_sync(FULL, f, o);
                const int nk = min(cnt_o, Pmax);
                const float *grow_o = grows + (size_t)start_o * RWc;
                bperm[lane] = select_first32(grow_o + F, RWc, cnt_o, lane);
                __syncwarp();
                if (p.voxels) {
                    float *vo = p.voxels + (size_t)f_o * Pmax * F;
                    for (int t = lane; t < Pmax * F; t += 32) {
                        const int s2 = t / F, kk = t - s2 * F;
                        vo[t] = (s2 < nk) ? __ldg(grow_o + (size_t)bperm[s2] * RWc + kk) : 0.f;
                    }
                }
                SlotSum sum;
                for (int s2 = 0; s2 < nk; ++s2) {
                    const float4 v = __ldg(reinterpret_cast<const float4 *>(grow_o + (size_t)bperm[s2] * RWc));
                    sum.add(s2, P4, v.x, v.y, v.z);
                }
                const float fn = (float)nk;
                const float hx = __fdiv_rn(sum.sx(), fn), hy = __fdiv_rn(sum.sy(), fn), hz = __fdiv_rn(sum.sz(), fn);
                const float cx = __fadd_rn(__fmul_rn((float)(x0 + o), vsx), vox);
                int v0 = 0, v1 = 0, v2 = 0, v3 = 0;
                if (nk < Pmax) { v0 = __float_as_int(pv.x); v1 = __float_as_int(pv.y); v2 = __float_as_int(pv.z); v3 = __float_as_int(pv.w); }
#pragma unroll 1
                for (int s2 = half; s2 < nk; s2 += 2) {
                    float row[RWc];
                    load_row(stg, -1 - start_o, bperm[s2], row);
                    eval_row(row, hx, hy, hz, cx, cy, v0, v1, v2, v3);
                }
                v0 = max(v0, __shfl_xor_sync(FULL, v0, 16)); v1 = max(v1, __shfl_xor_sync(FULL, v1, 16));
                v2 = max(v2, __shfl_xor_sync(FULL, v2, 16)); v3 = max(v3, __shfl_xor_sync(FULL, v3, 16));
                if (half == 0) {
                    if (p.feats)
                        st_f4_hint(p.feats + (size_t)f_o * C + c0,
                                   make_float4(__int_as_float(v0), __int_as_float(v1), __int_as_float(v2), __int_as_float(v3)), feats_policy);
                    put_tile(o, v0, v1, v2, v3);
                }
                __syncwarp();
            }
            // ---- optional contract output: the padded voxels tensor [M, P, F], coalesced, one pillar at a time ----
            if (p.voxels) {
                unsigned todo = __ballot_sync(FULL, live);
                while (todo) {
                    const int o = __ffs(todo) - 1;
                    todo &= todo - 1;
                    const int cnt_o = __shfl_sync(FULL, cnt, o), rel_o = __shfl_sync(FULL, rel, o), f_o = __shfl_sync(FULL, f, o);
                    const int nk = min(cnt_o, Pmax);
                    float *vo = p.voxels + (size_t)f_o * Pmax * F;
                    for (int t = lane; t < Pmax * F; t += 32) {
                        const int s2 = t / F, kk = t - s2 * F;
                        float v = 0.f;
                        if (s2 < nk) {
                            const int pos = (cnt_o == 1) ? 0 : (int)perm[o][s2];
                            v = (rel_o >= 0) ? stg[(size_t)(rel_o + pos) * RWc + kk] : __ldg(grows + (size_t)(-1 - rel_o + pos) * RWc + kk);
                        }
                        vo[t] = v;
                    }
                }
            }
            // ---- the tile goes out in one piece ----
            if (TMA) {
                fence_proxy_async_smem();
                __syncwarp();
#ifdef HGSF_EXPERIMENT
                if (lane == 0 && !(p.dbg & 2)) { if (p.dbg & 4) tma_store_3d(&tmap, tile, x0, y, b * C); else tma_store_3d_hint(&tmap, tile, x0, y, b * C, stream_policy); tma_commit(); }
#else
                if (lane == 0) { tma_store_3d_hint(&tmap, tile, x0, y, b * C, stream_policy); tma_commit(); }
#endif
                store_pending = true;
            } else if (STORE == 1) {
                __syncwarp();
                const int xc = x0 + 4 * (lane & 7);
                if (xc < p.nx) {
                    float *dst = p.canvas + (((size_t)b * C + (lane >> 3)) * p.ny + y) * p.nx + xc;
                    const size_t plane4 = (size_t)4 * p.ny * p.nx;
#pragma unroll
                    for (int i = 0; i < C / 4; ++i) {
                        const int row = (lane >> 3) + 4 * i;
                        const float4 v = *reinterpret_cast<const float4 *>(tile + row * 32 + (((lane & 7) ^ (row & 7)) << 2));
                        __stcs(reinterpret_cast<float4 *>(dst + i * plane4), v);
                    }
                }
            } else {
                __syncwarp();
                if (x0 + lane < p.nx)
                    for (int ch = 0; ch < C; ++ch) p.canvas[(((size_t)b * C + ch) * p.ny + y) * p.nx + x0 + lane] = tile[swz128(ch, lane)];
            }
            __syncwarp();
        }
        e_cur = e_nxt; e_nxt = e_nn;
        cur = nxt; nxt = nxt2; nxt2 = next_tile();
    }
    cp_async_wait<0>();
    if (TMA && lane <= 1) tma_wait_read<0>();        // shared memory must outlive the stores that read it
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

template <typename K>
static int launch_persistent(K kern, int threads, size_t smem, long long work_ctas, cudaStream_t stream, int *grid_out) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    int per_sm = 1;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem);
    if (e != cudaSuccess) return (int)e;
    if (per_sm < 1) per_sm = 1;
    *grid_out = (int)std::max<long long>(1, std::min<long long>(work_ctas, (long long)sm_count() * per_sm));
    return HGSF_OK;
}

template <int F, bool ABS, bool DIST, bool PFN>
static int launch_pfn_t(const PathParams &p, cudaStream_t stream) {
    const size_t smem = sizeof(int) * 2 * (size_t)(p.B + 1);
    const long long chunks = ((long long)p.n + 31) / 32;          // upper bound on pillar chunks
    const bool bn = p.bn_w != nullptr;
    auto go = [&](auto kern) -> int {
        int grid = 1;
        const int st = launch_persistent(kern, PFN_THREADS, smem, (chunks + PFN_WARPS - 1) / PFN_WARPS, stream, &grid);
        if (st != HGSF_OK) return st;
        kern<<<(unsigned)grid, PFN_THREADS, smem, stream>>>(p);
        return (int)cudaGetLastError();
    };
    if constexpr (PFN) return bn ? go(k_pfn<F, ABS, DIST, true, true>) : go(k_pfn<F, ABS, DIST, false, true>);
    return go(k_pfn<F, ABS, DIST, true, false>);
}

static int launch_pfn(const PathParams &p, bool with_pfn, bool abs_xyz, bool dist, cudaStream_t s) {
    if (!with_pfn) return launch_pfn_t<4, true, false, false>(p, s);   // F / RW are read from the params when PFN is off
    if (p.C != 64) return HGSF_ERR_UNSUPPORTED;
#define HGSF_CASE(FV, A, D) if (p.F == FV && abs_xyz == A && dist == D) return launch_pfn_t<FV, A, D, true>(p, s);
    HGSF_CASE(4, true, false) HGSF_CASE(5, true, false) HGSF_CASE(6, true, false) HGSF_CASE(7, true, false)
    HGSF_CASE(8, true, false) HGSF_CASE(7, false, false) HGSF_CASE(8, false, false)
    HGSF_CASE(7, true, true) HGSF_CASE(8, true, true)
#undef HGSF_CASE
    return HGSF_ERR_UNSUPPORTED;
}

template <int F, bool ABS, bool DIST>
static int launch_emit_t(const PathParams &p, cudaStream_t stream) {
    constexpr int C = 64;
    constexpr int RWc = (F + 1 + 3) / 4 * 4;
    const bool vec_ok = (p.nx % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.canvas) & 15) == 0);
    const char *env = getenv("HGSF_CANVAS_STORE");      // tma | vec : which tile store to use (experiments); default tma
    const bool tma = vec_ok && !(env && env[0] == 'v');
    CUtensorMap map, zmap;
    memset(&map, 0, sizeof(map));
    memset(&zmap, 0, sizeof(zmap));
    if (tma) {
        int st = make_canvas_map(&map, p.canvas, p.B, C, p.ny, p.nx, C);
        if (st == HGSF_OK) st = make_canvas_map(&zmap, p.canvas, p.B, C, p.ny, p.nx, C / 4);
        if (st != HGSF_OK) return st;
    }
    size_t smem = 1024 + sizeof(float) * (EMIT_WARPS * C * 32 + (C / 4) * 32 + EMIT_WARPS * 2 * STAGE_W * RWc) +
                  sizeof(int) * 2 * (size_t)(p.B + 1);
#ifdef HGSF_EXPERIMENT
    if (const char *ex = getenv("HGSF_EXTRA_SMEM")) smem += (size_t)atoi(ex);      // occupancy experiments: fewer CTAs per SM
#endif
    const long long n_tiles = (long long)p.B * p.ny * ((p.nx + 31) / 32);
    if (n_tiles == 0) return HGSF_OK;
    // tiles per ticket (see k_emit): sparse scenes -- fewer than 6 points per 32-cell tile on average, most tiles empty -- take 2
    // (measured, ms per step at chunk 1 / 2 / 4: VoD 2 000 points per frame 0.138 / 0.116 / 0.119, VoD 10 000 0.142 / 0.134 /
    // 0.138, TJ4D 30 000 clustered 0.429 / 0.310 / 0.324; VoD 30 000 0.182 / 0.195 / 0.203)
    int chunk = ((long long)p.n < 6 * n_tiles) ? 2 : 1;
    if (const char *tc = getenv("HGSF_TILE_CHUNK")) chunk = atoi(tc) >= 2 ? 2 : 1;
    const bool bn = p.bn_w != nullptr;
    auto go = [&](auto kern) -> int {
        int grid = 1;
        const int st = launch_persistent(kern, EMIT_THREADS, smem, (n_tiles + EMIT_WARPS - 1) / EMIT_WARPS, stream, &grid);
        if (st != HGSF_OK) return st;
        kern<<<(unsigned)grid, EMIT_THREADS, smem, stream>>>(map, zmap, p);
        return (int)cudaGetLastError();
    };
    if (tma && chunk == 2) return bn ? go(k_emit<F, ABS, DIST, true, 0, 2>) : go(k_emit<F, ABS, DIST, false, 0, 2>);
    if (tma) return bn ? go(k_emit<F, ABS, DIST, true, 0, 1>) : go(k_emit<F, ABS, DIST, false, 0, 1>);
    if (vec_ok) return bn ? go(k_emit<F, ABS, DIST, true, 1, 1>) : go(k_emit<F, ABS, DIST, false, 1, 1>);
    return bn ? go(k_emit<F, ABS, DIST, true, 2, 1>) : go(k_emit<F, ABS, DIST, false, 2, 1>);
}

static int launch_emit(const PathParams &p, bool abs_xyz, bool dist, cudaStream_t s) {
    if (p.C != 64) return HGSF_ERR_UNSUPPORTED;
#define HGSF_CASE(FV, A, D) if (p.F == FV && abs_xyz == A && dist == D) return launch_emit_t<FV, A, D>(p, s);
    HGSF_CASE(4, true, false) HGSF_CASE(5, true, false) HGSF_CASE(6, true, false) HGSF_CASE(7, true, false)
    HGSF_CASE(8, true, false) HGSF_CASE(7, false, false) HGSF_CASE(8, false, false)
    HGSF_CASE(7, true, true) HGSF_CASE(8, true, true)
#undef HGSF_CASE
    return HGSF_ERR_UNSUPPORTED;
}

// ---- optional per-launch timing of the dominant kernel (k_emit, or k_pfn without a canvas): bench.py's roofline leg ---------------------------------
// A ring of CUDA event pairs recorded on the launching stream around that launch.  Off by default.
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

int launch_pillar_path(const PathParams &p_in, bool with_pfn, bool abs_xyz, bool dist, size_t zero_bytes, void *zero_base,
                       cudaStream_t stream, int *launches) {
    PathParams p = p_in;
#ifdef HGSF_EXPERIMENT
    { static const int dbg = getenv("HGSF_DBG") ? atoi(getenv("HGSF_DBG")) : 0; p.dbg = dbg; }
#endif
    int nl = 0;
    (void)zero_bytes; (void)zero_base;      // the cell table is zeroed by k_front itself
    {
        // cooperative launch: every CTA must be resident, so the grid is the occupancy limit (capped: ~4 CTAs/SM is
        // plenty of parallelism for a latency-bound front end and keeps the grid barriers cheap)
        static int max_ctas = []() {
            int per_sm = 1;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_front, FRONT_THREADS, 0);
            const char *env = getenv("HGSF_FRONT_CTAS");
            const int cap = env ? atoi(env) : 4;
            return sm_count() * std::max(1, std::min(per_sm, cap));
        }();
        const long long want = std::max<long long>(((long long)p.B * p.cells + FRONT_THREADS * 8 - 1) / (FRONT_THREADS * 8),
                                                   ((long long)p.n + FRONT_THREADS - 1) / FRONT_THREADS);
        const int grid = (int)std::max<long long>(1, std::min<long long>(want, max_ctas));
        PathParams pp = p;
        void *args[] = {&pp};
        cudaError_t e = cudaLaunchCooperativeKernel((const void *)k_front, dim3(grid), dim3(FRONT_THREADS), args, 0, stream);
        if (e != cudaSuccess) return (int)e;
        ++nl;
    }
    // with a canvas: the fused k_emit (pillar rows + canvas in one pass); without: the pillar-major k_pfn
    const bool timed = g_timing.capacity > 0 && g_timing.count < g_timing.capacity;
    if (timed) cudaEventRecord(g_timing.ev[2 * g_timing.count], stream);
    const int st = (with_pfn && p.canvas) ? launch_emit(p, abs_xyz, dist, stream) : launch_pfn(p, with_pfn, abs_xyz, dist, stream);
    if (st != HGSF_OK) return st;
    if (timed) cudaEventRecord(g_timing.ev[2 * g_timing.count++ + 1], stream);
    ++nl;
    if (launches) *launches = nl;
    return HGSF_OK;
}

}  // namespace hgsf
