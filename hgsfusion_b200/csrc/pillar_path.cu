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

// A tile's plan: which of its 32 cells hold a (kept) pillar, and where that pillar's points sit in
// the staging buffer (off < 0: not staged -- more than 32 arrivals or the buffer is full -- read
// straight from global memory).
struct TilePlan {
    int nocc;
    int cell[32], m[32], cnt[32], start[32], off[32];
};
constexpr int STAGE_ROWS = 256;   // staged point rows per tile (typical tiles hold ~10, dense blobs ~200)

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Persistent CTAs walk the canvas tiles (32 cells of one BEV row x all C channels).  Software pipeline,
// per iteration i:   plan(i+1) from the prefetched table entries  ->  cp.async gather of tile i+1's point
// rows into stage[(i+1)&1]  ->  compute tile i from stage[i&1]  ->  one TMA tensor store of the tile.
template <int F, bool ABS, bool DIST, int C, int NWARPS, bool PFN, bool TMA>
__global__ void __launch_bounds__(NWARPS * 32) k_emit(const __grid_constant__ CUtensorMap tmap, const PathParams p) {
    using Lane = PfnLane<F, ABS, DIST, C>;
    constexpr int CPL = Lane::CPL;
    constexpr int RWc = (F + 1 + 3) / 4 * 4;   // F features + the point index, padded to float4
    constexpr int TILE = C * 32;               // floats per canvas tile: C channel rows of 32 cells (128 B each)
    constexpr int NT = NWARPS * 32;
    const int Fr = PFN ? F : p.F, RW = PFN ? RWc : p.RW, NV = RW >> 2;

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    float *tilebuf = reinterpret_cast<float *>(smem_raw);        // [2][TILE]
    float *zerobuf = tilebuf + 2 * TILE;                          // [TILE]
    float *stage = zerobuf + TILE;                                // [2][STAGE_ROWS * RW]
    int *s_R = reinterpret_cast<int *>(stage + 2 * STAGE_ROWS * RW);   // [B+1] raw pillar base per frame
    int *s_K = s_R + (p.B + 1);                                   // [B+1] kept (final) pillar base per frame
    __shared__ TilePlan s_plan[2];
    __shared__ int s_perm[NWARPS][32];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool canvas_on = PFN && (p.canvas != nullptr);

    for (int b = tid; b <= p.B; b += NT) s_R[b] = p.frame_raw_base[b];
    if (canvas_on)
        for (int t = tid; t < TILE; t += NT) zerobuf[t] = 0.f;
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

    // per-lane PFN constants: lane owns channels lane, lane+32, ...
    Lane pfn;
    if (PFN) pfn.load(PfnArgs{p.W, p.bias, p.bn_w, p.bn_b, p.bn_m, p.bn_v, p.eps}, lane);

    const int tiles_per_row = (p.nx + 31) >> 5;
    const int rows_per_frame = p.nz * p.ny;
    const int n_rows = p.B * rows_per_frame;
    const int P4 = (p.P >> 2) << 2;
    TileStep step;
    step.tiles_per_row = tiles_per_row; step.rows_per_frame = rows_per_frame;
    step.dr = (int)gridDim.x / tiles_per_row; step.dxt = (int)gridDim.x - step.dr * tiles_per_row;

    auto load_entry = [&](const TilePos &t) -> uint4 {
        const int x = t.xt * 32 + lane;
        // row r = (b*nz + z)*ny + y and the table is [b][z][y][x]: the cell index is r*nx + x
        return (t.r < n_rows && x < p.nx) ? __ldg(reinterpret_cast<const uint4 *>(p.table + (size_t)t.r * p.nx + x))
                                          : make_uint4(0, 0, 0, 0);
    };
    // warp 0: entries -> plan
    auto make_plan = [&](TilePlan &q, const uint4 e, const TilePos &t) {
        const int b = t.r < n_rows ? t.b : 0;
        const bool occ = (e.x != 0u) && ((int)(e.x - 1u) - s_R[b] < p.max_voxels);
        const unsigned bal = __ballot_sync(FULL, occ);
        const int need = (occ && e.y <= 32u) ? (int)e.y : 0;
        int incl = need;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(FULL, incl, d);
            if (lane >= d) incl += o;
        }
        if (occ) {
            const int k = __popc(bal & ((1u << lane) - 1u));
            q.cell[k] = lane; q.m[k] = (int)(e.x - 1u); q.cnt[k] = (int)e.y; q.start[k] = (int)e.z;
            q.off[k] = (need > 0 && incl <= STAGE_ROWS) ? (incl - need) : -1;
        }
        if (lane == 0) q.nocc = __popc(bal);
    };
    // all warps: start the gather of a planned tile (warp per pillar, lane per point)
    auto issue_gather = [&](const TilePlan &q, float *stg) {
        const int nocc = q.nocc;
        for (int k = warp; k < nocc; k += NWARPS) {
            const int off = q.off[k];
            if (off >= 0 && lane < q.cnt[k]) {
                const float *src = p.sorted_rows + (size_t)(q.start[k] + lane) * RW;
                float *dst = stg + (size_t)(off + lane) * RW;
                for (int v = 0; v < NV; ++v) cp_async16(dst + 4 * v, src + 4 * v);
            }
        }
        cp_async_commit();
    };

    // the only divisions of the kernel: where this CTA starts
    TilePos cur;
    cur.r = (int)blockIdx.x / tiles_per_row; cur.xt = (int)blockIdx.x - cur.r * tiles_per_row;
    cur.b = cur.r / rows_per_frame; cur.zy = cur.r - cur.b * rows_per_frame;
    TilePos nxt = cur;
    step.advance(nxt);
    TilePos nxt2 = nxt;
    step.advance(nxt2);
    uint4 e_next = make_uint4(0, 0, 0, 0);
    __syncthreads();                                       // s_R / s_K / zerobuf ready
    if (warp == 0) {
        make_plan(s_plan[0], load_entry(cur), cur);
        e_next = load_entry(nxt);
    }
    __syncthreads();
    issue_gather(s_plan[0], stage);

    int nb = 0;   // non-empty tiles so far (selects the smem tile buffer)
    for (int it = 0; cur.r < n_rows; ++it, cur = nxt, nxt = nxt2, step.advance(nxt2)) {
        const int slot = it & 1;
        const TilePlan &q = s_plan[slot];
        const float *stg = stage + (size_t)slot * STAGE_ROWS * RW;
        const int x0 = cur.xt * 32;
        const int b = cur.b, zy = cur.zy;
        const int z = (p.nz == 1) ? 0 : zy / p.ny;
        const int y = zy - z * p.ny;
        // (1) plan the next tile, prefetch the entries of the one after it
        if (warp == 0) {
            make_plan(s_plan[slot ^ 1], e_next, nxt);
            e_next = load_entry(nxt2);
            if (canvas_on && TMA && lane == 0 && q.nocc > 0) tma_wait_read<1>();   // tile buffer of two tiles ago is free
        }
        __syncthreads();   // (A)
        // (2) start the next tile's gather, clear this tile's buffer
        issue_gather(s_plan[slot ^ 1], stage + (size_t)(slot ^ 1) * STAGE_ROWS * RW);
        const int n_occ = q.nocc;
        float *tb = tilebuf + (nb & 1) * TILE;
        if (n_occ > 0) {
            ++nb;
            if (canvas_on)
                for (int t = tid * 4; t < TILE; t += NT * 4) *reinterpret_cast<float4 *>(tb + t) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        cp_async_wait<1>();   // this tile's rows have landed (for this thread's copies)
        __syncthreads();      // (B) ... and everybody else's; tile buffer cleared
        if (n_occ == 0) {
            if (canvas_on) {
                if (TMA) {
                    if (tid == 0) { tma_store_3d(&tmap, zerobuf, x0, zy, b * C); tma_commit(); }
                } else {
                    for (int c = warp; c < C; c += NWARPS)
                        if (x0 + lane < p.nx) p.canvas[(((size_t)b * C + c) * p.ny + y) * p.nx + x0 + lane] = 0.f;
                }
            }
            continue;
        }
        // (3) compute: warp per pillar
        for (int k = warp; k < n_occ; k += NWARPS) {
            const int cell = q.cell[k], m = q.m[k], cnt = q.cnt[k], start = q.start[k], off = q.off[k];
            const int f = s_K[b] + (m - s_R[b]);          // final pillar id (first-seen order, frames concatenated)
            const int n_keep = min(cnt, p.P);
            const bool staged = off >= 0;
            const float *grow = p.sorted_rows + (size_t)start * RW;   // the pillar's rows in global memory
            const float *srow = stg + (size_t)(staged ? off : 0) * RW; // ... and in the staging buffer
            // ---- order the cell's points by input index, keep the first P ----
            if (cnt == 1) {
                if (lane == 0) s_perm[warp][0] = 0;
            } else if (cnt <= 32) {
                uint32_t mine = 0xFFFFFFFFu;
                if (lane < cnt) mine = staged ? __float_as_uint(srow[lane * RW + Fr]) : __float_as_uint(__ldg(grow + (size_t)lane * RW + Fr));
                int rank = 0;
                for (int qq = 0; qq < cnt; ++qq) rank += (__shfl_sync(FULL, mine, qq) < mine) ? 1 : 0;
                if (lane < cnt) s_perm[warp][rank] = lane;
            } else {
                s_perm[warp][lane] = select_first32(grow + Fr, RW, cnt, lane);
            }
            __syncwarp();
            if (lane == 0) {
                p.num[f] = n_keep;
                *reinterpret_cast<int4 *>(p.coords + 4 * (size_t)f) = make_int4(b, z, y, x0 + cell);
            }
            if (p.voxels) {
                float *vo = p.voxels + (size_t)f * p.P * Fr;
                for (int t = lane; t < p.P * Fr; t += 32) {
                    const int s = t / Fr, kk = t - s * Fr;
                    float v = 0.f;
                    if (s < n_keep) v = staged ? srow[s_perm[warp][s] * RW + kk] : __ldg(grow + (size_t)s_perm[warp][s] * RW + kk);
                    vo[t] = v;
                }
            }
            if (PFN) {
                auto load_row = [&](int s, float (&rowf)[RWc]) {
                    const int pp = s_perm[warp][s];
#pragma unroll
                    for (int v = 0; v < RWc / 4; ++v) {
                        const float4 t4 = staged ? *reinterpret_cast<const float4 *>(srow + pp * RWc + 4 * v)
                                                 : __ldg(reinterpret_cast<const float4 *>(grow + (size_t)pp * RWc) + v);
                        rowf[4 * v] = t4.x; rowf[4 * v + 1] = t4.y; rowf[4 * v + 2] = t4.z; rowf[4 * v + 3] = t4.w;
                    }
                };
                // pillar centre: fl(fl(c*v)+off), two roundings, no FMA (pillar_vfe.py:101-103)
                const float cx = __fadd_rn(__fmul_rn((float)(x0 + cell), p.vsize[0]), p.voff[0]);
                const float cy = __fadd_rn(__fmul_rn((float)y, p.vsize[1]), p.voff[1]);
                const float cz = __fadd_rn(__fmul_rn((float)z, p.vsize[2]), p.voff[2]);
                float vmax[CPL];
                pfn.init_max(vmax, n_keep < p.P);
                float rowf[RWc];
                if (n_keep == 1) {
                    // mean of one point is the point (x/1 is exact): skip the slot sum
                    load_row(0, rowf);
                    pfn.point(rowf, rowf[0], rowf[1], rowf[2], cx, cy, cz, vmax);
                } else {
                    // ---- mean of the kept points (torch CPU sum order, pfn.cuh) ----
                    SlotSum sum;
                    for (int s = 0; s < n_keep; ++s) {
                        const int pp = s_perm[warp][s];
                        const float4 v = staged ? *reinterpret_cast<const float4 *>(srow + pp * RWc)
                                                : __ldg(reinterpret_cast<const float4 *>(grow + (size_t)pp * RWc));
                        sum.add(s, P4, v.x, v.y, v.z);
                    }
                    const float fn = (float)n_keep;
                    const float mx = __fdiv_rn(sum.sx(), fn), my = __fdiv_rn(sum.sy(), fn), mz = __fdiv_rn(sum.sz(), fn);
                    for (int s = 0; s < n_keep; ++s) {
                        load_row(s, rowf);
                        pfn.point(rowf, mx, my, mz, cx, cy, cz, vmax);
                    }
                }
#pragma unroll
                for (int j = 0; j < CPL; ++j) {
                    const int c = lane + 32 * j;
                    if (p.feats) p.feats[(size_t)f * C + c] = vmax[j];
                    if (canvas_on) tb[swz128(c, cell)] = vmax[j];
                }
            }
            __syncwarp();
        }
        // (4) the tile goes out in one piece
        if (canvas_on) {
            if (TMA) {
                fence_proxy_async_smem();
                __syncthreads();   // (C)
                if (tid == 0) { tma_store_3d(&tmap, tb, x0, zy, b * C); tma_commit(); }
            } else {
                __syncthreads();
                for (int c = warp; c < C; c += NWARPS)
                    if (x0 + lane < p.nx) p.canvas[(((size_t)b * C + c) * p.ny + y) * p.nx + x0 + lane] = tb[swz128(c, lane)];
            }
        } else {
            __syncthreads();       // plan / stage buffers are recycled two iterations later
        }
    }
    cp_async_wait<0>();
    if (canvas_on && TMA && tid == 0) tma_wait_read<0>();
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
int make_canvas_map(CUtensorMap *map, float *canvas, int B, int C, int ny, int nx) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return HGSF_ERR_DRIVER;
    const cuuint64_t gdim[3] = {(cuuint64_t)nx, (cuuint64_t)ny, (cuuint64_t)B * C};
    const cuuint64_t gstr[2] = {(cuuint64_t)nx * 4, (cuuint64_t)nx * ny * 4};
    const cuuint32_t box[3] = {32, 1, (cuuint32_t)C};
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

constexpr int EMIT_WARPS = 4;

template <int F, bool ABS, bool DIST, int C, bool PFN>
static int launch_emit_t(const PathParams &p, cudaStream_t stream) {
    const bool canvas_on = PFN && p.canvas;
    const bool tma = canvas_on && (p.nx % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.canvas) & 15) == 0) && C <= 256;
    CUtensorMap map;
    memset(&map, 0, sizeof(map));
    if (tma) {
        const int st = make_canvas_map(&map, p.canvas, p.B, C, p.ny, p.nx);
        if (st != HGSF_OK) return st;
    }
    const size_t smem = sizeof(float) * (3 * C * 32 + 2 * STAGE_ROWS * (size_t)p.RW) + sizeof(int) * 2 * (size_t)(p.B + 1);
    const long long n_tiles = (long long)p.B * p.nz * p.ny * ((p.nx + 31) / 32);
    if (n_tiles == 0) return HGSF_OK;
    auto go = [&](auto kern) -> int {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        int per_sm = 1;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, EMIT_WARPS * 32, smem);
        if (e != cudaSuccess) return (int)e;
        if (per_sm < 1) per_sm = 1;
        const long long grid = std::min<long long>(n_tiles, (long long)sm_count() * per_sm);
        kern<<<(unsigned)grid, EMIT_WARPS * 32, smem, stream>>>(map, p);
        return (int)cudaGetLastError();
    };
    if constexpr (PFN) {
        if (tma) return go(k_emit<F, ABS, DIST, C, EMIT_WARPS, PFN, true>);
    }
    return go(k_emit<F, ABS, DIST, C, EMIT_WARPS, PFN, false>);
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
    return launch_emit_t<4, true, false, 32, false>(p, s);   // F / RW are read from the params when PFN is off
}

int launch_pillar_path(const PathParams &p, bool with_pfn, bool abs_xyz, bool dist, size_t zero_bytes, void *zero_base,
                       cudaStream_t stream, int *launches) {
    int nl = 0;
    cudaError_t e = cudaMemsetAsync(zero_base, 0, zero_bytes, stream);
    if (e != cudaSuccess) return (int)e;
    ++nl;
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
    int st;
    if (with_pfn) {
        if (p.C == 64) st = launch_emit_pfn<64>(p, abs_xyz, dist, stream);
        else st = HGSF_ERR_UNSUPPORTED;
    } else {
        st = launch_emit_plain(p, stream);
    }
    if (st != HGSF_OK) return st;
    ++nl;
    if (launches) *launches = nl;
    return HGSF_OK;
}

}  // namespace hgsf
