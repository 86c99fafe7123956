// pillar_path.cu -- points -> pillars (first-seen order) -> decorate + PFN + max -> BEV canvas.
//
// Two kernels on one stream, no host sync, no allocation:
//   k_front  (cooperative)       cell key of every point; per cell min point index + count (warp-aggregated atomics into the
//                                direct-address cell table); two scans (first points -> raw pillar ids in first-seen order,
//                                cell counts -> CSR starts in cell order + one record per 32-cell tile); fill (features +
//                                point index of every point to its cell's CSR segment)
//   k_emit   (tile major)        the fused consumer, used whenever the canvas is requested: per 32-cell tile orders each
//                                pillar's points by index, keeps the first P, decorates, runs the PFN with the weights in
//                                registers, takes the max; writes voxel_coords / voxel_num_points / pillar_features rows AND
//                                the canvas tile (one TMA tensor store per tile, zeros included)
//   k_pfn    (pillar major)      the consumer without a canvas (hgsf_pillarize, or pillar_features only)
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
#include <climits>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace cg = cooperative_groups;

namespace hgsf {

// ------------------------------------------------------------------------------------------------
// k_front : ONE cooperative kernel for the whole front end (three grid barriers)
//   Every CTA owns a contiguous SLICE of the points and a contiguous slice of the cell table, processed in blocks of 1024
//   (4 per thread: four independent dependency chains in flight).  What a thread learns about the points / cells of its
//   first block stays in REGISTERS across the grid barriers, so with at most one block per CTA (the usual case: 480 000
//   points over 592 CTAs) no phase re-reads what an earlier phase computed.
//   phase 0  (only when the table is not known to be clean) zero the cell table.  Normally skipped: the consumer kernel
//            (k_emit / k_pfn) zeroes every entry it reads and then marks the table clean.
//   phase 1  count : cell key; per cell min point index + count through warp-aggregated atomics
//   -- barrier --
//   phase 2  reduce: (a) over POINTS, "is the first point of its cell" (one gather of the cell tag): per-CTA pillar totals and,
//            in registers, each first point's rank inside the block; (b) over CELLS in table order, the counts: per-CTA totals
//   -- barrier --
//   phase 3  apply : carry-in = sum of the slices before this one.  (a) a first point's exclusive prefix is the raw pillar
//            id -> first-seen order without a sort; written into the cell tag (+ pillar records and the raw id at each frame
//            start).  (b) the exclusive prefix of the counts is the cell's CSR start -> the point rows of a 32-cell canvas tile
//            are CONTIGUOUS in sorted_rows; every tile also gets a record {first row, rows}
//   -- barrier --
//   phase 4  fill  : copies each point's features (+ its index) to its cell's CSR segment
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

// phase 1, split in three so that several points per thread can be in flight at once:
//   point_key     loads the point, derives frame_offsets on the fly, returns the cell key (-1 = outside the grid)
//   count_issue   warp-aggregated atomics: lanes of the same cell elect the lowest lane (= lowest point index), which
//                 adds the group size to the cell count and maxes the inverted index into the tag
//   count_finish  broadcasts the leader's base and stores key + arrival rank
// every lane of the warp calls them (i >= n for the padding lanes)
__device__ __forceinline__ int point_key(const PathParams &p, int i) {
    int key = -1;
    if (i < p.n) {
        const float *row = p.pts + (size_t)i * p.stride;
        float x = __ldg(row + p.xyz_col), y = __ldg(row + p.xyz_col + 1);
        const float z = __ldg(row + p.xyz_col + 2);
        if (p.flags & HGSF_POINTS_FLIP_X) x = -x;        // DataProcessor.double_flip (data_processor.py:116-130): exact sign flips
        if (p.flags & HGSF_POINTS_FLIP_Y) y = -y;
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
        if (ok) key = b * p.cells + (__float2int_rz(qz) * p.ny + __float2int_rz(qy)) * p.nxp + __float2int_rz(qx);
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
__device__ __forceinline__ uint32_t count_finish(const PathParams &p, int i, int key, unsigned base, int leader, int rank) {
    base = __shfl_sync(FULL, base, leader) + (unsigned)rank;
    if (i < p.n) {
        p.key[i] = key;
        p.arrival[i] = base;
    }
    return base;
}

constexpr int FRONT_THREADS = 256;
constexpr int FRONT_WARPS = FRONT_THREADS / 32;
constexpr int PPT = 4;                                  // points per thread and block
constexpr int PBLOCK = FRONT_THREADS * PPT;             // 1024 points per block; thread t holds points base + u*256 + t
constexpr int KC = 1;                                   // cell blocks (1024 cells, one uint4 per thread) kept in registers
static_assert(PBLOCK == SCAN_TILE, "block size");

// block-wide sum of a 64-bit value; every thread gets the total
__device__ __forceinline__ uint64_t block_sum(uint64_t v, uint64_t *s_warp, int lane, int warp) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(FULL, v, d);
    __syncthreads();
    if (lane == 0) s_warp[warp] = v;
    __syncthreads();
    uint64_t t = 0;
#pragma unroll
    for (int w = 0; w < FRONT_WARPS; ++w) t += s_warp[w];
    return t;
}

// first-point flags of one block and their exclusive ranks inside the block (point order = u major, then thread):
// one ballot per 32 points, one 32-entry scan of the (slab, warp) counts by warp 0.  Returns the block's total.
__device__ __forceinline__ uint32_t block_first_ranks(const uint32_t (&f)[PPT], uint32_t (&lrank)[PPT], uint32_t (*s_wcnt)[FRONT_WARPS],
                                                      uint32_t *s_total, int lane, int warp) {
    const unsigned lt = (1u << lane) - 1u;
    uint32_t wpre[PPT];
#pragma unroll
    for (int u = 0; u < PPT; ++u) {
        const unsigned bal = __ballot_sync(FULL, f[u] != 0u);
        wpre[u] = __popc(bal & lt);
        if (lane == 0) s_wcnt[u][warp] = __popc(bal);
    }
    __syncthreads();
    if (warp == 0) {
        static_assert(PPT * FRONT_WARPS == 32, "one warp scans the (slab, warp) counts");
        const uint32_t v = s_wcnt[lane / FRONT_WARPS][lane % FRONT_WARPS];
        uint32_t incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(FULL, incl, d);
            if (lane >= d) incl += o;
        }
        s_wcnt[lane / FRONT_WARPS][lane % FRONT_WARPS] = incl - v;
        if (lane == 31) *s_total = incl;
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < PPT; ++u) lrank[u] = s_wcnt[u][warp] + wpre[u];
    const uint32_t total = *s_total;
    __syncthreads();               // s_wcnt / s_total are rewritten by the next block
    return total;
}

__global__ void __launch_bounds__(FRONT_THREADS, 4) k_front(const PathParams p) {
    cg::grid_group grid = cg::this_grid();
#ifdef HGSF_PHASE_TIMES
    auto stamp = [&](int k) { if (blockIdx.x == 0 && threadIdx.x == 0) { uint64_t t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); reinterpret_cast<uint64_t *>(p.ticket + 96)[k] = t; } };
#else
    auto stamp = [&](int) {};
#endif
    stamp(0);
    __shared__ uint64_t s_warp[FRONT_WARPS];
    __shared__ uint32_t s_wcnt[PPT][FRONT_WARPS];
    __shared__ uint32_t s_total;
    __shared__ uint32_t s_excl[PBLOCK];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cta = (int)blockIdx.x;

    // ---- phase 0: zero the cell table, unless the previous call's consumer kernel left it clean ----
    const bool clean = (ld_volatile_u64(p.state) == p.magic);        // grid-uniform: nobody writes p.state before the first barrier
    if (cta == 0 && tid == 0) { p.ticket[0] = 0u; p.ticket[32] = 0u; p.ticket[64] = 0u; p.ticket[16] = 0u; p.ticket[48] = 0u; }   // run ticket, heavy-tile
                                                          // count, finished CTAs; tickets / finished CTAs of the statistics pass
    if (p.stats && cta == 0 && tid < 16 * p.Cin) p.stats_S[tid] = 0.0;      // 16 * Cin <= 240 < FRONT_THREADS
    if ((p.flags & HGSF_POINTS_SPCONV1_BREAK) && cta == 0) for (int b = tid; b < p.B; b += FRONT_THREADS) p.cutoff[b] = INT_MAX;
    if (!clean) {
        uint4 *t4 = reinterpret_cast<uint4 *>(p.cell_tag);           // tag, cnt: two arrays back to back
        const long long n4 = (long long)(p.table_bytes >> 4);
        for (long long i = (long long)cta * FRONT_THREADS + tid; i < n4; i += (long long)gridDim.x * FRONT_THREADS)
            t4[i] = make_uint4(0u, 0u, 0u, 0u);
        grid.sync();
    }
    stamp(1);

    // this CTA's slice of the points: [lo, hi), multiples of 32
    const int n_pad = (p.n + 31) & ~31;
    const int lo = (int)min((long long)cta * p.pslice, (long long)n_pad);
    const int hi = (int)min((long long)lo + p.pslice, (long long)n_pad);
    // ... and of the cells: [clo, chi), multiples of 32
    const long long n_cells = (long long)p.B * p.cells;
    const long long clo = min((long long)cta * p.cslice, n_cells);
    const long long chi = min(clo + p.cslice, n_cells);

    // ---- phase 1: count ----
    int key_r[PPT];                 // first block: cell keys ...
    uint32_t arr_r[PPT];            // ... and arrival ranks, kept for the later phases
#pragma unroll
    for (int u = 0; u < PPT; ++u) { key_r[u] = -1; arr_r[u] = 0u; }
    for (int base = lo; base < hi; base += PBLOCK) {
        int k[PPT], l[PPT], r[PPT];
        unsigned bs[PPT];
        bool on[PPT];               // warp-uniform (base, hi and u*256 + warp*32 are multiples of 32): the match stays convergent
#pragma unroll
        for (int u = 0; u < PPT; ++u) {
            const int i = base + u * FRONT_THREADS + tid;
            on[u] = i < hi;
            k[u] = on[u] ? point_key(p, i) : -1;
        }
#pragma unroll
        for (int u = 0; u < PPT; ++u) {
            l[u] = 0; r[u] = 0; bs[u] = 0;
            if (on[u]) bs[u] = count_issue(p, base + u * FRONT_THREADS + tid, k[u], lane, l[u], r[u]);
        }
#pragma unroll
        for (int u = 0; u < PPT; ++u) {
            uint32_t a = 0u;
            if (on[u]) a = count_finish(p, base + u * FRONT_THREADS + tid, k[u], bs[u], l[u], r[u]);
            if (base == lo) { key_r[u] = k[u]; arr_r[u] = a; }
        }
    }
    grid.sync();
    stamp(2);
    if (cta == 0 && tid == 0) st_volatile_u64(p.state, 0ull);        // the table is in use: dirty until the consumer kernel has cleaned it

    // ---- phase 2: reduce ----
    // (b) first, so that its loads are in flight under (a)'s gathers: the counts of this CTA's cells, 4 consecutive cells per thread
    uint4 cv[KC];
    uint32_t csum = 0, osum = 0;    // points / occupied cells of this CTA's slice
    auto occ4 = [](const uint4 v) -> uint32_t { return (v.x ? 1u : 0u) + (v.y ? 1u : 0u) + (v.z ? 1u : 0u) + (v.w ? 1u : 0u); };
#pragma unroll
    for (int k = 0; k < KC; ++k) {
        const long long c0 = clo + (long long)k * PBLOCK + 4 * tid;
        cv[k] = make_uint4(0u, 0u, 0u, 0u);
        if (c0 < chi) cv[k] = *reinterpret_cast<const uint4 *>(p.cell_cnt + c0);
    }
    // (a) flags of a block: the point's cell tag still holds ~(smallest point index of the cell)
    auto flags_of = [&](int base, bool first, int (&keys)[PPT], uint32_t (&f)[PPT]) {
#pragma unroll
        for (int u = 0; u < PPT; ++u) {
            const int i = base + u * FRONT_THREADS + tid;
            keys[u] = first ? key_r[u] : ((i < hi && i < p.n) ? p.key[i] : -1);
        }
#pragma unroll
        for (int u = 0; u < PPT; ++u) {
            const int i = base + u * FRONT_THREADS + tid;
            f[u] = 0u;
            if (keys[u] >= 0) f[u] = (p.cell_tag[keys[u]] == 0xFFFFFFFFu - (uint32_t)i) ? 1u : 0u;
        }
    };
    uint32_t fbits = 0u, lrank_r[PPT];          // first block: flags and ranks inside the block
    uint32_t btotal_r = 0u;
    uint32_t ptotal = 0u;
#pragma unroll
    for (int u = 0; u < PPT; ++u) lrank_r[u] = 0u;
    for (int base = lo; base < hi; base += PBLOCK) {
        int keys[PPT];
        uint32_t f[PPT], lr[PPT];
        flags_of(base, base == lo, keys, f);
        const uint32_t bt = block_first_ranks(f, lr, s_wcnt, &s_total, lane, warp);
        if (base == lo) {
#pragma unroll
            for (int u = 0; u < PPT; ++u) { fbits |= f[u] << u; lrank_r[u] = lr[u]; }
            btotal_r = bt;
        }
        ptotal += bt;
    }
#pragma unroll
    for (int k = 0; k < KC; ++k) { csum += cv[k].x + cv[k].y + cv[k].z + cv[k].w; osum += occ4(cv[k]); }
    for (long long c0 = clo + (long long)KC * PBLOCK + 4 * tid; c0 < chi; c0 += PBLOCK) {
        const uint4 v = *reinterpret_cast<const uint4 *>(p.cell_cnt + c0);
        csum += v.x + v.y + v.z + v.w; osum += occ4(v);
    }
    {
        const uint64_t ctotal = block_sum(((uint64_t)osum << 32) | csum, s_warp, lane, warp);
        if (tid == 0) {
            p.scan_desc[cta] = ptotal; p.scan_desc[MAX_FRONT_CTAS + cta] = (uint32_t)ctotal;
            p.scan_desc[2 * MAX_FRONT_CTAS + cta] = (uint32_t)(ctotal >> 32);
        }
    }
    grid.sync();
    stamp(3);

    // ---- phase 3: apply ----
    uint32_t prun, crun, orun;      // pillars (point slices) / points / occupied cells (cell slices) of all slices before this one
    {
        uint64_t before = 0, obefore = 0;
        for (int c = tid; c < cta; c += FRONT_THREADS) {
            before += ((uint64_t)p.scan_desc[c] << 32) | (uint64_t)p.scan_desc[MAX_FRONT_CTAS + c];
            obefore += p.scan_desc[2 * MAX_FRONT_CTAS + c];
        }
        before = block_sum(before, s_warp, lane, warp);
        obefore = block_sum(obefore, s_warp, lane, warp);
        prun = (uint32_t)(before >> 32); crun = (uint32_t)before; orun = (uint32_t)obefore;
    }
    for (int base = lo; base < hi; base += PBLOCK) {
        int keys[PPT];
        uint32_t f[PPT], lr[PPT], bt;
        if (base == lo) {
#pragma unroll
            for (int u = 0; u < PPT; ++u) { keys[u] = key_r[u]; f[u] = (fbits >> u) & 1u; lr[u] = lrank_r[u]; }
            bt = btotal_r;
        } else {
            flags_of(base, false, keys, f);
            bt = block_first_ranks(f, lr, s_wcnt, &s_total, lane, warp);
        }
#pragma unroll
        for (int u = 0; u < PPT; ++u) {
            const uint32_t pillars = prun + lr[u];
            s_excl[u * FRONT_THREADS + tid] = pillars;
            if (f[u]) {
                p.cell_tag[keys[u]] = pillars + 1u;   // raw pillar id + 1 (disjoint from the ~i values phase 1 left)
                // beyond the first block the flag does not survive in a register: park it in the top bit of the arrival rank
                if (base != lo) p.arrival[base + u * FRONT_THREADS + tid] |= 0x80000000u;
            }
        }
        __syncthreads();
        // raw pillar id at each frame start
        const int bend = min(base + PBLOCK, hi);
        const bool last = (bend >= p.n);                 // this block holds the last point
        for (int b = tid; b <= p.B; b += FRONT_THREADS) {
            const int o = p.frame_offsets[b];
            if (o >= base && o < bend && o < p.n) p.frame_raw_base[b] = (int32_t)s_excl[o - base];
            else if (last && o >= p.n) p.frame_raw_base[b] = (int32_t)(prun + bt);
        }
        prun += bt;
        __syncthreads();
    }
    if (p.n == 0 && cta == 0)
        for (int b = tid; b <= p.B; b += FRONT_THREADS) p.frame_raw_base[b] = 0;
    // (b) cells: CSR start of every cell (empty ones too), and for every 32-cell tile a record {first CSR row, rows, pillars in
    //     cell order before the tile, occupancy mask}: the exclusive prefix of "occupied" numbers the pillars in CELL order
    //     (the pillar-major consumer walks them in that order, the canvas writer finds a cell's pillar from mask + base)
    {
        int k = 0;
        for (long long cb = clo; cb < chi; cb += PBLOCK, ++k) {
            const long long c0 = cb + 4 * tid;
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (k < KC) {
#pragma unroll
                for (int q = 0; q < KC; ++q) if (q == k) v = cv[q];
            } else if (c0 < chi) {
                v = *reinterpret_cast<const uint4 *>(p.cell_cnt + c0);
            }
            const uint32_t nib = (v.x ? 1u : 0u) | (v.y ? 2u : 0u) | (v.z ? 4u : 0u) | (v.w ? 8u : 0u);
            const uint64_t local = ((uint64_t)__popc(nib) << 32) | (uint64_t)(v.x + v.y + v.z + v.w);    // (occupied, points)
            uint64_t incl = local;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint64_t o = __shfl_up_sync(FULL, incl, d);
                if (lane >= d) incl += o;
            }
            __syncthreads();
            if (lane == 31) s_warp[warp] = incl;
            __syncthreads();
            uint64_t warp_off = 0, blk_total = 0;
#pragma unroll
            for (int w = 0; w < FRONT_WARPS; ++w) {
                const uint64_t t = s_warp[w];
                if (w < warp) warp_off += t;
                blk_total += t;
            }
            const uint64_t excl = warp_off + (incl - local);
            const uint32_t run = crun + (uint32_t)excl, orn = orun + (uint32_t)(excl >> 32);
            // the tile = the 8 threads of an aligned lane group (4 cells each)
            uint32_t tsum = (uint32_t)local, tmask = nib << (4 * (lane & 7));
#pragma unroll
            for (int d = 1; d < 8; d <<= 1) { tsum += __shfl_xor_sync(FULL, tsum, d); tmask |= __shfl_xor_sync(FULL, tmask, d); }
            if (c0 < chi) {
                *reinterpret_cast<uint4 *>(p.cell_start + c0) = make_uint4(run, run + v.x, run + v.x + v.y, run + v.x + v.y + v.z);
                if ((lane & 7) == 0) {
                    p.tile_rec[c0 >> 5] = make_uint4(run, tsum, orn, tmask);
                    // a tile with many points takes one warp a long time: listed, so that the consumer starts on those first
                    if (tsum > (uint32_t)p.heavy_pts) {
                        p.heavy_list[atomicAdd(p.ticket + 32, 1u)] = (uint32_t)(c0 >> 5);
                    }
                }
            }
            crun += (uint32_t)blk_total; orun += (uint32_t)(blk_total >> 32);
        }
    }
    grid.sync();
    stamp(4);
#ifdef HGSF_PDL_TRIGGER
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");      // k_pillars may be scheduled from here on
#endif

    // ---- phase 4: fill.  Per point: table entries of its cell -> destination row; then the row is copied ----
    for (int base = lo; base < hi; base += PBLOCK) {
        int idx[PPT], key[PPT];
        bool on[PPT];
        uint32_t tag[PPT], start[PPT], arr[PPT];
#pragma unroll
        for (int u = 0; u < PPT; ++u) {
            idx[u] = base + u * FRONT_THREADS + tid;
            const bool in = idx[u] < hi && idx[u] < p.n;
            key[u] = (base == lo) ? key_r[u] : (in ? p.key[idx[u]] : -1);
            on[u] = in && key[u] >= 0;
        }
        bool first[PPT];            // the first point of its cell: it also writes the cell's entry of the pillar list
#pragma unroll
        for (int u = 0; u < PPT; ++u) {
            tag[u] = start[u] = arr[u] = 0u;
            first[u] = false;
            if (on[u]) {
                tag[u] = p.cell_tag[key[u]]; start[u] = p.cell_start[key[u]];
                if (base == lo) { arr[u] = arr_r[u]; first[u] = (fbits >> u) & 1u; }
                else { const uint32_t a = p.arrival[idx[u]]; arr[u] = a & 0x7FFFFFFFu; first[u] = (a >> 31) != 0u; }
            }
        }
        // pillar list in CELL order: slot = pillars before the tile + occupied cells before this one inside the tile
        {
            uint4 rec[PPT];
            uint32_t cn[PPT];
#pragma unroll
            for (int u = 0; u < PPT; ++u)
                if (first[u]) { rec[u] = __ldg(p.tile_rec + (key[u] >> 5)); cn[u] = p.cell_cnt[key[u]]; }
#pragma unroll
            for (int u = 0; u < PPT; ++u)
                if (first[u]) {
                    const uint32_t slot = rec[u].z + (uint32_t)__popc(rec[u].w & ((1u << (key[u] & 31)) - 1u));
                    p.pil[slot] = make_int4(key[u], (int)(tag[u] - 1u), (int)cn[u], (int)start[u]);
                }
        }
        if (p.max_voxels < p.cells) {
#pragma unroll
            for (int u = 0; u < PPT; ++u) {
                if (on[u]) {
                    const int b = (int)fastdiv((uint32_t)key[u], p.div_cells);
                    const int local = (int)(tag[u] - 1u) - p.frame_raw_base[b];
                    on[u] = local < p.max_voxels;             // pillar beyond max_voxels: never created
                    // spconv 1.x: the first point of the first refused pillar is where the voxelization loop stopped
                    if ((p.flags & HGSF_POINTS_SPCONV1_BREAK) && first[u] && local == p.max_voxels) p.cutoff[b] = idx[u];
                }
            }
        }
        for (int k = 0; k < p.RW; k += 4) {
            float4 v[PPT];
#pragma unroll
            for (int u = 0; u < PPT; ++u) {
                if (!on[u]) continue;
                const float *src = p.pts + (size_t)idx[u] * p.stride + p.xyz_col;
                // F features, then the point index (slot F) that k_emit / k_pfn order the pillar by
                const float fi = __int_as_float(idx[u]);
                v[u].x = (k + 0 < p.F) ? __ldg(src + k + 0) : (k + 0 == p.F ? fi : 0.f);
                v[u].y = (k + 1 < p.F) ? __ldg(src + k + 1) : (k + 1 == p.F ? fi : 0.f);
                v[u].z = (k + 2 < p.F) ? __ldg(src + k + 2) : (k + 2 == p.F ? fi : 0.f);
                v[u].w = (k + 3 < p.F) ? __ldg(src + k + 3) : (k + 3 == p.F ? fi : 0.f);
                if (k == 0) {
                    if (p.flags & HGSF_POINTS_FLIP_X) v[u].x = -v[u].x;
                    if (p.flags & HGSF_POINTS_FLIP_Y) v[u].y = -v[u].y;
                }
            }
#pragma unroll
            for (int u = 0; u < PPT; ++u) {
                if (!on[u]) continue;
                const size_t pos = (size_t)start[u] + arr[u];
                reinterpret_cast<float4 *>(p.sorted_rows + pos * p.RW)[k >> 2] = v[u];
            }
        }
    }
    stamp(5);
}

// The consumer kernel (k_emit / k_pfn) has zeroed every table entry it read; the last CTA to finish marks the table clean
// for the next call's k_front.  Every thread of the CTA calls this at the end of the kernel.
__device__ __forceinline__ void mark_table_clean(const PathParams &p) {
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned prev = atomicAdd(p.ticket + 64, 1u);
        if (prev == gridDim.x - 1u) {
            __threadfence();
            st_volatile_u64(p.state, p.magic);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// k_pillars
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

// cnt > 32 arrivals in a cell: positions of the 32 smallest point indices above `lo` (all of them when !have_lo), ascending, one per
// lane; `key_out` = the lane's point index (0xFFFFFFFF when fewer than lane+1 are left).  Called once per 32 slots of the pillar.
__device__ __noinline__ int select_next32(const float *__restrict__ idx0, int stride, int cnt, int lane, bool have_lo, uint32_t lo,
                                          uint32_t &key_out) {
    uint32_t best = 0xFFFFFFFFu;
    int bestv = 0;
    for (int base = 0; base < cnt; base += 32) {
        const int j = base + lane;
        uint32_t k = (j < cnt) ? __float_as_uint(__ldg(idx0 + (size_t)j * stride)) : 0xFFFFFFFFu;
        if (have_lo && k <= lo) k = 0xFFFFFFFFu;          // taken by an earlier round
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
    key_out = best;
    return bestv;
}

// ---- k_pillars ------------------------------------------------------------------------------------
// The consumer of k_front's tables: one persistent kernel of autonomous warps.
//   A warp takes a RUN of consecutive 32-cell tiles of the cell table (one global ticket, fetched a run ahead; the tiles'
//   records {first CSR row, rows, pillars before the tile, occupancy mask} likewise) and cuts it into GROUPS of whole
//   tiles holding at most 32 pillars.  Per group:
//     * lane j OWNS pillar j (k_front's pillar list is in cell order, so the group's pillars are consecutive entries):
//       voxel_coords / voxel_num_points, ordering by point index, first P kept, mean in torch's summation order.  The CSR is in
//       cell order too: the group's point rows are ONE contiguous span of sorted_rows, staged with a single cooperative cp.async
//       copy issued a group ahead;
//     * the arithmetic is cut into UNITS of (pillar, 4 output channels): lane l always computes channels 4*(l&15)..+3, so its
//       Linear weights (packed pairs, one FFMA2 per channel pair and input feature) and BatchNorm constants stay in REGISTERS for
//       the whole kernel (CUDA-core FMA: a 13x64 contraction is far below a tensor-core tile).  Single-point pillars are paired
//       across the half-warps; a multi-point pillar is taken by both halves, which split its points and max-combine.  Results go
//       to pillar_features[f] (f = first-seen id) and, channel major, into the warp's feature block [64][32] in shared memory;
//     * the warp then writes the group's canvas tiles itself (64 channels x 32 cells each, zeros included: the canvas is written
//       exactly once): lane = (4 cells, 4 channel rows), values picked out of the feature block by the tile's occupancy mask,
//       coalesced 16-byte streaming stores -- no second pass over the features, no hand-shake between warps.
//   Runs are short (2 tiles, 1 in the last quarter of the table): the tiles in flight on the 1776 warps then form a compact window
//   moving through the canvas, which the L2 -> DRAM write-back rewards (measured, VoD clustered, us per launch at 1 / 2 / 4 / 8 /
//   16 tiles per run: 163 / 149 / 150 / 166 / 202), and the last runs handed out are short, so the warps finish together.
constexpr int PW = 4;                 // warps per CTA
constexpr int PT = PW * 32;
constexpr int SMALL_CNT = 6;          // up to this many arrivals the owning lane ranks them itself
#ifndef HGSF_MAX_P
#define HGSF_MAX_P 128
#endif
constexpr int MAX_P = HGSF_MAX_P;     // largest max_points_per_voxel (the rank list of a > 32-point pillar lives in shared memory)
#ifndef HGSF_RUN_TILES
#define HGSF_RUN_TILES 2
#endif
constexpr int CT = HGSF_RUN_TILES;    // tiles per run (= per global ticket), at most 32
#ifndef HGSF_RUN_TILES_TAIL
#define HGSF_RUN_TILES_TAIL 1
#endif
constexpr int CT_TAIL = HGSF_RUN_TILES_TAIL;   // ... for the last quarter of the table

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

struct Run {          // consecutive tiles of the cell table
    int t0, nt;       // first tile, tiles (0 = none)
    uint4 rec;        // lane i < nt: record of tile t0 + i
    bool heavy;       // a run out of k_front's list of heavy tiles (one tile)
};
struct Group {        // whole tiles of one run holding at most 32 pillars
    int t_a, nt;      // first tile, tiles (0 = none)
    int slotA, npil;  // first pillar (cell order, frames concatenated), pillars
    int rowA, nrows;  // first CSR row, rows
    unsigned tmask;   // lane i < nt: occupancy mask of tile t_a + i ...
    int tcol;         // ... and the block column of its first pillar
    bool skip;        // a listed heavy tile met inside the moving window: already taken care of, neither computed nor written
};

#ifndef HGSF_STAGE_ROWS
#define HGSF_STAGE_ROWS 96
#endif
#ifndef HGSF_PILLARS_MINB
#define HGSF_PILLARS_MINB 3
#endif
// CANVAS: also write spatial_features (p.canvas_vec: with 16-byte stores -- nx % 4 == 0 and an aligned canvas -- else scalar ones)
// GEN: the general build (spconv 1.x cut-off, more than 32 points per pillar); the shipped configs run the lean one
template <int F, bool ABS, bool DIST, bool BN, bool PFN, bool CANVAS, bool GEN>
__global__ void __launch_bounds__(PT, HGSF_PILLARS_MINB) k_pillars(const PathParams p) {
    constexpr int C = 64;
    constexpr int CIN = PFN ? ((ABS ? F : F - 3) + 6 + (DIST ? 1 : 0)) : 1;
    constexpr int RWc = (F + 1 + 3) / 4 * 4;   // F features + the point index, padded to float4
    constexpr int NV = RWc / 4;
    constexpr int SW = (RWc <= 8) ? HGSF_STAGE_ROWS : HGSF_STAGE_ROWS * 2 / 3;   // staged point rows per group (the rest is read through L1)
    constexpr int BLK = C * 32;
    const int Fr = PFN ? F : p.F, RW = PFN ? RWc : p.RW;      // without the PFN the kernel is generic in F (rows read from global)

    extern __shared__ __align__(16) uint8_t smem_raw[];
    float *blk_all = reinterpret_cast<float *>(smem_raw);                                   // [PW][BLK]          (CANVAS)
    float *stage_all = blk_all + (CANVAS ? PW * BLK : 0);                                   // [PW][2][SW * RWc]  (PFN)
    int *s_R = reinterpret_cast<int *>(stage_all + (PFN ? PW * 2 * SW * RWc : 0));          // [B+1] raw pillar base per frame
    int *s_K = s_R + (p.B + 1);                                                             // [B+1] kept (final) pillar base per frame
    __shared__ float4 s_rec_all[PW][32][2];                        // work lists: singles from the front, multis from the back
    __shared__ unsigned char s_perm_all[PW][32][32];               // per pillar: arrival position of its rank-th point
    __shared__ int s_bperm_all[PW][GEN ? MAX_P : 32];                         // same for a pillar with > 32 arrivals: its first min(cnt, P) points
    __shared__ int s_fcol_all[PW][32];                             // final pillar id of the chunk's pillar j (-1: never created)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float *blk = blk_all + warp * BLK;
    float *stage = stage_all + (size_t)warp * 2 * SW * RWc;
    float4(*rec)[2] = s_rec_all[warp];
    unsigned char(*perm)[32] = s_perm_all[warp];
    int *bperm = s_bperm_all[warp];
    int *fcol = s_fcol_all[warp];

    // ---- one-time setup (the only CTA barriers) ----
    if (CANVAS) for (int t = tid; t < PW * BLK; t += PT) blk_all[t] = 0.f;
    // launched as a programmatic dependent of k_front: everything below reads what that grid wrote
    asm volatile("griddepcontrol.wait;" ::: "memory");
    for (int b = tid; b <= p.B; b += PT) s_R[b] = p.frame_raw_base[b];
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
    const int half = lane >> 4;
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
    const unsigned lt = (1u << lane) - 1u;
    const uint64_t stream_policy = l2_policy_evict_first();   // pillar_features rows: written once, not re-read here

    // one point through decorate + Linear + BN, folded into the running max (integer max on the float bits: exact for
    // the non-negative post-ReLU values, drops negatives and -0 = the ReLU, lets a NaN 0x7fffffff win as torch.max does)
    auto eval_row = [&](const float (&row)[RWc], float mx, float my, float mz, float cx, float cy, float cz,
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
    auto row_ptr = [&](const float *stg, int rel, int pos, int rw) -> const float * {
        const float *base = (rel >= 0) ? stg + (size_t)rel * rw : grows + (size_t)(-1 - rel) * rw;
        return base + (size_t)pos * rw;
    };
    auto load_row = [&](const float *stg, int rel, int pos, float (&row)[RWc]) {
        const float4 *r4 = reinterpret_cast<const float4 *>(row_ptr(stg, rel, pos, RWc));
#pragma unroll
        for (int v = 0; v < NV; ++v) { const float4 t4 = r4[v]; row[4 * v] = t4.x; row[4 * v + 1] = t4.y; row[4 * v + 2] = t4.z; row[4 * v + 3] = t4.w; }
    };
    // channel c0+i of pillar column j sits at blk[(c0+i)*32 + (((j>>2) ^ ((c0+i)&7)) << 2 | (j&3))] (16-byte chunks XOR-swizzled by
    // the row so that the column writes spread over the banks); with c0 = 4*(lane&15): (c0+i)&7 = ((lane&1)<<2) ^ i
    float *const tbase = blk + c0 * 32;
    const int swb = (lane & 1) << 2;
    auto put_col = [&](int col, int v0, int v1, int v2, int v3) {
        const int xs = (col >> 2) ^ swb, xr = col & 3;
        tbase[0 * 32 + (((xs ^ 0) << 2) | xr)] = __int_as_float(v0);
        tbase[1 * 32 + (((xs ^ 1) << 2) | xr)] = __int_as_float(v1);
        tbase[2 * 32 + (((xs ^ 2) << 2) | xr)] = __int_as_float(v2);
        tbase[3 * 32 + (((xs ^ 3) << 2) | xr)] = __int_as_float(v3);
    };

    // ---- run hand-out: one global ticket per run, fetched a run before it is used.  Guided sizes: the first three quarters
    //      of the table go out in runs of CT tiles, the rest in runs of CT_TAIL.  [Measured and rejected: three tickets in flight
    //      per warp (the same-address atomics queue three times as long: +6 %); the first half of the runs assigned round-robin
    //      without tickets (the window of tiles in flight spreads: +16 %); 4 / 8 / 16 interleaved ticket counters (counter j
    //      hands out runs j, j + NQ, ...: the counters drift apart and the write window loses its order: +17 %, even on an almost
    //      empty scene, where the kernel is a pure zero-fill); one tile per group with a tile-layout buffer and one TMA tensor store
    //      per tile instead of the gather + 16 coalesced stores below (same time on VoD, +10 % on TJ4D and on sparse scenes: twice
    //      as many groups, and the group's fixed cost is what the sparse tiles pay).] ----
    //      Before all of that come the HEAVY tiles k_front listed (more than p.heavy_pts points), one tile per run: a single warp
    //      needs up to half the kernel's duration for the densest tile of a clustered scene, so it has to start at once; the moving
    //      window later steps over them.
    const int n_tt = (int)(((long long)p.B * p.cells) >> 5);       // tiles of the cell table
    const int n_heavy = (int)p.ticket[32];                          // the list holds one entry per tile at most: it cannot overflow
    const int n_big = (int)(((long long)n_tt * 3 / 4) / CT);       // runs of CT tiles
    const int t_tail = n_big * CT;                                 // first tile of the short runs
    const int n_runs = n_heavy + n_big + (n_tt - t_tail + CT_TAIL - 1) / CT_TAIL;
    unsigned tk_pending = 0u;            // lane 0: the ticket in flight
    bool runs_left = n_runs > 0;         // warp-uniform: a ticket below n_runs may still come
    auto fetch_ticket = [&]() { if (lane == 0) tk_pending = atomicAdd(p.ticket, 1u); };
    auto take_run = [&]() -> Run {
        Run r; r.t0 = 0; r.nt = 0; r.rec = make_uint4(0u, 0u, 0u, 0u); r.heavy = false;
        if (!runs_left) return r;
        int c = (int)__shfl_sync(FULL, tk_pending, 0);
        if (c >= n_runs) { runs_left = false; return r; }
        fetch_ticket();
        if (c < n_heavy) { r.t0 = (int)__ldg(p.heavy_list + c); r.nt = 1; r.heavy = true; }
        else {
            c -= n_heavy;
            if (c < n_big) { r.t0 = c * CT; r.nt = CT; }
            else { r.t0 = t_tail + (c - n_big) * CT_TAIL; r.nt = min(CT_TAIL, n_tt - r.t0); }
        }
        if (lane < r.nt) r.rec = __ldg(p.tile_rec + r.t0 + lane);
        return r;
    };
    Run run_cur, run_nxt;
    run_cur.t0 = run_cur.nt = 0; run_cur.rec = make_uint4(0u, 0u, 0u, 0u); run_cur.heavy = false; run_nxt = run_cur;
    int run_pos = 0;                     // next tile of run_cur that belongs to no group yet
    // the next group: tiles of the current run from run_pos on, as many whole tiles as hold at most 32 pillars (a tile never
    // holds more than 32, so at least one); moves on to the next run when this one is used up
    auto next_group = [&]() -> Group {
        Group g; g.t_a = 0; g.nt = 0; g.slotA = 0; g.npil = 0; g.rowA = 0; g.nrows = 0; g.tmask = 0u; g.tcol = 0; g.skip = false;
        if (run_pos >= run_cur.nt) {
            run_cur = run_nxt; run_nxt = take_run(); run_pos = 0;
            if (run_cur.nt == 0) return g;
        }
        const int a = run_pos;
        const unsigned slot_after = run_cur.rec.z + (unsigned)__popc(run_cur.rec.w);     // pillars before the NEXT tile
        const unsigned row_after = run_cur.rec.x + run_cur.rec.y;
        const int slotA = (int)__shfl_sync(FULL, run_cur.rec.z, a);
        // listed heavy tiles of a window run are stepped over (a group of their own that is neither computed nor written)
        unsigned hv = 0u;
        if (!run_cur.heavy && n_heavy > 0)
            hv = __ballot_sync(FULL, lane < run_cur.nt && run_cur.rec.y > (unsigned)p.heavy_pts) & (0xFFFFFFFFu << a);
        unsigned fits = __ballot_sync(FULL, lane >= a && lane < run_cur.nt && (int)slot_after - slotA <= 32);
        int nt;
        g.skip = false;
        if (hv & (1u << a)) { nt = 1; g.skip = true; }
        else {
            if (hv) fits &= (1u << (__ffs(hv) - 1)) - 1u;  // up to the first heavy tile
            nt = __popc(fits);                             // `fits` is a contiguous run of lanes starting at a
        }
        g.t_a = run_cur.t0 + a; g.nt = nt; g.slotA = slotA;
        g.npil = g.skip ? 0 : (int)__shfl_sync(FULL, slot_after, a + nt - 1) - slotA;
        g.rowA = (int)__shfl_sync(FULL, run_cur.rec.x, a);
        g.nrows = g.skip ? 0 : (int)__shfl_sync(FULL, row_after, a + nt - 1) - g.rowA;
        // lane i < nt keeps what the tile writer needs of tile a + i
        const int srcl = min(a + lane, 31);
        g.tmask = __shfl_sync(FULL, run_cur.rec.w, srcl);
        g.tcol = (int)__shfl_sync(FULL, run_cur.rec.z, srcl) - slotA;
        if (lane >= nt) { g.tmask = 0u; g.tcol = 0; }
        run_pos = a + nt;
        return g;
    };
    auto load_entry = [&](const Group &g) -> int4 {
        int4 e = make_int4(0, 0, 0, 0);
        if (lane < g.npil) e = __ldg(p.pil + g.slotA + lane);
        return e;
    };
    // the group's rows are sorted_rows[rowA, rowA + nrows): one cooperative async copy of (at most SW of) them
    auto issue_stage = [&](const Group &g, float *stg) {
        if (PFN) {
            const int pieces = min(g.nrows, SW) * NV;
            const float *src = grows + (size_t)g.rowA * RWc;
            for (int c = lane; c < pieces; c += 32) cp_async16(stg + 4 * c, src + 4 * c);
            cp_async_commit();
        }
    };

    // ---- the pillars of one group (all of one frame when there is a canvas; without one a group may straddle frames) ----
    auto process = [&](const Group &g, const int4 e, const float *stg) {
        const bool valid = lane < g.npil;
        const int key = e.x, cnt = e.z, start = e.w;
        const int b = (int)fastdiv((uint32_t)key, p.div_cells);
        const int local = e.y - s_R[b];
        const bool kept = valid && (local < maxv);            // pillars beyond max_voxels were never created
        const int f = s_K[b] + local;                          // final pillar id (first-seen order, frames concatenated)
        // spconv 1.x overflow: points from the frame's cut-off index on were never seen by the voxelizer
        const int cut = (GEN && (p.flags & HGSF_POINTS_SPCONV1_BREAK) && kept) ? p.cutoff[b] : INT_MAX;
        int cnt_eff = cnt;
        if (GEN && __any_sync(FULL, cut != INT_MAX)) {
            if (cut != INT_MAX && cnt <= 32) {
                cnt_eff = 0;
                for (int s2 = 0; s2 < cnt; ++s2)
                    cnt_eff += (__float_as_uint(__ldg(grows + (size_t)(start + s2) * RW + Fr)) < (uint32_t)cut) ? 1 : 0;
            }
        }
        const int n_keep = min(cnt_eff, Pmax);                 // (a pillar with more than 32 arrivals counts its own, below)
        // the pillar's cell: key = b*cells + (z*ny + y)*nxp + x
        const uint32_t rem = (uint32_t)(key - b * p.cells);
        const uint32_t pz = fastdiv(rem, p.div_plane), rem2 = rem - pz * (uint32_t)(p.ny * p.nxp);
        const uint32_t py = fastdiv(rem2, p.div_nxp), px = rem2 - py * (uint32_t)p.nxp;
        if (kept) {
            p.num[f] = n_keep;
            *reinterpret_cast<int4 *>(p.coords + 4 * (size_t)f) = make_int4(b, (int)pz, (int)py, (int)px);
        }
        // every pillar is visited exactly once: leave its table entry zero for the next call's k_front
        if (valid) { p.cell_cnt[key] = 0u; p.cell_tag[key] = 0u; }
        fcol[lane] = kept ? f : -1;
        const int rel0 = start - g.rowA;
        const bool staged = PFN && kept && cnt <= 32 && rel0 + cnt <= SW;
        const int rel = staged ? rel0 : (-1 - start);         // where the pillar's rows are (see row_ptr)
        if (PFN) cp_async_wait<1>();     // this chunk's rows have landed (this lane's copies) ...
        __syncwarp();                    // ... and every other lane's
#ifdef HGSF_EXPERIMENT
        if (p.dbg & 1) return;           // timing only (WRONG results): no ordering / arithmetic
#endif
        // ---- order the pillar's points by input index ----
        if (kept && cnt > 1 && cnt <= SMALL_CNT) {
            uint32_t idx[SMALL_CNT];
#pragma unroll
            for (int j = 0; j < SMALL_CNT; ++j) {
                idx[j] = 0xFFFFFFFFu;
                if (j < cnt) idx[j] = __float_as_uint(row_ptr(stg, rel, j, RW)[Fr]);
            }
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
            const int cnt_o = __shfl_sync(FULL, cnt, o), rel_o = __shfl_sync(FULL, rel, o);
            uint32_t mine = 0xFFFFFFFFu;
            if (lane < cnt_o) mine = __float_as_uint(row_ptr(stg, rel_o, lane, RW)[Fr]);
            int rank = 0;
            for (int q = 0; q < cnt_o; ++q) rank += (__shfl_sync(FULL, mine, q) < mine) ? 1 : 0;
            if (lane < cnt_o) perm[o][rank] = (unsigned char)lane;
        }
        __syncwarp();
        const bool live = kept && cnt <= 32;
        const unsigned huge = __ballot_sync(FULL, kept && cnt > 32);
        // pillar centre: fl(fl(c*v)+off), two roundings, no FMA (pillar_vfe.py:101-103)
        const float cx = __fadd_rn(__fmul_rn((float)px, vsx), vox);
        const float cy = __fadd_rn(__fmul_rn((float)py, vsy), voy);
        const float cz = __fadd_rn(__fmul_rn((float)pz, vsz), voz);
        if (PFN) {
            // mean of the kept points (torch CPU sum order) and the record the unit lanes read
            float mx = 0.f, my = 0.f, mz = 0.f;
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
            // two work lists: pillars with ONE point to evaluate (paired across the half-warps) and pillars with several
            // (taken one at a time, the half-warps splitting the points)
            const unsigned sbal = __ballot_sync(FULL, live && n_keep == 1);
            const unsigned mbal = __ballot_sync(FULL, live && n_keep > 1);
            const int n_s = __popc(sbal), n_m = __popc(mbal);
            if (live) {
                const int slot = (n_keep == 1) ? __popc(sbal & lt) : 31 - __popc(mbal & lt);   // singles from the front, multis from the back
                // position of the single evaluated point: 0, or the rank-0 arrival when P == 1 truncated a larger pillar
                const int pos0 = (cnt == 1) ? 0 : (int)perm[lane][0];
                // bit 24: the pillar was truncated to its first P points, so the unit lanes must go through the rank table; an
                // untruncated pillar is evaluated in arrival order (the max does not care), one dependent shared-memory load less
                rec[slot][0] = make_float4(mx, my, mz, __int_as_float(n_keep | (lane << 8) | (pos0 << 16) | ((cnt > n_keep) ? (1 << 24) : 0)));
                rec[slot][1] = make_float4(cx, cy, cz, __int_as_float(rel));
            }
            // the feature block: columns of pillars that were never created (beyond max_voxels) read as zero on the canvas
            if (CANVAS) {
                unsigned dropped = __ballot_sync(FULL, valid && !kept);
                while (dropped) {
                    const int col = __ffs(dropped) - 1;
                    dropped &= dropped - 1;
                    blk[swz128(lane, col)] = 0.f; blk[swz128(lane + 32, col)] = 0.f;
                }
            }
            __syncwarp();
            // ---- unit phase.  lane l always computes channels c0..c0+3 ----
            // (a) single-point pillars: half-warp h takes list entries 4*j + h and 4*j + 2 + h -- two independent points per
            //     iteration, so that their FMA chains interleave
#pragma unroll 1
            for (int j = 0; 4 * j < n_s; ++j) {
                const int eA = 4 * j + half;
                if (eA < n_s) {
                    const bool okB = eA + 2 < n_s;
                    const int eB = okB ? eA + 2 : eA;
                    const float4 rA0 = rec[eA][0], rA1 = rec[eA][1], rB0 = rec[eB][0], rB1 = rec[eB][1];
                    const int metaA = __float_as_int(rA0.w), metaB = __float_as_int(rB0.w);
                    const int colA = (metaA >> 8) & 0xFF, colB = (metaB >> 8) & 0xFF;
                    int a0 = 0, a1 = 0, a2 = 0, a3 = 0;
                    if (1 < Pmax) { a0 = __float_as_int(pv.x); a1 = __float_as_int(pv.y); a2 = __float_as_int(pv.z); a3 = __float_as_int(pv.w); }
                    int b0 = a0, b1 = a1, b2 = a2, b3 = a3;
                    float rowA[RWc], rowB[RWc];
                    load_row(stg, __float_as_int(rA1.w), (metaA >> 16) & 0xFF, rowA);
                    load_row(stg, __float_as_int(rB1.w), (metaB >> 16) & 0xFF, rowB);
                    eval_row(rowA, rA0.x, rA0.y, rA0.z, rA1.x, rA1.y, rA1.z, a0, a1, a2, a3);
                    eval_row(rowB, rB0.x, rB0.y, rB0.z, rB1.x, rB1.y, rB1.z, b0, b1, b2, b3);
                    if (p.feats) {
                        st_f4_hint(p.feats + (size_t)fcol[colA] * C + c0,
                                   make_float4(__int_as_float(a0), __int_as_float(a1), __int_as_float(a2), __int_as_float(a3)), stream_policy);
                        if (okB)
                            st_f4_hint(p.feats + (size_t)fcol[colB] * C + c0,
                                       make_float4(__int_as_float(b0), __int_as_float(b1), __int_as_float(b2), __int_as_float(b3)), stream_policy);
                    }
                    if (CANVAS) { put_col(colA, a0, a1, a2, a3); if (okB) put_col(colB, b0, b1, b2, b3); }
                }
            }
            // (b) multi-point pillars: both half-warps on the same pillar, half h takes slots h, h+2, ... (two per iteration);
            //     max-combined
#pragma unroll 1
            for (int j = 0; j < n_m; ++j) {
                const float4 r0 = rec[31 - j][0], r1 = rec[31 - j][1];
                const int meta = __float_as_int(r0.w);
                const int nk = meta & 0xFF, col = (meta >> 8) & 0xFF;
                const bool trunc = (meta >> 24) & 1;
                const int relp = __float_as_int(r1.w);
                int v0 = 0, v1 = 0, v2 = 0, v3 = 0;
                if (nk < Pmax) { v0 = __float_as_int(pv.x); v1 = __float_as_int(pv.y); v2 = __float_as_int(pv.z); v3 = __float_as_int(pv.w); }
                int u0 = v0, u1 = v1, u2 = v2, u3 = v3;
#pragma unroll 1
                for (int s2 = half; s2 < nk; s2 += 4) {
                    const int s3 = (s2 + 2 < nk) ? s2 + 2 : s2;        // the last odd one is evaluated twice: max is idempotent
                    float rowA[RWc], rowB[RWc];
                    load_row(stg, relp, trunc ? (int)perm[col][s2] : s2, rowA);
                    load_row(stg, relp, trunc ? (int)perm[col][s3] : s3, rowB);
                    eval_row(rowA, r0.x, r0.y, r0.z, r1.x, r1.y, r1.z, v0, v1, v2, v3);
                    eval_row(rowB, r0.x, r0.y, r0.z, r1.x, r1.y, r1.z, u0, u1, u2, u3);
                }
                v0 = max(v0, u0); v1 = max(v1, u1); v2 = max(v2, u2); v3 = max(v3, u3);
                v0 = max(v0, __shfl_xor_sync(FULL, v0, 16)); v1 = max(v1, __shfl_xor_sync(FULL, v1, 16));
                v2 = max(v2, __shfl_xor_sync(FULL, v2, 16)); v3 = max(v3, __shfl_xor_sync(FULL, v3, 16));
                if (half == 0) {
                    if (p.feats)
                        st_f4_hint(p.feats + (size_t)fcol[col] * C + c0,
                                   make_float4(__int_as_float(v0), __int_as_float(v1), __int_as_float(v2), __int_as_float(v3)), stream_policy);
                    if (CANVAS) put_col(col, v0, v1, v2, v3);
                }
            }
        }
        // ---- pillars with more than 32 arrivals: the warp selects the 32 smallest point indices, then as (b) ----
        unsigned hm = huge;
        while (hm) {
            const int o = __ffs(hm) - 1;
            hm &= hm - 1;
            const int cnt_o = __shfl_sync(FULL, cnt, o), start_o = __shfl_sync(FULL, start, o), f_o = __shfl_sync(FULL, f, o);
            const float cx_o = __shfl_sync(FULL, cx, o), cy_o = __shfl_sync(FULL, cy, o), cz_o = __shfl_sync(FULL, cz, o);
            const float *grow_o = grows + (size_t)start_o * RW;
            int cnt_e = cnt_o;
            const int cut_o = GEN ? __shfl_sync(FULL, cut, o) : INT_MAX;
            if (GEN && cut_o != INT_MAX) {
                cnt_e = 0;
                for (int j0 = 0; j0 < cnt_o; j0 += 32) {
                    const int j = j0 + lane;
                    cnt_e += __popc(__ballot_sync(FULL, j < cnt_o && __float_as_uint(__ldg(grow_o + (size_t)j * RW + Fr)) < (uint32_t)cut_o));
                }
            }
            const int nk = min(cnt_e, Pmax);
            if (GEN && lane == 0) p.num[f_o] = nk;
            {
                // the pillar's first nk points by index, 32 per round (the lean build has P <= 32: one round)
                uint32_t lo = 0u, key = 0u;
                if (GEN) {
                    for (int r0 = 0; r0 < nk; r0 += 32) {
                        const int pos = select_next32(grow_o + Fr, RW, cnt_o, lane, r0 > 0, lo, key);
                        if (r0 + lane < MAX_P) bperm[r0 + lane] = pos;
                        lo = __shfl_sync(FULL, key, 31);
                    }
                } else {
                    bperm[lane] = select_next32(grow_o + Fr, RW, cnt_o, lane, false, lo, key);
                }
            }
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
                const float hx = __fdiv_rn(sum.sx(), fn), hy = __fdiv_rn(sum.sy(), fn), hz = __fdiv_rn(sum.sz(), fn);
                int v0 = 0, v1 = 0, v2 = 0, v3 = 0;
                if (nk < Pmax) { v0 = __float_as_int(pv.x); v1 = __float_as_int(pv.y); v2 = __float_as_int(pv.z); v3 = __float_as_int(pv.w); }
#pragma unroll 1
                for (int s2 = half; s2 < nk; s2 += 2) {
                    float row[RWc];
                    load_row(stg, -1 - start_o, bperm[s2], row);
                    eval_row(row, hx, hy, hz, cx_o, cy_o, cz_o, v0, v1, v2, v3);
                }
                v0 = max(v0, __shfl_xor_sync(FULL, v0, 16)); v1 = max(v1, __shfl_xor_sync(FULL, v1, 16));
                v2 = max(v2, __shfl_xor_sync(FULL, v2, 16)); v3 = max(v3, __shfl_xor_sync(FULL, v3, 16));
                if (half == 0) {
                    if (p.feats)
                        st_f4_hint(p.feats + (size_t)f_o * C + c0,
                                   make_float4(__int_as_float(v0), __int_as_float(v1), __int_as_float(v2), __int_as_float(v3)), stream_policy);
                    if (CANVAS) put_col(o, v0, v1, v2, v3);
                }
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
                const int nk = __shfl_sync(FULL, n_keep, o);
                float *vo = p.voxels + (size_t)f_o * Pmax * Fr;
                for (int t = lane; t < Pmax * Fr; t += 32) {
                    const int s2 = t / Fr, kk = t - s2 * Fr;
                    float v = 0.f;
                    if (s2 < nk) v = row_ptr(stg, rel_o, (cnt_o == 1) ? 0 : (int)perm[o][s2], RW)[kk];
                    vo[t] = v;
                }
            }
        }
        __syncwarp();   // the feature block is complete; rec / perm / fcol are rewritten by the next group
    };

    // ---- the group's canvas tiles, out of the feature block.  lane -> cells 4*(lane&7)..+3 of the tile and channel rows
    //      (lane>>3) + 4*q ----
    const int tiles_per_row = p.tiles_per_row;
    auto write_tiles = [&](const Group &g) {
        const int quad = lane & 7, cs = lane >> 3;
        int r = (int)fastdiv((uint32_t)g.t_a, p.div_tpr), xt = g.t_a - r * tiles_per_row;
        int b = (int)fastdiv((uint32_t)r, p.div_ny), y = r - b * p.ny;
        const size_t plane4 = (size_t)4 * p.ny * p.nx;
        for (int i = 0; i < g.nt; ++i) {
            const unsigned mask = __shfl_sync(FULL, g.tmask, i);
            const int col0 = __shfl_sync(FULL, g.tcol, i);
            const int x = xt * 32 + 4 * quad;
            const unsigned bits = (mask >> (4 * quad)) & 0xFu;
            if (p.canvas_vec) {
                float *dst = p.canvas + (((size_t)b * C + cs) * p.ny + y) * p.nx + x;
                if (mask == 0u) {
                    if (x < p.nx) {
#pragma unroll
                        for (int q = 0; q < C / 4; ++q) __stcs(reinterpret_cast<float4 *>(dst + q * plane4), make_float4(0.f, 0.f, 0.f, 0.f));
                    }
                } else {
                    // block columns of this lane's (up to 4) pillars.  Channel row c = cs + 4q of column j sits at
                    // blk[c*32 + ((((j>>2) ^ (c&7)) << 2) | (j&3))] and c&7 = cs + 4*(q&1): two base pointers per column (even / odd q),
                    // then a compile-time offset q*128 per row.  Cell positions no quad of the tile occupies are skipped warp-wide.
                    const int cc0 = col0 + __popc(mask & ((1u << (4 * quad)) - 1u));
                    const float *rowbase = blk + cs * 32;
                    float4 v[C / 4];
#pragma unroll
                    for (int q = 0; q < C / 4; ++q) v[q] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (mask & (0x11111111u << k)) {                     // warp-uniform
                            if ((bits >> k) & 1u) {
                                const int j = cc0 + __popc(bits & ((1u << k) - 1u));
                                const float *pe = rowbase + ((((j >> 2) ^ cs) << 2) | (j & 3));
                                const float *po = rowbase + ((((j >> 2) ^ (cs + 4)) << 2) | (j & 3));
#pragma unroll
                                for (int q = 0; q < C / 4; ++q) {
                                    const float t = (q & 1) ? po[q * 128] : pe[q * 128];
                                    if (k == 0) v[q].x = t; else if (k == 1) v[q].y = t; else if (k == 2) v[q].z = t; else v[q].w = t;
                                }
                            }
                        }
                    }
                    if (x < p.nx) {
#pragma unroll
                        for (int q = 0; q < C / 4; ++q) __stcs(reinterpret_cast<float4 *>(dst + q * plane4), v[q]);
                    }
                }
            } else {
                // nx not a multiple of 4 (or a misaligned canvas): scalar stores, same mapping
                int col[4];
                {
                    int cc = col0 + __popc(mask & ((1u << (4 * quad)) - 1u));
#pragma unroll
                    for (int k = 0; k < 4; ++k) { col[k] = cc; cc += (bits >> k) & 1u; }
                }
                for (int q = 0; q < C / 4; ++q) {
                    const int c = cs + 4 * q;
                    float *dst = p.canvas + (((size_t)b * C + c) * p.ny + y) * p.nx;
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (x + k < p.nx) dst[x + k] = ((bits >> k) & 1u) ? blk[swz128(c, col[k])] : 0.f;
                }
            }
            if (++xt == tiles_per_row) { xt = 0; if (++y == p.ny) { y = 0; ++b; } }
        }
    };

    // ---- the work loop: the ticket and the records of the next run and the list entries + rows of the next group are in flight
    //      while the current group is computed and written ----
    fetch_ticket();
    run_cur = take_run(); run_nxt = take_run();
    Group g_cur = next_group(), g_nxt;
    int4 e_cur = load_entry(g_cur), e_nxt;
    issue_stage(g_cur, stage);
    for (int it = 0; g_cur.nt > 0; ++it) {
        g_nxt = next_group();
        e_nxt = load_entry(g_nxt);
        issue_stage(g_nxt, stage + (size_t)((it + 1) & 1) * SW * RWc);
        if (g_cur.npil > 0) process(g_cur, e_cur, stage + (size_t)(it & 1) * SW * RWc);
#ifdef HGSF_EXPERIMENT
        if (CANVAS && !g_cur.skip && !(p.dbg & 2)) write_tiles(g_cur);
#else
        if (CANVAS && !g_cur.skip) write_tiles(g_cur);
#endif
        __syncwarp();   // the feature block is rewritten by the next group
        g_cur = g_nxt; e_cur = e_nxt;
    }
    if (PFN) cp_async_wait<0>();
    mark_table_clean(p);
}

// ---- k_emit -------------------------------------------------------------------------------------
// The tile-major fused consumer (canvas requested, P <= 32, spconv-2 overflow rule, nx a multiple of 4): order + decorate +
// PFN + max AND the canvas tile, one pass.  Every WARP is an autonomous worker walking its own sequence of canvas tiles (a
// tile = 32 cells of one BEV row x 64 channels = 64 rows of 128 B); no CTA barrier in the loop.  Because k_front laid the
// CSR out in cell order, the point rows of a tile are ONE contiguous span of sorted_rows: they are staged with a single
// cooperative cp.async copy issued a tile ahead; the tile's record comes two tiles ahead, its table entries one -- no
// dependent gathers anywhere, and empty cells / empty tiles are never looked up in the table.
//   lane l OWNS cell l of the tile for the bookkeeping (ordering by point index, first P kept, mean in torch's
//   summation order, voxel_coords / voxel_num_points) and zeroes the table entry it read (k_front finds the table clean);
//   the arithmetic is cut into UNITS of (pillar, 4 output channels): lane l always computes channels 4*(l&15)..+3, so
//   its Linear weights (packed pairs, one FFMA2 per channel pair and input feature) and BatchNorm constants stay in
//   REGISTERS for the whole kernel (CUDA-core FMA: a 13x64 contraction is far below a tensor-core tile).  Single-point
//   pillars are paired across the half-warps; a multi-point pillar is taken by both halves, which split its points and
//   max-combine.
//   A finished tile leaves in ONE TMA tensor store (128-byte swizzle so the column writes spread over banks); an
//   empty tile is four stores of a shared 2 KB zero tile.  The canvas is written exactly once, zeros included.
// Against the pillar-major k_pillars (groups of <= 32 pillars, st.cs tile writes) on one B200, ms per step:
//   VoD clustered 0.178 / 0.190, VoD uniform 0.175 / 0.193, TJ4D clustered 0.315 / 0.320, stress 0.771 / 0.783.
// Measured on this kernel and NOT kept (VoD clustered, ms in k_emit, baseline 0.129):
//   tickets 2 ahead + records / entries 3 tiles ahead through a cp.async ring   0.144  (stalls gone, window wider: slower)
//   one ticket per CTA = 4 / 8 / 16 adjacent tiles popped from a shared queue   0.163 / 0.162 / 0.168
//   listed (heavy) tiles spread over 85 % / 100 % of the ticket sequence        0.141 / 0.146 (all first: 0.129)
//   16 / 15 / 14 warps per SM at 128 registers (4x4, 3x5, 2x7 warps; no spills) 0.141 / 0.137 / 0.136
//   listing threshold 40 / 28 / 20 / 12 points                                  0.132 / 0.133 / 0.133 / 0.134
//   two tiles per ticket half the table apart instead of adjacent               0.1323 vs 0.1318 (TJ4D 0.255 vs 0.247)
//   two tickets in flight, record + (unconditional) entries three tiles ahead, in registers (163, no spills)   0.144
//   last iteration's loads consumed before this iteration's instances are issued (address tied to them)       0.131
//   2 / 4 / 8 ticket counters on separate lines, interleaved sequence numbers    0.130 / 0.128 / 0.129 (TJ4D 0.249 vs 0.247):
//                                                  the same-address atomic is not what the warps wait for
// The statistics pass of the train mode (no Linear, no stores, no outputs) takes 115 us of the 143 us the full kernel takes
// under ncu: the walk itself -- ticket, record, entries, staging, ordering, means -- is the cost, not arithmetic or DRAM.
// Warm-cache L2 counters (ncu --cache-control none, one pass): 59 % of this kernel's read sectors miss the L2 and come from DRAM
// (20 MB per launch) although k_front wrote them just before -- the 432 MB write stream pushes them out.  Pinning the workspace
// with a persisting access-policy window (whole carve-out / 32 MB) halves the misses and makes the kernel 0.248 / 0.164 ms: the
// write stream needs the L2 capacity more than the reads need the hits.
// prefetch.global.L2 of the table lines / rows of the tile 2 048 / 4 096 / 8 192 positions down the window: the demand misses stay
// (562 k sectors), 0.135 / 0.134 / 0.134 ms.
// Entries out of k_front's pillar list (one coalesced 16-byte-per-pillar load + three shuffles to cell order) instead of the table
// arrays: no change on VoD (0.1284 vs 0.1286), 1 % slower on TJ4D / stress.
// Per-tile clocks (-DHGSF_TILE_CLOCKS, profiles/r02_tile_clocks_k_emit.txt): every tile, whatever its class, spends ~2 000 cycles
// waiting for the entries / record fetched one (mostly short) iteration earlier and for its ticket: L2 round trips take 2-4 k
// cycles under this kernel's own load, and fetching further ahead raised them further.
#ifndef HGSF_EMIT_WARPS
#define HGSF_EMIT_WARPS 4
#endif
#ifndef HGSF_EMIT_STAGE
#define HGSF_EMIT_STAGE 64
#endif
constexpr int EMIT_WARPS = HGSF_EMIT_WARPS;
constexpr int EMIT_THREADS = EMIT_WARPS * 32;
constexpr int STAGE_W = HGSF_EMIT_STAGE;           // staged point rows per tile (a typical tile holds ~10; the rest is read from L2)

#ifndef HGSF_EMIT_MINB
#define HGSF_EMIT_MINB 3
#endif
// STATS: the train-mode statistics pass (BatchNorm1d on batch statistics, pillar_vfe.py:29-42): the same walk over the same
// tiles with the same decoration, but instead of Linear + BN + max it accumulates the SECOND MOMENTS of the decorated features
// over every kept point, S = sum f f^T [Cin, Cin] and s = sum f [Cin] in fp64 (lane r of a half-warp owns row r of S, the lane
// after the last row owns s).  Everything BatchNorm's forward and backward need follows from them, because x = W f is linear:
// sum x_c = w_c . s,  sum x_c^2 = w_c^T S w_c,  T[c,k] = sum x_c f_k = (W S)[c,k].  Nothing is written but p.stats_S, the table
// is left as it is for the real pass that follows, and the last CTA expands S, s into the train_ops.cu statistics layout
// (Sx, Sxx, T, s: the backward reads it as it is), batch_mean / batch_var, and updates the running statistics.
template <int F, bool ABS, bool DIST, bool BN, int CHUNK, bool STATS>
__global__ void __launch_bounds__(EMIT_THREADS, HGSF_EMIT_MINB)
k_emit(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap zmap, const PathParams p) {
    constexpr int C = 64;
    constexpr int CIN = (ABS ? F : F - 3) + 6 + (DIST ? 1 : 0);
    static_assert(CIN <= 15, "STATS: rows of S on lanes 0..CIN-1 of a half-warp, s on lane CIN");
    constexpr int RWc = (F + 1 + 3) / 4 * 4;   // F features + the point index, padded to float4
    constexpr int NV = RWc / 4;
    constexpr int TILE = C * 32;
    constexpr int ZC = C / 4;
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
    for (int t = tid; t < ZC * 32; t += NT) zerobuf[t] = 0.f;
    for (int t = tid; t < EMIT_WARPS * TILE; t += NT) tiles[t] = 0.f;
    // launched as a programmatic dependent of k_front: nothing k_front wrote may be read before this returns
    asm volatile("griddepcontrol.wait;" ::: "memory");
    for (int b = tid; b <= p.B; b += NT) s_R[b] = p.frame_raw_base[b];
    __syncthreads();
    if (tid == 0) {
        int acc = 0;
        for (int b = 0; b < p.B; ++b) {
            s_K[b] = acc;
            const int m = min(s_R[b + 1] - s_R[b], p.max_voxels);
            if (!STATS && blockIdx.x == 0) p.num_pillars[1 + b] = m;
            acc += m;
        }
        s_K[p.B] = acc;
        if (!STATS && blockIdx.x == 0) p.num_pillars[0] = acc;
    }
    fence_proxy_async_smem();
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
    float *const feats_out = STATS ? nullptr : p.feats, *const voxels_out = STATS ? nullptr : p.voxels;
    double st_S[STATS ? CIN : 1];                    // STATS: row (lane & 15) of S = sum f f^T; the lane after the last row: s = sum f
#pragma unroll
    for (int k = 0; k < (STATS ? CIN : 1); ++k) st_S[k] = 0.0;
    const int st_role = lane & 15;
    auto eval_row = [&](const float (&row)[RWc], float mx, float my, float mz, float cx, float cy,
                        int &v0, int &v1, int &v2, int &v3, bool counted = true) {
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
        if (STATS) {                           // a row evaluated twice to fill a pair (below) counts once
            if (counted) {
                float fr = 1.f;                // the s lane multiplies by one
#pragma unroll
                for (int kk = 0; kk < CIN; ++kk) if (kk == st_role) fr = feat[kk];
                const double dr = (double)fr;
#pragma unroll
                for (int kk = 0; kk < CIN; ++kk) st_S[kk] = fma(dr, (double)feat[kk], st_S[kk]);
            }
            return;
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
        if (STATS) return;
        const int xs = (cell >> 2) ^ swb, xr = cell & 3;
        tbase[0 * 32 + (((xs ^ 0) << 2) | xr)] = __int_as_float(v0);
        tbase[1 * 32 + (((xs ^ 1) << 2) | xr)] = __int_as_float(v1);
        tbase[2 * 32 + (((xs ^ 2) << 2) | xr)] = __int_as_float(v2);
        tbase[3 * 32 + (((xs ^ 3) << 2) | xr)] = __int_as_float(v3);
    };

    // Tiles are handed out DYNAMICALLY, one global ticket per CHUNK consecutive tiles, fetched a chunk ahead: a dense tile costs
    // ten times a sparse one, and a static assignment leaves the unlucky warps running alone at the end.  Short chunks also keep
    // x-adjacent tiles -- adjacent 128-byte pieces of the same canvas rows -- in flight at the same time on different warps,
    // which the DRAM write stream rewards.  The first tickets are k_front's list of HEAVY tiles (more than p.heavy_pts points:
    // one warp is busy with such a tile for a long time, so they start first and the many light tiles fill in around them);
    // the moving window then steps over the listed tiles.
    const int n_tiles = (int)(((long long)p.B * p.cells) >> 5);
    const int n_heavy = (int)p.ticket[32];
    const unsigned heavy_pts = (unsigned)p.heavy_pts;
    constexpr int chunk = CHUNK;
    const int n_tickets = n_heavy + (n_tiles + chunk - 1) / chunk;
    // the ticket stays in lane 0's register until the chunk is actually started: broadcasting it right away would
    // stall the whole warp on the atomic's round trip
    auto fetch_raw = [&]() -> int {
        int v = 0;
        if (lane == 0) v = (int)atomicAdd(p.ticket + (STATS ? 16 : 0), 1u);
        return v;
    };
    struct Slot { int t; bool listed; };             // t == n_tiles: past the end
    Slot seq; seq.t = 0; seq.listed = false;
    int seq_left = 0, next_raw = 0;
    auto start_chunk = [&](int c) {
        if (c < n_heavy) { seq.t = (int)__ldg(p.heavy_list + c); seq_left = 1; seq.listed = true; }
        else if (c < n_tickets) { seq.t = (c - n_heavy) * chunk; seq_left = min(chunk, n_tiles - seq.t); seq.listed = false; }
        else { seq.t = n_tiles; seq_left = 0; seq.listed = false; }
    };
    start_chunk(__shfl_sync(FULL, fetch_raw(), 0));
    if (seq.t < n_tiles) next_raw = fetch_raw();
    auto next_tile = [&]() -> Slot {
        if (seq.t >= n_tiles) return seq;
        if (seq_left > 1) { --seq_left; ++seq.t; }
        else {
            start_chunk(__shfl_sync(FULL, next_raw, 0));
            if (seq.t < n_tiles) next_raw = fetch_raw();
        }
        return seq;
    };
    Slot cur = seq;
    Slot nxt = next_tile();
    Slot nxt2 = next_tile();
    // k_front's record of a tile: {first CSR row, rows, pillars before it, occupancy mask}
    auto load_rec = [&](const Slot &s) -> uint4 {
        return (s.t < n_tiles) ? __ldg(p.tile_rec + s.t) : make_uint4(0u, 0u, 0u, 0u);
    };
    // a listed tile met inside the window was handed out at the start: neither computed nor written here
    auto skipped = [&](const Slot &s, const uint4 &r) -> bool { return !s.listed && r.y > heavy_pts; };
    // lane l: table entry of cell 32*t + l (the table rows are padded to whole tiles); empty cells are not read
    auto load_entry = [&](const Slot &s, const uint4 &r) -> uint4 {
        uint4 e = make_uint4(0, 0, 0, 0);
        if (((r.w >> lane) & 1u) && !skipped(s, r)) {
            const size_t c = (size_t)s.t * 32 + lane;
            e.x = __ldg(p.cell_tag + c); e.y = __ldg(p.cell_cnt + c);     // (cell_start is not read: see `start` below)
        }
        return e;
    };
    // the tile's rows are sorted_rows[row0, row0 + total): one cooperative async copy of (at most STAGE_W of) them
    auto issue_stage = [&](const Slot &s, const uint4 &r, float *stg) {
        const int chunks = skipped(s, r) ? 0 : min((int)r.y, STAGE_W) * NV;
        const float *src = grows + (size_t)r.x * RWc;
        for (int c = lane; c < chunks; c += 32) cp_async16(stg + 4 * c, src + 4 * c);
        cp_async_commit();
    };

    uint4 r_cur = load_rec(cur);
    uint4 r_nxt = load_rec(nxt);
    uint4 e_cur = load_entry(cur, r_cur);
    issue_stage(cur, r_cur, stage);
    unsigned dirty = 0;                  // cells of the tile buffer that hold non-zero columns
    bool store_pending = false;          // a TMA store from the tile buffer may still be reading it

#ifdef HGSF_TILE_CLOCKS
    // measurement build (scripts/tile_clocks.py): cycles per tile by class {empty, 1..10 points, 11.., listed heavy, skipped} x
    // {front of the iteration, the tile itself, the hand-out at the end}, summed per warp, then into scan_desc's tail
    unsigned long long tc_cyc[5][3] = {}, tc_cnt[5] = {};
    const long long tc_begin = clock64();
#endif
    for (int it = 0; cur.t < n_tiles; ++it) {
#ifdef HGSF_TILE_CLOCKS
        const long long tc0 = clock64();
#endif
        const float *stg = stage + (size_t)(it & 1) * STAGE_W * RWc;
        // ---- pipeline: record of the tile after next, entries and rows of the next tile ----
        const uint4 r_nn = load_rec(nxt2);
        const uint4 e_nxt = load_entry(nxt, r_nxt);
        issue_stage(nxt, r_nxt, stage + (size_t)((it + 1) & 1) * STAGE_W * RWc);
#ifdef HGSF_TILE_CLOCKS
        const int tc_class = skipped(cur, r_cur) ? 4 : (cur.listed ? 3 : (r_cur.y == 0 ? 0 : (r_cur.y <= 10 ? 1 : 2)));
        const long long tc1 = clock64();
#endif

        // tile t = (frame b, BEV row y, 32 cells from x0)
        const uint32_t row_id = fastdiv((uint32_t)cur.t, p.div_tpr);
        const int b = (int)fastdiv(row_id, p.div_ny);
        const int y = (int)row_id - b * p.ny, x0 = (cur.t - (int)row_id * p.tiles_per_row) * 32;
        const bool skip = skipped(cur, r_cur);
        if (!STATS && e_cur.y) {         // the entry has been read: leave the table clean for the next call's k_front
            const size_t c = (size_t)cur.t * 32 + lane;
            p.cell_tag[c] = 0u; p.cell_cnt[c] = 0u;
        }
        // the CSR is in cell order and the tile's rows are one span from its record's first row: a cell's start is that row plus the
        // counts of the cells before it (one warp scan instead of a third table load per tile)
        const int m = (int)(e_cur.x - 1u), cnt = (int)e_cur.y;
        int start;
        {
            int incl = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int o = __shfl_up_sync(FULL, incl, d);
                if (lane >= d) incl += o;
            }
            start = (int)r_cur.x + incl - cnt;
        }
        const int local = m - s_R[b];
        const bool occ = (e_cur.x != 0u) && (local < maxv);     // pillars beyond max_voxels were never created
        const unsigned bal_occ = __ballot_sync(FULL, occ);
        if (skip) {
        } else if (bal_occ == 0u) {
            if (STATS) {
            } else
            // empty tile: four stores of the shared zero tile
            // issued by lane 1: bulk async-groups are per thread, so lane 0's wait for its tile store to have read the
            // tile buffer (below) does not also wait for the zero stores of the empty tiles that came after it
            if (lane == 1) {
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) tma_store_3d_hint(&zmap, zerobuf, x0, y, b * C + q4 * ZC, stream_policy);
                tma_commit();
            }
        } else {
            const int row0 = (int)r_cur.x;
            const int rel0 = start - row0;
            const bool staged = occ && cnt <= 32 && rel0 + cnt <= STAGE_W;
            const int rel = staged ? rel0 : (-1 - start);         // where the pillar's rows are (see load_row)
            const int n_keep = min(cnt, Pmax);
            const int f = s_K[b] + local;                          // final pillar id (first-seen order, frames concatenated)
            if (!STATS && occ) {
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
            if (!STATS && store_pending) {
                if (lane == 0) tma_wait_read<0>();
                store_pending = false;
            }
            __syncwarp();
            if (STATS) {
            } else if (__popc(dirty) > 2) {
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
                    eval_row(rowB, rB0.x, rB0.y, rB0.z, cxB, cy, b0, b1, b2, b3, okB);
                    if (feats_out) {
                        st_f4_hint(feats_out + (size_t)__float_as_int(rA1.y) * C + c0,
                                   make_float4(__int_as_float(a0), __int_as_float(a1), __int_as_float(a2), __int_as_float(a3)), feats_policy);
                        if (okB)
                            st_f4_hint(feats_out + (size_t)__float_as_int(rB1.y) * C + c0,
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
                    eval_row(rowB, r0.x, r0.y, r0.z, cx, cy, u0, u1, u2, u3, s2 + 2 < nk);
                }
                v0 = max(v0, u0); v1 = max(v1, u1); v2 = max(v2, u2); v3 = max(v3, u3);
                v0 = max(v0, __shfl_xor_sync(FULL, v0, 16)); v1 = max(v1, __shfl_xor_sync(FULL, v1, 16));
                v2 = max(v2, __shfl_xor_sync(FULL, v2, 16)); v3 = max(v3, __shfl_xor_sync(FULL, v3, 16));
                if (half == 0) {
                    if (feats_out)
                        st_f4_hint(feats_out + (size_t)__float_as_int(r1.y) * C + c0,
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
                { uint32_t key_unused; bperm[lane] = select_next32(grow_o + F, RWc, cnt_o, lane, false, 0u, key_unused); }
                __syncwarp();
                if (voxels_out) {
                    float *vo = voxels_out + (size_t)f_o * Pmax * F;
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
                    if (feats_out)
                        st_f4_hint(feats_out + (size_t)f_o * C + c0,
                                   make_float4(__int_as_float(v0), __int_as_float(v1), __int_as_float(v2), __int_as_float(v3)), feats_policy);
                    put_tile(o, v0, v1, v2, v3);
                }
                __syncwarp();
            }
            // ---- optional contract output: the padded voxels tensor [M, P, F], coalesced, one pillar at a time ----
            if (voxels_out) {
                unsigned todo = __ballot_sync(FULL, live);
                while (todo) {
                    const int o = __ffs(todo) - 1;
                    todo &= todo - 1;
                    const int cnt_o = __shfl_sync(FULL, cnt, o), rel_o = __shfl_sync(FULL, rel, o), f_o = __shfl_sync(FULL, f, o);
                    const int nk = min(cnt_o, Pmax);
                    float *vo = voxels_out + (size_t)f_o * Pmax * F;
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
            if (!STATS) {
                fence_proxy_async_smem();
                __syncwarp();
#ifdef HGSF_EXPERIMENT
                if (lane == 0 && !(p.dbg & 2)) { if (p.dbg & 4) tma_store_3d(&tmap, tile, x0, y, b * C); else tma_store_3d_hint(&tmap, tile, x0, y, b * C, stream_policy); tma_commit(); }
#else
                if (lane == 0) { tma_store_3d_hint(&tmap, tile, x0, y, b * C, stream_policy); tma_commit(); }
#endif
                store_pending = true;
            }
            __syncwarp();
        }
#ifdef HGSF_TILE_CLOCKS
        const long long tc2 = clock64();
#endif
        e_cur = e_nxt; r_cur = r_nxt; r_nxt = r_nn;
        cur = nxt; nxt = nxt2; nxt2 = next_tile();
#ifdef HGSF_TILE_CLOCKS
        const long long tc3 = clock64();
        tc_cyc[tc_class][0] += tc1 - tc0; tc_cyc[tc_class][1] += tc2 - tc1; tc_cyc[tc_class][2] += tc3 - tc2; ++tc_cnt[tc_class];
#endif
    }
#ifdef HGSF_TILE_CLOCKS
    if (lane == 0) {
        unsigned long long *dst = reinterpret_cast<unsigned long long *>(p.scan_desc + 3 * 2048 - 128);     // 64 u64 at the tail
        for (int c = 0; c < 5; ++c) {
            for (int q = 0; q < 3; ++q) atomicAdd(dst + c * 4 + q, tc_cyc[c][q]);
            atomicAdd(dst + c * 4 + 3, tc_cnt[c]);
        }
        atomicAdd(dst + 20, (unsigned long long)(clock64() - tc_begin));       // the warp's whole loop
        atomicAdd(dst + 21, 1ull);
    }
#endif
    cp_async_wait<0>();
    if (STATS) {
        // lanes l and l + 16 hold the same row; then the warps of the CTA, then one fp64 atomic per entry
#pragma unroll
        for (int k = 0; k < CIN; ++k) st_S[k] += __shfl_xor_sync(FULL, st_S[k], 16);
        __syncthreads();                                  // every warp is done with its tile buffer: reused for the reduction
        double *red = reinterpret_cast<double *>(tiles);  // [EMIT_WARPS][16][CIN]
        if (lane < 16) {
#pragma unroll
            for (int k = 0; k < CIN; ++k) red[(warp * 16 + lane) * CIN + k] = (lane <= CIN) ? st_S[k] : 0.0;
        }
        __syncthreads();
        for (int t = tid; t < 16 * CIN; t += NT) {
            double acc = 0.0;
            for (int w = 0; w < EMIT_WARPS; ++w) acc += red[w * 16 * CIN + t];
            if (acc != 0.0) atomicAdd(p.stats_S + t, acc);
        }
        // the last CTA to get here expands the moments: N = M * P rows per channel, the zero-padded rows included (they add
        // nothing to the sums); then train_ops.cu k_bn_finalize's arithmetic
        __shared__ int s_last;
        __threadfence();
        __syncthreads();
        if (tid == 0) s_last = (atomicAdd(p.ticket + 48, 1u) == gridDim.x - 1u) ? 1 : 0;
        __syncthreads();
        if (s_last) {
            __threadfence();
            double *Sm = red;                             // S rows 0..CIN-1, then s
            for (int t = tid; t < (CIN + 1) * CIN; t += NT) Sm[t] = __ldcg(p.stats_S + t);
            __syncthreads();
            if (tid < C) {
                double w[CIN];
#pragma unroll
                for (int k = 0; k < CIN; ++k) w[k] = (double)__ldg(p.W + tid * CIN + k);
                double sx = 0.0, sxx = 0.0;
#pragma unroll
                for (int k = 0; k < CIN; ++k) {
                    double tk = 0.0;                      // T[c,k] = sum_a W[c,a] S[a,k]
#pragma unroll
                    for (int a2 = 0; a2 < CIN; ++a2) tk = fma(w[a2], Sm[a2 * CIN + k], tk);
                    p.stats[2 * C + tid * CIN + k] = tk;
                    sxx = fma(tk, w[k], sxx);             // w^T S w
                    sx = fma(w[k], Sm[CIN * CIN + k], sx);
                }
                p.stats[tid] = sx;
                p.stats[C + tid] = sxx;
                const double n_rows = (double)s_K[p.B] * (double)p.P;
                if (n_rows > 0.0) {
                    const double mean = sx / n_rows;
                    double var = sxx / n_rows - mean * mean;
                    if (var < 0.0) var = 0.0;
                    p.batch_mean[tid] = (float)mean;
                    p.batch_var[tid] = (float)var;
                    if (p.run_mean) p.run_mean[tid] = (float)((1.0 - (double)p.momentum) * (double)p.run_mean[tid] + (double)p.momentum * mean);
                    if (p.run_var) {
                        const double unbiased = n_rows > 1.0 ? var * n_rows / (n_rows - 1.0) : var;
                        p.run_var[tid] = (float)((1.0 - (double)p.momentum) * (double)p.run_var[tid] + (double)p.momentum * unbiased);
                    }
                } else {                                  // no pillar at all: nothing to normalise, running statistics untouched
                    p.batch_mean[tid] = 0.f;
                    p.batch_var[tid] = 1.f;
                }
            }
            if (tid < CIN) p.stats[2 * C + C * CIN + tid] = Sm[CIN * CIN + tid];      // s
        }
        return;
    }
    if (lane <= 1) tma_wait_read<0>();        // shared memory must outlive the stores that read it
    mark_table_clean(p);
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
    // per device (a process may drive several GPUs, and they need not be the same part)
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
    if (cached[dev] == 0) {
        int v = 148;
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        cached[dev] = v;
    }
    return cached[dev];
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
static int launch_pillars_t(const PathParams &p, cudaStream_t stream) {
    constexpr int C = 64;
    constexpr int RWc = (F + 1 + 3) / 4 * 4;
    constexpr int SW = (RWc <= 8) ? HGSF_STAGE_ROWS : HGSF_STAGE_ROWS * 2 / 3;
    const bool canvas = PFN && p.canvas != nullptr;
    const size_t smem = (canvas ? sizeof(float) * PW * C * 32 : 0) + (PFN ? sizeof(float) * PW * 2 * SW * RWc : 0) +
                        sizeof(int) * 2 * (size_t)(p.B + 1);
    const long long n_runs = (((long long)p.B * p.cells >> 5) + CT_TAIL - 1) / CT_TAIL;    // upper bound
    const bool bn = p.bn_w != nullptr;
    auto go = [&](auto kern) -> int {
        int grid = 1;
        const int st = launch_persistent(kern, PT, smem, (n_runs + PW - 1) / PW, stream, &grid);
        if (st != HGSF_OK) return st;
        // programmatic dependent launch behind k_front (same stream): this kernel's launch latency overlaps k_front's last
        // phase; griddepcontrol.wait at the top of the kernel orders every dependent read after k_front's completion
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(PT); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
#ifdef HGSF_NO_PDL
        cfg.numAttrs = 0;
#else
        cfg.numAttrs = 1;
#endif
        return (int)cudaLaunchKernelEx(&cfg, kern, p);
    };
    const bool gen = (p.flags & HGSF_POINTS_SPCONV1_BREAK) || p.P > 32;
    if constexpr (PFN) {
        if (canvas) {
            if (gen) return bn ? go(k_pillars<F, ABS, DIST, true, true, true, true>) : go(k_pillars<F, ABS, DIST, false, true, true, true>);
            return bn ? go(k_pillars<F, ABS, DIST, true, true, true, false>) : go(k_pillars<F, ABS, DIST, false, true, true, false>);
        }
        // without a canvas (features only): the general build serves both
        return bn ? go(k_pillars<F, ABS, DIST, true, true, false, true>) : go(k_pillars<F, ABS, DIST, false, true, false, true>);
    }
    return go(k_pillars<F, ABS, DIST, true, false, false, true>);
}

template <int F, bool ABS, bool DIST>
static int launch_emit_t(const PathParams &p, cudaStream_t stream, bool stats_pass = false) {
    constexpr int C = 64;
    constexpr int RWc = (F + 1 + 3) / 4 * 4;
    CUtensorMap map, zmap;
    int st = make_canvas_map(&map, p.canvas, p.B, C, p.ny, p.nx, C);
    if (st == HGSF_OK) st = make_canvas_map(&zmap, p.canvas, p.B, C, p.ny, p.nx, C / 4);
    if (st != HGSF_OK) return st;
    const size_t smem = 1024 + sizeof(float) * (EMIT_WARPS * C * 32 + (C / 4) * 32 + EMIT_WARPS * 2 * STAGE_W * RWc) +
                        sizeof(int) * 2 * (size_t)(p.B + 1);
    const long long n_tiles = ((long long)p.B * p.cells) >> 5;
    if (n_tiles == 0) return HGSF_OK;
    int chunk = ((long long)p.n < 6 * n_tiles) ? 2 : 1;
    if (const char *tc = getenv("HGSF_TILE_CHUNK")) chunk = atoi(tc) >= 2 ? 2 : 1;
    const bool bn = p.bn_w != nullptr;
    auto go = [&](auto kern) -> int {
        int grid = 1;
        const int s2 = launch_persistent(kern, EMIT_THREADS, smem, (n_tiles + EMIT_WARPS - 1) / EMIT_WARPS, stream, &grid);
        if (s2 != HGSF_OK) return s2;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(EMIT_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
#ifdef HGSF_NO_PDL
        cfg.numAttrs = 0;
#else
        cfg.numAttrs = 1;
#endif
        return (int)cudaLaunchKernelEx(&cfg, kern, map, zmap, p);
    };
    if (stats_pass) return chunk == 2 ? go(k_emit<F, ABS, DIST, true, 2, true>) : go(k_emit<F, ABS, DIST, true, 1, true>);
    if (chunk == 2) return bn ? go(k_emit<F, ABS, DIST, true, 2, false>) : go(k_emit<F, ABS, DIST, false, 2, false>);
    return bn ? go(k_emit<F, ABS, DIST, true, 1, false>) : go(k_emit<F, ABS, DIST, false, 1, false>);
}

static int launch_pillars(const PathParams &p, bool with_pfn, bool abs_xyz, bool dist, cudaStream_t s) {
    if (!with_pfn) return launch_pillars_t<4, true, false, false>(p, s);   // F / RW are read from the params when the PFN is off
    if (p.C != 64) return HGSF_ERR_UNSUPPORTED;
    // the tile-major kernel takes the fused canvas case it was built for; everything else goes pillar-major
    static const bool force_pillars = getenv("HGSF_CONSUMER") && getenv("HGSF_CONSUMER")[0] == 'p';
    const bool tile_major = p.canvas != nullptr && p.canvas_vec && p.P <= 32 && !(p.flags & HGSF_POINTS_SPCONV1_BREAK) && !force_pillars;
#define HGSF_CASE(FV, A, D) if (p.F == FV && abs_xyz == A && dist == D) \
        return tile_major ? launch_emit_t<FV, A, D>(p, s) : launch_pillars_t<FV, A, D, true>(p, s);
    HGSF_CASE(4, true, false) HGSF_CASE(5, true, false) HGSF_CASE(6, true, false) HGSF_CASE(7, true, false)
    HGSF_CASE(8, true, false) HGSF_CASE(7, false, false) HGSF_CASE(8, false, false)
    HGSF_CASE(7, true, true) HGSF_CASE(8, true, true)
#undef HGSF_CASE
    return HGSF_ERR_UNSUPPORTED;
}

// ---- optional per-launch timing of the consumer kernel (k_pillars): bench.py's roofline leg ---------------------------------
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

// heavy-tile threshold, canvas store mode, CTA slices; then k_front (cooperative)
static int launch_front(PathParams &p, cudaStream_t stream) {
#ifdef HGSF_EXPERIMENT
    { static const int dbg = getenv("HGSF_DBG") ? atoi(getenv("HGSF_DBG")) : 0; p.dbg = dbg; }
#endif
    {
        // heavy tile = more than 3 times the average tile's points, at least 56 (HGSF_HEAVY_PTS overrides; huge = none).  Measured
        // (ms in k_emit at thresholds 146 / 40 / 28 / 20 / 12): stress 200 k 0.505 / 0.492 / 0.494 / 0.492 / 0.494;
        // VoD clustered at 56 / 40 / 28 / 20 / 12: 0.130 / 0.132 / 0.133 / 0.133 / 0.134
        const long long n_tt = ((long long)p.B * p.cells) >> 5;
        const long long avg3 = n_tt > 0 ? 3 * (long long)p.n / n_tt : 0;
        p.heavy_pts = (int)std::min<long long>(std::max<long long>(56, avg3), INT_MAX);
        if (const char *hp = getenv("HGSF_HEAVY_PTS")) p.heavy_pts = atoi(hp);
    }
    p.canvas_vec = (p.canvas != nullptr) && (p.nx % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.canvas) & 15) == 0) &&
                   !(getenv("HGSF_CANVAS_STORE") && getenv("HGSF_CANVAS_STORE")[0] == 's');       // 's': force scalar stores (tests)
    // cooperative launch: every CTA must be resident, so the grid is the occupancy limit (capped: ~4 CTAs/SM is
    // plenty of parallelism for a latency-bound front end and keeps the grid barriers cheap)
    static int per_sm_cached = []() {
        int per_sm = 1;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_front, FRONT_THREADS, 0);
        const char *env = getenv("HGSF_FRONT_CTAS");
        const int cap = env ? atoi(env) : 4;
        return std::max(1, std::min(per_sm, cap));
    }();
    const int max_ctas = std::min(sm_count() * per_sm_cached, MAX_FRONT_CTAS);
    const long long n_cells = (long long)p.B * p.cells;
    const long long want = std::max<long long>((n_cells + FRONT_THREADS * 8 - 1) / (FRONT_THREADS * 8),
                                               ((long long)p.n + FRONT_THREADS - 1) / FRONT_THREADS);
    const int grid = (int)std::max<long long>(1, std::min<long long>(want, max_ctas));
    // every CTA gets the same number of points / cells (multiples of 32)
    const long long n_pad = ((long long)p.n + 31) / 32 * 32;
    p.pslice = (int)(((n_pad + grid - 1) / grid + 31) / 32 * 32);
    p.cslice = (int)(((n_cells + grid - 1) / grid + 31) / 32 * 32);
    PathParams pp = p;
    void *args[] = {&pp};
    return (int)cudaLaunchCooperativeKernel((const void *)k_front, dim3(grid), dim3(FRONT_THREADS), args, 0, stream);
}

int launch_pillar_path(const PathParams &p_in, bool with_pfn, bool abs_xyz, bool dist, cudaStream_t stream, int *launches) {
    PathParams p = p_in;
    p.stats = nullptr;
    int st = launch_front(p, stream);
    if (st != HGSF_OK) return st;
    const bool timed = g_timing.capacity > 0 && g_timing.count < g_timing.capacity;
    if (timed) cudaEventRecord(g_timing.ev[2 * g_timing.count], stream);
    st = launch_pillars(p, with_pfn, abs_xyz, dist, stream);
    if (st != HGSF_OK) return st;
    if (timed) cudaEventRecord(g_timing.ev[2 * g_timing.count++ + 1], stream);
    if (launches) *launches = 2;
    return HGSF_OK;
}

int launch_pillar_path_train(const PathParams &p_in, bool abs_xyz, bool dist, cudaStream_t stream, int *launches) {
    PathParams p = p_in;
    if (!p.stats || !p.batch_mean || !p.batch_var || !p.bn_w || !p.bn_b || !p.canvas || p.C != 64) return HGSF_ERR_INVALID_ARG;
    if (p.nx % 4 != 0 || (reinterpret_cast<uintptr_t>(p.canvas) & 15) || p.P > 32 || (p.flags & HGSF_POINTS_SPCONV1_BREAK))
        return HGSF_ERR_UNSUPPORTED;
    int st = launch_front(p, stream);
    if (st != HGSF_OK) return st;
    if (!p.canvas_vec) return HGSF_ERR_UNSUPPORTED;      // HGSF_CANVAS_STORE=s (tests of the scalar store path)
    // pass 1 (statistics) normalises nothing: it needs valid pointers only; pass 2 normalises with what pass 1 left
    p.bn_m = p.batch_mean; p.bn_v = p.batch_var;
    st = HGSF_ERR_UNSUPPORTED;
#define HGSF_CASE(FV, A, D) if (p.F == FV && abs_xyz == A && dist == D) { \
        st = launch_emit_t<FV, A, D>(p, stream, true); if (st == HGSF_OK) st = launch_emit_t<FV, A, D>(p, stream, false); }
    HGSF_CASE(4, true, false) HGSF_CASE(5, true, false) HGSF_CASE(6, true, false) HGSF_CASE(7, true, false)
    HGSF_CASE(8, true, false) HGSF_CASE(7, false, false) HGSF_CASE(8, false, false)
    HGSF_CASE(7, true, true) HGSF_CASE(8, true, true)
#undef HGSF_CASE
    if (st != HGSF_OK) return st;
    if (launches) *launches = 3;
    return HGSF_OK;
}

}  // namespace hgsf
