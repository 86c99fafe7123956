// pillarnet_ops.cu -- sm_100a equivalents of the reference's "Path B" native ops (the PillarNet reader that the
// shipped HGSFusion YAMLs run): pybind module `pillar_cuda`, pcdet/ops/pillar_ops/src/pillar_api.cpp:10-22.
//
//   reference kernel (file:line under pcdet/ops/pillar_ops/src)            here
//   createPillarIndicesStackKernel      pillar_ops_gpu.cu:13-38   \
//   torch.cumsum + .item()              pillar_utils.py:106-110    |
//   createPillarIndicesKernel           pillar_ops_gpu.cu:55-72    |  k_pillarnet_indices: ONE cooperative kernel,
//   createPillarIndicePairsStackKernel  pillar_ops_gpu.cu:89-117   |  no host sync (counts stay on the device)
//   torch.cumsum + .item()              group_utils.py:21-24       |
//   flattenIndicePairsKernel            group_ops_gpu.cu:9-24     /
//   gather_feature_kernel               group_ops_gpu.cu:42-55       k_gather        (warp per row, coalesced)
//   gather_feature_grad_kernel          group_ops_gpu.cu:57-70       k_gather_grad   (red.add, coalesced)
//   scatter_max_kernel + atomics.cuh CAS group scatter_ops_gpu.cu:13-25  k_scatter_max  (native integer atomicMax on the float bits)
//   scatter_arg_max_kernel              scatter_ops_gpu.cu:27-46     k_scatter_arg   (deterministic: the largest qualifying id)
//   scatter_max_grad_kernel             scatter_ops_gpu.cu:48-58     k_scatter_max_grad
//
// Integer results (pillars, pillar_bev_indices, indice_pairs, point/pillar index lists, M, L) are bit-identical to
// the reference kernels; scatter_max is exact (max is order independent); `arg` matches up to the reference's own
// last-writer-wins ties.
#include <cooperative_groups.h>

#include <algorithm>

#include "pillarnet_ops.cuh"

namespace cg = cooperative_groups;

namespace hgsf {

constexpr int PB_THREADS = 256;
constexpr int PB_ITEMS = 4;
constexpr int PB_TILE = PB_THREADS * PB_ITEMS;

__device__ __forceinline__ uint32_t pb_block_sum(uint32_t v, uint32_t *s_warp, int lane, int warp) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(FULL, v, d);
    __syncthreads();
    if (lane == 0) s_warp[warp] = v;
    __syncthreads();
    uint32_t t = 0;
#pragma unroll
    for (int w = 0; w < PB_THREADS / 32; ++w) t += s_warp[w];
    return t;
}

// block-wide exclusive scan of 4 consecutive items per thread; returns the tile total
__device__ __forceinline__ uint32_t pb_tile_scan(const uint32_t (&v)[PB_ITEMS], uint32_t (&excl)[PB_ITEMS], uint32_t *s_warp,
                                                 int lane, int warp) {
    uint32_t local = 0;
#pragma unroll
    for (int j = 0; j < PB_ITEMS; ++j) local += v[j];
    uint32_t incl = local;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(FULL, incl, d);
        if (lane >= d) incl += o;
    }
    __syncthreads();
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t warp_off = 0, total = 0;
#pragma unroll
    for (int w = 0; w < PB_THREADS / 32; ++w) {
        const uint32_t t = s_warp[w];
        if (w < warp) warp_off += t;
        total += t;
    }
    uint32_t run = warp_off + incl - local;
#pragma unroll
    for (int j = 0; j < PB_ITEMS; ++j) { excl[j] = run; run += v[j]; }
    return total;
}

// One cooperative launch for gen_indice_pairs + flatten_indices.
//   phase 0  bev[B,H,W] = 0
//   phase 1  per point: frame (prefix of xyz_batch_cnt, as the reference scans it), cell = (int(y/s), int(x/s)) with
//            IEEE divide and truncation toward zero; bev[cell] = 1; key[p] = cell or -1
//   phase 2  two-level exclusive scan of bev over cells in (b, y, x) order: bev[cell] = pillar id or -1,
//            pillars[id] = (b, y, x); M
//   phase 3  two-level exclusive scan over points of "has a pillar": indice_pairs[p], and the compacted lists
//            point_idx[pos] = p, pillar_idx[pos] = id in input order; L
__global__ void __launch_bounds__(PB_THREADS) k_pillarnet_indices(const PillarNetParams q) {
    cg::grid_group grid = cg::this_grid();
    __shared__ uint32_t s_warp[PB_THREADS / 32];
    extern __shared__ int s_incl[];      // [B] inclusive prefix of xyz_batch_cnt
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long gtid = (long long)blockIdx.x * PB_THREADS + tid, nthr = (long long)gridDim.x * PB_THREADS;
    const long long n_cells = (long long)q.B * q.H * q.W;

    if (tid == 0) {
        int acc = 0;
        for (int b = 0; b < q.B; ++b) { acc += q.cnt[b]; s_incl[b] = acc; }
    }
    for (long long i = gtid; i < n_cells; i += nthr) q.bev[i] = 0;
    __syncthreads();
    grid.sync();
    // ---- phase 1 ----
    for (long long p = gtid; p < q.N; p += nthr) {
        // frame of point p exactly as pillar_ops_gpu.cu:22-27 finds it: the first b with p < cnt[0]+..+cnt[b], else B-1
        int bid = q.B - 1;
        for (int b = 0; b < q.B - 1; ++b)
            if (p < s_incl[b]) { bid = b; break; }
        const float x = __ldg(q.xyz + 3 * p), y = __ldg(q.xyz + 3 * p + 1);
        const int xid = __float2int_rz(__fdiv_rn(x, q.bev_size));     // int(x / bev_size): IEEE divide, trunc toward zero
        const int yid = __float2int_rz(__fdiv_rn(y, q.bev_size));
        int key = -1;
        if (!(xid < 0 || xid >= q.W || yid < 0 || yid >= q.H)) {
            key = (bid * q.H + yid) * q.W + xid;
            q.bev[key] = 1;
        }
        q.key[p] = key;
    }
    grid.sync();
    // ---- phase 2: cells ----
    const long long ctiles = (n_cells + PB_TILE - 1) / PB_TILE;
    const long long ctiles_per = (ctiles + gridDim.x - 1) / gridDim.x;
    const long long clo = min((long long)blockIdx.x * ctiles_per, ctiles) * PB_TILE;
    const long long chi = min(((long long)blockIdx.x + 1) * ctiles_per, ctiles) * PB_TILE;
    {
        uint32_t mine = 0;
        for (long long t0 = clo; t0 < chi; t0 += PB_TILE)
#pragma unroll
            for (int j = 0; j < PB_ITEMS; ++j) {
                const long long c = t0 + tid * PB_ITEMS + j;
                if (c < n_cells) mine += (uint32_t)q.bev[c];
            }
        const uint32_t total = pb_block_sum(mine, s_warp, lane, warp);
        if (tid == 0) q.partial[blockIdx.x] = total;
    }
    grid.sync();
    {
        uint32_t before = 0;
        for (int c = tid; c < (int)blockIdx.x; c += PB_THREADS) before += q.partial[c];
        uint32_t carry = pb_block_sum(before, s_warp, lane, warp);
        const int HW = q.H * q.W;
        for (long long t0 = clo; t0 < chi; t0 += PB_TILE) {
            const long long c0 = t0 + tid * PB_ITEMS;
            uint32_t v[PB_ITEMS], ex[PB_ITEMS];
#pragma unroll
            for (int j = 0; j < PB_ITEMS; ++j) v[j] = (c0 + j < n_cells) ? (uint32_t)q.bev[c0 + j] : 0u;
            const uint32_t total = pb_tile_scan(v, ex, s_warp, lane, warp);
#pragma unroll
            for (int j = 0; j < PB_ITEMS; ++j) {
                const long long c = c0 + j;
                if (c < n_cells) {
                    if (v[j]) {
                        const int id = (int)(carry + ex[j]);
                        q.bev[c] = id;
                        const int b = (int)(c / HW), rem = (int)(c - (long long)b * HW);
                        const int yy = rem / q.W;
                        q.pillars[3 * (size_t)id + 0] = b;
                        q.pillars[3 * (size_t)id + 1] = yy;
                        q.pillars[3 * (size_t)id + 2] = rem - yy * q.W;
                    } else {
                        q.bev[c] = -1;
                    }
                }
            }
            carry += total;
        }
        if (blockIdx.x == gridDim.x - 1 && tid == 0) q.counts[0] = (int)carry;   // the last slice ends at the last cell
    }
    grid.sync();
    // ---- phase 3: points ----
    const long long ptiles = (q.N + PB_TILE - 1) / PB_TILE;
    const long long ptiles_per = (ptiles + gridDim.x - 1) / gridDim.x;
    const long long plo = min((long long)blockIdx.x * ptiles_per, ptiles) * PB_TILE;
    const long long phi = min(((long long)blockIdx.x + 1) * ptiles_per, ptiles) * PB_TILE;
    {
        uint32_t mine = 0;
        for (long long t0 = plo; t0 < phi; t0 += PB_TILE)
#pragma unroll
            for (int j = 0; j < PB_ITEMS; ++j) {
                const long long p = t0 + tid * PB_ITEMS + j;
                if (p < q.N && q.key[p] >= 0) ++mine;
            }
        const uint32_t total = pb_block_sum(mine, s_warp, lane, warp);
        if (tid == 0) q.partial[4096 + blockIdx.x] = total;
    }
    grid.sync();
    {
        uint32_t before = 0;
        for (int c = tid; c < (int)blockIdx.x; c += PB_THREADS) before += q.partial[4096 + c];
        uint32_t carry = pb_block_sum(before, s_warp, lane, warp);
        for (long long t0 = plo; t0 < phi; t0 += PB_TILE) {
            const long long p0 = t0 + tid * PB_ITEMS;
            uint32_t v[PB_ITEMS], ex[PB_ITEMS];
            int id[PB_ITEMS];
#pragma unroll
            for (int j = 0; j < PB_ITEMS; ++j) {
                id[j] = -1;
                if (p0 + j < q.N) { const int k = q.key[p0 + j]; if (k >= 0) id[j] = q.bev[k]; }
                v[j] = id[j] >= 0 ? 1u : 0u;
            }
            const uint32_t total = pb_tile_scan(v, ex, s_warp, lane, warp);
#pragma unroll
            for (int j = 0; j < PB_ITEMS; ++j) {
                const long long p = p0 + j;
                if (p < q.N) {
                    if (q.pairs) q.pairs[p] = id[j];
                    if (v[j]) {
                        const size_t pos = carry + ex[j];
                        q.point_idx[pos] = (int)p;
                        q.pillar_idx[pos] = id[j];
                    }
                }
            }
            carry += total;
        }
        if (blockIdx.x == gridDim.x - 1 && tid == 0) q.counts[1] = (int)carry;
    }
}

int launch_pillarnet_indices(const PillarNetParams &q, cudaStream_t stream) {
    static int max_ctas = []() {
        int per_sm = 1, dev = 0, sms = 148;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pillarnet_indices, PB_THREADS, 4096);
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        return sms * std::max(1, std::min(per_sm, 4));
    }();
    const long long n_cells = (long long)q.B * q.H * q.W;
    const long long want = std::max<long long>((n_cells + PB_TILE - 1) / PB_TILE, (q.N + PB_THREADS - 1) / PB_THREADS);
    const int grid = (int)std::max<long long>(1, std::min<long long>(std::min<long long>(want, max_ctas), 4096));
    PillarNetParams qq = q;
    void *args[] = {&qq};
    const size_t smem = sizeof(int) * (size_t)q.B;
    if (smem > 4096) return HGSF_ERR_UNSUPPORTED;
    return (int)cudaLaunchCooperativeKernel((const void *)k_pillarnet_indices, dim3(grid), dim3(PB_THREADS), args, smem, stream);
}

// ------------------------------------------------------------------------------------------------
// gather / gather_grad: a warp per output row, lanes over the C columns (the reference walks a row per thread)
__global__ void __launch_bounds__(256) k_gather(long long L, int C, const int *__restrict__ idx, const float *__restrict__ f,
                                                float *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const long long w0 = ((long long)blockIdx.x * 256 + threadIdx.x) >> 5, nw = ((long long)gridDim.x * 256) >> 5;
    for (long long l = w0; l < L; l += nw) {
        const float *src = f + (size_t)__ldg(idx + l) * C;
        float *dst = out + (size_t)l * C;
        for (int c = lane; c < C; c += 32) dst[c] = __ldg(src + c);
    }
}

__global__ void __launch_bounds__(256) k_gather_grad(long long L, int C, const int *__restrict__ idx,
                                                     const float *__restrict__ gout, float *__restrict__ gin) {
    const int lane = threadIdx.x & 31;
    const long long w0 = ((long long)blockIdx.x * 256 + threadIdx.x) >> 5, nw = ((long long)gridDim.x * 256) >> 5;
    for (long long l = w0; l < L; l += nw) {
        float *dst = gin + (size_t)__ldg(idx + l) * C;
        const float *src = gout + (size_t)l * C;
        for (int c = lane; c < C; c += 32) atomicAdd(dst + c, __ldg(src + c));
    }
}

// ------------------------------------------------------------------------------------------------
// scatter_max: out[c, index[p]] = max(0, max_p src[c, p]).  out starts at 0, so only positive values can win and for
// them the float order is the integer order of the bit patterns: one native atomicMax instead of a CAS loop.
__global__ void __launch_bounds__(256) k_scatter_max(int C, long long L, long long M, const int *__restrict__ index,
                                                     const float *__restrict__ src, float *__restrict__ out) {
    const long long n = (long long)C * L;
    for (long long t = (long long)blockIdx.x * 256 + threadIdx.x; t < n; t += (long long)gridDim.x * 256) {
        const long long c = t / L, p = t - c * L;
        const float v = __ldg(src + t);
        if (v > 0.f) atomicMax(reinterpret_cast<int *>(out + c * M + __ldg(index + p)), __float_as_int(v));
    }
}

// arg[c, m] = some flat id c*L+p with |src - out| < 1e-5 (the reference: last writer wins); here: the largest such id
__global__ void __launch_bounds__(256) k_scatter_arg(int C, long long L, long long M, const int *__restrict__ index,
                                                     const float *__restrict__ src, const float *__restrict__ out, int *__restrict__ arg) {
    const long long n = (long long)C * L;
    for (long long t = (long long)blockIdx.x * 256 + threadIdx.x; t < n; t += (long long)gridDim.x * 256) {
        const long long c = t / L, p = t - c * L;
        const long long o = c * M + __ldg(index + p);
        if (fabsf(__ldg(src + t) - __ldg(out + o)) < 1e-5f) atomicMax(arg + o, (int)t);
    }
}

__global__ void __launch_bounds__(256) k_scatter_max_grad(long long n, const int *__restrict__ arg, const float *__restrict__ gout,
                                                          float *__restrict__ gsrc) {
    for (long long t = (long long)blockIdx.x * 256 + threadIdx.x; t < n; t += (long long)gridDim.x * 256) {
        const int a = __ldg(arg + t);
        if (a >= 0) gsrc[a] = __ldg(gout + t);
    }
}

static unsigned grid_for(long long work_threads) {
    const long long blocks = (work_threads + 255) / 256;
    return (unsigned)std::max<long long>(1, std::min<long long>(blocks, 148 * 32));
}

int launch_gather(long long L, int C, const int *idx, const float *f, float *out, cudaStream_t s) {
    if (L == 0) return HGSF_OK;
    k_gather<<<grid_for(L * 32), 256, 0, s>>>(L, C, idx, f, out);
    return (int)cudaGetLastError();
}
int launch_gather_grad(long long L, int C, const int *idx, const float *gout, float *gin, cudaStream_t s) {
    if (L == 0) return HGSF_OK;
    k_gather_grad<<<grid_for(L * 32), 256, 0, s>>>(L, C, idx, gout, gin);
    return (int)cudaGetLastError();
}
int launch_scatter_max(int C, long long L, long long M, const int *index, const float *src, int *arg, float *out, cudaStream_t s,
                       int *launches) {
    cudaError_t e = cudaMemsetAsync(out, 0, sizeof(float) * (size_t)C * M, s);
    if (e == cudaSuccess && arg) e = cudaMemsetAsync(arg, 0xFF, sizeof(int) * (size_t)C * M, s);   // -1
    if (e != cudaSuccess) return (int)e;
    int nl = 0;
    if (L > 0 && M > 0) {
        k_scatter_max<<<grid_for((long long)C * L), 256, 0, s>>>(C, L, M, index, src, out);
        ++nl;
        if (arg) { k_scatter_arg<<<grid_for((long long)C * L), 256, 0, s>>>(C, L, M, index, src, out, arg); ++nl; }
    }
    if (launches) *launches = nl;
    return (int)cudaGetLastError();
}
int launch_scatter_max_grad(int C, long long M, const int *arg, const float *gout, float *gsrc, cudaStream_t s) {
    if ((long long)C * M == 0) return HGSF_OK;
    k_scatter_max_grad<<<grid_for((long long)C * M), 256, 0, s>>>((long long)C * M, arg, gout, gsrc);
    return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Reader input prep (vfe/pillarnet.py:51-58 + dynamic_pillar_encoder.py:55-118): split the collated points by frame, build
// the "split" real/virtual feature rows and the range-relative xyz.  A warp stages 32 input rows in shared memory (one
// coalesced span when the rows are taken in input order) and writes 32 output rows as one coalesced span.
//   info[0] bit 0: the emitted order is not the reference's (rows not grouped by ascending frame, or a dropped row before
//                  a kept one) -> the caller must pass `order`;  info[1] = rows kept.
constexpr int ENC_THREADS = 256;
constexpr int ENC_MAX_W = 40;     // 1 + Fin

__global__ void __launch_bounds__(ENC_THREADS) k_split_encode(const SplitEncodeParams q) {
    __shared__ float s_rows[ENC_THREADS / 32][32 * ENC_MAX_W];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    float *rows = s_rows[wib];
    const int W = q.Fin + 1, Fo = q.Fout, n = q.n_split;
    const long long nw = ((long long)gridDim.x * ENC_THREADS) >> 5;
    for (long long r0 = ((((long long)blockIdx.x * ENC_THREADS) >> 5) + wib) * 32; r0 < q.L; r0 += nw * 32) {
        const int nr = (int)min((long long)32, q.L - r0);
        if (q.order == nullptr) {
            const float *src = q.points + (size_t)r0 * W;
            for (int e = lane; e < nr * W; e += 32) rows[e] = __ldg(src + e);
        } else {
            for (int e = lane; e < nr * W; e += 32) {
                const int r = e / W, c = e - r * W;
                rows[e] = __ldg(q.points + (size_t)__ldg(q.order + r0 + r) * W + c);
            }
        }
        __syncwarp();
        // frame key of my row: the frame index when the reference keeps the row (`points[:,0] == i` for an i in [0, B)), else B
        int key = q.B;
        if (lane < nr) {
            const float b = rows[lane * W];
            if (b >= 0.f && b < (float)q.B && b == truncf(b)) key = (int)b;
        }
        int prev = __shfl_up_sync(FULL, key, 1);
        if (lane == 0) {
            prev = 0;
            if (r0 > 0) {
                const long long pr = q.order ? (long long)__ldg(q.order + r0 - 1) : r0 - 1;
                const float b = __ldg(q.points + (size_t)pr * W);
                prev = (b >= 0.f && b < (float)q.B && b == truncf(b)) ? (int)b : q.B;
            }
        }
        const bool keep = key < q.B;
        const unsigned bad = __ballot_sync(FULL, lane < nr && key < prev);
        const unsigned peers = __match_any_sync(FULL, key);
        if (keep && lane == __ffs(peers) - 1) atomicAdd(q.cnt + key, __popc(peers));
        const unsigned kept = __ballot_sync(FULL, keep);
        if (lane == 0) {
            if (bad) atomicOr(q.info, 1);
            if (kept) atomicAdd(q.info + 1, __popc(kept));
        }
        // xyz relative to the range minimum: one rounded fp32 subtract per coordinate (absl_to_relative, :46-53)
        for (int e = lane; e < nr * 3; e += 32) {
            const int r = e / 3, c = e - r * 3;
            if ((kept >> r) & 1) q.xyz[(size_t)r0 * 3 + e] = __fsub_rn(rows[r * W + 1 + c], c == 0 ? q.pc_min[0] : (c == 1 ? q.pc_min[1] : q.pc_min[2]));
        }
        for (int e = lane; e < nr * Fo; e += 32) {
            const int r = e / Fo, c = e - r * Fo;
            if (!((kept >> r) & 1)) continue;
            const float *row = rows + r * W + 1;
            float v;
            if (q.mode == HGSF_ENCODE_SPLIT) {
                const bool virt = row[q.Fin - 2] < 0.5f;                               // :68
                if (c < 3) v = row[c];                                                  // :69
                else if (c >= Fo - 2) v = row[q.Fin - 2 + (c - (Fo - 2))];              // :80-81
                else if (c < 3 + n) v = virt ? 0.f : row[c];                            // :72 / :75
                else if (c < 3 + 2 * n) v = virt ? row[c - n] : 0.f;                    // :73 / :76
                else v = 0.f;
            } else {
                v = row[c];                                                             // mixed / plain: all columns; direct: Fout = Fin - 2
            }
            q.feat[(size_t)r0 * Fo + e] = v;
        }
        __syncwarp();
    }
}

int launch_split_encode(const SplitEncodeParams &q, cudaStream_t s) {
    cudaError_t e = cudaMemsetAsync(q.cnt, 0, sizeof(int) * (size_t)q.B, s);
    if (e == cudaSuccess) e = cudaMemsetAsync(q.info, 0, sizeof(int) * 2, s);
    if (e != cudaSuccess) return (int)e;
    if (q.L == 0) return HGSF_OK;
    k_split_encode<<<grid_for(q.L), ENC_THREADS, 0, s>>>(q);
    return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// The reader's forward in ONE kernel (eval mode): PillarQueryAndGroup's gather + centre offsets (pillar_utils.py:31-54),
// the shared MLP Linear(no bias) + BatchNorm1d(eval) + ReLU (pillar_modules.py:19-26,76) and scatter_max
// (scatter_ops_gpu.cu:13-25) -- the reference materialises group_features [L, Cf+6], the MLP output [L, 32] and its
// transpose on the way.  A warp takes 32 consecutive grouped points; lane = output channel (C = 32) with its Linear
// row in registers; the staged point rows are read as shared-memory broadcasts; consecutive points of the same
// pillar are max-combined in registers and leave as one 128-byte integer atomicMax (values are post-ReLU >= 0, the
// output starts at 0: float order = integer order of the bits).
constexpr int RD_WARPS = 8;
constexpr int RD_MAXCIN = 40;

template <int CIN4>   // Cin padded to a multiple of 4, in float4s
__global__ void __launch_bounds__(RD_WARPS * 32) k_reader_fused(const ReaderParams q) {
    constexpr int CINP = CIN4 * 4;
    __shared__ __align__(16) float s_rows[RD_WARPS][32][CINP];
    __shared__ int s_pid[RD_WARPS][32], s_pt[RD_WARPS][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float(*rows)[CINP] = s_rows[warp];
    int *pid = s_pid[warp], *spt = s_pt[warp];
    const int Cf = q.Cf, Cin = Cf + 6;
    float w[CINP];
#pragma unroll
    for (int k = 0; k < CINP; ++k) w[k] = (k < Cin) ? __ldg(q.W + (size_t)lane * Cin + k) : 0.f;
    const float mu = __ldg(q.bn_m + lane), iv = __frcp_rn(__fsqrt_rn(__fadd_rn(__ldg(q.bn_v + lane), q.eps)));
    const float ga = __ldg(q.bn_w + lane), be = __ldg(q.bn_b + lane);
    const long long nchunks = (q.L + 31) >> 5;
    const long long nw = (long long)gridDim.x * RD_WARPS;
    for (long long ch = (long long)blockIdx.x * RD_WARPS + warp; ch < nchunks; ch += nw) {
        const long long l0 = ch << 5;
        const int np = (int)min((long long)32, q.L - l0);
        // ---- stage: lane l owns grouped point l0 + l ----
        int my_pt = 0, my_pid = -1;
        if (lane < np) { my_pt = __ldg(q.point_idx + l0 + lane); my_pid = __ldg(q.pillar_idx + l0 + lane); }
        pid[lane] = my_pid;
        spt[lane] = my_pt;
        __syncwarp();
        // feature rows: the 32 x Cf elements as one flat index space, so that the loads are independent and batch up
        const int total = np * Cf;
#pragma unroll 4
        for (int e = lane; e < total; e += 32) {
            const int r = e / Cf, k = e - r * Cf;
            rows[r][k] = __ldg(q.feat + (size_t)spt[r] * Cf + k);
        }
        if (lane < np) {
            const float x = __ldg(q.xyz + (size_t)my_pt * 3), y = __ldg(q.xyz + (size_t)my_pt * 3 + 1), z = __ldg(q.xyz + (size_t)my_pt * 3 + 2);
            const int py = __ldg(q.pillars + (size_t)my_pid * 3 + 1), px = __ldg(q.pillars + (size_t)my_pid * 3 + 2);
            // centres (pillar_utils.py:117-121): (idx + 0.5) * pillar_size in fp32; z centre is the range's absolute mid height
            const float cx = __fmul_rn(__fadd_rn((float)px, 0.5f), q.bev_size), cy = __fmul_rn(__fadd_rn((float)py, 0.5f), q.bev_size);
            float *g = rows[lane];
            g[Cf] = x; g[Cf + 1] = y; g[Cf + 2] = z;
            g[Cf + 3] = __fsub_rn(x, cx); g[Cf + 4] = __fsub_rn(y, cy); g[Cf + 5] = __fsub_rn(z, q.z_center);
            for (int k = Cin; k < CINP; ++k) g[k] = 0.f;
        }
        __syncwarp();
        // ---- lane = channel: Linear + BN + ReLU, max over runs of equal pillar ----
        int run_pid = pid[0];
        float run_max = 0.f;
        for (int r = 0; r < np; ++r) {
            const int p_r = pid[r];
            if (p_r != run_pid) {
                if (run_max > 0.f) atomicMax(reinterpret_cast<int *>(q.out + (size_t)run_pid * 32 + lane), __float_as_int(run_max));
                run_pid = p_r; run_max = 0.f;
            }
            const float4 *g4 = reinterpret_cast<const float4 *>(rows[r]);
            float acc = 0.f;
#pragma unroll
            for (int v = 0; v < CIN4; ++v) {
                const float4 g = g4[v];
                acc = fmaf(g.x, w[4 * v], acc); acc = fmaf(g.y, w[4 * v + 1], acc);
                acc = fmaf(g.z, w[4 * v + 2], acc); acc = fmaf(g.w, w[4 * v + 3], acc);
            }
            const float y = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(acc, mu), iv), ga), be);
            run_max = fmaxf(run_max, y);
        }
        if (run_max > 0.f) atomicMax(reinterpret_cast<int *>(q.out + (size_t)run_pid * 32 + lane), __float_as_int(run_max));
        __syncwarp();
    }
}

int launch_reader_fused(const ReaderParams &q, cudaStream_t s, int *launches) {
    if (launches) *launches = 0;
    cudaError_t e = cudaMemsetAsync(q.out, 0, sizeof(float) * (size_t)q.M * 32, s);
    if (e != cudaSuccess) return (int)e;
    if (q.L == 0 || q.M == 0) return HGSF_OK;
    const int cin4 = (q.Cf + 6 + 3) / 4;
    const long long chunks = (q.L + 31) / 32;
    const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>((chunks + RD_WARPS - 1) / RD_WARPS, 148 * 3));
    switch (cin4) {
#define HGSF_RD(N) case N: k_reader_fused<N><<<grid, RD_WARPS * 32, 0, s>>>(q); break;
        HGSF_RD(2) HGSF_RD(3) HGSF_RD(4) HGSF_RD(5) HGSF_RD(6) HGSF_RD(7) HGSF_RD(8) HGSF_RD(9) HGSF_RD(10)
#undef HGSF_RD
        default: return HGSF_ERR_UNSUPPORTED;
    }
    if (launches) *launches = 1;
    return (int)cudaGetLastError();
}

}  // namespace hgsf
