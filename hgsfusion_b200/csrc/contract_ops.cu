// contract_ops.cu -- the batch_dict contract entry points as stand-alone ops:
//   k_vfe            PillarVFE.forward on padded voxels [M,P,F]      (pillar_vfe.py:94-123)
//   k_scatter_map /  PointPillarScatter.forward                       (pointpillar_scatter.py:14-41)
//   k_scatter_canvas
// Same arithmetic as the fused path (pfn.cuh); these exist so that a detector that still voxelizes
// on the CPU (DATA_PROCESSOR: transform_points_to_voxels) can swap in the two modules one at a time.
#include "contract_ops.cuh"
#include "pfn.cuh"

#include <algorithm>
#include <cstring>

namespace hgsf {

// ------------------------------------------------------------------------------------------------
template <int F, bool ABS, bool DIST, int C, int NWARPS>
__global__ void __launch_bounds__(NWARPS * 32) k_vfe(const VfeParams q) {
    using Lane = PfnLane<F, ABS, DIST, C>;
    constexpr int CPL = Lane::CPL;
    extern __shared__ float s_vox[];                    // [NWARPS][P*F]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *buf = s_vox + (size_t)warp * q.P * F;
    Lane pfn;
    pfn.load(q.pfn, lane);
    const int P4 = (q.P >> 2) << 2;
    const long long nwarps = (long long)gridDim.x * NWARPS;
    for (long long m = (long long)blockIdx.x * NWARPS + warp; m < q.M; m += nwarps) {
        const float *src = q.voxels + (size_t)m * q.P * F;
        for (int t = lane; t < q.P * F; t += 32) buf[t] = __ldg(src + t);
        __syncwarp();
        // voxel_num_points / voxel_coords arrive as float32 in the reference (pcdet/models/__init__.py:36)
        const float nf = q.num_float ? __ldg(reinterpret_cast<const float *>(q.num) + m)
                                     : (float)__ldg(reinterpret_cast<const int32_t *>(q.num) + m);
        const int cnt = min(max((int)nf, 0), q.P);      // mask = num.int() > arange(P)  (pillar_vfe.py:87-91)
        float cz_i, cy_i, cx_i;
        if (q.coords_float) {
            const float4 c = __ldg(reinterpret_cast<const float4 *>(q.coords) + m);
            cz_i = c.y; cy_i = c.z; cx_i = c.w;
        } else {
            const int4 c = __ldg(reinterpret_cast<const int4 *>(q.coords) + m);
            cz_i = (float)c.y; cy_i = (float)c.z; cx_i = (float)c.w;
        }
        // the reference sums all P slots, padding included (pillar_vfe.py:97)
        SlotSum sum;
        for (int s = 0; s < q.P; ++s) sum.add(s, P4, buf[s * F], buf[s * F + 1], buf[s * F + 2]);
        const float mx = __fdiv_rn(sum.sx(), nf), my = __fdiv_rn(sum.sy(), nf), mz = __fdiv_rn(sum.sz(), nf);
        const float cx = __fadd_rn(__fmul_rn(cx_i, q.vsize[0]), q.voff[0]);
        const float cy = __fadd_rn(__fmul_rn(cy_i, q.vsize[1]), q.voff[1]);
        const float cz = __fadd_rn(__fmul_rn(cz_i, q.vsize[2]), q.voff[2]);
        float vmax[CPL];
        pfn.init_max(vmax, cnt < q.P);
        for (int s = 0; s < cnt; ++s) {
            float rowf[F];
#pragma unroll
            for (int k = 0; k < F; ++k) rowf[k] = buf[s * F + k];
            pfn.point(rowf, mx, my, mz, cx, cy, cz, vmax);
        }
#pragma unroll
        for (int j = 0; j < CPL; ++j) q.out[(size_t)m * C + lane + 32 * j] = vmax[j];
        __syncwarp();
    }
}

template <int F, bool ABS, bool DIST, int C>
static int launch_vfe_t(const VfeParams &q, cudaStream_t stream) {
    constexpr int NW = 4;
    auto kern = k_vfe<F, ABS, DIST, C, NW>;
    const size_t smem = sizeof(float) * NW * (size_t)q.P * F;
    if (smem > 200 * 1024) return HGSF_ERR_UNSUPPORTED;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    const long long want = (q.M + NW - 1) / NW;
    const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>(want, (long long)sm_count() * 16));
    kern<<<grid, NW * 32, smem, stream>>>(q);
    return (int)cudaGetLastError();
}

// out_channels 32 / 64 / 128 (lane l owns channels l, l+32, ...); point features 4..8 as the fused path
int launch_vfe(const VfeParams &q, bool abs_xyz, bool dist, cudaStream_t s) {
    if (q.C != 32 && q.C != 64 && q.C != 128) return HGSF_ERR_UNSUPPORTED;
    if (q.M == 0) return HGSF_OK;
#define HGSF_CASE(FV, A, D) if (q.F == FV && abs_xyz == A && dist == D) { \
        if (q.C == 64) return launch_vfe_t<FV, A, D, 64>(q, s); \
        if (q.C == 32) return launch_vfe_t<FV, A, D, 32>(q, s); \
        return launch_vfe_t<FV, A, D, 128>(q, s); }
    HGSF_CASE(4, true, false) HGSF_CASE(5, true, false) HGSF_CASE(6, true, false) HGSF_CASE(7, true, false)
    HGSF_CASE(8, true, false) HGSF_CASE(7, false, false) HGSF_CASE(8, false, false)
    HGSF_CASE(7, true, true) HGSF_CASE(8, true, true)
#undef HGSF_CASE
    return HGSF_ERR_UNSUPPORTED;
}

// ------------------------------------------------------------------------------------------------
// Stacked PFN, two layers (NUM_FILTERS = [2H, C1], pillar_vfe.py:63-74): layer 0 = Linear(Cin -> H) + BN + ReLU per slot and
// its max over ALL P slots (a padded slot is a zero row: relu(BN(0)) per channel), concatenated per slot as [h(s), max]
// (PFNLayer.forward :47-49); layer 1 = Linear(2H -> C1) + BN + ReLU per slot, max over all P slots (a padded slot's input is
// [relu(BN0(0)), max0]: evaluated once per pillar).  A warp per pillar: lane l owns channels l, l+32, ...; layer 0's rows live in
// registers (pfn.cuh), layer 1's weights transposed in shared memory ([2H][C1]: the lanes read consecutive words), its inputs are
// shared-memory broadcasts.  Sequential fp32 FMA in input-channel order.
template <int F, bool ABS, bool DIST, int H, int NWARPS>
__global__ void __launch_bounds__(NWARPS * 32) k_vfe_stacked(const VfeParams q) {
    using Lane = PfnLane<F, ABS, DIST, H>;
    constexpr int CPL = Lane::CPL;
    constexpr int K1 = 2 * H;
    constexpr int MAXC1 = 4;                              // C1 <= 128
    extern __shared__ float s_all[];
    const int C1 = q.C1, cpl1 = C1 >> 5;
    float *w1t = s_all;                                   // [K1][C1]
    float *bn1 = w1t + (size_t)K1 * C1;                   // [4][C1]: mean, invstd, gamma, beta (or 0, 0, 0, bias)
    float *per_warp = bn1 + 4 * (size_t)C1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *buf = per_warp + (size_t)warp * ((size_t)q.P * F + (size_t)(q.P + 1) * K1);     // voxel rows [P][F]
    float *hin = buf + (size_t)q.P * F;                   // layer-1 inputs [P + 1][2H] (row P = the padded slot)
    const bool bn1_on = q.pfn1.bn_w != nullptr;
    for (int t = threadIdx.x; t < K1 * C1; t += NWARPS * 32) {
        const int j = t / C1, c = t - j * C1;
        w1t[t] = __ldg(q.pfn1.W + (size_t)c * K1 + j);
    }
    for (int c = threadIdx.x; c < C1; c += NWARPS * 32) {
        if (bn1_on) {
            bn1[c] = __ldg(q.pfn1.bn_m + c);
            bn1[C1 + c] = __frcp_rn(__fsqrt_rn(__fadd_rn(__ldg(q.pfn1.bn_v + c), q.pfn1.eps)));
            bn1[2 * C1 + c] = __ldg(q.pfn1.bn_w + c);
            bn1[3 * C1 + c] = __ldg(q.pfn1.bn_b + c);
        } else {
            bn1[c] = 0.f; bn1[C1 + c] = 0.f; bn1[2 * C1 + c] = 0.f; bn1[3 * C1 + c] = __ldg(q.pfn1.bias + c);
        }
    }
    __syncthreads();
    Lane pfn;
    pfn.load(q.pfn, lane);
    const int P4 = (q.P >> 2) << 2;
    const long long nwarps = (long long)gridDim.x * NWARPS;
    for (long long m = (long long)blockIdx.x * NWARPS + warp; m < q.M; m += nwarps) {
        const float *src = q.voxels + (size_t)m * q.P * F;
        for (int t = lane; t < q.P * F; t += 32) buf[t] = __ldg(src + t);
        __syncwarp();
        const float nf = q.num_float ? __ldg(reinterpret_cast<const float *>(q.num) + m)
                                     : (float)__ldg(reinterpret_cast<const int32_t *>(q.num) + m);
        const int cnt = min(max((int)nf, 0), q.P);
        float cz_i, cy_i, cx_i;
        if (q.coords_float) {
            const float4 c = __ldg(reinterpret_cast<const float4 *>(q.coords) + m);
            cz_i = c.y; cy_i = c.z; cx_i = c.w;
        } else {
            const int4 c = __ldg(reinterpret_cast<const int4 *>(q.coords) + m);
            cz_i = (float)c.y; cy_i = (float)c.z; cx_i = (float)c.w;
        }
        SlotSum sum;
        for (int s = 0; s < q.P; ++s) sum.add(s, P4, buf[s * F], buf[s * F + 1], buf[s * F + 2]);
        const float mx = __fdiv_rn(sum.sx(), nf), my = __fdiv_rn(sum.sy(), nf), mz = __fdiv_rn(sum.sz(), nf);
        const float cx = __fadd_rn(__fmul_rn(cx_i, q.vsize[0]), q.voff[0]);
        const float cy = __fadd_rn(__fmul_rn(cy_i, q.vsize[1]), q.voff[1]);
        const float cz = __fadd_rn(__fmul_rn(cz_i, q.vsize[2]), q.voff[2]);
        // ---- layer 0: h(s) per slot into hin[s][0..H), its max over all P slots ----
        float vmax[CPL];
        pfn.init_max(vmax, cnt < q.P);
        for (int s = 0; s < cnt; ++s) {
            float rowf[F], y[CPL];
#pragma unroll
            for (int k = 0; k < F; ++k) rowf[k] = buf[s * F + k];
            pfn.value(rowf, mx, my, mz, cx, cy, cz, y);
#pragma unroll
            for (int j = 0; j < CPL; ++j) {
                hin[(size_t)s * K1 + lane + 32 * j] = y[j];
                vmax[j] = (y[j] > vmax[j] || y[j] != y[j]) ? y[j] : vmax[j];
            }
        }
#pragma unroll
        for (int j = 0; j < CPL; ++j) hin[(size_t)q.P * K1 + lane + 32 * j] = pfn.padv[j];      // the padded slot's h
        // the repeated max (second half of every slot's input)
        for (int s = 0; s <= cnt; ++s) {
            const int row = (s < cnt) ? s : q.P;
#pragma unroll
            for (int j = 0; j < CPL; ++j) hin[(size_t)row * K1 + H + lane + 32 * j] = vmax[j];
        }
        __syncwarp();
        // ---- layer 1: per slot, max over all P slots ----
        float omax[MAXC1];
#pragma unroll
        for (int t = 0; t < MAXC1; ++t) omax[t] = 0.f;
        const int n_eval = (cnt < q.P) ? cnt + 1 : cnt;           // + the padded slot, once
        for (int e = 0; e < n_eval; ++e) {
            const float *in = hin + (size_t)((e < cnt) ? e : q.P) * K1;
            float acc[MAXC1];
#pragma unroll
            for (int t = 0; t < MAXC1; ++t) acc[t] = 0.f;
            for (int j = 0; j < K1; ++j) {
                const float x = in[j];
#pragma unroll
                for (int t = 0; t < MAXC1; ++t)
                    if (t < cpl1) acc[t] = fmaf(x, w1t[(size_t)j * C1 + lane + 32 * t], acc[t]);
            }
#pragma unroll
            for (int t = 0; t < MAXC1; ++t) {
                if (t < cpl1) {
                    const int c = lane + 32 * t;
                    float y;
                    if (bn1_on) y = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(acc[t], bn1[c]), bn1[C1 + c]), bn1[2 * C1 + c]), bn1[3 * C1 + c]);
                    else        y = __fadd_rn(acc[t], bn1[3 * C1 + c]);
                    y = (y > 0.f || y != y) ? y : 0.f;
                    omax[t] = (y > omax[t] || y != y) ? y : omax[t];
                }
            }
        }
#pragma unroll
        for (int t = 0; t < MAXC1; ++t)
            if (t < cpl1) q.out[(size_t)m * C1 + lane + 32 * t] = omax[t];
        __syncwarp();
    }
}

template <int F, bool ABS, bool DIST, int H>
static int launch_vfe_stacked_t(const VfeParams &q, cudaStream_t stream) {
    constexpr int NW = 4;
    auto kern = k_vfe_stacked<F, ABS, DIST, H, NW>;
    const size_t smem = sizeof(float) * ((size_t)2 * H * q.C1 + 4 * (size_t)q.C1 + NW * ((size_t)q.P * F + (size_t)(q.P + 1) * 2 * H));
    if (smem > 200 * 1024) return HGSF_ERR_UNSUPPORTED;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    const long long want = (q.M + NW - 1) / NW;
    const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>(want, (long long)sm_count() * 4));
    kern<<<grid, NW * 32, smem, stream>>>(q);
    return (int)cudaGetLastError();
}

// first layer out_channels H = NUM_FILTERS[0] / 2 in {32, 64}; last layer out_channels C1 in {32, 64, 128}
int launch_vfe_stacked(const VfeParams &q, bool abs_xyz, bool dist, cudaStream_t s) {
    if ((q.C != 32 && q.C != 64) || (q.C1 != 32 && q.C1 != 64 && q.C1 != 128)) return HGSF_ERR_UNSUPPORTED;
    if (q.M == 0) return HGSF_OK;
#define HGSF_CASE(FV, A, D) if (q.F == FV && abs_xyz == A && dist == D) { \
        if (q.C == 32) return launch_vfe_stacked_t<FV, A, D, 32>(q, s); \
        return launch_vfe_stacked_t<FV, A, D, 64>(q, s); }
    HGSF_CASE(4, true, false) HGSF_CASE(7, true, false) HGSF_CASE(8, true, false) HGSF_CASE(7, false, false)
    HGSF_CASE(8, false, false) HGSF_CASE(7, true, true) HGSF_CASE(8, true, true)
#undef HGSF_CASE
    return HGSF_ERR_UNSUPPORTED;
}

// ------------------------------------------------------------------------------------------------
// cell -> (pillar row + 1); duplicates resolve to the last row as the CPU index_put does
__global__ void __launch_bounds__(256) k_scatter_map(const ScatterParams q) {
    const long long m = (long long)blockIdx.x * 256 + threadIdx.x;
    if (m >= q.M) return;
    int b;
    long long idx;
    if (q.coords_float) {
        const float4 c = __ldg(reinterpret_cast<const float4 *>(q.coords) + m);
        b = (int)c.x;
        // indices = c1 + c2 * nx + c3 evaluated in fp32 like the reference (pointpillar_scatter.py:31-32)
        idx = (long long)__fadd_rn(__fadd_rn(c.y, __fmul_rn(c.z, (float)q.nx)), c.w);
        if (!(c.x >= 0.f) || !(c.x < (float)q.B)) return;
    } else if (q.coord_cols == 3) {
        // SparseConvTensor indices (b, y, x) -- the .dense() of the PillarNet branch (lss_fpn.py:111-113)
        const int *c = reinterpret_cast<const int *>(q.coords) + 3 * m;
        const int y = __ldg(c + 1), x = __ldg(c + 2);
        b = __ldg(c);
        if (b < 0 || b >= q.B || y < 0 || y >= q.ny || x < 0 || x >= q.nx) return;
        idx = (long long)y * q.nx + x;
    } else {
        const int4 c = __ldg(reinterpret_cast<const int4 *>(q.coords) + m);
        b = c.x;
        idx = (long long)c.y + (long long)c.z * q.nx + c.w;
        if (b < 0 || b >= q.B) return;
    }
    if (idx < 0 || idx >= q.plane) return;
    atomicMax(q.map + (size_t)b * q.plane + idx, (unsigned)(m + 1));
}

template <int NWARPS, bool TMA>
__global__ void __launch_bounds__(NWARPS * 32) k_scatter_canvas(const __grid_constant__ CUtensorMap tmap, const ScatterParams q) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const int C = q.C, TILE = C * 32, NT = NWARPS * 32;
    // the TMA swizzle works on absolute shared-memory address bits: start the tiles on a 1024-byte boundary
    uint8_t *smem_al = smem_raw + ((1024u - ((uint32_t)__cvta_generic_to_shared(smem_raw) & 1023u)) & 1023u);
    float *tilebuf = reinterpret_cast<float *>(smem_al);    // [2][TILE]
    float *zerobuf = tilebuf + 2 * TILE;
    __shared__ int s_nocc[2], s_cell[2][32], s_m[2][32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int t = tid; t < TILE; t += NT) zerobuf[t] = 0.f;
    if (TMA) fence_proxy_async_smem();
    __syncthreads();
    const int tiles_per_row = (q.nx + 31) >> 5;
    const long long n_tiles = (long long)q.B * q.ny * tiles_per_row;
    auto load_entry = [&](long long t) -> unsigned {
        const int r = (int)(t / tiles_per_row);
        const int x = (int)(t - (long long)r * tiles_per_row) * 32 + lane;
        return (x < q.nx) ? __ldg(q.map + (size_t)r * q.nx + x) : 0u;
    };
    long long tile = blockIdx.x;
    unsigned e_next = 0;
    if (warp == 0 && tile < n_tiles) e_next = load_entry(tile);
    int nb = 0;
    for (int it = 0; tile < n_tiles; ++it, tile += gridDim.x) {
        const int slot = it & 1;
        const int r = (int)(tile / tiles_per_row);
        const int x0 = (int)(tile - (long long)r * tiles_per_row) * 32;
        const int b = r / q.ny, y = r - b * q.ny;
        if (warp == 0) {
            const unsigned e = e_next;
            const long long nt = tile + gridDim.x;
            if (nt < n_tiles) e_next = load_entry(nt);
            const unsigned bal = __ballot_sync(FULL, e != 0u);
            if (e) {
                const int k = __popc(bal & ((1u << lane) - 1u));
                s_cell[slot][k] = lane; s_m[slot][k] = (int)(e - 1u);
            }
            if (lane == 0) {
                s_nocc[slot] = __popc(bal);
                if (TMA && bal) tma_wait_read<1>();
            }
        }
        __syncthreads();
        const int n_occ = s_nocc[slot];
        if (n_occ == 0) {
            if (TMA) {
                if (tid == 0) { tma_store_3d(&tmap, zerobuf, x0, y, b * C); tma_commit(); }
            } else {
                for (int c = warp; c < C; c += NWARPS)
                    if (x0 + lane < q.nx) q.canvas[(((size_t)b * C + c) * q.ny + y) * q.nx + x0 + lane] = 0.f;
            }
            continue;
        }
        float *tb = tilebuf + (nb & 1) * TILE;
        ++nb;
        for (int t = tid * 4; t < TILE; t += NT * 4) *reinterpret_cast<float4 *>(tb + t) = make_float4(0.f, 0.f, 0.f, 0.f);
        __syncthreads();
        for (int k = warp; k < n_occ; k += NWARPS) {
            const int cell = s_cell[slot][k];
            const float *src = q.feats + (size_t)s_m[slot][k] * C;
            for (int c = lane; c < C; c += 32) tb[swz128(c, cell)] = __ldg(src + c);
        }
        if (TMA) {
            fence_proxy_async_smem();
            __syncthreads();
            if (tid == 0) { tma_store_3d(&tmap, tb, x0, y, b * C); tma_commit(); }
        } else {
            __syncthreads();
            for (int c = warp; c < C; c += NWARPS)
                if (x0 + lane < q.nx) q.canvas[(((size_t)b * C + c) * q.ny + y) * q.nx + x0 + lane] = tb[swz128(c, lane)];
        }
    }
    if (TMA && tid == 0) tma_wait_read<0>();
}

int launch_scatter(const ScatterParams &q, cudaStream_t stream, int *launches) {
    int nl = 0;
    cudaError_t e = cudaMemsetAsync(q.map, 0, sizeof(unsigned) * (size_t)q.B * q.plane, stream);   // not counted as a kernel
    if (e != cudaSuccess) return (int)e;
    if (q.M > 0) {
        k_scatter_map<<<(unsigned)((q.M + 255) / 256), 256, 0, stream>>>(q);
        if ((e = cudaGetLastError()) != cudaSuccess) return (int)e;
        ++nl;
    }
    constexpr int NW = 4;
    const bool tma = (q.nx % 4 == 0) && ((reinterpret_cast<uintptr_t>(q.canvas) & 15) == 0) && q.C <= 256;
    CUtensorMap map;
    memset(&map, 0, sizeof(map));
    if (tma) {
        const int st = make_canvas_map(&map, q.canvas, q.B, q.C, q.ny, q.nx, q.C);
        if (st != HGSF_OK) return st;
    }
    const size_t smem = 1024 + sizeof(float) * 3 * (size_t)q.C * 32;
    const long long n_tiles = (long long)q.B * q.ny * ((q.nx + 31) / 32);
    if (n_tiles > 0) {
        auto go = [&](auto kern) -> int {
            cudaError_t e2 = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e2 != cudaSuccess) return (int)e2;
            int per_sm = 1;
            e2 = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NW * 32, smem);
            if (e2 != cudaSuccess) return (int)e2;
            const long long grid = std::min<long long>(n_tiles, (long long)sm_count() * std::max(per_sm, 1));
            kern<<<(unsigned)grid, NW * 32, smem, stream>>>(map, q);
            return (int)cudaGetLastError();
        };
        const int st = tma ? go(k_scatter_canvas<NW, true>) : go(k_scatter_canvas<NW, false>);
        if (st != HGSF_OK) return st;
        ++nl;
    }
    if (launches) *launches = nl;
    return HGSF_OK;
}

}  // namespace hgsf
