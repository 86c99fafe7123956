// train_ops.cu -- training through PillarVFE and PointPillarScatter, on the batch_dict contract
// layout (voxels [M,P,F], voxel_coords [M,4], voxel_num_points [M]).
//
// What the reference does in train mode (pcdet/models/backbones_3d/vfe/pillar_vfe.py:29-49): x = Linear(features)
// [M,P,C]; BatchNorm1d(eps 1e-3, momentum 0.01) with BATCH statistics over all N = M*P rows of each channel -- the
// zero-padded rows (x = 0) included; ReLU; max over P.  Autograd then routes d(pillar_features) to the arg-max row,
// through ReLU, the batch-norm backward (which couples every row of a channel) and the Linear.
//
//   k_vfe_stats     one pass over the real rows: per channel  Sx = sum x, Sxx = sum x^2, T[c,k] = sum x[c]*feat[k];
//                   s[k] = sum feat[k].  Padded rows add nothing to any of them (x = 0, feat = 0); they count in N.
//   k_bn_finalize   mean = Sx/N, var = Sxx/N - mean^2 (biased, what the forward normalises with); running stats
//                   <- (1-momentum)*running + momentum*{mean, var*N/(N-1)}  (torch BatchNorm1d)
//   (forward)       hgsf_pillar_vfe with bn_mean/bn_var = the batch statistics: the eval kernel, same arithmetic form
//   k_vfe_backward  per (pillar, channel): recompute the rows, find the arg-max (first maximum, real rows before the
//                   padding as torch.max does); with g = d out[m,c] and y* > 0:
//                   dBeta[c] += g, dGammaRaw[c] += g*xhat*, A[c,:] += g*feat(arg-max row)
//   k_vfe_combine   batch statistics:  dW[c,k] = gamma*invstd*( A - dBeta/N * s[k] - dGamma/N * invstd*(T - mean*s[k]) )
//                   running statistics (eval-mode BN, frozen): dW = gamma*invstd*A
//                   no BN (USE_NORM False): dW = A, dBias = sum g
//   k_scatter_grad  d pillar_features[m, c] = d spatial_features[b, c, y, x]   (PointPillarScatter backward: a gather)
//
// Sums are accumulated per thread in fp64 and merged with fp64 atomics (B200 keeps a usable FP64 pipe): the variance
// and the (T - mean*s) term are differences of large sums.
#include "train_ops.cuh"

#include <algorithm>

namespace hgsf {

namespace {

template <int F, bool ABS, bool DIST>
struct Deco {
    static constexpr int CIN = (ABS ? F : F - 3) + 6 + (DIST ? 1 : 0);
    // decorated features of one point (pillar_vfe.py:94-118), same roundings as pfn.cuh
    __device__ static __forceinline__ void feat(const float *row, float mx, float my, float mz, float cx, float cy, float cz,
                                                float (&f)[CIN]) {
        int kf = 0;
#pragma unroll
        for (int q = ABS ? 0 : 3; q < F; ++q) f[kf++] = row[q];
        f[kf++] = __fsub_rn(row[0], mx); f[kf++] = __fsub_rn(row[1], my); f[kf++] = __fsub_rn(row[2], mz);
        f[kf++] = __fsub_rn(row[0], cx); f[kf++] = __fsub_rn(row[1], cy); f[kf++] = __fsub_rn(row[2], cz);
        if (DIST) f[kf++] = __fsqrt_rn(fmaf(row[2], row[2], fmaf(row[1], row[1], __fmul_rn(row[0], row[0]))));
    }
};

// what every pillar needs before its rows: the staged voxel, its point count, slot mean and centre
struct PillarHead {
    int cnt;
    float mx, my, mz, cx, cy, cz;
};

__device__ __forceinline__ void cp_async4(void *smem_dst, const void *gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// count, slot mean and centre of pillar m, whose voxel [P, F] is staged in buf
template <int F>
__device__ __forceinline__ PillarHead pillar_head(const VfeParams &q, long long m, const float *buf) {
    PillarHead h;
    const float nf = q.num_float ? __ldg(reinterpret_cast<const float *>(q.num) + m)
                                 : (float)__ldg(reinterpret_cast<const int32_t *>(q.num) + m);
    h.cnt = min(max((int)nf, 0), q.P);
    float cz_i, cy_i, cx_i;
    if (q.coords_float) {
        const float4 c = __ldg(reinterpret_cast<const float4 *>(q.coords) + m);
        cz_i = c.y; cy_i = c.z; cx_i = c.w;
    } else {
        const int4 c = __ldg(reinterpret_cast<const int4 *>(q.coords) + m);
        cz_i = (float)c.y; cy_i = (float)c.z; cx_i = (float)c.w;
    }
    const int P4 = (q.P >> 2) << 2;
    SlotSum sum;
    for (int s = 0; s < q.P; ++s) sum.add(s, P4, buf[s * F], buf[s * F + 1], buf[s * F + 2]);
    h.mx = __fdiv_rn(sum.sx(), nf); h.my = __fdiv_rn(sum.sy(), nf); h.mz = __fdiv_rn(sum.sz(), nf);
    h.cx = __fadd_rn(__fmul_rn(cx_i, q.vsize[0]), q.voff[0]);
    h.cy = __fadd_rn(__fmul_rn(cy_i, q.vsize[1]), q.voff[1]);
    h.cz = __fadd_rn(__fmul_rn(cz_i, q.vsize[2]), q.voff[2]);
    return h;
}

constexpr int TR_WARPS = 8;
constexpr int TR_C = 64;

// block-level merge of per-warp partial sums, then one fp64 atomic per CTA and address
template <int NV>
__device__ __forceinline__ void merge_and_add(double (&v)[NV], double *s_red, double *dst, int stride_lane, int lane, int warp,
                                              bool lane_active) {
    // s_red: [TR_WARPS][32] doubles; one value index at a time keeps the shared footprint small
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        __syncthreads();
        s_red[warp * 32 + lane] = v[i];
        __syncthreads();
        if (warp == 0 && lane_active) {
            double t = 0.0;
#pragma unroll
            for (int w = 0; w < TR_WARPS; ++w) t += s_red[w * 32 + lane];
            if (t != 0.0) atomicAdd(dst + (size_t)i * stride_lane + lane, t);
        }
    }
}

__device__ __forceinline__ long long pillar_count(const VfeParams &q) {
    return q.M_dev ? min((long long)max(__ldg(q.M_dev), 0), q.M) : q.M;
}

template <int F, bool ABS, bool DIST>
__global__ void __launch_bounds__(TR_WARPS * 32) k_vfe_stats(const VfeParams q, double *stats) {
    using D = Deco<F, ABS, DIST>;
    constexpr int CIN = D::CIN;
    extern __shared__ __align__(16) uint8_t tr_raw[];
    double *s_red = reinterpret_cast<double *>(tr_raw);                          // [TR_WARPS*32]
    float *s_vox = reinterpret_cast<float *>(tr_raw + sizeof(double) * TR_WARPS * 32);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int PF = q.P * F;
    float *buf0 = s_vox + (size_t)warp * 2 * PF;                 // two staging buffers (as in k_vfe_backward)
    float w[2][CIN];
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int k = 0; k < CIN; ++k) w[j][k] = __ldg(q.pfn.W + (lane + 32 * j) * CIN + k);
    double sx[2] = {0.0, 0.0}, sxx[2] = {0.0, 0.0}, T[2 * CIN], sf[CIN];
#pragma unroll
    for (int k = 0; k < 2 * CIN; ++k) T[k] = 0.0;
#pragma unroll
    for (int k = 0; k < CIN; ++k) sf[k] = 0.0;
    const long long nwarps = (long long)gridDim.x * TR_WARPS;
    const long long M = pillar_count(q);
    auto issue = [&](long long m, float *b) {
        if (m < M) {
            const float *src = q.voxels + (size_t)m * PF;
            for (int t = lane; t < PF; t += 32) cp_async4(b + t, src + t);
        }
        cp_async_commit();
    };
    int cur = 0;
    issue((long long)blockIdx.x * TR_WARPS + warp, buf0);
    for (long long m = (long long)blockIdx.x * TR_WARPS + warp; m < M; m += nwarps, cur ^= 1) {
        const float *buf = buf0 + cur * PF;
        issue(m + nwarps, buf0 + (cur ^ 1) * PF);            // the next pillar's voxel arrives under this one's rows
        cp_async_wait<1>();
        __syncwarp();
        const PillarHead h = pillar_head<F>(q, m, buf);
        for (int s = 0; s < h.cnt; ++s) {
            float f[CIN];
            D::feat(buf + s * F, h.mx, h.my, h.mz, h.cx, h.cy, h.cz, f);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                float acc = 0.f;
#pragma unroll
                for (int k = 0; k < CIN; ++k) acc = fmaf(f[k], w[j][k], acc);     // the forward's Linear, same order
                const double x = (double)acc;
                sx[j] += x; sxx[j] += x * x;
#pragma unroll
                for (int k = 0; k < CIN; ++k) T[j * CIN + k] += x * (double)f[k];
            }
#pragma unroll
            for (int k = 0; k < CIN; ++k) sf[k] += (double)f[k];
        }
        __syncwarp();
    }
    cp_async_wait<0>();
    // layout of stats: Sx [C], Sxx [C], T [C][CIN], s [CIN]
    {
        double v[2];
        v[0] = sx[0]; v[1] = sx[1];
        merge_and_add<2>(v, s_red, stats, 32, lane, warp, true);                 // channel lane + 32 j -> index j*32 + lane
        v[0] = sxx[0]; v[1] = sxx[1];
        merge_and_add<2>(v, s_red, stats + TR_C, 32, lane, warp, true);
    }
    // T[c][k] at stats[2C + c*CIN + k]: lane's channel c = lane + 32 j
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int k = 0; k < CIN; ++k) {
            __syncthreads();
            s_red[warp * 32 + lane] = T[j * CIN + k];
            __syncthreads();
            if (warp == 0) {
                double t = 0.0;
#pragma unroll
                for (int ww = 0; ww < TR_WARPS; ++ww) t += s_red[ww * 32 + lane];
                if (t != 0.0) atomicAdd(stats + 2 * TR_C + (size_t)(lane + 32 * j) * CIN + k, t);
            }
        }
    // s[k]: every lane of a warp holds the same sums; lane 0 of each warp contributes
    __syncthreads();
    if (lane < CIN) {
        double mine = 0.0;
#pragma unroll
        for (int k = 0; k < CIN; ++k) if (k == lane) mine = sf[k];
        s_red[warp * 32 + lane] = mine;
    }
    __syncthreads();
    if (warp == 0 && lane < CIN) {
        double t = 0.0;
#pragma unroll
        for (int ww = 0; ww < TR_WARPS; ++ww) t += s_red[ww * 32 + lane];
        if (t != 0.0) atomicAdd(stats + 2 * TR_C + TR_C * CIN + lane, t);
    }
}

__global__ void k_bn_finalize(const double *stats, double n_rows, int C, float momentum, float *running_mean, float *running_var,
                              float *batch_mean, float *batch_var) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const double mean = stats[c] / n_rows;
    double var = stats[C + c] / n_rows - mean * mean;
    if (var < 0.0) var = 0.0;
    batch_mean[c] = (float)mean;
    batch_var[c] = (float)var;
    if (running_mean) running_mean[c] = (float)((1.0 - (double)momentum) * (double)running_mean[c] + (double)momentum * mean);
    if (running_var) {
        const double unbiased = n_rows > 1.0 ? var * n_rows / (n_rows - 1.0) : var;
        running_var[c] = (float)((1.0 - (double)momentum) * (double)running_var[c] + (double)momentum * unbiased);
    }
}

template <int F, bool ABS, bool DIST, bool BN>
// (two CTAs per SM at 128 registers with 56 B of spills: 0.414 vs 0.426 ms for the backward of config 2 -- not worth the spills)
__global__ void __launch_bounds__(TR_WARPS * 32) k_vfe_backward(const VfeParams q, const float *__restrict__ grad_out, double *acc) {
    using D = Deco<F, ABS, DIST>;
    constexpr int CIN = D::CIN;
    extern __shared__ __align__(16) uint8_t tr_raw[];
    double *s_red = reinterpret_cast<double *>(tr_raw);
    float *s_vox = reinterpret_cast<float *>(tr_raw + sizeof(double) * TR_WARPS * 32);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int PF = q.P * F;
    float *buf0 = s_vox + (size_t)warp * 2 * PF;                 // two staging buffers: the next pillar's voxel arrives under this one's rows
    float w[2][CIN], mu[2], iv[2], ga[2], be[2], ypad[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int c = lane + 32 * j;
#pragma unroll
        for (int k = 0; k < CIN; ++k) w[j][k] = __ldg(q.pfn.W + c * CIN + k);
        if (BN) {
            mu[j] = __ldg(q.pfn.bn_m + c);
            iv[j] = __frcp_rn(__fsqrt_rn(__fadd_rn(__ldg(q.pfn.bn_v + c), q.pfn.eps)));
            ga[j] = __ldg(q.pfn.bn_w + c);
            be[j] = __ldg(q.pfn.bn_b + c);
            ypad[j] = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(0.f, mu[j]), iv[j]), ga[j]), be[j]);
        } else {
            mu[j] = 0.f; iv[j] = 1.f; ga[j] = 1.f;
            be[j] = __ldg(q.pfn.bias + c);
            ypad[j] = __fadd_rn(0.f, be[j]);
        }
    }
    double A[2 * CIN], dB[2] = {0.0, 0.0}, dG[2] = {0.0, 0.0};
#pragma unroll
    for (int k = 0; k < 2 * CIN; ++k) A[k] = 0.0;
    const long long nwarps = (long long)gridDim.x * TR_WARPS;
    const long long M = pillar_count(q);
    auto issue = [&](long long m, float *b) {
        if (m < M) {
            const float *src = q.voxels + (size_t)m * PF;
            for (int t = lane; t < PF; t += 32) cp_async4(b + t, src + t);
        }
        cp_async_commit();
    };
    int cur = 0;
    issue((long long)blockIdx.x * TR_WARPS + warp, buf0);
    for (long long m = (long long)blockIdx.x * TR_WARPS + warp; m < M; m += nwarps, cur ^= 1) {
        const float *buf = buf0 + cur * PF;
        issue(m + nwarps, buf0 + (cur ^ 1) * PF);
        // this pillar's cotangents: independent of its rows, so requested before them
        const float g_in[2] = {__ldg(grad_out + (size_t)m * TR_C + lane), __ldg(grad_out + (size_t)m * TR_C + lane + 32)};
        cp_async_wait<1>();
        __syncwarp();
        const PillarHead h = pillar_head<F>(q, m, buf);
        float best_y[2], best_x[2];
        int best_s[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) { best_s[j] = -1; best_y[j] = 0.f; best_x[j] = 0.f; }
        for (int s = 0; s < h.cnt; ++s) {
            float f[CIN];
            D::feat(buf + s * F, h.mx, h.my, h.mz, h.cx, h.cy, h.cz, f);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                float x = 0.f;
#pragma unroll
                for (int k = 0; k < CIN; ++k) x = fmaf(f[k], w[j][k], x);
                const float y = BN ? __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(x, mu[j]), iv[j]), ga[j]), be[j]) : __fadd_rn(x, be[j]);
                if (best_s[j] < 0 || y > best_y[j]) { best_s[j] = s; best_y[j] = y; best_x[j] = x; }   // first maximum
            }
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            // the padded rows come after the real ones in the slot order: they win only when strictly larger
            if (h.cnt < q.P && (best_s[j] < 0 || ypad[j] > best_y[j])) { best_s[j] = -2; best_y[j] = ypad[j]; best_x[j] = 0.f; }
            if (best_s[j] == -1 || !(best_y[j] > 0.f)) continue;                  // ReLU gate (threshold_backward: x > 0)
            const double g = (double)g_in[j];
            dB[j] += g;
            if (BN) dG[j] += g * (double)__fmul_rn(__fsub_rn(best_x[j], mu[j]), iv[j]);
            if (best_s[j] >= 0) {
                float f[CIN];
                D::feat(buf + best_s[j] * F, h.mx, h.my, h.mz, h.cx, h.cy, h.cz, f);
#pragma unroll
                for (int k = 0; k < CIN; ++k) A[j * CIN + k] += g * (double)f[k];
            }
        }
        __syncwarp();
    }
    cp_async_wait<0>();
    // layout of acc: A [C][CIN], dGammaRaw [C], dBeta [C]
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int k = 0; k < CIN; ++k) {
            __syncthreads();
            s_red[warp * 32 + lane] = A[j * CIN + k];
            __syncthreads();
            if (warp == 0) {
                double t = 0.0;
#pragma unroll
                for (int ww = 0; ww < TR_WARPS; ++ww) t += s_red[ww * 32 + lane];
                if (t != 0.0) atomicAdd(acc + (size_t)(lane + 32 * j) * CIN + k, t);
            }
        }
    merge_and_add<2>(dG, s_red, acc + TR_C * CIN, 32, lane, warp, true);
    merge_and_add<2>(dB, s_red, acc + TR_C * CIN + TR_C, 32, lane, warp, true);
}

// mode 0: batch statistics (train);  1: running statistics (frozen BN);  2: no BN (Linear with bias)
__global__ void k_vfe_combine(const double *acc, const double *stats, double n_rows, const int32_t *M_dev, long long cap, int P,
                              int C, int CIN, int mode, const float *gamma, const float *bn_mean, const float *bn_var, float eps,
                              float *grad_weight, float *grad_gamma, float *grad_beta) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= C * CIN) return;
    if (M_dev) n_rows = (double)min((long long)max(__ldg(M_dev), 0), cap) * (double)P;
    const int c = t / CIN, k = t - c * CIN;
    const double A = acc[t], dG = acc[C * CIN + c], dB = acc[C * CIN + C + c];
    double dW;
    if (mode == 2) {
        dW = A;
    } else {
        const double invstd = 1.0 / sqrt((double)bn_var[c] + (double)eps);
        const double gi = (double)gamma[c] * invstd;
        if (mode == 1) {
            dW = gi * A;
        } else {
            const double s = stats[2 * C + C * CIN + k], T = stats[2 * C + (size_t)c * CIN + k], mean = stats[c] / n_rows;
            dW = gi * (A - dB / n_rows * s - dG / n_rows * invstd * (T - mean * s));
        }
    }
    grad_weight[t] = (float)dW;
    if (k == 0) {
        if (grad_gamma && mode != 2) grad_gamma[c] = (float)dG;
        if (grad_beta) grad_beta[c] = (float)dB;
    }
}

__global__ void __launch_bounds__(256) k_scatter_grad(const float *__restrict__ grad_canvas, const void *coords, int coords_float,
                                                      long long M, const int32_t *M_dev, int C, int B, int ny, int nx,
                                                      float *__restrict__ grad_feats, const float *__restrict__ add) {
    const int lane = threadIdx.x & 31;
    if (M_dev) M = min((long long)max(__ldg(M_dev), 0), M);
    const long long w0 = ((long long)blockIdx.x * 256 + threadIdx.x) >> 5, nw = ((long long)gridDim.x * 256) >> 5;
    for (long long m = w0; m < M; m += nw) {
        int b, y, x;
        if (coords_float) {
            const float4 c = __ldg(reinterpret_cast<const float4 *>(coords) + m);
            b = (int)c.x; y = (int)c.z; x = (int)c.w;
        } else {
            const int4 c = __ldg(reinterpret_cast<const int4 *>(coords) + m);
            b = c.x; y = c.z; x = c.w;
        }
        const bool ok = grad_canvas && b >= 0 && b < B && y >= 0 && y < ny && x >= 0 && x < nx;
        const float *src = grad_canvas + (((size_t)(ok ? b : 0) * C) * ny + (ok ? y : 0)) * nx + (ok ? x : 0);
        for (int c = lane; c < C; c += 32) {
            float g = ok ? __ldg(src + (size_t)c * ny * nx) : 0.f;
            if (add) g = __fadd_rn(g, __ldg(add + (size_t)m * C + c));       // a cotangent on pillar_features itself
            grad_feats[(size_t)m * C + c] = g;
        }
    }
}

template <typename K, typename... Extra>
int launch_rows(K kern, const VfeParams &q, cudaStream_t s, Extra... extra) {
    const size_t smem = sizeof(double) * TR_WARPS * 32 + sizeof(float) * TR_WARPS * 2 * (size_t)q.P * q.F;   // (two voxel buffers per warp)
    if (smem > 200 * 1024) return HGSF_ERR_UNSUPPORTED;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    const long long want = (q.M + TR_WARPS - 1) / TR_WARPS;
    const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>(want, (long long)sm_count() * 2));
    kern<<<grid, TR_WARPS * 32, smem, s>>>(q, extra...);
    return (int)cudaGetLastError();
}

}  // namespace

int launch_vfe_stats(const VfeParams &q, bool abs_xyz, bool dist, double *stats, cudaStream_t s) {
    if (q.C != TR_C) return HGSF_ERR_UNSUPPORTED;
    const int cin = (abs_xyz ? q.F : q.F - 3) + 6 + (dist ? 1 : 0);
    cudaError_t e = cudaMemsetAsync(stats, 0, sizeof(double) * train_stats_len(q.C, cin), s);
    if (e != cudaSuccess) return (int)e;
    if (q.M == 0) return HGSF_OK;
#define HGSF_CASE(FV, A, D) if (q.F == FV && abs_xyz == A && dist == D) return launch_rows(k_vfe_stats<FV, A, D>, q, s, stats);
    HGSF_CASE(4, true, false) HGSF_CASE(5, true, false) HGSF_CASE(6, true, false) HGSF_CASE(7, true, false)
    HGSF_CASE(8, true, false) HGSF_CASE(7, false, false) HGSF_CASE(8, false, false)
    HGSF_CASE(7, true, true) HGSF_CASE(8, true, true)
#undef HGSF_CASE
    return HGSF_ERR_UNSUPPORTED;
}

int launch_bn_finalize(const double *stats, double n_rows, int C, float momentum, float *running_mean, float *running_var,
                       float *batch_mean, float *batch_var, cudaStream_t s) {
    k_bn_finalize<<<(C + 63) / 64, 64, 0, s>>>(stats, n_rows, C, momentum, running_mean, running_var, batch_mean, batch_var);
    return (int)cudaGetLastError();
}

int launch_vfe_backward(const VfeParams &q, bool abs_xyz, bool dist, const float *grad_out, const double *stats, int mode,
                        double *acc, float *grad_weight, float *grad_gamma, float *grad_beta, cudaStream_t s, int *launches) {
    if (q.C != TR_C) return HGSF_ERR_UNSUPPORTED;
    const int cin = (abs_xyz ? q.F : q.F - 3) + 6 + (dist ? 1 : 0);
    cudaError_t e = cudaMemsetAsync(acc, 0, sizeof(double) * train_acc_len(q.C, cin), s);
    if (e != cudaSuccess) return (int)e;
    int st = HGSF_OK, nl = 0;
    if (q.M > 0) {
        st = HGSF_ERR_UNSUPPORTED;
        const bool bn = mode != 2;
#define HGSF_CASE(FV, A, D)                                                                                     \
    if (q.F == FV && abs_xyz == A && dist == D)                                                                  \
        st = bn ? launch_rows(k_vfe_backward<FV, A, D, true>, q, s, grad_out, acc)                               \
                : launch_rows(k_vfe_backward<FV, A, D, false>, q, s, grad_out, acc);
        HGSF_CASE(4, true, false) HGSF_CASE(5, true, false) HGSF_CASE(6, true, false) HGSF_CASE(7, true, false)
        HGSF_CASE(8, true, false) HGSF_CASE(7, false, false) HGSF_CASE(8, false, false)
        HGSF_CASE(7, true, true) HGSF_CASE(8, true, true)
#undef HGSF_CASE
        if (st != HGSF_OK) return st;
        ++nl;
    }
    const int n = q.C * cin;
    k_vfe_combine<<<(n + 127) / 128, 128, 0, s>>>(acc, stats, (double)q.M * q.P, q.M_dev, q.M, q.P, q.C, cin, mode, q.pfn.bn_w,
                                                  q.pfn.bn_m, q.pfn.bn_v, q.pfn.eps, grad_weight, grad_gamma, grad_beta);
    ++nl;
    if (launches) *launches = nl;
    return (int)cudaGetLastError();
}

int launch_scatter_grad(const float *grad_canvas, const void *coords, int coords_float, long long M, int C, int B, int ny, int nx,
                        float *grad_feats, cudaStream_t s, const int32_t *M_dev, const float *add) {
    if (M == 0) return HGSF_OK;
    const long long blocks = (M * 32 + 255) / 256;
    k_scatter_grad<<<(unsigned)std::max<long long>(1, std::min<long long>(blocks, 148 * 32)), 256, 0, s>>>(
        grad_canvas, coords, coords_float, M, M_dev, C, B, ny, nx, grad_feats, add);
    return (int)cudaGetLastError();
}

}  // namespace hgsf
