// pfn.cuh -- per-lane PillarVFE arithmetic shared by the fused path (k_emit) and the batch_dict
// contract kernel (k_vfe).  A lane owns output channels lane, lane+32, ... and keeps their Linear
// rows and BatchNorm constants in registers; the Cin contraction is CUDA-core FMA (13x64 / 14x64 is
// far below anything a tensor core tile could use).
//
// Operation order = the order measured on the reference's CPU path (oracle/pillar_oracle.c):
//   Linear   acc = fmaf(feat[k], W[c][k], acc), k ascending           pillar_vfe.py:37
//   BN eval  (((acc - mean) * invstd) * gamma) + beta, 4 roundings     pillar_vfe.py:39
//   ReLU, max over the pillar's slots; NaN propagates                   pillar_vfe.py:41-42
#pragma once
#include "common.cuh"

namespace hgsf {

struct PfnArgs {
    const float *W, *bias, *bn_w, *bn_b, *bn_m, *bn_v;
    float eps;
};

template <int F, bool ABS, bool DIST, int C>
struct PfnLane {
    static constexpr int CIN = (ABS ? F : F - 3) + 6 + (DIST ? 1 : 0);
    static constexpr int CPL = C / 32;
    float w[CPL][CIN];
    float bn_m[CPL], bn_i[CPL], bn_g[CPL], bn_b[CPL], padv[CPL];
    bool has_bn;

    __device__ __forceinline__ void load(const PfnArgs &a, int lane) {
        has_bn = (a.bn_w != nullptr);
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
            const int c = lane + 32 * j;
#pragma unroll
            for (int k = 0; k < CIN; ++k) w[j][k] = __ldg(a.W + c * CIN + k);
            float y;
            if (has_bn) {
                bn_m[j] = __ldg(a.bn_m + c);
                bn_i[j] = __frcp_rn(__fsqrt_rn(__fadd_rn(__ldg(a.bn_v + c), a.eps)));
                bn_g[j] = __ldg(a.bn_w + c);
                bn_b[j] = __ldg(a.bn_b + c);
                // a zero (padded) row still goes through BN + ReLU and joins the max (pillar_vfe.py:37-42)
                y = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(0.f, bn_m[j]), bn_i[j]), bn_g[j]), bn_b[j]);
            } else {
                bn_m[j] = 0.f; bn_i[j] = 0.f; bn_g[j] = 0.f;
                bn_b[j] = __ldg(a.bias + c);
                y = __fadd_rn(0.f, bn_b[j]);
            }
            padv[j] = (y > 0.f || y != y) ? y : 0.f;
        }
    }

    __device__ __forceinline__ void init_max(float (&vmax)[CPL], bool has_padding) const {
#pragma unroll
        for (int j = 0; j < CPL; ++j) vmax[j] = has_padding ? padv[j] : 0.f;
    }

    // one point, NOT folded into a maximum: y[j] = relu(BN(Linear(decorated row))) of this lane's channels (the input of the next
    // layer of a stacked PFN, pillar_vfe.py:47-49)
    template <int RW>
    __device__ __forceinline__ void value(const float (&row)[RW], float mx, float my, float mz,
                                          float cx, float cy, float cz, float (&yout)[CPL]) const {
        float feat[CIN];
        int kf = 0;
#pragma unroll
        for (int q = ABS ? 0 : 3; q < F; ++q) feat[kf++] = row[q];
        feat[kf++] = __fsub_rn(row[0], mx); feat[kf++] = __fsub_rn(row[1], my); feat[kf++] = __fsub_rn(row[2], mz);
        feat[kf++] = __fsub_rn(row[0], cx); feat[kf++] = __fsub_rn(row[1], cy); feat[kf++] = __fsub_rn(row[2], cz);
        if (DIST) feat[kf++] = __fsqrt_rn(fmaf(row[2], row[2], fmaf(row[1], row[1], __fmul_rn(row[0], row[0]))));
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
            float acc = 0.f;
#pragma unroll
            for (int q = 0; q < CIN; ++q) acc = fmaf(feat[q], w[j][q], acc);
            float y;
            if (has_bn) y = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(acc, bn_m[j]), bn_i[j]), bn_g[j]), bn_b[j]);
            else        y = __fadd_rn(acc, bn_b[j]);
            yout[j] = (y > 0.f || y != y) ? y : 0.f;
        }
    }

    // one point: row[0..F) raw features (x, y, z first), pillar mean and centre
    template <int RW>
    __device__ __forceinline__ void point(const float (&row)[RW], float mx, float my, float mz,
                                          float cx, float cy, float cz, float (&vmax)[CPL]) const {
        float feat[CIN];
        int kf = 0;
#pragma unroll
        for (int q = ABS ? 0 : 3; q < F; ++q) feat[kf++] = row[q];
        feat[kf++] = __fsub_rn(row[0], mx); feat[kf++] = __fsub_rn(row[1], my); feat[kf++] = __fsub_rn(row[2], mz);
        feat[kf++] = __fsub_rn(row[0], cx); feat[kf++] = __fsub_rn(row[1], cy); feat[kf++] = __fsub_rn(row[2], cz);
        // torch.norm(xyz, 2, 2) on the CPU: sqrt(fma(z,z, fma(y,y, x*x)))  (pillar_vfe.py:110-112)
        if (DIST) feat[kf++] = __fsqrt_rn(fmaf(row[2], row[2], fmaf(row[1], row[1], __fmul_rn(row[0], row[0]))));
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
            float acc = 0.f;
#pragma unroll
            for (int q = 0; q < CIN; ++q) acc = fmaf(feat[q], w[j][q], acc);
            float y;
            if (has_bn) y = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(acc, bn_m[j]), bn_i[j]), bn_g[j]), bn_b[j]);
            else        y = __fadd_rn(acc, bn_b[j]);
            y = (y > 0.f || y != y) ? y : 0.f;
            vmax[j] = (y > vmax[j] || y != y) ? y : vmax[j];
        }
    }
};

// torch CPU sum(dim=1) over the P slots: four interleaved partial sums over the leading 4*floor(P/4)
// slots, the tail added into partial 0, combined as ((a0+a1)+a2)+a3  (pillar_vfe.py:97)
struct SlotSum {
    float ax[4] = {0.f, 0.f, 0.f, 0.f}, ay[4] = {0.f, 0.f, 0.f, 0.f}, az[4] = {0.f, 0.f, 0.f, 0.f};
    __device__ __forceinline__ void add(int s, int P4, float x, float y, float z) {
        const int q = (s < P4) ? (s & 3) : 0;
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (q == u) { ax[u] = __fadd_rn(ax[u], x); ay[u] = __fadd_rn(ay[u], y); az[u] = __fadd_rn(az[u], z); }
    }
    __device__ __forceinline__ float sx() const { return __fadd_rn(__fadd_rn(__fadd_rn(ax[0], ax[1]), ax[2]), ax[3]); }
    __device__ __forceinline__ float sy() const { return __fadd_rn(__fadd_rn(__fadd_rn(ay[0], ay[1]), ay[2]), ay[3]); }
    __device__ __forceinline__ float sz() const { return __fadd_rn(__fadd_rn(__fadd_rn(az[0], az[1]), az[2]), az[3]); }
};

}  // namespace hgsf
