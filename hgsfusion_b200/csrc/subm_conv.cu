// subm_conv.cu -- the pillar-list consumer, sm_100a.
//
// The PillarNet branch hands the reader's SparseConvTensor (pillar_modules.py:82) to SpMiddlePillarEncoder18, whose first
// stage `conv1` (pillarnet_modules/pcnres18.py:212-215) is five submanifold 3x3 convolutions 32 -> 32 (conv2D3x3 with
// stride 1 -> spconv.SubMConv2d, :82-95), each followed by BatchNorm1d(eps 1e-3) and, per block, a residual add and ReLU
// (Sparse2DBasicBlockV.forward :139-151, Sparse2DBasicBlock.forward :176-187).  A submanifold convolution keeps the
// active set: out[m] = bias + sum over the 3x3 taps of W[tap] . in[pillar at (y + ky - 1, x + kx - 1)], absent neighbours
// contributing nothing -- the dense cross-correlation with padding 1, evaluated at the active cells only.
//
//   k_subm_neighbors : one thread per (pillar, tap): the rule book [M, 9] from the reader's pillar_bev_indices table; built
//                      once per indice_key ("res1") and reused by all five convolutions, as spconv caches its indice pairs
//   k_subm_conv      : persistent CTAs of 128 threads; a tile = 64 pillars x Cout.  The whole weight tensor (9 x Cin x Cout
//                      fp32: 36 KB for 32 -> 32) stays in shared memory for the kernel's lifetime; per tap the 64 neighbour
//                      rows are gathered with cp.async (16-byte pieces, zero rows for absent neighbours; the neighbour ids
//                      are read one gather early so no dependent shared-memory load sits in front of the copies) into a
//                      three-stage ring, two taps ahead of the contraction, ONE barrier per tap.  A thread holds a
//                      4-pillar x 4-channel accumulator tile as packed pairs: one FFMA2 (fma.rn.f32x2) per channel pair,
//                      bit-identical to two fmaf in (tap, ci) order.  spconv's default arithmetic for fp32 features is
//                      fp32 (no TF32), so CUDA-core FMA is the faithful precision; the epilogue fuses bias, BatchNorm1d
//                      (eval), the residual add and the ReLU, so a block's activations cross HBM once.
//
// Bound: fp32 FMA (9 * Cin * Cout = 9 216 FMA per pillar for 32 -> 32 against 2 * 128 B of HBM traffic per pillar; the
// gathers hit L2).  Measured on B200 (config 2, 182 k pillars): 0.103 ms per convolution = 32.8 TFLOP/s, 44 % of the
// 74.5 TFLOP/s fp32 peak; ncu: shared-memory wavefronts at ~70 % of capacity -- every gathered 16-byte piece is re-read by
// the 8 threads sharing the pillar, every weight by the 16 thread rows -- so shared-memory bandwidth and FMA issue bound
// it about equally.  Thread tiles 4x2 / 8x2 / 8x4, tiles of 32 / 128 / 256 pillars and 2 / 4 stages measured within 10 %
// below this one; a row-owner mapping (a thread owns whole pillars, weights as warp-wide broadcast loads) was built and
// measured slower (26.8 TFLOP/s: 4 warps per SM at 166 registers).  [Cout,3,3,Cin] weights are transposed by every CTA
// on the fly (-8 %): convert to [3,3,Cin,Cout] once when the weights are frozen (the Python mirror does).
// 64 -> 64 (the weights take 147 KB of the SM's shared memory, one CTA per SM): 8-pillar x 4-channel thread tiles over 128-pillar
// tiles, 256 threads, 2 stages: 0.278 ms for the 143 k pillars of conv2's active set = 38.0 TFLOP/s, 51 % of the peak (4 x 4 over 64
// pillars 27.4; 3 stages 30.5; 16 x 4 35.6; 8 x 8 with 128 threads 32.2 -- too few warps).
// 128 and 256 channels (conv3 / conv4, pcnres18.py:227-245): 9 x Cin x Cout weights no longer fit shared memory (590 KB / 2.4 MB), so
// the output channels are cut into SLICES whose weights do (147 KB each: 64 of 64 -> 128, 32 of 128 -> {128, 256}, 16 of 256 -> 256);
// blockIdx.y = slice, every slice gathers the input rows again (L2 hits: these layers run on the stride-4 / stride-8 active sets,
// at most a few ten thousand pillars).  Same arithmetic (fp32 FMA in (tap, ci) order), same fused epilogue.
#include "subm_conv.cuh"

#include <algorithm>
#include <cstdlib>

#ifndef SUBM_UNROLL
#define SUBM_UNROLL 2
#endif

namespace hgsf {

namespace {

__global__ void __launch_bounds__(256) k_subm_neighbors(const SubmNeighborParams q) {
    const long long M = q.m_dev ? min((long long)max(q.m_dev[0], 0), q.M) : q.M;
    const long long total = M * 9;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long m = i / 9;
        const int k = (int)(i - m * 9);
        const int b = q.pillars[m * 3], y = q.pillars[m * 3 + 1], x = q.pillars[m * 3 + 2];
        const int yy = y * q.stride + k / 3 - 1, xx = x * q.stride + k % 3 - 1;
        int v = -1;
        if (b >= 0 && b < q.B && yy >= 0 && yy < q.H && xx >= 0 && xx < q.W) v = q.bev[((long long)b * q.H + yy) * q.W + xx];
        q.nbr[i] = v;
    }
}

// SparseConv2d(kernel 3, stride 2, padding 1): output cell (yo, xo) is active iff an input pillar lies in rows 2yo-1 .. 2yo+1 and
// columns 2xo-1 .. 2xo+1, i.e. an input pillar at an even coordinate reaches one output coordinate (y / 2) and at an odd one two
// ((y - 1) / 2 and (y + 1) / 2).  Every input pillar emits the centres of the (up to four) output cells it reaches as candidate
// "points"; hgsf_pillarnet_indices' machinery then turns them into the raster-ordered output pillar list and its cell table.
__global__ void __launch_bounds__(256) k_stride2_candidates(const Stride2CandidateParams q) {
    const long long M = q.m_dev ? min((long long)max(q.m_dev[0], 0), q.M) : q.M;
    const long long m_pad = (q.M + 31) & ~31ll;
    for (long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x; m < m_pad; m += (long long)gridDim.x * blockDim.x) {
        const bool live = m < M;
        int b = -1, y = 0, x = 0;
        if (live) { b = q.pillars[m * 3]; y = q.pillars[m * 3 + 1]; x = q.pillars[m * 3 + 2]; }
        const bool ok = live && b >= 0 && b < q.B && y >= 0 && x >= 0;
        if (m < q.M) {
            const int y0 = (y & 1) ? (y - 1) / 2 : y / 2, x0 = (x & 1) ? (x - 1) / 2 : x / 2;
            const int y1 = (y & 1) ? (y + 1) / 2 : -1, x1 = (x & 1) ? (x + 1) / 2 : -1;
            const int ys[2] = {y0, y1}, xs[2] = {x0, x1};
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                const int yo = ys[s >> 1], xo = xs[s & 1];
                const bool v = ok && yo >= 0 && yo < q.Ho && xo >= 0 && xo < q.Wo;
                float *c = q.cand + (m * 4 + s) * 3;
                c[0] = v ? (float)xo + 0.5f : -10.f;
                c[1] = v ? (float)yo + 0.5f : -10.f;
                c[2] = 0.f;
            }
        }
        // 4 candidate slots per pillar of frame b: one atomic per distinct frame in the warp
        const unsigned peers = __match_any_sync(FULL, ok ? b : -1);
        if (ok && (threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(q.cnt + b, 4 * __popc(peers));
    }
}

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// PT pillars x NCO output channels per thread; a tile = TM pillars x COUT channels, (COUT / NCO) * (TM / PT) threads
template <int CIN, int COUT, int PT, int NCO, int TM, int NBUF>
struct SubmCfg {
    static constexpr int TXN = COUT / NCO;                   // threads along the channels
    static constexpr int THREADS = TXN * (TM / PT);
    static constexpr int AS = CIN + 4;                       // padded row of the gathered tile
    static constexpr size_t w_bytes = sizeof(float) * 9 * CIN * COUT;
    static constexpr size_t a_bytes = sizeof(float) * NBUF * TM * AS;
    static constexpr size_t n_bytes = sizeof(int) * TM * 9;
    static constexpr size_t total = w_bytes + a_bytes + n_bytes;
};

template <int CIN, int COUT, int PT, int NCO, int TM, int NBUF>
__global__ void __launch_bounds__(SubmCfg<CIN, COUT, PT, NCO, TM, NBUF>::THREADS) k_subm_conv(const SubmConvParams q) {
    using S = SubmCfg<CIN, COUT, PT, NCO, TM, NBUF>;
    constexpr int AS = S::AS, THREADS = S::THREADS, TXN = S::TXN;
    constexpr int C4 = CIN / 4;               // 16-byte pieces per row
    constexpr int NP = NCO / 2;               // channel pairs per thread (packed fp32: one FFMA2 per pair)
    static_assert(NCO == 2 || NCO == 4 || NCO == 8, "channel pairs");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *Ws = reinterpret_cast<float *>(smem_raw);                        // [9][CIN][COUT]
    float *As = reinterpret_cast<float *>(smem_raw + S::w_bytes);           // [NBUF][TM][AS]
    int *Ns = reinterpret_cast<int *>(smem_raw + S::w_bytes + S::a_bytes);  // [TM][9]

    const int tid = threadIdx.x;
    const long long M = q.m_dev ? min((long long)max(q.m_dev[0], 0), q.M) : q.M;
    const long long n_tiles = (M + TM - 1) / TM;
    if ((long long)blockIdx.x >= n_tiles) return;
    // COUT is the width of this CTA's slice of the q.Cout output channels: channels co0 .. co0 + COUT - 1
    const int CT = q.Cout, co0 = (int)blockIdx.y * COUT;

    // the slice's weights, once per CTA, into [tap][ci][co]
    if (q.layout == 1) {
        if (CT == COUT) {
            for (int i = tid; i < 9 * CIN * COUT / 4; i += THREADS)
                reinterpret_cast<float4 *>(Ws)[i] = __ldg(reinterpret_cast<const float4 *>(q.W) + i);
        } else {
            for (int i = tid; i < 9 * CIN * COUT / 4; i += THREADS) {      // rows of COUT floats out of rows of CT
                const int r = i / (COUT / 4), c = i - r * (COUT / 4);
                reinterpret_cast<float4 *>(Ws)[i] = __ldg(reinterpret_cast<const float4 *>(q.W + (size_t)r * CT + co0) + c);
            }
        }
    } else {
        for (int i = tid; i < 9 * CIN * COUT; i += THREADS) {      // source order [co][tap][ci]: coalesced reads
            const int ci = i % CIN, k = (i / CIN) % 9, co = i / (9 * CIN);
            Ws[(k * CIN + ci) * COUT + co] = __ldg(q.W + (size_t)co0 * 9 * CIN + i);
        }
    }

    const int tx = tid % TXN, ty = tid / TXN;   // channels tx*NCO .. +NCO-1, pillars ty*PT .. +PT-1 of the tile
    float bnm[NCO], bni[NCO], bnw[NCO], bnb[NCO], bia[NCO];
#pragma unroll
    for (int c = 0; c < NCO; ++c) {
        const int co = co0 + tx * NCO + c;
        bia[c] = q.bias ? q.bias[co] : 0.f;
        if (q.bn_w) {
            bnm[c] = q.bn_m[co];
            bni[c] = __fdiv_rn(1.f, __fsqrt_rn(__fadd_rn(q.bn_v[co], q.eps)));
            bnw[c] = q.bn_w[co];
            bnb[c] = q.bn_b[co];
        } else { bnm[c] = 0.f; bni[c] = 1.f; bnw[c] = 1.f; bnb[c] = 0.f; }
    }

    constexpr int UNR = SUBM_UNROLL;
    constexpr int CH = TM * C4 / THREADS;      // 16-byte pieces a thread copies per tap
    static_assert(TM * C4 % THREADS == 0, "whole pieces per thread");
    int nbn[CH];                               // neighbour ids of the NEXT gather, read one gather early (no dependent LDS in front of the copies)
    auto load_nb = [&](int k) {
#pragma unroll
        for (int i = 0; i < CH; ++i) nbn[i] = Ns[((tid + i * THREADS) / C4) * 9 + k];
    };
    auto gather = [&](int k, int buf) {
        float *dst = As + (size_t)buf * TM * AS;
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            const int c = tid + i * THREADS;
            const int row = c / C4, col = c - row * C4;
            float *d = dst + row * AS + col * 4;
            if (nbn[i] >= 0) cp_async16(d, q.in + (size_t)nbn[i] * CIN + col * 4);
            else *reinterpret_cast<float4 *>(d) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (k + 1 < 9) load_nb(k + 1);
    };

    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long row0 = tile * TM;
        __syncthreads();                       // the previous tile's buffers and rule-book slice are free; Ws is written
        for (int i = tid; i < TM * 9; i += THREADS) {
            const long long g = row0 * 9 + i;
            Ns[i] = (g < M * 9) ? q.nbr[g] : -1;
        }
        __syncthreads();
        uint64_t acc[PT][NP];
#pragma unroll
        for (int i = 0; i < PT; ++i)
#pragma unroll
            for (int c = 0; c < NP; ++c) acc[i][c] = 0ull;       // (+0.f, +0.f)

        // NBUF-stage pipeline over the taps, one barrier per tap: the gather of tap k + NBUF - 1 is issued right after the
        // barrier that proves every warp has finished tap k - 1, into the buffer tap k - 1 used
        load_nb(0);
#pragma unroll
        for (int t = 0; t < NBUF - 1; ++t) { gather(t, t); cp_commit(); }
#pragma unroll 1
        for (int k = 0; k < 9; ++k) {
            cp_wait<NBUF - 2>();               // this thread's pieces of tap k have landed
            __syncthreads();                   // ... and everybody's; tap k - 1 is done with its buffer
            if (k + NBUF - 1 < 9) gather(k + NBUF - 1, (k + NBUF - 1) % NBUF);
            cp_commit();                       // (an empty group at the end keeps the group count uniform)
            const float *a0 = As + (size_t)(k % NBUF) * TM * AS + (ty * PT) * AS;
            const float *wk = Ws + (size_t)k * CIN * COUT + tx * NCO;
#pragma unroll UNR
            for (int c4 = 0; c4 < C4; ++c4) {
                float4 a[PT];
#pragma unroll
                for (int i = 0; i < PT; ++i) a[i] = *reinterpret_cast<const float4 *>(a0 + i * AS + c4 * 4);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint64_t w[NP];
                    if (NP == 1) w[0] = *reinterpret_cast<const uint64_t *>(wk + (c4 * 4 + j) * COUT);
                    else {
#pragma unroll
                        for (int c = 0; c < NP; c += 2) {
                            const ulonglong2 t = *reinterpret_cast<const ulonglong2 *>(wk + (c4 * 4 + j) * COUT + c * 2);
                            w[c] = t.x; w[c + 1 < NP ? c + 1 : c] = t.y;
                        }
                    }
#pragma unroll
                    for (int i = 0; i < PT; ++i) {
                        const float av = (j == 0) ? a[i].x : (j == 1) ? a[i].y : (j == 2) ? a[i].z : a[i].w;
                        const uint64_t a2 = pack_f2(av, av);
#pragma unroll
                        for (int c = 0; c < NP; ++c) acc[i][c] = fma2_rn(a2, w[c], acc[i][c]);   // two fmaf, bit-identical
                    }
                }
            }
        }

        // epilogue: bias, BatchNorm1d (eval), residual, ReLU
#pragma unroll
        for (int i = 0; i < PT; ++i) {
            const long long row = row0 + ty * PT + i;
            if (row >= M) continue;
            float v[NCO];
#pragma unroll
            for (int c = 0; c < NP; ++c) unpack_f2(acc[i][c], v[2 * c], v[2 * c + 1]);
#pragma unroll
            for (int c = 0; c < NCO; ++c) {
                float t = __fadd_rn(v[c], bia[c]);
                if (q.bn_w) t = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(t, bnm[c]), bni[c]), bnw[c]), bnb[c]);
                v[c] = t;
            }
            const size_t o = (size_t)row * CT + co0 + tx * NCO;
            if (q.residual) {
#pragma unroll
                for (int c = 0; c < NCO; ++c) v[c] = __fadd_rn(v[c], q.residual[o + c]);
            }
            if (q.relu) {
#pragma unroll
                for (int c = 0; c < NCO; ++c) v[c] = (v[c] > 0.f || v[c] != v[c]) ? v[c] : 0.f;   // NaN propagates like torch's ReLU
            }
            if (NCO == 2) *reinterpret_cast<float2 *>(q.out + o) = make_float2(v[0], v[1]);
            else {
#pragma unroll
                for (int c = 0; c + 3 < NCO; c += 4) *reinterpret_cast<float4 *>(q.out + o + c) = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
            }
        }
    }
}

// COUT = slice width (== q.Cout for the layers whose weights fit shared memory whole)
template <int CIN, int COUT, int PT, int NCO, int TM, int NBUF>
int launch_conv_t(const SubmConvParams &q, cudaStream_t stream) {
    using S = SubmCfg<CIN, COUT, PT, NCO, TM, NBUF>;
    auto kern = k_subm_conv<CIN, COUT, PT, NCO, TM, NBUF>;
    const int slices = q.Cout / COUT;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::total);
    if (e != cudaSuccess) return (int)e;
    int dev = 0, sms = 0, per_sm = 0;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return (int)e;
    if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return (int)e;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, S::THREADS, S::total)) != cudaSuccess) return (int)e;
    if (per_sm < 1) return HGSF_ERR_UNSUPPORTED;
    const long long tiles = (q.M + TM - 1) / TM;
    long long grid = tiles < (long long)sms * per_sm ? tiles : (long long)sms * per_sm;
    if (slices > 1) grid = std::max<long long>(1, std::min<long long>(tiles, ((long long)sms * per_sm + slices - 1) / slices));
    kern<<<dim3((unsigned)grid, (unsigned)slices), S::THREADS, S::total, stream>>>(q);
    return (int)cudaGetLastError();
}

}  // namespace

int launch_subm_neighbors(const SubmNeighborParams &q, cudaStream_t stream) {
    if (q.M == 0) return HGSF_OK;
    const long long total = q.M * 9;
    long long grid = (total + 255) / 256;
    if (grid > 148 * 16) grid = 148 * 16;
    k_subm_neighbors<<<(unsigned)grid, 256, 0, stream>>>(q);
    return (int)cudaGetLastError();
}

int launch_stride2_candidates(const Stride2CandidateParams &q, cudaStream_t stream) {
    if (q.M == 0) return HGSF_OK;
    long long grid = (q.M + 255) / 256;
    if (grid > 148 * 8) grid = 148 * 8;
    k_stride2_candidates<<<(unsigned)grid, 256, 0, stream>>>(q);
    return (int)cudaGetLastError();
}

int launch_subm_conv(const SubmConvParams &q, cudaStream_t stream) {
    if (q.M == 0) return HGSF_OK;
    if (q.Cin == 32 && q.Cout == 32) {
#ifdef HGSF_EXPERIMENT
        const char *v = getenv("HGSF_SUBM_VARIANT");
        const int var = v ? atoi(v) : 0;
        if (var == 1) return launch_conv_t<32, 32, 8, 2, 128, 2>(q, stream);
        if (var == 2) return launch_conv_t<32, 32, 8, 2, 128, 3>(q, stream);
        if (var == 3) return launch_conv_t<32, 32, 4, 4, 64, 2>(q, stream);
        if (var == 4) return launch_conv_t<32, 32, 4, 4, 64, 4>(q, stream);
        if (var == 5) return launch_conv_t<32, 32, 4, 2, 64, 3>(q, stream);
        if (var == 6) return launch_conv_t<32, 32, 8, 4, 128, 3>(q, stream);
        if (var == 7) return launch_conv_t<32, 32, 4, 4, 32, 3>(q, stream);
        if (var == 8) return launch_conv_t<32, 32, 8, 8, 128, 3>(q, stream);
        if (var == 9) return launch_conv_t<32, 32, 8, 8, 256, 3>(q, stream);
#endif
        return launch_conv_t<32, 32, 4, 4, 64, 3>(q, stream);
    }
    if (q.Cin == 64 && q.Cout == 64) {
#ifdef HGSF_EXPERIMENT
        const char *v = getenv("HGSF_SUBM_VARIANT");
        const int var = v ? atoi(v) : 0;
        if (var == 1) return launch_conv_t<64, 64, 4, 4, 64, 2>(q, stream);
        if (var == 2) return launch_conv_t<64, 64, 4, 4, 64, 3>(q, stream);
        if (var == 3) return launch_conv_t<64, 64, 8, 4, 64, 3>(q, stream);
        if (var == 4) return launch_conv_t<64, 64, 4, 4, 32, 4>(q, stream);
        if (var == 5) return launch_conv_t<64, 64, 8, 8, 128, 2>(q, stream);
        if (var == 6) return launch_conv_t<64, 64, 8, 8, 64, 3>(q, stream);
        if (var == 7) return launch_conv_t<64, 64, 16, 4, 128, 2>(q, stream);
#endif
        return launch_conv_t<64, 64, 8, 4, 128, 2>(q, stream);
    }
    if (q.Cin == 32 && q.Cout == 64) return launch_conv_t<32, 64, 4, 4, 64, 3>(q, stream);
    // conv3 / conv4: output channels in slices whose weights fit shared memory (see the header)
    if (q.Cin == 64 && q.Cout == 128) return launch_conv_t<64, 64, 8, 4, 128, 2>(q, stream);
    if (q.Cin == 128 && (q.Cout == 128 || q.Cout == 256)) return launch_conv_t<128, 32, 4, 4, 64, 2>(q, stream);
    if (q.Cin == 256 && q.Cout == 256) return launch_conv_t<256, 16, 2, 2, 32, 2>(q, stream);
    return HGSF_ERR_UNSUPPORTED;
}

}  // namespace hgsf
