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
//   k_subm_conv      : persistent CTAs; a tile = 64 pillars x Cout.  The whole weight tensor (9 x Cin x Cout fp32: 36 KB
//                      for 32 -> 32) stays in shared memory for the kernel's lifetime; per tap the 64 neighbour rows are
//                      gathered with cp.async (16-byte pieces, zero rows for absent neighbours) into a double buffer
//                      while the previous tap is being contracted; a thread holds a 4-pillar x (Cout/16)-channel
//                      accumulator tile.  fp32 FMA on the CUDA cores in (tap, ci) order: spconv's default arithmetic
//                      for fp32 features is fp32 (no TF32), so this is the faithful precision; the epilogue fuses bias,
//                      BatchNorm1d (eval), the residual add and the ReLU, so a block's activations cross HBM once.
//
// Bound: fp32 FMA issue (9 * Cin * Cout FMA per pillar = 9 216 for 32 -> 32 against 2 * 128 B of HBM traffic per pillar,
// the gathers hit L2).
#include "subm_conv.cuh"

namespace hgsf {

namespace {

constexpr int SUBM_THREADS = 256;
constexpr int SUBM_TM = 64;      // pillars per tile

__global__ void __launch_bounds__(256) k_subm_neighbors(const SubmNeighborParams q) {
    const long long M = q.m_dev ? min((long long)max(q.m_dev[0], 0), q.M) : q.M;
    const long long total = M * 9;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long m = i / 9;
        const int k = (int)(i - m * 9);
        const int b = q.pillars[m * 3], y = q.pillars[m * 3 + 1], x = q.pillars[m * 3 + 2];
        const int yy = y + k / 3 - 1, xx = x + k % 3 - 1;
        int v = -1;
        if (b >= 0 && b < q.B && yy >= 0 && yy < q.H && xx >= 0 && xx < q.W) v = q.bev[((long long)b * q.H + yy) * q.W + xx];
        q.nbr[i] = v;
    }
}

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int CIN, int COUT>
struct SubmSmem {
    static constexpr int AS = CIN + 4;                       // padded row of the gathered tile
    static constexpr size_t w_bytes = sizeof(float) * 9 * CIN * COUT;
    static constexpr size_t a_bytes = sizeof(float) * 2 * SUBM_TM * AS;
    static constexpr size_t n_bytes = sizeof(int) * SUBM_TM * 9;
    static constexpr size_t total = w_bytes + a_bytes + n_bytes;
};

template <int CIN, int COUT>
__global__ void __launch_bounds__(SUBM_THREADS) k_subm_conv(const SubmConvParams q) {
    using S = SubmSmem<CIN, COUT>;
    constexpr int AS = S::AS;
    constexpr int NCO = COUT / 16;            // output channels per thread
    constexpr int C4 = CIN / 4;               // 16-byte pieces per row
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *Ws = reinterpret_cast<float *>(smem_raw);                        // [9][CIN][COUT]
    float *As = reinterpret_cast<float *>(smem_raw + S::w_bytes);           // [2][TM][AS]
    int *Ns = reinterpret_cast<int *>(smem_raw + S::w_bytes + S::a_bytes);  // [TM][9]

    const int tid = threadIdx.x;
    const long long M = q.m_dev ? min((long long)max(q.m_dev[0], 0), q.M) : q.M;
    const long long n_tiles = (M + SUBM_TM - 1) / SUBM_TM;
    if ((long long)blockIdx.x >= n_tiles) return;

    // the weights, once per CTA, into [tap][ci][co]
    if (q.layout == 1) {
        for (int i = tid; i < 9 * CIN * COUT / 4; i += SUBM_THREADS)
            reinterpret_cast<float4 *>(Ws)[i] = __ldg(reinterpret_cast<const float4 *>(q.W) + i);
    } else {
        for (int i = tid; i < 9 * CIN * COUT; i += SUBM_THREADS) {      // source order [co][tap][ci]: coalesced reads
            const int ci = i % CIN, k = (i / CIN) % 9, co = i / (9 * CIN);
            Ws[(k * CIN + ci) * COUT + co] = __ldg(q.W + i);
        }
    }

    const int tx = tid & 15, ty = tid >> 4;   // channels tx*NCO .. +NCO-1, pillars ty*4 .. +3 of the tile
    float bnm[NCO], bni[NCO], bnw[NCO], bnb[NCO], bia[NCO];
#pragma unroll
    for (int c = 0; c < NCO; ++c) {
        const int co = tx * NCO + c;
        bia[c] = q.bias ? q.bias[co] : 0.f;
        if (q.bn_w) {
            bnm[c] = q.bn_m[co];
            bni[c] = __fdiv_rn(1.f, __fsqrt_rn(__fadd_rn(q.bn_v[co], q.eps)));
            bnw[c] = q.bn_w[co];
            bnb[c] = q.bn_b[co];
        } else { bnm[c] = 0.f; bni[c] = 1.f; bnw[c] = 1.f; bnb[c] = 0.f; }
    }

    auto gather = [&](int k, int buf) {
        float *dst = As + (size_t)buf * SUBM_TM * AS;
#pragma unroll
        for (int c = tid; c < SUBM_TM * C4; c += SUBM_THREADS) {
            const int row = c / C4, col = c - row * C4;
            const int nb = Ns[row * 9 + k];
            float *d = dst + row * AS + col * 4;
            if (nb >= 0) cp_async16(d, q.in + (size_t)nb * CIN + col * 4);
            else *reinterpret_cast<float4 *>(d) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        cp_commit();
    };

    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long row0 = tile * SUBM_TM;
        __syncthreads();                       // the previous tile's buffers and rule-book slice are free; Ws is written
        for (int i = tid; i < SUBM_TM * 9; i += SUBM_THREADS) {
            const long long g = row0 * 9 + i;
            Ns[i] = (g < M * 9) ? q.nbr[g] : -1;
        }
        __syncthreads();
        float acc[4][NCO];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int c = 0; c < NCO; ++c) acc[i][c] = 0.f;

        gather(0, 0);
#pragma unroll 1
        for (int k = 0; k < 9; ++k) {
            if (k + 1 < 9) { gather(k + 1, (k + 1) & 1); cp_wait<1>(); } else cp_wait<0>();
            __syncthreads();
            const float *a0 = As + (size_t)(k & 1) * SUBM_TM * AS + (ty * 4) * AS;
            const float *wk = Ws + (size_t)k * CIN * COUT + tx * NCO;
#pragma unroll 2
            for (int c4 = 0; c4 < C4; ++c4) {
                float4 a[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4 *>(a0 + i * AS + c4 * 4);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float w[NCO];
                    if (NCO == 2) {
                        const float2 t = *reinterpret_cast<const float2 *>(wk + (c4 * 4 + j) * COUT);
                        w[0] = t.x; w[1] = t.y;
                    } else {
#pragma unroll
                        for (int c = 0; c < NCO; c += 4) {
                            const float4 t = *reinterpret_cast<const float4 *>(wk + (c4 * 4 + j) * COUT + c);
                            w[c] = t.x; w[c + 1] = t.y; w[c + 2] = t.z; w[c + 3] = t.w;
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float av = (j == 0) ? a[i].x : (j == 1) ? a[i].y : (j == 2) ? a[i].z : a[i].w;
#pragma unroll
                        for (int c = 0; c < NCO; ++c) acc[i][c] = fmaf(av, w[c], acc[i][c]);
                    }
                }
            }
            __syncthreads();                   // buffer k & 1 is refilled by the gather of tap k + 2
        }

        // epilogue: bias, BatchNorm1d (eval), residual, ReLU
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const long long row = row0 + ty * 4 + i;
            if (row >= M) continue;
            float v[NCO];
#pragma unroll
            for (int c = 0; c < NCO; ++c) {
                float t = __fadd_rn(acc[i][c], bia[c]);
                if (q.bn_w) t = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(t, bnm[c]), bni[c]), bnw[c]), bnb[c]);
                v[c] = t;
            }
            const size_t o = (size_t)row * COUT + tx * NCO;
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
                for (int c = 0; c < NCO; c += 4) *reinterpret_cast<float4 *>(q.out + o + c) = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
            }
        }
    }
}

template <int CIN, int COUT>
int launch_conv_t(const SubmConvParams &q, cudaStream_t stream) {
    using S = SubmSmem<CIN, COUT>;
    auto kern = k_subm_conv<CIN, COUT>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::total);
    if (e != cudaSuccess) return (int)e;
    int dev = 0, sms = 0, per_sm = 0;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return (int)e;
    if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return (int)e;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, SUBM_THREADS, S::total)) != cudaSuccess) return (int)e;
    if (per_sm < 1) return HGSF_ERR_UNSUPPORTED;
    const long long tiles = (q.M + SUBM_TM - 1) / SUBM_TM;
    const long long grid = tiles < (long long)sms * per_sm ? tiles : (long long)sms * per_sm;
    kern<<<(unsigned)grid, SUBM_THREADS, S::total, stream>>>(q);
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

int launch_subm_conv(const SubmConvParams &q, cudaStream_t stream) {
    if (q.M == 0) return HGSF_OK;
    if (q.Cin == 32 && q.Cout == 32) return launch_conv_t<32, 32>(q, stream);
    if (q.Cin == 64 && q.Cout == 64) return launch_conv_t<64, 64>(q, stream);
    if (q.Cin == 32 && q.Cout == 64) return launch_conv_t<32, 64>(q, stream);
    return HGSF_ERR_UNSUPPORTED;
}

}  // namespace hgsf
