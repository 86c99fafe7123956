// hybrid_points.cu -- the step in front of the pillar path on the device:
// raw radar sweep + real points inside instance masks + virtual (RHGM) points -> one collated [sum N', 1 + F] array with
// the two flag columns, NO_DUP filter, camera field-of-view filter, range mask -- stable (input order kept), no host sync.
//
// What it reproduces (file:line under the reference):
//   assembly + flags   pcdet/datasets/kitti/vod_dataset.py:498-522, pcdet/datasets/kitti/tj4d_dataset.py:588-610
//   NO_DUP             vod_dataset.py:13-19 (calc_dist), :511-514
//   FOV filter         vod_dataset.py:181-197,525-528; pcdet/utils/calibration_kitti.py:68-88
//   range mask         pcdet/utils/common_utils.py:78-81 (data_processor.py:83-85)
//   batch column       pcdet/datasets/dataset.py:237-244;  float32 cast  pcdet/models/__init__.py:23-36
// The reference does all of it in float64 numpy per sample inside DataLoader workers; the filters here are evaluated in
// double with separately rounded multiplies and adds in the same order (oracle/hybrid_oracle.py), so the kept set is identical.
//
// Two launches: k_hybrid_flags (1 thread / candidate row: keep flag + per-chunk counts) and k_hybrid_write (per chunk:
// carry-in = sum of the chunk counts before it, block scan, row copy).  Candidates are numbered frame-major:
// frame b = its sweep rows, then its mask rows, then its virtual rows.
#include "hybrid_points.cuh"

namespace hgsf {

constexpr int HY_THREADS = 256;

struct FrameSpan {
    int b, nr, ng, nv, r0, g0, v0;
    long long first;       // candidate index of the frame's first row
};

__device__ __forceinline__ long long cand_off(const HybridParams &q, int b) {
    long long v = (long long)__ldg(q.real_off + b);
    if (q.W) v += (long long)__ldg(q.gt_off + b) + (long long)__ldg(q.virt_off + b);
    return v;
}

// largest b in [0, B) with cand_off(b) <= g
__device__ __forceinline__ FrameSpan find_span(const HybridParams &q, long long g) {
    int lo = 0, hi = q.B;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (cand_off(q, mid) <= g) lo = mid; else hi = mid;
    }
    FrameSpan s;
    s.b = lo;
    s.r0 = __ldg(q.real_off + lo); s.nr = __ldg(q.real_off + lo + 1) - s.r0;
    s.g0 = s.v0 = s.ng = s.nv = 0;
    if (q.W) {
        s.g0 = __ldg(q.gt_off + lo); s.ng = __ldg(q.gt_off + lo + 1) - s.g0;
        s.v0 = __ldg(q.virt_off + lo); s.nv = __ldg(q.virt_off + lo + 1) - s.v0;
    }
    s.first = (long long)s.r0 + s.g0 + s.v0;
    return s;
}

// source row of candidate g: kind 0 = sweep, 1 = mask (gt_real), 2 = virtual
__device__ __forceinline__ const float *source_row(const HybridParams &q, const FrameSpan &s, long long g, int &kind) {
    const int local = (int)(g - s.first);
    if (local < s.nr) { kind = 0; return q.real + (size_t)(s.r0 + local) * q.Fr; }
    if (local < s.nr + s.ng) { kind = 1; return q.gt + (size_t)(s.g0 + local - s.nr) * q.W; }
    kind = 2;
    return q.virt + (size_t)(s.v0 + local - s.nr - s.ng) * q.W;
}

__device__ __forceinline__ bool keep_row(const HybridParams &q, const FrameSpan &s, const float *row, int kind) {
    // a frame without mask points keeps only its sweep (vod_dataset.py:507-509)
    if (q.W && s.ng == 0 && kind != 0) return false;
    const double x = (double)__ldg(row), y = (double)__ldg(row + 1), z = (double)__ldg(row + 2);
    if (q.no_dup && kind == 0 && s.ng > 0) {
        // calc_dist: min over the mask points of ((dx*dx + dy*dy) + dz*dz); kept iff |min| > threshold
        double best = __longlong_as_double(0x7ff0000000000000ll);
        bool any_nan = false;
        for (int j = 0; j < s.ng; ++j) {
            const float *gr = q.gt + (size_t)(s.g0 + j) * q.W;
            const double dx = __dsub_rn((double)__ldg(gr), x), dy = __dsub_rn((double)__ldg(gr + 1), y),
                         dz = __dsub_rn((double)__ldg(gr + 2), z);
            const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
            any_nan |= (d2 != d2);                   // numpy's min propagates NaN; |NaN| > t is False -> dropped
            best = fmin(best, d2);
        }
        if (any_nan || !(fabs(best) > q.dup_threshold)) return false;
    }
    if (q.calib) {
        const float *c = q.calib + (size_t)s.b * 26;        // 12: lidar->rect [4,3]; 12: P2 [3,4]; image h, w
        double rect[3], hom[3];
#pragma unroll
        for (int j = 0; j < 3; ++j)
            rect[j] = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(x, (double)__ldg(c + j)), __dmul_rn(y, (double)__ldg(c + 3 + j))),
                                          __dmul_rn(z, (double)__ldg(c + 6 + j))), (double)__ldg(c + 9 + j));
#pragma unroll
        for (int j = 0; j < 3; ++j)
            hom[j] = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(rect[0], (double)__ldg(c + 12 + 4 * j)),
                                                   __dmul_rn(rect[1], (double)__ldg(c + 13 + 4 * j))),
                                         __dmul_rn(rect[2], (double)__ldg(c + 14 + 4 * j))), (double)__ldg(c + 15 + 4 * j));
        const double u = __ddiv_rn(hom[0], rect[2]), v = __ddiv_rn(hom[1], rect[2]);
        const double depth = __dsub_rn(hom[2], (double)__ldg(c + 23));
        const double img_h = (double)__ldg(c + 24), img_w = (double)__ldg(c + 25);
        if (!(u >= 0.0 && u < img_w && v >= 0.0 && v < img_h && depth >= 0.0)) return false;
    }
    if (q.mask_range && !(x >= q.range_xy[0] && x <= q.range_xy[2] && y >= q.range_xy[1] && y <= q.range_xy[3])) return false;
    return true;
}

__device__ __forceinline__ int block_reduce_int(int v, int *s_warp) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(FULL, v, d);
    __syncthreads();
    if (lane == 0) s_warp[warp] = v;
    __syncthreads();
    int t = 0;
#pragma unroll
    for (int w = 0; w < HY_THREADS / 32; ++w) t += s_warp[w];
    return t;
}

__global__ void __launch_bounds__(HY_THREADS) k_hybrid_flags(const HybridParams q) {
    __shared__ int s_warp[HY_THREADS / 32];
    const long long g = (long long)blockIdx.x * HY_THREADS + threadIdx.x;
    int keep = 0;
    if (g < q.n) {
        const FrameSpan s = find_span(q, g);
        int kind;
        const float *row = source_row(q, s, g, kind);
        keep = keep_row(q, s, row, kind) ? 1 : 0;
        q.flags[g] = (unsigned char)keep;
    }
    const int total = block_reduce_int(keep, s_warp);
    if (threadIdx.x == 0) q.chunk_counts[blockIdx.x] = total;
}

__global__ void __launch_bounds__(HY_THREADS) k_hybrid_write(const HybridParams q) {
    __shared__ int s_warp[HY_THREADS / 32];
    __shared__ int s_scan[HY_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // carry-in: rows kept by all chunks before this one
    int before = 0;
    for (int c = tid; c < (int)blockIdx.x; c += HY_THREADS) before += q.chunk_counts[c];
    const int carry = block_reduce_int(before, s_warp);
    const long long g = (long long)blockIdx.x * HY_THREADS + tid;
    const int keep = (g < q.n) ? (int)q.flags[g] : 0;
    int incl = keep;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(FULL, incl, d);
        if (lane >= d) incl += o;
    }
    if (lane == 31) s_scan[warp] = incl;
    __syncthreads();
    int warp_off = 0;
#pragma unroll
    for (int w = 0; w < HY_THREADS / 32; ++w) warp_off += (w < warp) ? s_scan[w] : 0;
    const int pos = carry + warp_off + incl - keep;           // rows kept before candidate g
    if (g >= q.n) return;
    const FrameSpan s = find_span(q, g);
    // frame offsets of the output: the first candidate of a frame knows how many rows precede the frame
    if (g == s.first)
        for (int b = s.b; b >= 0 && cand_off(q, b) == g; --b) q.frame_offsets_out[b] = pos;
    if (g == q.n - 1)
        for (int b = q.B; b > s.b && cand_off(q, b) >= q.n; --b) q.frame_offsets_out[b] = pos + keep;
    if (!keep) return;
    int kind;
    const float *row = source_row(q, s, g, kind);
    const int Wout = q.W ? q.W + 2 : q.Fr;
    float *dst = q.out + (size_t)pos * (1 + Wout);
    dst[0] = (float)s.b;
    if (!q.W) {
        for (int k = 0; k < q.Fr; ++k) dst[1 + k] = __ldg(row + k);
        return;
    }
    if (kind == 0) {
        for (int k = 0; k < q.Fr; ++k) dst[1 + k] = __ldg(row + k);
        for (int k = q.Fr; k < q.W + 2; ++k) dst[1 + k] = 1.f;              // label columns and both flags stay 1 (np.ones)
    } else {
        for (int k = 0; k < q.W; ++k) dst[1 + k] = __ldg(row + k);
        dst[1 + q.W] = 0.f;
        // vod_dataset.py:521: points[-Nv:, -1] = 1 -- with Nv == 0 the slice is the whole array (mask rows get 1 too)
        dst[2 + q.W] = (kind == 2 || s.nv == 0) ? 1.f : 0.f;
    }
}

size_t hybrid_workspace_bytes(long long n) {
    const long long chunks = (n + HY_THREADS - 1) / HY_THREADS;
    return align_up((size_t)n, 256) + align_up(sizeof(int) * (size_t)(chunks + 1), 256);
}

int launch_hybrid(HybridParams q, void *ws, size_t ws_bytes, cudaStream_t s, int *launches) {
    *launches = 0;
    if (q.n == 0) {
        const cudaError_t e = cudaMemsetAsync(q.frame_offsets_out, 0, sizeof(int) * (size_t)(q.B + 1), s);
        return (int)e;
    }
    if (ws_bytes < hybrid_workspace_bytes(q.n) || (reinterpret_cast<uintptr_t>(ws) & 255)) return HGSF_ERR_WORKSPACE;
    q.flags = static_cast<unsigned char *>(ws);
    q.chunk_counts = reinterpret_cast<int *>(static_cast<unsigned char *>(ws) + align_up((size_t)q.n, 256));
    const unsigned chunks = (unsigned)((q.n + HY_THREADS - 1) / HY_THREADS);
    k_hybrid_flags<<<chunks, HY_THREADS, 0, s>>>(q);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    k_hybrid_write<<<chunks, HY_THREADS, 0, s>>>(q);
    e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    *launches = 2;
    return HGSF_OK;
}

}  // namespace hgsf
