// Micro-benchmark: how fast can a [B,64,ny,nx] fp32 canvas be written with 16-byte stores in different orders?
//  pattern 0: linear (memset-like)                         -- every warp instruction writes 512 contiguous bytes
//  pattern 1: tile order: warp writes a 32-cell x 64-channel tile = 64 pieces of 128 B, 409600 B apart (k_canvas today)
//  pattern 2: strip order: warp writes 8 channel rows of one BEV row (8 x 1280 B contiguous runs)
//  pattern 3: tile order but 4 consecutive x-tiles per warp (64 x 512 B pieces)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_write(float4 *out, int B, int C, int ny, int nx, int pattern) {
    const int lane = threadIdx.x & 31, gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    const size_t plane4 = (size_t)ny * nx / 4, row4 = nx / 4;
    if (pattern == 0) {
        const size_t n4 = (size_t)B * C * plane4;
        for (size_t i = (size_t)gw * 32 + lane; i < n4; i += (size_t)nw * 32) __stcs(out + i, z);
    } else if (pattern == 1) {
        const int tpr = nx / 32; const long long tiles = (long long)B * ny * tpr;
        for (long long t = gw; t < tiles; t += nw) {
            const int r = (int)(t / tpr), xt = (int)(t - (long long)r * tpr); const int b = r / ny, y = r - b * ny;
            float4 *dst = out + ((size_t)b * C + (lane >> 3)) * plane4 + (size_t)y * row4 + xt * 8 + (lane & 7);
            for (int i = 0; i < C / 4; ++i) __stcs(dst + (size_t)i * 4 * plane4, z);
        }
    } else if (pattern == 2) {
        const long long strips = (long long)B * ny * (C / 8);
        for (long long s = gw; s < strips; s += nw) {
            const int cb = (int)(s % (C / 8)); const long long r = s / (C / 8); const int b = (int)(r / ny), y = (int)(r - (long long)b * ny);
            for (int ch = 0; ch < 8; ++ch) {
                float4 *dst = out + ((size_t)b * C + cb * 8 + ch) * plane4 + (size_t)y * row4;
                for (int x = lane; x < row4; x += 32) __stcs(dst + x, z);
            }
        }
    } else {
        const int tpr = nx / 128; const long long tiles = (long long)B * ny * tpr;
        for (long long t = gw; t < tiles; t += nw) {
            const int r = (int)(t / tpr), xt = (int)(t - (long long)r * tpr); const int b = r / ny, y = r - b * ny;
            float4 *dst = out + (size_t)b * C * plane4 + (size_t)y * row4 + xt * 32 + lane;
            for (int c = 0; c < C; ++c) __stcs(dst + (size_t)c * plane4, z);
        }
    }
}
int main() {
    const int B = 16, C = 64, ny = 320, nx = 384;   // nx multiple of 128 so that every pattern covers the canvas
    const size_t bytes = (size_t)B * C * ny * nx * 4;
    float4 *d; cudaMalloc(&d, bytes);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int threads : {128, 256}) for (int per_sm : {1, 2, 3, 4, 8, 16}) for (int pat = 0; pat < 4; ++pat) {
        const int grid = 148 * per_sm;
        for (int i = 0; i < 3; ++i) k_write<<<grid, threads>>>(d, B, C, ny, nx, pat);
        cudaEventRecord(e0);
        for (int i = 0; i < 20; ++i) k_write<<<grid, threads>>>(d, B, C, ny, nx, pat);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 20;
        printf("threads %3d ctas/sm %2d pattern %d: %7.1f us  %7.1f GB/s\n", threads, per_sm, pat, ms * 1e3, bytes / ms / 1e6);
    }
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
