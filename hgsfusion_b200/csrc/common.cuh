// common.cuh -- shared device helpers for the pillar path (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/hgsfusion_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "hgsfusion_b200 is written for sm_100a (B200) only"
#endif

namespace hgsf {

constexpr uint32_t FULL = 0xffffffffu;

// One entry per BEV cell (b, z, y, x) of the direct-address table -- the perfect hash of a
// dense grid.  16 B so a tile of 32 cells is 512 contiguous bytes.
//   after k_count : tag = 0xFFFFFFFF - (smallest point index in the cell), cnt = points in the cell
//   after k_scan  : tag = raw pillar id + 1 (first-seen rank over the whole batch), start = CSR offset
//   tag == 0      : empty cell
struct __align__(16) CellEntry {
    uint32_t tag;
    uint32_t cnt;
    uint32_t start;
    uint32_t pad;
};

__host__ __device__ inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---- TMA (bulk tensor) store helpers ------------------------------------------------------
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap *map, const void *smem, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(map), "r"((uint32_t)__cvta_generic_to_shared(smem)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tma_store_3d_hint(const CUtensorMap *map, const void *smem, int c0, int c1, int c2, uint64_t policy) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3, %4}], [%1], %5;"
                 ::"l"(map), "r"((uint32_t)__cvta_generic_to_shared(smem)), "r"(c0), "r"(c1), "r"(c2), "l"(policy)
                 : "memory");
}
// L2 eviction policies: the canvas is written once and never re-read by this path (evict first); the pillar rows are
// re-read by k_canvas right after k_pfn wrote them (evict last)
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_normal() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void st_f4_hint(float *ptr, const float4 v, uint64_t policy) {
    asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(ptr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(policy)
                 : "memory");
}
__device__ __forceinline__ float4 ld_f4_hint(const float *ptr, uint64_t policy) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(ptr), "l"(policy));
    return r;
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}

// ---- streaming loads / stores ---------------------------------------------------------------
__device__ __forceinline__ float4 ld_nc_f4(const float *p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint64_t ld_volatile_u64(const uint64_t *p) {
    uint64_t v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_u64(uint64_t *p, uint64_t v) {
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// ---- packed fp32 pairs (sm_100 FFMA2 / FMUL2 / FADD2): two independent IEEE round-to-nearest operations per
// instruction, bit-identical to the scalar forms.  ptxas 12.9 contracts mul.rn.f32x2 -> add.rn.f32x2 into FFMA2 even
// under -fmad=false (scalar .rn ops are left alone), so a rounded multiply must never feed a PACKED add: the callers
// finish such chains with scalar __fadd_rn.
__device__ __forceinline__ uint64_t pack_f2(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack_f2(uint64_t v, float &lo, float &hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t fma2_rn(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ uint64_t sub2_rn(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t mul2_rn(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

// 128-byte TMA swizzle: 16-byte chunk index (address bits 4..6) XOR row index (bits 7..9).
// `row` = 128-byte row of a 1024-byte aligned tile, `col` = float column 0..31.
__device__ __forceinline__ int swz128(int row, int col) {
    return row * 32 + ((((col >> 2) ^ (row & 7)) << 2) | (col & 3));
}

}  // namespace hgsf
