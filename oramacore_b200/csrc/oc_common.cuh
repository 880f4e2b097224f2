// oc_common.cuh — shared device helpers: ordered score keys, bitonic sorts, mbarrier /
// bulk-copy (TMA 1-D) PTX wrappers.  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "oramacore_b200.h"

namespace oc {

// ---------------------------------------------------------------------------------------
// 64-bit rank keys.  Larger key == better hit: (order-preserving score bits << 32) | ~idx,
// so ties on score resolve to the LOWER index (rows are stored in ascending DocumentId
// order, matching the oracle's "ties by ascending doc id").  Keys are unique per index.
// NaN scores never become keys (callers test `score == score`), mirroring
// NotNan::new(..) => continue in top_n (read/sort.rs:264-267).
// ---------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t f32_ordered(float f) {
#ifdef __CUDA_ARCH__
    uint32_t u = __float_as_uint(f);
#else
    union { float f; uint32_t u; } c; c.f = f; uint32_t u = c.u;
#endif
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float f32_unordered(uint32_t o) {
    uint32_t u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    union { float f; uint32_t u; } c; c.u = u; return c.f;
#endif
}
__host__ __device__ __forceinline__ uint64_t make_key(float score, uint32_t idx) {
    score = score + 0.0f;  // -0.0 -> +0.0 so that it ties with +0.0 like a float compare
    return (uint64_t(f32_ordered(score)) << 32) | uint64_t(0xffffffffu - idx);
}
__host__ __device__ __forceinline__ float key_score(uint64_t k) { return f32_unordered(uint32_t(k >> 32)); }
__host__ __device__ __forceinline__ uint32_t key_idx(uint64_t k) { return 0xffffffffu - uint32_t(k); }

constexpr uint64_t KEY_NONE = 0ull;  // below every real key (f32_ordered(x) >= 0x007fffff for non-NaN)

__host__ __device__ __forceinline__ uint32_t next_pow2(uint32_t v) {
    uint32_t p = 1;
    while (p < v) p <<= 1;
    return p;
}

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------
// Bitonic sort, DESCENDING, of n (power of two) u64 keys in shared memory.
// ---------------------------------------------------------------------------------------
// by one warp; callers guarantee only this warp touches buf.
__device__ __forceinline__ void warp_bitonic_desc(uint64_t *buf, uint32_t n, uint32_t lane) {
    for (uint32_t k = 2; k <= n; k <<= 1) {
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            __syncwarp();
            for (uint32_t i = lane; i < n; i += 32) {
                uint32_t ixj = i ^ j;
                if (ixj > i) {
                    uint64_t a = buf[i], b = buf[ixj];
                    bool desc = ((i & k) == 0);
                    if (desc ? (a < b) : (a > b)) { buf[i] = b; buf[ixj] = a; }
                }
            }
        }
    }
    __syncwarp();
}

// by `nthreads` threads that all call it (tid in [0,nthreads)); sync via named barrier `bar`.
__device__ __forceinline__ void named_bar_sync(uint32_t bar, uint32_t nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(bar), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void group_bitonic_desc(uint64_t *buf, uint32_t n, uint32_t tid,
                                                   uint32_t nthreads, uint32_t bar) {
    for (uint32_t k = 2; k <= n; k <<= 1) {
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            named_bar_sync(bar, nthreads);
            for (uint32_t i = tid; i < n; i += nthreads) {
                uint32_t ixj = i ^ j;
                if (ixj > i) {
                    uint64_t a = buf[i], b = buf[ixj];
                    bool desc = ((i & k) == 0);
                    if (desc ? (a < b) : (a > b)) { buf[i] = b; buf[ixj] = a; }
                }
            }
        }
    }
    named_bar_sync(bar, nthreads);
}

// ---------------------------------------------------------------------------------------
// mbarrier + 1-D bulk async copy (TMA engine; SASS: UBLKCP / SYNCS).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}
// global -> shared::cta bulk copy, completion signalled on `bar` as complete_tx(bytes).
// dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
// same with an L2 evict-first policy: the matrix is streamed once per sweep.
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulk_g2s_hint(void *dst_smem, const void *src_gmem, uint32_t bytes,
                                              uint64_t *bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], "
        "%4;" ::"r"(smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}

__device__ __forceinline__ float warp_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 16);
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v;
}
#endif  // __CUDACC__

}  // namespace oc
