// ring_common.cuh -- mbarrier / bulk-copy (cp.async.bulk, SASS UBLKCP) primitives and the ring position shared by the
// ring kernels (kernels_ring.cu: 240-column windows, dedicated producer warp; kernels_ring2.cu: 256-column windows, 8 consumer warps).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace dwtb200 {

constexpr int RING_SLOTS = 6;    // row pairs in flight per CTA

// ---- mbarrier / bulk-copy primitives -------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
        "selp.u32 %0, 1, 0, P1;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// global -> shared, completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}

template <class T, int N> __device__ __forceinline__ void lds_vec(uint32_t addr, T *v)
{
    static_assert(N * sizeof(T) == 32 || N * sizeof(T) == 16, "lane width");
#pragma unroll
    for (int i = 0; i < (int)(N * sizeof(T)) / 16; i++) {
        int4 r;
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr + 16 * i));
        *reinterpret_cast<int4 *>(reinterpret_cast<char *>(v) + 16 * i) = r;
    }
}
template <class T> __device__ __forceinline__ T lds_one(uint32_t addr)
{
    T v;
    if constexpr (sizeof(T) == 4) {
        uint32_t r;
        asm volatile("ld.shared.b32 %0, [%1];" : "=r"(r) : "r"(addr));
        v = *reinterpret_cast<T *>(&r);
    } else {
        unsigned long long r;
        asm volatile("ld.shared.b64 %0, [%1];" : "=l"(r) : "r"(addr));
        v = *reinterpret_cast<T *>(&r);
    }
    return v;
}

struct RingState {   // position in a consumer's ring, advanced identically by producer and consumer
    int slot = 0;
    uint32_t phase = 0;
    __device__ __forceinline__ void next()
    {
        if (++slot == RING_SLOTS) {
            slot = 0;
            phase ^= 1;
        }
    }
};

}  // namespace dwtb200
