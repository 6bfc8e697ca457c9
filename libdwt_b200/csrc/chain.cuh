// chain.cuh -- device side of the dataflow links between the kernels of one transform (struct Chain, kernels.h).
#pragma once
#include "kernels.h"

namespace dwtb200 {

__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// every thread, first thing in a chained kernel: the generation, then let the successor be scheduled.  A kernel
// whose input is not covered by counters waits for its predecessor in the ordinary way.
__device__ __forceinline__ uint32_t chain_begin(const Chain &c)
{
    const uint32_t gen = c.gen ? *(const volatile uint32_t *)c.gen : 0u;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (!c.in) asm volatile("griddepcontrol.wait;" ::: "memory");
    return gen;
}
__device__ __forceinline__ int chain_block(const Chain &c, int row)
{
    const int b = (row + c.in_bias) / c.in_div;
    return b < c.in_nblocks ? b : c.in_nblocks - 1;
}
// spin until the producer of input block `block` of `frame` has finished it in this generation
__device__ __forceinline__ void chain_wait(const Chain &c, uint32_t gen, int frame, int block)
{
    const uint32_t target = (gen + 1u) * (uint32_t)c.in_need;
    const uint32_t *f = c.in + (size_t)frame * c.in_nblocks + block;
    while ((int32_t)(ld_acquire_gpu(f) - target) < 0) __nanosleep(40);
}
// the same question without waiting
__device__ __forceinline__ bool chain_test(const Chain &c, uint32_t gen, int frame, int block)
{
    const uint32_t target = (gen + 1u) * (uint32_t)c.in_need;
    return (int32_t)(ld_acquire_gpu(c.in + (size_t)frame * c.in_nblocks + block) - target) >= 0;
}
// one thread, after a CTA-wide barrier that follows the CTA's stores
__device__ __forceinline__ void chain_signal(const Chain &c, int frame, int block)
{
    if (c.out) {
        __threadfence();
        atomicAdd(c.out + (size_t)frame * c.out_nblocks + block, 1u);
    }
    if (c.done && atomicAdd(c.done, 1u) == c.total - 1u) {   // last CTA of the transform's last chained kernel
        *c.done = 0u;
        __threadfence();
        atomicAdd(const_cast<uint32_t *>(c.gen), 1u);
    }
}

// a running window of verified input blocks (one thread): rows are asked for in nearly monotonic order
struct ChainWindow {
    int lo = 0x7fffffff, hi = -1;
    __device__ __forceinline__ void need_row(const Chain &c, uint32_t gen, int frame, int row)
    {
        const int b = chain_block(c, row);
        if (b >= lo && b <= hi) return;
        if (hi < 0) {
            chain_wait(c, gen, frame, b);
            lo = hi = b;
        } else if (b > hi) {
            for (int x = hi + 1; x <= b; x++) chain_wait(c, gen, frame, x);
            hi = b;
        } else {
            for (int x = b; x < lo; x++) chain_wait(c, gen, frame, x);
            lo = b;
        }
        asm volatile("fence.proxy.async;" ::: "memory");   // the data is read next by the bulk-copy (async) proxy
    }
    // need_row without waiting: false when a block the row needs is not complete yet (what was verified so far is kept)
    __device__ __forceinline__ bool try_row(const Chain &c, uint32_t gen, int frame, int row)
    {
        const int b = chain_block(c, row);
        if (b >= lo && b <= hi) return true;
        if (hi < 0) {
            if (!chain_test(c, gen, frame, b)) return false;
            lo = hi = b;
        } else if (b > hi) {
            for (int x = hi + 1; x <= b; x++) {
                if (!chain_test(c, gen, frame, x)) return false;
                hi = x;
            }
        } else {
            for (int x = lo - 1; x >= b; x--) {
                if (!chain_test(c, gen, frame, x)) return false;
                lo = x;
            }
        }
        asm volatile("fence.proxy.async;" ::: "memory");
        return true;
    }
};

}  // namespace dwtb200
