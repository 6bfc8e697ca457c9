// kernels_tail.cu -- all remaining coarse levels of a plane in ONE launch, one CTA per frame, the
// shrinking LL band resident in shared memory (sm_100a).
//
// The last ~7 levels of a pyramid hold < 0.1 % of the samples but would cost a launch (and an HBM
// round trip of the LL band) each; here they cost one launch and the LL band never leaves the SM.
// Semantics are those of the reference drivers level by level, including the degenerate shapes a
// `decompose_one` pyramid runs into (/root/reference/src/libdwt.c:12837, 12867: float skips a pass
// when the line count is <= 1; :2036, :11437: double scales a length-1 line; :10961: int leaves it).
// Lines are evaluated pair-wise with the mirrored window of lifting.cuh, so results are
// bit-identical to the streaming kernels and to the reference.
#include "chain.cuh"
#include "tail_body.cuh"

namespace dwtb200 {

constexpr int TAIL_THREADS = 1024;
constexpr int TAIL_CAP_BYTES = 64 * 1024;                        // largest LL band a tail takes (samples x size)
constexpr int TAIL_BUF_BYTES = TAIL_CAP_BYTES / 16 * 17;         // per buffer, two buffers: room for the odd row pitch (tail_pitch)

template <class WV> __global__ void __launch_bounds__(TAIL_THREADS) k_fwd_tail(const TailParams p)
{
    using T = typename WV::T;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t gen = chain_begin(p.chain);
    T *bufA = reinterpret_cast<T *>(smem_raw), *bufB = reinterpret_cast<T *>(smem_raw + TAIL_BUF_BYTES);
    if (p.chain.in) {   // chained: wait for every row block of the LL band, read it through L2
        for (int b = threadIdx.x; b < p.chain.in_nblocks; b += TAIL_THREADS) chain_wait(p.chain, gen, blockIdx.x, b);
        __syncthreads();
        fwd_tail_body<WV>(p, blockIdx.x, bufA, bufB, LdCg());
    } else {
        fwd_tail_body<WV>(p, blockIdx.x, bufA, bufB, LdNc());
    }
    if (p.chain.gen) {
        __syncthreads();
        if (threadIdx.x == 0) chain_signal(p.chain, blockIdx.x, 0);
    }
}
template <class WV> __global__ void __launch_bounds__(TAIL_THREADS) k_inv_tail(const TailParams p)
{
    using T = typename WV::T;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    chain_begin(p.chain);   // first kernel of an inverse chain: its input is complete at launch
    inv_tail_body<WV>(p, blockIdx.x, reinterpret_cast<T *>(smem_raw), reinterpret_cast<T *>(smem_raw + TAIL_BUF_BYTES), LdNc(),
                      reinterpret_cast<T *>(smem_raw + 2 * TAIL_BUF_BYTES));
    if (p.chain.gen) {
        __syncthreads();
        if (threadIdx.x == 0) chain_signal(p.chain, blockIdx.x, 0);
    }
}

int tail_max_elems(int kind) { return TAIL_CAP_BYTES / kind_elem_size(kind); }

// Kernels are loaded lazily by the CUDA runtime; loading (and cudaFuncSetAttribute) is not allowed while
// a stream is being captured, so dwtb200_init() calls this once before any graph is built.
template <class K> static cudaError_t prep(K kern, int bufs)
{
    cudaFuncAttributes a;
    cudaError_t e = cudaFuncGetAttributes(&a, kern);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bufs * TAIL_BUF_BYTES);
    return e;
}
cudaError_t preload_tail()
{
    cudaError_t e = cudaSuccess;
    for (int kind = 0; kind < K_COUNT; kind++)
        dispatch_kind(kind, [&](auto wv) {
            using WV = decltype(wv);
            if (e == cudaSuccess) e = prep(k_fwd_tail<WV>, 2);
            if (e == cudaSuccess) e = prep(k_inv_tail<WV>, 3);   // third buffer: the staged Mallat block
        });
    return e;
}
void launch_fwd_tail(int kind, const TailParams &p, int frames, cudaStream_t st)
{
    const size_t sm = 2 * TAIL_BUF_BYTES;
    dispatch_kind(kind, [&](auto wv) { launch_pdl(k_fwd_tail<decltype(wv)>, dim3(frames), dim3(TAIL_THREADS), sm, st, p.chain.pdl, p); });
}
void launch_inv_tail(int kind, const TailParams &p, int frames, cudaStream_t st)
{
    const size_t sm = 3 * TAIL_BUF_BYTES;
    dispatch_kind(kind, [&](auto wv) { launch_pdl(k_inv_tail<decltype(wv)>, dim3(frames), dim3(TAIL_THREADS), sm, st, p.chain.pdl, p); });
}

}  // namespace dwtb200
