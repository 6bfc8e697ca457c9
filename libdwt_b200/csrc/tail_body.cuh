// tail_body.cuh -- all remaining coarse levels of ONE frame inside one CTA's shared memory.
// Device bodies shared by the stand-alone tail kernels (kernels_tail.cu) and the persistent
// mid-pyramid kernels (kernels_tile.cu).  Semantics: see kernels_tail.cu.
#pragma once
#include "kernels.h"
#include "lifting.cuh"

namespace dwtb200 {

struct LdNc {   // read-only for the whole launch
    template <class T> __device__ __forceinline__ T operator()(const T *p) const { return __ldg(p); }
};
struct LdCg {   // written earlier in the same launch by other SMs: L2 only, never a stale L1 line
    template <class T> __device__ __forceinline__ T operator()(const T *p) const { return __ldcg(p); }
};

__host__ __device__ inline int cdiv_pow2(int v, int j) { return (int)(((int64_t)v + ((int64_t)1 << j) - 1) >> j); }

// length-1 line: only the unguarded double driver scales it
template <class WV> __device__ __forceinline__ typename WV::T one_fwd(typename WV::T v)
{
    return (WV::HAS_ONE && !WV::GUARD) ? WV::one_f(v) : v;
}
template <class WV> __device__ __forceinline__ typename WV::T one_inv(typename WV::T v)
{
    return (WV::HAS_ONE && !WV::GUARD) ? WV::one_i(v) : v;
}

// Both bodies run on ONE CTA with the LL band of `frame` resident in the two shared buffers bufA / bufB
// (each at least (W0>>j0 rounded up) * (H0>>j0 rounded up) elements).  LD loads global memory: __ldg
// for a stand-alone launch, __ldcg when the input was written earlier in the SAME launch (kernels_tile.cu).
template <class WV, class LD> __device__ __forceinline__ void fwd_tail_body(const TailParams &p, int frame, typename WV::T *bufA, typename WV::T *bufB, LD ld)
{
    using T = typename WV::T;
    const T *src = (const T *)p.src + (int64_t)frame * p.src_frame;
    T *dst = (T *)p.dst + (int64_t)frame * p.dst_frame;
    const int tid = threadIdx.x;
    const int TAIL_THREADS = blockDim.x;

    int w = cdiv_pow2(p.W0, p.j0), h = cdiv_pow2(p.H0, p.j0);
    for (int t = tid; t < w * h; t += TAIL_THREADS) bufA[t] = ld(src + (int64_t)(t / w) * p.src_pitch + (t % w));
    T *in = bufA, *other = bufB;
    __syncthreads();

    for (int j = p.j0; j < p.j1; j++) {
        const int nlx = (w + 1) >> 1, nhx = w >> 1, nly = (h + 1) >> 1, nhy = h >> 1;
        // ---- rows: in -> other, row y = [L (nlx) | H (nhx)] ----
        T *rows = in;
        if (!(WV::GUARD && w <= 1)) {
            rows = other;
            if (w >= 2) {
                for (int t = tid; t < h * nlx; t += TAIL_THREADS) {
                    const int y = t / nlx, k = t % nlx;
                    const T *line = in + y * w;
                    T win[2 * WV::HALO + 2];
#pragma unroll
                    for (int i = 0; i < 2 * WV::HALO + 2; i++) win[i] = line[reflect(2 * k - WV::HALO + i, w)];
                    T L, H;
                    window_fwd<WV>(win, L, H);
                    rows[y * w + k] = L;
                    if (k < nhx) rows[y * w + nlx + k] = H;
                }
            } else {
                for (int t = tid; t < h; t += TAIL_THREADS) rows[t] = one_fwd<WV>(in[t]);
            }
            __syncthreads();
        }
        // ---- columns: rows -> LL' (smem) + HL/LH/HH (global, Mallat positions) ----
        T *next = (rows == in) ? other : in;
        if (!(WV::GUARD && h <= 1) && h >= 2) {
            for (int t = tid; t < nly * w; t += TAIL_THREADS) {
                const int k = t / w, x = t % w;
                T win[2 * WV::HALO + 2];
#pragma unroll
                for (int i = 0; i < 2 * WV::HALO + 2; i++) win[i] = rows[reflect(2 * k - WV::HALO + i, h) * w + x];
                T L, H;
                window_fwd<WV>(win, L, H);
                if (x < nlx) next[k * nlx + x] = L;
                else dst[(int64_t)k * p.dst_pitch + x] = L;
                if (k < nhy) dst[(int64_t)(nly + k) * p.dst_pitch + x] = H;
            }
        } else {   // h == 1 (or guarded): the single row passes through
            for (int t = tid; t < h * w; t += TAIL_THREADS) {
                const int y = t / w, x = t % w;
                T v = rows[t];
                if (!(WV::GUARD && h <= 1)) v = one_fwd<WV>(v);
                if (x < nlx) next[y * nlx + x] = v;
                else dst[(int64_t)y * p.dst_pitch + x] = v;
            }
        }
        __syncthreads();
        in = next;
        other = (next == bufA) ? bufB : bufA;
        w = nlx;
        h = nly;
    }
    for (int t = tid; t < w * h; t += TAIL_THREADS) dst[(int64_t)(t / w) * p.dst_pitch + (t % w)] = in[t];
}

template <class WV, class LD> __device__ __forceinline__ void inv_tail_body(const TailParams &p, int frame, typename WV::T *bufA, typename WV::T *bufB, LD ld)
{
    using T = typename WV::T;
    const T *src = (const T *)p.src + (int64_t)frame * p.src_frame;
    T *dst = (T *)p.dst + (int64_t)frame * p.dst_frame;
    const int tid = threadIdx.x;
    const int TAIL_THREADS = blockDim.x;

    {   // coarsest LL band
        const int w = cdiv_pow2(p.W0, p.j1), h = cdiv_pow2(p.H0, p.j1);
        for (int t = tid; t < w * h; t += TAIL_THREADS) bufA[t] = ld(src + (int64_t)(t / w) * p.src_pitch + (t % w));
    }
    T *in = bufA, *tmp = bufB;
    __syncthreads();

    for (int j = p.j1; j > p.j0; j--) {
        const int w = cdiv_pow2(p.W0, j - 1), h = cdiv_pow2(p.H0, j - 1);   // size being reconstructed
        const int nlx = (w + 1) >> 1, nly = (h + 1) >> 1;
        // Mallat-arranged input of this level: LL from shared memory, the rest from the plane
        auto M = [&](int y, int x) -> T {
            return (y < nly && x < nlx) ? in[y * nlx + x] : ld(src + (int64_t)y * p.src_pitch + x);
        };
        const bool do_rows = !(WV::GUARD && w <= 1), do_cols = !(WV::GUARD && h <= 1);
        if constexpr (!WV::INV_COLS_FIRST) {
            // rows: M -> tmp (h x w, rows still in Mallat order)
            if (do_rows && w >= 2) {
                for (int t = tid; t < h * nlx; t += TAIL_THREADS) {
                    const int y = t / nlx, k = t % nlx;
                    T win[2 * WV::HALO + 2];
#pragma unroll
                    for (int i = 0; i < 2 * WV::HALO + 2; i++) {
                        const int c = reflect(2 * k - WV::HALO + i, w);
                        win[i] = M(y, (c & 1) ? nlx + (c >> 1) : (c >> 1));
                    }
                    T E, O;
                    window_inv<WV>(win, E, O);
                    tmp[y * w + 2 * k] = E;
                    if (2 * k + 1 < w) tmp[y * w + 2 * k + 1] = O;
                }
            } else {
                for (int t = tid; t < h * w; t += TAIL_THREADS) {
                    const T v = M(t / w, t % w);
                    tmp[t] = do_rows ? one_inv<WV>(v) : v;
                }
            }
            __syncthreads();
            // columns: tmp -> out (next LL in shared memory, or the destination plane at the end)
            T *out = in;   // `in` is dead once the row pass has finished
            const bool last = (j - 1 == p.j0);
            for (int t = tid; t < nly * w; t += TAIL_THREADS) {
                const int k = t / w, x = t % w;
                T E, O = T(0);
                if (do_cols && h >= 2) {
                    T win[2 * WV::HALO + 2];
#pragma unroll
                    for (int i = 0; i < 2 * WV::HALO + 2; i++) {
                        const int c = reflect(2 * k - WV::HALO + i, h);
                        win[i] = tmp[((c & 1) ? nly + (c >> 1) : (c >> 1)) * w + x];
                    }
                    window_inv<WV>(win, E, O);
                } else {
                    E = do_cols ? one_inv<WV>(tmp[x]) : tmp[x];   // h == 1
                }
                if (last) {
                    dst[(int64_t)(2 * k) * p.dst_pitch + x] = E;
                    if (2 * k + 1 < h) dst[(int64_t)(2 * k + 1) * p.dst_pitch + x] = O;
                } else {
                    out[(2 * k) * w + x] = E;
                    if (2 * k + 1 < h) out[(2 * k + 1) * w + x] = O;
                }
            }
            __syncthreads();
        } else {
            // columns first (int 5/3): M -> tmp (rows de-interleaved, columns still in Mallat order)
            for (int t = tid; t < nly * w; t += TAIL_THREADS) {
                const int k = t / w, x = t % w;
                if (h >= 2) {
                    T win[2 * WV::HALO + 2];
#pragma unroll
                    for (int i = 0; i < 2 * WV::HALO + 2; i++) {
                        const int c = reflect(2 * k - WV::HALO + i, h);
                        win[i] = M((c & 1) ? nly + (c >> 1) : (c >> 1), x);
                    }
                    T E, O;
                    window_inv<WV>(win, E, O);
                    tmp[(2 * k) * w + x] = E;
                    if (2 * k + 1 < h) tmp[(2 * k + 1) * w + x] = O;
                } else {
                    tmp[x] = one_inv<WV>(M(0, x));
                }
            }
            __syncthreads();
            T *out = in;
            const bool last = (j - 1 == p.j0);
            for (int t = tid; t < h * nlx; t += TAIL_THREADS) {
                const int y = t / nlx, k = t % nlx;
                T E, O = T(0);
                if (w >= 2) {
                    T win[2 * WV::HALO + 2];
#pragma unroll
                    for (int i = 0; i < 2 * WV::HALO + 2; i++) {
                        const int c = reflect(2 * k - WV::HALO + i, w);
                        win[i] = tmp[y * w + ((c & 1) ? nlx + (c >> 1) : (c >> 1))];
                    }
                    window_inv<WV>(win, E, O);
                } else {
                    E = one_inv<WV>(tmp[y]);
                }
                if (last) {
                    dst[(int64_t)y * p.dst_pitch + 2 * k] = E;
                    if (2 * k + 1 < w) dst[(int64_t)y * p.dst_pitch + 2 * k + 1] = O;
                } else {
                    out[y * w + 2 * k] = E;
                    if (2 * k + 1 < w) out[y * w + 2 * k + 1] = O;
                }
            }
            __syncthreads();
        }
    }
    if (p.j1 == p.j0) {   // nothing to do: pass the band through
        const int w = cdiv_pow2(p.W0, p.j0), h = cdiv_pow2(p.H0, p.j0);
        for (int t = tid; t < w * h; t += TAIL_THREADS) dst[(int64_t)(t / w) * p.dst_pitch + (t % w)] = in[t];
    }
}

}  // namespace dwtb200
