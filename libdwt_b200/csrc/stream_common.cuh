// stream_common.cuh -- register-level building blocks of the streaming kernels (kernels_stream.cu, kernels_vol.cu):
// 16-byte vector access, row lifting across the lanes of a warp, column lifting as a register pipeline.
#pragma once
#include "kernels.h"
#include "lifting.cuh"

namespace dwtb200 {

constexpr unsigned FULL = 0xffffffffu;

// cache policy of the streaming accesses: every sample is read once and written once per level
#ifndef DWT_STREAM_HINTS
#define DWT_STREAM_HINTS 0   // measured: evict-first hints cost 6 % on level 0 (L1 serves the second half-row load, L2 merges the subband rows)
#endif
#if DWT_STREAM_HINTS
#define DWT_LD(ptr) __ldcs(ptr)          // ld.global.cs: evict-first
#define DWT_ST(ptr, val) __stcs(ptr, val)   // st.global.cs
#else
#define DWT_LD(ptr) __ldg(ptr)
#define DWT_ST(ptr, val) (*(ptr) = (val))
#endif

template <class T, int VPL> struct Row {
    T v[VPL];
};

// ---- 16-byte vector access -----------------------------------------------------------------------
template <class T, int N> __device__ __forceinline__ void ld_vec(const T *p, T *v)
{
    constexpr int BYTES = N * (int)sizeof(T);
    static_assert(BYTES == 8 || BYTES % 16 == 0, "vector width");
    if constexpr (BYTES == 8) {
        const int2 r = DWT_LD(reinterpret_cast<const int2 *>(p));
        *reinterpret_cast<int2 *>(v) = r;
    } else {
        constexpr int PER = 16 / sizeof(T);
#pragma unroll
        for (int i = 0; i < N / PER; i++) {
            const int4 r = DWT_LD(reinterpret_cast<const int4 *>(p) + i);
            *reinterpret_cast<int4 *>(v + i * PER) = r;
        }
    }
}
template <class T, int N> __device__ __forceinline__ void st_vec(T *p, const T *v)
{
    constexpr int BYTES = N * (int)sizeof(T);
    static_assert(BYTES == 8 || BYTES % 16 == 0, "vector width");
    if constexpr (BYTES == 8) {
        DWT_ST(reinterpret_cast<int2 *>(p), *reinterpret_cast<const int2 *>(v));
    } else {
        constexpr int PER = 16 / sizeof(T);
#pragma unroll
        for (int i = 0; i < N / PER; i++) DWT_ST(reinterpret_cast<int4 *>(p) + i, *reinterpret_cast<const int4 *>(v + i * PER));
    }
}

// ---- row lifting in registers: lane holds VPL consecutive samples, v[0] at an even column ------
template <class WV, int S, int VPL, bool INV> __device__ __forceinline__ void hstep_odd(typename WV::T (&v)[VPL])
{
    using T = typename WV::T;
    const T nxt = __shfl_down_sync(FULL, v[0], 1);
#pragma unroll
    for (int i = 1; i < VPL; i += 2) {
        const T r = (i + 1 < VPL) ? v[(i + 1) % VPL] : nxt;
        v[i] = INV ? WV::template i<S>(v[i], v[i - 1], r) : WV::template f<S>(v[i], v[i - 1], r);
    }
}
template <class WV, int S, int VPL, bool INV> __device__ __forceinline__ void hstep_even(typename WV::T (&v)[VPL])
{
    using T = typename WV::T;
    const T prv = __shfl_up_sync(FULL, v[VPL - 1], 1);
#pragma unroll
    for (int i = 0; i < VPL; i += 2) {
        const T l = i ? v[(i + VPL - 1) % VPL] : prv;
        v[i] = INV ? WV::template i<S>(v[i], l, v[i + 1]) : WV::template f<S>(v[i], l, v[i + 1]);
    }
}
template <class WV, int VPL> __device__ __forceinline__ void hfwd(typename WV::T (&v)[VPL])
{
    hstep_odd<WV, 0, VPL, false>(v);
    hstep_even<WV, 1, VPL, false>(v);
    if constexpr (WV::NS == 4) {
        hstep_odd<WV, 2, VPL, false>(v);
        hstep_even<WV, 3, VPL, false>(v);
    }
#pragma unroll
    for (int i = 0; i < VPL; i += 2) {
        v[i] = WV::fse(v[i]);
        v[i + 1] = WV::fso(v[i + 1]);
    }
}
template <class WV, int VPL> __device__ __forceinline__ void hinv(typename WV::T (&v)[VPL])
{
#pragma unroll
    for (int i = 0; i < VPL; i += 2) {
        v[i] = WV::ise(v[i]);
        v[i + 1] = WV::iso(v[i + 1]);
    }
    hstep_even<WV, 0, VPL, true>(v);
    hstep_odd<WV, 1, VPL, true>(v);
    if constexpr (WV::NS == 4) {
        hstep_even<WV, 2, VPL, true>(v);
        hstep_odd<WV, 3, VPL, true>(v);
    }
}

// ---- column lifting: one pipeline step per row pair --------------------------------------------------
// forward: consumes row-lifted rows a = 2m+1, b = 2m+2; state st[0] = row 2m, st[1..] = d1, s1, d2 of
// the rows above; produces the low / high output rows of pair m - DELAY
template <class WV, int VPL>
__device__ __forceinline__ void vfwd(const typename WV::T (&a)[VPL], const typename WV::T (&b)[VPL], typename WV::T (&st)[WV::NS][VPL],
                                     typename WV::T (&oL)[VPL], typename WV::T (&oH)[VPL])
{
    using T = typename WV::T;
#pragma unroll
    for (int i = 0; i < VPL; i++) {
        if constexpr (WV::NS == 4) {
            const T d1n = WV::template f<0>(a[i], st[0][i], b[i]);
            const T s1n = WV::template f<1>(st[0][i], st[1][i], d1n);
            const T d2n = WV::template f<2>(st[1][i], st[2][i], s1n);
            const T s2n = WV::template f<3>(st[2][i], st[3][i], d2n);
            oL[i] = WV::fse(s2n);
            oH[i] = WV::fso(d2n);
            st[0][i] = b[i];
            st[1][i] = d1n;
            st[2][i] = s1n;
            st[3][i] = d2n;
        } else {
            const T d1n = WV::template f<0>(a[i], st[0][i], b[i]);
            const T s1n = WV::template f<1>(st[0][i], st[1][i], d1n);
            oL[i] = WV::fse(s1n);
            oH[i] = WV::fso(d1n);
            st[0][i] = b[i];
            st[1][i] = d1n;
        }
    }
}
// inverse: consumes coefficient rows a = 2k (L row), b = 2k+1 (H row); produces output rows 2q-1 (oO)
// and 2q (oE), q = k - DELAY
template <class WV, int VPL>
__device__ __forceinline__ void vinv(const typename WV::T (&a)[VPL], const typename WV::T (&b)[VPL], typename WV::T (&st)[WV::NS][VPL],
                                     typename WV::T (&oO)[VPL], typename WV::T (&oE)[VPL])
{
    using T = typename WV::T;
#pragma unroll
    for (int i = 0; i < VPL; i++) {
        if constexpr (WV::NS == 4) {
            const T s2k = WV::ise(a[i]), d2k = WV::iso(b[i]);
            const T s1k = WV::template i<0>(s2k, st[0][i], d2k);
            const T d1m = WV::template i<1>(st[0][i], st[1][i], s1k);
            const T xen = WV::template i<2>(st[1][i], st[2][i], d1m);
            const T xo = WV::template i<3>(st[2][i], st[3][i], xen);
            oO[i] = xo;
            oE[i] = xen;
            st[0][i] = d2k;
            st[1][i] = s1k;
            st[2][i] = d1m;
            st[3][i] = xen;
        } else {
            const T ck = WV::ise(a[i]), cn = WV::iso(b[i]);
            const T xen = WV::template i<0>(ck, st[0][i], cn);
            const T xo = WV::template i<1>(st[0][i], st[1][i], xen);
            oO[i] = xo;
            oE[i] = xen;
            st[0][i] = cn;
            st[1][i] = xen;
        }
    }
}

}  // namespace dwtb200
