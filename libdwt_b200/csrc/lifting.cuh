// lifting.cuh -- per-wavelet lifting arithmetic shared by every kernel (sm_100a).
//
// Each `Wxx` struct is the arithmetic of one reference line function, written as single lifting
// steps so that the same code serves the streaming level kernels, the tail kernel and the generic
// pass kernels.  Everything is evaluated with explicitly rounded intrinsics (__fadd_rn/__fmul_rn,
// __dadd_rn/__dmul_rn): the compiler may not contract them into FMAs, so a sample goes through
// exactly the operations of the reference's x86-64 SSE2 build and float/double output is
// bit-identical, not merely within tolerance.
//
//   W97F  dwt_cdf97_f_ex_stride_s / dwt_cdf97_i_ex_stride_s   /root/reference/src/libdwt.c:10744, 11530
//         (core accel_lift_op4s_main_s :2264, edges :9510 :9844 :10199)
//   W97D  dwt_cdf97_f_ex_stride_d / dwt_cdf97_i_ex_stride_d   src/libdwt.c:2024, 11424
//   W53I  dwt_cdf53_f_ex_stride_i / dwt_cdf53_i_ex_stride_i   src/libdwt.c:10950, 11749
//   constants                                                  src/inline.h:310-341
//
// Step numbering: forward step S (0..NS-1) updates ODD samples when S is even and EVEN samples when
// S is odd; inverse step S updates EVEN samples when S is even.  A step is x' = step(x, left, right).
// Boundaries are whole-sample mirrors (t[-1]=t[1], t[N]=t[N-2]); the reference writes the edge as
// (2c)*nb, which equals c*(nb+nb) bit for bit, and for int ((nb+nb)>>1)==nb, ((2nb+2)>>2)==((nb+1)>>1)
// as long as 2*nb does not overflow int32 (documented domain: |coefficient| < 2^30).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace dwtb200 {

enum Kind : int { K_CDF97_F32 = 0, K_CDF97_F64 = 1, K_CDF53_I32 = 2, K_CDF53_F32 = 3, K_CDF53_F64 = 4, K_CDF97_I32 = 5, K_COUNT = 6 };

struct W97F {
    using T = float;
    static constexpr int NS = 4;       // lifting steps
    static constexpr int HALO = 4;     // input samples needed either side of an output sample
    static constexpr bool HAS_ONE = true;   // a line of length 1 is scaled (libdwt.c:10757, 11546)
    static constexpr bool GUARD = true;     // driver skips a pass when the line count is <= 1 (libdwt.c:12837)
    static constexpr bool INV_COLS_FIRST = false;
    // double literals rounded to float, exactly as `static const float x = <double literal>` does
    static constexpr float P1 = (float)1.58613434342059, U1 = (float)-0.0529801185729;
    static constexpr float P2 = (float)-0.8829110755309, U2 = (float)0.4435068520439;
    static constexpr float Z = (float)1.1496043988602;
    static constexpr float IZ = 1.0f / Z;                          // 1/zeta evaluated in float (libdwt.c:2289)
    static constexpr float S2 = (float)(1 / 1.1496043988602);      // dwt_cdf97_s2_s, 1 ulp above IZ
    template <int S> static __device__ __forceinline__ T cf() { return S == 0 ? -P1 : S == 1 ? U1 : S == 2 ? -P2 : U2; }
    template <int S> static __device__ __forceinline__ T ci() { return S == 0 ? -U2 : S == 1 ? P2 : S == 2 ? -U1 : P1; }
    template <int S> static __device__ __forceinline__ T f(T x, T l, T r) { return __fadd_rn(x, __fmul_rn(cf<S>(), __fadd_rn(l, r))); }
    template <int S> static __device__ __forceinline__ T i(T x, T l, T r) { return __fadd_rn(x, __fmul_rn(ci<S>(), __fadd_rn(l, r))); }
    static __device__ __forceinline__ T fse(T x) { return __fmul_rn(x, Z); }    // forward scale, even sample
    static __device__ __forceinline__ T fso(T x) { return __fmul_rn(x, IZ); }   // forward scale, odd sample
    static __device__ __forceinline__ T ise(T x) { return __fmul_rn(x, IZ); }
    static __device__ __forceinline__ T iso(T x) { return __fmul_rn(x, Z); }
    static __device__ __forceinline__ T one_f(T x) { return __fmul_rn(x, Z); }
    static __device__ __forceinline__ T one_i(T x) { return __fmul_rn(x, S2); }
};

struct W97D {
    using T = double;
    static constexpr int NS = 4;
    static constexpr int HALO = 4;
    static constexpr bool HAS_ONE = true;    // libdwt.c:2036, 11437
    static constexpr bool GUARD = false;     // libdwt.c:12481: no `lines > 1` test
    static constexpr bool INV_COLS_FIRST = false;
    static constexpr double P1 = 1.58613434342059, U1 = -0.0529801185729;
    static constexpr double P2 = -0.8829110755309, U2 = 0.4435068520439;
    static constexpr double Z = 1.1496043988602;
    static constexpr double S2 = 1 / 1.1496043988602;
    template <int S> static __device__ __forceinline__ T cf() { return S == 0 ? -P1 : S == 1 ? U1 : S == 2 ? -P2 : U2; }
    template <int S> static __device__ __forceinline__ T ci() { return S == 0 ? -U2 : S == 1 ? P2 : S == 2 ? -U1 : P1; }
    template <int S> static __device__ __forceinline__ T f(T x, T l, T r) { return __dadd_rn(x, __dmul_rn(cf<S>(), __dadd_rn(l, r))); }
    template <int S> static __device__ __forceinline__ T i(T x, T l, T r) { return __dadd_rn(x, __dmul_rn(ci<S>(), __dadd_rn(l, r))); }
    static __device__ __forceinline__ T fse(T x) { return __dmul_rn(x, Z); }
    static __device__ __forceinline__ T fso(T x) { return __dmul_rn(x, S2); }
    static __device__ __forceinline__ T ise(T x) { return __dmul_rn(x, S2); }
    static __device__ __forceinline__ T iso(T x) { return __dmul_rn(x, Z); }
    static __device__ __forceinline__ T one_f(T x) { return __dmul_rn(x, Z); }
    static __device__ __forceinline__ T one_i(T x) { return __dmul_rn(x, S2); }
};

struct W53I {
    using T = int32_t;
    static constexpr int NS = 2;
    static constexpr int HALO = 2;
    static constexpr bool HAS_ONE = false;   // libdwt.c:10961: N < 2 leaves the line untouched
    static constexpr bool GUARD = false;
    static constexpr bool INV_COLS_FIRST = true;   // libdwt.c:18178, 18187
    static __device__ __forceinline__ T wadd(T a, T b) { return (T)((uint32_t)a + (uint32_t)b); }
    static __device__ __forceinline__ T wsub(T a, T b) { return (T)((uint32_t)a - (uint32_t)b); }
    template <int S> static __device__ __forceinline__ T f(T x, T l, T r)
    {
        if (S == 0) return wsub(x, wadd(l, r) >> 1);          // predict, odd samples
        return wadd(x, wadd(wadd(l, r), 2) >> 2);             // update, even samples
    }
    template <int S> static __device__ __forceinline__ T i(T x, T l, T r)
    {
        if (S == 0) return wsub(x, wadd(wadd(l, r), 2) >> 2); // undo update, even samples
        return wadd(x, wadd(l, r) >> 1);                      // undo predict, odd samples
    }
    static __device__ __forceinline__ T fse(T x) { return x; }
    static __device__ __forceinline__ T fso(T x) { return x; }
    static __device__ __forceinline__ T ise(T x) { return x; }
    static __device__ __forceinline__ T iso(T x) { return x; }
    static __device__ __forceinline__ T one_f(T x) { return x; }
    static __device__ __forceinline__ T one_i(T x) { return x; }
};

// ---- sibling transforms sharing the same kernels -------------------------
//   W53F  dwt_cdf53_f_ex_stride_s / dwt_cdf53_i_ex_stride_s   src/libdwt.c:10986, 11785   (drivers :16470, 18296)
//   W53D  dwt_cdf53_f_ex_stride_d / dwt_cdf53_i_ex_stride_d   src/libdwt.c:2085, 11484    (drivers :12535, 16962)
//   W97I  dwt_cdf97_f_ex_stride_i / dwt_cdf97_i_ex_stride_i   src/libdwt.c:10901, 11699   (drivers :16387, 18219)
// `x -= c*(l+r)` of the reference equals x + (-c)*(l+r) bit for bit (negation is exact).
template <class FT> struct W53Float {
    using T = FT;
    static constexpr int NS = 2;
    static constexpr int HALO = 2;
    static constexpr bool HAS_ONE = true;    // libdwt.c:10997, 11797: a line of length 1 is scaled
    static constexpr bool GUARD = false;     // libdwt.c:16508: no `lines > 1` test
    static constexpr bool INV_COLS_FIRST = false;   // libdwt.c:18333 rows, then :18342 columns
    static constexpr FT P1 = (FT)0.5, U1 = (FT)0.25;
    static constexpr FT S1 = (FT)1.41421356237309504880, S2 = (FT)0.70710678118654752440;   // inline.h:333-341
    static __device__ __forceinline__ T add(T a, T b)
    {
        if constexpr (sizeof(T) == 4) return __fadd_rn(a, b);
        else return __dadd_rn(a, b);
    }
    static __device__ __forceinline__ T mul(T a, T b)
    {
        if constexpr (sizeof(T) == 4) return __fmul_rn(a, b);
        else return __dmul_rn(a, b);
    }
    template <int S> static __device__ __forceinline__ T f(T x, T l, T r) { return add(x, mul(S == 0 ? -P1 : U1, add(l, r))); }
    template <int S> static __device__ __forceinline__ T i(T x, T l, T r) { return add(x, mul(S == 0 ? -U1 : P1, add(l, r))); }
    static __device__ __forceinline__ T fse(T x) { return mul(x, S1); }
    static __device__ __forceinline__ T fso(T x) { return mul(x, S2); }
    static __device__ __forceinline__ T ise(T x) { return mul(x, S2); }
    static __device__ __forceinline__ T iso(T x) { return mul(x, S1); }
    static __device__ __forceinline__ T one_f(T x) { return mul(x, S1); }
    static __device__ __forceinline__ T one_i(T x) { return mul(x, S2); }
};
using W53F = W53Float<float>;
using W53D = W53Float<double>;

struct W97I {
    using T = int32_t;
    static constexpr int NS = 4;
    static constexpr int HALO = 4;
    static constexpr bool HAS_ONE = false;   // libdwt.c:10912: N < 2 leaves the line untouched
    static constexpr bool GUARD = false;
    static constexpr bool INV_COLS_FIRST = true;   // libdwt.c:18256 columns, then :18265 rows
    // 32-bit two's-complement wrap-around like the compiled reference; >> is arithmetic
    static __device__ __forceinline__ T q7(int c, T l, T r)    // ( c*(l+r) - (1<<6) ) >> 7
    {
        return (T)((uint32_t)c * ((uint32_t)l + (uint32_t)r) - 64u) >> 7;
    }
    static __device__ __forceinline__ T q12(int c, T l, T r)   // ( c*(l+r) + (1<<11) ) >> 12
    {
        return (T)((uint32_t)c * ((uint32_t)l + (uint32_t)r) + 2048u) >> 12;
    }
    static __device__ __forceinline__ T wadd(T a, T b) { return (T)((uint32_t)a + (uint32_t)b); }
    static __device__ __forceinline__ T wsub(T a, T b) { return (T)((uint32_t)a - (uint32_t)b); }
    template <int S> static __device__ __forceinline__ T f(T x, T l, T r)
    {
        if (S == 0) return wsub(x, q7(203, l, r));
        if (S == 1) return wadd(x, q12(-217, l, r));
        if (S == 2) return wsub(x, q7(-113, l, r));
        return wadd(x, q12(1817, l, r));
    }
    template <int S> static __device__ __forceinline__ T i(T x, T l, T r)
    {
        if (S == 0) return wsub(x, q12(1817, l, r));
        if (S == 1) return wadd(x, q7(-113, l, r));
        if (S == 2) return wsub(x, q12(-217, l, r));
        return wadd(x, q7(203, l, r));
    }
    static __device__ __forceinline__ T fse(T x) { return x; }
    static __device__ __forceinline__ T fso(T x) { return x; }
    static __device__ __forceinline__ T ise(T x) { return x; }
    static __device__ __forceinline__ T iso(T x) { return x; }
    static __device__ __forceinline__ T one_f(T x) { return x; }
    static __device__ __forceinline__ T one_i(T x) { return x; }
};

// run f(WV{}) for the wavelet/type of `kind`
template <class F> inline void dispatch_kind(int kind, F &&f)
{
    switch (kind) {
    case K_CDF97_F32: f(W97F{}); break;
    case K_CDF97_F64: f(W97D{}); break;
    case K_CDF53_I32: f(W53I{}); break;
    case K_CDF53_F32: f(W53F{}); break;
    case K_CDF53_F64: f(W53D{}); break;
    default: f(W97I{}); break;
    }
}
inline int kind_lifting_steps(int kind) { return (kind == K_CDF97_F32 || kind == K_CDF97_F64 || kind == K_CDF97_I32) ? 4 : 2; }
inline int kind_elem_size(int kind) { return (kind == K_CDF97_F64 || kind == K_CDF53_F64) ? 8 : 4; }
inline int kind_elem_class(int kind) { return (kind == K_CDF97_F32 || kind == K_CDF53_F32) ? 1 : kind_elem_size(kind) == 8 ? 2 : 0; }   // 0 int32, 1 float, 2 double

// Programmatic dependent launch: every dense-path kernel lets its successor be scheduled at once
// (launch_dependents) and waits for its predecessor's results before touching memory (wait).  A
// dependent launch then costs a fraction of the ~5 us of a fully serialised kernel boundary.
__device__ __forceinline__ void pdl_begin()
{
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}
template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int pdl, Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, args...);
}
extern int g_use_pdl;   // dwtb200.cu (DWTB200_TUNE_PDL)

// Whole-sample mirror of index i into [0, n), n >= 2; period 2(n-1) so windows wider than the
// line (tiny coarse levels) fold repeatedly, exactly like the reference's N = 2, 3, 4 cases.
__device__ __forceinline__ int reflect(int i, int n)
{
    if ((unsigned)i < (unsigned)n) return i;
    // one fold is enough unless the line is shorter than the window overhang (the last few levels of a pyramid);
    // folding in a loop is much cheaper than the integer modulo it replaces
    do {
        if (i < 0) i = -i;
        if (i >= n) i = 2 * (n - 1) - i;
    } while ((unsigned)i >= (unsigned)n);
    return i;
}

// ---- window evaluation: one (even, odd) output pair from a 2*HALO+2 sample window -------------
// forward:  w[q] = x[2k-HALO+q]  ->  L[k] (even), H[k] (odd)
// inverse:  w[q] = c[2k-HALO+q]  (interleaved coefficients)  ->  x[2k], x[2k+1]
template <class WV> __device__ __forceinline__ void window_fwd(typename WV::T (&w)[2 * WV::HALO + 2], typename WV::T &L, typename WV::T &H)
{
    if constexpr (WV::NS == 4) {
#pragma unroll
        for (int q = 1; q <= 7; q += 2) w[q] = WV::template f<0>(w[q], w[q - 1], w[q + 1]);
#pragma unroll
        for (int q = 2; q <= 6; q += 2) w[q] = WV::template f<1>(w[q], w[q - 1], w[q + 1]);
#pragma unroll
        for (int q = 3; q <= 5; q += 2) w[q] = WV::template f<2>(w[q], w[q - 1], w[q + 1]);
        w[4] = WV::template f<3>(w[4], w[3], w[5]);
        L = WV::fse(w[4]);
        H = WV::fso(w[5]);
    } else {
        w[1] = WV::template f<0>(w[1], w[0], w[2]);
        w[3] = WV::template f<0>(w[3], w[2], w[4]);
        w[2] = WV::template f<1>(w[2], w[1], w[3]);
        L = WV::fse(w[2]);
        H = WV::fso(w[3]);
    }
}
template <class WV> __device__ __forceinline__ void window_inv(typename WV::T (&w)[2 * WV::HALO + 2], typename WV::T &E, typename WV::T &O)
{
    constexpr int N = 2 * WV::HALO + 2;
#pragma unroll
    for (int q = 0; q < N; q += 2) { w[q] = WV::ise(w[q]); w[q + 1] = WV::iso(w[q + 1]); }
    if constexpr (WV::NS == 4) {
#pragma unroll
        for (int q = 2; q <= 8; q += 2) w[q] = WV::template i<0>(w[q], w[q - 1], w[q + 1]);
#pragma unroll
        for (int q = 3; q <= 7; q += 2) w[q] = WV::template i<1>(w[q], w[q - 1], w[q + 1]);
#pragma unroll
        for (int q = 4; q <= 6; q += 2) w[q] = WV::template i<2>(w[q], w[q - 1], w[q + 1]);
        w[5] = WV::template i<3>(w[5], w[4], w[6]);
        E = w[4];
        O = w[5];
    } else {
        w[2] = WV::template i<0>(w[2], w[1], w[3]);
        w[4] = WV::template i<0>(w[4], w[3], w[5]);
        w[3] = WV::template i<1>(w[3], w[2], w[4]);
        E = w[2];
        O = w[3];
    }
}

// ---- P output pairs from one window of 2P + 2*HALO samples (kernels_pyr.cu) -------------------------------
// The same lifting steps as window_fwd / window_inv, evaluated on a longer window so that the redundant work at the
// window edges is shared by P pairs: forward step S updates the samples of its parity in [S+1, n-2-S], inverse step S
// those in [2+S, n-2-S]; outputs are w[HALO + 2i] (even) and w[HALO + 1 + 2i] (odd), i < P.
template <class WV, int S, int N> __device__ __forceinline__ void lift_range_fwd(typename WV::T (&w)[N])
{
#pragma unroll
    for (int q = S + 1; q <= N - 2 - S; q += 2) w[q] = WV::template f<S>(w[q], w[q - 1], w[q + 1]);
}
template <class WV, int S, int N> __device__ __forceinline__ void lift_range_inv(typename WV::T (&w)[N])
{
#pragma unroll
    for (int q = 2 + S; q <= N - 2 - S; q += 2) w[q] = WV::template i<S>(w[q], w[q - 1], w[q + 1]);
}
template <class WV, int P>
__device__ __forceinline__ void window_fwd_p(typename WV::T (&w)[2 * P + 2 * WV::HALO], typename WV::T (&L)[P], typename WV::T (&H)[P])
{
    constexpr int N = 2 * P + 2 * WV::HALO;
    lift_range_fwd<WV, 0, N>(w);
    lift_range_fwd<WV, 1, N>(w);
    if constexpr (WV::NS == 4) {
        lift_range_fwd<WV, 2, N>(w);
        lift_range_fwd<WV, 3, N>(w);
    }
#pragma unroll
    for (int i = 0; i < P; i++) {
        L[i] = WV::fse(w[WV::HALO + 2 * i]);
        H[i] = WV::fso(w[WV::HALO + 1 + 2 * i]);
    }
}
template <class WV, int P>
__device__ __forceinline__ void window_inv_p(typename WV::T (&w)[2 * P + 2 * WV::HALO], typename WV::T (&E)[P], typename WV::T (&O)[P])
{
    constexpr int N = 2 * P + 2 * WV::HALO;
#pragma unroll
    for (int q = 0; q < N; q += 2) {
        w[q] = WV::ise(w[q]);
        w[q + 1] = WV::iso(w[q + 1]);
    }
    lift_range_inv<WV, 0, N>(w);
    lift_range_inv<WV, 1, N>(w);
    if constexpr (WV::NS == 4) {
        lift_range_inv<WV, 2, N>(w);
        lift_range_inv<WV, 3, N>(w);
    }
#pragma unroll
    for (int i = 0; i < P; i++) {
        E[i] = w[WV::HALO + 2 * i];
        O[i] = w[WV::HALO + 1 + 2 * i];
    }
}

}  // namespace dwtb200
