// tail_body.cuh -- all remaining coarse levels of ONE frame inside one CTA's shared memory.
// Device bodies of the tail kernels (kernels_tail.cu).  Semantics: see kernels_tail.cu.
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

constexpr int TAIL_P = 4;            // output pairs per window evaluation on the fast paths
constexpr int TAIL_FAST_MIN = 16;    // lines at least this long take them (>= 2 * TAIL_P, so a shifted last group exists)
// row pitch of a w-wide band in shared memory: odd for the bands the fast paths walk, so that neither a walk down a
// column nor one thread per row hits a single bank (costs at most 1/16 more room: see tail_max_elems)
__host__ __device__ inline int tail_pitch(int w) { return w >= TAIL_FAST_MIN ? (w | 1) : w; }

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
    // tasks are indexed (tx, ty) = (lane, warp) by nested strided loops: a flat index split by / and % of a run-time size costs an
    // integer division (a ~150-cycle dependent chain) per task in passes that have only a few hundred cycles of work
    const int tx = tid & 31, ty = tid >> 5, nty = TAIL_THREADS >> 5;

    // Lines shorter than TAIL_FAST_MIN are evaluated pair by pair from a mirrored 2 * HALO + 2 window: folding every tap index
    // (lifting.cuh: reflect, a loop when the line is shorter than the window) made each of the last levels cost ~1.5 us of
    // dependent integer work on a handful of threads.  The folded indices of a level come from a table instead, built once per
    // level by as many threads as it has entries: tap q of pair k is tab[2 k + q] = reflect(2 k + q - HALO, n).
    __shared__ short tabx[TAIL_FAST_MIN + 2 * WV::HALO + 2], taby[TAIL_FAST_MIN + 2 * WV::HALO + 2];
    int w = cdiv_pow2(p.W0, p.j0), h = cdiv_pow2(p.H0, p.j0);
    for (int y = ty; y < h; y += nty)
        for (int x = tx; x < w; x += 32) bufA[y * tail_pitch(w) + x] = ld(src + (int64_t)y * p.src_pitch + x);
    T *in = bufA, *other = bufB;
    __syncthreads();

    for (int j = p.j0; j < p.j1; j++) {
        const int nlx = (w + 1) >> 1, nhx = w >> 1, nly = (h + 1) >> 1, nhy = h >> 1;
        const int pw = tail_pitch(w), pn = tail_pitch(nlx);   // pitches of this level's band and of the next one
        if ((w < TAIL_FAST_MIN && w >= 2) || (h < TAIL_FAST_MIN && h >= 2)) {
            if (w < TAIL_FAST_MIN && w >= 2 && tid < w + 2 * WV::HALO + 2) tabx[tid] = (short)reflect(tid - WV::HALO, w);
            if (h < TAIL_FAST_MIN && h >= 2 && tid >= 64 && tid - 64 < h + 2 * WV::HALO + 2) taby[tid - 64] = (short)reflect(tid - 64 - WV::HALO, h);
            __syncthreads();
        }
        // ---- rows: in -> other, row y = [L (nlx) | H (nhx)] ----
        T *rows = in;
        if (!(WV::GUARD && w <= 1)) {
            rows = other;
            if (w >= TAIL_FAST_MIN) {
                // P pairs per window evaluation (lifting.cuh), one integer division per task; the last group of a line
                // is shifted back so that it ends with the line (recomputes a few pairs, same values)
                constexpr int P = TAIL_P, NT = 2 * P + 2 * WV::HALO;
                const int ng = (nlx + P - 1) / P;
                for (int wt = ty; wt < ng * ((h + 31) >> 5); wt += nty) {   // a warp per (group of pairs, block of 32 lines): no division per task
                    int g = 0, yb = wt;
                    while (yb >= ((h + 31) >> 5)) {
                        yb -= (h + 31) >> 5;
                        g++;
                    }
                    const int y = (yb << 5) + tx;   // consecutive threads walk down a column of windows
                    if (y >= h) continue;
                    int k = g * P;
                    if (k + P > nlx) k = nlx - P;
                    const T *line = in + y * pw;
                    T win[NT], L[P], H[P];
                    const int t0 = 2 * k - WV::HALO;
                    if (t0 >= 0 && t0 + NT <= w) {
#pragma unroll
                        for (int i = 0; i < NT; i++) win[i] = line[t0 + i];
                    } else {
#pragma unroll
                        for (int i = 0; i < NT; i++) win[i] = line[reflect(t0 + i, w)];
                    }
                    window_fwd_p<WV, P>(win, L, H);
#pragma unroll
                    for (int i = 0; i < P; i++) {
                        rows[y * pw + k + i] = L[i];
                        if (k + i < nhx) rows[y * pw + nlx + k + i] = H[i];
                    }
                }
            } else if (w >= 2) {
                for (int y = ty; y < h; y += nty)
                for (int k = tx; k < nlx; k += 32) {
                    const T *line = in + y * pw;
                    T win[2 * WV::HALO + 2];
#pragma unroll
                    for (int i = 0; i < 2 * WV::HALO + 2; i++) win[i] = line[tabx[2 * k + i]];
                    T L, H;
                    window_fwd<WV>(win, L, H);
                    rows[y * pw + k] = L;
                    if (k < nhx) rows[y * pw + nlx + k] = H;
                }
            } else {
                for (int t = tid; t < h; t += TAIL_THREADS) rows[t * pw] = one_fwd<WV>(in[t * pw]);
            }
            __syncthreads();
        }
        // ---- columns: rows -> LL' (smem) + HL/LH/HH (global, Mallat positions) ----
        T *next = (rows == in) ? other : in;
        if (!(WV::GUARD && h <= 1) && h >= TAIL_FAST_MIN) {
            constexpr int P = TAIL_P, NT = 2 * P + 2 * WV::HALO;
            const int ng = (nly + P - 1) / P;
            for (int wt = ty; wt < ng * ((w + 31) >> 5); wt += nty) {
                int g = 0, xb = wt;
                while (xb >= ((w + 31) >> 5)) {
                    xb -= (w + 31) >> 5;
                    g++;
                }
                const int x = (xb << 5) + tx;
                if (x >= w) continue;
                int k = g * P;
                if (k + P > nly) k = nly - P;
                T win[NT], L[P], H[P];
                const int t0 = 2 * k - WV::HALO;
                if (t0 >= 0 && t0 + NT <= h) {
#pragma unroll
                    for (int i = 0; i < NT; i++) win[i] = rows[(t0 + i) * pw + x];
                } else {
#pragma unroll
                    for (int i = 0; i < NT; i++) win[i] = rows[reflect(t0 + i, h) * pw + x];
                }
                window_fwd_p<WV, P>(win, L, H);
#pragma unroll
                for (int i = 0; i < P; i++) {
                    if (x < nlx) next[(k + i) * pn + x] = L[i];
                    else dst[(int64_t)(k + i) * p.dst_pitch + x] = L[i];
                    if (k + i < nhy) dst[(int64_t)(nly + k + i) * p.dst_pitch + x] = H[i];
                }
            }
        } else if (!(WV::GUARD && h <= 1) && h >= 2) {
            for (int k = ty; k < nly; k += nty)
            for (int x = tx; x < w; x += 32) {
                T win[2 * WV::HALO + 2];
#pragma unroll
                for (int i = 0; i < 2 * WV::HALO + 2; i++) win[i] = rows[taby[2 * k + i] * pw + x];
                T L, H;
                window_fwd<WV>(win, L, H);
                if (x < nlx) next[k * pn + x] = L;
                else dst[(int64_t)k * p.dst_pitch + x] = L;
                if (k < nhy) dst[(int64_t)(nly + k) * p.dst_pitch + x] = H;
            }
        } else {   // h == 1 (or guarded): the single row passes through
            for (int y = ty; y < h; y += nty)
            for (int x = tx; x < w; x += 32) {
                T v = rows[y * pw + x];
                if (!(WV::GUARD && h <= 1)) v = one_fwd<WV>(v);
                if (x < nlx) next[y * pn + x] = v;
                else dst[(int64_t)y * p.dst_pitch + x] = v;
            }
        }
        __syncthreads();
        in = next;
        other = (next == bufA) ? bufB : bufA;
        w = nlx;
        h = nly;
    }
    for (int y = ty; y < h; y += nty)
        for (int x = tx; x < w; x += 32) dst[(int64_t)y * p.dst_pitch + x] = in[y * tail_pitch(w) + x];
}

// P output pairs k .. k+P-1 of one inverse line of n samples: tap(c) returns interleaved coefficient c, emit(q, E, O)
// receives samples 2q and 2q+1
template <class WV, int P, class TAP, class EMIT> __device__ __forceinline__ void inv_group(int k, int n, TAP tap, EMIT emit)
{
    using T = typename WV::T;
    constexpr int NT = 2 * P + 2 * WV::HALO;
    T win[NT], E[P], O[P];
    const int t0 = 2 * k - WV::HALO;
    if (t0 >= 0 && t0 + NT <= n) {
#pragma unroll
        for (int i = 0; i < NT; i++) win[i] = tap(t0 + i);
    } else {
#pragma unroll
        for (int i = 0; i < NT; i++) win[i] = tap(reflect(t0 + i, n));
    }
    window_inv_p<WV, P>(win, E, O);
#pragma unroll
    for (int i = 0; i < P; i++) emit(k + i, E[i], O[i]);
}

// `mal` != nullptr: a third shared-memory buffer that receives the whole Mallat block of the tail (all its subbands, w0 x h0 samples
// in the top left corner of the plane) in ONE load phase at the start, instead of one dependent round trip to L2 per level
template <class WV, class LD>
__device__ __forceinline__ void inv_tail_body(const TailParams &p, int frame, typename WV::T *bufA, typename WV::T *bufB, LD ld,
                                              typename WV::T *mal = nullptr)
{
    using T = typename WV::T;
    const T *src = (const T *)p.src + (int64_t)frame * p.src_frame;
    T *dst = (T *)p.dst + (int64_t)frame * p.dst_frame;
    const int tid = threadIdx.x;
    const int TAIL_THREADS = blockDim.x;
    // tasks are indexed (tx, ty) = (lane, warp) by nested strided loops: a flat index split by / and % of a run-time size costs an
    // integer division (a ~150-cycle dependent chain) per task in passes that have only a few hundred cycles of work
    const int tx = tid & 31, ty = tid >> 5, nty = TAIL_THREADS >> 5;
    const int w0 = cdiv_pow2(p.W0, p.j0), pm = tail_pitch(w0);
    __shared__ short itabx[TAIL_FAST_MIN + 2 * WV::HALO + 2], itaby[TAIL_FAST_MIN + 2 * WV::HALO + 2];
    if (mal) {
        const int h0 = cdiv_pow2(p.H0, p.j0);
        for (int y = ty; y < h0; y += nty)
            for (int x = tx; x < w0; x += 32) mal[y * pm + x] = ld(src + (int64_t)y * p.src_pitch + x);
        __syncthreads();
    }

    {   // coarsest LL band
        const int w = cdiv_pow2(p.W0, p.j1), h = cdiv_pow2(p.H0, p.j1);
        for (int y = ty; y < h; y += nty)
            for (int x = tx; x < w; x += 32) bufA[y * tail_pitch(w) + x] = mal ? mal[y * pm + x] : ld(src + (int64_t)y * p.src_pitch + x);
    }
    T *in = bufA, *tmp = bufB;
    __syncthreads();

    for (int j = p.j1; j > p.j0; j--) {
        const int w = cdiv_pow2(p.W0, j - 1), h = cdiv_pow2(p.H0, j - 1);   // size being reconstructed
        const int nlx = (w + 1) >> 1, nly = (h + 1) >> 1;
        const int pw = tail_pitch(w), pl = tail_pitch(nlx);   // pitches of the band being reconstructed and of the LL band
        // Mallat-arranged input of this level: LL from shared memory, the rest from the plane
        auto M = [&](int y, int x) -> T {
            return (y < nly && x < nlx) ? in[y * pl + x] : (mal ? mal[y * pm + x] : ld(src + (int64_t)y * p.src_pitch + x));
        };
        const bool do_rows = !(WV::GUARD && w <= 1), do_cols = !(WV::GUARD && h <= 1);
        // short lines: the Mallat position of every (mirrored) tap from a table, see fwd_tail_body
        if ((w < TAIL_FAST_MIN && w >= 2) || (h < TAIL_FAST_MIN && h >= 2)) {
            if (w < TAIL_FAST_MIN && w >= 2 && tid < w + 2 * WV::HALO + 2) {
                const int c = reflect(tid - WV::HALO, w);
                itabx[tid] = (short)((c & 1) ? nlx + (c >> 1) : (c >> 1));
            }
            if (h < TAIL_FAST_MIN && h >= 2 && tid >= 64 && tid - 64 < h + 2 * WV::HALO + 2) {
                const int c = reflect(tid - 64 - WV::HALO, h);
                itaby[tid - 64] = (short)((c & 1) ? nly + (c >> 1) : (c >> 1));
            }
            __syncthreads();
        }
        if constexpr (!WV::INV_COLS_FIRST) {
            // rows: M -> tmp (h x w, rows still in Mallat order)
            if (do_rows && w >= TAIL_FAST_MIN) {
                const int ng = (nlx + TAIL_P - 1) / TAIL_P;
                for (int wt = ty; wt < ng * ((h + 31) >> 5); wt += nty) {   // a warp per (group of pairs, block of 32 lines): no division per task
                    int g = 0, yb = wt;
                    while (yb >= ((h + 31) >> 5)) {
                        yb -= (h + 31) >> 5;
                        g++;
                    }
                    const int y = (yb << 5) + tx;   // consecutive threads walk down a column of windows
                    if (y >= h) continue;
                    const int k = (g + 1) * TAIL_P > nlx ? nlx - TAIL_P : g * TAIL_P;
                    inv_group<WV, TAIL_P>(
                        k, w, [&](int c) { return M(y, (c & 1) ? nlx + (c >> 1) : (c >> 1)); },
                        [&](int q, T E, T O) {
                            tmp[y * pw + 2 * q] = E;
                            if (2 * q + 1 < w) tmp[y * pw + 2 * q + 1] = O;
                        });
                }
            } else if (do_rows && w >= 2) {
                for (int y = ty; y < h; y += nty)
                for (int k = tx; k < nlx; k += 32) {
                    T win[2 * WV::HALO + 2];
#pragma unroll
                    for (int i = 0; i < 2 * WV::HALO + 2; i++) win[i] = M(y, itabx[2 * k + i]);
                    T E, O;
                    window_inv<WV>(win, E, O);
                    tmp[y * pw + 2 * k] = E;
                    if (2 * k + 1 < w) tmp[y * pw + 2 * k + 1] = O;
                }
            } else {
                for (int y = ty; y < h; y += nty)
                for (int x = tx; x < w; x += 32) {
                    const T v = M(y, x);
                    tmp[y * pw + x] = do_rows ? one_inv<WV>(v) : v;
                }
            }
            __syncthreads();
            // columns: tmp -> out (next LL in shared memory, or the destination plane at the end)
            T *out = in;   // `in` is dead once the row pass has finished
            const bool last = (j - 1 == p.j0);
            if (do_cols && h >= TAIL_FAST_MIN) {
                const int ng = (nly + TAIL_P - 1) / TAIL_P;
                for (int wt = ty; wt < ng * ((w + 31) >> 5); wt += nty) {
                    int g = 0, xb = wt;
                    while (xb >= ((w + 31) >> 5)) {
                        xb -= (w + 31) >> 5;
                        g++;
                    }
                    const int x = (xb << 5) + tx;
                    if (x >= w) continue;
                    const int k = (g + 1) * TAIL_P > nly ? nly - TAIL_P : g * TAIL_P;
                    inv_group<WV, TAIL_P>(
                        k, h, [&](int c) { return tmp[((c & 1) ? nly + (c >> 1) : (c >> 1)) * pw + x]; },
                        [&](int q, T E, T O) {
                            if (last) {
                                dst[(int64_t)(2 * q) * p.dst_pitch + x] = E;
                                if (2 * q + 1 < h) dst[(int64_t)(2 * q + 1) * p.dst_pitch + x] = O;
                            } else {
                                out[(2 * q) * pw + x] = E;
                                if (2 * q + 1 < h) out[(2 * q + 1) * pw + x] = O;
                            }
                        });
                }
            } else
            for (int k = ty; k < nly; k += nty)
            for (int x = tx; x < w; x += 32) {
                T E, O = T(0);
                if (do_cols && h >= 2) {
                    T win[2 * WV::HALO + 2];
#pragma unroll
                    for (int i = 0; i < 2 * WV::HALO + 2; i++) win[i] = tmp[itaby[2 * k + i] * pw + x];
                    window_inv<WV>(win, E, O);
                } else {
                    E = do_cols ? one_inv<WV>(tmp[x]) : tmp[x];   // h == 1
                }
                if (last) {
                    dst[(int64_t)(2 * k) * p.dst_pitch + x] = E;
                    if (2 * k + 1 < h) dst[(int64_t)(2 * k + 1) * p.dst_pitch + x] = O;
                } else {
                    out[(2 * k) * pw + x] = E;
                    if (2 * k + 1 < h) out[(2 * k + 1) * pw + x] = O;
                }
            }
            __syncthreads();
        } else {
            // columns first (int 5/3): M -> tmp (rows de-interleaved, columns still in Mallat order)
            if (h >= TAIL_FAST_MIN) {
                const int ng = (nly + TAIL_P - 1) / TAIL_P;
                for (int wt = ty; wt < ng * ((w + 31) >> 5); wt += nty) {
                    int g = 0, xb = wt;
                    while (xb >= ((w + 31) >> 5)) {
                        xb -= (w + 31) >> 5;
                        g++;
                    }
                    const int x = (xb << 5) + tx;
                    if (x >= w) continue;
                    const int k = (g + 1) * TAIL_P > nly ? nly - TAIL_P : g * TAIL_P;
                    inv_group<WV, TAIL_P>(
                        k, h, [&](int c) { return M((c & 1) ? nly + (c >> 1) : (c >> 1), x); },
                        [&](int q, T E, T O) {
                            tmp[(2 * q) * pw + x] = E;
                            if (2 * q + 1 < h) tmp[(2 * q + 1) * pw + x] = O;
                        });
                }
            } else
            for (int k = ty; k < nly; k += nty)
            for (int x = tx; x < w; x += 32) {
                if (h >= 2) {
                    T win[2 * WV::HALO + 2];
#pragma unroll
                    for (int i = 0; i < 2 * WV::HALO + 2; i++) win[i] = M(itaby[2 * k + i], x);
                    T E, O;
                    window_inv<WV>(win, E, O);
                    tmp[(2 * k) * pw + x] = E;
                    if (2 * k + 1 < h) tmp[(2 * k + 1) * pw + x] = O;
                } else {
                    tmp[x] = one_inv<WV>(M(0, x));   // h == 1: row 0
                }
            }
            __syncthreads();
            T *out = in;
            const bool last = (j - 1 == p.j0);
            if (w >= TAIL_FAST_MIN) {
                const int ng = (nlx + TAIL_P - 1) / TAIL_P;
                for (int wt = ty; wt < ng * ((h + 31) >> 5); wt += nty) {   // a warp per (group of pairs, block of 32 lines): no division per task
                    int g = 0, yb = wt;
                    while (yb >= ((h + 31) >> 5)) {
                        yb -= (h + 31) >> 5;
                        g++;
                    }
                    const int y = (yb << 5) + tx;   // consecutive threads walk down a column of windows
                    if (y >= h) continue;
                    const int k = (g + 1) * TAIL_P > nlx ? nlx - TAIL_P : g * TAIL_P;
                    inv_group<WV, TAIL_P>(
                        k, w, [&](int c) { return tmp[y * pw + ((c & 1) ? nlx + (c >> 1) : (c >> 1))]; },
                        [&](int q, T E, T O) {
                            if (last) {
                                dst[(int64_t)y * p.dst_pitch + 2 * q] = E;
                                if (2 * q + 1 < w) dst[(int64_t)y * p.dst_pitch + 2 * q + 1] = O;
                            } else {
                                out[y * pw + 2 * q] = E;
                                if (2 * q + 1 < w) out[y * pw + 2 * q + 1] = O;
                            }
                        });
                }
            } else
            for (int y = ty; y < h; y += nty)
            for (int k = tx; k < nlx; k += 32) {
                T E, O = T(0);
                if (w >= 2) {
                    T win[2 * WV::HALO + 2];
#pragma unroll
                    for (int i = 0; i < 2 * WV::HALO + 2; i++) win[i] = tmp[y * pw + itabx[2 * k + i]];
                    window_inv<WV>(win, E, O);
                } else {
                    E = one_inv<WV>(tmp[y * pw]);
                }
                if (last) {
                    dst[(int64_t)y * p.dst_pitch + 2 * k] = E;
                    if (2 * k + 1 < w) dst[(int64_t)y * p.dst_pitch + 2 * k + 1] = O;
                } else {
                    out[y * pw + 2 * k] = E;
                    if (2 * k + 1 < w) out[y * pw + 2 * k + 1] = O;
                }
            }
            __syncthreads();
        }
    }
    if (p.j1 == p.j0) {   // nothing to do: pass the band through
        const int w = cdiv_pow2(p.W0, p.j0), h = cdiv_pow2(p.H0, p.j0);
        for (int y = ty; y < h; y += nty)
            for (int x = tx; x < w; x += 32) dst[(int64_t)y * p.dst_pitch + x] = in[y * tail_pitch(w) + x];
    }
}

}  // namespace dwtb200
