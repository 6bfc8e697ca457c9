// kernels_stream.cu -- the hot kernels: one decomposition level per launch, row and column lifting
// fused in registers (sm_100a).
//
// Replaces, per level, the two OpenMP line loops of the reference drivers
// (/root/reference/src/libdwt.c:12837-12893, 17098-17154, 16342-16361, 18178-18195) and the
// gather / 4 sweeps / scatter of every 1-D line call beneath them.
//
// Shape of the computation (forward; the inverse is its mirror image):
//   * a WARP owns a column group of 30*VPL output columns and a strip of output row pairs; lanes 0
//     and 31 are halo lanes (they hold the lifting-depth overlap with the neighbouring groups and
//     store nothing), so every lane runs the same instruction stream and every 16-byte load/store
//     of the 30 producing lanes is aligned: 32*VPL columns loaded, 30*VPL produced (6.25 % overlap,
//     served by L1/L2 because neighbouring groups run in neighbouring warps);
//   * the warp streams DOWN its strip two rows at a time.  Row lifting happens in registers with one
//     __shfl per lifting step for the +-1 neighbour in the adjacent lane; column lifting is a
//     register pipeline carried from row pair to row pair (NS state values per column), so a
//     sample is read from HBM once and each of the four subband samples is written once -- no
//     shared memory, no intermediate plane;
//   * image borders: whole-sample mirror on the row index of the load and, for groups touching the
//     left/right border, per-lane mirrored column indices (warp-uniform slow path);
//   * the next row pair is prefetched into registers while the current one is lifted.
// Bit-exactness: every sample still sees "row lifting + scale, then column lifting + scale" in the
// reference's operation order (int inverse: columns first), only the schedule differs.
#include "stream_common.cuh"

namespace dwtb200 {

// =====================================================================================================
// forward level
// =====================================================================================================
template <class WV, int VPL, int PFD = 1>
__global__ void __launch_bounds__(128, PFD == 2 ? 3 : (VPL * sizeof(typename WV::T) >= 32) ? 4 : 8) k_fwd_level(const LevelParams p)
{
    using T = typename WV::T;
    constexpr int OUTW = 30 * VPL, HV = VPL / 2;
    pdl_begin();
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (gw >= p.ncg * p.nstrips) return;
    const int cg = gw % p.ncg, strip = gw / p.ncg + p.strip0;
    const int xl = cg * OUTW - VPL + lane * VPL;   // first column held by this lane (even)
    const int k0 = strip * p.pps, k1 = min(k0 + p.pps, p.nLy);
    const int W = p.W, H = p.H;

    const T *src = (const T *)p.src + (int64_t)blockIdx.y * p.src_frame;
    const bool fast = __all_sync(FULL, xl >= 0 && xl + VPL <= W);
    int cx[VPL];
    if (!fast) {
#pragma unroll
        for (int i = 0; i < VPL; i++) cx[i] = reflect(xl + i, W);
    }
    auto load = [&](int r, T(&v)[VPL]) {
        const T *rp = src + (int64_t)reflect(r, H) * p.src_pitch;
        if (fast) {
            ld_vec<T, VPL>(rp + xl, v);
        } else {
#pragma unroll
            for (int i = 0; i < VPL; i++) v[i] = __ldg(rp + cx[i]);
        }
    };

    // destinations: even columns -> L half (ll / lh), odd columns -> H half (hl / hh)
    const int cb = xl >> 1;
    const bool producer = lane >= 1 && lane <= 30 && xl < W;
    const bool whole = xl + VPL <= W;   // all VPL/2 L and H columns exist
    T *ll = (T *)p.ll + (int64_t)blockIdx.y * p.ll_frame + cb;
    T *hl = (T *)p.hl + (int64_t)blockIdx.y * p.sub_frame + cb;
    T *lh = (T *)p.lh + (int64_t)blockIdx.y * p.sub_frame + cb;
    T *hh = (T *)p.hh + (int64_t)blockIdx.y * p.sub_frame + cb;
    const bool vec_sub = whole && p.sub_aligned;
    auto put = [&](T *base, int64_t pitch, int row, const T(&o)[HV], int limit, bool vec) {
        T *q = base + (int64_t)row * pitch;
        if (vec) {
            st_vec<T, HV>(q, o);
        } else {
#pragma unroll
            for (int i = 0; i < HV; i++)
                if (cb + i < limit) q[i] = o[i];
        }
    };

    constexpr int NSTATE = WV::NS;   // carried values per column
    T st[NSTATE][VPL];               // NS==4: xe, d1, s1, d2     NS==2: xe, d1
    T a[VPL], b[VPL], na[VPL], nb[VPL];

    constexpr int WARM = WV::NS / 2 + (WV::NS == 4 ? 1 : 0);   // warm-up iterations: 3 (9/7) or 1 (5/3)
    constexpr int DELAY = WV::NS / 2 - 1;                      // iteration m emits pair m - DELAY
    const int m0 = k0 + DELAY - WARM;                          // first iteration
    const int m1 = k1 - 1 + DELAY;                             // last iteration (inclusive)

    load(2 * m0, st[0]);
    hfwd<WV, VPL>(st[0]);
    load(2 * m0 + 1, na);
    load(2 * m0 + 2, nb);
    T n2a[PFD == 2 ? VPL : 1], n2b[PFD == 2 ? VPL : 1];   // second pair in flight (PFD == 2)
    if constexpr (PFD == 2) {
        if (m0 < m1) {
            load(2 * m0 + 3, n2a);
            load(2 * m0 + 4, n2b);
        }
    }
#pragma unroll
    for (int s = 1; s < NSTATE; s++)
#pragma unroll
        for (int i = 0; i < VPL; i++) st[s][i] = T(0);

    for (int m = m0; m <= m1; m++) {
#pragma unroll
        for (int i = 0; i < VPL; i++) {
            a[i] = na[i];
            b[i] = nb[i];
        }
        if constexpr (PFD == 2) {
#pragma unroll
            for (int i = 0; i < VPL; i++) {
                na[i] = n2a[i];
                nb[i] = n2b[i];
            }
            if (m + 1 < m1) {   // two pairs ahead
                load(2 * m + 5, n2a);
                load(2 * m + 6, n2b);
            }
        } else if (m < m1) {   // prefetch the next pair while this one is lifted
            load(2 * m + 3, na);
            load(2 * m + 4, nb);
        }
#ifdef DWTB200_DEBUG_KEYS
        if (p.dbg != 2)
#endif
        {
            hfwd<WV, VPL>(a);
            hfwd<WV, VPL>(b);
        }

        T oL[VPL], oH[VPL];   // column-lifted low / high outputs of this iteration
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
#ifdef DWTB200_DEBUG_KEYS   // measurement only (profiles/dbg_level0.py): 2 = no lifting arithmetic, 1 = no stores
        if (p.dbg == 2) {
#pragma unroll
            for (int i = 0; i < VPL; i++) {
                oL[i] = a[i];
                oH[i] = b[i];
            }
        }
#endif
        const int kk = m - DELAY;
#ifdef DWTB200_DEBUG_KEYS
        if (p.dbg == 1 && oL[0] != T(123457)) continue;   // the compare keeps the arithmetic alive
#endif
        if (kk >= k0 && producer) {
            T o[HV];
#pragma unroll
            for (int i = 0; i < HV; i++) o[i] = oL[2 * i];
            put(ll, p.ll_pitch, kk, o, p.nLx, whole);
#pragma unroll
            for (int i = 0; i < HV; i++) o[i] = oL[2 * i + 1];
            put(hl, p.sub_pitch, kk, o, p.nHx, vec_sub);
            if (kk < p.nHy) {
#pragma unroll
                for (int i = 0; i < HV; i++) o[i] = oH[2 * i];
                put(lh, p.sub_pitch, kk, o, p.nLx, whole);
#pragma unroll
                for (int i = 0; i < HV; i++) o[i] = oH[2 * i + 1];
                put(hh, p.sub_pitch, kk, o, p.nHx, vec_sub);
            }
        }
    }
}

// =====================================================================================================
// inverse level
// =====================================================================================================
template <class WV, int VPL> __global__ void __launch_bounds__(128, (VPL * sizeof(typename WV::T) >= 32) ? 4 : 8) k_inv_level(const LevelParams p)
{
    using T = typename WV::T;
    constexpr int OUTW = 30 * VPL, HV = VPL / 2;
    pdl_begin();
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (gw >= p.ncg * p.nstrips) return;
    const int cg = gw % p.ncg, strip = gw / p.ncg + p.strip0;
    const int xl = cg * OUTW - VPL + lane * VPL;   // first OUTPUT column of this lane (even)
    const int cb = xl >> 1;                        // first subband column
    const int W = p.W, H = p.H;

    const T *ll = (const T *)p.ll + (int64_t)blockIdx.y * p.ll_frame;
    const T *hl = (const T *)p.hl + (int64_t)blockIdx.y * p.sub_frame;
    const T *lh = (const T *)p.lh + (int64_t)blockIdx.y * p.sub_frame;
    const T *hh = (const T *)p.hh + (int64_t)blockIdx.y * p.sub_frame;
    T *dst = (T *)p.dst + (int64_t)blockIdx.y * p.dst_frame;

    const bool inside = xl >= 0 && xl + VPL <= W;
    const bool fast = __all_sync(FULL, inside);
    const bool fast_sub = fast && p.sub_aligned;
    int cL[HV], cH[HV];   // mirrored subband columns for the border path
    if (!fast_sub) {
#pragma unroll
        for (int i = 0; i < HV; i++) {
            cL[i] = reflect(xl + 2 * i, W) >> 1;
            cH[i] = reflect(xl + 2 * i + 1, W) >> 1;
        }
    }
    // one interleaved row: even positions from `lo` (L half), odd positions from `hi` (H half)
    auto load = [&](const T *lo, int64_t lpitch, const T *hi, int64_t hpitch, int row, T(&v)[VPL]) {
        const T *lp = lo + (int64_t)row * lpitch, *hp = hi + (int64_t)row * hpitch;
        T l[HV], h[HV];
        if (fast) {
            ld_vec<T, HV>(lp + cb, l);
        } else {
#pragma unroll
            for (int i = 0; i < HV; i++) l[i] = __ldg(lp + cL[i]);
        }
        if (fast_sub) {
            ld_vec<T, HV>(hp + cb, h);
        } else if (fast) {
#pragma unroll
            for (int i = 0; i < HV; i++) h[i] = __ldg(hp + cb + i);
        } else {
#pragma unroll
            for (int i = 0; i < HV; i++) h[i] = __ldg(hp + cH[i]);
        }
#pragma unroll
        for (int i = 0; i < HV; i++) {
            v[2 * i] = l[i];
            v[2 * i + 1] = h[i];
        }
    };
    // iteration k consumes interleaved rows 2k (an L row) and 2k+1 (an H row), mirrored
    auto load_pair = [&](int k, T(&a)[VPL], T(&b)[VPL]) {
        const int ra = reflect(2 * k, H) >> 1, rb = reflect(2 * k + 1, H) >> 1;
        load(ll, p.ll_pitch, hl, p.sub_pitch, ra, a);
        load(lh, p.sub_pitch, hh, p.sub_pitch, rb, b);
    };
    const bool producer = lane >= 1 && lane <= 30 && xl < W;
    const bool whole = xl + VPL <= W;
    auto put = [&](int row, const T(&o)[VPL]) {
        if (row < 0 || row >= H || !producer) return;
        T *q = dst + (int64_t)row * p.dst_pitch + xl;
        if (whole) {
            st_vec<T, VPL>(q, o);
        } else {
#pragma unroll
            for (int i = 0; i < VPL; i++)
                if (xl + i < W) q[i] = o[i];
        }
    };

    constexpr int NSTATE = WV::NS;
    T st[NSTATE][VPL];   // NS==4: d2p, s1p, d1p, xep      NS==2: cp, xep
    T a[VPL], b[VPL], na[VPL], nb[VPL];
    constexpr int DELAY = WV::NS / 2 - 1;   // iteration k emits rows 2(k-DELAY)-1 and 2(k-DELAY)
    constexpr int WARM = WV::NS;            // warm-up iterations
    const int q0 = strip * p.pps, q1 = min(q0 + p.pps, (H >> 1) + 1);   // emitted q = k - DELAY in [q0, q1)
    const int ka = q0 + DELAY - WARM, kb = q1 - 1 + DELAY;

#pragma unroll
    for (int s = 0; s < NSTATE; s++)
#pragma unroll
        for (int i = 0; i < VPL; i++) st[s][i] = T(0);
    load_pair(ka, na, nb);

    for (int k = ka; k <= kb; k++) {
#pragma unroll
        for (int i = 0; i < VPL; i++) {
            a[i] = na[i];
            b[i] = nb[i];
        }
        if (k < kb) load_pair(k + 1, na, nb);
        if constexpr (!WV::INV_COLS_FIRST) {   // rows first (float / double): libdwt.c:17098 then 17127
            hinv<WV, VPL>(a);
            hinv<WV, VPL>(b);
        }
        T oO[VPL], oE[VPL];   // output rows 2q-1 (odd) and 2q (even)
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
        const int q = k - DELAY;
        if (q >= q0) {   // warp-uniform
            if constexpr (WV::INV_COLS_FIRST) {   // columns first (int): libdwt.c:18178 then 18187
                hinv<WV, VPL>(oO);
                hinv<WV, VPL>(oE);
            }
            put(2 * q - 1, oO);
            put(2 * q, oE);
        }
    }
}

// ---- launchers -------------------------------------------------------------------------------
template <class WV, int VPL> static void go_fwd(const LevelParams &p, int frames, cudaStream_t st)
{
    const int warps = p.ncg * p.nstrips;
    const dim3 grid((warps + 3) / 4, frames);
    if (p.pfd == 2 && VPL * sizeof(typename WV::T) >= 32) launch_pdl(k_fwd_level<WV, VPL, 2>, grid, dim3(128), 0, st, g_use_pdl, p);
    else launch_pdl(k_fwd_level<WV, VPL>, grid, dim3(128), 0, st, g_use_pdl, p);
}
template <class WV, int VPL> static void go_inv(const LevelParams &p, int frames, cudaStream_t st)
{
    const int warps = p.ncg * p.nstrips;
    const dim3 grid((warps + 3) / 4, frames);
    launch_pdl(k_inv_level<WV, VPL>, grid, dim3(128), 0, st, g_use_pdl, p);
}
template <class K> static cudaError_t touch(K kern)
{
    cudaFuncAttributes a;
    return cudaFuncGetAttributes(&a, kern);
}
cudaError_t preload_stream()
{
    cudaError_t e = cudaSuccess;
    for (int kind = 0; kind < K_COUNT; kind++)
        dispatch_kind(kind, [&](auto wv) {
            using WV = decltype(wv);
            constexpr int V = 32 / (int)sizeof(typename WV::T);
            if (e == cudaSuccess) e = touch(k_fwd_level<WV, V>);
            if (e == cudaSuccess) e = touch(k_inv_level<WV, V>);
            if (e == cudaSuccess) e = touch(k_fwd_level<WV, V, 2>);
            if (e == cudaSuccess) e = touch(k_fwd_level<WV, 4>);
            if (e == cudaSuccess) e = touch(k_inv_level<WV, 4>);
        });
    return e;
}
// two widths per type: 32 bytes per lane (16 warps per SM, fewest instructions per sample) and 16 bytes
// per lane (half the registers -> 32 warps per SM, more latency hiding); p.narrow selects.  8-byte samples
// always take 4 per lane (a lane must hold >= HALO samples).
int stream_out_width(int kind, int narrow) { return kind_elem_size(kind) == 8 ? 30 * 4 : (30 * 8) >> (narrow ? 1 : 0); }
int stream_warps_per_sm(int kind, int narrow, int pfd) { return (narrow && kind_elem_size(kind) != 8) ? 32 : pfd == 2 ? 12 : 16; }

void launch_fwd_level(int kind, const LevelParams &p, int frames, cudaStream_t st)
{
    dispatch_kind(kind, [&](auto wv) {
        using WV = decltype(wv);
        if (p.narrow) go_fwd<WV, 4>(p, frames, st);
        else go_fwd<WV, 32 / (int)sizeof(typename WV::T)>(p, frames, st);
    });
}
void launch_inv_level(int kind, const LevelParams &p, int frames, cudaStream_t st)
{
    dispatch_kind(kind, [&](auto wv) {
        using WV = decltype(wv);
        if (p.narrow) go_inv<WV, 4>(p, frames, st);
        else go_inv<WV, 32 / (int)sizeof(typename WV::T)>(p, frames, st);
    });
}

}  // namespace dwtb200
