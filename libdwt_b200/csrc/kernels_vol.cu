// kernels_vol.cu -- one-level 3-D CDF 9/7 (float) on a volume, interleaved subbands (sm_100a).
//
// Replaces cdf97_3f_op_sep_horizontal_s / cdf97_3f_ip_sep_horizontal_s (/root/reference/src/volume-dwt.c:727,
// 677; the 1-D kernel is fdwt1_single_cdf97_horizontal_min5_s, src/dwt-simple.c:2166) and
// cdf97_3i_ip_sep_horizontal_s (src/volume-dwt.c:1115 -> dwt_cdf97_1i_inplace_s, src/libdwt.c:17182).
// The reference lifts along x for every (y, z), then along y, then along z -- three passes over the
// volume through the cache hierarchy, the z pass with a 4 MiB stride.  Here:
//   k_vol_xy   x and y lifting of one slice fused in registers (the streaming scheme of kernels_stream.cu:
//              a warp owns 30*VPL columns and streams down a strip of row pairs), slices are independent
//              -> blockIdx.y = z; subbands stay interleaved, so rows are stored where they were read
//   k_vol_z    z lifting: a warp owns 32*VPL columns of one row y and streams through a strip of slice
//              pairs with the same register pipeline (no horizontal dependency -> no halo lanes)
// Two passes over the volume (4*S bytes) instead of three; both directions apply x, then y, then z like
// the reference (the inverse is NOT the mirrored order there either), so results are bit-identical.
#include "stream_common.cuh"

namespace dwtb200 {

using WV = W97F;
using T = float;
constexpr int VPL = 8, OUTW = 30 * VPL;

template <bool INV> __global__ void __launch_bounds__(128, 4) k_vol_xy(const VolParams p)
{
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (gw >= p.ncg * p.nstrips) return;
    const int cg = gw % p.ncg, strip = gw / p.ncg;
    const int xl = cg * OUTW - VPL + lane * VPL;
    const int W = p.nx, H = p.ny;
    const T *src = p.src + (int64_t)blockIdx.y * p.s_slice;
    T *dst = p.dst + (int64_t)blockIdx.y * p.d_slice;

    const bool fast = __all_sync(FULL, xl >= 0 && xl + VPL <= W);
    int cx[VPL];
    if (!fast) {
#pragma unroll
        for (int i = 0; i < VPL; i++) cx[i] = reflect(xl + i, W);
    }
    auto load = [&](int r, T(&v)[VPL]) {
        const T *rp = src + (int64_t)reflect(r, H) * p.s_pitch;
        if (fast) {
            ld_vec<T, VPL>(rp + xl, v);
        } else {
#pragma unroll
            for (int i = 0; i < VPL; i++) v[i] = __ldg(rp + cx[i]);
        }
    };
    const bool producer = lane >= 1 && lane <= 30 && xl < W;
    const bool whole = xl + VPL <= W;
    auto put = [&](int row, const T(&o)[VPL]) {
        if (row < 0 || row >= H || !producer) return;
        T *q = dst + (int64_t)row * p.d_pitch + xl;
        if (whole) {
            st_vec<T, VPL>(q, o);
        } else {
#pragma unroll
            for (int i = 0; i < VPL; i++)
                if (xl + i < W) q[i] = o[i];
        }
    };

    T st[WV::NS][VPL], a[VPL], b[VPL], na[VPL], nb[VPL];
    if constexpr (!INV) {
        constexpr int WARM = 3, DELAY = 1;
        const int nLy = (H + 1) >> 1;
        const int k0 = strip * p.pps, k1 = min(k0 + p.pps, nLy);
        const int m0 = k0 + DELAY - WARM, m1 = k1 - 1 + DELAY;
        load(2 * m0, st[0]);
        hfwd<WV, VPL>(st[0]);
        load(2 * m0 + 1, na);
        load(2 * m0 + 2, nb);
#pragma unroll
        for (int s = 1; s < WV::NS; s++)
#pragma unroll
            for (int i = 0; i < VPL; i++) st[s][i] = 0.f;
        for (int m = m0; m <= m1; m++) {
#pragma unroll
            for (int i = 0; i < VPL; i++) {
                a[i] = na[i];
                b[i] = nb[i];
            }
            if (m < m1) {
                load(2 * m + 3, na);
                load(2 * m + 4, nb);
            }
            hfwd<WV, VPL>(a);
            hfwd<WV, VPL>(b);
            T oL[VPL], oH[VPL];
            vfwd<WV, VPL>(a, b, st, oL, oH);
            const int kk = m - DELAY;
            if (kk >= k0) {
                put(2 * kk, oL);
                put(2 * kk + 1, oH);   // dropped by put() when H is odd and this is the last pair
            }
        }
    } else {
        constexpr int WARM = 4, DELAY = 1;
        const int q0 = strip * p.pps, q1 = min(q0 + p.pps, (H >> 1) + 1);
        const int ka = q0 + DELAY - WARM, kb = q1 - 1 + DELAY;
#pragma unroll
        for (int s = 0; s < WV::NS; s++)
#pragma unroll
            for (int i = 0; i < VPL; i++) st[s][i] = 0.f;
        load(2 * ka, na);
        load(2 * ka + 1, nb);
        for (int k = ka; k <= kb; k++) {
#pragma unroll
            for (int i = 0; i < VPL; i++) {
                a[i] = na[i];
                b[i] = nb[i];
            }
            if (k < kb) {
                load(2 * k + 2, na);
                load(2 * k + 3, nb);
            }
            hinv<WV, VPL>(a);   // x first, then y (src/volume-dwt.c:1115)
            hinv<WV, VPL>(b);
            T oO[VPL], oE[VPL];
            vinv<WV, VPL>(a, b, st, oO, oE);
            const int q = k - DELAY;
            if (q >= q0) {
                put(2 * q - 1, oO);
                put(2 * q, oE);
            }
        }
    }
}

template <bool INV> __global__ void __launch_bounds__(128, 4) k_vol_z(const VolParams p)
{
    const int lane = threadIdx.x & 31;
    const int cg = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);   // column group of 32*VPL columns
    if (cg >= p.ncg) return;
    const int y = blockIdx.y, strip = blockIdx.z;
    const int xl = cg * 32 * VPL + lane * VPL;
    const int W = p.nx, N = p.nz;
    if (xl >= W) return;
    const bool whole = xl + VPL <= W;
    const T *src = p.src + (int64_t)y * p.s_pitch + xl;
    T *dst = p.dst + (int64_t)y * p.d_pitch + xl;
    auto load = [&](int z, T(&v)[VPL]) {
        const T *rp = src + (int64_t)reflect(z, N) * p.s_slice;
        if (whole) {
            ld_vec<T, VPL>(rp, v);
        } else {
#pragma unroll
            for (int i = 0; i < VPL; i++) v[i] = (xl + i < W) ? __ldg(rp + i) : 0.f;
        }
    };
    auto put = [&](int z, const T(&o)[VPL]) {
        if (z < 0 || z >= N) return;
        T *q = dst + (int64_t)z * p.d_slice;
        if (whole) {
            st_vec<T, VPL>(q, o);
        } else {
#pragma unroll
            for (int i = 0; i < VPL; i++)
                if (xl + i < W) q[i] = o[i];
        }
    };
    T st[WV::NS][VPL], a[VPL], b[VPL], na[VPL], nb[VPL];
    if constexpr (!INV) {
        constexpr int WARM = 3, DELAY = 1;
        const int nL = (N + 1) >> 1;
        const int k0 = strip * p.pps, k1 = min(k0 + p.pps, nL);
        if (k0 >= k1) return;
        const int m0 = k0 + DELAY - WARM, m1 = k1 - 1 + DELAY;
        load(2 * m0, st[0]);
        load(2 * m0 + 1, na);
        load(2 * m0 + 2, nb);
#pragma unroll
        for (int s = 1; s < WV::NS; s++)
#pragma unroll
            for (int i = 0; i < VPL; i++) st[s][i] = 0.f;
        for (int m = m0; m <= m1; m++) {
#pragma unroll
            for (int i = 0; i < VPL; i++) {
                a[i] = na[i];
                b[i] = nb[i];
            }
            if (m < m1) {
                load(2 * m + 3, na);
                load(2 * m + 4, nb);
            }
            T oL[VPL], oH[VPL];
            vfwd<WV, VPL>(a, b, st, oL, oH);
            const int kk = m - DELAY;
            if (kk >= k0) {
                put(2 * kk, oL);
                put(2 * kk + 1, oH);
            }
        }
    } else {
        constexpr int WARM = 4, DELAY = 1;
        const int q0 = strip * p.pps, q1 = min(q0 + p.pps, (N >> 1) + 1);
        if (q0 >= q1) return;
        const int ka = q0 + DELAY - WARM, kb = q1 - 1 + DELAY;
#pragma unroll
        for (int s = 0; s < WV::NS; s++)
#pragma unroll
            for (int i = 0; i < VPL; i++) st[s][i] = 0.f;
        load(2 * ka, na);
        load(2 * ka + 1, nb);
        for (int k = ka; k <= kb; k++) {
#pragma unroll
            for (int i = 0; i < VPL; i++) {
                a[i] = na[i];
                b[i] = nb[i];
            }
            if (k < kb) {
                load(2 * k + 2, na);
                load(2 * k + 3, nb);
            }
            T oO[VPL], oE[VPL];
            vinv<WV, VPL>(a, b, st, oO, oE);
            const int q = k - DELAY;
            if (q >= q0) {
                put(2 * q - 1, oO);
                put(2 * q, oE);
            }
        }
    }
}

static int pick_pps(int units, int64_t other_warps, int sm_count)
{
    // enough warps for ~16 per SM, strips between 8 and 64 pairs (3-4 warm-up pairs are re-read per strip)
    int64_t per_col = ((int64_t)sm_count * 16) / (other_warps > 0 ? other_warps : 1);
    if (per_col < 1) per_col = 1;
    int pps = (int)((units + per_col - 1) / per_col);
    if (pps < 8) pps = 8;
    if (pps > 64) pps = 64;
    return pps;
}

void launch_vol_xy(VolParams p, int inverse, int sm_count, cudaStream_t st)
{
    p.ncg = (p.nx + OUTW - 1) / OUTW;
    const int units = inverse ? (p.ny >> 1) + 1 : (p.ny + 1) >> 1;
    p.pps = pick_pps(units, (int64_t)p.ncg * p.nz, sm_count);
    p.nstrips = (units + p.pps - 1) / p.pps;
    const dim3 grid((p.ncg * p.nstrips + 3) / 4, p.nz);
    if (inverse) k_vol_xy<true><<<grid, 128, 0, st>>>(p);
    else k_vol_xy<false><<<grid, 128, 0, st>>>(p);
}
void launch_vol_z(VolParams p, int inverse, int sm_count, cudaStream_t st)
{
    p.ncg = (p.nx + 32 * VPL - 1) / (32 * VPL);
    const int units = inverse ? (p.nz >> 1) + 1 : (p.nz + 1) >> 1;
    p.pps = pick_pps(units, (int64_t)p.ncg * p.ny, sm_count);
    p.nstrips = (units + p.pps - 1) / p.pps;
    const dim3 grid((p.ncg + 3) / 4, p.ny, p.nstrips);
    if (inverse) k_vol_z<true><<<grid, 128, 0, st>>>(p);
    else k_vol_z<false><<<grid, 128, 0, st>>>(p);
}

}  // namespace dwtb200
