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
//   k_vol3t    volumes of at least 64 x 32 x 16: all three axes in ONE pass (tiles marching along z, staged by tensor copies;
//              k_vol3 is the same pass staged by cp.async, kept as DWTB200_TUNE_VOL3 = 2 and as the fallback without a tensor map)
// Two passes over the volume (4*S bytes) instead of three; both directions apply x, then y, then z like
// the reference (the inverse is NOT the mirrored order there either), so results are bit-identical.
#include <cuda.h>
#include <cstdlib>

#include "ring_common.cuh"
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

// =====================================================================================================
// forward, single pass: x, y and z lifting of a (V3_TX x V3_TY) column of the volume marching along z
// =====================================================================================================
// Two passes move every voxel four times through HBM; this kernel reads and writes it once.  A CTA owns a tile of
// V3_TX x V3_TY (x, y) positions and a range of slice pairs.  Per slice: the tile plus a 4-sample halo is staged in
// shared memory with cp.async (three buffers: two slices in flight), lifted along x by window evaluation (16 outputs
// from 24 staged samples per task), then along y (8 outputs from 16 x-lifted samples) -- each thread ends up holding
// the x/y-lifted values of ITS 8 positions (one column, 8 rows) -- and those go straight into the z register pipeline
// of k_vol_z (vfwd), whose state (4 values per position) stays in the thread's registers from slice to slice.
// Same operations in the same order per sample as the two-pass kernels and the reference (x, then y, then z).
constexpr int V3_TX = 64, V3_TY = 32, V3_THREADS = 256, V3_HALO = 4, V3_NCTA = 2;
constexpr int V3_SW = V3_TX + 2 * V3_HALO, V3_SH = V3_TY + 2 * V3_HALO;   // staged tile: 136 x 40
constexpr int V3_SP = V3_SW + 4;      // staged row pitch (140 words: conflict-free 16-byte accesses of consecutive rows)
constexpr int V3_XP = V3_TX + 12;     // pitch of the x-lifted buffer (140)
constexpr int V3_NBUF = 3;
constexpr int V3_SMEM = (V3_NBUF * V3_SH * V3_SP + V3_SH * V3_XP) * (int)sizeof(float);

__device__ __forceinline__ void cp_async16(float *smem, const float *gmem)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async4(float *smem, const float *gmem)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}

template <bool INV> __global__ void __launch_bounds__(V3_THREADS, V3_NCTA) k_vol3(const VolParams p)
{
    extern __shared__ __align__(16) float v3_smem[];
    float *stage = v3_smem, *xb = v3_smem + V3_NBUF * V3_SH * V3_SP;
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * V3_TX, y0 = blockIdx.y * V3_TY;
    const int nx = p.nx, ny = p.ny, N = p.nz;

    // the slice schedule of k_vol_z: forward pairs (2m+1, 2m+2) after a seed slice 2 m0, inverse pairs (2k, 2k+1)
    constexpr int WARM = INV ? 4 : 3, DELAY = 1;
    const int units = INV ? (N >> 1) + 1 : (N + 1) >> 1;
    const int k0 = (blockIdx.z + p.strip0) * p.pps, k1 = min(k0 + p.pps, units);
    if (k0 >= k1) return;
    const int m0 = k0 + DELAY - WARM, m1 = k1 - 1 + DELAY;
    const int zfirst = 2 * m0, npairs = m1 - m0 + 1, nsl = 2 * npairs + (INV ? 0 : 1);   // slices zfirst .. zfirst + nsl - 1 (mirrored into the volume)

    // staging: groups of four columns; a thread copies the same (up to three) groups of every slice, so their offsets are
    // computed once.  goff < 0: the group straddles the volume's edge and is copied element by element through the mirror.
    constexpr int NG = (V3_SH * (V3_SW / 4) + V3_THREADS - 1) / V3_THREADS;
    int soff[NG];
    int64_t goff[NG];
#pragma unroll
    for (int k = 0; k < NG; k++) {
        const int g = tid + k * V3_THREADS;
        const int r = g / (V3_SW / 4), c = (g - r * (V3_SW / 4)) * 4;
        const int gx = x0 - V3_HALO + c;
        const int64_t row = (int64_t)reflect(y0 - V3_HALO + r, ny) * p.s_pitch;
        soff[k] = g < V3_SH * (V3_SW / 4) ? r * V3_SP + c : -1;
        goff[k] = (gx >= 0 && gx + 3 < nx) ? row + gx : -(row + 1);   // slow path keeps the row offset, biased to stay negative
    }
    auto issue = [&](int i) {
        if (i < nsl) {
            const float *sl = p.src + (int64_t)reflect(zfirst + i, N) * p.s_slice;
            float *buf = stage + (i % V3_NBUF) * (V3_SH * V3_SP);
#pragma unroll
            for (int k = 0; k < NG; k++) {
                if (soff[k] < 0) continue;
                if (goff[k] >= 0) {
                    cp_async16(buf + soff[k], sl + goff[k]);
                } else {
                    const float *row = sl + (-goff[k] - 1);
                    const int c = soff[k] % V3_SP, gx = x0 - V3_HALO + c;
#pragma unroll
                    for (int j = 0; j < 4; j++) cp_async4(buf + soff[k] + j, row + reflect(gx + j, nx));
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    // this thread's 8 positions: column x0 + px, rows y0 + 8 * ps .. + 7
    const int px = tid % V3_TX, ps = tid / V3_TX;
    T st[WV::NS][8];
#pragma unroll
    for (int s = 0; s < WV::NS; s++)
#pragma unroll
        for (int i = 0; i < 8; i++) st[s][i] = 0.f;
    const int rows_ok = (x0 + px < nx) ? min(max(ny - (y0 + 8 * ps), 0), 8) : 0;   // how many of the 8 rows exist
    float *out = p.dst + (int64_t)(y0 + 8 * ps) * p.d_pitch + x0 + px;
    const int dp = (int)p.d_pitch;
    auto put = [&](int z, const T(&o)[8]) {
        if (z < 0 || z >= N) return;
        float *q = out + (int64_t)z * p.d_slice;
        if (rows_ok == 8) {
#pragma unroll
            for (int j = 0; j < 8; j++) q[j * dp] = o[j];
        } else {
#pragma unroll
            for (int j = 0; j < 8; j++)
                if (j < rows_ok) q[j * dp] = o[j];
        }
    };
    // x and y lifting of slice number i of the sequence; leaves this thread's 8 values in v
    const int xr = tid % V3_SH, xsg = tid / V3_SH;   // x task of this thread (threads >= V3_SH * 8 have none)
    auto xy = [&](int i, T(&v)[8]) {
        issue(i + 2);
        asm volatile("cp.async.wait_group 2;" ::: "memory");
        __syncthreads();   // slice i is staged (and every thread is done with the x-lifted buffer of slice i - 1)
        const float *buf = stage + (i % V3_NBUF) * (V3_SH * V3_SP);
        if (xsg < V3_TX / 16) {   // x: (staged row, run of 16 output columns); consecutive threads take consecutive rows
            T w[24], L[8], H[8];
            const float4 *src4 = reinterpret_cast<const float4 *>(buf + xr * V3_SP + 16 * xsg);
#pragma unroll
            for (int j = 0; j < 6; j++) {
                const float4 q4 = src4[j];
                w[4 * j] = q4.x; w[4 * j + 1] = q4.y; w[4 * j + 2] = q4.z; w[4 * j + 3] = q4.w;
            }
            if constexpr (INV) window_inv_p<WV, 8>(w, L, H);
            else window_fwd_p<WV, 8>(w, L, H);
            float4 *dst4 = reinterpret_cast<float4 *>(xb + xr * V3_XP + 16 * xsg);
#pragma unroll
            for (int j = 0; j < 4; j++) dst4[j] = make_float4(L[2 * j], H[2 * j], L[2 * j + 1], H[2 * j + 1]);
        }
        __syncthreads();
        T w[16], L[4], H[4];   // y: 8 output rows of this thread's column from 16 x-lifted rows
        const float *col = xb + (8 * ps) * V3_XP + px;
#pragma unroll
        for (int j = 0; j < 16; j++) w[j] = col[j * V3_XP];
        if constexpr (INV) window_inv_p<WV, 4>(w, L, H);
        else window_fwd_p<WV, 4>(w, L, H);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            v[2 * j] = L[j];
            v[2 * j + 1] = H[j];
        }
    };

    issue(0);
    issue(1);
    if constexpr (!INV) {
        xy(0, st[0]);   // z: the register pipeline of k_vol_z -- slice 0 of the sequence seeds st[0], then pairs (2m+1, 2m+2)
        for (int q = 0; q < npairs; q++) {
            T a[8], b[8], oL[8], oH[8];
            xy(2 * q + 1, a);
            xy(2 * q + 2, b);
            vfwd<WV, 8>(a, b, st, oL, oH);
            const int kk = m0 + q - DELAY;
            if (kk >= k0) {
                put(2 * kk, oL);
                put(2 * kk + 1, oH);
            }
        }
    } else {
        for (int q = 0; q < npairs; q++) {
            T a[8], b[8], oO[8], oE[8];
            xy(2 * q, a);
            xy(2 * q + 1, b);
            vinv<WV, 8>(a, b, st, oO, oE);
            const int kk = m0 + q - DELAY;
            if (kk >= k0) {
                put(2 * kk - 1, oO);
                put(2 * kk, oE);
            }
        }
    }
}


// =====================================================================================================
// the same pass with the tile staged by the tensor-copy engine (round 2)
// =====================================================================================================
// k_vol3 spends about a fifth of its instructions on staging (720 cp.async of 16 bytes per slice and CTA, their addresses and
// mirror tests).  Here ONE thread issues ONE cp.async.bulk.tensor per slice: a 76 x 40 x 1 box of a 3-D tensor map over the
// source volume (coordinates may leave the volume: the engine fills zeros there), completion counted on an mbarrier.  CTAs
// whose staged window leaves the volume replace the zeros by the mirrored samples (<= 4 shared-memory moves per thread and
// slice, listed once).  The x pass of slice i and the y pass of slice i - 1 share a phase (the x-lifted buffer is double
// buffered), so a slice costs one __syncthreads instead of two and every warp has work in every phase.  NBUF staging
// buffers: NBUF - 1 slices in flight.  1024^3: 2.53 -> 1.70 ms forward, 2.62 -> 1.72 ms inverse (60 -> 41 instructions per
// voxel; profiles/ncu_vol3t_r2.txt).
constexpr int V3T_BW = V3_SW + 4;                         // box width: 76 columns, so that the staged pitch keeps 16-byte row accesses conflict-free
constexpr int V3T_STAGE = V3_SH * V3T_BW;                 // floats per staged slice (12160 bytes, a multiple of 128)
constexpr int V3T_XB = V3_SH * V3_XP;                     // floats per x-lifted buffer
constexpr int v3t_smem(int nbuf) { return (nbuf * V3T_STAGE + 2 * V3T_XB) * (int)sizeof(float) + 64; }
constexpr int V3T_NFIX = 4;                               // (4 + 4) x 40 + (4 + 4) x 72 = 896 halo positions for 256 threads

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, int c0, int c1, int c2, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
                 : "memory");
}

template <bool INV, int NCTA, int NBUF> __global__ void __launch_bounds__(V3_THREADS, NCTA) k_vol3t(const VolParams p, const __grid_constant__ CUtensorMap tmap)
{
    extern __shared__ __align__(128) float v3_smem[];
    float *stage = v3_smem, *xb = v3_smem + NBUF * V3T_STAGE;
    const uint32_t bars = smem_u32(xb + 2 * V3T_XB);
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * V3_TX, y0 = blockIdx.y * V3_TY;
    const int nx = p.nx, ny = p.ny, N = p.nz;

    constexpr int WARM = INV ? 4 : 3, DELAY = 1;
    const int units = INV ? (N >> 1) + 1 : (N + 1) >> 1;
    const int k0 = (blockIdx.z + p.strip0) * p.pps, k1 = min(k0 + p.pps, units);
    if (k0 >= k1) return;
    const int m0 = k0 + DELAY - WARM, m1 = k1 - 1 + DELAY;
    const int zfirst = 2 * m0, npairs = m1 - m0 + 1, nsl = 2 * npairs + (INV ? 0 : 1);

    if (tid == 0) {
        for (int b = 0; b < NBUF; b++) mbar_init(bars + 8 * b, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // halo positions of the staged window that lie outside the volume, and the staged position of their mirror image
    const bool edge = x0 < V3_HALO || x0 + V3_TX + V3_HALO > nx || y0 < V3_HALO || y0 + V3_TY + V3_HALO > ny;
    int fix[V3T_NFIX];
#pragma unroll
    for (int k = 0; k < V3T_NFIX; k++) {
        fix[k] = -1;
        if (!edge) continue;
        int id = tid + k * V3_THREADS, r, c;
        if (id < 8 * V3_SH) {   // four columns left of the volume, four right of it
            r = id % V3_SH;
            const int s = id / V3_SH;
            c = s < 4 ? s : nx - (x0 - V3_HALO) + (s - 4);
        } else {
            id -= 8 * V3_SH;
            if (id >= 8 * V3_SW) continue;
            c = id % V3_SW;
            const int s = id / V3_SW;
            r = s < 4 ? s : ny - (y0 - V3_HALO) + (s - 4);
        }
        if (r < 0 || r >= V3_SH || c < 0 || c >= V3_SW) continue;
        const int gx = x0 - V3_HALO + c, gy = y0 - V3_HALO + r;
        if (gx >= 0 && gx < nx && gy >= 0 && gy < ny) continue;
        const int sc = reflect(gx, nx) - (x0 - V3_HALO), sr = reflect(gy, ny) - (y0 - V3_HALO);
        if (sc < 0 || sc >= V3_SW || sr < 0 || sr >= V3_SH) continue;
        fix[k] = (r * V3T_BW + c) << 16 | (sr * V3T_BW + sc);
    }
    __syncthreads();
    auto issue = [&](int i) {
        if (tid == 0 && i < nsl) {
            const int b = i % NBUF;
            mbar_expect_tx(bars + 8 * b, V3T_STAGE * (uint32_t)sizeof(float));
            tma_load_3d(smem_u32(stage + b * V3T_STAGE), &tmap, x0 - V3_HALO, y0 - V3_HALO, reflect(zfirst + i, N), bars + 8 * b);
        }
    };

    const int px = tid % V3_TX, ps = tid / V3_TX;
    T st[WV::NS][8];
#pragma unroll
    for (int s = 0; s < WV::NS; s++)
#pragma unroll
        for (int i = 0; i < 8; i++) st[s][i] = 0.f;
    const int rows_ok = (x0 + px < nx) ? min(max(ny - (y0 + 8 * ps), 0), 8) : 0;
    float *out = p.dst + (int64_t)(y0 + 8 * ps) * p.d_pitch + x0 + px;
    const int dp = (int)p.d_pitch;
    auto put = [&](int z, const T(&o)[8]) {
        if (z < 0 || z >= N) return;
        float *q = out + (int64_t)z * p.d_slice;
        if (rows_ok == 8) {
#pragma unroll
            for (int j = 0; j < 8; j++) q[j * dp] = o[j];
        } else {
#pragma unroll
            for (int j = 0; j < 8; j++)
                if (j < rows_ok) q[j * dp] = o[j];
        }
    };
    const int xr = tid % V3_SH, xsg = tid / V3_SH;
    // phase i: x lifting of slice i (if there is one) into xb[i & 1], y lifting of slice i - 1 from xb[(i - 1) & 1] into v
    auto phase = [&](int i, T(&v)[8], bool have_y) {
        issue(i + NBUF - 1);   // into the buffer slice i - 1 was read from (every thread is past the barrier of phase i - 1)
        if (i < nsl) {
            float *buf = stage + (i % NBUF) * V3T_STAGE;
            mbar_wait(bars + 8 * (i % NBUF), (uint32_t)(i / NBUF) & 1u);
            if (edge) {
#pragma unroll
                for (int k = 0; k < V3T_NFIX; k++)
                    if (fix[k] >= 0) buf[fix[k] >> 16] = buf[fix[k] & 0xffff];
                __syncthreads();
            }
            if (xsg < V3_TX / 16) {
                T w[24], L[8], H[8];
                const float4 *src4 = reinterpret_cast<const float4 *>(buf + xr * V3T_BW + 16 * xsg);
#pragma unroll
                for (int j = 0; j < 6; j++) {
                    const float4 q4 = src4[j];
                    w[4 * j] = q4.x; w[4 * j + 1] = q4.y; w[4 * j + 2] = q4.z; w[4 * j + 3] = q4.w;
                }
                if constexpr (INV) window_inv_p<WV, 8>(w, L, H);
                else window_fwd_p<WV, 8>(w, L, H);
                float4 *dst4 = reinterpret_cast<float4 *>(xb + (i & 1) * V3T_XB + xr * V3_XP + 16 * xsg);
#pragma unroll
                for (int j = 0; j < 4; j++) dst4[j] = make_float4(L[2 * j], H[2 * j], L[2 * j + 1], H[2 * j + 1]);
            }
        }
        if (have_y) {
            T w[16], L[4], H[4];
            const float *col = xb + ((i - 1) & 1) * V3T_XB + (8 * ps) * V3_XP + px;
#pragma unroll
            for (int j = 0; j < 16; j++) w[j] = col[j * V3_XP];
            if constexpr (INV) window_inv_p<WV, 4>(w, L, H);
            else window_fwd_p<WV, 4>(w, L, H);
#pragma unroll
            for (int j = 0; j < 4; j++) {
                v[2 * j] = L[j];
                v[2 * j + 1] = H[j];
            }
        }
        if (edge) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the mirror writes, before the engine refills the buffer
        __syncthreads();
    };

#pragma unroll
    for (int i = 0; i < NBUF - 1; i++) issue(i);
    {
        T none[8];
        phase(0, none, false);
    }
    if constexpr (!INV) {
        phase(1, st[0], true);   // slice 0 of the sequence seeds the z pipeline, then pairs (2q+1, 2q+2)
#pragma unroll 2
        for (int q = 0; q < npairs; q++) {
            T a[8], b[8], oL[8], oH[8];
            phase(2 * q + 2, a, true);
            phase(2 * q + 3, b, true);
            vfwd<WV, 8>(a, b, st, oL, oH);
            const int kk = m0 + q - DELAY;
            if (kk >= k0) {
                put(2 * kk, oL);
                put(2 * kk + 1, oH);
            }
        }
    } else {
#pragma unroll 2
        for (int q = 0; q < npairs; q++) {
            T a[8], b[8], oO[8], oE[8];
            phase(2 * q + 1, a, true);
            phase(2 * q + 2, b, true);
            vinv<WV, 8>(a, b, st, oO, oE);
            const int kk = m0 + q - DELAY;
            if (kk >= k0) {
                put(2 * kk - 1, oO);
                put(2 * kk, oE);
            }
        }
    }
}

// tensor map over the source volume (driver entry point fetched through the runtime: no link against libcuda)
static bool vol3_tensor_map(const VolParams &p, CUtensorMap *map)
{
    typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static encode_fn encode = nullptr;
    static bool looked = false;
    if (!looked) {
        looked = true;
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess)
            encode = reinterpret_cast<encode_fn>(fn);
        else
            cudaGetLastError();
    }
    if (!encode) return false;
    if ((reinterpret_cast<uintptr_t>(p.src) & 15) || (p.s_pitch & 3) || (p.s_slice & 3)) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)p.nx, (cuuint64_t)p.ny, (cuuint64_t)p.nz};
    const cuuint64_t strides[2] = {(cuuint64_t)p.s_pitch * sizeof(float), (cuuint64_t)p.s_slice * sizeof(float)};
    const cuuint32_t box[3] = {(cuuint32_t)V3T_BW, (cuuint32_t)V3_SH, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    // L2 promotion none / 64 / 128 / 256 bytes: no measurable difference (1024^3 forward 1.819 - 1.822 ms)
    return encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(p.src), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool vol3_applies(const VolParams &p) { return p.nx >= V3_TX && p.ny >= V3_TY && p.nz >= 16; }   // at least one full tile
static void vol3_prepare()
{
    static bool prepared = false;
    if (!prepared) {
        cudaFuncSetAttribute(k_vol3<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, V3_SMEM);
        cudaFuncSetAttribute(k_vol3<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, V3_SMEM);
        prepared = true;
    }
}
static void vol3_launch(const VolParams &p, int inverse, int variant, int count, cudaStream_t st)
{
    const int tx = (p.nx + V3_TX - 1) / V3_TX, ty = (p.ny + V3_TY - 1) / V3_TY;
    const dim3 grid(tx, ty, count);
    CUtensorMap map;
    if (variant != 2 && vol3_tensor_map(p, &map)) {   // variant 2 (DWTB200_TUNE_VOL3 = 2): the cp.async staging of round 1
        // forward: 2 CTAs per SM (108 registers; at 80 it spills and 1024^3 takes 1.99 instead of 1.82 ms), 4 staging buffers;
        // inverse: 3 CTAs per SM (80 registers, no spills: 1.91 -> 1.83 ms), 3 buffers.  The slice-pair loop is unrolled twice so that
        // the z state is not moved back into place every iteration (forward 1.82 -> 1.77 ms; profiles/ncu_vol3t_r2.txt)
        static bool prepared_t = false;
        if (!prepared_t) {
            cudaFuncSetAttribute(k_vol3t<false, 2, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, v3t_smem(4));
            cudaFuncSetAttribute(k_vol3t<true, 3, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, v3t_smem(3));
            cudaFuncSetAttribute(k_vol3t<true, 3, 3>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            prepared_t = true;
        }
        if (inverse) k_vol3t<true, 3, 3><<<grid, V3_THREADS, v3t_smem(3), st>>>(p, map);
        else k_vol3t<false, 2, 4><<<grid, V3_THREADS, v3t_smem(4), st>>>(p, map);
        return;
    }
    if (inverse) k_vol3<true><<<grid, V3_THREADS, V3_SMEM, st>>>(p);
    else k_vol3<false><<<grid, V3_THREADS, V3_SMEM, st>>>(p);
}

void launch_vol3(VolParams p, int inverse, int variant, int sm_count, cudaStream_t st)
{
    vol3_prepare();
    const int tx = (p.nx + V3_TX - 1) / V3_TX, ty = (p.ny + V3_TY - 1) / V3_TY;
    const int units = inverse ? (p.nz >> 1) + 1 : (p.nz + 1) >> 1;
    const int ncta = variant != 2 && inverse ? 3 : V3_NCTA;   // CTAs per SM (see below)
    const double warm = inverse ? 4.0 : 3.0;
    // z ranges.  Every range recomputes `warm` slice pairs, yet many short ranges beat few long ones: the CTAs of a wave start together
    // and stay in phase (x / y / z lifting, staging waits), later waves drift apart and fill each other's gaps.  Fitted on 512^3, 768^3
    // and 1024^3 (profiles/ncu_vol3t_r2.txt, "z ranges"): time ~ (1 + warm / pps) (1 + B / waves), B = 1.44 with three CTAs per SM
    // (inverse), 0.7 with two, and half of the idle share of the last wave; 1024^3 inverse: 7 ranges of 74 pairs 1.81 ms, 13 of 40 1.70 ms
    int best = 1;
    double best_cost = 1e30;
    for (int zs = 1; zs <= 32; zs++) {
        const int pps = (units + zs - 1) / zs;
        if (zs > 1 && pps < 26) break;   // shorter ranges cost more than the model says (512^3: 16 ranges of 16 pairs 0.273 ms, 9 of 29 0.254 ms)
        const int64_t n = (int64_t)tx * ty * ((units + pps - 1) / pps);
        const double slots = (double)sm_count * ncta, waves = (double)n / slots, whole = (double)((n + (int64_t)slots - 1) / (int64_t)slots);
        const double cost = (1.0 + warm / pps) * (1.0 + (ncta == 3 ? 1.44 : 0.7) / waves) * (1.0 + 0.5 * (whole / waves - 1.0));
        if (cost < best_cost - 1e-9) {
            best_cost = cost;
            best = zs;
        }
    }
#ifdef DWTB200_DEBUG_KEYS   // measurement only: number of z ranges forced from the environment
    if (const char *e = getenv("DWTB200_VOL3_ZS"))
        if (atoi(e) > 0) best = atoi(e);
#endif
    p.pps = (units + best - 1) / best;
    p.nstrips = (units + p.pps - 1) / p.pps;
    p.strip0 = 0;
    vol3_launch(p, inverse, variant, p.nstrips, st);
}
void launch_vol3_ranges(VolParams p, int inverse, int variant, int pps, int strip0, int count, cudaStream_t st)
{
    vol3_prepare();
    const int units = inverse ? (p.nz >> 1) + 1 : (p.nz + 1) >> 1;
    p.pps = pps;
    p.nstrips = (units + pps - 1) / pps;
    p.strip0 = strip0;
    if (strip0 + count > p.nstrips) count = p.nstrips - strip0;
    if (count > 0) vol3_launch(p, inverse, variant, count, st);
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
