// kernels_tile.cu -- low-latency level kernels for the middle of the pyramid (sm_100a).
//
// The streaming kernels (kernels_stream.cu) reach HBM speed on the big levels but every warp walks
// down a strip serially, so a level can never finish in less than ~12 us; once a level's LL band is
// a few MB (L2-resident, written by the previous level) that floor dominates.  These kernels cut the
// same level into small 2-D tiles -- one CTA stages a (TH + 2*HALO) x (TW + 2*HALO) tile with its
// lifting-depth halo in shared memory, lifts rows and then columns with the mirrored-window
// evaluation of lifting.cuh (no intra-pass barriers, two __syncthreads in total), and writes each
// subband row as one 128-byte segment -- so the critical path is one tile (~1-2 us) and thousands
// of tiles run concurrently.
//
// Same reference semantics as the streaming kernels (/root/reference/src/libdwt.c:12837-12893 forward,
// 17098-17154 / 18178-18195 inverse): every sample sees "row lifting + scale, then column lifting +
// scale" (int inverse: columns first) with whole-sample mirrored borders, so the result is
// bit-identical to the reference whichever kernel family handles a level.
#include "chain.cuh"
#include "tail_body.cuh"

namespace dwtb200 {

constexpr int TILE_THREADS = 1024;

template <class WV> struct TileCfg {
    static constexpr int TW = 64, TH = 32;
    static constexpr int SW = TW + 2 * WV::HALO, SH = TH + 2 * WV::HALO;
};

// =====================================================================================================
// forward
// =====================================================================================================
template <class WV> struct TileSmem {
    using T = typename WV::T;
    using C = TileCfg<WV>;
    static constexpr int BELEMS = (C::SH * C::TW > C::TH * C::SW) ? C::SH * C::TW : C::TH * C::SW;
    T A[C::SH][C::SW];   // tile with halo (mirrored at the image borders)
    T B[BELEMS];         // after the first pass
};

// one tile (bx, by) of frame bz; LD: see tail_body.cuh
template <class WV, class LD>
__device__ __forceinline__ void fwd_tile_body(const LevelParams &p, int bx, int by, int bz, TileSmem<WV> &sm, LD ld)
{
    using T = typename WV::T;
    using C = TileCfg<WV>;
    constexpr int HALO = WV::HALO, TW = C::TW, TH = C::TH, SW = C::SW, SH = C::SH, WIN = 2 * HALO + 2;
    T(*A)[SW] = sm.A;
    T(*B)[TW] = reinterpret_cast<T(*)[TW]>(sm.B);   // after the row pass: [L half | H half] of every row

    const int tid = threadIdx.x;
    const int x0 = bx * TW, y0 = by * TH;
    const int W = p.W, H = p.H;
    const T *src = (const T *)p.src + (int64_t)bz * p.src_frame;

    const bool interior = x0 - HALO >= 0 && x0 + TW + HALO <= W && y0 - HALO >= 0 && y0 + TH + HALO <= H;
    if (interior) {
        constexpr int PER = 16 / sizeof(T), VW = SW / PER;
        if constexpr (SW % PER == 0 && HALO % PER == 0) {
            for (int i = tid; i < SH * VW; i += TILE_THREADS) {
                const int r = i / VW, v = i % VW;
                const int4 q = ld(reinterpret_cast<const int4 *>(src + (int64_t)(y0 - HALO + r) * p.src_pitch + (x0 - HALO)) + v);
                *reinterpret_cast<int4 *>(&A[r][v * PER]) = q;
            }
        } else {
            for (int i = tid; i < SH * SW; i += TILE_THREADS) {
                const int r = i / SW, c = i % SW;
                A[r][c] = ld(src + (int64_t)(y0 - HALO + r) * p.src_pitch + (x0 - HALO + c));
            }
        }
    } else {
        for (int i = tid; i < SH * SW; i += TILE_THREADS) {
            const int r = i / SW, c = i % SW;
            A[r][c] = ld(src + (int64_t)reflect(y0 - HALO + r, H) * p.src_pitch + reflect(x0 - HALO + c, W));
        }
    }
    __syncthreads();

    // rows: pair k of row r <- window A[r][2k .. 2k+WIN)
    for (int i = tid; i < SH * (TW / 2); i += TILE_THREADS) {
        const int r = i / (TW / 2), k = i % (TW / 2);
        T w[WIN];
#pragma unroll
        for (int q = 0; q < WIN; q++) w[q] = A[r][2 * k + q];
        T L, Hh;
        window_fwd<WV>(w, L, Hh);
        B[r][k] = L;
        B[r][TW / 2 + k] = Hh;
    }
    __syncthreads();

    // columns: pair kk of column x <- window B[2kk .. 2kk+WIN)[x]; x < TW/2 is the L half (-> LL, LH)
    T *ll = (T *)p.ll + (int64_t)bz * p.ll_frame;
    T *hl = (T *)p.hl + (int64_t)bz * p.sub_frame;
    T *lh = (T *)p.lh + (int64_t)bz * p.sub_frame;
    T *hh = (T *)p.hh + (int64_t)bz * p.sub_frame;
    for (int i = tid; i < (TH / 2) * TW; i += TILE_THREADS) {
        const int kk = i / TW, x = i % TW;
        const bool low = x < TW / 2;
        const int gx = x0 / 2 + (low ? x : x - TW / 2), gy = y0 / 2 + kk;
        if (gx >= (low ? p.nLx : p.nHx) || gy >= p.nLy) continue;
        T w[WIN];
#pragma unroll
        for (int q = 0; q < WIN; q++) w[q] = B[2 * kk + q][x];
        T L, Hh;
        window_fwd<WV>(w, L, Hh);
        if (low) {
            ll[(int64_t)gy * p.ll_pitch + gx] = L;
            if (gy < p.nHy) lh[(int64_t)gy * p.sub_pitch + gx] = Hh;
        } else {
            hl[(int64_t)gy * p.sub_pitch + gx] = L;
            if (gy < p.nHy) hh[(int64_t)gy * p.sub_pitch + gx] = Hh;
        }
    }
}

// =====================================================================================================
// inverse
// =====================================================================================================
template <class WV, class LD>
__device__ __forceinline__ void inv_tile_body(const LevelParams &p, int bx, int by, int bz, TileSmem<WV> &sm, LD ld)
{
    using T = typename WV::T;
    using C = TileCfg<WV>;
    constexpr int HALO = WV::HALO, TW = C::TW, TH = C::TH, SW = C::SW, SH = C::SH, WIN = 2 * HALO + 2;
    T(*A)[SW] = sm.A;   // interleaved coefficients with halo (mirrored at the borders)
    T *Bs = sm.B;

    const int tid = threadIdx.x;
    const int x0 = bx * TW, y0 = by * TH;
    const int W = p.W, H = p.H;
    const T *ll = (const T *)p.ll + (int64_t)bz * p.ll_frame;
    const T *hl = (const T *)p.hl + (int64_t)bz * p.sub_frame;
    const T *lh = (const T *)p.lh + (int64_t)bz * p.sub_frame;
    const T *hh = (const T *)p.hh + (int64_t)bz * p.sub_frame;
    T *dst = (T *)p.dst + (int64_t)bz * p.dst_frame;

    // stage: local (r, c) <-> interleaved coefficient (y0-HALO+r, x0-HALO+c), mirrored; parity picks the
    // subband.  Consecutive threads take consecutive columns of ONE subband (coalesced).
    for (int i = tid; i < SH * SW; i += TILE_THREADS) {
        const int r = i / SW, cc = i % SW;
        const int half = cc / (SW / 2), c = 2 * (cc % (SW / 2)) + half;   // first the even columns, then the odd ones
        const int gy = reflect(y0 - HALO + r, H), gx = reflect(x0 - HALO + c, W);
        const T *base;
        int64_t pitch;
        if (gy & 1) {
            base = (gx & 1) ? hh : lh;
            pitch = p.sub_pitch;
        } else {
            base = (gx & 1) ? hl : ll;
            pitch = (gx & 1) ? p.sub_pitch : p.ll_pitch;
        }
        A[r][c] = ld(base + (int64_t)(gy >> 1) * pitch + (gx >> 1));
    }
    __syncthreads();

    if constexpr (!WV::INV_COLS_FIRST) {
        // rows first (float, double: libdwt.c:17098 then 17127): B[SH][TW]
        T(*B)[TW] = reinterpret_cast<T(*)[TW]>(Bs);
        for (int i = tid; i < SH * (TW / 2); i += TILE_THREADS) {
            const int r = i / (TW / 2), k = i % (TW / 2);
            T w[WIN];
#pragma unroll
            for (int q = 0; q < WIN; q++) w[q] = A[r][2 * k + q];
            T E, O;
            window_inv<WV>(w, E, O);
            B[r][2 * k] = E;
            B[r][2 * k + 1] = O;
        }
        __syncthreads();
        for (int i = tid; i < (TH / 2) * TW; i += TILE_THREADS) {
            const int kk = i / TW, x = i % TW;
            const int gx = x0 + x, gy = y0 + 2 * kk;
            if (gx >= W || gy >= H) continue;
            T w[WIN];
#pragma unroll
            for (int q = 0; q < WIN; q++) w[q] = B[2 * kk + q][x];
            T E, O;
            window_inv<WV>(w, E, O);
            dst[(int64_t)gy * p.dst_pitch + gx] = E;
            if (gy + 1 < H) dst[(int64_t)(gy + 1) * p.dst_pitch + gx] = O;
        }
    } else {
        // columns first (int 5/3: libdwt.c:18178 then 18187): B[TH][SW]
        T(*B)[SW] = reinterpret_cast<T(*)[SW]>(Bs);
        for (int i = tid; i < (TH / 2) * SW; i += TILE_THREADS) {
            const int kk = i / SW, c = i % SW;
            T w[WIN];
#pragma unroll
            for (int q = 0; q < WIN; q++) w[q] = A[2 * kk + q][c];
            T E, O;
            window_inv<WV>(w, E, O);
            B[2 * kk][c] = E;
            B[2 * kk + 1][c] = O;
        }
        __syncthreads();
        for (int i = tid; i < TH * (TW / 2); i += TILE_THREADS) {
            const int r = i / (TW / 2), k = i % (TW / 2);
            const int gx = x0 + 2 * k, gy = y0 + r;
            if (gx >= W || gy >= H) continue;
            T w[WIN];
#pragma unroll
            for (int q = 0; q < WIN; q++) w[q] = B[r][2 * k + q];
            T E, O;
            window_inv<WV>(w, E, O);
            T *q = dst + (int64_t)gy * p.dst_pitch + gx;
            q[0] = E;
            if (gx + 1 < W) q[1] = O;
        }
    }
}

// ---- stand-alone kernels: one tile per CTA ---------------------------------------------------------
// Chained (struct Chain): the tile waits for the row blocks of its input the producing kernel has not
// finished yet, reads them through L2 (they were written while this kernel was already running), and bumps
// its own row block's counter when its LL rows (forward) / output rows (inverse) are written.
template <class WV> __device__ __forceinline__ void tile_wait(const LevelParams &p, uint32_t gen, int row_lo, int row_hi)
{
    const int b0 = chain_block(p.chain, row_lo), b1 = chain_block(p.chain, row_hi);
    for (int b = b0 + (int)threadIdx.x; b <= b1; b += TILE_THREADS) chain_wait(p.chain, gen, blockIdx.z, b);
    __syncthreads();
}
template <class WV> __global__ void __launch_bounds__(TILE_THREADS) k_fwd_tile(const LevelParams p)
{
    using C = TileCfg<WV>;
    __shared__ __align__(16) TileSmem<WV> sm;
    const uint32_t gen = chain_begin(p.chain);
    if (p.chain.in) {
        const int y0 = blockIdx.y * C::TH;
        tile_wait<WV>(p, gen, max(y0 - WV::HALO, 0), min(y0 + C::TH + WV::HALO, p.H) - 1);
        fwd_tile_body<WV>(p, blockIdx.x, blockIdx.y, blockIdx.z, sm, LdCg());
    } else {
        fwd_tile_body<WV>(p, blockIdx.x, blockIdx.y, blockIdx.z, sm, LdNc());
    }
    if (p.chain.gen) {
        __syncthreads();
        if (threadIdx.x == 0) chain_signal(p.chain, blockIdx.z, blockIdx.y);
    }
}
template <class WV> __global__ void __launch_bounds__(TILE_THREADS) k_inv_tile(const LevelParams p)
{
    using C = TileCfg<WV>;
    __shared__ __align__(16) TileSmem<WV> sm;
    const uint32_t gen = chain_begin(p.chain);
    if (p.chain.in) {   // the LL band comes from the previous kernel of the chain, the other subbands from the source plane
        const int y0 = blockIdx.y * C::TH;
        tile_wait<WV>(p, gen, max(y0 - WV::HALO, 0) >> 1, (min(y0 + C::TH + WV::HALO, p.H) - 1) >> 1);
        inv_tile_body<WV>(p, blockIdx.x, blockIdx.y, blockIdx.z, sm, LdCg());
    } else {
        inv_tile_body<WV>(p, blockIdx.x, blockIdx.y, blockIdx.z, sm, LdNc());
    }
    if (p.chain.gen) {
        __syncthreads();
        if (threadIdx.x == 0) chain_signal(p.chain, blockIdx.z, blockIdx.y);
    }
}

// ---- launchers -------------------------------------------------------------------------------
template <class WV> static dim3 tile_grid(const LevelParams &p, int frames)
{
    using C = TileCfg<WV>;
    return dim3((p.W + C::TW - 1) / C::TW, (p.H + C::TH - 1) / C::TH, frames);
}
int tile_rows() { return TileCfg<W97F>::TH; }
dim3 tile_grid_of(int kind, const LevelParams &p, int frames)
{
    dim3 g;
    dispatch_kind(kind, [&](auto wv) { g = tile_grid<decltype(wv)>(p, frames); });
    return g;
}
void launch_fwd_tile(int kind, const LevelParams &p, int frames, cudaStream_t st)
{
    dispatch_kind(kind, [&](auto wv) {
        using WV = decltype(wv);
        launch_pdl(k_fwd_tile<WV>, tile_grid<WV>(p, frames), dim3(TILE_THREADS), 0, st, p.chain.pdl, p);
    });
}
void launch_inv_tile(int kind, const LevelParams &p, int frames, cudaStream_t st)
{
    dispatch_kind(kind, [&](auto wv) {
        using WV = decltype(wv);
        launch_pdl(k_inv_tile<WV>, tile_grid<WV>(p, frames), dim3(TILE_THREADS), 0, st, p.chain.pdl, p);
    });
}

template <class K> static cudaError_t touch(K kern)
{
    cudaFuncAttributes a;
    return cudaFuncGetAttributes(&a, kern);
}
cudaError_t preload_tile()
{
    cudaError_t e = cudaSuccess;
    for (int kind = 0; kind < K_COUNT; kind++)
        dispatch_kind(kind, [&](auto wv) {
            using WV = decltype(wv);
            if (e == cudaSuccess) e = touch(k_fwd_tile<WV>);
            if (e == cudaSuccess) e = touch(k_inv_tile<WV>);
        });
    return e;
}

}  // namespace dwtb200
