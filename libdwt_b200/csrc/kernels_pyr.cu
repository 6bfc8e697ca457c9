// kernels_pyr.cu -- several consecutive L2-resident levels of the pyramid in ONE launch (sm_100a).
//
// Below ~1024^2 samples a level is pure latency: a launch plus a dependent global round trip (~5-9 us) for a few
// hundred nanoseconds of work.  Here a CTA carries a square tile through up to three levels without leaving shared
// memory: it stages the tile of the group's input band together with the lifting halo of ALL the fused levels
// (4 + 8 + 16 samples per side for three 9/7 levels), runs level after level on the shrinking window -- rows, then
// columns, each as mirrored-window evaluations of P output pairs per thread (lifting.cuh) -- writes the H subbands
// of the samples it owns to their Mallat positions after every level, and the LL band of the last level at the end.
// The halo is recomputed by the neighbouring tiles (same operations on the same inputs -> same bits), no
// inter-CTA communication.
//
// Semantics per level are those of the tile / tail kernels, i.e. of the reference drivers
// (/root/reference/src/libdwt.c:12837-12893 forward, 17098-17154 / 18178-18195 inverse): rows then columns
// (integer inverse: columns then rows), whole-sample mirror at the true borders of every level, L/H split at
// ceil/floor halves.  Used only for dense layouts and levels whose sides are all >= PYR_MIN_SIDE.
#include "tail_body.cuh"

namespace dwtb200 {

constexpr int PYR_THREADS = 512;
constexpr int PYR_P = 4;   // output pairs per window evaluation
constexpr size_t PYR_SMEM_MAX = 113 * 1024;   // two CTAs per SM

struct Rng {
    int a, b;   // [a, b)
    __host__ __device__ int n() const { return b - a; }
};

// input range of a level needed to compute the output pairs [r.a, r.b) of that level (taps 2k-HALO .. 2k+HALO+1),
// clipped to the band [0, n); mirrored taps fall inside as long as the range keeps 2*HALO+2 samples at a border
__host__ __device__ inline Rng pyr_fwd_input(Rng r, int n, int halo)
{
    Rng o;
    o.a = 2 * r.a - halo;
    o.b = 2 * r.b + halo;
    if (o.a < 0) o.a = 0;
    if (o.b > n) o.b = n;
    const int minw = 2 * halo + 2 < n ? 2 * halo + 2 : n;
    if (o.b - o.a < minw) {
        if (o.a == 0) o.b = minw;
        else o.a = o.b - minw;
    }
    return o;
}

// ---- one pass of window evaluations over a shared-memory window -----------------------------------------
// lines x pairs: line l (stride ls), pair k in [ka, kb) of a band of length n whose samples [ia, ..) sit at in + (idx - ia) * es.
// fn(l, k, L, H) consumes the outputs.
template <class WV, int P, class FN>
__device__ __forceinline__ void pyr_pass_fwd(const typename WV::T *in, int ls, int es, int nlines, int ia, int n, int ka, int kb, FN fn)
{
    using T = typename WV::T;
    constexpr int NT = 2 * P + 2 * WV::HALO;
    const int npair = kb - ka, ng = (npair + P - 1) / P;
    for (int t = threadIdx.x; t < nlines * ng; t += PYR_THREADS) {
        const int l = t % nlines, g = t / nlines;
        int k = ka + g * P;
        if (k + P > kb) k = kb - P;   // last group: shifted back (recomputes a few pairs, same values)
        const T *line = in + l * ls;
        T w[NT], L[P], H[P];
        const int t0 = 2 * k - WV::HALO;
        if (t0 >= 0 && t0 + NT <= n) {
#pragma unroll
            for (int q = 0; q < NT; q++) w[q] = line[(t0 + q - ia) * es];
        } else {
#pragma unroll
            for (int q = 0; q < NT; q++) w[q] = line[(reflect(t0 + q, n) - ia) * es];
        }
        window_fwd_p<WV, P>(w, L, H);
#pragma unroll
        for (int i = 0; i < P; i++) fn(l, k + i, L[i], H[i]);
    }
}

struct PyrParams {
    const void *in;          // forward: input band of level j0;   inverse: LL band of level j0 + F - 1
    void *out;               // forward: LL band of level j0+F-1;  inverse: output band (LL of level j0 - 1)
    void *plane;             // Mallat plane: H subbands written (forward) / read (inverse)
    int64_t in_pitch, in_frame, out_pitch, out_frame, plane_pitch, plane_frame;
    int W0, H0;              // full image
    int j0, F;               // levels j0 .. j0+F-1
    int T;                   // edge of the owned tile: final LL samples (forward), output samples (inverse)
    int pitchA, pitchB;      // shared-memory row pitches (odd: conflict-free column walks)
    int elemsA;              // elements of buffer A (buffer B follows)
};

template <class WV> __global__ void __launch_bounds__(PYR_THREADS) k_fwd_pyr(const PyrParams p)
{
    using T = typename WV::T;
    constexpr int HALO = WV::HALO;
    extern __shared__ __align__(16) unsigned char pyr_smem[];
    T *bufA = reinterpret_cast<T *>(pyr_smem), *bufB = bufA + p.elemsA;
    pdl_begin();
    const int F = p.F;
    int w[4], h[4];
    for (int f = 0; f <= F; f++) {
        w[f] = cdiv_pow2(p.W0, p.j0 + f);
        h[f] = cdiv_pow2(p.H0, p.j0 + f);
    }
    Rng rx[4], ry[4];   // r[f]: input range of level f;  r[f+1]: the pairs level f computes;  r[F]: the owned tile
    rx[F].a = blockIdx.x * p.T;
    rx[F].b = min(rx[F].a + p.T, w[F]);
    ry[F].a = blockIdx.y * p.T;
    ry[F].b = min(ry[F].a + p.T, h[F]);
    for (int f = F - 1; f >= 0; f--) {
        rx[f] = pyr_fwd_input(rx[f + 1], w[f], HALO);
        ry[f] = pyr_fwd_input(ry[f + 1], h[f], HALO);
    }
    // stage the level-j0 window
    {
        const T *src = (const T *)p.in + (int64_t)blockIdx.z * p.in_frame + (int64_t)ry[0].a * p.in_pitch + rx[0].a;
        const int ww = rx[0].n(), hh = ry[0].n();
        for (int t = threadIdx.x; t < ww * hh; t += PYR_THREADS) {
            const int yy = t / ww, xx = t % ww;
            bufA[yy * p.pitchA + xx] = __ldg(src + (int64_t)yy * p.in_pitch + xx);
        }
    }
    __syncthreads();
    T *plane = (T *)p.plane + (int64_t)blockIdx.z * p.plane_frame;
    T *out = (T *)p.out + (int64_t)blockIdx.z * p.out_frame;
    int pitch_in = p.pitchA;
    for (int f = 0; f < F; f++) {
        const int nKx = rx[f + 1].n();
        const int pB = p.pitchB;
        // ---- rows: bufA (ry[f] x rx[f]) -> bufB (ry[f] x [L pairs | H pairs]) ----
        {
            const int kax = rx[f + 1].a;
            auto put = [&](int l, int k, T L, T H) {
                bufB[l * pB + (k - kax)] = L;
                bufB[l * pB + nKx + (k - kax)] = H;
            };
            if (nKx >= PYR_P) pyr_pass_fwd<WV, PYR_P>(bufA, pitch_in, 1, ry[f].n(), rx[f].a, w[f], kax, rx[f + 1].b, put);
            else pyr_pass_fwd<WV, 1>(bufA, pitch_in, 1, ry[f].n(), rx[f].a, w[f], kax, rx[f + 1].b, put);
        }
        __syncthreads();
        // ---- columns: bufB -> next window (bufA) / LL band, H subbands of the owned samples -> Mallat plane ----
        {
            const int s = F - 1 - f;
            const int oxa = (blockIdx.x * p.T) << s, oxb = min(((blockIdx.x + 1) * p.T) << s, w[f + 1]);
            const int oya = (blockIdx.y * p.T) << s, oyb = min(((blockIdx.y + 1) * p.T) << s, h[f + 1]);
            const int nHx = w[f] >> 1, nHy = h[f] >> 1, odx = w[f + 1], ody = h[f + 1];
            const int kax = rx[f + 1].a, kay = ry[f + 1].a;
            const bool last = f == F - 1;
            const int pn = nKx | 1;   // odd pitch of the next window
            auto put = [&](int c, int ky, T lo, T hi) {
                const bool left = c < nKx;
                const int kx = kax + (left ? c : c - nKx);
                const bool own = kx >= oxa && kx < oxb && ky >= oya && ky < oyb;
                if (left) {
                    if (!last) bufA[(ky - kay) * pn + c] = lo;
                    else if (own) out[(int64_t)ky * p.out_pitch + kx] = lo;
                    if (own && ky < nHy) plane[(int64_t)(ody + ky) * p.plane_pitch + kx] = hi;
                } else if (own && kx < nHx) {
                    plane[(int64_t)ky * p.plane_pitch + odx + kx] = lo;
                    if (ky < nHy) plane[(int64_t)(ody + ky) * p.plane_pitch + odx + kx] = hi;
                }
            };
            const int nKy = ry[f + 1].n();
            if (nKy >= PYR_P) pyr_pass_fwd<WV, PYR_P>(bufB, 1, pB, 2 * nKx, ry[f].a, h[f], kay, ry[f + 1].b, put);
            else pyr_pass_fwd<WV, 1>(bufB, 1, pB, 2 * nKx, ry[f].a, h[f], kay, ry[f + 1].b, put);
        }
        __syncthreads();
        pitch_in = nKx | 1;
    }
}

// ---- host side ---------------------------------------------------------------------------------------
// worst-case window sizes of a group (interior tile)
static void pyr_fwd_extent(int T, int F, int halo, int &e0, int &e1)
{
    int n = T;
    e1 = T;
    for (int f = F - 1; f >= 0; f--) {
        if (f == 0) e1 = n;
        n = 2 * n + 2 * halo;
    }
    e0 = n;
}
static int odd(int v) { return v | 1; }

template <class K> static cudaError_t prep(K kern)
{
    cudaFuncAttributes a;
    cudaError_t e = cudaFuncGetAttributes(&a, kern);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PYR_SMEM_MAX);
    return e;
}
cudaError_t preload_pyr()
{
    cudaError_t e = cudaSuccess;
    for (int kind = 0; kind < K_COUNT; kind++)
        dispatch_kind(kind, [&](auto wv) {
            using WV = decltype(wv);
            if (e == cudaSuccess) e = prep(k_fwd_pyr<WV>);
        });
    return e;
}
int pyr_max_levels(int kind) { return kind_elem_size(kind) == 8 ? 2 : 3; }
int pyr_min_side() { return 16; }

// levels j0 .. j0+F-1 of the forward transform: `in` = their input band, `out` = where the LL band of the last one goes
void launch_fwd_pyr(int kind, const void *in, int64_t in_pitch, int64_t in_frame, void *out, int64_t out_pitch, int64_t out_frame,
                    void *plane, int64_t plane_pitch, int64_t plane_frame, int W0, int H0, int j0, int F, int frames, int T, cudaStream_t st)
{
    dispatch_kind(kind, [&](auto wv) {
        using WV = decltype(wv);
        PyrParams p;
        p.in = in;
        p.out = out;
        p.plane = plane;
        p.in_pitch = in_pitch;
        p.in_frame = in_frame;
        p.out_pitch = out_pitch;
        p.out_frame = out_frame;
        p.plane_pitch = plane_pitch;
        p.plane_frame = plane_frame;
        p.W0 = W0;
        p.H0 = H0;
        p.j0 = j0;
        p.F = F;
        int e0, e1;
        size_t smem;
        for (;; T /= 2) {   // the largest tile edge <= T whose two windows fit (about half an SM's shared memory each)
            pyr_fwd_extent(T, F, WV::HALO, e0, e1);
            p.pitchA = odd(e0);
            p.pitchB = odd(2 * e1);
            p.elemsA = (p.pitchA * e0 + 3) / 4 * 4;
            smem = ((size_t)p.elemsA + (size_t)p.pitchB * e0) * sizeof(typename WV::T);
            if (smem <= PYR_SMEM_MAX || T <= 2) break;
        }
        p.T = T;
        const int wF = cdiv_pow2(W0, j0 + F), hF = cdiv_pow2(H0, j0 + F);
        const dim3 grid((wF + T - 1) / T, (hF + T - 1) / T, frames);
        launch_pdl(k_fwd_pyr<WV>, grid, dim3(PYR_THREADS), smem, st, 0, p);
    });
}

}  // namespace dwtb200
