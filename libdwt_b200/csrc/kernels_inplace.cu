// kernels_inplace.cu -- the interleaved in-place family (sm_100a): dwt_cdf97_2f_inplace_s and its _sep / _sdl twins,
// dwt_cdf97_2i_inplace_s, dwt_cdf53_2f_inplace_s / dwt_cdf53_2i_inplace_s
// (/root/reference/src/libdwt.c:12926, 13485, 13641, 14847, 17474, 16553, 17886).
//
// Coefficients of this family stay where the lifting leaves them: level j works on the samples at stride 2^j, even = L,
// odd = H.  On the device the pyramid is still computed level by level on DENSE LL bands with the streaming / tile
// kernels of the Mallat family; two cheap kernels translate between the layouts (k_ip_pack / k_ip_unpack).
//
// CDF 5/3 of this family is the Mallat result bit for bit (rows, then columns; :16583).  CDF 9/7 is not: the reference
// runs a level as up to eight sweeps -- "exceptions" (lines of 2..4 samples forward, 2..3 inverse: the whole 1-D
// transform at once; :10803, :11574), prolog, core, epilog -- each over all rows and then over all columns
// (:12975-13451, :17512-17598).  The top rows therefore get their column prolog BEFORE the row core, and the right
// columns their row epilog AFTER the column core; float rounding makes that order visible in the first 7 (forward) /
// 8 (inverse) rows and the last 5 columns of every level.  k_ip_phase evaluates exactly that schedule for any
// rectangle of a level: a tile plus a 4-sample halo is staged in shared memory and the eight sweeps are applied
// operation by operation, each lifting operation (step, position) belonging to the part the reference does it in
// (:9591 prolog, :9929 epilog, the rest core).  The host runs the ordinary level kernel and then k_ip_phase over the
// top and right frame (and over whole levels once they are small), so the result is bit-identical everywhere.
#include "kernels.h"
#include "lifting.cuh"
#include "tail_body.cuh"

namespace dwtb200 {

constexpr int IP_THREADS = 512;
constexpr int IP_HALO = 4;   // lifting depth: four steps per axis, whatever the order of the sweeps

enum { IP_X = 0, IP_P = 1, IP_C = 2, IP_E = 3 };

// forward steps: 0 alpha (odd), 1 beta (even), 2 gamma (odd), 3 delta (even), 4 scale; lines start at offset 1 (:10831)
__device__ __forceinline__ int ip_part_fwd(int N, int step, int i)
{
    if (N < 5) return IP_X;
    if ((step == 0 && (i == 1 || i == 3)) || (step == 1 && (i == 0 || i == 2)) || (step == 2 && i == 1) || (step == 3 && i == 0) ||
        (step == 4 && i == 0))
        return IP_P;
    if (N & 1) {
        if ((step == 1 && i == N - 1) || (step == 2 && i == N - 2) || (step == 3 && (i == N - 1 || i == N - 3)) || (step == 4 && i >= N - 4))
            return IP_E;
    } else {
        if ((step == 0 && i == N - 1) || (step == 1 && i == N - 2) || (step == 2 && (i == N - 1 || i == N - 3)) ||
            (step == 3 && (i == N - 2 || i == N - 4)) || (step == 4 && i >= N - 5))
            return IP_E;
    }
    return IP_C;
}
// inverse steps: 0 scale, 1 even += -u2, 2 odd += p2, 3 even += -u1, 4 odd += p1; offset 0 (:11604)
__device__ __forceinline__ int ip_part_inv(int N, int step, int i)
{
    if (N < 4) return IP_X;
    if ((step == 0 && i <= 3) || (step == 1 && (i == 0 || i == 2)) || (step == 2 && i == 1) || (step == 3 && i == 0)) return IP_P;
    if (N & 1) {
        if ((step == 0 && i == N - 1) || (step == 1 && i == N - 1) || (step == 2 && i == N - 2) || (step == 3 && (i == N - 1 || i == N - 3)) ||
            (step == 4 && (i == N - 2 || i == N - 4)))
            return IP_E;
    } else {
        if ((step == 2 && i == N - 1) || (step == 3 && i == N - 2) || (step == 4 && (i == N - 1 || i == N - 3))) return IP_E;
    }
    return IP_C;
}

template <bool INV> __device__ __forceinline__ float ip_coef(int lift)
{
    if (INV) return lift == 0 ? -W97F::U2 : lift == 1 ? W97F::P2 : lift == 2 ? -W97F::U1 : W97F::P1;
    return lift == 0 ? -W97F::P1 : lift == 1 ? W97F::U1 : lift == 2 ? -W97F::P2 : W97F::U2;
}

// positions [a0, a1) of a line of N samples that can hold operations of `part`
template <bool INV> __device__ __forceinline__ void ip_part_range(int N, int part, int &a0, int &a1)
{
    const int small = INV ? 4 : 5;
    a0 = 0;
    a1 = N;
    if (N < small) {
        if (part != IP_X) a1 = 0;
        return;
    }
    if (part == IP_X) a1 = 0;
    else if (part == IP_P) a1 = min(N, 4);
    else if (part == IP_E) a0 = max(0, N - 5);
}

// Per position of a line: which part does each of its five steps belong to (3 bits per step; 7 = the step does not
// touch this position).  Built once per staged window so that the sweeps are a table lookup per sample.
template <bool INV> __device__ __forceinline__ uint32_t ip_codes(int N, int i)
{
    uint32_t m = 0;
#pragma unroll
    for (int step = 0; step < 5; step++) {
        const bool scale = INV ? step == 0 : step == 4;
        const int lift = INV ? step - 1 : step;
        const int par = INV ? (lift & 1) : !(lift & 1);   // parity of the positions a lifting step updates
        const uint32_t code = (!scale && (i & 1) != par) ? 7u : (uint32_t)(INV ? ip_part_inv(N, step, i) : ip_part_fwd(N, step, i));
        m |= code << (3 * step);
    }
    return m;
}
// tx[0..w) / ty[0..h): codes of the window's columns / rows
template <bool INV> __device__ __forceinline__ void ip_build_tables(uint32_t *tx, uint32_t *ty, int w, int h, int lx0, int ly0, int nx, int ny)
{
    for (int e = threadIdx.x; e < w + h; e += blockDim.x) {
        if (e < w) tx[e] = ip_codes<INV>(nx, lx0 + e);
        else ty[e - w] = ip_codes<INV>(ny, ly0 + e - w);
    }
}

// One sweep (part, axis) over samples held in shared memory: sample (y, x) of the level lives at sm[y * ys + x * xs],
// the staged window covers level rows [ly0, ly0 + h) and columns [lx0, lx0 + w) (a whole level, or a tile plus its halo).
// Warps walk the rows, lanes the columns.
template <bool INV, bool ALONG_X>
__device__ __forceinline__ void ip_sweep(float *sm, const uint32_t *tab, int ys, int xs, int w, int h, int lx0, int ly0, int N, int part)
{
    int a0, a1;
    ip_part_range<INV>(N, part, a0, a1);
    const int o = ALONG_X ? lx0 : ly0, tn = ALONG_X ? w : h, d = ALONG_X ? xs : ys;
    const int t0 = max(a0 - o, 0), t1 = min(a1 - o, tn);   // window indices along the line
    if (t1 <= t0) return;   // block-uniform
    const int x0 = ALONG_X ? t0 : 0, x1 = ALONG_X ? t1 : w, y0 = ALONG_X ? 0 : t0, y1 = ALONG_X ? h : t1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (int step = 0; step < 5; step++) {
        const bool scale = INV ? step == 0 : step == 4;
        const float c = scale ? 0.f : ip_coef<INV>(INV ? step - 1 : step);
        for (int y = y0 + warp; y < y1; y += nwarps) {
            for (int x = x0 + lane; x < x1; x += 32) {
                const int t = ALONG_X ? x : y, i = o + t;   // window index and position along the line
                if (((tab[t] >> (3 * step)) & 7u) != (uint32_t)part) continue;
                float *q = sm + y * ys + x * xs;
                if (scale) {
                    *q = __fmul_rn(*q, ((i & 1) != INV) ? W97F::IZ : W97F::Z);   // forward: even * zeta, odd / zeta; inverse the other way
                    continue;
                }
                float l, r;
                if (i == 0) {   // whole-sample mirror: (2c) * neighbour == c * (nb + nb)
                    if (t + 1 >= tn) continue;
                    l = r = q[d];
                } else if (i == N - 1) {
                    if (t < 1) continue;
                    l = r = q[-d];
                } else {
                    if (t < 1 || t + 1 >= tn) continue;   // neighbour outside the staged window: this sample is halo, its value is not used
                    l = q[-d];
                    r = q[d];
                }
                *q = __fadd_rn(*q, __fmul_rn(c, __fadd_rn(l, r)));
            }
        }
        __syncthreads();
    }
}

// the eight sweeps of one level, in the reference's order (:12975-13451, :17512-17598)
template <bool INV>
__device__ __forceinline__ void ip_level_sweeps(float *sm, uint32_t *tx, uint32_t *ty, int ys, int xs, int w, int h, int lx0, int ly0, int nx, int ny)
{
    ip_build_tables<INV>(tx, ty, w, h, lx0, ly0, nx, ny);
    __syncthreads();
    for (int part = IP_X; part <= IP_E; part++) {
        if (nx > 1) ip_sweep<INV, true>(sm, tx, ys, xs, w, h, lx0, ly0, nx, part);
        if (ny > 1) ip_sweep<INV, false>(sm, ty, ys, xs, w, h, lx0, ly0, ny, part);
    }
}

// up to two rectangles of a level per launch (the top and the right frame), each cut into tiles of tw x th outputs
struct IpRects {
    int n;
    int x0[2], y0[2], x1[2], y1[2], tw[2], th[2], ntx[2], nt[2];
};

// One sample per thread: the window (tile + halo) holds at most IP_THREADS samples, every thread keeps its sample's
// position codes in registers and the 8 x 5 rounds are straight-line code -- the kernel is latency-bound (a few hundred
// samples per CTA), so what counts is instructions per round, not throughput.
template <bool INV, bool ALONG_X>
__device__ __forceinline__ void ip_round1(float *q, bool live, uint32_t codes, int i, int t, int tn, int d, int N, int part)
{
#pragma unroll
    for (int step = 0; step < 5; step++) {
        const bool scale = INV ? step == 0 : step == 4;
        if (live && ((codes >> (3 * step)) & 7u) == (uint32_t)part) {
            if (scale) {
                *q = __fmul_rn(*q, ((i & 1) != INV) ? W97F::IZ : W97F::Z);
            } else {
                const float c = ip_coef<INV>(INV ? step - 1 : step);
                bool ok = true;
                float l, r;
                if (i == 0) {
                    ok = t + 1 < tn;
                    l = r = ok ? q[d] : 0.f;
                } else if (i == N - 1) {
                    ok = t >= 1;
                    l = r = ok ? q[-d] : 0.f;
                } else {
                    ok = t >= 1 && t + 1 < tn;   // neighbour outside the staged window: this sample is halo, its value is not used
                    l = ok ? q[-d] : 0.f;
                    r = ok ? q[d] : 0.f;
                }
                if (ok) *q = __fadd_rn(*q, __fmul_rn(c, __fadd_rn(l, r)));
            }
        }
        __syncthreads();
    }
}

template <bool INV> __global__ void __launch_bounds__(IP_THREADS) k_ip_phase(const LevelParams p, const IpRects rc)
{
    __shared__ float sm[IP_THREADS];
    const int frame = blockIdx.z;
    const int nx = p.W, ny = p.H;
    int b = blockIdx.x, r = 0;
    if (b >= rc.nt[0]) {
        b -= rc.nt[0];
        r = 1;
    }
    const int tw = rc.tw[r], th = rc.th[r];
    const int ox0 = rc.x0[r] + (b % rc.ntx[r]) * tw, oy0 = rc.y0[r] + (b / rc.ntx[r]) * th;
    const int ox1 = min(ox0 + tw, rc.x1[r]), oy1 = min(oy0 + th, rc.y1[r]);
    const int lx0 = max(ox0 - IP_HALO, 0), lx1 = min(ox1 + IP_HALO, nx), ly0 = max(oy0 - IP_HALO, 0), ly1 = min(oy1 + IP_HALO, ny);
    const int w = lx1 - lx0, h = ly1 - ly0;
    const int e = threadIdx.x, ty = e / w, tx = e - ty * w, gy = ly0 + ty, gx = lx0 + tx;
    const bool live = e < w * h;
    float *q = sm + e;   // the window is stored densely: pitch w
    if (live) {
        float v;
        if (!INV) {
            v = ((const float *)p.src)[(size_t)frame * p.src_frame + (size_t)gy * p.src_pitch + gx];
        } else if (p.il && ((gy | gx) & 1)) {   // interleaved source (level 0 of the ring path): HL, LH, HH lie where they belong
            v = ((const float *)p.il)[(size_t)frame * p.il_frame + (size_t)gy * p.il_pitch + gx];
        } else {
            const int by = gy >> 1, bx = gx >> 1;
            const float *bp = (gy & 1) ? ((gx & 1) ? (const float *)p.hh : (const float *)p.lh) : ((gx & 1) ? (const float *)p.hl : nullptr);
            if (bp) v = bp[(size_t)frame * p.sub_frame + (size_t)by * p.sub_pitch + bx];
            else v = ((const float *)p.ll)[(size_t)frame * p.ll_frame + (size_t)by * p.ll_pitch + bx];
        }
        *q = v;
    }
    const uint32_t cx = live ? ip_codes<INV>(nx, gx) : 0x7fffu, cy = live ? ip_codes<INV>(ny, gy) : 0x7fffu;
    __syncthreads();
    for (int part = IP_X; part <= IP_E; part++) {
        int a0, a1;
        ip_part_range<INV>(nx, part, a0, a1);
        if (nx > 1 && max(a0, lx0) < min(a1, lx1)) ip_round1<INV, true>(q, live, cx, gx, tx, w, 1, nx, part);   // block-uniform tests
        ip_part_range<INV>(ny, part, a0, a1);
        if (ny > 1 && max(a0, ly0) < min(a1, ly1)) ip_round1<INV, false>(q, live, cy, gy, ty, h, w, ny, part);
    }
    if (live && gx >= ox0 && gx < ox1 && gy >= oy0 && gy < oy1) {
        const float v = *q;
        if (INV) {
            ((float *)p.dst)[(size_t)frame * p.dst_frame + (size_t)gy * p.dst_pitch + gx] = v;
        } else if (p.il) {
            ((float *)p.il)[(size_t)frame * p.il_frame + (size_t)gy * p.il_pitch + gx] = v;
            if (!((gy | gx) & 1)) ((float *)p.ll)[(size_t)frame * p.ll_frame + (size_t)(gy >> 1) * p.ll_pitch + (gx >> 1)] = v;
        } else {
            const int by = gy >> 1, bx = gx >> 1;
            float *o;
            if (gy & 1) o = ((gx & 1) ? (float *)p.hh : (float *)p.lh) + (size_t)frame * p.sub_frame + (size_t)by * p.sub_pitch + bx;
            else o = (gx & 1) ? (float *)p.hl + (size_t)frame * p.sub_frame + (size_t)by * p.sub_pitch + bx
                              : (float *)p.ll + (size_t)frame * p.ll_frame + (size_t)by * p.ll_pitch + bx;
            *o = v;
        }
    }
}

static void ip_add_rect(IpRects &rc, int x0, int y0, int x1, int y1)
{
    if (x1 <= x0 || y1 <= y0) return;
    const int rw = x1 - x0, rh = y1 - y0, i = rc.n++;
    int tw, th;   // (tw + 8) * (th + 8) <= IP_THREADS samples
    if (rh <= 8) { th = rh; tw = IP_THREADS / (rh + 2 * IP_HALO) - 2 * IP_HALO; }        // the top frame: a few full-width rows
    else if (rw <= 8) { tw = rw; th = IP_THREADS / (rw + 2 * IP_HALO) - 2 * IP_HALO; }   // the right frame
    else { tw = 14; th = 14; }                                                            // a whole level
    if (tw > rw) tw = rw;
    if (th > rh) th = rh;
    rc.x0[i] = x0; rc.y0[i] = y0; rc.x1[i] = x1; rc.y1[i] = y1;
    rc.tw[i] = tw; rc.th[i] = th;
    rc.ntx[i] = (rw + tw - 1) / tw;
    rc.nt[i] = rc.ntx[i] * ((rh + th - 1) / th);
}
// rectangle [rx0, rx1) x [ry0, ry1) of a level, and optionally a second one, in one launch
void launch_ip_phase(bool inverse, const LevelParams &p, int frames, int rx0, int ry0, int rx1, int ry1, int sx0, int sy0, int sx1, int sy1,
                     cudaStream_t st)
{
    IpRects rc;
    memset(&rc, 0, sizeof rc);
    ip_add_rect(rc, rx0, ry0, rx1, ry1);
    ip_add_rect(rc, sx0, sy0, sx1, sy1);
    if (!rc.n) return;
    const dim3 grid(rc.nt[0] + rc.nt[1], 1, frames);
    if (inverse) k_ip_phase<true><<<grid, IP_THREADS, 0, st>>>(p, rc);
    else k_ip_phase<false><<<grid, IP_THREADS, 0, st>>>(p, rc);
}

// ---- all remaining small levels in one launch: one CTA per frame, truly in place in shared memory ------------------
// `buf` holds LL_{j0-1} (w0 x h0, dense) on entry of the forward kernel and the interleaved pyramid of the levels j0 .. J-1
// on exit (level j at stride 2^(j - j0)); the inverse kernel goes the other way.
constexpr int IP_TAIL_THREADS = 1024;
constexpr int IP_TAIL_CAP = IP_TAIL_THREADS;   // samples of the tail's first level (32 x 32): one per thread
template <bool INV> __global__ void __launch_bounds__(IP_TAIL_THREADS) k_ip_tail(float *buf, int64_t pitch, int64_t frame, int w0, int h0, int nlev)
{
    __shared__ float sm[2 * IP_TAIL_CAP];   // (w0 | 1) * h0 <= 2 * w0 * h0
    float *g = buf + (size_t)blockIdx.x * frame;
    const int e = threadIdx.x;
    const int sp = w0 | 1;   // odd pitch
    {
        const int y = e / w0, x = e - y * w0;
        if (e < w0 * h0) sm[y * sp + x] = g[(size_t)y * pitch + x];
    }
    __syncthreads();
    for (int qq = 0; qq < nlev; qq++) {
        const int k = INV ? nlev - 1 - qq : qq;
        const int w = cdiv_pow2(w0, k), h = cdiv_pow2(h0, k);   // this level's samples sit at stride 2^k
        const int y = e / w, x = e - y * w;
        const bool live = e < w * h;
        float *q = sm + (y << k) * sp + (x << k);
        const uint32_t cx = live ? ip_codes<INV>(w, x) : 0x7fffu, cy = live ? ip_codes<INV>(h, y) : 0x7fffu;
        for (int part = IP_X; part <= IP_E; part++) {
            int a0, a1;
            ip_part_range<INV>(w, part, a0, a1);
            if (w > 1 && a0 < a1) ip_round1<INV, true>(q, live, cx, x, x, w, 1 << k, w, part);
            ip_part_range<INV>(h, part, a0, a1);
            if (h > 1 && a0 < a1) ip_round1<INV, false>(q, live, cy, y, y, h, sp << k, h, part);
        }
    }
    {
        const int y = e / w0, x = e - y * w0;
        if (e < w0 * h0) g[(size_t)y * pitch + x] = sm[y * sp + x];
    }
}
int ip_tail_cap() { return IP_TAIL_CAP; }
void launch_ip_tail(bool inverse, void *buf, int64_t pitch, int64_t frame, int w0, int h0, int nlev, int frames, cudaStream_t st)
{
    if (inverse) k_ip_tail<true><<<frames, IP_TAIL_THREADS, 0, st>>>((float *)buf, pitch, frame, w0, h0, nlev);
    else k_ip_tail<false><<<frames, IP_TAIL_THREADS, 0, st>>>((float *)buf, pitch, frame, w0, h0, nlev);
}

// ---- layout translation: Mallat pyramid of J levels <-> interleaved (4-byte elements) ------------------------------
// Interleaved sample (Y, X): the lowest set bit of Y | X names its level; below bit J it is a sample of LL_J.  Samples
// whose level is >= jt (multiples of 2^jt in both coordinates) live in the dense "tail block" instead (k_ip_tail),
// at (Y >> jt, X >> jt).
struct IpPack {
    const void *src;
    void *dst;
    void *tail;            // tail block (nullptr: none), dense with pitch / frame stride tpitch / tframe
    int64_t pitch, frame, tpitch, tframe;
    int ox, oy, J, jt;
    int shift;             // only the samples at multiples of 2^shift in both coordinates are translated (1: level 0 is already interleaved)
    int wl[32], hl[32];    // size of LL_l+1: column / row origin of the H bands of level l
};
constexpr int IP_PACK_ROWS = 4;   // rows per thread: the column's part of the index arithmetic is shared (the kernel is issue-bound)
template <bool UNPACK> __global__ void __launch_bounds__(256) k_ip_pack(const IpPack p)
{
    const int X = (blockIdx.x * 256 + threadIdx.x) << p.shift;
    if (X >= p.ox) return;
    const size_t f = (size_t)blockIdx.z * p.frame, ft = (size_t)blockIdx.z * p.tframe;
    const uint32_t *src = (const uint32_t *)p.src;
    uint32_t *dst = (uint32_t *)p.dst, *tail = (uint32_t *)p.tail;
    const int cx = X ? __ffs(X) - 1 : 31;
    const int jt = tail ? p.jt : 32;
#pragma unroll
    for (int r = 0; r < IP_PACK_ROWS; r++) {
        const int Y = (blockIdx.y * IP_PACK_ROWS + r) << p.shift;
        if (Y >= p.oy) break;
        const int lvl = min(cx, Y ? __ffs(Y) - 1 : 31);
        const size_t i = f + (size_t)Y * p.pitch + X;
        if (lvl >= jt) {
            const size_t m = ft + (size_t)(Y >> jt) * p.tpitch + (X >> jt);
            if (UNPACK) tail[m] = src[i];
            else dst[i] = tail[m];
            continue;
        }
        int row, col;
        if (lvl >= p.J) {
            row = Y >> p.J;
            col = X >> p.J;
        } else {
            row = Y >> (lvl + 1);
            col = X >> (lvl + 1);
            if ((Y >> lvl) & 1) row += p.hl[lvl];
            if ((X >> lvl) & 1) col += p.wl[lvl];
        }
        const size_t m = f + (size_t)row * p.pitch + col;
        if (UNPACK) dst[m] = src[i];
        else dst[i] = src[m];
    }
}
void launch_ip_pack(bool unpack, const void *src, void *dst, int64_t pitch, int64_t frame, int ox, int oy, int J, void *tail, int64_t tpitch,
                    int64_t tframe, int jt, int shift, int frames, cudaStream_t st)
{
    IpPack p;
    p.src = src; p.dst = dst; p.tail = tail;
    p.pitch = pitch; p.frame = frame; p.tpitch = tpitch; p.tframe = tframe;
    p.ox = ox; p.oy = oy; p.J = J; p.jt = jt; p.shift = shift;
    for (int l = 0; l < 32; l++) {
        p.wl[l] = cdiv_pow2(ox, l + 1);
        p.hl[l] = cdiv_pow2(oy, l + 1);
    }
    const int nx = cdiv_pow2(ox, shift), ny = cdiv_pow2(oy, shift);
    const dim3 grid((nx + 255) / 256, (ny + IP_PACK_ROWS - 1) / IP_PACK_ROWS, frames);
    if (unpack) k_ip_pack<true><<<grid, 256, 0, st>>>(p);
    else k_ip_pack<false><<<grid, 256, 0, st>>>(p);
}

cudaError_t preload_inplace()
{
    cudaFuncAttributes a;
    cudaError_t e = cudaFuncGetAttributes(&a, k_ip_phase<false>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, k_ip_phase<true>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, k_ip_pack<false>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, k_ip_pack<true>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, k_ip_tail<false>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, k_ip_tail<true>);
    return e;
}

}  // namespace dwtb200
