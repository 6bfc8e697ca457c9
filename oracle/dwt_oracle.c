/*
 * oracle/dwt_oracle.c -- CPU restatement of libdwt's separable lifting DWT hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under libdwt_b200/ may link, load or call this
 * file; it exists so tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg can
 * check the CUDA path.  Parity status: PINNED -- tests/test_oracle.py compares
 * every function here bit-for-bit against the compiled reference (oracle/_ref, built
 * from /root/reference by oracle/Makefile) and tests/golden/ holds digests of reference
 * outputs produced by tests/golden/make_golden.py.
 *
 * It is a restatement, not a copy: one textbook lifting routine per (wavelet, type,
 * direction) working on a gathered line with whole-sample mirror boundaries, 64-bit
 * addressing everywhere (the reference computes y*stride_x in `int`, src/inline.h:188,
 * and cannot address images >= 2 GiB), and one generic level driver.
 *
 * Reference lines followed (relative to /root/reference/):
 *   drivers   src/libdwt.c:12776 (2f_s) 17040 (2i_s) 12451 (2f_d) 16884 (2i_d)
 *             16304 (53 2f_i) 18142 (53 2i_i)
 *   lines     src/libdwt.c:10744 11530 (float; cores 2264, 9510, 9844, 10199)
 *             2024 11424 (double)   10950 11749 (int 5/3)
 *   siblings  drivers src/libdwt.c:16470 18296 (5/3 float) 12535 16962 (5/3 double) 16387 18219 (9/7 int);
 *             lines 10986 11785, 2085 11484, 10901 11699; constants src/inline.h:333-341
 *   padding   src/libdwt.c:12080-12215
 *   constants src/inline.h:310-323, helpers 443-460, 590-607
 *   patterns  src/libdwt.c:1112-1244, fills 1247-1385, src/volume.c:41
 *   in-place  src/libdwt.c:12926 13485 13641 14847 (9/7 fwd) 17474 (9/7 inv) 16553 17886 (5/3 float); parts 9591 9929
 *             10803-10898 11574-11670; lines 11032 11831
 *   3-D       src/volume-dwt.c:727 (fwd, via src/dwt-simple.c:2166, 580, 981, 1469)
 *             src/volume-dwt.c:1115 (inv, via src/libdwt.c:17182)
 */
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ---- constants (src/inline.h:310-323); float ones are the double literals rounded to float ---- */
static const float  P1s = 1.58613434342059;
static const float  U1s = -0.0529801185729;
static const float  P2s = -0.8829110755309;
static const float  U2s = 0.4435068520439;
static const float  S1s = 1.1496043988602;
static const float  S2s = 1 / 1.1496043988602; /* double division, then rounded: 0x3f5eaf70 */
static const double P1d = 1.58613434342059;
static const double U1d = -0.0529801185729;
static const double P2d = -0.8829110755309;
static const double U2d = 0.4435068520439;
static const double S1d = 1.1496043988602;
static const double S2d = 1 / 1.1496043988602;

/* ---- integer helpers (src/inline.h:443-460, 590-607) ---- */
static int orc_ceil_div_pow2(int i, int j) { return (int)(((int64_t)i + ((int64_t)1 << j) - 1) >> j); }
int orc_ceil_log2(int x) { int j = 0; while (((int64_t)1 << j) < x) j++; return j; }
static int half_up(int n) { return (n + 1) >> 1; }
static int half_dn(int n) { return n >> 1; }

/* =====================================================================================
 * 1-D lifting on a gathered line t[0..N), N >= 2, interleaved in place (even = L, odd = H).
 * Mirror boundary written the way the reference writes it: edge sample gets (2c)*neighbour.
 * ===================================================================================== */
#define DEF_LIFT(T, SUF)                                                                   \
    static void lift_odd_##SUF(T *t, int N, T c)                                           \
    {                                                                                      \
        for (int i = 1; i + 1 < N; i += 2) t[i] += c * (t[i - 1] + t[i + 1]);              \
        if (!(N & 1)) t[N - 1] += (2 * c) * t[N - 2];                                      \
    }                                                                                      \
    static void lift_even_##SUF(T *t, int N, T c)                                          \
    {                                                                                      \
        t[0] += (2 * c) * t[1];                                                            \
        for (int i = 2; i + 1 < N; i += 2) t[i] += c * (t[i - 1] + t[i + 1]);              \
        if (N & 1) t[N - 1] += (2 * c) * t[N - 2];                                         \
    }
DEF_LIFT(float, s)
DEF_LIFT(double, d)

/* float forward: libdwt.c:10780 coefficients (-p1,u1,-p2,u2), then even*=zeta, odd*=1/zeta
 * with 1/zeta evaluated in float (libdwt.c:2289) */
static void fwd97_s(void *v, int N)
{
    float *t = (float *)v;
    lift_odd_s(t, N, -P1s);
    lift_even_s(t, N, U1s);
    lift_odd_s(t, N, -P2s);
    lift_even_s(t, N, U2s);
    const float z = S1s, iz = 1 / z;
    for (int i = 0; i < N; i += 2) t[i] *= z;
    for (int i = 1; i < N; i += 2) t[i] *= iz;
}
/* float inverse: libdwt.c:11561 (-u2,p2,-u1,p1), descale first (libdwt.c:2289-2301) */
static void inv97_s(void *v, int N)
{
    float *t = (float *)v;
    const float z = S1s, iz = 1 / z;
    for (int i = 0; i < N; i += 2) t[i] *= iz;
    for (int i = 1; i < N; i += 2) t[i] *= z;
    lift_even_s(t, N, -U2s);
    lift_odd_s(t, N, P2s);
    lift_even_s(t, N, -U1s);
    lift_odd_s(t, N, P1s);
}
static void one_fwd97_s(void *v) { *(float *)v = *(float *)v * S1s; } /* libdwt.c:10757 */
static void one_inv97_s(void *v) { *(float *)v = *(float *)v * S2s; } /* libdwt.c:11546 */

/* double: libdwt.c:2024, 11424.  t -= p*(l+r) is bit-equal to t += (-p)*(l+r). */
static void fwd97_d(void *v, int N)
{
    double *t = (double *)v;
    lift_odd_d(t, N, -P1d);
    lift_even_d(t, N, U1d);
    lift_odd_d(t, N, -P2d);
    lift_even_d(t, N, U2d);
    for (int i = 0; i < N; i += 2) t[i] = t[i] * S1d;
    for (int i = 1; i < N; i += 2) t[i] = t[i] * S2d;
}
static void inv97_d(void *v, int N)
{
    double *t = (double *)v;
    for (int i = 0; i < N; i += 2) t[i] = t[i] * S2d;
    for (int i = 1; i < N; i += 2) t[i] = t[i] * S1d;
    lift_even_d(t, N, -U2d);
    lift_odd_d(t, N, P2d);
    lift_even_d(t, N, -U1d);
    lift_odd_d(t, N, P1d);
}
static void one_fwd97_d(void *v) { *(double *)v = *(double *)v * S1d; } /* libdwt.c:2036 */
static void one_inv97_d(void *v) { *(double *)v = *(double *)v * S2d; } /* libdwt.c:11437 */

/* int 5/3: libdwt.c:10950, 11749.  Adds are done in uint32_t so that wrap-around is
 * defined; >> on a negative int32_t is arithmetic with gcc, as in the reference. */
static int32_t wadd(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
static int32_t wsub(int32_t a, int32_t b) { return (int32_t)((uint32_t)a - (uint32_t)b); }
static void fwd53_i(void *v, int N)
{
    int32_t *t = (int32_t *)v;
    for (int i = 1; i < N - 2 + (N & 1); i += 2) t[i] = wsub(t[i], wadd(t[i - 1], t[i + 1]) >> 1);
    if (N & 1) t[N - 1] = wadd(t[N - 1], wadd(t[N - 2], 1) >> 1);
    else       t[N - 1] = wsub(t[N - 1], t[N - 2]);
    t[0] = wadd(t[0], wadd(t[1], 1) >> 1);
    for (int i = 2; i < N - (N & 1); i += 2) t[i] = wadd(t[i], wadd(wadd(t[i - 1], t[i + 1]), 2) >> 2);
}
static void inv53_i(void *v, int N)
{
    int32_t *t = (int32_t *)v;
    for (int i = 2; i < N - (N & 1); i += 2) t[i] = wsub(t[i], wadd(wadd(t[i - 1], t[i + 1]), 2) >> 2);
    t[0] = wsub(t[0], wadd(t[1], 1) >> 1);
    if (N & 1) t[N - 1] = wsub(t[N - 1], wadd(t[N - 2], 1) >> 1);
    else       t[N - 1] = wadd(t[N - 1], t[N - 2]);
    for (int i = 1; i < N - 2 + (N & 1); i += 2) t[i] = wadd(t[i], wadd(t[i - 1], t[i + 1]) >> 1);
}

/* ---- sibling transforms ----
 * 5/3 float / double: libdwt.c:10986, 11785, 2085, 11484.  p1 = 0.5, u1 = 0.25, even *= sqrt2, odd *= 1/sqrt2
 * (both written as decimal literals, src/inline.h:333-341); edges are (2c)*neighbour like 9/7. */
static const float  P53s = 0.5, U53s = 0.25, S53_1s = 1.41421356237309504880, S53_2s = 0.70710678118654752440;
static const double P53d = 0.5, U53d = 0.25, S53_1d = 1.41421356237309504880, S53_2d = 0.70710678118654752440;
#define DEF_53F(T, SUF, P, U, S1, S2)                                  \
    static void fwd53_##SUF(void *v, int N)                            \
    {                                                                  \
        T *t = (T *)v;                                                 \
        lift_odd_##SUF(t, N, -P);                                      \
        lift_even_##SUF(t, N, U);                                      \
        for (int i = 0; i < N; i += 2) t[i] = t[i] * S1;               \
        for (int i = 1; i < N; i += 2) t[i] = t[i] * S2;               \
    }                                                                  \
    static void inv53_##SUF(void *v, int N)                            \
    {                                                                  \
        T *t = (T *)v;                                                 \
        for (int i = 0; i < N; i += 2) t[i] = t[i] * S2;               \
        for (int i = 1; i < N; i += 2) t[i] = t[i] * S1;               \
        lift_even_##SUF(t, N, -U);                                     \
        lift_odd_##SUF(t, N, P);                                       \
    }                                                                  \
    static void one_fwd53_##SUF(void *v) { *(T *)v = *(T *)v * S1; }   \
    static void one_inv53_##SUF(void *v) { *(T *)v = *(T *)v * S2; }
DEF_53F(float, s, P53s, U53s, S53_1s, S53_2s)
DEF_53F(double, d, P53d, U53d, S53_1d, S53_2d)

/* 9/7 with integer lifting: libdwt.c:10901, 11699.  Steps ( c*(l+r) - 64 ) >> 7 subtracted from odd samples
 * (c = 203, -113) and ( c*(l+r) + 2048 ) >> 12 added to even samples (c = -217, 1817); mirrored edges use
 * (nb + nb); 32-bit wrap-around made explicit; no scaling; N < 2 untouched. */
static int32_t q7(int32_t c, int32_t l, int32_t r) { return (int32_t)((uint32_t)c * ((uint32_t)l + (uint32_t)r) - 64u) >> 7; }
static int32_t q12(int32_t c, int32_t l, int32_t r) { return (int32_t)((uint32_t)c * ((uint32_t)l + (uint32_t)r) + 2048u) >> 12; }
static int32_t left_of(const int32_t *t, int i) { return t[i ? i - 1 : 1]; }
static int32_t right_of(const int32_t *t, int i, int N) { return t[i + 1 < N ? i + 1 : N - 2]; }
static void fwd97_i(void *v, int N)
{
    int32_t *t = (int32_t *)v;
    for (int i = 1; i < N; i += 2) t[i] = wsub(t[i], q7(203, left_of(t, i), right_of(t, i, N)));
    for (int i = 0; i < N; i += 2) t[i] = wadd(t[i], q12(-217, left_of(t, i), right_of(t, i, N)));
    for (int i = 1; i < N; i += 2) t[i] = wsub(t[i], q7(-113, left_of(t, i), right_of(t, i, N)));
    for (int i = 0; i < N; i += 2) t[i] = wadd(t[i], q12(1817, left_of(t, i), right_of(t, i, N)));
}
static void inv97_i(void *v, int N)
{
    int32_t *t = (int32_t *)v;
    for (int i = 0; i < N; i += 2) t[i] = wsub(t[i], q12(1817, left_of(t, i), right_of(t, i, N)));
    for (int i = 1; i < N; i += 2) t[i] = wadd(t[i], q7(-113, left_of(t, i), right_of(t, i, N)));
    for (int i = 0; i < N; i += 2) t[i] = wsub(t[i], q12(-217, left_of(t, i), right_of(t, i, N)));
    for (int i = 1; i < N; i += 2) t[i] = wadd(t[i], q7(203, left_of(t, i), right_of(t, i, N)));
}

/* =====================================================================================
 * Generic Mallat-layout level driver.
 * ===================================================================================== */
typedef struct {
    size_t esz;
    void (*fwd)(void *, int);
    void (*inv)(void *, int);
    void (*one_fwd)(void *); /* N==1 action, NULL = leave untouched (int, libdwt.c:10961) */
    void (*one_inv)(void *);
    int guard; /* float drivers skip a pass when the OUTER line length is <= 1 (libdwt.c:12837) */
    int inv_cols_first; /* int inverse: columns then rows (libdwt.c:18178) */
} kind_t;

static const kind_t K97S = {4, fwd97_s, inv97_s, one_fwd97_s, one_inv97_s, 1, 0};
static const kind_t K97D = {8, fwd97_d, inv97_d, one_fwd97_d, one_inv97_d, 0, 0};
static const kind_t K53I = {4, fwd53_i, inv53_i, NULL, NULL, 0, 1};
static const kind_t K53S = {4, fwd53_s, inv53_s, one_fwd53_s, one_inv53_s, 0, 0}; /* libdwt.c:16508, 18333: no guard, rows first */
static const kind_t K53D = {8, fwd53_d, inv53_d, one_fwd53_d, one_inv53_d, 0, 0}; /* libdwt.c:12574, 16999 */
static const kind_t K97I = {4, fwd97_i, inv97_i, NULL, NULL, 0, 1};               /* libdwt.c:18256: columns first */

static void gather(void *dst, ptrdiff_t dstep, const void *src, ptrdiff_t sstep, int n, size_t esz)
{
    char *d = (char *)dst;
    const char *s = (const char *)src;
    for (int i = 0; i < n; i++) memcpy(d + (ptrdiff_t)i * dstep, s + (ptrdiff_t)i * sstep, esz);
}

/* forward line: src -> (L at dst_l, H at dst_h), element step `step` bytes */
static void line_fwd(const kind_t *k, char *src, char *dst_l, char *dst_h, void *tmp, int N, ptrdiff_t step)
{
    if (N < 2) {
        if (N == 1 && k->one_fwd) {
            memcpy(tmp, src, k->esz);
            k->one_fwd(tmp);
            memcpy(dst_l, tmp, k->esz);
        }
        return;
    }
    gather(tmp, (ptrdiff_t)k->esz, src, step, N, k->esz);
    k->fwd(tmp, N);
    gather(dst_l, step, tmp, 2 * (ptrdiff_t)k->esz, half_up(N), k->esz);
    gather(dst_h, step, (char *)tmp + k->esz, 2 * (ptrdiff_t)k->esz, half_dn(N), k->esz);
}
static void line_inv(const kind_t *k, char *src_l, char *src_h, char *dst, void *tmp, int N, ptrdiff_t step)
{
    if (N < 2) {
        if (N == 1 && k->one_inv) {
            memcpy(tmp, src_l, k->esz);
            k->one_inv(tmp);
            memcpy(dst, tmp, k->esz);
        }
        return;
    }
    gather(tmp, 2 * (ptrdiff_t)k->esz, src_l, step, half_up(N), k->esz);
    gather((char *)tmp + k->esz, 2 * (ptrdiff_t)k->esz, src_h, step, half_dn(N), k->esz);
    k->inv(tmp, N);
    gather(dst, step, tmp, (ptrdiff_t)k->esz, N, k->esz);
}
static void zero_run(char *p, ptrdiff_t step, int n, size_t esz)
{
    for (int i = 0; i < n; i++) memset(p + (ptrdiff_t)i * step, 0, esz);
}

/* lines along x (rows): line y starts at ptr + y*sx, element step sy; along y: swap roles */
#define AT(p, y, x) ((char *)(p) + (ptrdiff_t)(y) * sx + (ptrdiff_t)(x) * sy)

static void pass_fwd(const kind_t *k, void *ptr, ptrdiff_t sx, ptrdiff_t sy, int rows, int nlines, int N, int off_h)
{
    /* rows!=0: nlines rows, transform along x; else nlines columns, transform along y */
#pragma omp parallel
    {
        void *tmp = malloc((size_t)(N > 2 ? N : 2) * k->esz);
#pragma omp for schedule(static)
        for (int l = 0; l < nlines; l++) {
            if (rows) line_fwd(k, AT(ptr, l, 0), AT(ptr, l, 0), AT(ptr, l, off_h), tmp, N, sy);
            else      line_fwd(k, AT(ptr, 0, l), AT(ptr, 0, l), AT(ptr, off_h, l), tmp, N, sx);
        }
        free(tmp);
    }
}
static void pass_inv(const kind_t *k, void *ptr, ptrdiff_t sx, ptrdiff_t sy, int rows, int nlines, int N, int off_h)
{
#pragma omp parallel
    {
        void *tmp = malloc((size_t)(N > 2 ? N : 2) * k->esz);
#pragma omp for schedule(static)
        for (int l = 0; l < nlines; l++) {
            if (rows) line_inv(k, AT(ptr, l, 0), AT(ptr, l, off_h), AT(ptr, l, 0), tmp, N, sy);
            else      line_inv(k, AT(ptr, 0, l), AT(ptr, off_h, l), AT(ptr, 0, l), tmp, N, sx);
        }
        free(tmp);
    }
}

static void drive_fwd(const kind_t *k, void *ptr, ptrdiff_t sx, ptrdiff_t sy, int ox, int oy, int ix, int iy,
                      int *j_max_ptr, int decompose_one, int zero_padding)
{
    const int omin = ox < oy ? ox : oy, omax = ox < oy ? oy : ox;
    const int j_limit = orc_ceil_log2(decompose_one ? omax : omin);
    if (*j_max_ptr < 0 || *j_max_ptr > j_limit) *j_max_ptr = j_limit;
    for (int j = 0; j < *j_max_ptr; j++) {
        const int osx = orc_ceil_div_pow2(ox, j), osy = orc_ceil_div_pow2(oy, j);
        const int odx = orc_ceil_div_pow2(ox, j + 1), ody = orc_ceil_div_pow2(oy, j + 1);
        const int isx = orc_ceil_div_pow2(ix, j), isy = orc_ceil_div_pow2(iy, j);
        if (!k->guard || osx > 1) pass_fwd(k, ptr, sx, sy, 1, osy, isx, odx);
        if (!k->guard || osy > 1) pass_fwd(k, ptr, sx, sy, 0, osx, isy, ody);
        if (zero_padding) {
            /* libdwt.c:12896-12916 + 12080: only when a destination length is non-zero */
            if (odx || (osx - odx))
                for (int y = 0; y < osy; y++) {
                    zero_run(AT(ptr, y, half_up(isx)), sy, odx - half_up(isx), k->esz);
                    zero_run(AT(ptr, y, odx + half_dn(isx)), sy, (osx - odx) - half_dn(isx), k->esz);
                }
            if (ody || (osy - ody))
                for (int x = 0; x < osx; x++) {
                    zero_run(AT(ptr, half_up(isy), x), sx, ody - half_up(isy), k->esz);
                    zero_run(AT(ptr, ody + half_dn(isy), x), sx, (osy - ody) - half_dn(isy), k->esz);
                }
        }
    }
}
static void drive_inv(const kind_t *k, void *ptr, ptrdiff_t sx, ptrdiff_t sy, int ox, int oy, int ix, int iy,
                      int j_max, int decompose_one, int zero_padding)
{
    const int omin = ox < oy ? ox : oy, omax = ox < oy ? oy : ox;
    int j = orc_ceil_log2(decompose_one ? omax : omin);
    if (j_max >= 0 && j_max < j) j = j_max;
    for (; j > 0; j--) {
        const int osx = orc_ceil_div_pow2(ox, j), osy = orc_ceil_div_pow2(oy, j);
        const int odx = orc_ceil_div_pow2(ox, j - 1), ody = orc_ceil_div_pow2(oy, j - 1);
        const int idx = orc_ceil_div_pow2(ix, j - 1), idy = orc_ceil_div_pow2(iy, j - 1);
        if (k->inv_cols_first) {
            pass_inv(k, ptr, sx, sy, 0, odx, idy, osy);
            pass_inv(k, ptr, sx, sy, 1, ody, idx, osx);
        } else {
            if (!k->guard || odx > 1) pass_inv(k, ptr, sx, sy, 1, ody, idx, osx);
            if (!k->guard || ody > 1) pass_inv(k, ptr, sx, sy, 0, odx, idy, osy);
        }
        if (zero_padding) { /* libdwt.c:17156-17176 */
            for (int y = 0; y < ody; y++) zero_run(AT(ptr, y, idx), sy, odx - idx, k->esz);
            for (int x = 0; x < odx; x++) zero_run(AT(ptr, idy, x), sx, ody - idy, k->esz);
        }
    }
}

/* exported entry points: same argument meaning as the reference prototypes (src/libdwt.h:526-992)
 * except that strides are 64-bit */
#define EXPORT_2D(NAME, K)                                                                              \
    void orc_##NAME##_2f(void *ptr, int64_t sx, int64_t sy, int ox, int oy, int ix, int iy, int *j_max, \
                         int decompose_one, int zero_padding)                                           \
    { drive_fwd(&K, ptr, (ptrdiff_t)sx, (ptrdiff_t)sy, ox, oy, ix, iy, j_max, decompose_one, zero_padding); } \
    void orc_##NAME##_2i(void *ptr, int64_t sx, int64_t sy, int ox, int oy, int ix, int iy, int j_max,  \
                         int decompose_one, int zero_padding)                                           \
    { drive_inv(&K, ptr, (ptrdiff_t)sx, (ptrdiff_t)sy, ox, oy, ix, iy, j_max, decompose_one, zero_padding); }
EXPORT_2D(cdf97_s, K97S)
EXPORT_2D(cdf97_d, K97D)
EXPORT_2D(cdf53_i, K53I)
EXPORT_2D(cdf53_s, K53S)
EXPORT_2D(cdf53_d, K53D)
EXPORT_2D(cdf97_i, K97I)

/* =====================================================================================
 * Interleaved in-place family (src/libdwt.c:12926 dwt_cdf97_2f_inplace_s and its _sep / _sdl twins :13485,
 * :13641, :14847 -- all four bit-identical on every input tried; inverse :17474).  Coefficients stay where
 * the lifting leaves them: level j works on the samples at stride 2^j, even = L, odd = H.
 *
 * The reference runs every level as up to eight sweeps -- "exceptions" over rows then columns (lines of
 * 2..4 / 2..3 samples, the whole 1-D transform at once, :10803, :11574), then prolog, core and epilog, each over
 * all rows and then over all columns (:12975-13451, :17512-17598).  In one dimension the three parts add up to
 * the textbook lifting; in two dimensions the top rows (column prolog before the row core) and the right
 * columns (row epilog after the column core) get their row and column steps in a different order than
 * rows-then-columns, which changes the rounding there.  Restated as a table: every lifting operation
 * (step, position) of a line belongs to exactly one part (:9591 prolog, :9929 epilog, the rest core), and a sweep
 * applies the operations of its part step by step.  A line of one sample is not touched at all (:12975, :12997).
 * ===================================================================================== */
enum { PH_X = 0, PH_P = 1, PH_C = 2, PH_E = 3 };
/* forward steps: 0 alpha (odd), 1 beta (even), 2 gamma (odd), 3 delta (even), 4 scale; offset 1 (:10831) */
static int ip_phase_fwd(int N, int step, int i)
{
    if (N < 5) return PH_X;
    if ((step == 0 && (i == 1 || i == 3)) || (step == 1 && (i == 0 || i == 2)) || (step == 2 && i == 1) ||
        (step == 3 && i == 0) || (step == 4 && i == 0))
        return PH_P;
    if (N & 1) {
        if ((step == 1 && i == N - 1) || (step == 2 && i == N - 2) || (step == 3 && (i == N - 1 || i == N - 3)) ||
            (step == 4 && i >= N - 4))
            return PH_E;
    } else {
        if ((step == 0 && i == N - 1) || (step == 1 && i == N - 2) || (step == 2 && (i == N - 1 || i == N - 3)) ||
            (step == 3 && (i == N - 2 || i == N - 4)) || (step == 4 && i >= N - 5))
            return PH_E;
    }
    return PH_C;
}
/* inverse steps: 0 scale, 1 even += -u2, 2 odd += p2, 3 even += -u1, 4 odd += p1; offset 0 (:11604) */
static int ip_phase_inv(int N, int step, int i)
{
    if (N < 4) return PH_X;
    if ((step == 0 && i <= 3) || (step == 1 && (i == 0 || i == 2)) || (step == 2 && i == 1) || (step == 3 && i == 0))
        return PH_P;
    if (N & 1) {
        if ((step == 0 && i == N - 1) || (step == 1 && i == N - 1) || (step == 2 && i == N - 2) ||
            (step == 3 && (i == N - 1 || i == N - 3)) || (step == 4 && (i == N - 2 || i == N - 4)))
            return PH_E;
    } else {
        if ((step == 2 && i == N - 1) || (step == 3 && i == N - 2) || (step == 4 && (i == N - 1 || i == N - 3)))
            return PH_E;
    }
    return PH_C;
}
#define IP(i) (*(float *)(line + (ptrdiff_t)(i) * st))
static void ip_lift(char *line, ptrdiff_t st, int N, int i, float c)
{
    if (i == 0) IP(0) += (2 * c) * IP(1);
    else if (i == N - 1) IP(N - 1) += (2 * c) * IP(N - 2);
    else IP(i) += c * (IP(i - 1) + IP(i + 1));
}
static void ip_sweep(char *line, ptrdiff_t st, int N, int inverse, int part)
{
    const float z = S1s, iz = 1 / z;
    static const float cf[4] = {-P1s, U1s, -P2s, U2s}, ci[4] = {-U2s, P2s, -U1s, P1s};
    for (int step = 0; step < 5; step++) {
        const int scale = inverse ? step == 0 : step == 4;
        const int lift = inverse ? step - 1 : step;            /* index into cf / ci */
        const int par = scale ? -1 : (inverse ? !(lift & 1) ? 0 : 1 : !(lift & 1) ? 1 : 0);
        for (int i = 0; i < N; i++) {
            if (par >= 0 && (i & 1) != par) continue;
            if ((inverse ? ip_phase_inv(N, step, i) : ip_phase_fwd(N, step, i)) != part) continue;
            if (scale) IP(i) *= ((i & 1) != inverse) ? iz : z;   /* fwd: even*z odd*iz; inv: even*iz odd*z */
            else ip_lift(line, st, N, i, inverse ? ci[lift] : cf[lift]);
        }
    }
}
#undef IP
static void ip_level(char *ptr, ptrdiff_t sxj, ptrdiff_t syj, int nx, int ny, int inverse)
{
    for (int part = PH_X; part <= PH_E; part++) {
        if (nx > 1)
            for (int y = 0; y < ny; y++) ip_sweep(ptr + (ptrdiff_t)y * sxj, syj, nx, inverse, part);
        if (ny > 1)
            for (int x = 0; x < nx; x++) ip_sweep(ptr + (ptrdiff_t)x * syj, sxj, ny, inverse, part);
    }
}
/* sx = bytes between rows, sy = bytes between samples of a row (the reference's stride_x / stride_y) */
void orc_cdf97_s_2f_inplace(void *ptr, int64_t sx, int64_t sy, int ox, int oy, int ix, int iy, int *j_max_ptr,
                            int decompose_one)
{
    const int omin = ox < oy ? ox : oy, omax = ox < oy ? oy : ox;
    const int j_limit = orc_ceil_log2(decompose_one ? omax : omin);
    if (*j_max_ptr < 0 || *j_max_ptr > j_limit) *j_max_ptr = j_limit;
    for (int j = 0; j < *j_max_ptr; j++)
        ip_level((char *)ptr, (ptrdiff_t)sx << j, (ptrdiff_t)sy << j, orc_ceil_div_pow2(ix, j), orc_ceil_div_pow2(iy, j), 0);
}
void orc_cdf97_s_2i_inplace(void *ptr, int64_t sx, int64_t sy, int ox, int oy, int ix, int iy, int j_max,
                            int decompose_one)
{
    const int omin = ox < oy ? ox : oy, omax = ox < oy ? oy : ox;
    int j = orc_ceil_log2(decompose_one ? omax : omin);
    if (j_max >= 0 && j_max < j) j = j_max;
    for (; j > 0; j--)
        ip_level((char *)ptr, (ptrdiff_t)sx << (j - 1), (ptrdiff_t)sy << (j - 1), orc_ceil_div_pow2(ix, j - 1),
                 orc_ceil_div_pow2(iy, j - 1), 1);
}

/* CDF 5/3 float of the same family (src/libdwt.c:16553, 17886; lines :11032, :11831): every level transforms all rows,
 * then all columns, in place at stride 2^j -- the textbook lifting of fwd53_s / inv53_s; a line of one sample is scaled. */
static void ip53_lines(char *ptr, ptrdiff_t line_step, int nlines, ptrdiff_t st, int N, int inverse)
{
    float *tmp = (float *)malloc((size_t)(N > 2 ? N : 2) * sizeof(float));
    for (int l = 0; l < nlines; l++) {
        char *line = ptr + (ptrdiff_t)l * line_step;
        if (N == 1) {
            if (inverse) one_inv53_s(line); else one_fwd53_s(line);
            continue;
        }
        gather(tmp, 4, line, st, N, 4);
        if (inverse) inv53_s(tmp, N); else fwd53_s(tmp, N);
        gather(line, st, tmp, 4, N, 4);
    }
    free(tmp);
}
void orc_cdf53_s_2f_inplace(void *ptr, int64_t sx, int64_t sy, int ox, int oy, int ix, int iy, int *j_max_ptr,
                            int decompose_one)
{
    const int omin = ox < oy ? ox : oy, omax = ox < oy ? oy : ox;
    const int j_limit = orc_ceil_log2(decompose_one ? omax : omin);
    if (*j_max_ptr < 0 || *j_max_ptr > j_limit) *j_max_ptr = j_limit;
    for (int j = 0; j < *j_max_ptr; j++) {
        const int nx = orc_ceil_div_pow2(ix, j), ny = orc_ceil_div_pow2(iy, j);
        ip53_lines((char *)ptr, (ptrdiff_t)sx << j, ny, (ptrdiff_t)sy << j, nx, 0);
        ip53_lines((char *)ptr, (ptrdiff_t)sy << j, nx, (ptrdiff_t)sx << j, ny, 0);
    }
}
void orc_cdf53_s_2i_inplace(void *ptr, int64_t sx, int64_t sy, int ox, int oy, int ix, int iy, int j_max,
                            int decompose_one)
{
    const int omin = ox < oy ? ox : oy, omax = ox < oy ? oy : ox;
    int j = orc_ceil_log2(decompose_one ? omax : omin);
    if (j_max >= 0 && j_max < j) j = j_max;
    for (; j > 0; j--) {
        const int nx = orc_ceil_div_pow2(ix, j - 1), ny = orc_ceil_div_pow2(iy, j - 1);
        ip53_lines((char *)ptr, (ptrdiff_t)sx << (j - 1), ny, (ptrdiff_t)sy << (j - 1), nx, 1);
        ip53_lines((char *)ptr, (ptrdiff_t)sy << (j - 1), nx, (ptrdiff_t)sx << (j - 1), ny, 1);
    }
}

/* =====================================================================================
 * Test patterns (src/libdwt.c:1112-1244).  wrap32 != 0 reproduces the reference's 32-bit
 * `int` products (two's-complement wrap, what the compiled reference does); wrap32 == 0 is
 * the 64-bit formula used where the reference would be undefined (65536^2).
 * ===================================================================================== */
static int64_t w32(int64_t v, int wrap32) { return wrap32 ? (int64_t)(int32_t)(uint32_t)(uint64_t)v : v; }

static float pat_s(int x, int y, int rnd, int type, int wrap32)
{
    x++; y++;
    switch (type) {
    case 0: {
        x >>= rnd;
        const int64_t num = w32(w32((int64_t)2 * x, wrap32) * y, wrap32);
        const int64_t den = w32(w32(w32((int64_t)x * x, wrap32) + w32((int64_t)y * y, wrap32), wrap32) + 1, wrap32);
        return (float)num / (float)den;
    }
    case 2: return (float)((x ^ y) & 0xff) / 32;
    case 3: return (float)((((x & 1) << 1) | (y & 1)) + 1) / 4.f;
    default: return 0.f;
    }
}
static double pat_d(int x, int y, int rnd, int wrap32)
{
    x >>= rnd;
    const int64_t num = w32(w32((int64_t)2 * x, wrap32) * y, wrap32);
    const int64_t den = w32(w32(w32((int64_t)x * x, wrap32) + w32((int64_t)y * y, wrap32), wrap32) + 1, wrap32);
    return (double)num / (double)den;
}
static int32_t pat_i(int x, int y, int rnd, int type, int wrap32)
{
    if (type == 2) return (x ^ y) & 0xff;
    x >>= rnd;
    const int64_t num = w32(255 * w32(w32((int64_t)2 * x, wrap32) * y, wrap32), wrap32);
    const int64_t den = w32(w32(w32((int64_t)x * x, wrap32) + w32((int64_t)y * y, wrap32), wrap32) + 1, wrap32);
    return den ? (int32_t)(num / den) : 0;
}
void orc_fill_s(void *ptr, int64_t sx, int64_t sy, int nx, int ny, int rnd, int type, int wrap32)
{
#pragma omp parallel for schedule(static)
    for (int y = 0; y < ny; y++)
        for (int x = 0; x < nx; x++) *(float *)AT(ptr, y, x) = pat_s(x, y, rnd, type, wrap32);
}
void orc_fill_d(void *ptr, int64_t sx, int64_t sy, int nx, int ny, int rnd, int wrap32)
{
#pragma omp parallel for schedule(static)
    for (int y = 0; y < ny; y++)
        for (int x = 0; x < nx; x++) *(double *)AT(ptr, y, x) = pat_d(x, y, rnd, wrap32);
}
void orc_fill_i(void *ptr, int64_t sx, int64_t sy, int nx, int ny, int rnd, int type, int wrap32)
{
#pragma omp parallel for schedule(static)
    for (int y = 0; y < ny; y++)
        for (int x = 0; x < nx; x++) *(int32_t *)AT(ptr, y, x) = pat_i(x, y, rnd, type, wrap32);
}
/* src/volume.c:41: slice z = 2-D type-0 pattern with rand = fold(z & 11) */
void orc_volume_fill_s(void *data, int64_t sx_pix, int64_t sy_row, int64_t sz_slice, int nx, int ny, int nz)
{
    for (int z = 0; z < nz; z++) {
        int rnd = z & 11;
        if (rnd > 11 / 2) rnd = 11 - rnd;
        orc_fill_s((char *)data + (ptrdiff_t)z * sz_slice, sy_row, sx_pix, nx, ny, rnd, 0, 1);
    }
}

/* =====================================================================================
 * 3-D, one level, interleaved subbands, in place on dst after copying src
 * (src/volume-dwt.c:727; 1-D kernel src/dwt-simple.c:2166 = same lifting as fwd97_s).
 * volume_t naming: stride_x = pixel, stride_y = row, stride_z = slice (src/volume.h:14).
 * The reference requires every size >= 5 (dwt-simple.c:2172).
 * ===================================================================================== */
static void lines3(char *base, int n0, ptrdiff_t s0, int n1, ptrdiff_t s1, int N, ptrdiff_t step,
                   void (*fn)(void *, int))
{
#pragma omp parallel
    {
        float *tmp = (float *)malloc((size_t)(N > 2 ? N : 2) * sizeof(float));
#pragma omp for schedule(static) collapse(2)
        for (int a = 0; a < n0; a++)
            for (int b = 0; b < n1; b++) {
                char *p = base + (ptrdiff_t)a * s0 + (ptrdiff_t)b * s1;
                gather(tmp, 4, p, step, N, 4);
                fn(tmp, N);
                gather(p, step, tmp, 4, N, 4);
            }
        free(tmp);
    }
}
void orc_cdf97_3f_s(const void *src, int64_t ssx, int64_t ssy, int64_t ssz, void *dst, int64_t dsx, int64_t dsy,
                    int64_t dsz, int nx, int ny, int nz)
{
    if (src != dst)
        for (int z = 0; z < nz; z++)
            for (int y = 0; y < ny; y++)
                gather((char *)dst + (ptrdiff_t)y * dsy + (ptrdiff_t)z * dsz, (ptrdiff_t)dsx,
                       (const char *)src + (ptrdiff_t)y * ssy + (ptrdiff_t)z * ssz, (ptrdiff_t)ssx, nx, 4);
    lines3((char *)dst, ny, dsy, nz, dsz, nx, dsx, fwd97_s);
    lines3((char *)dst, nx, dsx, nz, dsz, ny, dsy, fwd97_s);
    lines3((char *)dst, nx, dsx, ny, dsy, nz, dsz, fwd97_s);
}
/* src/volume-dwt.c:1115 -> dwt_cdf97_1i_inplace_s(ptr, stride, size, 1): x, then y, then z */
void orc_cdf97_3i_s(void *vol, int64_t sx, int64_t sy, int64_t sz, int nx, int ny, int nz)
{
    if (nx > 1) lines3((char *)vol, ny, sy, nz, sz, nx, sx, inv97_s);
    if (ny > 1) lines3((char *)vol, nx, sx, nz, sz, ny, sy, inv97_s);
    if (nz > 1) lines3((char *)vol, nx, sx, ny, sy, nz, sz, inv97_s);
}

void orc_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
int orc_get_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
