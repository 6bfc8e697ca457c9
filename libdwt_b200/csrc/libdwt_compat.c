/*
 * libdwt_compat.c -- the reference's hot-path entry points (C99), forwarding to libdwtb200.so.
 * See include/libdwt_compat.h for the list and the reference lines each one replaces.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/dwtb200.h"
#include "../../include/libdwt_compat.h"

/* the reference's error convention: log + abort (src/libdwt.c:20410-20421, 19200-19215) */
static void die(const char *what, int rc)
{
    fprintf(stderr, "ERROR: %s failed (%d): %s\n", what, rc, dwtb200_last_error());
    abort();
}

#define FWD(NAME, KIND)                                                                                          \
    void NAME(void *ptr, int stride_x, int stride_y, int size_o_big_x, int size_o_big_y, int size_i_big_x,      \
              int size_i_big_y, int *j_max_ptr, int decompose_one, int zero_padding)                            \
    {                                                                                                            \
        const int rc = dwtb200_fwd2_host(KIND, ptr, stride_x, stride_y, size_o_big_x, size_o_big_y, size_i_big_x, \
                                         size_i_big_y, j_max_ptr, decompose_one, zero_padding);                 \
        if (rc) die(#NAME, rc);                                                                                  \
    }
#define INV(NAME, KIND)                                                                                          \
    void NAME(void *ptr, int stride_x, int stride_y, int size_o_big_x, int size_o_big_y, int size_i_big_x,      \
              int size_i_big_y, int j_max, int decompose_one, int zero_padding)                                 \
    {                                                                                                            \
        const int rc = dwtb200_inv2_host(KIND, ptr, stride_x, stride_y, size_o_big_x, size_o_big_y, size_i_big_x, \
                                         size_i_big_y, j_max, decompose_one, zero_padding);                     \
        if (rc) die(#NAME, rc);                                                                                  \
    }

FWD(dwt_cdf97_2f_s, DWTB200_CDF97_F32)
INV(dwt_cdf97_2i_s, DWTB200_CDF97_F32)
FWD(dwt_cdf97_2f_d, DWTB200_CDF97_F64)
INV(dwt_cdf97_2i_d, DWTB200_CDF97_F64)
FWD(dwt_cdf53_2f_i, DWTB200_CDF53_I32)
INV(dwt_cdf53_2i_i, DWTB200_CDF53_I32)
/* sibling drivers on the same kernels */
FWD(dwt_cdf53_2f_s, DWTB200_CDF53_F32)
INV(dwt_cdf53_2i_s, DWTB200_CDF53_F32)
FWD(dwt_cdf53_2f_d, DWTB200_CDF53_F64)
INV(dwt_cdf53_2i_d, DWTB200_CDF53_F64)
FWD(dwt_cdf97_2f_i, DWTB200_CDF97_I32)
INV(dwt_cdf97_2i_i, DWTB200_CDF97_I32)

/* interleaved in-place family: zero_padding is ignored, as in the reference */
#define FWD_IP(NAME, KIND)                                                                                       \
    void NAME(void *ptr, int stride_x, int stride_y, int size_o_big_x, int size_o_big_y, int size_i_big_x,      \
              int size_i_big_y, int *j_max_ptr, int decompose_one, int zero_padding)                            \
    {                                                                                                            \
        (void)zero_padding;                                                                                      \
        const int rc = dwtb200_fwd2_inplace_host(KIND, ptr, stride_x, stride_y, size_o_big_x, size_o_big_y,      \
                                                 size_i_big_x, size_i_big_y, j_max_ptr, decompose_one);          \
        if (rc) die(#NAME, rc);                                                                                  \
    }
#define INV_IP(NAME, KIND)                                                                                       \
    void NAME(void *ptr, int stride_x, int stride_y, int size_o_big_x, int size_o_big_y, int size_i_big_x,      \
              int size_i_big_y, int j_max, int decompose_one, int zero_padding)                                 \
    {                                                                                                            \
        (void)zero_padding;                                                                                      \
        const int rc = dwtb200_inv2_inplace_host(KIND, ptr, stride_x, stride_y, size_o_big_x, size_o_big_y,      \
                                                 size_i_big_x, size_i_big_y, j_max, decompose_one);              \
        if (rc) die(#NAME, rc);                                                                                  \
    }
FWD_IP(dwt_cdf97_2f_inplace_s, DWTB200_CDF97_F32)
FWD_IP(dwt_cdf97_2f_inplace_sep_s, DWTB200_CDF97_F32)
FWD_IP(dwt_cdf97_2f_inplace_sdl_s, DWTB200_CDF97_F32)
FWD_IP(dwt_cdf97_2f_inplace_sep_sdl_s, DWTB200_CDF97_F32)
INV_IP(dwt_cdf97_2i_inplace_s, DWTB200_CDF97_F32)
FWD_IP(dwt_cdf53_2f_inplace_s, DWTB200_CDF53_F32)
INV_IP(dwt_cdf53_2i_inplace_s, DWTB200_CDF53_F32)

#define PERF_IP(NAME)                                                                                                       \
    void NAME(int stride_x, int stride_y, int size_o_big_x, int size_o_big_y, int size_i_big_x, int size_i_big_y, int j_max,    \
              int decompose_one, int zero_padding, int M, int N, int clock_type, float *fwd_secs, float *inv_secs)              \
    {                                                                                                                           \
        (void)stride_x; (void)stride_y; (void)zero_padding; (void)clock_type;                                                   \
        const int rc = dwtb200_perf2_inplace(DWTB200_CDF97_F32, size_o_big_x, size_o_big_y, size_i_big_x, size_i_big_y, j_max,  \
                                             decompose_one, M, N, fwd_secs, inv_secs);                                          \
        if (rc) die(#NAME, rc);                                                                                                 \
    }
PERF_IP(dwt_util_perf_cdf97_2_inplace_s)
PERF_IP(dwt_util_perf_cdf97_2_inplace_sep_s)
PERF_IP(dwt_util_perf_cdf97_2_inplace_sdl_s)
PERF_IP(dwt_util_perf_cdf97_2_inplace_sep_sdl_s)

/* dwt_util_measure_perf_cdf97_2_s / _inplace_s and twins (src/libdwt.c:22559, 22646): the perf harness over a range of square sizes,
 * x growing by g_growth_factor_s = 1.13 (:22385), sizes per dwt_util_get_sizes_s (:22296: outer size rounded up to a power of two for
 * DWT_ARR_SIMPLE / DWT_ARR_SPARSE), one "pixels <TAB> seconds" line per size and direction.  The reference's versions call its own
 * CPU harness from inside libdwt.a, so they are restated here on top of the device harness. */
static int pow2_not_less(int x)
{
    int p = 1;
    while (p < x) p <<= 1;
    return p;
}
static void measure_perf(int inplace, int array_type, int min_x, int max_x, int j_max, int decompose_one, int zero_padding, int M, int N,
                         FILE *fwd_plot_data, FILE *inv_plot_data, const char *name)
{
    for (int x = min_x; x <= max_x;) {
        const int o = (array_type == 0 || array_type == 1) ? pow2_not_less(x) : x;   /* DWT_ARR_SIMPLE, DWT_ARR_SPARSE */
        float f = 0, i = 0;
        const int rc = inplace ? dwtb200_perf2_inplace(DWTB200_CDF97_F32, o, o, x, x, j_max, decompose_one, M, N, &f, &i)
                               : dwtb200_perf2(DWTB200_CDF97_F32, o, o, x, x, j_max, decompose_one, zero_padding, M, N, &f, &i);
        if (rc) die(name, rc);
        fprintf(fwd_plot_data, "%i\t%.10f\n", x * x, f);
        fprintf(inv_plot_data, "%i\t%.10f\n", x * x, i);
        const float t = x * 1.13f;   /* x = ceilf(x * growth_factor) */
        int nx = (int)t;
        if ((float)nx < t) nx++;
        x = nx;
    }
}
#define MEASURE(NAME, INPLACE)                                                                                                 \
    void NAME(int array_type, int min_x, int max_x, int opt_stride, int j_max, int decompose_one, int zero_padding, int M, int N, \
              int clock_type, FILE *fwd_plot_data, FILE *inv_plot_data)                                                        \
    {                                                                                                                          \
        (void)opt_stride; (void)clock_type;                                                                                    \
        measure_perf(INPLACE, array_type, min_x, max_x, j_max, decompose_one, zero_padding, M, N, fwd_plot_data, inv_plot_data, #NAME); \
    }
MEASURE(dwt_util_measure_perf_cdf97_2_s, 0)
MEASURE(dwt_util_measure_perf_cdf97_2_inplace_s, 1)
MEASURE(dwt_util_measure_perf_cdf97_2_inplace_sep_s, 1)
MEASURE(dwt_util_measure_perf_cdf97_2_inplace_sdl_s, 1)
MEASURE(dwt_util_measure_perf_cdf97_2_inplace_sep_sdl_s, 1)

void dwt_cdf97_2f_s2(const void *src, void *dst, int stride_x, int stride_y, int size_o_big_x, int size_o_big_y, int size_i_big_x,
                     int size_i_big_y, int *j_max_ptr, int decompose_one, int zero_padding)
{
    const int rc = dwtb200_fwd2_host2(DWTB200_CDF97_F32, src, dst, stride_x, stride_y, size_o_big_x, size_o_big_y, size_i_big_x,
                                      size_i_big_y, j_max_ptr, decompose_one, zero_padding);
    if (rc) die("dwt_cdf97_2f_s2", rc);
}
void dwt_cdf97_2i_s2(const void *src, void *dst, int stride_x, int stride_y, int size_o_big_x, int size_o_big_y, int size_i_big_x,
                     int size_i_big_y, int j_max, int decompose_one, int zero_padding)
{
    const int rc = dwtb200_inv2_host2(DWTB200_CDF97_F32, src, dst, stride_x, stride_y, size_o_big_x, size_o_big_y, size_i_big_x,
                                      size_i_big_y, j_max, decompose_one, zero_padding);
    if (rc) die("dwt_cdf97_2i_s2", rc);
}

/* the reference's perf harness, timed on the device (strides and clock type have no meaning for HBM-resident images) */
void dwt_util_perf_cdf97_2_s(int stride_x, int stride_y, int size_o_big_x, int size_o_big_y, int size_i_big_x, int size_i_big_y,
                             int j_max, int decompose_one, int zero_padding, int M, int N, int clock_type, float *fwd_secs,
                             float *inv_secs)
{
    (void)stride_x; (void)stride_y; (void)clock_type;
    const int rc = dwtb200_perf2(DWTB200_CDF97_F32, size_o_big_x, size_o_big_y, size_i_big_x, size_i_big_y, j_max, decompose_one,
                                 zero_padding, M, N, fwd_secs, inv_secs);
    if (rc) die("dwt_util_perf_cdf97_2_s", rc);
}
void dwt_util_perf_cdf53_2_i(int stride_x, int stride_y, int size_o_big_x, int size_o_big_y, int size_i_big_x, int size_i_big_y,
                             int j_max, int decompose_one, int zero_padding, int M, int N, int clock_type, float *fwd_secs,
                             float *inv_secs)
{
    (void)stride_x; (void)stride_y; (void)clock_type;
    const int rc = dwtb200_perf2(DWTB200_CDF53_I32, size_o_big_x, size_o_big_y, size_i_big_x, size_i_big_y, j_max, decompose_one,
                                 zero_padding, M, N, fwd_secs, inv_secs);
    if (rc) die("dwt_util_perf_cdf53_2_i", rc);
}

void dwt_util_alloc_image(void **pptr, int stride_x, int stride_y, int size_o_big_x, int size_o_big_y)
{
    (void)stride_y;
    (void)size_o_big_x;
    /* the reference sizes the block as stride_x * size_y in `int` (src/libdwt.c:1419); 64-bit here */
    *pptr = dwtb200_host_alloc((size_t)stride_x * (size_t)size_o_big_y);
    if (!*pptr) die("dwt_util_alloc_image", DWTB200_ENOMEM);
}

void dwt_util_free_image(void **pptr)
{
    dwtb200_host_free(*pptr);
    *pptr = NULL;
}

void cdf97_3f_op_sep_horizontal_s(struct volume_t *src, struct volume_t *dst)
{
    const int rc = dwtb200_fwd3_host(src->data, src->stride_x, src->stride_y, src->stride_z, dst->data, dst->stride_x,
                                     dst->stride_y, dst->stride_z, src->size_x, src->size_y, src->size_z);
    if (rc) die("cdf97_3f_op_sep_horizontal_s", rc);
}
void cdf97_3f_ip_sep_horizontal_s(struct volume_t *v)
{
    cdf97_3f_op_sep_horizontal_s(v, v);
}
void cdf97_3i_ip_sep_horizontal_s(struct volume_t *v)
{
    const int rc = dwtb200_inv3_host(v->data, v->stride_x, v->stride_y, v->stride_z, v->size_x, v->size_y, v->size_z);
    if (rc) die("cdf97_3i_ip_sep_horizontal_s", rc);
}

/* src/volume-dwt.h:227.  Measured on the compiled reference: VOL_SEP_HORIZONTAL (0) and VOL_SEP_VERTICAL (1) give identical bits; the
 * blocked / fused schedules 2..9 round differently in most samples and overrun their buffers on sizes that are not multiples of their
 * block, 10..12 are single-axis passes.  Only the two parity-checked ones are served. */
void cdf97_3f_op_wrapper_s(struct volume_t *src, struct volume_t *dst, int approach)
{
    if (approach != 0 && approach != 1) {
        fprintf(stderr, "ERROR: cdf97_3f_op_wrapper_s: approach %d is a CPU loop schedule with its own rounding; only VOL_SEP_HORIZONTAL / VOL_SEP_VERTICAL run on the device\n", approach);
        abort();
    }
    cdf97_3f_op_sep_horizontal_s(src, dst);
}

/* src/volume-dwt.c:2810: every `approach` is a CPU schedule of the same transform; the device has one */
int volume_perftest_fwd97op_s(int size, int opt_stride, int approach, int N, double *secs, long unsigned *faults)
{
    (void)opt_stride;
    (void)approach;
    int errors = 0;
    const int rc = dwtb200_perf3(size, N, secs, &errors);
    if (rc) die("volume_perftest_fwd97op_s", rc);
    if (faults) *faults = 0;
    return errors;
}

/* ---- volumes in page-locked host memory: volume_alloc_realiably / volume_alloc_realiably_locked / volume_free
 * (src/volume.c:10, 194, 34).  Strides per dwt_util_get_stride (src/libdwt.c:20688-20728): 0 packed, 1 next prime, 2 prime << 6,
 * 3 one cache line past a page multiple, 4 multiple of 64, 5, 6 odd, 7 odd << 6.  The reference mlock()s the block; here it is
 * cudaHostAlloc'ed, so the slice-by-slice transfers of cdf97_3f_op_sep_horizontal_s run at PCIe speed instead of through a staging copy. */
static int is_prime_(int n)
{
    if (n < 2) return 0;
    if (n % 2 == 0) return n == 2;
    for (int q = 3; (long long)q * q <= n; q += 2)
        if (n % q == 0) return 0;
    return 1;
}
static int next_prime_(int n)
{
    if (n <= 2) return 2;
    n |= 1;
    while (!is_prime_(n)) n += 2;
    return n;
}
static int ceil_log2_(int x)
{
    int j = 0;
    while ((1LL << j) < x) j++;
    return j;
}
static int get_stride_(int min_stride, int opt)
{
    const int a64 = (min_stride + 63) & ~63, a4096 = (min_stride + 4095) & ~4095;
    switch (opt) {
    case 1: return next_prime_(min_stride);
    case 2: return next_prime_(a64 >> 6) << 6;
    case 3: return ((a4096 >> 6) + 1) << 6;
    case 4: return a64;
    case 5: return a4096 + (1 << ceil_log2_(min_stride));
    case 6: return min_stride | 1;
    case 7: return ((a64 >> 6) | 1) << 6;
    default: return min_stride;
    }
}
struct volume_t *volume_alloc_realiably_locked(size_t pix_size, int size_x, int size_y, int size_z, int opt_stride)
{
    struct volume_t *v = (struct volume_t *)malloc(sizeof *v);
    if (!v) die("volume_alloc_realiably_locked", DWTB200_ENOMEM);
    v->size_x = size_x;
    v->size_y = size_y;
    v->size_z = size_z;
    v->stride_x = pix_size;
    v->stride_y = (size_t)get_stride_((int)(v->stride_x * (size_t)size_x), opt_stride);
    v->stride_z = (size_t)get_stride_((int)(v->stride_y * (size_t)size_y), opt_stride);
    v->data = dwtb200_host_alloc(v->stride_z * (size_t)size_z);
    if (!v->data) die("volume_alloc_realiably_locked", DWTB200_ENOMEM);
    return v;
}
struct volume_t *volume_alloc_realiably(size_t pix_size, int size_x, int size_y, int size_z, int opt_stride)
{
    return volume_alloc_realiably_locked(pix_size, size_x, size_y, size_z, opt_stride);
}
void volume_free(struct volume_t *volume)
{
    if (!volume) return;
    dwtb200_host_free(volume->data);
    free(volume);
}

/* volume_measure_fwd97op_s (src/volume-dwt.c:2898): volume_perftest_fwd97op_s over cube edges size_min, size_grow(size) ... below
 * size_max (growth factor 1.13, rounded up to a multiple of size_step: :2884), one "voxels <TAB> seconds per voxel" line per size into
 * data/perftest/time-stride=S-approach=A.txt and the page-fault twin (always 0 here: the volumes live in HBM) */
int volume_measure_fwd97op_s(int size_min, int size_max, int size_step, int N, int opt_stride, int approach)
{
    char path[4096];
    snprintf(path, sizeof path, "data/perftest/time-stride=%i-approach=%i.txt", opt_stride, approach);
    FILE *file_time = fopen(path, "w");
    if (!file_time) {
        fprintf(stderr, "ERROR: unable to open file: %s\n", path);
        abort();
    }
    snprintf(path, sizeof path, "data/perftest/faults-stride=%i-approach=%i.txt", opt_stride, approach);
    FILE *file_faults = fopen(path, "w");
    if (!file_faults) {
        fprintf(stderr, "ERROR: unable to open file: %s\n", path);
        abort();
    }
    fprintf(file_time, "# voxels secs/pel\n");
    fprintf(file_faults, "# voxels page_faults\n");
    int total_errors = 0;
    for (int size = size_min; size < size_max;) {
        double secs = 0;
        long unsigned faults = 0;
        const int errors = volume_perftest_fwd97op_s(size, opt_stride, approach, N, &secs, &faults);
        const int voxels = size * size * size;
        fprintf(stderr, "perftest: size=%4i opt_stride=%i approach=%2i (N=%2i): time=%f [nsecs/pel]; errors=%i; faults=%lu\n", size, opt_stride,
                approach, N, secs * 1e9, errors, faults);
        fprintf(file_time, "%i\t%.20f\n", voxels, secs);
        fprintf(file_faults, "%i\t%lu\n", voxels, faults);
        total_errors += errors;
        size = (int)(size * 1.13f);   /* size_grow */
        size += 1;
        size += size_step - 1;
        size &= ~(size_step - 1);
    }
    fclose(file_time);
    fclose(file_faults);
    return total_errors;
}

double dwt_b200_last_transform_ms(void) { return dwtb200_last_transform_ms(); }
