/*
 * libdwt_compat.h -- the reference's own symbol names for the hot path, implemented on the B200 by
 * libdwt_compat.so (libdwt_b200/csrc/libdwt_compat.c, plain C99) on top of the C ABI in dwtb200.h.
 *
 * Every prototype is the reference's, verbatim in meaning and argument order (paths relative to the
 * libdwt tree, /root/reference/): existing programs keep including the reference's libdwt.h and are
 * linked against libdwt_compat.so ahead of the reference library.
 *
 * Semantics kept: byte strides (possibly unaligned / multi-channel), outer and inner sizes, j_max in/out,
 * decompose_one, zero_padding, symmetric boundary extension, Mallat subband layout, synchronous
 * in-place update of the caller's host buffer, void return, errors -> message on stderr + abort()
 * (src/libdwt.c:20410-20421, 19200-19215).  There is no CPU fallback.
 */
#ifndef LIBDWT_COMPAT_H
#define LIBDWT_COMPAT_H

#include <stddef.h>
#include <stdio.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* src/libdwt.h:562-573, 867-878 (src/libdwt.c:12776, 17040) */
void dwt_cdf97_2f_s(void *ptr, int stride_x, int stride_y, int size_o_big_x, int size_o_big_y, int size_i_big_x,
                    int size_i_big_y, int *j_max_ptr, int decompose_one, int zero_padding);
void dwt_cdf97_2i_s(void *ptr, int stride_x, int stride_y, int size_o_big_x, int size_o_big_y, int size_i_big_x,
                    int size_i_big_y, int j_max, int decompose_one, int zero_padding);
/* src/libdwt.h:526-537, 831-842 (src/libdwt.c:12451, 16884) */
void dwt_cdf97_2f_d(void *ptr, int stride_x, int stride_y, int size_o_big_x, int size_o_big_y, int size_i_big_x,
                    int size_i_big_y, int *j_max_ptr, int decompose_one, int zero_padding);
void dwt_cdf97_2i_d(void *ptr, int stride_x, int stride_y, int size_o_big_x, int size_o_big_y, int size_i_big_x,
                    int size_i_big_y, int j_max, int decompose_one, int zero_padding);
/* src/libdwt.h:686-697, 981-992 (src/libdwt.c:16304, 18142) */
void dwt_cdf53_2f_i(void *ptr, int stride_x, int stride_y, int size_o_big_x, int size_o_big_y, int size_i_big_x,
                    int size_i_big_y, int *j_max_ptr, int decompose_one, int zero_padding);
void dwt_cdf53_2i_i(void *ptr, int stride_x, int stride_y, int size_o_big_x, int size_o_big_y, int size_i_big_x,
                    int size_i_big_y, int j_max, int decompose_one, int zero_padding);

/* sibling drivers sharing the same kernels.  CDF 5/3 float: src/libdwt.h:722-733, 1053-1064 (src/libdwt.c:16470, 18296) */
void dwt_cdf53_2f_s(void *ptr, int stride_x, int stride_y, int size_o_big_x, int size_o_big_y, int size_i_big_x,
                    int size_i_big_y, int *j_max_ptr, int decompose_one, int zero_padding);
void dwt_cdf53_2i_s(void *ptr, int stride_x, int stride_y, int size_o_big_x, int size_o_big_y, int size_i_big_x,
                    int size_i_big_y, int j_max, int decompose_one, int zero_padding);
/* CDF 5/3 double: src/libdwt.h:544-555, 849-860 (src/libdwt.c:12535, 16962) */
void dwt_cdf53_2f_d(void *ptr, int stride_x, int stride_y, int size_o_big_x, int size_o_big_y, int size_i_big_x,
                    int size_i_big_y, int *j_max_ptr, int decompose_one, int zero_padding);
void dwt_cdf53_2i_d(void *ptr, int stride_x, int stride_y, int size_o_big_x, int size_o_big_y, int size_i_big_x,
                    int size_i_big_y, int j_max, int decompose_one, int zero_padding);
/* CDF 9/7 with integer lifting: src/libdwt.h:704-715, 999-1010 (src/libdwt.c:16387, 18219) */
void dwt_cdf97_2f_i(void *ptr, int stride_x, int stride_y, int size_o_big_x, int size_o_big_y, int size_i_big_x,
                    int size_i_big_y, int *j_max_ptr, int decompose_one, int zero_padding);
void dwt_cdf97_2i_i(void *ptr, int stride_x, int stride_y, int size_o_big_x, int size_o_big_y, int size_i_big_x,
                    int size_i_big_y, int j_max, int decompose_one, int zero_padding);

/* interleaved in-place family: src/libdwt.h:586-597, 612-662, 889-900 (src/libdwt.c:12926, 13485, 13641, 14847, 17474) and the
 * 5/3 float pair src/libdwt.h:599-610, 944-955 (src/libdwt.c:16553, 17886).  The four forward 9/7 variants are the same
 * arithmetic in a different loop order and give bit-identical results in the reference. */
void dwt_cdf97_2f_inplace_s(void *ptr, int stride_x, int stride_y, int size_o_big_x, int size_o_big_y, int size_i_big_x,
                            int size_i_big_y, int *j_max_ptr, int decompose_one, int zero_padding);
void dwt_cdf97_2f_inplace_sep_s(void *ptr, int stride_x, int stride_y, int size_o_big_x, int size_o_big_y, int size_i_big_x,
                                int size_i_big_y, int *j_max_ptr, int decompose_one, int zero_padding);
void dwt_cdf97_2f_inplace_sdl_s(void *ptr, int stride_x, int stride_y, int size_o_big_x, int size_o_big_y, int size_i_big_x,
                                int size_i_big_y, int *j_max_ptr, int decompose_one, int zero_padding);
void dwt_cdf97_2f_inplace_sep_sdl_s(void *ptr, int stride_x, int stride_y, int size_o_big_x, int size_o_big_y, int size_i_big_x,
                                    int size_i_big_y, int *j_max_ptr, int decompose_one, int zero_padding);
void dwt_cdf97_2i_inplace_s(void *ptr, int stride_x, int stride_y, int size_o_big_x, int size_o_big_y, int size_i_big_x,
                            int size_i_big_y, int j_max, int decompose_one, int zero_padding);
void dwt_cdf53_2f_inplace_s(void *ptr, int stride_x, int stride_y, int size_o_big_x, int size_o_big_y, int size_i_big_x,
                            int size_i_big_y, int *j_max_ptr, int decompose_one, int zero_padding);
void dwt_cdf53_2i_inplace_s(void *ptr, int stride_x, int stride_y, int size_o_big_x, int size_o_big_y, int size_i_big_x,
                            int size_i_big_y, int j_max, int decompose_one, int zero_padding);

/* out-of-place 9/7 float: src/libdwt.h:667-679, 962-974 (src/libdwt.c:12619, 17985) */
void dwt_cdf97_2f_s2(const void *src, void *dst, int stride_x, int stride_y, int size_o_big_x, int size_o_big_y, int size_i_big_x,
                     int size_i_big_y, int *j_max_ptr, int decompose_one, int zero_padding);
void dwt_cdf97_2i_s2(const void *src, void *dst, int stride_x, int stride_y, int size_o_big_x, int size_o_big_y, int size_i_big_x,
                     int size_i_big_y, int j_max, int decompose_one, int zero_padding);
/* the performance harness re-pointed at the device: src/libdwt.h:2498-2513 (src/libdwt.c:21391, 21262); M images, N loops,
 * minimum of the mean seconds per transform, measured with CUDA events on HBM-resident images */
void dwt_util_perf_cdf97_2_s(int stride_x, int stride_y, int size_o_big_x, int size_o_big_y, int size_i_big_x, int size_i_big_y,
                             int j_max, int decompose_one, int zero_padding, int M, int N, int clock_type, float *fwd_secs,
                             float *inv_secs);
void dwt_util_perf_cdf53_2_i(int stride_x, int stride_y, int size_o_big_x, int size_o_big_y, int size_i_big_x, int size_i_big_y,
                             int j_max, int decompose_one, int zero_padding, int M, int N, int clock_type, float *fwd_secs,
                             float *inv_secs);

/* the same harness for the in-place family: src/libdwt.h:2520-2590 (four loop orders of one transform) */
void dwt_util_perf_cdf97_2_inplace_s(int stride_x, int stride_y, int size_o_big_x, int size_o_big_y, int size_i_big_x, int size_i_big_y,
                                     int j_max, int decompose_one, int zero_padding, int M, int N, int clock_type, float *fwd_secs,
                                     float *inv_secs);
void dwt_util_perf_cdf97_2_inplace_sep_s(int stride_x, int stride_y, int size_o_big_x, int size_o_big_y, int size_i_big_x, int size_i_big_y,
                                         int j_max, int decompose_one, int zero_padding, int M, int N, int clock_type, float *fwd_secs,
                                         float *inv_secs);
void dwt_util_perf_cdf97_2_inplace_sdl_s(int stride_x, int stride_y, int size_o_big_x, int size_o_big_y, int size_i_big_x, int size_i_big_y,
                                         int j_max, int decompose_one, int zero_padding, int M, int N, int clock_type, float *fwd_secs,
                                         float *inv_secs);
void dwt_util_perf_cdf97_2_inplace_sep_sdl_s(int stride_x, int stride_y, int size_o_big_x, int size_o_big_y, int size_i_big_x,
                                             int size_i_big_y, int j_max, int decompose_one, int zero_padding, int M, int N, int clock_type,
                                             float *fwd_secs, float *inv_secs);

/* the harness over a range of sizes, one "pixels <TAB> seconds" line per size: src/libdwt.h:2700-2815 (src/libdwt.c:22559, 22646 ...);
 * array_type is the reference's enum dwt_array (0 DWT_ARR_SIMPLE, 1 DWT_ARR_SPARSE, 2 DWT_ARR_PACKED), passed as int */
void dwt_util_measure_perf_cdf97_2_s(int array_type, int min_x, int max_x, int opt_stride, int j_max, int decompose_one, int zero_padding,
                                     int M, int N, int clock_type, FILE *fwd_plot_data, FILE *inv_plot_data);
void dwt_util_measure_perf_cdf97_2_inplace_s(int array_type, int min_x, int max_x, int opt_stride, int j_max, int decompose_one,
                                             int zero_padding, int M, int N, int clock_type, FILE *fwd_plot_data, FILE *inv_plot_data);
void dwt_util_measure_perf_cdf97_2_inplace_sep_s(int array_type, int min_x, int max_x, int opt_stride, int j_max, int decompose_one,
                                                 int zero_padding, int M, int N, int clock_type, FILE *fwd_plot_data, FILE *inv_plot_data);
void dwt_util_measure_perf_cdf97_2_inplace_sdl_s(int array_type, int min_x, int max_x, int opt_stride, int j_max, int decompose_one,
                                                 int zero_padding, int M, int N, int clock_type, FILE *fwd_plot_data, FILE *inv_plot_data);
void dwt_util_measure_perf_cdf97_2_inplace_sep_sdl_s(int array_type, int min_x, int max_x, int opt_stride, int j_max, int decompose_one,
                                                     int zero_padding, int M, int N, int clock_type, FILE *fwd_plot_data,
                                                     FILE *inv_plot_data);

/* src/libdwt.h:1382-1409 (src/libdwt.c:1437, 1482): page-locked host memory instead of memalign(16, ...) */
void dwt_util_alloc_image(void **pptr, int stride_x, int stride_y, int size_o_big_x, int size_o_big_y);
void dwt_util_free_image(void **pptr);

/* src/volume.h:14-24 */
struct volume_t {
    int size_x, size_y, size_z;
    size_t stride_x, stride_y, stride_z; /* sizeof(pixel), sizeof(row), sizeof(slice) */
    void *data;
};
/* src/volume-dwt.h:21, 38 and the inverse (src/volume-dwt.c:727, 677, 1115) */
void cdf97_3f_op_sep_horizontal_s(struct volume_t *src, struct volume_t *dst);
void cdf97_3f_ip_sep_horizontal_s(struct volume_t *volume);
void cdf97_3i_ip_sep_horizontal_s(struct volume_t *volume);
/* src/volume-dwt.h:227; approach: enum volume_approach, only VOL_SEP_HORIZONTAL (0) / VOL_SEP_VERTICAL (1), whose results are identical */
void cdf97_3f_op_wrapper_s(struct volume_t *src, struct volume_t *dst, int approach);

/* src/volume-dwt.h:234 (src/volume-dwt.c:2810): `approach` is enum volume_approach there (an int here: every approach is a
 * CPU schedule of the same transform); seconds per voxel of the forward transform on a device-resident volume */
int volume_perftest_fwd97op_s(int size, int opt_stride, int approach, int N, double *secs, long unsigned *faults);

/* src/volume.h:29, 81, 34 (src/volume.c:10, 194): strides per dwt_util_get_stride(opt_stride); page-locked host memory, so the
 * transfers of the 3-D entry points above run at PCIe speed */
struct volume_t *volume_alloc_realiably(size_t pix_size, int size_x, int size_y, int size_z, int opt_stride);
struct volume_t *volume_alloc_realiably_locked(size_t pix_size, int size_x, int size_y, int size_z, int opt_stride);
void volume_free(struct volume_t *volume);
/* src/volume-dwt.h:246 (src/volume-dwt.c:2898): the perf test over a range of cube sizes, results into data/perftest/ */
int volume_measure_fwd97op_s(int size_min, int size_max, int size_step, int N, int opt_stride, int approach);

/* device-event timing around the transforms (replaces dwt_util_get_clock in benchmarks, src/libdwt.c:18701):
 * milliseconds the device spent in the last transform call, excluding host<->device copies */
double dwt_b200_last_transform_ms(void);

#ifdef __cplusplus
}
#endif
#endif
