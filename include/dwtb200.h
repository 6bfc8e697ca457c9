/*
 * dwtb200.h -- C ABI of libdwtb200.so, the B200 (sm_100a) implementation of libdwt's separable
 * lifting DWT hot path.  Plain pointers and sizes only; no C++/torch types cross this boundary.
 *
 * Every entry point names the reference interface it replaces (paths relative to the libdwt
 * source tree, /root/reference/).  The reference-named symbols (dwt_cdf97_2f_s, ...) are exported
 * by the thin C99 layer libdwt_b200/csrc/libdwt_compat.c, which calls the functions below; see
 * INTEGRATION.md for how the unmodified examples link against it.
 *
 * Conventions
 *   - "kind" selects wavelet and sample type (DWTB200_CDF97_F32, ...).
 *   - host strides are in BYTES exactly as in libdwt: stride_x between rows, stride_y between
 *     columns (src/libdwt.h:528-529); they may be unaligned (dwt_util_get_opt_stride returns primes).
 *   - functions return 0 on success, a negative DWTB200_E* code otherwise; dwtb200_last_error()
 *     describes the failure.  There is no CPU fallback: without a usable CUDA device every compute
 *     entry point fails with DWTB200_ENODEV.
 *   - threading: the library keeps process-global state like the reference (src/libdwt.c:478-756), but unlike the
 *     reference every entry point below takes one process-wide (recursive) lock, so calls from several host threads
 *     are safe and are serialised; device work of different images still overlaps (each image owns a stream).
 *     One process drives one GPU; use one process per GPU (LOCAL_RANK) for several.
 */
#ifndef DWTB200_H
#define DWTB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    DWTB200_CDF97_F32 = 0, /* dwt_cdf97_2f_s / dwt_cdf97_2i_s   src/libdwt.c:12776, 17040 */
    DWTB200_CDF97_F64 = 1, /* dwt_cdf97_2f_d / dwt_cdf97_2i_d   src/libdwt.c:12451, 16884 */
    DWTB200_CDF53_I32 = 2, /* dwt_cdf53_2f_i / dwt_cdf53_2i_i   src/libdwt.c:16304, 18142 */
    /* sibling drivers sharing the same kernels */
    DWTB200_CDF53_F32 = 3, /* dwt_cdf53_2f_s / dwt_cdf53_2i_s   src/libdwt.c:16470, 18296 */
    DWTB200_CDF53_F64 = 4, /* dwt_cdf53_2f_d / dwt_cdf53_2i_d   src/libdwt.c:12535, 16962 */
    DWTB200_CDF97_I32 = 5, /* dwt_cdf97_2f_i / dwt_cdf97_2i_i   src/libdwt.c:16387, 18219 (9/7 with integer lifting) */
    DWTB200_KIND_COUNT = 6
};

enum {
    DWTB200_OK = 0,
    DWTB200_ENODEV = -1,  /* no CUDA device / driver */
    DWTB200_ECUDA = -2,   /* a CUDA call failed */
    DWTB200_EINVAL = -3,  /* bad argument */
    DWTB200_ENOMEM = -4
};

/* ---- lifecycle: dwt_util_init / dwt_util_finish / dwt_util_abort (src/libdwt.c:19158, 19186, 19200) ---- */
int dwtb200_init(int device);      /* device < 0: LOCAL_RANK env or 0.  Idempotent. */
void dwtb200_finish(void);
const char *dwtb200_last_error(void);
int dwtb200_device_count(void);
int dwtb200_device(void);          /* device in use, -1 before init */

/* ---- pinned host images: dwt_util_alloc_image / dwt_util_free_image (src/libdwt.c:1437, 1482) ---- */
/* page-locked, page-aligned; on a multi-socket host the pages are bound to the NUMA node of the GPU in use (DWTB200_NUMA=0: off) */
void *dwtb200_host_alloc(size_t bytes);
int dwtb200_host_numa_node(void);         /* that node; -1: not bound (single node, unknown, or switched off) */
void dwtb200_host_free(void *ptr);

/* ---- level arithmetic shared with the reference drivers (src/libdwt.c:12807-12810, inline.h:443-460) ---- */
int dwtb200_ceil_log2(int x);
int dwtb200_clamp_j(int j_max, int size_o_big_x, int size_o_big_y, int decompose_one);

/* ---- reference-semantics transforms on HOST memory, in place, synchronous ----------------------
 * dwt_cdf97_2f_s/_d, dwt_cdf53_2f_i (src/libdwt.h:526-537, 562-573, 686-697) and the inverses
 * (src/libdwt.h:831-842, 867-878, 981-992): same argument meaning, Mallat layout, *j_max_ptr
 * clamped to ceil_log2(decompose_one ? max : min) and left holding the achieved depth.
 * H2D copy, transform, D2H copy; results are visible in ptr on return. */
int dwtb200_fwd2_host(int kind, void *ptr, int64_t stride_x, int64_t stride_y, int size_o_big_x, int size_o_big_y,
                      int size_i_big_x, int size_i_big_y, int *j_max_ptr, int decompose_one, int zero_padding);
int dwtb200_inv2_host(int kind, void *ptr, int64_t stride_x, int64_t stride_y, int size_o_big_x, int size_o_big_y,
                      int size_i_big_x, int size_i_big_y, int j_max, int decompose_one, int zero_padding);

/* out-of-place variants, dwt_cdf97_2f_s2 / dwt_cdf97_2i_s2 (src/libdwt.h:667-679, 962-974; src/libdwt.c:12619, 17985):
 * `src` is read, `dst` (same strides) receives the result; what the reference leaves of dst's old content in a sparse
 * layout is left here too */
int dwtb200_fwd2_host2(int kind, const void *src, void *dst, int64_t stride_x, int64_t stride_y, int size_o_big_x,
                       int size_o_big_y, int size_i_big_x, int size_i_big_y, int *j_max_ptr, int decompose_one, int zero_padding);
int dwtb200_inv2_host2(int kind, const void *src, void *dst, int64_t stride_x, int64_t stride_y, int size_o_big_x,
                       int size_o_big_y, int size_i_big_x, int size_i_big_y, int j_max, int decompose_one, int zero_padding);
/* interleaved in-place family: dwt_cdf97_2f_inplace_s and its bit-identical twins _inplace_sep_s, _inplace_sdl_s,
 * _inplace_sep_sdl_s, dwt_cdf97_2i_inplace_s (src/libdwt.h:586-662, 889-900; src/libdwt.c:12926, 13485, 13641, 14847,
 * 17474) with kind DWTB200_CDF97_F32, and dwt_cdf53_2f_inplace_s / dwt_cdf53_2i_inplace_s (src/libdwt.h:599, 944;
 * src/libdwt.c:16553, 17886) with kind DWTB200_CDF53_F32.  Coefficients stay interleaved (level j at stride 2^j,
 * even = L, odd = H); level sizes come from the inner size, the outer size only bounds the level count; no zero
 * padding (the reference ignores the flag).  The 9/7 results reproduce the reference's prolog / core / epilog sweep
 * order, which is NOT the rounding of the Mallat family in the top rows and right columns of every level. */
int dwtb200_fwd2_inplace_host(int kind, void *ptr, int64_t stride_x, int64_t stride_y, int size_o_big_x, int size_o_big_y,
                              int size_i_big_x, int size_i_big_y, int *j_max_ptr, int decompose_one);
int dwtb200_inv2_inplace_host(int kind, void *ptr, int64_t stride_x, int64_t stride_y, int size_o_big_x, int size_o_big_y,
                              int size_i_big_x, int size_i_big_y, int j_max, int decompose_one);
/* the reference's performance protocol on the device, dwt_util_perf_cdf97_2_s / dwt_util_perf_cdf53_2_i
 * (src/libdwt.h:2498-2513, src/libdwt.c:21391, 21262): M device-resident test images, N loops of M forward then M
 * inverse transforms, minimum over the loops of the mean seconds per transform (CUDA events, no host copies) */
int dwtb200_perf2(int kind, int size_o_big_x, int size_o_big_y, int size_i_big_x, int size_i_big_y, int j_max,
                  int decompose_one, int zero_padding, int M, int N, float *fwd_secs, float *inv_secs);

/* the same protocol for the interleaved in-place family, dwt_util_perf_cdf97_2_inplace_s and its _sep_s / _sdl_s / _sep_sdl_s
 * twins (src/libdwt.h:2520-2590) */
int dwtb200_perf2_inplace(int kind, int size_o_big_x, int size_o_big_y, int size_i_big_x, int size_i_big_y, int j_max,
                          int decompose_one, int M, int N, float *fwd_secs, float *inv_secs);

/* device time (CUDA events on the library stream) of the transform inside the last *_host call, in
 * milliseconds, host<->device copies excluded; replaces dwt_util_get_clock() brackets (src/libdwt.c:18701) */
double dwtb200_last_transform_ms(void);
/* the *_host calls keep a device mirror per sample type between calls; this frees them */
void dwtb200_release_host_cache(void);

/* ---- device-resident images (what the roofline numbers are measured on) ---------------------------
 * A dwtb200_image is `frames` independent planes of size_o_big_x x size_o_big_y samples living in
 * HBM (two ping-pong planes plus LL scratch).  fwd2/inv2 have the semantics above, leave the result
 * on the device and return without synchronising (work is ordered on the library stream). */
typedef struct dwtb200_image dwtb200_image;
dwtb200_image *dwtb200_image_create(int kind, int size_o_big_x, int size_o_big_y, int frames);
void dwtb200_image_destroy(dwtb200_image *img);
/* host <-> device, one frame; arbitrary byte strides (dwt_util_memcpy_stride_*, src/system.c:90-180).
 * upload is ASYNCHRONOUS: it enqueues the copy on the image's stream and returns; a pinned `host` buffer must stay
 * valid and unmodified until dwtb200_sync() or a download of the same image returns.  download synchronises. */
int dwtb200_image_upload(dwtb200_image *img, int frame, const void *host, int64_t stride_x, int64_t stride_y);
int dwtb200_image_download(dwtb200_image *img, int frame, void *host, int64_t stride_x, int64_t stride_y);
/* dwt_util_test_image_fill{,2}_{s,d,i} on the device (src/libdwt.c:1247-1385); frame k uses
 * rand = (rand_mod > 0 ? (rand + k) % rand_mod : rand), cf. volume_fill_s (src/volume.c:41); with rand_mod > 0, `rand` is
 * the global index of frame 0 when the batch is one shard of a larger one */
int dwtb200_image_fill(dwtb200_image *img, int rand, int type, int rand_mod);
/* the same with the pattern's row coordinate shifted by y_offset (a row strip of a larger image) and, with
 * wide != 0, the products evaluated in 64 bits: the extension of the patterns to images the reference cannot
 * address (src/libdwt.c:1154, 1216 overflow `int` from ~32768) */
int dwtb200_image_fill_ex(dwtb200_image *img, int rand, int type, int rand_mod, int y_offset, int wide);
/* rows [row0, row0+rows) of a frame <-> a dense buffer on the host, this device or a peer device (halo rows
 * of a row-strip partition; replaces nothing in the reference, which has no multi-device code) */
int dwtb200_image_copy_rows(dwtb200_image *img, int frame, int row0, int rows, void *buf, int64_t buf_pitch_bytes, int to_image);
/* CUDA IPC handle (64 bytes) of the image's current plane, and mapping / unmapping it in another process */
int dwtb200_image_ipc_export(dwtb200_image *img, void *handle64);
void *dwtb200_ipc_open(const void *handle64);
int dwtb200_ipc_close(void *ptr);
int dwtb200_image_fwd2(dwtb200_image *img, int size_i_big_x, int size_i_big_y, int *j_max_ptr, int decompose_one,
                       int zero_padding);
int dwtb200_image_inv2(dwtb200_image *img, int size_i_big_x, int size_i_big_y, int j_max, int decompose_one,
                       int zero_padding);
/* the interleaved in-place family (see dwtb200_fwd2_inplace_host) on a device-resident image, inner == outer size */
int dwtb200_image_fwd2_inplace(dwtb200_image *img, int *j_max_ptr, int decompose_one);
int dwtb200_image_inv2_inplace(dwtb200_image *img, int j_max, int decompose_one);
/* current plane of frame 0 (device pointer) and its pitch in bytes; frames are frame_bytes apart */
void *dwtb200_image_devptr(dwtb200_image *img, size_t *pitch_bytes, size_t *frame_bytes);
/* dwt_util_subband_{s,d,i} (src/libdwt.h:2286-2340, src/libdwt.c:20731) for a device-resident image: device pointer, pitch and
 * inner size of subband `band` (0 LL, 1 HL, 2 LH, 3 HH: enum dwt_subbands) of level j inside the current Mallat plane */
int dwtb200_image_subband(dwtb200_image *img, int frame, int size_i_big_x, int size_i_big_y, int j, int band, void **dev_ptr,
                          size_t *pitch_bytes, int *size_x, int *size_y);
/* sum, sum of squares and max |x| of that subband, accumulated in double on the device: the inputs of the per-subband
 * feature reductions (dwt_util_wps_s, _var_s, _norm_s ..., src/libdwt.c:23086-23786) without a device-to-host copy */
int dwtb200_image_subband_moments(dwtb200_image *img, int frame, int size_i_big_x, int size_i_big_y, int j, int band, double *sum,
                                  double *sum_sq, double *max_abs);
/* the reference's per-subband feature vectors on a device-resident Mallat plane: dwt_util_wps_s, dwt_util_mean_s, dwt_util_var_s,
 * dwt_util_stdev_s, dwt_util_maxnorm_s, dwt_util_norm_s (src/libdwt.h; src/libdwt.c:23201, 23515, 23549, 23583, 23686, 23754): one
 * value per non-empty HL, LH, HH band of the levels j = 1 .. j_max - 1, in that order; *count receives their number.  Sums are
 * accumulated in double on the device (the reference sums sequentially in float), only the feature vector crosses PCIe. */
enum { DWTB200_FEAT_WPS = 0, DWTB200_FEAT_MEAN = 1, DWTB200_FEAT_VAR = 2, DWTB200_FEAT_STDEV = 3, DWTB200_FEAT_MAXNORM = 4, DWTB200_FEAT_NORM = 5 };
int dwtb200_image_features(dwtb200_image *img, int frame, int size_i_big_x, int size_i_big_y, int j_max, int feature, float *fv, int *count);
/* dwt_util_conv_show_{s,d,i} (src/libdwt.h; src/libdwt.c:21075, 21120, 21020) on device-resident planes: the current plane of every
 * frame of `dst` receives log(1 + |c| * 100) / 10 (float, double; logarithm in double as the reference's log_i_s) or |c| (int) of
 * the top-left size_i_big_x x size_i_big_y samples of `src`; src == dst is allowed */
int dwtb200_image_conv_show(dwtb200_image *src, dwtb200_image *dst, int size_i_big_x, int size_i_big_y);
/* dwt_util_save_to_pgm_s / _d (src/libdwt.c:19794, 19877) for a device-resident frame: grey values computed on the device (one byte
 * per sample crosses PCIe instead of the plane), the same "P2" text file written by the host */
int dwtb200_image_save_pgm(dwtb200_image *img, int frame, const char *filename, double max_value, int size_i_big_x, int size_i_big_y);
/* dwt_util_save_sym_to_pgm_s (src/libdwt.c:26184): samples in [-max_value, +max_value], shifted by +max_value and written against
 * twice the maximum (both in the image's own precision, as dwt_util_shift_s and the reference's call do) */
int dwtb200_image_save_sym_pgm(dwtb200_image *img, int frame, const char *filename, double max_value, int size_i_big_x, int size_i_big_y);
/* dwt_util_save_to_mat_s (src/libdwt.c:24430) for a device-resident frame: "%f" values separated by commas, one row per line */
int dwtb200_image_save_mat(dwtb200_image *img, int frame, const char *filename, int size_i_big_x, int size_i_big_y);
/* bit-exact comparison of the current planes of two images on the device: number of differing samples */
int64_t dwtb200_image_diff(dwtb200_image *a, dwtb200_image *b);
/* max |a-b| over the current planes (float/double kinds), cf. dwt_util_compare_s (src/libdwt.c:1593) */
double dwtb200_image_maxabs(dwtb200_image *a, dwtb200_image *b);
int dwtb200_image_copy(dwtb200_image *dst, dwtb200_image *src);
/* kernels launched by the last fwd2/inv2 on this image (graph nodes) */
int dwtb200_image_last_launches(dwtb200_image *img);
/* 0: streaming + tail kernels (dense planes), 1: generic pass kernels; which one the last call used */
int dwtb200_image_last_path(dwtb200_image *img);
/* testing hook: force the generic pass kernels (1) or let the library choose (0) */
void dwtb200_force_generic(int on);
/* tuning hook: output rows per strip of the streaming kernels (0 = heuristic) */
void dwtb200_set_strip_rows(int rows);
/* tuning hooks for the kernel selection per level (defaults in parentheses):
 *   DWTB200_TUNE_TILE_MAX  a level with <= value samples over all frames takes the tile kernels, larger
 *                          levels the streaming kernels (2048*2048)
 *   DWTB200_TUNE_TAIL_MAX  the single-launch tail starts at the first level with <= value samples per
 *                          frame (64*64; 0 disables the tail)
 *   DWTB200_TUNE_MID_MAX   accepted and ignored (the persistent mid-level kernels of round 1 were never faster and are gone)
 *   DWTB200_TUNE_PDL       1: kernels of a pyramid are chained by programmatic dependent launch (0: measured no gain)
 *   DWTB200_TUNE_NARROW    1: streaming kernels hold 16 instead of 32 bytes per lane: twice the warps per SM (0)
 *   DWTB200_TUNE_PIPELINE  1: the *_host calls overlap upload, level-0 strips / z ranges and download for large dense images and volumes (1)
 *   DWTB200_TUNE_RING      bit 0 / bit 1: forward / inverse streaming levels stage their input through a shared-memory
 *                          ring filled by the bulk-copy engine (cp.async.bulk + mbarrier) instead of a register double buffer;
 *                          bits 4-6 force a CTA shape (0 = chosen per level: 7 consumer warps x 2 CTAs per SM, or 5 x 3 for
 *                          large batches of 2048-wide frames; 1 = 15 x 1, 2 = 8 x 2, 3 = 5 x 3, 5 = 7 x 2)
 *   DWTB200_TUNE_VOL3      1 (default): the 3-D transforms of volumes of at least 64 x 32 x 16 run in ONE pass over the volume, tiles staged by
 *                          tensor copies (k_vol3t); 2: the same pass staged by cp.async (k_vol3, round 1);
 *                          0: always x + y per slice, then z (two passes)
 *   DWTB200_TUNE_PYR       accepted and ignored (the fused tile-pyramid kernels of round 1 were never faster and are gone)
 *   DWTB200_TUNE_CHAIN     1: the kernels of a pyramid are launched with programmatic stream serialization and wait for
 *                          their input row block by row block on completion counters, so consecutive levels overlap (1) */
enum { DWTB200_TUNE_TILE_MAX = 0, DWTB200_TUNE_TAIL_MAX = 1, DWTB200_TUNE_MID_MAX = 2, DWTB200_TUNE_PDL = 3, DWTB200_TUNE_NARROW = 4,
       DWTB200_TUNE_PIPELINE = 5, DWTB200_TUNE_RING = 6, DWTB200_TUNE_CHAIN = 7, DWTB200_TUNE_PYR = 8, DWTB200_TUNE_VOL3 = 9 };
int dwtb200_set_tuning(int key, long long value);

/* ---- 3-D, one level, interleaved subbands (src/volume-dwt.c:727, 677, 1115; struct volume_t
 * src/volume.h:14-24: stride_x = pixel, stride_y = row, stride_z = slice, all in bytes) --------
 * The host calls are synchronous (results in the caller's memory at return); src and dst may be the same volume.  Volumes of
 * 64 MiB and more go through a pipeline: z ranges are uploaded, transformed and downloaded concurrently (pin the volumes --
 * volume_alloc_realiably_locked of the compat layer, dwtb200_host_alloc -- for the copies to overlap). */
int dwtb200_fwd3_host(const void *src, size_t s_stride_x, size_t s_stride_y, size_t s_stride_z, void *dst,
                      size_t d_stride_x, size_t d_stride_y, size_t d_stride_z, int size_x, int size_y, int size_z);
int dwtb200_inv3_host(void *vol, size_t stride_x, size_t stride_y, size_t stride_z, int size_x, int size_y,
                      int size_z);
typedef struct dwtb200_volume dwtb200_volume;
dwtb200_volume *dwtb200_volume_create(int size_x, int size_y, int size_z);
void dwtb200_volume_destroy(dwtb200_volume *v);
/* upload is asynchronous like dwtb200_image_upload; download synchronises */
int dwtb200_volume_upload(dwtb200_volume *v, const void *host, size_t stride_x, size_t stride_y, size_t stride_z);
int dwtb200_volume_download(dwtb200_volume *v, void *host, size_t stride_x, size_t stride_y, size_t stride_z);
int dwtb200_volume_fill(dwtb200_volume *v);   /* volume_fill_s, src/volume.c:41 */
int dwtb200_volume_fwd3(dwtb200_volume *v);   /* cdf97_3f_ip_sep_horizontal_s */
int dwtb200_volume_inv3(dwtb200_volume *v);   /* cdf97_3i_ip_sep_horizontal_s */

/* volume_perftest_fwd97op_s (src/volume-dwt.h:234, src/volume-dwt.c:2810) on the device: N runs of fill / forward (timed) /
 * inverse / compare; minimum forward seconds per voxel, number of failed round trips */
int dwtb200_perf3(int size, int N, double *secs_per_voxel, int *errors);

/* ---- ONE image as row strips over the GPUs of a box (BASELINE config 5a) ----------------------------
 * Replaces nothing in the reference, which has no multi-device code and cannot address such an image (`y * stride_x` in int,
 * src/inline.h:188); the transform is dwt_cdf97_2f_s / dwt_cdf97_2i_s and siblings (src/libdwt.c:12776-12924, 17040-17180) of
 * the WHOLE picture, full depth, Mallat layout, bit-identical to the single-device result.
 * One process per GPU.  Rank r owns rows [own0, own1) and holds the extended strip [ext0, ext1) (owned rows + halo) in an
 * ordinary dwtb200_image (dwtb200_strips_image: fill / upload / download it like any image; local row = global row - ext0).
 * The first levels_distributed levels run on every rank's strip; the halo rows and LL band move between the ranks as
 * peer-to-peer copies over NVLink (CUDA IPC mappings) ordered by device-side flags -- the calls enqueue work and return, there is
 * no host synchronisation inside them; rank 0 holds the remaining levels in its `top` image (the Mallat image of LL_Jd).
 * The calls are collective: every rank makes the same sequence of dwtb200_strips_fwd2 / _inv2 calls. */
typedef struct dwtb200_strips dwtb200_strips;
typedef struct {
    int halo;                 /* rows of level-0 input held beyond each end of the owned rows: HALO * 2^levels_distributed */
    int own0, own1;           /* image rows this rank owns */
    int ext0, ext1;           /* image rows it holds */
    int ll_w, ll_h;           /* size of LL after the distributed levels (the top image) */
    int ll_own0, ll_own1;     /* rows of that band this rank owns ... */
    int ll_ext0, ll_ext1;     /* ... and holds */
    int neighbours_only;      /* 1: every strip is at least as tall as the halo (required) */
} dwtb200_strip_plan;
/* geometry only (no device needed); halo_lines = lifting reach per level: 4 for CDF 9/7, 2 for CDF 5/3 */
int dwtb200_strips_plan(int width, int height, int world, int levels_distributed, int halo_lines, int rank, dwtb200_strip_plan *out);
/* rows of distributed level j in its output resolution: out[11] = { off, nly_global, nly_local, extL0, extL1, extH0, extH1,
 * ownL0, ownL1, ownH0, ownH1 } (L: LL/HL rows, H: LH/HH rows; global row numbers; local row = global - off, H rows after nly_local) */
int dwtb200_strips_band(int width, int height, int world, int levels_distributed, int halo_lines, int rank, int j, int *out);
/* levels_distributed <= 0: chosen so that LL has at most 2048^2 samples.  `session`: a name unique to this strips object and
 * common to its ranks (a POSIX shared-memory segment of that name carries the IPC handles); rank 0 must be created first when
 * several ranks live in one process (single-GPU tests).  Returns NULL on failure (dwtb200_last_error). */
dwtb200_strips *dwtb200_strips_create(int kind, int width, int height, int levels_distributed, int rank, int world, const char *session);
/* waits for every rank's create and maps their planes.  Called explicitly (before the strip holds data: it runs warm-up transforms)
 * it also lets level 0 of the forward transform read the halo rows straight from the neighbours' planes over NVLink instead of
 * copying them first (DWTB200_STRIPS_DIRECT=0: off); the first transform implies a plain connect without that */
int dwtb200_strips_connect(dwtb200_strips *s);
void dwtb200_strips_destroy(dwtb200_strips *s);
dwtb200_image *dwtb200_strips_image(dwtb200_strips *s);   /* the extended strip */
dwtb200_image *dwtb200_strips_top(dwtb200_strips *s);     /* rank 0: Mallat image of LL_Jd; NULL elsewhere */
int dwtb200_strips_levels(dwtb200_strips *s, int *j_total, int *j_distributed);
int dwtb200_strips_get_plan(dwtb200_strips *s, dwtb200_strip_plan *out);
int dwtb200_strips_fwd2(dwtb200_strips *s, int *j_max_ptr);   /* *j_max_ptr receives the depth (ceil_log2(min(width, height))) */
int dwtb200_strips_inv2(dwtb200_strips *s, int j_max);        /* j_max: that depth, or -1 */
int dwtb200_strips_sync(dwtb200_strips *s);   /* waits for this rank's queued work; fails if a wait for a peer timed out */
unsigned long long dwtb200_strips_last_peer_bytes(dwtb200_strips *s);   /* bytes the last call moved between this rank and its peers */
/* verification: differing samples (bit-wise) between the rows this rank owns and a single-device image of the whole picture on
 * this device (mallat != 0: forward coefficients incl. rank 0's top; 0: image rows); -1 on error */
int64_t dwtb200_strips_compare_owned(dwtb200_strips *s, dwtb200_image *full, int mallat);
/* the owned rows written into a host image of the whole picture (stride_x bytes between rows) at their final positions */
int dwtb200_strips_download_owned(dwtb200_strips *s, void *host_full, int64_t stride_x, int mallat);

/* ---- device-event timing: replaces dwt_util_get_clock around transforms (src/libdwt.c:18701) ---- */
int dwtb200_sync(void);                 /* wait for the library stream */
int dwtb200_timer_start(void);          /* record an event on the library stream */
double dwtb200_timer_stop_ms(void);     /* record, synchronise, elapsed milliseconds (< 0 on error) */
void *dwtb200_stream(void);             /* cudaStream_t of the library stream */
/* the same for the calls made on ONE image: both events are recorded on the image's own stream (the one its kernels are
 * launched on), so nothing but that image's work lies between them */
int dwtb200_image_timer_start(dwtb200_image *img);
double dwtb200_image_timer_stop_ms(dwtb200_image *img);
/* a series of marks recorded on the image's stream WITHOUT synchronising (the host runs ahead, so no host latency falls between two
 * marks); _read waits for the last mark, stores the milliseconds between consecutive marks in ms[0 .. ) and returns their number
 * (< 0: error); dwtb200_image_wait orders `img`'s later work after everything queued so far on `other` (no host wait) */
int dwtb200_image_timer_mark(dwtb200_image *img);
int dwtb200_image_timer_read(dwtb200_image *img, double *ms, int capacity);
int dwtb200_image_wait(dwtb200_image *img, dwtb200_image *other);
/* write `bytes` of device scratch so the 126 MB L2 holds none of the previous step's data */
int dwtb200_flush_l2(size_t bytes);

#ifdef __cplusplus
}
#endif
#endif /* DWTB200_H */
