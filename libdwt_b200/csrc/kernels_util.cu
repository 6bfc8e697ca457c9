// kernels_util.cu -- test-pattern generators, layout repacking, 3-D axis lifting (sm_100a).
#include "kernels.h"
#include "lifting.cuh"

namespace dwtb200 {

// ---- test patterns: dwt_util_test_image_value_i_{s,d,i} (/root/reference/src/libdwt.c:1112-1244) ----
// The reference evaluates the products in 32-bit `int`; they wrap for large coordinates and the
// wrapped values are part of its observable output (4096^2 int image), so the products here are
// done in uint32_t and reinterpreted.
__device__ __forceinline__ int32_t wmul(int32_t a, int32_t b) { return (int32_t)((uint32_t)a * (uint32_t)b); }
__device__ __forceinline__ int32_t wadd32(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }

// WIDE = true: the products are evaluated in 64 bits (no wrap): the documented extension of the patterns
// to images the reference cannot address (>= 2 GiB, e.g. 65536^2), see oracle/dwt_oracle.c pat_*.
template <class T> __device__ __forceinline__ T pattern(int x, int y, int rnd, int type);
template <class T> __device__ __forceinline__ T pattern_wide(int x, int y, int rnd, int type);
template <> __device__ __forceinline__ float pattern_wide<float>(int x, int y, int rnd, int type)
{
    x++;
    y++;
    if (type == 2) return __fdiv_rn((float)((x ^ y) & 0xff), 32.0f);
    if (type == 3) return __fdiv_rn((float)((((x & 1) << 1) | (y & 1)) + 1), 4.0f);
    x >>= rnd;
    const long long num = 2ll * x * y, den = (long long)x * x + (long long)y * y + 1;
    return __fdiv_rn(__ll2float_rn(num), __ll2float_rn(den));
}
template <> __device__ __forceinline__ double pattern_wide<double>(int x, int y, int rnd, int)
{
    x >>= rnd;
    const long long num = 2ll * x * y, den = (long long)x * x + (long long)y * y + 1;
    return __ddiv_rn(__ll2double_rn(num), __ll2double_rn(den));
}
template <> __device__ __forceinline__ int32_t pattern_wide<int32_t>(int x, int y, int rnd, int type)
{
    if (type == 2) return (x ^ y) & 0xff;
    x >>= rnd;
    const long long num = 255ll * (2ll * x * y), den = (long long)x * x + (long long)y * y + 1;
    return (int32_t)(num / den);
}
template <> __device__ __forceinline__ float pattern<float>(int x, int y, int rnd, int type)
{
    x++;
    y++;
    if (type == 2) return __fdiv_rn((float)((x ^ y) & 0xff), 32.0f);
    if (type == 3) return __fdiv_rn((float)((((x & 1) << 1) | (y & 1)) + 1), 4.0f);
    x >>= rnd;
    const int32_t num = wmul(wmul(2, x), y);
    const int32_t den = wadd32(wadd32(wmul(x, x), wmul(y, y)), 1);
    return __fdiv_rn(__int2float_rn(num), __int2float_rn(den));
}
template <> __device__ __forceinline__ double pattern<double>(int x, int y, int rnd, int)
{
    x >>= rnd;
    const int32_t num = wmul(wmul(2, x), y);
    const int32_t den = wadd32(wadd32(wmul(x, x), wmul(y, y)), 1);
    return __ddiv_rn((double)num, (double)den);
}
template <> __device__ __forceinline__ int32_t pattern<int32_t>(int x, int y, int rnd, int type)
{
    if (type == 2) return (x ^ y) & 0xff;
    x >>= rnd;
    const int32_t num = wmul(255, wmul(wmul(2, x), y));
    const int32_t den = wadd32(wadd32(wmul(x, x), wmul(y, y)), 1);
    if (den == 0) return 0;
    if (num == INT32_MIN && den == -1) return INT32_MIN;
    return num / den;
}

template <class T>
__global__ void __launch_bounds__(256)
k_fill(T *buf, int64_t pitch, int64_t frame, int nx, int ny, int rnd, int type, int mod, int y_offset, int wide)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= nx || y >= ny) return;
    const int r = mod > 0 ? (int)((blockIdx.z + (unsigned)rnd) % mod) : rnd;   // frame k of a batch that starts at global frame `rnd`
    buf[(int64_t)blockIdx.z * frame + (int64_t)y * pitch + x] =
        wide ? pattern_wide<T>(x, y + y_offset, r, type) : pattern<T>(x, y + y_offset, r, type);
}

void launch_fill(int kind, void *buf, int64_t pitch, int64_t frame, int nx, int ny, int rnd, int type, int mod,
                 int frames, int y_offset, int wide, cudaStream_t st)
{
    const dim3 b(32, 8), g((nx + 31) / 32, (ny + 7) / 8, frames);
    const int cls = kind_elem_class(kind);
    if (cls == 1) k_fill<float><<<g, b, 0, st>>>((float *)buf, pitch, frame, nx, ny, rnd, type, mod, y_offset, wide);
    else if (cls == 2) k_fill<double><<<g, b, 0, st>>>((double *)buf, pitch, frame, nx, ny, rnd, type, mod, y_offset, wide);
    else k_fill<int32_t><<<g, b, 0, st>>>((int32_t *)buf, pitch, frame, nx, ny, rnd, type, mod, y_offset, wide);
}

// ---- repack: caller layout (arbitrary byte strides, staged verbatim on the device) <-> dense plane ----
template <int ES>
__global__ void __launch_bounds__(256) k_repack(unsigned char *plane, int64_t pitch_bytes, unsigned char *staged, int64_t sx,
                                                int64_t sy, int nx, int ny, int to_plane)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= nx || y >= ny) return;
    unsigned char *a = plane + (int64_t)y * pitch_bytes + (int64_t)x * ES;
    unsigned char *b = staged + (int64_t)y * sx + (int64_t)x * sy;
    unsigned char *d = to_plane ? a : b, *s = to_plane ? b : a;
#pragma unroll
    for (int i = 0; i < ES; i++) d[i] = s[i];   // byte-wise: caller elements may be unaligned
}
void launch_repack(int es, void *plane, int64_t pitch_elems, void *staged, int64_t sx, int64_t sy, int nx, int ny,
                   int to_plane, cudaStream_t st)
{
    const dim3 b(32, 8), g((nx + 31) / 32, (ny + 7) / 8);
    if (es == 8) k_repack<8><<<g, b, 0, st>>>((unsigned char *)plane, pitch_elems * 8, (unsigned char *)staged, sx, sy, nx, ny, to_plane);
    else k_repack<4><<<g, b, 0, st>>>((unsigned char *)plane, pitch_elems * 4, (unsigned char *)staged, sx, sy, nx, ny, to_plane);
}

template <class T>
__global__ void __launch_bounds__(256) k_copy2d(T *dst, int64_t dp, const T *src, int64_t sp, int w, int h, int64_t df, int64_t sf)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    dst[(int64_t)blockIdx.z * df + (int64_t)y * dp + x] = src[(int64_t)blockIdx.z * sf + (int64_t)y * sp + x];
}
cudaError_t preload_util()
{
    cudaFuncAttributes a;
    cudaError_t e = cudaFuncGetAttributes(&a, k_copy2d<double>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, k_copy2d<int32_t>);
    return e;
}
void launch_copy2d(int es, void *dst, int64_t dp, const void *src, int64_t sp, int w, int h, int64_t df, int64_t sf,
                   int frames, cudaStream_t st)
{
    if (w <= 0 || h <= 0) return;
    const dim3 b(32, 8), g((w + 31) / 32, (h + 7) / 8, frames);
    if (es == 8) k_copy2d<double><<<g, b, 0, st>>>((double *)dst, dp, (const double *)src, sp, w, h, df, sf);
    else k_copy2d<int32_t><<<g, b, 0, st>>>((int32_t *)dst, dp, (const int32_t *)src, sp, w, h, df, sf);
}

// ---- comparison on the device: differing samples (bit patterns) and max |a-b| ---------------------
// dwt_util_compare_{s,d,i} (src/libdwt.c:1593, 1502, 1531) without the D2H of both images.
// out[0] = number of samples whose bit patterns differ, out[1] = max |a-b| as the bits of a double
// (non-negative doubles order like unsigned integers; a NaN difference is reported as +inf).
template <class T>
__global__ void __launch_bounds__(256) k_compare(const T *a, const T *b, int64_t pitch, int64_t frame, int nx, int ny,
                                                 unsigned long long *out)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    unsigned long long cnt = 0;
    double m = 0.0;
    if (x < nx && y < ny) {
        const int64_t o = (int64_t)blockIdx.z * frame + (int64_t)y * pitch + x;
        const T va = a[o], vb = b[o];
        if constexpr (sizeof(T) == 8) cnt = __double_as_longlong((double)va) != __double_as_longlong((double)vb);
        else if constexpr (sizeof(T) == 4 && !(T(1) / T(2) > T(0))) cnt = va != vb;
        else cnt = __float_as_int((float)va) != __float_as_int((float)vb);
        double d = fabs((double)va - (double)vb);
        if (!(d == d)) d = __longlong_as_double(0x7ff0000000000000ll);
        m = d;
    }
    cnt = __reduce_add_sync(0xffffffffu, (unsigned)cnt);
    for (int o = 16; o; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) {
        if (cnt) atomicAdd(out, cnt);
        if (m > 0.0) atomicMax(out + 1, (unsigned long long)__double_as_longlong(m));
    }
}
void launch_compare(int es, const void *a, const void *b, int64_t pitch, int64_t frame, int nx, int ny, int frames, int mode,
                    unsigned long long *out, cudaStream_t st)
{
    (void)es;
    const dim3 blk(32, 8), g((nx + 31) / 32, (ny + 7) / 8, frames);
    if (mode == 2) k_compare<double><<<g, blk, 0, st>>>((const double *)a, (const double *)b, pitch, frame, nx, ny, out);
    else if (mode == 1) k_compare<float><<<g, blk, 0, st>>>((const float *)a, (const float *)b, pitch, frame, nx, ny, out);
    else k_compare<int32_t><<<g, blk, 0, st>>>((const int32_t *)a, (const int32_t *)b, pitch, frame, nx, ny, out);
}

// ---- visualisation of a device-resident plane without bringing the coefficients to the host -----------------------------
// dwt_util_conv_show_{s,d,i} (src/libdwt.c:21075, 21120, 21020): float / double: log(1 + |c| * 100) / 10 -- the reference takes
// the logarithm in double (log_i_s, :21010) and rounds it to the sample type before the division --, a non-finite result
// becomes 0; int: |c|.
template <class T> __global__ void __launch_bounds__(256) k_conv_show(const T *src, int64_t sp, T *dst, int64_t dp, int nx, int ny)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= nx || y >= ny) return;
    const T c = src[(int64_t)y * sp + x];
    T r;
    if constexpr (sizeof(T) == 8) {
        r = __ddiv_rn(log(__dadd_rn(1.0, __dmul_rn(fabs((double)c), 100.0))), 10.0);
        if (!isfinite((double)r)) r = 0;
    } else if constexpr (T(1) / T(2) > T(0)) {
        const float t = (float)log((double)__fadd_rn(1.f, __fmul_rn(fabsf((float)c), 100.f)));
        r = __fdiv_rn(t, 10.f);
        if (!isfinite((float)r)) r = 0;
    } else {
        r = c < 0 ? (T)(0u - (uint32_t)c) : c;   // abs(); INT_MIN stays INT_MIN as in C
    }
    dst[(int64_t)y * dp + x] = r;
}
void launch_conv_show(int elem_class, const void *src, int64_t sp, void *dst, int64_t dp, int nx, int ny, cudaStream_t st)
{
    if (nx <= 0 || ny <= 0) return;
    const dim3 b(32, 8), g((nx + 31) / 32, (ny + 7) / 8);
    if (elem_class == 2) k_conv_show<double><<<g, b, 0, st>>>((const double *)src, sp, (double *)dst, dp, nx, ny);
    else if (elem_class == 1) k_conv_show<float><<<g, b, 0, st>>>((const float *)src, sp, (float *)dst, dp, nx, ny);
    else k_conv_show<int32_t><<<g, b, 0, st>>>((const int32_t *)src, sp, (int32_t *)dst, dp, nx, ny);
}
// the grey value dwt_util_save_to_pgm_s writes for a sample (src/libdwt.c:19794-19866): (int)(255 * px / max_value) in float,
// 255 above max_value, 0 below zero and for NaN; one byte per sample is all that crosses PCIe
// `shift` is added first (dwt_util_shift_s, src/libdwt.c:25530: the symmetric writer dwt_util_save_sym_to_pgm_s, :26184, shifts by
// +max and saves with 2 max)
template <class T> __global__ void __launch_bounds__(256) k_pgm_quant(const T *src, int64_t sp, unsigned char *dst, int nx, int ny, T maxv, T shift, int shifted)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= nx || y >= ny) return;
    T px = src[(int64_t)y * sp + x];
    if (shifted) {
        if constexpr (sizeof(T) == 8) px = __dadd_rn(px, shift);
        else px = __fadd_rn(px, shift);
    }
    int val;
    if constexpr (sizeof(T) == 8) val = (int)__ddiv_rn(__dmul_rn(255.0, px), maxv);
    else val = (int)__fdiv_rn(__fmul_rn(255.f, px), maxv);
    if (px != px) val = 0;
    if (px > maxv) val = 255;
    if (px < T(0)) val = 0;
    dst[(int64_t)y * nx + x] = (unsigned char)min(max(val, 0), 255);
}
void launch_pgm_quant(int elem_class, const void *src, int64_t sp, unsigned char *dst, int nx, int ny, double maxv, double shift, int shifted,
                      cudaStream_t st)
{
    if (nx <= 0 || ny <= 0) return;
    const dim3 b(32, 8), g((nx + 31) / 32, (ny + 7) / 8);
    if (elem_class == 2) k_pgm_quant<double><<<g, b, 0, st>>>((const double *)src, sp, dst, nx, ny, maxv, shift, shifted);
    else k_pgm_quant<float><<<g, b, 0, st>>>((const float *)src, sp, dst, nx, ny, (float)maxv, (float)shift, shifted);
}

// ---- moments of a rectangle (a subband of the Mallat plane): sum, sum of squares, max |x|, in double ----
// what dwt_util_band_wps_s / _var_s / _norm_s style feature reductions need (src/libdwt.c:23086-23786) without bringing the
// coefficients back to the host
template <class T> __global__ void __launch_bounds__(256) k_moments(const T *a, int64_t pitch, int nx, int ny, double *out)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    double v = 0.0;
    if (x < nx && y < ny) v = (double)a[(int64_t)y * pitch + x];
    double s = v, q = v * v, m = fabs(v);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_down_sync(0xffffffffu, s, o);
        q += __shfl_down_sync(0xffffffffu, q, o);
        m = fmax(m, __shfl_down_sync(0xffffffffu, m, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(out, s);
        atomicAdd(out + 1, q);
        atomicMax((unsigned long long *)(out + 2), (unsigned long long)__double_as_longlong(m));   // m >= 0: bit order = value order
    }
}
void launch_moments(int elem_class, const void *a, int64_t pitch, int nx, int ny, double *out, cudaStream_t st)
{
    if (nx <= 0 || ny <= 0) return;
    const dim3 blk(32, 8), g((nx + 31) / 32, (ny + 7) / 8);
    if (elem_class == 2) k_moments<double><<<g, blk, 0, st>>>((const double *)a, pitch, nx, ny, out);
    else if (elem_class == 1) k_moments<float><<<g, blk, 0, st>>>((const float *)a, pitch, nx, ny, out);
    else k_moments<int32_t><<<g, blk, 0, st>>>((const int32_t *)a, pitch, nx, ny, out);
}

// ---- volume_fill_s (src/volume.c:41): slice z = 2-D type-0 float pattern with rand = fold(z & 11) ----
__global__ void __launch_bounds__(256) k_volume_fill(float *buf, int64_t pitch, int64_t slice, int nx, int ny)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= nx || y >= ny) return;
    const int z = blockIdx.z;
    int rnd = z & 11;
    if (rnd > 11 / 2) rnd = 11 - rnd;
    buf[(int64_t)z * slice + (int64_t)y * pitch + x] = pattern<float>(x, y, rnd, 0);
}
void launch_volume_fill(float *buf, int64_t pitch, int64_t slice, int nx, int ny, int nz, cudaStream_t st)
{
    const dim3 b(32, 8), g((nx + 31) / 32, (ny + 7) / 8, nz);
    k_volume_fill<<<g, b, 0, st>>>(buf, pitch, slice, nx, ny);
}

// ---- 3-D: one interleaved lifting pass along one axis of a float volume ---------------------------
// fdwt1_single_cdf97_horizontal_min5_s (src/dwt-simple.c:2166) forward, dwt_cdf97_1i_inplace_s
// (src/libdwt.c:17182, one level) inverse: subbands stay interleaved (even = L, odd = H).
template <bool INV> __global__ void __launch_bounds__(256) k_axis3(Axis3Params p)
{
    const int a = blockIdx.x * blockDim.x + threadIdx.x;   // fastest line index
    const int b = blockIdx.y;
    const int k = blockIdx.z * blockDim.z + threadIdx.z;   // pair along the lifting axis
    const int nl = (p.N + 1) >> 1;
    if (a >= p.n0 || b >= p.n1 || k >= nl) return;
    const float *s = p.src + a * p.s_line0 + b * p.s_line1;
    float *d = p.dst + a * p.d_line0 + b * p.d_line1;
    if (p.N < 2) {
        d[0] = s[0];
        return;
    }
    float w[10];
#pragma unroll
    for (int i = 0; i < 10; i++) w[i] = s[(int64_t)reflect(2 * k - 4 + i, p.N) * p.s_elem];
    float e, o;
    if (INV) window_inv<W97F>(w, e, o);
    else window_fwd<W97F>(w, e, o);
    d[(int64_t)(2 * k) * p.d_elem] = e;
    if (2 * k + 1 < p.N) d[(int64_t)(2 * k + 1) * p.d_elem] = o;
}
void launch_axis3(const Axis3Params &p, int inverse, cudaStream_t st)
{
    const int nl = (p.N + 1) >> 1;
    if (p.n0 <= 0 || p.n1 <= 0 || nl <= 0) return;
    const dim3 b(64, 1, 4), g((p.n0 + 63) / 64, p.n1, (nl + 3) / 4);
    if (inverse) k_axis3<true><<<g, b, 0, st>>>(p);
    else k_axis3<false><<<g, b, 0, st>>>(p);
}

}  // namespace dwtb200
