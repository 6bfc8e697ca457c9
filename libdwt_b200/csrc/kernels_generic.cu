// kernels_generic.cu -- exact-semantics pass kernels for sparse layouts (outer != inner), skinny
// planes and every other shape the streaming/tail kernels do not take.
//
// One pass = what one `for(y) dwt_cdfXX_{f,i}_ex_stride_T(...)` loop of a reference level driver does
// (/root/reference/src/libdwt.c:12837-12893 forward, 17098-17154 inverse), written out of place: the
// kernel rewrites the WHOLE outer region of the level in `dst` -- transformed samples where the
// reference writes them, everything else copied through -- so the host driver can ping-pong two
// full-size planes and stay equivalent to the reference's in-place update, including the parts of a
// sparse array that the reference transforms as a side effect.
#include "kernels.h"
#include "lifting.cuh"

namespace dwtb200 {

template <class WV> __global__ void __launch_bounds__(256) k_pass_fwd(PassParams p)
{
    using T = typename WV::T;
    const int px = blockIdx.x * blockDim.x + threadIdx.x;
    const int py = blockIdx.y * blockDim.y + threadIdx.y;
    if (px >= p.region_w || py >= p.region_h) return;
    const T *src = (const T *)p.src + (int64_t)blockIdx.z * p.src_frame;
    T *dst = (T *)p.dst + (int64_t)blockIdx.z * p.dst_frame;
    const int q = p.along_x ? px : py;        // position along the line
    const int line = p.along_x ? py : px;
    const int64_t se = p.along_x ? 1 : p.src_pitch, sl = p.along_x ? p.src_pitch : 1;
    const int64_t de = p.along_x ? 1 : p.dst_pitch, dl = p.along_x ? p.dst_pitch : 1;
    const T *s = src + line * sl;
    T *d = dst + line * dl;
    const int N = p.N, nl = (N + 1) >> 1, nh = N >> 1;

    if (N >= 2) {
        if (q < nl) {
            T w[2 * WV::HALO + 2];
#pragma unroll
            for (int i = 0; i < 2 * WV::HALO + 2; i++) w[i] = s[reflect(2 * q - WV::HALO + i, N) * se];
            T L, H;
            window_fwd<WV>(w, L, H);
            d[q * de] = L;
            if (q < nh) d[(p.off_h + q) * de] = H;
        }
        const bool in_l = q < nl, in_h = q >= p.off_h && q < p.off_h + nh;
        if (!in_l && !in_h && !p.keep_dst) d[q * de] = s[q * se];
    } else if (!p.keep_dst) {
        T v = s[q * se];
        if (N == 1 && q == 0 && WV::HAS_ONE) v = WV::one_f(v);
        d[q * de] = v;
    } else if (N == 1 && q == 0 && WV::HAS_ONE) {
        d[0] = WV::one_f(s[0]);
    }
}

template <class WV> __global__ void __launch_bounds__(256) k_pass_inv(PassParams p)
{
    using T = typename WV::T;
    const int px = blockIdx.x * blockDim.x + threadIdx.x;
    const int py = blockIdx.y * blockDim.y + threadIdx.y;
    if (px >= p.region_w || py >= p.region_h) return;
    const T *src = (const T *)p.src + (int64_t)blockIdx.z * p.src_frame;
    T *dst = (T *)p.dst + (int64_t)blockIdx.z * p.dst_frame;
    const int q = p.along_x ? px : py;
    const int line = p.along_x ? py : px;
    const int64_t se = p.along_x ? 1 : p.src_pitch, sl = p.along_x ? p.src_pitch : 1;
    const int64_t de = p.along_x ? 1 : p.dst_pitch, dl = p.along_x ? p.dst_pitch : 1;
    const T *s = src + line * sl;
    T *d = dst + line * dl;
    const int N = p.N, nl = (N + 1) >> 1;

    if (N >= 2) {
        if (q < nl) {
            T w[2 * WV::HALO + 2];
#pragma unroll
            for (int i = 0; i < 2 * WV::HALO + 2; i++) {
                const int c = reflect(2 * q - WV::HALO + i, N);   // interleaved index; parity survives the mirror
                const int pos = (c & 1) ? p.off_h + (c >> 1) : (c >> 1);
                w[i] = s[pos * se];
            }
            T E, O;
            window_inv<WV>(w, E, O);
            d[(2 * q) * de] = E;
            if (2 * q + 1 < N) d[(2 * q + 1) * de] = O;
        }
        if (q >= N) d[q * de] = s[q * se];
    } else {
        T v = s[q * se];
        if (N == 1 && q == 0 && WV::HAS_ONE) v = WV::one_i(v);
        d[q * de] = v;
    }
}

template <class T> __global__ void __launch_bounds__(256) k_zero(ZeroParams p)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= p.region_w || y >= p.region_h) return;
    const bool zx = (x >= p.x0a && x < p.x0b) || (x >= p.x1a && x < p.x1b);
    const bool zy = (y >= p.y0a && y < p.y0b) || (y >= p.y1a && y < p.y1b);
    if (zx || zy) ((T *)p.buf)[(int64_t)blockIdx.z * p.frame + (int64_t)y * p.pitch + x] = T(0);
}

template <class K> static cudaError_t touch(K kern)
{
    cudaFuncAttributes a;
    return cudaFuncGetAttributes(&a, kern);
}
cudaError_t preload_generic()
{
    cudaError_t e = cudaSuccess;
    for (int kind = 0; kind < K_COUNT; kind++)
        dispatch_kind(kind, [&](auto wv) {
            using WV = decltype(wv);
            if (e == cudaSuccess) e = touch(k_pass_fwd<WV>);
            if (e == cudaSuccess) e = touch(k_pass_inv<WV>);
        });
    if (e == cudaSuccess) e = touch(k_zero<double>);
    if (e == cudaSuccess) e = touch(k_zero<int32_t>);
    return e;
}
static dim3 grid2(int w, int h, int frames, dim3 b) { return dim3((w + b.x - 1) / b.x, (h + b.y - 1) / b.y, frames); }

void launch_pass_fwd(int kind, const PassParams &p, int frames, cudaStream_t st)
{
    if (p.region_w <= 0 || p.region_h <= 0) return;
    const dim3 b(32, 8), g = grid2(p.region_w, p.region_h, frames, b);
    dispatch_kind(kind, [&](auto wv) { k_pass_fwd<decltype(wv)><<<g, b, 0, st>>>(p); });
}
void launch_pass_inv(int kind, const PassParams &p, int frames, cudaStream_t st)
{
    if (p.region_w <= 0 || p.region_h <= 0) return;
    const dim3 b(32, 8), g = grid2(p.region_w, p.region_h, frames, b);
    dispatch_kind(kind, [&](auto wv) { k_pass_inv<decltype(wv)><<<g, b, 0, st>>>(p); });
}
void launch_zero(int kind, const ZeroParams &p, int frames, cudaStream_t st)
{
    if (p.region_w <= 0 || p.region_h <= 0) return;
    const dim3 b(32, 8), g = grid2(p.region_w, p.region_h, frames, b);
    if (kind_elem_size(kind) == 8) k_zero<double><<<g, b, 0, st>>>(p);
    else k_zero<int32_t><<<g, b, 0, st>>>(p);   // float 0.0f and int 0 share a bit pattern
}

}  // namespace dwtb200
