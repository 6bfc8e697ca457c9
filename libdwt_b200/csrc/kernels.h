// kernels.h -- parameter blocks and launchers of the sm_100a kernels (internal to libdwtb200.so).
//
// All pitches / strides / offsets are in ELEMENTS of the plane's type.  Every kernel takes a frame
// index from gridDim.z (or .y where noted) so a batch of independent frames is one launch.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace dwtb200 {

// force-load the kernels of each translation unit (lazy loading is illegal during stream capture)
cudaError_t preload_stream();
cudaError_t preload_tail();
cudaError_t preload_generic();
cudaError_t preload_util();
cudaError_t preload_tile();

// ---- dataflow links between the kernels of one transform (chain.cuh) ---------------------------------
// The kernels of a pyramid are launched back to back with programmatic stream serialization: a kernel may
// start as soon as every CTA of its predecessor has STARTED, and waits for its input row block by row block
// on completion counters its predecessor bumps, instead of waiting for the predecessor to finish.  The tail
// of level j then overlaps the head of level j+1 and the small levels cost a flag round trip, not a launch.
// Counters are never reset: a block is complete when its counter reaches (gen + 1) * need, and the
// transform's last chained kernel increments the generation word when its last CTA retires.
struct Chain {
    const uint32_t *gen;   // generation word (nullptr: this launch is not chained at all)
    const uint32_t *in;    // counters of the kernel producing this kernel's LL input (nullptr: input complete at launch)
    uint32_t *out;         // counters this kernel's CTAs bump once their LL output rows are written (nullptr: nobody waits)
    uint32_t *done;        // != nullptr: last chained kernel of the transform, its last CTA bumps *gen
    int in_nblocks, in_bias, in_div, in_need;   // input row r lives in block (r + in_bias) / in_div, complete after in_need arrivals
    int out_nblocks;       // counters per frame of `out`
    int pdl;               // launch attribute: programmatic stream serialization
    unsigned total;        // CTAs of this launch
};

// ---- streaming level kernels (dense planes): one launch = one decomposition level ---------------
// Forward: reads the level's LL input (W x H) once, writes LL' (to `ll`) and HL/LH/HH (Mallat
// positions inside the output plane) once.  Inverse is the mirror image.
struct LevelParams {
    const void *src;      // fwd: level input plane;      inv: unused
    void *dst;            // fwd: unused;                  inv: level output plane (W x H)
    void *ll, *hl, *lh, *hh;   // fwd: outputs; inv: inputs.  hl/lh/hh share `sub_pitch`
    int64_t src_pitch, dst_pitch, ll_pitch, sub_pitch;
    int64_t src_frame, dst_frame, ll_frame, sub_frame;   // frame strides
    int W, H;             // size of the full-resolution side of this level
    int nLx, nHx, nLy, nHy;   // ceil/floor halves
    int ncg;              // column groups (one warp wide each)
    int nstrips;          // row strips covered by this launch ...
    int strip0;           // ... starting with this one (the host path pipelines level 0 strip range by strip range)
    int pps;              // output row pairs (fwd) / iterations (inv) per strip
    int sub_aligned;      // 1: hl/hh column offsets allow 16-byte vector access
    int h_room;           // elements from the hl / hh column origin to the end of the pitched plane row
    int bw, nbands;       // ring kernels: column groups per CTA (band), bands per row of CTAs
    int cfg;              // ring kernels: CTA shape (kernels_ring.cu, dispatch_cfg)
    Chain chain;
    int pfd;              // row pairs prefetched ahead in registers: 1 (4 CTAs/SM) or 2 (3 CTAs/SM)
    int dbg;              // measurement only: 1 = no stores, 2 = no lifting arithmetic (forward streaming kernel)
    int narrow;           // 1: 16 bytes per lane instead of 32 (half the registers, twice the warps per SM)
    // interleaved layout (in-place family, ring kernels only): forward writes its four subbands as rows 2k / 2k+1 of `il` (even
    // = L, odd = H in both directions) and LL additionally to `ll`; inverse reads HL, LH, HH from `il` and LL from `ll`
    void *il;
    int64_t il_pitch, il_frame;
    // row strips over several GPUs (strips.cu), forward ring kernels only, one frame: input rows [0, up_end) are read from `src_up`
    // (row r lives at row r + up_row0 there) and rows >= dn_begin (0: none) from `src_dn` (row r - dn_begin + dn_row0) -- the
    // neighbour ranks' planes, peer-mapped -- with the pitch of `src`.  The bulk copies of the ring producer fetch them over NVLink.
    const void *src_up, *src_dn;
    int up_end, dn_begin;
    int64_t up_row0, dn_row0;
};
// the plane and row a forward ring producer reads input row r (already mirrored into [0, H)) from
template <class T> __device__ __forceinline__ const T *level_src_row(const LevelParams &p, const T *local, int r)
{
    if (r < p.up_end) return (const T *)p.src_up + ((int64_t)r + p.up_row0) * p.src_pitch;
    if (p.dn_begin && r >= p.dn_begin) return (const T *)p.src_dn + ((int64_t)(r - p.dn_begin) + p.dn_row0) * p.src_pitch;
    return local + (int64_t)r * p.src_pitch;
}
void launch_fwd_level(int kind, const LevelParams &p, int frames, cudaStream_t st);
void launch_inv_level(int kind, const LevelParams &p, int frames, cudaStream_t st);
// the same level with its input staged through a shared-memory ring by the bulk-copy engine (kernels_ring.cu)
cudaError_t preload_ring();
void launch_fwd_ring(int kind, const LevelParams &p, int frames, int cfg, cudaStream_t st);
void launch_inv_ring(int kind, const LevelParams &p, int frames, int cfg, cudaStream_t st);
// second generation (kernels_ring2.cu, cfg 4): 256-column windows, 8 consumer warps, producer folded into warp 0; inverse for the
// rows-first (float / double) wavelets only
cudaError_t preload_ring2();
void launch_fwd_ring2(int kind, const LevelParams &p, int frames, int cta_warps, cudaStream_t st);
void launch_inv_ring2(int kind, const LevelParams &p, int frames, int cta_warps, cudaStream_t st);
int ring2_cfg_for(int ncg);      // cfg number of the CTA shape for a row of ncg windows: 8, 4, 2 or 1 warps per CTA (16 warps per SM each)
int ring2_cfg_warps(int cfg);
bool ring2_cfg(int cfg);
bool ring2_inverse_ok(int kind);
bool ring2_width_ok(int kind, int W);
int ring2_out_width(int kind);
constexpr int RING_CFG_V2 = 4;
bool ring_interleaved_ok(int kind);   // p.il != nullptr is supported for this kind ...
bool ring_interleaved_cfg_ok(int cfg);   // ... and this CTA shape
int ring_warps_per_sm(int cfg);
int ring_cta_warps(int cfg);
int ring_ctas_per_sm(int cfg);
int stream_out_width(int kind, int narrow);   // output columns per warp
int stream_warps_per_sm(int kind, int narrow, int pfd);
// tile kernels (kernels_tile.cu): same LevelParams (ncg/nstrips/pps/sub_aligned unused), low latency
int tile_rows();   // output rows of the full-resolution side per tile (forward tiles emit tile_rows()/2 LL rows)
dim3 tile_grid_of(int kind, const LevelParams &p, int frames);
void launch_fwd_tile(int kind, const LevelParams &p, int frames, cudaStream_t st);
void launch_inv_tile(int kind, const LevelParams &p, int frames, cudaStream_t st);

// ---- tail kernels: all remaining coarse levels of a plane inside one CTA's shared memory -----------
struct TailParams {
    const void *src;   // fwd: LL input of level j0 (w0 x h0); inv: Mallat plane holding the subbands
    void *dst;         // fwd: Mallat plane receiving subbands + final LL; inv: receives LL of level j0
    int64_t src_pitch, dst_pitch, src_frame, dst_frame;
    int W0, H0;        // full image size (level 0) -> Mallat offsets
    int j0, j1;        // levels j0 .. j1-1 are transformed (fwd ascending, inv descending)
    Chain chain;
};
void launch_fwd_tail(int kind, const TailParams &p, int frames, cudaStream_t st);
void launch_inv_tail(int kind, const TailParams &p, int frames, cudaStream_t st);
int tail_max_elems(int kind);

// ---- interleaved in-place family (kernels_inplace.cu) --------------------------------------------------------
cudaError_t preload_inplace();
// exact evaluation of the reference's prolog / core / epilog sweep order for the rectangle [rx0, rx1) x [ry0, ry1) of a
// level, and optionally a second rectangle [sx0, sx1) x [sy0, sy1), in one launch (CDF 9/7 float): forward reads p.src and
// writes the four subbands, inverse reads them and writes p.dst
void launch_ip_phase(bool inverse, const LevelParams &p, int frames, int rx0, int ry0, int rx1, int ry1, int sx0, int sy0, int sx1, int sy1,
                     cudaStream_t st);
// all levels of a small LL band (w0 x h0 <= ip_tail_cap() samples) in one launch, in place: dense LL on one side, the
// interleaved pyramid of `nlev` levels on the other
int ip_tail_cap();
void launch_ip_tail(bool inverse, void *buf, int64_t pitch, int64_t frame, int w0, int h0, int nlev, int frames, cudaStream_t st);
// Mallat pyramid of J levels -> interleaved layout (unpack: the other way), 4-byte samples, same pitch on both sides; samples of
// the levels >= jt come from / go to the dense tail block instead (tail == nullptr: no tail block)
// (shift = 1: only the samples at even rows and columns, level 0 being interleaved already)
void launch_ip_pack(bool unpack, const void *src, void *dst, int64_t pitch, int64_t frame, int ox, int oy, int J, void *tail, int64_t tpitch,
                    int64_t tframe, int jt, int shift, int frames, cudaStream_t st);

// ---- generic pass kernels: exact reference semantics for sparse (outer != inner) layouts ------
struct PassParams {
    const void *src;
    void *dst;
    int64_t src_pitch, dst_pitch, src_frame, dst_frame;
    int region_w, region_h;   // outer extent of this level (everything inside is rewritten in dst)
    int along_x;              // 1: lines are rows, 0: lines are columns
    int N;                    // inner line length
    int off_h;                // position of the H half along the line
    int keep_dst;             // 1: samples the line function does not write keep dst's content instead of being copied from src
};
void launch_pass_fwd(int kind, const PassParams &p, int frames, cudaStream_t st);
void launch_pass_inv(int kind, const PassParams &p, int frames, cudaStream_t st);

struct ZeroParams {
    void *buf;
    int64_t pitch, frame;
    int region_w, region_h;
    // zero x in [x0a,x0b) U [x1a,x1b) for every row of the region, y likewise for every column
    int x0a, x0b, x1a, x1b, y0a, y0b, y1a, y1b;
};
void launch_zero(int kind, const ZeroParams &p, int frames, cudaStream_t st);

// ---- utilities -------------------------------------------------------------------------------
// test patterns of dwt_util_test_image_fill{,2}_{s,d,i}  (/root/reference/src/libdwt.c:1112-1244)
void launch_fill(int kind, void *buf, int64_t pitch, int64_t frame, int nx, int ny, int rnd, int type,
                 int rnd_frame_mod, int frames, int y_offset, int wide, cudaStream_t st);
// strided element gather/scatter between a byte-addressed staging copy of the caller's layout and a
// dense plane (dwt_util_memcpy_stride_* semantics, src/system.c:90-180)
void launch_repack(int elem_size, void *plane, int64_t pitch_elems, void *staged, int64_t stride_x_bytes,
                   int64_t stride_y_bytes, int nx, int ny, int to_plane, cudaStream_t st);
void launch_copy2d(int elem_size, void *dst, int64_t dpitch, const void *src, int64_t spitch, int w, int h,
                   int64_t dframe, int64_t sframe, int frames, cudaStream_t st);

void launch_compare(int elem_size, const void *a, const void *b, int64_t pitch, int64_t frame, int nx, int ny, int frames,
                    int mode, unsigned long long *out, cudaStream_t st);
void launch_conv_show(int elem_class, const void *src, int64_t sp, void *dst, int64_t dp, int nx, int ny, cudaStream_t st);
void launch_pgm_quant(int elem_class, const void *src, int64_t sp, unsigned char *dst, int nx, int ny, double maxv, double shift, int shifted,
                      cudaStream_t st);
void launch_moments(int elem_class, const void *a, int64_t pitch_elems, int nx, int ny, double *out3, cudaStream_t st);
void launch_volume_fill(float *buf, int64_t pitch, int64_t slice, int nx, int ny, int nz, cudaStream_t st);

// ---- 3-D (one level, interleaved, in place): lifting along one axis --------------------------
// fused 3-D passes (kernels_vol.cu): x+y per slice, then z; strides in elements
struct VolParams {
    const float *src;
    float *dst;
    int64_t s_pitch, s_slice, d_pitch, d_slice;
    int nx, ny, nz;
    int ncg, nstrips, pps;   // filled by the launchers
    int strip0;              // k_vol3 / k_vol3t: first z range of this launch (pipelined host path; 0 otherwise)
};
void launch_vol_xy(VolParams p, int inverse, int sm_count, cudaStream_t st);
void launch_vol_z(VolParams p, int inverse, int sm_count, cudaStream_t st);
// all three axes in ONE pass over the volume (tiles of 64 x 32 positions marching along z); volumes it applies to
bool vol3_applies(const VolParams &p);
// variant 1: tile staged by one tensor copy per slice (k_vol3t); 2: by cp.async (k_vol3, also the fallback without a tensor map)
void launch_vol3(VolParams p, int inverse, int variant, int sm_count, cudaStream_t st);
// the z ranges [strip0, strip0 + count) of the same pass cut into ranges of pps slice pairs (the pipelined host path launches it range by range)
void launch_vol3_ranges(VolParams p, int inverse, int variant, int pps, int strip0, int count, cudaStream_t st);

struct Axis3Params {
    const float *src;
    float *dst;
    int64_t s_line0, s_line1, s_elem;   // src strides (elements): two line-index axes and the lifting axis
    int64_t d_line0, d_line1, d_elem;
    int n0, n1, N;
};
void launch_axis3(const Axis3Params &p, int inverse, cudaStream_t st);

}  // namespace dwtb200
