// dwtb200.cu -- host side of libdwtb200.so: the C ABI of include/dwtb200.h.
//
// This file is the B200 replacement of the reference's level drivers
//   dwt_cdf97_2f_s / 2i_s   /root/reference/src/libdwt.c:12776, 17040
//   dwt_cdf97_2f_d / 2i_d   src/libdwt.c:12451, 16884
//   dwt_cdf53_2f_i / 2i_i   src/libdwt.c:16304, 18142
//   cdf97_3f_{op,ip}_sep_horizontal_s / cdf97_3i_ip_sep_horizontal_s   src/volume-dwt.c:727, 677, 1115
// and of the image allocation / transfer around them (dwt_util_alloc_image src/libdwt.c:1437,
// dwt_util_memcpy_stride_* src/system.c:90-180).
//
// Data layout in HBM (one dwtb200_image = `frames` independent planes):
//   plane[0], plane[1]   two full planes, pitch = width rounded up to 32 elements (128 B / 256 B rows);
//                        a transform reads plane[cur] and writes plane[cur^1] (Mallat layout forbids
//                        in-place tiling: level-j H subbands land where other tiles still read), then cur flips
//   llpool               LL scratch, one band per level (S/3 in total): level j writes LL_j, level j+1 reads it; separate
//                        bands because the kernels of a pyramid overlap (struct Chain, kernels.h)
// Each level of the dense path is ONE kernel launch (kernels_stream.cu) reading its LL input once and
// writing its four subbands once; the coarse levels whose LL band fits one CTA's shared memory are ONE
// launch in total (kernels_tail.cu).  The launch sequence of a call is captured into a CUDA graph and
// cached per image and argument set.  Sparse layouts (outer != inner) and degenerate shapes take the
// generic pass kernels (kernels_generic.cu), two launches per level.
//
// There is no CPU fallback anywhere in this file: without a CUDA device every compute entry point
// returns DWTB200_ENODEV.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <set>
#include <string>
#include <tuple>
#include <algorithm>
#include <vector>

#include <cctype>
#include <sys/mman.h>
#include <sys/syscall.h>
#include <unistd.h>

#include "../../include/dwtb200.h"
#include "internal.h"
#include "kernels.h"
#include "lifting.cuh"

using namespace dwtb200;

namespace dwtb200 {
int g_use_pdl = 0;
}

// ---- global context (the reference keeps process-global state too, src/libdwt.c:478-756) ----
namespace {
struct Ctx {
    int dev = -1;
    cudaStream_t st = nullptr;    // the stream the current call issues on: the image's own stream inside an image call (ImageScope)
    cudaStream_t st0 = nullptr;   // the library stream: volumes, timing events, everything that is not one image's work
    std::set<dwtb200_image *> live;   // images alive: dwtb200_sync / the timer join their streams
    cudaEvent_t stage_ev = nullptr;   // last use of the shared staging buffer
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    char err[512] = "";
    int force_generic = 0;
    int strip_rows = 0;
    int use_graph = 1;
    int launches = 0;   // counts kernel launches issued by the drivers
    void *flush = nullptr;
    size_t flush_bytes = 0;
    void *stage = nullptr;   // device staging for strided repack
    size_t stage_bytes = 0;
    int sm_count = 148;
    // tuning (dwtb200_set_tuning): a level with at most tile_max samples (all frames) takes the tile
    // kernels instead of the streaming ones; the tail kernel starts at the first level with at most
    // tail_max samples per frame
    int64_t tile_max = (int64_t)1024 * 1024;
    int tail_max = 64 * 64;
    int narrow = 0;
    int dbg = 0;
    int pfd = 1;
    int pipeline = 1;   // pipelined host path for large dense images
    int ring = 3;       // bit 0 / 1: forward / inverse streaming levels take the bulk-copy ring kernels (kernels_ring.cu)
    int ring_v2 = 1;    // the ring levels take the 256-column kernels (kernels_ring2.cu) where they apply; scoped off for the interleaved layout
    int ring_waves = 3, ring_pps_min = 0, ring_pps_max = 0;   // strip-length search range of the ring kernels (0: defaults)
    int vol3 = 1;       // 3-D: one pass over the volume where the tile kernel applies (kernels_vol.cu: 1 = k_vol3t, 2 = k_vol3, 0 = two passes)
    int chain = 1;      // kernels of a pyramid overlap through completion counters (struct Chain): bit 0 ring levels, bit 1 tile / tail
    int epoch = 0;   // bumped by every tuning change: part of the graph cache key
} g;

// Threading contract (include/dwtb200.h): every extern "C" entry point takes this process-wide lock, so concurrent callers
// (an OpenMP loop over images around dwt_cdf97_2f_s, say) are serialised instead of corrupting the shared context --
// g.st, g.err, the staging buffer, the cached device mirrors of the *_host calls.  Recursive: entry points call each other.
std::recursive_mutex g_api_mutex;
#define API_LOCK() std::lock_guard<std::recursive_mutex> api_lock_(g_api_mutex)

int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g.err, sizeof g.err, fmt, ap);
    va_end(ap);
    return code;
}
#define CK(call)                                                                                      \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess)                                                                        \
            return fail(DWTB200_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

inline int cdiv_pow2(int v, int j) { return (int)(((int64_t)v + ((int64_t)1 << j) - 1) >> j); }
inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }
inline size_t esize(int kind) { return (size_t)kind_elem_size(kind); }
inline bool guard(int kind) { return kind == DWTB200_CDF97_F32; }            // src/libdwt.c:12837 (no other driver tests `lines > 1`)
inline bool inv_cols_first(int kind) { return kind == DWTB200_CDF53_I32 || kind == DWTB200_CDF97_I32; }   // src/libdwt.c:18178, 18256
inline bool kind_ok(int kind) { return kind >= 0 && kind < DWTB200_KIND_COUNT; }

int ensure_stage(size_t bytes)
{
    if (bytes <= g.stage_bytes) return 0;
    if (g.stage_ev) cudaEventSynchronize(g.stage_ev);   // the old buffer may still be in use on another image's stream
    if (g.stage) cudaFree(g.stage);
    g.stage = nullptr;
    g.stage_bytes = 0;
    CK(cudaMalloc(&g.stage, bytes));
    g.stage_bytes = bytes;
    return 0;
}
}  // namespace

struct dwtb200_image {
    cudaStream_t st = nullptr;   // independent images overlap: the small levels of one run in the shadow of another's big ones
    cudaEvent_t ev = nullptr;
    cudaEvent_t t0 = nullptr, t1 = nullptr;   // dwtb200_image_timer_*: events on the image's own stream
    // row strips (strips.cu): where level 0 of a forward transform reads the rows it does not own, per plane (see LevelParams)
    const void *up[2] = {nullptr, nullptr}, *dn[2] = {nullptr, nullptr};
    int up_end = 0, dn_begin = 0;
    int64_t up_row0 = 0, dn_row0 = 0;
    std::vector<cudaEvent_t> marks;           // dwtb200_image_timer_mark
    size_t nmarks = 0;
    int kind = 0, ox = 0, oy = 0, frames = 0;
    size_t es = 4;
    int64_t pitch = 0, frame = 0;   // elements
    void *plane[2] = {nullptr, nullptr};
    int cur = 0;
    void *llpool = nullptr;           // LL_j bands of the levels j = 0 .. (one buffer per level: the kernels of a pyramid overlap)
    int64_t ll_off[40] = {0};         // element offset of LL_j inside llpool
    uint32_t *sync = nullptr;         // chain counters of un-captured launch sequences (captured graphs own theirs)
    size_t sync_words = 0;
    int last_launches = 0, last_path = 0;
    int kind_class() const { return dwtb200::kind_elem_class(kind); }
    typedef std::tuple<int, int, int, int, int, int, int, int, int> Key;
    struct Entry {
        cudaGraphExec_t exec;
        int launches, path, flips;
        uint32_t *sync;   // generation word + completion counters of this graph's kernel chain
    };
    std::map<Key, Entry> graphs;
};

struct dwtb200_volume {
    int nx = 0, ny = 0, nz = 0;
    int64_t pitch = 0, slice = 0;   // elements
    float *buf[2] = {nullptr, nullptr};
    int cur = 0;
};

namespace {
// every call that works on one image issues on that image's stream
struct ImageScope {
    cudaStream_t prev;
    explicit ImageScope(const dwtb200_image *im) : prev(g.st)
    {
        if (im && im->st) g.st = im->st;
    }
    ~ImageScope() { g.st = prev; }
};
// make stream `waiter` wait for everything queued so far on the stream of `im`
void wait_for_image(cudaStream_t waiter, dwtb200_image *im)
{
    if (!im || !im->st || im->st == waiter) return;
    cudaEventRecord(im->ev, im->st);
    cudaStreamWaitEvent(waiter, im->ev, 0);
}
// the shared staging buffer is used by one copy at a time, whatever stream it runs on
void stage_acquire()
{
    if (g.stage_ev) cudaStreamWaitEvent(g.st, g.stage_ev, 0);
}
void stage_release()
{
    if (!g.stage_ev) cudaEventCreateWithFlags(&g.stage_ev, cudaEventDisableTiming);
    cudaEventRecord(g.stage_ev, g.st);
}
}  // namespace

namespace dwtb200 {
ImageView image_view(const dwtb200_image *im)
{
    ImageView v;
    v.plane[0] = im->plane[0];
    v.plane[1] = im->plane[1];
    v.cur = im->cur;
    v.pitch = im->pitch;
    v.frame = im->frame;
    v.es = im->es;
    v.ox = im->ox;
    v.oy = im->oy;
    v.frames = im->frames;
    v.kind = im->kind;
    v.st = im->st;
    return v;
}
int set_error(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g.err, sizeof g.err, fmt, ap);
    va_end(ap);
    return code;
}
std::recursive_mutex &api_mutex() { return g_api_mutex; }
}  // namespace dwtb200
namespace dwtb200 {
bool image_level0_is_ring(dwtb200_image *im, int J);   // defined below, next to the planner
void image_set_row_sources(dwtb200_image *im, const void *const up[2], const void *const dn[2], int up_end, int dn_begin, int64_t up_row0, int64_t dn_row0)
{
    for (int i = 0; i < 2; i++) {
        im->up[i] = up ? up[i] : nullptr;
        im->dn[i] = dn ? dn[i] : nullptr;
    }
    im->up_end = up_end;
    im->dn_begin = dn_begin;
    im->up_row0 = up_row0;
    im->dn_row0 = dn_row0;
    // the cached graphs hold the kernel parameters of the old sources
    if (im->st) cudaStreamSynchronize(im->st);
    for (auto &kv : im->graphs) {
        cudaGraphExecDestroy(kv.second.exec);
        if (kv.second.sync) cudaFree(kv.second.sync);
    }
    im->graphs.clear();
}
}  // namespace dwtb200

extern "C" {

// =====================================================================================================
// lifecycle
// =====================================================================================================
int dwtb200_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int dwtb200_init(int device)
{
    API_LOCK();
    if (g.dev >= 0) return DWTB200_OK;
    const int n = dwtb200_device_count();
    if (n <= 0) return fail(DWTB200_ENODEV, "no CUDA device (libdwtb200 has no CPU fallback)");
    if (device < 0) {
        const char *e = getenv("DWT_DEVICE");
        if (!e) e = getenv("LOCAL_RANK");
        device = e ? atoi(e) % n : 0;
    }
    if (device >= n) return fail(DWTB200_EINVAL, "device %d out of range (%d devices)", device, n);
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(DWTB200_ENODEV, "device %d is sm_%d%d; libdwtb200 is built for sm_100a only", device, prop.major, prop.minor);
    g.sm_count = prop.multiProcessorCount;
    CK(preload_stream());
    CK(preload_ring());
    CK(preload_ring2());
    CK(preload_inplace());
    CK(preload_tail());
    CK(preload_generic());
    CK(preload_util());
    CK(preload_tile());
    CK(cudaStreamCreateWithFlags(&g.st0, cudaStreamNonBlocking));
    g.st = g.st0;
    CK(cudaEventCreate(&g.e0));
    CK(cudaEventCreate(&g.e1));
    const char *ng = getenv("DWTB200_NO_GRAPH");
    g.use_graph = !(ng && atoi(ng));
    g.dev = device;
    return DWTB200_OK;
}

void dwtb200_release_host_cache(void);
void dwtb200_finish(void)
{
    API_LOCK();
    if (g.dev < 0) return;
    cudaDeviceSynchronize();
    dwtb200_release_host_cache();
    if (g.stage_ev) cudaEventDestroy(g.stage_ev);
    g.stage_ev = nullptr;
    if (g.flush) cudaFree(g.flush);
    if (g.stage) cudaFree(g.stage);
    g.flush = g.stage = nullptr;
    g.flush_bytes = g.stage_bytes = 0;
    cudaEventDestroy(g.e0);
    cudaEventDestroy(g.e1);
    cudaStreamDestroy(g.st0);
    g.st = g.st0 = nullptr;
    g.dev = -1;
}

const char *dwtb200_last_error(void) { return g.err; }
int dwtb200_device(void) { return g.dev; }

#define NEED_DEV()                                       \
    do {                                                 \
        if (g.dev < 0) {                                 \
            const int r_ = dwtb200_init(-1);             \
            if (r_) return r_;                           \
        }                                                \
    } while (0)

// Page-locked host memory on the NUMA node of the GPU.  cudaHostAlloc places the pages wherever the calling thread happens to run;
// with one process per GPU on a two-socket host half the ranks then move every byte across the socket interconnect and the eight
// ranks' copies pile up on one memory controller (round 1: aggregate H2D + D2H saturated at 100-130 GB/s whatever the number of
// GPUs).  Here the block is mmap'ed, bound to the node the GPU's PCIe root port hangs off (sysfs numa_node, mbind MPOL_PREFERRED)
// and then registered with the driver.  DWTB200_NUMA=0, a single-node host or a refused mbind fall back to cudaHostAlloc.
namespace {
std::map<void *, size_t> g_numa_blocks;   // blocks obtained through mmap + cudaHostRegister
int gpu_numa_node()
{
    static int node = -2;
    if (node != -2) return node;
    node = -1;
    const char *e = getenv("DWTB200_NUMA");
    if (e && !atoi(e)) return node;
    char bdf[32] = "";
    if (cudaDeviceGetPCIBusId(bdf, sizeof bdf, g.dev) != cudaSuccess) {
        cudaGetLastError();
        return node;
    }
    for (char *c = bdf; *c; c++) *c = (char)tolower(*c);
    char path[128];
    snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/numa_node", bdf);
    if (FILE *f = fopen(path, "r")) {
        int n = -1;
        if (fscanf(f, "%d", &n) == 1) node = n;
        fclose(f);
    }
    if (node >= 0 && access("/sys/devices/system/node/node1", F_OK) != 0) node = -1;   // one node: nothing to choose
    return node;
}
}  // namespace

int dwtb200_host_numa_node(void)
{
    API_LOCK();
    if (g.dev < 0 && dwtb200_init(-1)) return -1;
    return gpu_numa_node();
}

void *dwtb200_host_alloc(size_t bytes)
{
    API_LOCK();
    if (g.dev < 0 && dwtb200_init(-1)) return nullptr;
    if (!bytes) bytes = 1;
    const int node = gpu_numa_node();
    if (node >= 0 && node < 1024) {
        const size_t len = (bytes + 4095) & ~(size_t)4095;
        void *p = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        if (p != MAP_FAILED) {
            unsigned long mask[16] = {0};
            mask[node / 64] = 1ul << (node % 64);
            const long rc = syscall(SYS_mbind, p, len, 1 /* MPOL_PREFERRED */, mask, 1025ul, 0u);
            if (rc == 0 && cudaHostRegister(p, len, cudaHostRegisterPortable) == cudaSuccess) {
                g_numa_blocks[p] = len;
                return p;
            }
            cudaGetLastError();
            munmap(p, len);
        }
    }
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) {
        fail(DWTB200_ENOMEM, "cudaHostAlloc(%zu) failed: %s", bytes, cudaGetErrorString(cudaGetLastError()));
        return nullptr;
    }
    return p;
}
void dwtb200_host_free(void *ptr)
{
    API_LOCK();
    if (!ptr) return;
    auto it = g_numa_blocks.find(ptr);
    if (it != g_numa_blocks.end()) {
        cudaHostUnregister(ptr);
        munmap(ptr, it->second);
        g_numa_blocks.erase(it);
        return;
    }
    // the compat layer interposes dwt_util_free_image process-wide: a pointer that did not come from dwtb200_host_alloc
    // (memalign in the reference's own allocator) goes back to free(), and no CUDA error is left pending
    if (cudaFreeHost(ptr) != cudaSuccess) {
        cudaGetLastError();
        free(ptr);
    }
}

int dwtb200_ceil_log2(int x)
{
    int j = 0;
    while (((int64_t)1 << j) < x) j++;
    return j;
}
int dwtb200_clamp_j(int j_max, int ox, int oy, int decompose_one)
{
    const int omin = ox < oy ? ox : oy, omax = ox < oy ? oy : ox;
    const int lim = dwtb200_ceil_log2(decompose_one ? omax : omin);
    return (j_max < 0 || j_max > lim) ? lim : j_max;
}

void dwtb200_force_generic(int on) { g.force_generic = on; }
void dwtb200_set_strip_rows(int rows) { g.strip_rows = rows; }
int dwtb200_set_tuning(int key, long long value)
{
    API_LOCK();
    switch (key) {
    case DWTB200_TUNE_TILE_MAX: g.tile_max = value; break;
    case DWTB200_TUNE_MID_MAX: break;   // the persistent mid-level kernels were removed in round 2 (never faster than one launch per level)
    case DWTB200_TUNE_PIPELINE: g.pipeline = value != 0; break;
    case DWTB200_TUNE_RING: g.ring = (int)value; break;
    case DWTB200_TUNE_CHAIN: g.chain = (int)value; break;
    case DWTB200_TUNE_PYR: break;   // the fused tile-pyramid kernels were removed in round 2 (never faster than one tile launch per level)
    case DWTB200_TUNE_VOL3: g.vol3 = value < 0 || value > 2 ? 1 : value; break;
#ifdef DWTB200_DEBUG_KEYS   // measurement-only knobs (profiles/scripts): not part of the release ABI
    case 97: g.ring_waves = (int)(value & 0xff); g.ring_pps_min = (int)((value >> 8) & 0xff); g.ring_pps_max = (int)((value >> 16) & 0xfff); break;
    case 98: g.pfd = (int)value; break;
    case 99: g.dbg = (int)value; break;   // see kernels.h
#endif
    case DWTB200_TUNE_NARROW: g.narrow = value != 0; break;
    case DWTB200_TUNE_PDL: dwtb200::g_use_pdl = value != 0; break;
    case DWTB200_TUNE_TAIL_MAX:
        if (value < 0 || value > tail_max_elems(DWTB200_CDF97_F32)) return fail(DWTB200_EINVAL, "tail_max out of range");
        g.tail_max = (int)value;
        break;
    default: return fail(DWTB200_EINVAL, "unknown tuning key %d", key);
    }
    g.epoch++;
    return DWTB200_OK;
}

// =====================================================================================================
// images
// =====================================================================================================
dwtb200_image *dwtb200_image_create(int kind, int ox, int oy, int frames)
{
    API_LOCK();
    if (g.dev < 0 && dwtb200_init(-1)) return nullptr;
    if (!kind_ok(kind) || ox < 1 || oy < 1 || frames < 1) {
        fail(DWTB200_EINVAL, "image_create: bad arguments");
        return nullptr;
    }
    dwtb200_image *im = new dwtb200_image;
    if (cudaStreamCreateWithFlags(&im->st, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&im->ev, cudaEventDisableTiming) != cudaSuccess) {
        fail(DWTB200_ECUDA, "image_create: stream: %s", cudaGetErrorString(cudaGetLastError()));
        delete im;
        return nullptr;
    }
    g.live.insert(im);
    ImageScope scope(im);
    im->kind = kind;
    im->ox = ox;
    im->oy = oy;
    im->frames = frames;
    im->es = esize(kind);
    im->pitch = align_up(ox, 32);
    im->frame = im->pitch * oy;
    const size_t plane_bytes = (size_t)im->frame * frames * im->es;
    int64_t pool = 0;
    for (int j = 0; j < 40; j++) {   // LL_j = (w_{j+1} x h_{j+1}) per frame, 128-byte aligned
        im->ll_off[j] = pool;
        const int w = cdiv_pow2(ox, j + 1), h = cdiv_pow2(oy, j + 1);
        if (w * (int64_t)h > 1 || j == 0) pool += align_up(align_up(w, 32) * (int64_t)h * frames, 32);
    }
    // chain counters: at most one per 8 output rows of every level, per frame
    im->sync_words = 2 + (size_t)frames * ((size_t)oy / 4 + 4 * 40 + 64);
    bool ok = cudaMalloc(&im->plane[0], plane_bytes) == cudaSuccess && cudaMalloc(&im->plane[1], plane_bytes) == cudaSuccess &&
              cudaMalloc(&im->llpool, (size_t)pool * im->es) == cudaSuccess &&
              cudaMalloc((void **)&im->sync, im->sync_words * sizeof(uint32_t)) == cudaSuccess;
    if (ok) ok = cudaMemsetAsync(im->plane[0], 0, plane_bytes, g.st) == cudaSuccess &&
                 cudaMemsetAsync(im->plane[1], 0, plane_bytes, g.st) == cudaSuccess;
    if (!ok) {
        fail(DWTB200_ENOMEM, "image_create(%d x %d x %d): %s", ox, oy, frames, cudaGetErrorString(cudaGetLastError()));
        dwtb200_image_destroy(im);
        return nullptr;
    }
    return im;
}

void dwtb200_image_destroy(dwtb200_image *im)
{
    API_LOCK();
    if (!im) return;
    if (im->st) cudaStreamSynchronize(im->st);
    g.live.erase(im);
    if (im->ev) cudaEventDestroy(im->ev);
    if (im->t0) cudaEventDestroy(im->t0);
    if (im->t1) cudaEventDestroy(im->t1);
    for (cudaEvent_t e : im->marks) cudaEventDestroy(e);
    if (im->st) cudaStreamDestroy(im->st);
    for (auto &kv : im->graphs) {
        cudaGraphExecDestroy(kv.second.exec);
        if (kv.second.sync) cudaFree(kv.second.sync);
    }
    for (int i = 0; i < 2; i++)
        if (im->plane[i]) cudaFree(im->plane[i]);
    if (im->llpool) cudaFree(im->llpool);
    if (im->sync) cudaFree(im->sync);
    delete im;
}

static inline char *frame_ptr(dwtb200_image *im, int which, int frame)
{
    return (char *)im->plane[which] + (size_t)frame * im->frame * im->es;
}

// host span touched by an (ox x oy) image with byte strides sx (rows) / sy (columns)
static inline size_t host_span(int ox, int oy, int64_t sx, int64_t sy, size_t es)
{
    return (size_t)((int64_t)(oy - 1) * sx + (int64_t)(ox - 1) * sy) + es;
}

// the top-left w x h samples of a host image -> the current plane of `frame`
static int upload_region(dwtb200_image *im, int frame, const void *host, int64_t sx, int64_t sy, int w, int h)
{
    if (w <= 0 || h <= 0) return DWTB200_OK;
    char *d = frame_ptr(im, im->cur, frame);
    if (sy == (int64_t)im->es && sx >= (int64_t)(w * im->es)) {
        CK(cudaMemcpy2DAsync(d, im->pitch * im->es, host, (size_t)sx, w * im->es, h, cudaMemcpyDefault, g.st));
    } else {
        const size_t span = host_span(w, h, sx, sy, im->es);
        int r = ensure_stage(span);
        if (r) return r;
        stage_acquire();
        CK(cudaMemcpyAsync(g.stage, host, span, cudaMemcpyDefault, g.st));
        launch_repack((int)im->es, d, im->pitch, g.stage, sx, sy, w, h, 1, g.st);
        stage_release();
        CK(cudaGetLastError());
    }
    return DWTB200_OK;
}

int dwtb200_image_upload(dwtb200_image *im, int frame, const void *host, int64_t sx, int64_t sy)
{
    API_LOCK();
    ImageScope scope(im);
    NEED_DEV();
    if (!im || !host || frame < 0 || frame >= im->frames || sx <= 0 || sy <= 0) return fail(DWTB200_EINVAL, "image_upload: bad arguments");
    return upload_region(im, frame, host, sx, sy, im->ox, im->oy);
}

int dwtb200_image_download(dwtb200_image *im, int frame, void *host, int64_t sx, int64_t sy)
{
    API_LOCK();
    ImageScope scope(im);
    NEED_DEV();
    if (!im || !host || frame < 0 || frame >= im->frames || sx <= 0 || sy <= 0) return fail(DWTB200_EINVAL, "image_download: bad arguments");
    char *d = frame_ptr(im, im->cur, frame);
    if (sy == (int64_t)im->es && sx >= (int64_t)(im->ox * im->es)) {
        CK(cudaMemcpy2DAsync(host, (size_t)sx, d, im->pitch * im->es, im->ox * im->es, im->oy, cudaMemcpyDefault, g.st));
    } else {
        // bytes of the caller's buffer that do not belong to this image (other channels, padding) must
        // survive: stage the whole span, scatter the samples into it, copy the span back
        const size_t span = host_span(im->ox, im->oy, sx, sy, im->es);
        int r = ensure_stage(span);
        if (r) return r;
        stage_acquire();
        CK(cudaMemcpyAsync(g.stage, host, span, cudaMemcpyDefault, g.st));
        launch_repack((int)im->es, d, im->pitch, g.stage, sx, sy, im->ox, im->oy, 0, g.st);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(host, g.stage, span, cudaMemcpyDefault, g.st));
        stage_release();
    }
    CK(cudaStreamSynchronize(g.st));
    return DWTB200_OK;
}

int dwtb200_image_fill_ex(dwtb200_image *im, int rnd, int type, int rand_mod, int y_offset, int wide)
{
    API_LOCK();
    ImageScope scope(im);
    NEED_DEV();
    if (!im) return fail(DWTB200_EINVAL, "image_fill: null image");
    launch_fill(im->kind, im->plane[im->cur], im->pitch, im->frame, im->ox, im->oy, rnd, type, rand_mod, im->frames, y_offset,
                wide, g.st);
    CK(cudaGetLastError());
    return DWTB200_OK;
}
int dwtb200_image_fill(dwtb200_image *im, int rnd, int type, int rand_mod) { return dwtb200_image_fill_ex(im, rnd, type, rand_mod, 0, 0); }

// rows [row0, row0+rows) of one frame of the current plane <-> a dense device or host buffer (cudaMemcpyDefault):
// the halo rows of a row-strip partition travel through this (peer-mapped pointers included)
int dwtb200_image_copy_rows(dwtb200_image *im, int frame, int row0, int rows, void *buf, int64_t buf_pitch_bytes, int to_image)
{
    API_LOCK();
    ImageScope scope(im);
    NEED_DEV();
    if (!im || !buf || frame < 0 || frame >= im->frames || row0 < 0 || rows < 0 || row0 + rows > im->oy)
        return fail(DWTB200_EINVAL, "image_copy_rows: bad arguments");
    if (rows == 0) return DWTB200_OK;
    char *d = (char *)im->plane[im->cur] + ((size_t)frame * im->frame + (size_t)row0 * im->pitch) * im->es;
    if (to_image) CK(cudaMemcpy2DAsync(d, im->pitch * im->es, buf, (size_t)buf_pitch_bytes, im->ox * im->es, rows, cudaMemcpyDefault, g.st));
    else CK(cudaMemcpy2DAsync(buf, (size_t)buf_pitch_bytes, d, im->pitch * im->es, im->ox * im->es, rows, cudaMemcpyDefault, g.st));
    return DWTB200_OK;
}

// CUDA IPC: let another process on the same node map this image's current plane (NVLink P2P between ranks)
int dwtb200_image_ipc_export(dwtb200_image *im, void *handle64)
{
    API_LOCK();
    NEED_DEV();
    if (!im || !handle64) return fail(DWTB200_EINVAL, "ipc_export: null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    CK(cudaIpcGetMemHandle((cudaIpcMemHandle_t *)handle64, im->plane[im->cur]));
    return DWTB200_OK;
}
void *dwtb200_ipc_open(const void *handle64)
{
    API_LOCK();
    if (g.dev < 0 && dwtb200_init(-1)) return nullptr;
    void *p = nullptr;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof h);
    if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        fail(DWTB200_ECUDA, "cudaIpcOpenMemHandle: %s", cudaGetErrorString(cudaGetLastError()));
        return nullptr;
    }
    return p;
}
int dwtb200_ipc_close(void *ptr)
{
    API_LOCK();
    NEED_DEV();
    CK(cudaIpcCloseMemHandle(ptr));
    return DWTB200_OK;
}

void *dwtb200_image_devptr(dwtb200_image *im, size_t *pitch_bytes, size_t *frame_bytes)
{
    API_LOCK();
    if (!im) return nullptr;
    if (pitch_bytes) *pitch_bytes = (size_t)im->pitch * im->es;
    if (frame_bytes) *frame_bytes = (size_t)im->frame * im->es;
    return im->plane[im->cur];
}

int dwtb200_image_copy(dwtb200_image *dst, dwtb200_image *src)
{
    API_LOCK();
    NEED_DEV();
    if (!dst || !src || dst->kind != src->kind || dst->ox != src->ox || dst->oy != src->oy || dst->frames != src->frames)
        return fail(DWTB200_EINVAL, "image_copy: shape mismatch");
    ImageScope scope(dst);
    wait_for_image(g.st, src);   // the copy reads src after everything queued on it ...
    CK(cudaMemcpyAsync(dst->plane[dst->cur], src->plane[src->cur], (size_t)src->frame * src->frames * src->es,
                       cudaMemcpyDeviceToDevice, g.st));
    wait_for_image(src->st, dst);   // ... and later work on src must not overwrite it before the copy has read it
    return DWTB200_OK;
}

int dwtb200_image_last_launches(dwtb200_image *im) { return im ? im->last_launches : 0; }
int dwtb200_image_last_path(dwtb200_image *im) { return im ? im->last_path : 0; }

// ---- level plans -------------------------------------------------------------------------------
namespace {

struct Band {   // where an LL band lives
    void *p;
    int64_t pitch, frame;
};

// ---- kernel selection for the dense path -----------------------------------------------------------
// Level j (input w_j x h_j, all frames) is handled by
//   STREAM  kernels_stream.cu   big, HBM-bound levels: one launch per level
//   TILE    kernels_tile.cu     one launch per level, small tiles
//   TAIL    kernels_tail.cu     every remaining level once the LL band fits one CTA's shared memory
enum { PLAN_STREAM = 0, PLAN_TILE = 1 };
struct DensePlan {
    int jt;                 // first level of the tail (== J: no tail); -1: the dense kernels cannot take this pyramid
    int type[40];
};

DensePlan dense_plan(const dwtb200_image *im, int J)
{
    DensePlan pl;
    int cap = tail_max_elems(im->kind);
    if (g.tail_max < cap) cap = g.tail_max;
    pl.jt = J;
    for (int j = 0; j < J; j++) {
        const int w = cdiv_pow2(im->ox, j), h = cdiv_pow2(im->oy, j);
        if ((int64_t)w * h <= cap) {
            pl.jt = j;
            break;
        }
        if (w < 2 || h < 2) {
            pl.jt = -1;
            return pl;
        }
    }
    for (int j = 0; j < pl.jt; j++) {
        const int64_t n = (int64_t)cdiv_pow2(im->ox, j) * cdiv_pow2(im->oy, j) * im->frames;
        pl.type[j] = n <= g.tile_max ? PLAN_TILE : PLAN_STREAM;
    }
    return pl;
}

Band ll_band(const dwtb200_image *im, int j)   // LL_j = output of level j, (w_{j+1} x h_{j+1})
{
    const int w = cdiv_pow2(im->ox, j + 1), h = cdiv_pow2(im->oy, j + 1);
    Band b;
    b.p = (char *)im->llpool + (size_t)im->ll_off[j] * im->es;
    b.pitch = align_up(w, 32);
    b.frame = b.pitch * h;
    return b;
}

void level_geometry(const dwtb200_image *im, int j, bool inverse, LevelParams &p)
{
    const int W = cdiv_pow2(im->ox, j), H = cdiv_pow2(im->oy, j);
    p.W = W;
    p.H = H;
    p.nLx = (W + 1) >> 1;
    p.nHx = W >> 1;
    p.nLy = (H + 1) >> 1;
    p.nHy = H >> 1;
    p.narrow = g.narrow;
    p.dbg = g.dbg;
    p.pfd = inverse ? 1 : g.pfd;
    // second-generation ring kernels (256-column warp windows): every forward level, inverse levels of the rows-first wavelets
    const int vec = im->es == 8 ? 2 : (p.narrow ? 2 : 4);   // elements per subband store of a lane
    p.sub_aligned = (p.nLx % vec) == 0;
    // (an inverse level whose HL / HH column origin is not 16-byte aligned falls back to the register kernels and their geometry)
    const bool ring_here = !p.narrow && (inverse ? ((g.ring & 2) && p.sub_aligned) : (g.ring & 1));
    int forced = (g.ring >> 4) & 7;   // DWTB200_TUNE_RING bits 4-6 force a shape
    if (forced > 5) forced = 0;
    const bool v2 = ring_here && g.ring_v2 && (forced == 0 || forced == RING_CFG_V2) && (!inverse || ring2_inverse_ok(im->kind)) &&
                    ring2_width_ok(im->kind, W);
    const int outw = v2 ? ring2_out_width(im->kind) : stream_out_width(im->kind, p.narrow);
    p.ncg = (W + outw - 1) / outw;
    if (v2) {
        p.cfg = ring2_cfg_for(p.ncg);   // 8 warps per CTA, or 4 / 2 / 1 for rows of at most 4 / 2 / 1 windows
        const int cw = ring_cta_warps(p.cfg), nb = (p.ncg + cw - 1) / cw;
        p.bw = (p.ncg + nb - 1) / nb;
        p.nbands = (p.ncg + p.bw - 1) / p.bw;
    } else
    {   // ring kernels: bands of at most ring_cta_warps() column groups, as equal as possible
        // CTA shape: 7 consumer warps x 2 CTAs per SM, unless the row splits into bands of exactly 5 column groups (2048-
        // or 1080p-wide frames) AND the batch is large enough for >= 2 waves of the smaller CTAs: then 5 x 3 keeps 15 instead
        // of 10 consumer warps per SM busy (2048^2 x 64: 419 -> 445 Gpixel/s; a 4-frame level of 2048^2 is better off with 7 x 2)
        p.cfg = forced == RING_CFG_V2 ? 0 : forced;
        if (p.cfg == 0) {
            const int nb7 = (p.ncg + 6) / 7, bw7 = (p.ncg + nb7 - 1) / nb7;   // bands of the default shape: 5 of 7 warps busy?
            const int nb5 = (p.ncg + 4) / 5, bw5 = (p.ncg + nb5 - 1) / nb5;
            const int64_t est = (int64_t)nb5 * ((inverse ? (H >> 1) + 1 : p.nLy) / 28 + 1) * im->frames;
            if (bw7 <= 5 && bw5 == 5 && est >= 2 * (int64_t)g.sm_count * ring_ctas_per_sm(3)) p.cfg = 3;
        }
        const int cw = ring_cta_warps(p.cfg), nb = (p.ncg + cw - 1) / cw;
        p.bw = (p.ncg + nb - 1) / nb;
        p.nbands = (p.ncg + p.bw - 1) / p.bw;
    }
    const int units = inverse ? (H >> 1) + 1 : p.nLy;   // row pairs to emit
    int pps;
    if (g.strip_rows > 0) {
        pps = g.strip_rows / 2 > 0 ? g.strip_rows / 2 : 1;
    } else {
        // enough warps for ~16 per SM, but strips of at least 8 and at most 64 pairs (warm-up rows are
        // re-read per strip: 3 pairs for 9/7 forward, 4 for inverse)
        const bool ring = !p.narrow && (inverse ? (g.ring & 2) : (g.ring & 1));
        if (ring) {
            // Measured (profiles/ring_pps_r1.txt, pps_single_r1.txt): CTAs run at visibly different speeds, so many short
            // CTAs handed out dynamically beat one long CTA per slot in spite of the warm-up rows every strip re-reads;
            // and the last wave matters: 3.85 or 2.9 waves of CTAs run 5-10 % faster than 4.3 or 2.2.  Pick the strip
            // length with the lowest modelled time: (re-read overhead) / (occupancy of the last wave) + imbalance.
            const int cfg = p.cfg;
            const int64_t slots = (int64_t)g.sm_count * ring_ctas_per_sm(cfg);
            const int ns = kind_lifting_steps(im->kind), warm = inverse ? ns : ns / 2 + (ns == 4 ? 1 : 0);
            int lo = g.ring_pps_min > 0 ? g.ring_pps_min : (ns == 4 ? 12 : 8), hi = g.ring_pps_max > 0 ? g.ring_pps_max : 40;
            // Level 0 is not fed by a predecessor through the chain, its own ramp and ragged end are paid in full: measured on one
            // 8192^2 float image (profiles/single_r2_pps.txt), strips of 14 / 18 / 24 pairs (4.95 / 3.85 / 2.9 waves) take 100.3 /
            // 100.4 / 101.7 us, the 35 pairs (2.0 waves) the model below used to pick 109.6 us, 16 / 22 / 30 pairs (4.3 / 3.2 / 2.3
            // waves) 105 - 108 us.  Fitted: the re-read of the warm-up rows costs a quarter of its bytes (reads only, half of them
            // L2 hits), an incomplete last wave 30 % of its idle slots (the running CTAs get the bandwidth), and few waves a ragged
            // end of 0.36 / waves^2.  The levels behind level 0 run at the pace of their predecessor and keep the old model.
            const bool head = j == 0 && g.ring_waves != 99;
            double best = 1e30;
            pps = lo;
            // a level that cannot fill one wave of CTAs even with the shortest regular strips is latency-bound: its time is the
            // (warm-up + pps) iterations one CTA walks through, so it gets the shortest strips that still fit one wave
            if (!head && g.ring_pps_min == 0 && (int64_t)p.nbands * ((units + lo - 1) / lo) * im->frames <= slots) {
                int c = lo;
                while (c > 4 && (int64_t)p.nbands * ((units + (c - 1) - 1) / (c - 1)) * im->frames <= slots) c--;
                lo = hi = c;
            }
            for (int c = lo; c <= hi; c++) {
                const int64_t n = (int64_t)p.nbands * ((units + c - 1) / c) * im->frames;
                const int64_t waves = (n + slots - 1) / slots;
                const double eff = (double)n / (double)(waves * slots);
                const double wv = (double)n / (double)slots;
                const double cost = head ? (1.0 + 0.25 * warm / c) * (1.0 + 0.3 * (1.0 / eff - 1.0)) * (1.0 + 0.36 / (wv * wv))
                                         : (1.0 + 0.5 * warm / c) / eff + 0.002 * c;
                if (cost < best - 1e-9) {
                    best = cost;
                    pps = c;
                }
            }
        } else {
            const int64_t want = (int64_t)g.sm_count * stream_warps_per_sm(im->kind, p.narrow, p.pfd);
            int64_t per_col = want / ((int64_t)p.ncg * im->frames);   // strips per column group: never more warps than fit at once
            if (per_col < 1) per_col = 1;
            pps = (int)((units + per_col - 1) / per_col);
            if (pps < 8) pps = 8;
            if (pps > 64) pps = 64;
        }
    }
    p.pps = pps;
    p.nstrips = (units + pps - 1) / pps;
}

}  // namespace
// true when level 0 of a dense forward transform of `im` is served by a bulk-copy ring kernel (the kernels that can take rows from
// neighbour planes, LevelParams::src_up / src_dn)
extern "C++" bool dwtb200::image_level0_is_ring(dwtb200_image *im, int J)
{
    if (J < 1 || g.force_generic || !(g.ring & 1) || g.narrow || !g.use_graph) return false;
    const DensePlan pl = dense_plan(im, J);
    return pl.jt > 0 && pl.type[0] == PLAN_STREAM;
}
namespace {
// parameters of forward level j reading `in` (LL_{j-1} or the source plane); returns where LL_j goes
Band fwd_level_params(const dwtb200_image *im, int j, int J, const Band &in, char *dst_plane, LevelParams &p)
{
    memset(&p, 0, sizeof p);
    level_geometry(im, j, false, p);
    const Band out = (j == J - 1) ? Band{dst_plane, im->pitch, im->frame} : ll_band(im, j);
    p.src = in.p;
    p.src_pitch = in.pitch;
    p.src_frame = in.frame;
    p.ll = out.p;
    p.ll_pitch = out.pitch;
    p.ll_frame = out.frame;
    const int ody = cdiv_pow2(im->oy, j + 1), odx = cdiv_pow2(im->ox, j + 1);
    p.hl = dst_plane + (size_t)odx * im->es;
    p.lh = dst_plane + (size_t)ody * im->pitch * im->es;
    p.hh = dst_plane + ((size_t)ody * im->pitch + odx) * im->es;
    p.sub_pitch = im->pitch;
    p.sub_frame = im->frame;
    if (j == 0 && (im->up[im->cur] || im->dn[im->cur])) {
        p.src_up = im->up[im->cur];
        p.src_dn = im->dn[im->cur];
        p.up_end = im->up[im->cur] ? im->up_end : 0;
        p.dn_begin = im->dn[im->cur] ? im->dn_begin : 0;
        p.up_row0 = im->up_row0;
        p.dn_row0 = im->dn_row0;
    }
    return out;
}

void inv_level_params(const dwtb200_image *im, int j, int J, char *src_plane, char *dst_plane, LevelParams &p)
{
    memset(&p, 0, sizeof p);
    level_geometry(im, j, true, p);
    const Band in = (j == J - 1) ? Band{src_plane, im->pitch, im->frame} : ll_band(im, j);
    const Band out = (j == 0) ? Band{dst_plane, im->pitch, im->frame} : ll_band(im, j - 1);
    p.ll = in.p;
    p.ll_pitch = in.pitch;
    p.ll_frame = in.frame;
    const int ody = cdiv_pow2(im->oy, j + 1), odx = cdiv_pow2(im->ox, j + 1);
    p.hl = src_plane + (size_t)odx * im->es;
    p.lh = src_plane + (size_t)ody * im->pitch * im->es;
    p.hh = src_plane + ((size_t)ody * im->pitch + odx) * im->es;
    p.sub_pitch = im->pitch;
    p.sub_frame = im->frame;
    p.h_room = (int)(im->pitch - odx);
    p.dst = out.p;
    p.dst_pitch = out.pitch;
    p.dst_frame = out.frame;
}

void stream_fwd(int kind, const LevelParams &p, int frames, cudaStream_t st)
{
    if ((g.ring & 1) && !p.narrow) launch_fwd_ring(kind, p, frames, p.cfg, st);
    else launch_fwd_level(kind, p, frames, st);
}
void stream_inv(int kind, const LevelParams &p, int frames, cudaStream_t st)
{
    if ((g.ring & 2) && !p.narrow && p.sub_aligned) launch_inv_ring(kind, p, frames, p.cfg, st);
    else launch_inv_level(kind, p, frames, st);
}

// ---- the dense path as a list of launches, linked into a chain (struct Chain, kernels.h) ----------------
struct Launch {
    enum { RING_F, REG_F, TILE_F, TAIL_F, RING_I, REG_I, TILE_I, TAIL_I } type;
    LevelParams lp;
    TailParams tp;
    bool chainable = false;
    // how consumers find the row block of an LL row this launch produces
    int out_nblocks = 0, out_div = 1, out_bias = 0, out_need = 0;
    unsigned total = 0;
    Chain *chain() { return (type == TAIL_F || type == TAIL_I) ? &tp.chain : &lp.chain; }
};

void plan_level(dwtb200_image *im, Launch &L, bool inverse, int plan_type)
{
    const LevelParams &p = L.lp;
    if (plan_type == PLAN_TILE) {
        L.type = inverse ? Launch::TILE_I : Launch::TILE_F;
        const dim3 gr = tile_grid_of(im->kind, p, im->frames);
        L.chainable = (g.chain & 2) != 0;
        L.out_nblocks = (int)gr.y;
        L.out_div = inverse ? tile_rows() : tile_rows() / 2;
        L.out_need = (int)gr.x;
        L.total = gr.x * gr.y * gr.z;
        return;
    }
    const bool ring = !p.narrow && (inverse ? ((g.ring & 2) && p.sub_aligned) : (g.ring & 1));
    L.type = inverse ? (ring ? Launch::RING_I : Launch::REG_I) : (ring ? Launch::RING_F : Launch::REG_F);
    L.chainable = ring && (g.chain & 1);
    if (ring) {
        L.out_nblocks = p.nstrips;
        L.out_div = inverse ? 2 * p.pps : p.pps;
        L.out_bias = inverse ? 1 : 0;
        L.out_need = p.nbands;
        L.total = (unsigned)p.nbands * p.nstrips * im->frames;
    }
}

size_t chain_words(const std::vector<Launch> &ls, int frames)
{
    size_t w = 2;
    for (const Launch &l : ls) w += (size_t)l.out_nblocks * frames;
    return w;
}

// link consecutive chainable launches; `sync` = generation word, done counter, then the launches' counters
void link_chain(std::vector<Launch> &ls, uint32_t *sync, int frames)
{
    const int n = (int)ls.size();
    size_t off = 2;
    int last = -1;
    std::vector<uint32_t *> flags(n, nullptr);
    for (int i = 0; i < n; i++) {
        flags[i] = sync + off;
        off += (size_t)ls[i].out_nblocks * frames;
    }
    for (int i = 0; i < n; i++) {
        Chain &c = *ls[i].chain();
        memset(&c, 0, sizeof c);
        if (!g.chain || !sync) continue;
        const bool in = i > 0 && ls[i - 1].chainable && ls[i].chainable;
        const bool out = i + 1 < n && ls[i].chainable && ls[i + 1].chainable;
        if (!in && !out) continue;
        c.gen = sync;
        c.total = ls[i].total;
        if (in) {
            c.in = flags[i - 1];
            c.in_nblocks = ls[i - 1].out_nblocks;
            c.in_div = ls[i - 1].out_div;
            c.in_bias = ls[i - 1].out_bias;
            c.in_need = ls[i - 1].out_need;
            c.pdl = 1;
        }
        if (out) {
            c.out = flags[i];
            c.out_nblocks = ls[i].out_nblocks;
        }
        last = i;
    }
    if (last >= 0) ls[last].chain()->done = sync + 1;
}

int issue(dwtb200_image *im, std::vector<Launch> &ls)
{
    for (Launch &l : ls) {
        cudaError_t e = cudaSuccess;
        switch (l.type) {
        case Launch::RING_F: launch_fwd_ring(im->kind, l.lp, im->frames, l.lp.cfg, g.st); break;
        case Launch::REG_F: launch_fwd_level(im->kind, l.lp, im->frames, g.st); break;
        case Launch::TILE_F: launch_fwd_tile(im->kind, l.lp, im->frames, g.st); break;
        case Launch::TAIL_F: launch_fwd_tail(im->kind, l.tp, im->frames, g.st); break;
        case Launch::RING_I: launch_inv_ring(im->kind, l.lp, im->frames, l.lp.cfg, g.st); break;
        case Launch::REG_I: launch_inv_level(im->kind, l.lp, im->frames, g.st); break;
        case Launch::TILE_I: launch_inv_tile(im->kind, l.lp, im->frames, g.st); break;
        case Launch::TAIL_I: launch_inv_tail(im->kind, l.tp, im->frames, g.st); break;
        }
        if (e != cudaSuccess) return fail(DWTB200_ECUDA, "cooperative launch: %s", cudaGetErrorString(e));
        g.launches++;
    }
    return 0;
}

// levels jstart .. J-1 of the forward transform (jstart > 0: the caller ran the levels below it itself)
void plan_fwd_dense(dwtb200_image *im, int J, const DensePlan &pl, int jstart, std::vector<Launch> &ls)
{
    char *src_plane = (char *)im->plane[im->cur], *dst_plane = (char *)im->plane[im->cur ^ 1];
    Band in = {src_plane, im->pitch, im->frame};
    if (jstart > 0) in = ll_band(im, jstart - 1);
    auto tail_params = [&](int j, const Band &from) {
        TailParams t;
        memset(&t, 0, sizeof t);
        t.src = from.p;
        t.src_pitch = from.pitch;
        t.src_frame = from.frame;
        t.dst = dst_plane;
        t.dst_pitch = im->pitch;
        t.dst_frame = im->frame;
        t.W0 = im->ox;
        t.H0 = im->oy;
        t.j0 = j;
        t.j1 = J;
        return t;
    };
    for (int j = jstart; j < J; j++) {
        ls.emplace_back();
        Launch &L = ls.back();
        if (j == pl.jt) {   // stand-alone tail
            L.type = Launch::TAIL_F;
            L.tp = tail_params(j, in);
            L.chainable = (g.chain & 2) != 0;
            L.total = im->frames;
            return;
        }
        in = fwd_level_params(im, j, J, in, dst_plane, L.lp);
        plan_level(im, L, false, pl.type[j]);
    }
}

// the tail (if any) and the levels jtop-1 .. jstop of the inverse transform
void plan_inv_dense(dwtb200_image *im, int J, const DensePlan &pl, int jstop, std::vector<Launch> &ls)
{
    char *src_plane = (char *)im->plane[im->cur], *dst_plane = (char *)im->plane[im->cur ^ 1];
    const int jt = pl.jt;
    auto tail_params = [&]() {
        const Band out = (jt == 0) ? Band{dst_plane, im->pitch, im->frame} : ll_band(im, jt - 1);
        TailParams t;
        memset(&t, 0, sizeof t);
        t.src = src_plane;
        t.src_pitch = im->pitch;
        t.src_frame = im->frame;
        t.dst = out.p;
        t.dst_pitch = out.pitch;
        t.dst_frame = out.frame;
        t.W0 = im->ox;
        t.H0 = im->oy;
        t.j0 = jt;
        t.j1 = J;
        return t;
    };
    int jtop = jt < J ? jt : J;   // levels jtop-1 .. 0 remain after the tail
    if (jt < J) {
        ls.emplace_back();
        Launch &L = ls.back();
        L.type = Launch::TAIL_I;
        L.tp = tail_params();
        L.chainable = (g.chain & 2) != 0;
        L.out_nblocks = 1;
        L.out_div = 1 << 30;
        L.out_need = 1;
        L.total = im->frames;
    }
    for (int j = jtop - 1; j >= jstop; j--) {   // jstop > 0: the caller runs the levels below it (pipelined host path)
        ls.emplace_back();
        Launch &L = ls.back();
        inv_level_params(im, j, J, src_plane, dst_plane, L.lp);
        plan_level(im, L, true, pl.type[j]);
    }
}

// un-captured launch sequences (no-graph mode, pipelined host path) share the image's own counters, zeroed per call
int run_dense_uncaptured(dwtb200_image *im, bool inverse, int J, const DensePlan &pl, int jlimit)
{
    std::vector<Launch> ls;
    if (inverse) plan_inv_dense(im, J, pl, jlimit, ls);
    else plan_fwd_dense(im, J, pl, jlimit, ls);
    uint32_t *sync = nullptr;
    if (g.chain && chain_words(ls, im->frames) <= im->sync_words) {
        sync = im->sync;
        CK(cudaMemsetAsync(sync, 0, chain_words(ls, im->frames) * sizeof(uint32_t), g.st));
    }
    link_chain(ls, sync, im->frames);
    return issue(im, ls);
}

// one generic pass A -> B over the level's outer region, then the second pass B -> A; a pass the
// reference skips is replaced by nothing, and the result is copied back so that A is always complete
void generic_level(dwtb200_image *im, bool inverse, bool first_along_x, bool do_first, bool do_second, int region_w,
                   int region_h, int Nx, int Ny, int offx, int offy)
{
    char *A = (char *)im->plane[im->cur], *B = (char *)im->plane[im->cur ^ 1];
    auto pass = [&](char *s, char *d, bool along_x) {
        PassParams p;
        memset(&p, 0, sizeof p);
        p.src = s;
        p.dst = d;
        p.src_pitch = p.dst_pitch = im->pitch;
        p.src_frame = p.dst_frame = im->frame;
        p.region_w = region_w;
        p.region_h = region_h;
        p.along_x = along_x ? 1 : 0;
        p.N = along_x ? Nx : Ny;
        p.off_h = along_x ? offx : offy;
        if (inverse) launch_pass_inv(im->kind, p, im->frames, g.st);
        else launch_pass_fwd(im->kind, p, im->frames, g.st);
        g.launches++;
    };
    auto copy_back = [&]() {
        launch_copy2d((int)im->es, A, im->pitch, B, im->pitch, region_w, region_h, im->frame, im->frame, im->frames, g.st);
        g.launches++;
    };
    if (do_first && do_second) {
        pass(A, B, first_along_x);
        pass(B, A, !first_along_x);
    } else if (do_first) {
        pass(A, B, first_along_x);
        copy_back();
    } else if (do_second) {
        pass(A, B, !first_along_x);
        copy_back();
    }
}

// level 0 of the out-of-place forward transform on a sparse layout (dwt_cdf97_2f_s2, src/libdwt.c:12678-12745): plane A
// (current) holds dst's old content, plane B the source samples; the first pass that runs reads B and writes only its L / H
// outputs into A, the rest of A keeps dst's content; the second pass runs in place on A
void s2_level0(dwtb200_image *im, int ix, int iy)
{
    char *A = (char *)im->plane[im->cur], *B = (char *)im->plane[im->cur ^ 1];
    const bool gd = guard(im->kind);
    const bool rows = !gd || im->ox > 1, cols = !gd || im->oy > 1;
    const int odx = cdiv_pow2(im->ox, 1), ody = cdiv_pow2(im->oy, 1);
    auto pass = [&](char *s, char *d, bool along_x, int keep) {
        PassParams p;
        memset(&p, 0, sizeof p);
        p.src = s;
        p.dst = d;
        p.src_pitch = p.dst_pitch = im->pitch;
        p.src_frame = p.dst_frame = im->frame;
        p.region_w = im->ox;
        p.region_h = im->oy;
        p.along_x = along_x ? 1 : 0;
        p.N = along_x ? ix : iy;
        p.off_h = along_x ? odx : ody;
        p.keep_dst = keep;
        launch_pass_fwd(im->kind, p, im->frames, g.st);
        g.launches++;
    };
    if (rows) {
        pass(B, A, true, 1);
        if (cols) {
            pass(A, B, false, 0);
            launch_copy2d((int)im->es, A, im->pitch, B, im->pitch, im->ox, im->oy, im->frame, im->frame, im->frames, g.st);
            g.launches++;
        }
    } else if (cols) {
        pass(B, A, false, 1);
    }
}

// src/libdwt.c:12896-12916: after level j, zero what lies between the inner subbands and the outer ones
void fwd_zero_padding(dwtb200_image *im, int ix, int iy, int j)
{
    const int osx = cdiv_pow2(im->ox, j), osy = cdiv_pow2(im->oy, j);
    const int odx = cdiv_pow2(im->ox, j + 1), ody = cdiv_pow2(im->oy, j + 1);
    const int isx = cdiv_pow2(ix, j), isy = cdiv_pow2(iy, j);
    ZeroParams z;
    z.buf = im->plane[im->cur];
    z.pitch = im->pitch;
    z.frame = im->frame;
    z.region_w = osx;
    z.region_h = osy;
    z.x0a = (isx + 1) >> 1;
    z.x0b = odx;
    z.x1a = odx + (isx >> 1);
    z.x1b = osx;
    z.y0a = (isy + 1) >> 1;
    z.y0b = ody;
    z.y1a = ody + (isy >> 1);
    z.y1b = osy;
    launch_zero(im->kind, z, im->frames, g.st);
    g.launches++;
}

void run_fwd_generic(dwtb200_image *im, int ix, int iy, int J, int zero_padding, int jstart = 0)
{
    const bool gd = guard(im->kind);
    for (int j = jstart; j < J; j++) {
        const int osx = cdiv_pow2(im->ox, j), osy = cdiv_pow2(im->oy, j);
        const int odx = cdiv_pow2(im->ox, j + 1), ody = cdiv_pow2(im->oy, j + 1);
        const int isx = cdiv_pow2(ix, j), isy = cdiv_pow2(iy, j);
        generic_level(im, false, true, !gd || osx > 1, !gd || osy > 1, osx, osy, isx, isy, odx, ody);
        if (zero_padding) fwd_zero_padding(im, ix, iy, j);
    }
}

void run_inv_generic(dwtb200_image *im, int ix, int iy, int J, int zero_padding)
{
    const bool gd = guard(im->kind), cf = inv_cols_first(im->kind);
    for (int j = J; j > 0; j--) {
        const int osx = cdiv_pow2(im->ox, j), osy = cdiv_pow2(im->oy, j);
        const int odx = cdiv_pow2(im->ox, j - 1), ody = cdiv_pow2(im->oy, j - 1);
        const int idx = cdiv_pow2(ix, j - 1), idy = cdiv_pow2(iy, j - 1);
        if (cf) generic_level(im, true, false, true, true, odx, ody, idx, idy, osx, osy);
        else generic_level(im, true, true, !gd || odx > 1, !gd || ody > 1, odx, ody, idx, idy, osx, osy);
        if (zero_padding) {   // src/libdwt.c:17156-17176
            ZeroParams z;
            z.buf = im->plane[im->cur];
            z.pitch = im->pitch;
            z.frame = im->frame;
            z.region_w = odx;
            z.region_h = ody;
            z.x0a = idx;
            z.x0b = odx;
            z.x1a = z.x1b = 0;
            z.y0a = idy;
            z.y0b = ody;
            z.y1a = z.y1b = 0;
            launch_zero(im->kind, z, im->frames, g.st);
            g.launches++;
        }
    }
}

int transform(dwtb200_image *im, bool inverse, int ix, int iy, int J, int zero_padding)
{
    if (ix < 1 || iy < 1 || ix > im->ox || iy > im->oy) return fail(DWTB200_EINVAL, "inner size %d x %d outside outer %d x %d", ix, iy, im->ox, im->oy);
    im->last_launches = 0;
    if (J == 0) return DWTB200_OK;
    const bool dense_shape = ix == im->ox && iy == im->oy && !g.force_generic;
    DensePlan pl;
    pl.jt = -1;
    if (dense_shape) pl = dense_plan(im, J);
    const bool dense = pl.jt >= 0;

    const dwtb200_image::Key key(inverse, ix, iy, J, zero_padding, im->cur, g.force_generic, g.strip_rows, g.epoch);
    auto it = g.use_graph ? im->graphs.find(key) : im->graphs.end();
    if (it == im->graphs.end()) {
        g.launches = 0;
        cudaGraph_t graph = nullptr;
        std::vector<Launch> ls;
        uint32_t *sync = nullptr;
        if (dense && g.use_graph) {   // a captured chain owns its counters: zeroed once, generations thereafter
            if (inverse) plan_inv_dense(im, J, pl, 0, ls);
            else plan_fwd_dense(im, J, pl, 0, ls);
            if (g.chain) {
                const size_t bytes = chain_words(ls, im->frames) * sizeof(uint32_t);
                CK(cudaMalloc((void **)&sync, bytes));
                CK(cudaMemsetAsync(sync, 0, bytes, g.st));
                CK(cudaStreamSynchronize(g.st));
            }
            link_chain(ls, sync, im->frames);
        }
        if (g.use_graph) CK(cudaStreamBeginCapture(g.st, cudaStreamCaptureModeThreadLocal));
        int rr = 0;
        if (dense) {
            rr = g.use_graph ? issue(im, ls) : run_dense_uncaptured(im, inverse, J, pl, 0);
        } else {
            if (inverse) run_inv_generic(im, ix, iy, J, zero_padding);
            else run_fwd_generic(im, ix, iy, J, zero_padding);
        }
        const cudaError_t le = cudaGetLastError();
        if (g.use_graph) {
            const cudaError_t ce = cudaStreamEndCapture(g.st, &graph);
            if (rr) {
                if (graph) cudaGraphDestroy(graph);
                if (sync) cudaFree(sync);
                return rr;
            }
            if (le != cudaSuccess || ce != cudaSuccess) {
                if (graph) cudaGraphDestroy(graph);
                if (sync) cudaFree(sync);
                return fail(DWTB200_ECUDA, "launch/capture failed: %s / %s", cudaGetErrorString(le), cudaGetErrorString(ce));
            }
            dwtb200_image::Entry e;
            e.launches = g.launches;
            e.path = dense ? 0 : 1;
            e.flips = dense ? 1 : 0;
            e.sync = sync;
            const cudaError_t ie = cudaGraphInstantiate(&e.exec, graph, 0);
            cudaGraphDestroy(graph);
            if (ie != cudaSuccess && sync) cudaFree(sync);
            if (ie != cudaSuccess) return fail(DWTB200_ECUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(ie));
            it = im->graphs.emplace(key, e).first;
        } else {
            if (rr) return rr;
            if (le != cudaSuccess) return fail(DWTB200_ECUDA, "kernel launch failed: %s", cudaGetErrorString(le));
            im->last_launches = g.launches;
            im->last_path = dense ? 0 : 1;
            if (dense) im->cur ^= 1;
            return DWTB200_OK;
        }
    }
    CK(cudaGraphLaunch(it->second.exec, g.st));
    im->last_launches = it->second.launches;
    im->last_path = it->second.path;
    if (it->second.flips) im->cur ^= 1;
    return DWTB200_OK;
}
}  // namespace

int dwtb200_image_fwd2(dwtb200_image *im, int ix, int iy, int *j_max_ptr, int decompose_one, int zero_padding)
{
    API_LOCK();
    ImageScope scope(im);
    NEED_DEV();
    if (!im || !j_max_ptr) return fail(DWTB200_EINVAL, "image_fwd2: null argument");
    *j_max_ptr = dwtb200_clamp_j(*j_max_ptr, im->ox, im->oy, decompose_one);   // src/libdwt.c:12807-12810
    return transform(im, false, ix, iy, *j_max_ptr, zero_padding);
}

int dwtb200_image_inv2(dwtb200_image *im, int ix, int iy, int j_max, int decompose_one, int zero_padding)
{
    API_LOCK();
    ImageScope scope(im);
    NEED_DEV();
    if (!im) return fail(DWTB200_EINVAL, "image_inv2: null argument");
    const int J = dwtb200_clamp_j(j_max, im->ox, im->oy, decompose_one);   // src/libdwt.c:17063-17066
    return transform(im, true, ix, iy, J, zero_padding);
}

// The per-subband feature vectors of the reference (dwt_util_wps_s, _mean_s, _var_s, _stdev_s, _maxnorm_s, _norm_s:
// src/libdwt.c:23201, 23515, 23549, 23583, 23686, 23754; band functions :23086-23480) for a device-resident Mallat plane:
// levels j = 1 .. j_max - 1 (the reference's loop bound), bands HL, LH, HH of each, skipping empty bands.  The reference
// accumulates sequentially in float; here the sums come from the double-precision device reduction.
int dwtb200_image_features(dwtb200_image *im, int frame, int ix, int iy, int j_max, int feature, float *fv, int *count)
{
    API_LOCK();
    NEED_DEV();
    if (!im || !fv || feature < 0 || feature > DWTB200_FEAT_NORM) return fail(DWTB200_EINVAL, "image_features: bad arguments");
    int n = 0;
    for (int j = 1; j < j_max; j++)
        for (int band = 1; band <= 3; band++) {   // DWT_HL, DWT_LH, DWT_HH
            int sx = 0, sy = 0;
            void *p = nullptr;
            int r = dwtb200_image_subband(im, frame, ix, iy, j, band, &p, nullptr, &sx, &sy);
            if (r) return r;
            if (!sx || !sy) continue;
            double sum = 0, sq = 0, mx = 0;
            r = dwtb200_image_subband_moments(im, frame, ix, iy, j, band, &sum, &sq, &mx);
            if (r) return r;
            const double cnt = (double)sx * sy, mean = sum / cnt;
            double v = 0;
            switch (feature) {
            case DWTB200_FEAT_WPS: v = sq / (double)((int64_t)1 << j); break;      // :23112 rectification by 2^j
            case DWTB200_FEAT_MEAN: v = mean; break;
            case DWTB200_FEAT_VAR: v = std::max(sq / cnt - mean * mean, 0.0); break;   // second central moment (:23349)
            case DWTB200_FEAT_STDEV: v = sqrt(std::max(sq / cnt - mean * mean, 0.0)); break;
            case DWTB200_FEAT_MAXNORM: v = mx; break;
            default: v = sqrt(sq); break;                                           // l2 norm (:23470)
            }
            fv[n++] = (float)v;
        }
    if (count) *count = n;
    return DWTB200_OK;
}

// =====================================================================================================
// interleaved in-place family (src/libdwt.c:12926, 13485, 13641, 14847, 17474, 16553, 17886; kernels_inplace.cu)
// =====================================================================================================
namespace {
constexpr int IP_STD_MIN = 32;            // a level with a side below this is evaluated by k_ip_phase alone
constexpr int IP_WHOLE_MAX = 256 * 256;   // samples (all frames) up to which a level is evaluated by k_ip_phase alone
constexpr int IP_TOP = 8, IP_RIGHT = 6;   // frame whose sweep order differs from rows-then-columns (7 / 8 rows, 5 columns)

// the interleaved layout is served by the first-generation ring kernels: their geometry must be the one level_geometry computes
struct RingV2Off {
    int prev;
    RingV2Off() : prev(g.ring_v2) { g.ring_v2 = 0; }
    ~RingV2Off() { g.ring_v2 = prev; }
};

// one level of the 9/7 family: the ordinary level kernel, then the exact schedule over the top and right frame
int ip_level(dwtb200_image *im, bool inverse, const LevelParams &lp)
{
    const int nx = lp.W, ny = lp.H, top = std::min(IP_TOP, ny);
    // small levels: the exact kernel alone (one launch of ~7 us) beats tile kernel + frame kernel (5 + 7 us)
    if (std::min(nx, ny) < IP_STD_MIN || (int64_t)nx * ny * im->frames <= IP_WHOLE_MAX) {
        launch_ip_phase(inverse, lp, im->frames, 0, 0, nx, ny, 0, 0, 0, 0, g.st);
        g.launches++;
        return 0;
    }
    std::vector<Launch> one(1);
    one[0].lp = lp;
    plan_level(im, one[0], inverse, (int64_t)nx * ny * im->frames <= g.tile_max ? PLAN_TILE : PLAN_STREAM);
    memset(&one[0].lp.chain, 0, sizeof(Chain));
    const int r = issue(im, one);
    if (r) return r;
    launch_ip_phase(inverse, lp, im->frames, 0, 0, nx, top, std::max(nx - IP_RIGHT, 0), top, nx, ny, g.st);
    g.launches++;
    return 0;
}

// `flips`: the result is left in the other plane
int ip_run97(dwtb200_image *im, bool inverse, int J, int &flips)
{
    char *A = (char *)im->plane[im->cur], *B = (char *)im->plane[im->cur ^ 1];   // A: the image (interleaved), B: scratch
    flips = 0;
    // levels jt .. J-1 run inside one CTA per frame once their input fits its shared memory (k_ip_tail)
    int jt = J;
    for (int j = 0; j < J; j++)
        if ((int64_t)cdiv_pow2(im->ox, j) * cdiv_pow2(im->oy, j) <= ip_tail_cap()) {
            jt = j;
            break;
        }
    if (jt == 0) {   // the whole image: in place, nothing to translate
        launch_ip_tail(inverse, A, im->pitch, im->frame, im->ox, im->oy, J, im->frames, g.st);
        g.launches++;
        return 0;
    }
    const bool tail = jt < J;
    const Band tb = ll_band(im, jt - 1);   // LL_{jt-1}: dense input of level jt, and the tail block
    const int tw = cdiv_pow2(im->ox, jt), th = cdiv_pow2(im->oy, jt);
    LevelParams lp;
    // Level 0 on the ring kernels reads / writes the interleaved layout directly (LevelParams::il): the image is then
    // translated only at its even rows and columns, where the deeper levels live.
    level_geometry(im, 0, inverse, lp);
    const bool il0 = std::min(im->ox, im->oy) >= IP_STD_MIN && (int64_t)im->ox * im->oy * im->frames > g.tile_max && !lp.narrow &&
                     (g.ring & (inverse ? 2 : 1)) && ring_interleaved_ok(im->kind) && ring_interleaved_cfg_ok(lp.cfg) && !g.force_generic;
    char *M = il0 ? A : B;   // Mallat scratch of the levels the translation kernels handle (forward: level 0 has been read by then)
    const int shift = il0 ? 1 : 0;
    void *tailp = tail ? tb.p : nullptr;
    if (!inverse) {
        Band in = {A, im->pitch, im->frame};
        for (int j = 0; j < jt; j++) {
            Band out = fwd_level_params(im, j, J, in, il0 ? (j == 0 ? B : A) : B, lp);
            if (il0 && j == 0) {
                lp.il = B;
                lp.il_pitch = im->pitch;
                lp.il_frame = im->frame;
                out = ll_band(im, 0);   // also when J == 1: the interleaved rows hold LL already
                lp.ll = out.p;
                lp.ll_pitch = out.pitch;
                lp.ll_frame = out.frame;
            }
            const int r = ip_level(im, false, lp);
            if (r) return r;
            in = out;
        }
        if (tail) launch_ip_tail(false, tb.p, tb.pitch, tb.frame, tw, th, J - jt, im->frames, g.st);
        if (!il0 || J > 1) launch_ip_pack(false, M, il0 ? B : A, im->pitch, im->frame, im->ox, im->oy, J, tailp, tb.pitch, tb.frame, jt, shift, im->frames, g.st);
    } else {
        // (J == 1 on the interleaved route: LL_1 goes to the dense band of LL_0, not to the plane the level writes its output to)
        const Band l0 = ll_band(im, 0);
        if (il0 && J == 1) launch_ip_pack(true, A, B, im->pitch, im->frame, im->ox, im->oy, J, l0.p, l0.pitch, l0.frame, 1, shift, im->frames, g.st);
        else launch_ip_pack(true, A, B, im->pitch, im->frame, im->ox, im->oy, J, tailp, tb.pitch, tb.frame, jt, shift, im->frames, g.st);
        if (tail) launch_ip_tail(true, tb.p, tb.pitch, tb.frame, tw, th, J - jt, im->frames, g.st);
        for (int j = jt - 1; j >= 0; j--) {
            inv_level_params(im, j, J, B, il0 ? B : A, lp);
            if (il0 && j == 0) {   // HL, LH, HH from the interleaved plane, LL_0 from its dense band
                lp.ll = l0.p;
                lp.ll_pitch = l0.pitch;
                lp.ll_frame = l0.frame;
                lp.il = A;
                lp.il_pitch = im->pitch;
                lp.il_frame = im->frame;
                lp.sub_aligned = 1;   // no HL / HH column origins to align in this layout: always the ring kernel
            }
            const int r = ip_level(im, true, lp);
            if (r) return r;
        }
    }
    g.launches += tail ? 2 : 1;
    flips = il0 ? 1 : 0;
    return 0;
}

// CDF 5/3 float of the family is the Mallat transform bit for bit; with level 0 on the ring kernels it takes the same
// route as above without the frame kernels: interleaved level 0, the other levels in a Mallat scratch plane (streaming / tile
// kernels and the Mallat tail kernel), translation of the even rows and columns only.  false: this route does not apply.
bool ip_53_fast(const dwtb200_image *im, bool inverse, int J, DensePlan &pl)
{
    if (g.force_generic || std::min(im->ox, im->oy) < IP_STD_MIN || (int64_t)im->ox * im->oy * im->frames <= g.tile_max) return false;
    LevelParams lp;
    level_geometry(im, 0, inverse, lp);
    if (lp.narrow || !(g.ring & (inverse ? 2 : 1)) || !ring_interleaved_ok(im->kind) || !ring_interleaved_cfg_ok(lp.cfg)) return false;
    pl = dense_plan(im, J);
    if (pl.jt < 1) return false;   // degenerate pyramid
    return true;
}
int ip_run53(dwtb200_image *im, bool inverse, int J, const DensePlan &pl)
{
    char *A = (char *)im->plane[im->cur], *B = (char *)im->plane[im->cur ^ 1];
    const int jt = pl.jt;
    std::vector<Launch> ls;
    auto level = [&](LevelParams &lp, int j) {
        ls.emplace_back();
        ls.back().lp = lp;
        plan_level(im, ls.back(), inverse, j == 0 ? PLAN_STREAM : pl.type[j]);
    };
    auto run = [&]() {
        for (Launch &l : ls) memset(l.chain(), 0, sizeof(Chain));
        const int r = issue(im, ls);
        ls.clear();
        return r;
    };
    LevelParams lp;
    TailParams t;
    memset(&t, 0, sizeof t);
    t.W0 = im->ox;
    t.H0 = im->oy;
    t.j0 = jt;
    t.j1 = J;
    if (!inverse) {
        Band in = {A, im->pitch, im->frame};
        for (int j = 0; j < jt; j++) {
            Band out = fwd_level_params(im, j, J, in, j == 0 ? B : A, lp);
            if (j == 0) {
                lp.il = B;
                lp.il_pitch = im->pitch;
                lp.il_frame = im->frame;
                out = ll_band(im, 0);
                lp.ll = out.p;
                lp.ll_pitch = out.pitch;
                lp.ll_frame = out.frame;
            }
            level(lp, j);
            in = out;
        }
        if (jt < J) {
            ls.emplace_back();
            ls.back().type = Launch::TAIL_F;
            t.src = in.p;
            t.src_pitch = in.pitch;
            t.src_frame = in.frame;
            t.dst = A;
            t.dst_pitch = im->pitch;
            t.dst_frame = im->frame;
            ls.back().tp = t;
        }
        const int r = run();
        if (r) return r;
        if (J > 1) launch_ip_pack(false, A, B, im->pitch, im->frame, im->ox, im->oy, J, nullptr, 0, 0, 0, 1, im->frames, g.st);
    } else {
        const Band l0 = ll_band(im, 0);   // J == 1: LL_1 goes straight to the dense band the level reads, not into the plane it writes
        launch_ip_pack(true, A, B, im->pitch, im->frame, im->ox, im->oy, J, J == 1 ? l0.p : nullptr, l0.pitch, l0.frame, 1, 1, im->frames, g.st);
        if (jt < J) {
            const Band out = ll_band(im, jt - 1);
            ls.emplace_back();
            ls.back().type = Launch::TAIL_I;
            t.src = B;
            t.src_pitch = im->pitch;
            t.src_frame = im->frame;
            t.dst = out.p;
            t.dst_pitch = out.pitch;
            t.dst_frame = out.frame;
            ls.back().tp = t;
        }
        for (int j = jt - 1; j >= 1; j--) {
            inv_level_params(im, j, J, B, B, lp);
            level(lp, j);
        }
        int r = run();
        if (r) return r;
        inv_level_params(im, 0, J, B, B, lp);
        lp.ll = l0.p;
        lp.ll_pitch = l0.pitch;
        lp.ll_frame = l0.frame;
        lp.il = A;   // HL, LH, HH from the interleaved plane, LL_0 from its dense band
        lp.il_pitch = im->pitch;
        lp.il_frame = im->frame;
        lp.sub_aligned = 1;
        level(lp, 0);
        r = run();
        if (r) return r;
    }
    g.launches++;
    return 0;
}

// J levels of the family on the whole image (J as given: the host entry points clamp it against the caller's OUTER size)
int inplace_transform(dwtb200_image *im, bool inverse, int J)
{
    if (im->kind != DWTB200_CDF97_F32 && im->kind != DWTB200_CDF53_F32)
        return fail(DWTB200_EINVAL, "in-place family: CDF 9/7 float and CDF 5/3 float only (kind %d)", im->kind);
    RingV2Off v2off;
    im->last_launches = 0;
    if (J <= 0) return DWTB200_OK;
    DensePlan pl53;
    const bool is53 = im->kind == DWTB200_CDF53_F32;
    if (is53 && !ip_53_fast(im, inverse, J, pl53)) {   // the Mallat transform and a translation of the whole plane (:16583; a lone sample is scaled)
        int r = 0;
        if (inverse) {
            launch_ip_pack(true, im->plane[im->cur], im->plane[im->cur ^ 1], im->pitch, im->frame, im->ox, im->oy, J, nullptr, 0, 0, 0, 0, im->frames, g.st);
            im->cur ^= 1;
            r = transform(im, true, im->ox, im->oy, J, 0);
        } else {
            r = transform(im, false, im->ox, im->oy, J, 0);
            if (r) return r;
            launch_ip_pack(false, im->plane[im->cur], im->plane[im->cur ^ 1], im->pitch, im->frame, im->ox, im->oy, J, nullptr, 0, 0, 0, 0, im->frames, g.st);
            im->cur ^= 1;
        }
        im->last_launches++;
        CK(cudaGetLastError());
        return r;
    }
    const int jcap = dwtb200_ceil_log2(std::max(im->ox, im->oy));   // beyond it every line has one sample: 9/7 touches nothing (:12975)
    if (!is53 && J > jcap) J = jcap;
    if (J <= 0) return DWTB200_OK;
    auto body = [&](int &flips) {
        if (is53) {
            flips = 1;
            return ip_run53(im, inverse, J, pl53);
        }
        return ip_run97(im, inverse, J, flips);
    };
    const dwtb200_image::Key key(inverse, im->ox, im->oy, J, 0x100, im->cur, g.force_generic, g.strip_rows, g.epoch);
    auto it = g.use_graph ? im->graphs.find(key) : im->graphs.end();
    if (it == im->graphs.end()) {
        g.launches = 0;
        if (!g.use_graph) {
            int flips = 0;
            const int r = body(flips);
            if (r) return r;
            CK(cudaGetLastError());
            im->last_launches = g.launches;
            if (flips) im->cur ^= 1;
            return DWTB200_OK;
        }
        cudaGraph_t graph = nullptr;
        CK(cudaStreamBeginCapture(g.st, cudaStreamCaptureModeThreadLocal));
        int flips = 0;
        const int rr = body(flips);
        const cudaError_t le = cudaGetLastError(), ce = cudaStreamEndCapture(g.st, &graph);
        if (rr || le != cudaSuccess || ce != cudaSuccess) {
            if (graph) cudaGraphDestroy(graph);
            return rr ? rr : fail(DWTB200_ECUDA, "in-place launch/capture failed: %s / %s", cudaGetErrorString(le), cudaGetErrorString(ce));
        }
        dwtb200_image::Entry e;
        e.launches = g.launches;
        e.path = 0;
        e.flips = flips;
        e.sync = nullptr;
        const cudaError_t ie = cudaGraphInstantiate(&e.exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ie != cudaSuccess) return fail(DWTB200_ECUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(ie));
        it = im->graphs.emplace(key, e).first;
    }
    CK(cudaGraphLaunch(it->second.exec, g.st));
    im->last_launches = it->second.launches;
    im->last_path = 0;
    if (it->second.flips) im->cur ^= 1;
    return DWTB200_OK;
}
}  // namespace

int dwtb200_image_fwd2_inplace(dwtb200_image *im, int *j_max_ptr, int decompose_one)
{
    API_LOCK();
    ImageScope scope(im);
    NEED_DEV();
    if (!im || !j_max_ptr) return fail(DWTB200_EINVAL, "image_fwd2_inplace: null argument");
    *j_max_ptr = dwtb200_clamp_j(*j_max_ptr, im->ox, im->oy, decompose_one);   // src/libdwt.c:12948-12951
    return inplace_transform(im, false, *j_max_ptr);
}

int dwtb200_image_inv2_inplace(dwtb200_image *im, int j_max, int decompose_one)
{
    API_LOCK();
    ImageScope scope(im);
    NEED_DEV();
    if (!im) return fail(DWTB200_EINVAL, "image_inv2_inplace: null argument");
    return inplace_transform(im, true, dwtb200_clamp_j(j_max, im->ox, im->oy, decompose_one));   // src/libdwt.c:17493-17496
}

// dwt_util_subband (src/libdwt.c:20731): where subband `band` of level j lives inside the Mallat plane, and its inner size
int dwtb200_image_subband(dwtb200_image *im, int frame, int ix, int iy, int j, int band, void **dev_ptr, size_t *pitch_bytes,
                          int *size_x, int *size_y)
{
    API_LOCK();
    if (!im || frame < 0 || frame >= im->frames || j < 0 || band < 0 || band > 3 || ix < 0 || iy < 0 || ix > im->ox || iy > im->oy)
        return fail(DWTB200_EINVAL, "image_subband: bad arguments");
    int hx = 0, hy = 0, lx = ix, ly = iy, ox = im->ox, oy = im->oy;
    for (int l = 1; l <= j; l++) {
        hx = lx >> 1;
        hy = ly >> 1;
        lx = (lx + 1) >> 1;
        ly = (ly + 1) >> 1;
        ox = (ox + 1) >> 1;
        oy = (oy + 1) >> 1;
    }
    const int col = (band == 1 || band == 3) ? ox : 0, row = (band == 2 || band == 3) ? oy : 0;   // LL, HL, LH, HH (enum dwt_subbands)
    if (dev_ptr) *dev_ptr = frame_ptr(im, im->cur, frame) + ((size_t)row * im->pitch + col) * im->es;
    if (pitch_bytes) *pitch_bytes = (size_t)im->pitch * im->es;
    if (size_x) *size_x = (band == 1 || band == 3) ? hx : lx;
    if (size_y) *size_y = (band == 2 || band == 3) ? hy : ly;
    return DWTB200_OK;
}

// sum, sum of squares and max |x| of a subband, accumulated in double on the device
int dwtb200_image_subband_moments(dwtb200_image *im, int frame, int ix, int iy, int j, int band, double *sum, double *sum_sq,
                                  double *max_abs)
{
    API_LOCK();
    ImageScope scope(im);
    NEED_DEV();
    void *p = nullptr;
    int sx = 0, sy = 0;
    int r = dwtb200_image_subband(im, frame, ix, iy, j, band, &p, nullptr, &sx, &sy);
    if (r) return r;
    double *d = nullptr, h[3] = {0, 0, 0};
    CK(cudaMalloc((void **)&d, sizeof h));
    CK(cudaMemsetAsync(d, 0, sizeof h, g.st));
    launch_moments(kind_elem_class(im->kind), p, im->pitch, sx, sy, d, g.st);
    cudaMemcpyAsync(h, d, sizeof h, cudaMemcpyDeviceToHost, g.st);
    const cudaError_t e = cudaStreamSynchronize(g.st);
    cudaFree(d);
    if (e != cudaSuccess) return fail(DWTB200_ECUDA, "subband_moments: %s", cudaGetErrorString(e));
    if (sum) *sum = h[0];
    if (sum_sq) *sum_sq = h[1];
    if (max_abs) *max_abs = h[2];
    return DWTB200_OK;
}

// dwt_util_conv_show_{s,d,i} (src/libdwt.c:21075, 21120, 21020) on device-resident planes: frame `frame` of src -> frame 0.. of dst
int dwtb200_image_conv_show(dwtb200_image *src, dwtb200_image *dst, int ix, int iy)
{
    API_LOCK();
    NEED_DEV();
    if (!src || !dst || src->kind_class() != dst->kind_class() || src->frames != dst->frames || ix < 0 || iy < 0 || ix > src->ox || iy > src->oy ||
        ix > dst->ox || iy > dst->oy)
        return fail(DWTB200_EINVAL, "image_conv_show: images must hold the same sample type and cover %d x %d", ix, iy);
    ImageScope scope(dst);
    wait_for_image(g.st, src);
    for (int f = 0; f < src->frames; f++)
        launch_conv_show(kind_elem_class(src->kind), frame_ptr(src, src->cur, f), src->pitch, frame_ptr(dst, dst->cur, f), dst->pitch, ix, iy, g.st);
    CK(cudaGetLastError());
    wait_for_image(src->st, dst);
    return DWTB200_OK;
}

// dwt_util_save_to_pgm_s / _d (src/libdwt.c:19794, 19877): the grey values are computed on the device, one byte per sample is copied
// to the host, the text file ("P2", one value per line) is written there
static int save_pgm(dwtb200_image *im, int frame, const char *filename, double max_value, double shift, int shifted, int ix, int iy)
{
    if (!im || !filename || frame < 0 || frame >= im->frames || ix < 0 || iy < 0 || ix > im->ox || iy > im->oy || max_value == 0.0 ||
        kind_elem_class(im->kind) == 0)
        return fail(DWTB200_EINVAL, "image_save_pgm: float or double image, a frame, a file name and a nonzero maximum");
    ImageScope scope(im);
    const size_t n = (size_t)ix * (size_t)iy;
    unsigned char *d = nullptr;
    std::vector<unsigned char> h(n ? n : 1);
    if (n) {
        CK(cudaMalloc((void **)&d, n));
        launch_pgm_quant(kind_elem_class(im->kind), frame_ptr(im, im->cur, frame), im->pitch, d, ix, iy, max_value, shift, shifted, g.st);
        cudaMemcpyAsync(h.data(), d, n, cudaMemcpyDeviceToHost, g.st);
        const cudaError_t e = cudaStreamSynchronize(g.st);
        cudaFree(d);
        if (e != cudaSuccess) return fail(DWTB200_ECUDA, "image_save_pgm: %s", cudaGetErrorString(e));
    }
    FILE *file = fopen(filename, "w");
    if (!file) return fail(DWTB200_EINVAL, "image_save_pgm: cannot open %s", filename);
    fprintf(file, "P2\n%i %i\n%i\n", ix, iy, 255);
    std::string out;
    out.reserve(n * 4);
    char buf[8];
    for (size_t i = 0; i < n; i++) {
        const int len = snprintf(buf, sizeof buf, "%i\n", (int)h[i]);
        out.append(buf, (size_t)len);
    }
    const bool ok = fwrite(out.data(), 1, out.size(), file) == out.size();
    fclose(file);
    return ok ? DWTB200_OK : fail(DWTB200_EINVAL, "image_save_pgm: error writing %s", filename);
}
int dwtb200_image_save_pgm(dwtb200_image *im, int frame, const char *filename, double max_value, int ix, int iy)
{
    API_LOCK();
    NEED_DEV();
    return save_pgm(im, frame, filename, max_value, 0.0, 0, ix, iy);
}
// dwt_util_save_sym_to_pgm_s (src/libdwt.c:26184): coefficients in [-max, +max] shifted by +max (in the image's own precision, as
// dwt_util_shift_s does) and written against 2 max
int dwtb200_image_save_sym_pgm(dwtb200_image *im, int frame, const char *filename, double max_value, int ix, int iy)
{
    API_LOCK();
    NEED_DEV();
    if (!(max_value > 0.0)) return fail(DWTB200_EINVAL, "image_save_sym_pgm: the maximum must be positive");
    const bool f32 = im && kind_elem_class(im->kind) == 1;
    const double twice = f32 ? (double)(2.f * (float)max_value) : 2.0 * max_value;
    return save_pgm(im, frame, filename, twice, max_value, 1, ix, iy);
}
// dwt_util_save_to_mat_s (src/libdwt.c:24430): the top-left ix x iy samples of a frame as text, "%f" separated by commas, one row per
// line.  Text output is host work; the samples are brought over once, packed.
int dwtb200_image_save_mat(dwtb200_image *im, int frame, const char *filename, int ix, int iy)
{
    API_LOCK();
    NEED_DEV();
    if (!im || !filename || frame < 0 || frame >= im->frames || ix < 0 || iy < 0 || ix > im->ox || iy > im->oy || kind_elem_class(im->kind) == 0)
        return fail(DWTB200_EINVAL, "image_save_mat: float or double image, a frame and a file name");
    ImageScope scope(im);
    const size_t n = (size_t)ix * (size_t)iy;
    std::vector<char> h(n ? n * im->es : 1);
    if (n) {
        CK(cudaMemcpy2DAsync(h.data(), (size_t)ix * im->es, frame_ptr(im, im->cur, frame), im->pitch * im->es, (size_t)ix * im->es, iy, cudaMemcpyDeviceToHost,
                             g.st));
        CK(cudaStreamSynchronize(g.st));
    }
    FILE *file = fopen(filename, "w");
    if (!file) return fail(DWTB200_EINVAL, "image_save_mat: cannot open %s", filename);
    std::string out;
    char buf[400];   // "%f" of a double can be 300+ characters
    for (int y = 0; y < iy; y++) {
        for (int x = 0; x < ix; x++) {
            const size_t i = (size_t)y * ix + x;
            const double v = im->es == 8 ? ((const double *)h.data())[i] : (double)((const float *)h.data())[i];
            const int len = snprintf(buf, sizeof buf, "%f", v);
            out.append(buf, (size_t)len);
            if (x + 1 != ix) out.push_back(',');
        }
        out.push_back('\n');
        if (out.size() > (1u << 20)) {
            if (fwrite(out.data(), 1, out.size(), file) != out.size()) {
                fclose(file);
                return fail(DWTB200_EINVAL, "image_save_mat: error writing %s", filename);
            }
            out.clear();
        }
    }
    const bool ok = fwrite(out.data(), 1, out.size(), file) == out.size();
    fclose(file);
    return ok ? DWTB200_OK : fail(DWTB200_EINVAL, "image_save_mat: error writing %s", filename);
}

int64_t dwtb200_image_diff(dwtb200_image *a, dwtb200_image *b)
{
    API_LOCK();
    if (g.dev < 0 || !a || !b || a->kind != b->kind || a->ox != b->ox || a->oy != b->oy || a->frames != b->frames) {
        fail(DWTB200_EINVAL, "image_diff: shape mismatch");
        return -1;
    }
    ImageScope scope(a);
    wait_for_image(g.st, b);
    unsigned long long *d = nullptr, h = 0;
    if (cudaMalloc(&d, 16) != cudaSuccess) return -1;
    cudaMemsetAsync(d, 0, 16, g.st);
    launch_compare((int)a->es, a->plane[a->cur], b->plane[b->cur], a->pitch, a->frame, a->ox, a->oy, a->frames,
                   kind_elem_class(a->kind), d, g.st);
    cudaMemcpyAsync(&h, d, 8, cudaMemcpyDeviceToHost, g.st);
    const cudaError_t e = cudaStreamSynchronize(g.st);
    cudaFree(d);
    if (e != cudaSuccess) {
        fail(DWTB200_ECUDA, "image_diff: %s", cudaGetErrorString(e));
        return -1;
    }
    return (int64_t)h;
}

double dwtb200_image_maxabs(dwtb200_image *a, dwtb200_image *b)
{
    API_LOCK();
    if (g.dev < 0 || !a || !b || a->kind != b->kind || a->ox != b->ox || a->oy != b->oy || a->frames != b->frames) {
        fail(DWTB200_EINVAL, "image_maxabs: shape mismatch");
        return -1.0;
    }
    ImageScope scope(a);
    wait_for_image(g.st, b);
    unsigned long long *d = nullptr, h[2] = {0, 0};
    if (cudaMalloc(&d, 16) != cudaSuccess) return -1.0;
    cudaMemsetAsync(d, 0, 16, g.st);
    launch_compare((int)a->es, a->plane[a->cur], b->plane[b->cur], a->pitch, a->frame, a->ox, a->oy, a->frames,
                   kind_elem_class(a->kind), d, g.st);
    cudaMemcpyAsync(h, d, 16, cudaMemcpyDeviceToHost, g.st);
    const cudaError_t e = cudaStreamSynchronize(g.st);
    cudaFree(d);
    if (e != cudaSuccess) {
        fail(DWTB200_ECUDA, "image_maxabs: %s", cudaGetErrorString(e));
        return -1.0;
    }
    double r;
    memcpy(&r, &h[1], 8);
    return r;
}

// =====================================================================================================
// host-memory entry points with the reference's semantics
// =====================================================================================================
namespace {
// device mirror of the caller's host image, cached per sample type between calls of the same shape (the
// reference mallocs its temps per call, src/libdwt.c:12801; here the planes and the captured graphs persist)
dwtb200_image *g_host_img[DWTB200_KIND_COUNT] = {};
cudaEvent_t g_t0 = nullptr, g_t1 = nullptr;
float g_last_ms = -1.f;

dwtb200_image *host_image(int kind, int ox, int oy)
{
    if (!kind_ok(kind)) {
        fail(DWTB200_EINVAL, "bad kind %d", kind);
        return nullptr;
    }
    dwtb200_image *&im = g_host_img[kind];
    if (im && im->ox == ox && im->oy == oy) return im;
    if (im) dwtb200_image_destroy(im);
    im = dwtb200_image_create(kind, ox, oy, 1);
    return im;
}

// ---- pipelined host path ---------------------------------------------------------------------------------
// A synchronous in-place call on host memory is PCIe-bound (2 x 256 MiB for an 8192^2 float image against
// ~0.2 ms of kernels), so the large dense case overlaps the two directions of the link: level 0 is run strip
// range by strip range while the image is still arriving, and each range's finished rows go back while the
// next range is uploaded.  Forward: the H subbands of level 0 (3/4 of the output) and of level 1 (3/16) leave early, then
// levels 2..J run on the LL band of level 1 and its quadrant follows.  Inverse: the LL quadrant goes first and is inverted down
// to level 1, then the level-0 subbands arrive range by range and the reconstructed rows leave.  The call is
// in place on the caller's buffer, so a download may only overwrite host rows whose old content has already
// been uploaded; the waits below encode exactly that.
struct Pipe {
    cudaStream_t up = nullptr, dn = nullptr, dn2 = nullptr, dn3 = nullptr;
    std::vector<cudaEvent_t> ev;
    cudaEvent_t get(size_t i)
    {
        while (ev.size() <= i) {
            cudaEvent_t e;
            cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
            ev.push_back(e);
        }
        return ev[i];
    }
} g_pipe;

// DWTB200_PIPE_TRACE=1: time stamps (ms after the call's start) of every stage of the pipelined host path, on stderr
struct PipeTrace {
    bool on = false;
    cudaEvent_t t0 = nullptr;
    struct Stamp { const char *what; int chunk; cudaEvent_t e; };
    std::vector<Stamp> stamps;
    void begin(cudaStream_t st)
    {
        static const bool want = getenv("DWTB200_PIPE_TRACE") && atoi(getenv("DWTB200_PIPE_TRACE")) != 0;
        on = want;
        if (!on) return;
        cudaEventCreate(&t0);
        cudaEventRecord(t0, st);
    }
    void mark(const char *what, int chunk, cudaStream_t st)
    {
        if (!on) return;
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, st);
        stamps.push_back({what, chunk, e});
    }
    void end()
    {
        if (!on) return;
        for (const Stamp &s : stamps) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, t0, s.e);
            fprintf(stderr, "pipe %-8s %2d %8.3f ms\n", s.what, s.chunk, ms);
            cudaEventDestroy(s.e);
        }
        cudaEventDestroy(t0);
        stamps.clear();
    }
};

bool pipeline_applies(const dwtb200_image *im, int64_t sx, int64_t sy, int ix, int iy, int J, DensePlan &pl)
{
    if (!g.pipeline || g.force_generic || im->frames != 1 || sy != (int64_t)im->es || ix != im->ox || iy != im->oy || J < 2) return false;
    if ((int64_t)im->ox * im->oy * (int64_t)im->es < ((int64_t)32 << 20) || sx < (int64_t)im->ox * (int64_t)im->es) return false;
    pl = dense_plan(im, J);
    return pl.jt > 1 && pl.type[0] == PLAN_STREAM;
}

int host_pipelined(bool inverse, dwtb200_image *im, char *host, int64_t sx, int J, const DensePlan &pl)
{
    if (!g_pipe.up) {
        CK(cudaStreamCreateWithFlags(&g_pipe.up, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&g_pipe.dn, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&g_pipe.dn2, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&g_pipe.dn3, cudaStreamNonBlocking));
    }
    const size_t es = im->es;
    const int W = im->ox, H = im->oy;
    const int nLx = (W + 1) >> 1, nLy = (H + 1) >> 1, nHy = H >> 1;
    const int nLx1 = (nLx + 1) >> 1, nLy1 = (nLy + 1) >> 1, nHy1 = nLy >> 1;   // level 1 works on the nLx x nLy LL band
    char *src_plane = (char *)im->plane[im->cur], *dst_plane = (char *)im->plane[im->cur ^ 1];
    const size_t dpitch = (size_t)im->pitch * es;
    LevelParams lp, lp1;
    memset(&lp1, 0, sizeof lp1);
    if (inverse) inv_level_params(im, 0, J, src_plane, dst_plane, lp);
    else fwd_level_params(im, 0, J, Band{src_plane, im->pitch, im->frame}, dst_plane, lp);
    // forward: level 1 is pipelined behind level 0 when it is a streaming level with levels beyond it, so that three quarters of the LL
    // quadrant leave while the image is still arriving (8192^2 float: 7.70 -> 7.35 ms).  The inverse keeps level 1 with the rest of
    // the pyramid: with level 1 in front of level 0 range by range the first rows leave after 0.8 instead of 1.6 ms, but every
    // reconstructed row overwrites coefficients of both levels, the uploads have to run further ahead and the call takes 7.6
    // instead of 7.35 ms (profiles/pipe_trace_r2.txt)
    const bool two = !inverse && J >= 3 && pl.jt >= 2 && pl.type[1] == PLAN_STREAM;
    if (two) fwd_level_params(im, 1, J, ll_band(im, 0), dst_plane, lp1);
    const int nstrips = lp.nstrips, pps = lp.pps;
    const int nstrips1 = two ? lp1.nstrips : 0, pps1 = two ? lp1.pps : 1;
    const int nch = nstrips < 16 ? nstrips : 16;
    // events: [0, nch) level kernels of chunk c done; nch: previous work on g.st done; nch + 1 + i: upload piece i done
    const size_t EV_PREV = (size_t)nch, EV_UP = (size_t)nch + 1;
    struct Piece {
        int r0, r1, c0, c1;
    };
    std::vector<Piece> pieces;   // host rectangles in upload order (rows, columns in elements)
    auto upload = [&](int r0, int r1, int c0, int c1, char *plane) -> cudaError_t {   // rows [r0,r1) x columns [c0,c1)
        if (r1 <= r0 || c1 <= c0) return cudaSuccess;
        const cudaError_t e = cudaMemcpy2DAsync(plane + (size_t)r0 * dpitch + (size_t)c0 * es, dpitch, host + (size_t)r0 * sx + (size_t)c0 * es, (size_t)sx,
                                                (size_t)(c1 - c0) * es, r1 - r0, cudaMemcpyDefault, g_pipe.up);
        if (e != cudaSuccess) return e;
        pieces.push_back(Piece{r0, r1, c0, c1});
        return cudaEventRecord(g_pipe.get(EV_UP + pieces.size() - 1), g_pipe.up);
    };
    // the call is in place: a download may only overwrite host memory whose old content has been uploaded.  Index of the last upload
    // piece that overlaps the rectangle (-1: none); uploads are one stream, so waiting for that piece's event covers the earlier ones
    auto landing = [&](int r0, int r1, int c0, int c1) -> int {
        for (int i = (int)pieces.size() - 1; i >= 0; i--)
            if (pieces[i].r0 < r1 && r0 < pieces[i].r1 && pieces[i].c0 < c1 && c0 < pieces[i].c1) return i;
        return -1;
    };
    // download of rows [r0,r1) x columns [c0,c1) on stream `on`, after the kernels of chunk c and the uploads it would overwrite
    auto download = [&](int c, int r0, int r1, int c0, int c1, char *plane, cudaStream_t on, int have_piece) -> cudaError_t {
        if (r1 <= r0 || c1 <= c0) return cudaSuccess;
        cudaError_t e = cudaStreamWaitEvent(on, g_pipe.get((size_t)c), 0);
        if (e != cudaSuccess) return e;
        const int k = landing(r0, r1, c0, c1);
        if (k > have_piece && (e = cudaStreamWaitEvent(on, g_pipe.get(EV_UP + (size_t)k), 0)) != cudaSuccess) return e;
        return cudaMemcpy2DAsync(host + (size_t)r0 * sx + (size_t)c0 * es, (size_t)sx, plane + (size_t)r0 * dpitch + (size_t)c0 * es, dpitch,
                                 (size_t)(c1 - c0) * es, r1 - r0, cudaMemcpyDefault, on);
    };
    std::vector<int> s_lo(nch + 1), up_last(nch, -1);
    for (int c = 0; c <= nch; c++) s_lo[c] = (int)((int64_t)nstrips * c / nch);
    CK(cudaEventRecord(g_pipe.get(EV_PREV), g.st));
    CK(cudaStreamWaitEvent(g_pipe.up, g_pipe.get(EV_PREV), 0));
    CK(cudaStreamWaitEvent(g_pipe.dn, g_pipe.get(EV_PREV), 0));
    CK(cudaStreamWaitEvent(g_pipe.dn2, g_pipe.get(EV_PREV), 0));
    CK(cudaStreamWaitEvent(g_pipe.dn3, g_pipe.get(EV_PREV), 0));
    CK(cudaEventRecord(g_t0, g.st));
    PipeTrace tr;
    tr.begin(g.st);
    auto range = [&](const LevelParams &of, int s0, int s1) {
        LevelParams q = of;
        q.strip0 = s0;
        q.nstrips = s1 - s0;
        if (inverse) stream_inv(im->kind, q, 1, g.st);
        else stream_fwd(im->kind, q, 1, g.st);
    };

    if (!inverse) {
        int row = 0;
        for (int c = 0; c < nch; c++) {
            const int k1 = std::min(s_lo[c + 1] * pps, nLy);
            const int r1 = (c == nch - 1) ? H : std::min(H, 2 * k1 + 4);   // the range's kernel reads rows up to 2 k1 + 2
            CK(upload(row, r1, 0, W, src_plane));
            row = std::max(row, r1);
            up_last[c] = (int)pieces.size() - 1;
            tr.mark("up", c, g_pipe.up);
        }
        int s1_done = 0;   // level-1 strips launched so far
        for (int c = 0; c < nch; c++) {
            const int k0 = s_lo[c] * pps, k1 = std::min(s_lo[c + 1] * pps, nLy), kh = std::min(k1, nHy);
            CK(cudaStreamWaitEvent(g.st, g_pipe.get(EV_UP + (size_t)up_last[c]), 0));
            range(lp, s_lo[c], s_lo[c + 1]);
            // level 1 on the LL rows that are complete now: a strip range [.., e) reads LL rows up to 2 e pps1 + 2
            int s1_hi = s1_done;
            if (two) {
                if (c == nch - 1) s1_hi = nstrips1;
                else
                    while (s1_hi < nstrips1 && std::min(2 * (s1_hi + 1) * pps1 + 4, nLy) <= k1) s1_hi++;
                if (s1_hi > s1_done) range(lp1, s1_done, s1_hi);
            }
            CK(cudaEventRecord(g_pipe.get((size_t)c), g.st));
            tr.mark("kernel", c, g.st);
            CK(download(c, k0, k1, nLx, W, dst_plane, g_pipe.dn, up_last[c]));   // HL rows: these host rows were uploaded before the kernel ran
            tr.mark("dn-HL", c, g_pipe.dn);
            // LH | HH rows land in host rows [nLy + k0, nLy + kh): not before those were uploaded -- on a stream of their own, so that the
            // rows of the following ranges that may land at once do not queue up behind that wait
            CK(download(c, nLy + k0, nLy + kh, 0, W, dst_plane, g_pipe.dn2, up_last[c]));
            tr.mark("dn-LHHH", c, g_pipe.dn2);
            if (s1_hi > s1_done) {   // the same for the level-1 subbands inside the LL quadrant
                const int j0 = s1_done * pps1, j1 = std::min(s1_hi * pps1, nLy1), jh = std::min(j1, nHy1);
                CK(download(c, j0, j1, nLx1, nLx, dst_plane, g_pipe.dn, up_last[c]));
                CK(download(c, nLy1 + j0, nLy1 + jh, 0, nLx, dst_plane, g_pipe.dn3, up_last[c]));
                tr.mark("dn-lvl1", c, g_pipe.dn3);
                s1_done = s1_hi;
            }
        }
        {   // the levels behind the pipelined ones on their LL band (stream order after the last range)
            const int rr = run_dense_uncaptured(im, false, J, pl, two ? 2 : 1);
            if (rr) return rr;
        }
        CK(cudaGetLastError());
        CK(cudaEventRecord(g_t1, g.st));
        tr.mark("levels", 0, g.st);
        CK(cudaStreamWaitEvent(g_pipe.dn, g_t1, 0));
        if (two) CK(cudaMemcpy2DAsync(host, (size_t)sx, dst_plane, dpitch, (size_t)nLx1 * es, nLy1, cudaMemcpyDefault, g_pipe.dn));
        else CK(cudaMemcpy2DAsync(host, (size_t)sx, dst_plane, dpitch, (size_t)nLx * es, nLy, cudaMemcpyDefault, g_pipe.dn));
        tr.mark("dn-LL", 0, g_pipe.dn);
    } else {
        // the LL quadrant first, inverted down to level 1
        CK(upload(0, nLy, 0, nLx, src_plane));
        CK(cudaStreamWaitEvent(g.st, g_pipe.get(EV_UP), 0));
        {
            const int rr = run_dense_uncaptured(im, true, J, pl, 1);
            if (rr) return rr;
        }
        CK(cudaGetLastError());
        const int units = (H >> 1) + 1;
        // uploads of chunk c: the subband rows its kernel reads (< q1 + 2), and the coefficients its reconstructed rows [.., r1) will
        // overwrite (host row r holds HL row r, or LH|HH row r - nLy) -- the HL rows therefore run ahead at twice the pace -- so that
        // no download has to wait for a later chunk (8.35 -> 7.25 ms)
        int hl0 = 0, hh0 = 0;
        auto upto = [](int cur, int limit, int a, int b) { return std::max(cur, std::min(limit, std::max(a, b))); };
        for (int c = 0; c < nch; c++) {
            const bool last = c == nch - 1;
            const int q1 = std::min(s_lo[c + 1] * pps, units), r1 = last ? H : std::min(H, 2 * q1 - 1);
            const int a = last ? nLy : upto(hl0, nLy, q1 + 2, r1), b = last ? nHy : upto(hh0, nHy, q1 + 2, r1 - nLy);
            CK(upload(hl0, a, nLx, W, src_plane));
            CK(upload(nLy + hh0, nLy + b, 0, W, src_plane));
            hl0 = a;
            hh0 = b;
            up_last[c] = (int)pieces.size() - 1;
            tr.mark("up", c, g_pipe.up);
        }
        for (int c = 0; c < nch; c++) {
            const int q0 = s_lo[c] * pps, q1 = std::min(s_lo[c + 1] * pps, units);
            CK(cudaStreamWaitEvent(g.st, g_pipe.get(EV_UP + (size_t)up_last[c]), 0));
            range(lp, s_lo[c], s_lo[c + 1]);
            CK(cudaEventRecord(g_pipe.get((size_t)c), g.st));
            tr.mark("kernel", c, g.st);
            const int r0 = std::max(0, 2 * q0 - 1), r1 = std::min(H, 2 * q1 - 1);   // rows this range reconstructs
            CK(download(c, r0, r1, 0, W, dst_plane, g_pipe.dn, up_last[c]));
            tr.mark("dn", c, g_pipe.dn);
        }
        CK(cudaEventRecord(g_t1, g.st));
    }
    CK(cudaStreamSynchronize(g_pipe.dn));
    CK(cudaStreamSynchronize(g_pipe.dn2));
    CK(cudaStreamSynchronize(g_pipe.dn3));
    CK(cudaStreamSynchronize(g_pipe.up));
    CK(cudaStreamSynchronize(g.st));
    CK(cudaGetLastError());
    tr.end();
    CK(cudaEventElapsedTime(&g_last_ms, g_t0, g_t1));
    im->cur ^= 1;
    im->last_path = 0;
    return DWTB200_OK;
}

int host_transform(bool inverse, int kind, void *ptr, int64_t sx, int64_t sy, int ox, int oy, int ix, int iy, int *j_io,
                   int decompose_one, int zero_padding)
{
    dwtb200_image *im = host_image(kind, ox, oy);
    if (!im) return DWTB200_ENOMEM;
    ImageScope scope(im);
    if (!g_t0) {
        CK(cudaEventCreate(&g_t0));
        CK(cudaEventCreate(&g_t1));
    }
    if (ix >= 1 && iy >= 1 && ix <= ox && iy <= oy) {
        const int J = dwtb200_clamp_j(*j_io, ox, oy, decompose_one);
        DensePlan pl;
        if (pipeline_applies(im, sx, sy, ix, iy, J, pl)) {
            if (!inverse) *j_io = J;
            return host_pipelined(inverse, im, (char *)ptr, sx, J, pl);
        }
    }
    int r = dwtb200_image_upload(im, 0, ptr, sx, sy);
    if (r) return r;
    CK(cudaEventRecord(g_t0, g.st));
    r = inverse ? dwtb200_image_inv2(im, ix, iy, *j_io, decompose_one, zero_padding)
                : dwtb200_image_fwd2(im, ix, iy, j_io, decompose_one, zero_padding);
    if (r) return r;
    CK(cudaEventRecord(g_t1, g.st));
    r = dwtb200_image_download(im, 0, ptr, sx, sy);   // synchronises the stream
    if (r) return r;
    CK(cudaEventElapsedTime(&g_last_ms, g_t0, g_t1));
    return DWTB200_OK;
}
}  // namespace

double dwtb200_last_transform_ms(void) { return (double)g_last_ms; }
static void release_host_volume();
void dwtb200_release_host_cache(void)
{
    API_LOCK();
    // the pipelined host path's streams and events, and the transform timing events, go with the cache
    for (cudaEvent_t e : g_pipe.ev) cudaEventDestroy(e);
    g_pipe.ev.clear();
    for (cudaStream_t *s : {&g_pipe.up, &g_pipe.dn, &g_pipe.dn2, &g_pipe.dn3}) {
        if (*s) cudaStreamDestroy(*s);
        *s = nullptr;
    }
    if (g_t0) {
        cudaEventDestroy(g_t0);
        cudaEventDestroy(g_t1);
        g_t0 = g_t1 = nullptr;
    }
    for (int k = 0; k < DWTB200_KIND_COUNT; k++) {
        if (g_host_img[k]) dwtb200_image_destroy(g_host_img[k]);
        g_host_img[k] = nullptr;
    }
    release_host_volume();
}

int dwtb200_fwd2_host(int kind, void *ptr, int64_t sx, int64_t sy, int ox, int oy, int ix, int iy, int *j_max_ptr,
                      int decompose_one, int zero_padding)
{
    API_LOCK();
    NEED_DEV();
    if (!ptr || !j_max_ptr) return fail(DWTB200_EINVAL, "fwd2_host: null argument");
    return host_transform(false, kind, ptr, sx, sy, ox, oy, ix, iy, j_max_ptr, decompose_one, zero_padding);
}

int dwtb200_inv2_host(int kind, void *ptr, int64_t sx, int64_t sy, int ox, int oy, int ix, int iy, int j_max,
                      int decompose_one, int zero_padding)
{
    API_LOCK();
    NEED_DEV();
    if (!ptr) return fail(DWTB200_EINVAL, "inv2_host: null argument");
    return host_transform(true, kind, ptr, sx, sy, ox, oy, ix, iy, &j_max, decompose_one, zero_padding);
}

// The in-place family on host memory.  Its level geometry comes from the INNER size alone and the outer size only
// bounds the level count (src/libdwt.c:12948, 12964), so the device works on an image of the inner size.
static int host_inplace(bool inverse, int kind, void *ptr, int64_t sx, int64_t sy, int ox, int oy, int ix, int iy, int *j_io, int decompose_one)
{
    if (ix < 1 || iy < 1 || ix > ox || iy > oy) return fail(DWTB200_EINVAL, "inner size %d x %d outside outer %d x %d", ix, iy, ox, oy);
    const int J = dwtb200_clamp_j(*j_io, ox, oy, decompose_one);
    if (!inverse) *j_io = J;
    dwtb200_image *im = host_image(kind, ix, iy);
    if (!im) return DWTB200_ENOMEM;
    ImageScope scope(im);
    if (!g_t0) {
        CK(cudaEventCreate(&g_t0));
        CK(cudaEventCreate(&g_t1));
    }
    int r = dwtb200_image_upload(im, 0, ptr, sx, sy);
    if (r) return r;
    CK(cudaEventRecord(g_t0, g.st));
    r = inplace_transform(im, inverse, J);
    if (r) return r;
    CK(cudaEventRecord(g_t1, g.st));
    r = dwtb200_image_download(im, 0, ptr, sx, sy);   // synchronises the stream
    if (r) return r;
    CK(cudaEventElapsedTime(&g_last_ms, g_t0, g_t1));
    return DWTB200_OK;
}

int dwtb200_fwd2_inplace_host(int kind, void *ptr, int64_t sx, int64_t sy, int ox, int oy, int ix, int iy, int *j_max_ptr, int decompose_one)
{
    API_LOCK();
    NEED_DEV();
    if (!ptr || !j_max_ptr) return fail(DWTB200_EINVAL, "fwd2_inplace_host: null argument");
    return host_inplace(false, kind, ptr, sx, sy, ox, oy, ix, iy, j_max_ptr, decompose_one);
}

int dwtb200_inv2_inplace_host(int kind, void *ptr, int64_t sx, int64_t sy, int ox, int oy, int ix, int iy, int j_max, int decompose_one)
{
    API_LOCK();
    NEED_DEV();
    if (!ptr) return fail(DWTB200_EINVAL, "inv2_inplace_host: null argument");
    return host_inplace(true, kind, ptr, sx, sy, ox, oy, ix, iy, &j_max, decompose_one);
}

// Out of place (dwt_cdf97_2f_s2 / 2i_s2, src/libdwt.c:12619, 17985): the first pass of the first level reads `src`
// and writes `dst`, everything after that runs in place on `dst` (src/libdwt.c:12700 "src = dst"); the inverse copies
// the inner region of `src` into `dst` first (:18000).  Both are therefore the in-place transform of `dst` with the
// region that first pass reads replaced by `src` -- which is what is done here, on the device copy of `dst`.
static int host_transform2(bool inverse, int kind, const void *src, void *dst, int64_t sx, int64_t sy, int ox, int oy, int ix,
                           int iy, int *j_io, int decompose_one, int zero_padding)
{
    if (ix < 1 || iy < 1 || ix > ox || iy > oy) return fail(DWTB200_EINVAL, "inner size %d x %d outside outer %d x %d", ix, iy, ox, oy);
    const int J = dwtb200_clamp_j(*j_io, ox, oy, decompose_one);
    if (!inverse) *j_io = J;
    int rw = ix, rh = iy;   // region of `src` that reaches `dst`
    if (!inverse) {
        if (J == 0) return DWTB200_OK;                  // no level: dst is not touched at all
        if (!guard(kind) || ox > 1) rh = oy;            // row pass first: every outer row, inner columns (:12680)
        else rw = ox;                                   // row pass skipped: the column pass reads src (:12713)
    }
    dwtb200_image *im = host_image(kind, ox, oy);
    if (!im) return DWTB200_ENOMEM;
    ImageScope scope(im);
    int r = DWTB200_OK;
    const bool covers = rw == ox && rh == oy;
    if (inverse || covers) {   // the region replaces dst's content before anything is computed: overlay, then in place
        if (!covers) r = dwtb200_image_upload(im, 0, dst, sx, sy);
        if (!r) r = upload_region(im, 0, src, sx, sy, rw, rh);
        if (!r) r = inverse ? dwtb200_image_inv2(im, ix, iy, J, decompose_one, zero_padding)
                            : dwtb200_image_fwd2(im, ix, iy, j_io, decompose_one, zero_padding);
    } else {
        // sparse forward: the first pass writes only its L / H outputs into dst, whose other samples keep their content
        r = dwtb200_image_upload(im, 0, dst, sx, sy);
        if (!r) {
            im->cur ^= 1;
            r = upload_region(im, 0, src, sx, sy, rw, rh);   // the source samples go to the other plane
            im->cur ^= 1;
        }
        if (!r) {
            s2_level0(im, ix, iy);
            if (zero_padding) fwd_zero_padding(im, ix, iy, 0);
            run_fwd_generic(im, ix, iy, J, zero_padding, 1);
            im->last_path = 1;
            CK(cudaGetLastError());
        }
    }
    if (!r) r = dwtb200_image_download(im, 0, dst, sx, sy);
    return r;
}
int dwtb200_fwd2_host2(int kind, const void *src, void *dst, int64_t sx, int64_t sy, int ox, int oy, int ix, int iy, int *j_max_ptr,
                       int decompose_one, int zero_padding)
{
    API_LOCK();
    NEED_DEV();
    if (!src || !dst || !j_max_ptr) return fail(DWTB200_EINVAL, "fwd2_host2: null argument");
    return host_transform2(false, kind, src, dst, sx, sy, ox, oy, ix, iy, j_max_ptr, decompose_one, zero_padding);
}
int dwtb200_inv2_host2(int kind, const void *src, void *dst, int64_t sx, int64_t sy, int ox, int oy, int ix, int iy, int j_max,
                       int decompose_one, int zero_padding)
{
    API_LOCK();
    NEED_DEV();
    if (!src || !dst) return fail(DWTB200_EINVAL, "inv2_host2: null argument");
    return host_transform2(true, kind, src, dst, sx, sy, ox, oy, ix, iy, &j_max, decompose_one, zero_padding);
}

// The reference's performance protocol (dwt_util_perf_cdf97_2_s / dwt_util_perf_cdf53_2_i, src/libdwt.c:21391, 21262)
// on the device: M images filled with the test pattern, N loops of "M forward transforms, then M inverse transforms",
// the minimum over the loops of the mean time per transform -- timed with CUDA events on the library stream instead
// of dwt_util_get_clock, the images resident in HBM (M of them cycled, like the reference's cache flush).
int dwtb200_perf2(int kind, int ox, int oy, int ix, int iy, int j_max, int decompose_one, int zero_padding, int M, int N,
                  float *fwd_secs, float *inv_secs)
{
    API_LOCK();
    NEED_DEV();
    if (M < 1 || N < 1 || !fwd_secs || !inv_secs) return fail(DWTB200_EINVAL, "perf2: bad arguments");
    std::vector<dwtb200_image *> imgs;
    int r = DWTB200_OK;
    for (int m = 0; m < M && !r; m++) {
        dwtb200_image *im = dwtb200_image_create(kind, ox, oy, 1);
        if (!im) r = DWTB200_ENOMEM;
        else {
            imgs.push_back(im);
            r = dwtb200_image_fill(im, 0, 0, 0);
        }
    }
    float best_f = 1e30f, best_i = 1e30f;
    std::vector<int> j(M, j_max);
    for (int n = -1; n < N && !r; n++) {   // loop -1 warms up (graph capture)
        float ms = 0;
        if (!r) r = dwtb200_timer_start();
        for (int m = 0; m < M && !r; m++) {
            j[m] = j_max;
            r = dwtb200_image_fwd2(imgs[m], ix, iy, &j[m], decompose_one, zero_padding);
        }
        if (!r) {
            ms = (float)dwtb200_timer_stop_ms();
            if (n >= 0 && ms / M < best_f) best_f = ms / M;
        }
        if (!r) r = dwtb200_timer_start();
        for (int m = 0; m < M && !r; m++) r = dwtb200_image_inv2(imgs[m], ix, iy, j[m], decompose_one, zero_padding);
        if (!r) {
            ms = (float)dwtb200_timer_stop_ms();
            if (n >= 0 && ms / M < best_i) best_i = ms / M;
        }
    }
    for (dwtb200_image *im : imgs) dwtb200_image_destroy(im);
    *fwd_secs = best_f * 1e-3f;
    *inv_secs = best_i * 1e-3f;
    return r;
}

// =====================================================================================================
// 3-D, one level, interleaved subbands
// dwt_util_perf_cdf97_2_inplace_s and its _sep / _sdl twins (src/libdwt.h:2520-2590): the protocol of dwtb200_perf2 for the
// in-place family.  Level sizes come from the inner size, the level count from the outer one.
int dwtb200_perf2_inplace(int kind, int ox, int oy, int ix, int iy, int j_max, int decompose_one, int M, int N, float *fwd_secs,
                          float *inv_secs)
{
    API_LOCK();
    NEED_DEV();
    if (M < 1 || N < 1 || !fwd_secs || !inv_secs || ix < 1 || iy < 1 || ix > ox || iy > oy) return fail(DWTB200_EINVAL, "perf2_inplace: bad arguments");
    const int J = dwtb200_clamp_j(j_max, ox, oy, decompose_one);
    std::vector<dwtb200_image *> imgs;
    int r = DWTB200_OK;
    for (int m = 0; m < M && !r; m++) {
        dwtb200_image *im = dwtb200_image_create(kind, ix, iy, 1);
        if (!im) r = DWTB200_ENOMEM;
        else {
            imgs.push_back(im);
            r = dwtb200_image_fill(im, 0, 0, 0);
        }
    }
    float best_f = 1e30f, best_i = 1e30f;
    for (int n = -1; n < N && !r; n++) {   // loop -1 warms up (graph capture)
        for (int inverse = 0; inverse < 2 && !r; inverse++) {
            r = dwtb200_timer_start();
            for (int m = 0; m < M && !r; m++) {
                ImageScope scope(imgs[m]);
                r = inplace_transform(imgs[m], inverse != 0, J);
            }
            if (!r) {
                const float ms = (float)dwtb200_timer_stop_ms();
                float &best = inverse ? best_i : best_f;
                if (n >= 0 && ms / M < best) best = ms / M;
            }
        }
    }
    for (dwtb200_image *im : imgs) dwtb200_image_destroy(im);
    *fwd_secs = best_f * 1e-3f;
    *inv_secs = best_i * 1e-3f;
    return r;
}

// =====================================================================================================
dwtb200_volume *dwtb200_volume_create(int nx, int ny, int nz)
{
    API_LOCK();
    if (g.dev < 0 && dwtb200_init(-1)) return nullptr;
    if (nx < 1 || ny < 1 || nz < 1) {
        fail(DWTB200_EINVAL, "volume_create: bad size");
        return nullptr;
    }
    dwtb200_volume *v = new dwtb200_volume;
    v->nx = nx;
    v->ny = ny;
    v->nz = nz;
    v->pitch = align_up(nx, 32);
    v->slice = v->pitch * ny;
    const size_t bytes = (size_t)v->slice * nz * sizeof(float);
    if (cudaMalloc(&v->buf[0], bytes) != cudaSuccess || cudaMalloc(&v->buf[1], bytes) != cudaSuccess) {
        fail(DWTB200_ENOMEM, "volume_create(%d,%d,%d): %s", nx, ny, nz, cudaGetErrorString(cudaGetLastError()));
        dwtb200_volume_destroy(v);
        return nullptr;
    }
    return v;
}
void dwtb200_volume_destroy(dwtb200_volume *v)
{
    API_LOCK();
    if (!v) return;
    if (g.st) cudaStreamSynchronize(g.st);
    for (int i = 0; i < 2; i++)
        if (v->buf[i]) cudaFree(v->buf[i]);
    delete v;
}

static int volume_xfer(dwtb200_volume *v, void *host, size_t sx, size_t sy, size_t sz, bool up)
{
    if (!v || !host) return fail(DWTB200_EINVAL, "volume transfer: null argument");
    if (sx != sizeof(float)) return fail(DWTB200_EINVAL, "volume transfer: stride_x must be sizeof(float)");
    cudaMemcpy3DParms p;
    memset(&p, 0, sizeof p);
    // pitched pointers: host rows are sy bytes apart, slices sz bytes apart (sz must be a multiple of sy
    // for cudaMemcpy3D; otherwise fall back to one 2-D copy per slice)
    float *d = v->buf[v->cur];
    if (sz % sy == 0) {
        p.srcPtr = up ? make_cudaPitchedPtr(host, sy, v->nx, sz / sy) : make_cudaPitchedPtr(d, v->pitch * 4, v->nx, v->ny);
        p.dstPtr = up ? make_cudaPitchedPtr(d, v->pitch * 4, v->nx, v->ny) : make_cudaPitchedPtr(host, sy, v->nx, sz / sy);
        p.extent = make_cudaExtent((size_t)v->nx * 4, v->ny, v->nz);
        p.kind = up ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost;
        CK(cudaMemcpy3DAsync(&p, g.st));
    } else {
        for (int z = 0; z < v->nz; z++) {
            char *hp = (char *)host + (size_t)z * sz;
            float *dp = d + (size_t)z * v->slice;
            if (up) CK(cudaMemcpy2DAsync(dp, v->pitch * 4, hp, sy, (size_t)v->nx * 4, v->ny, cudaMemcpyHostToDevice, g.st));
            else CK(cudaMemcpy2DAsync(hp, sy, dp, v->pitch * 4, (size_t)v->nx * 4, v->ny, cudaMemcpyDeviceToHost, g.st));
        }
    }
    if (!up) CK(cudaStreamSynchronize(g.st));
    return DWTB200_OK;
}
int dwtb200_volume_upload(dwtb200_volume *v, const void *host, size_t sx, size_t sy, size_t sz)
{
    API_LOCK();
    NEED_DEV();
    return volume_xfer(v, (void *)host, sx, sy, sz, true);
}
int dwtb200_volume_download(dwtb200_volume *v, void *host, size_t sx, size_t sy, size_t sz)
{
    API_LOCK();
    NEED_DEV();
    return volume_xfer(v, host, sx, sy, sz, false);
}
int dwtb200_volume_fill(dwtb200_volume *v)
{
    API_LOCK();
    NEED_DEV();
    if (!v) return fail(DWTB200_EINVAL, "volume_fill: null");
    launch_volume_fill(v->buf[v->cur], v->pitch, v->slice, v->nx, v->ny, v->nz, g.st);
    CK(cudaGetLastError());
    return DWTB200_OK;
}

static int volume_axes(dwtb200_volume *v, int inverse)
{
    if (v->nx >= 2 && v->ny >= 2 && v->nz >= 2 && !g.force_generic) {
        // x and y fused per slice (buf[cur] -> buf[cur^1]), then z (buf[cur^1] -> buf[cur]): two passes over the volume
        VolParams p;
        memset(&p, 0, sizeof p);
        p.nx = v->nx;
        p.ny = v->ny;
        p.nz = v->nz;
        p.s_pitch = p.d_pitch = v->pitch;
        p.s_slice = p.d_slice = v->slice;
        p.src = v->buf[v->cur];
        p.dst = v->buf[v->cur ^ 1];
        if (g.vol3 && vol3_applies(p)) {   // all three axes in one pass: buf[cur] -> buf[cur^1]
            launch_vol3(p, inverse, g.vol3, g.sm_count, g.st);
            v->cur ^= 1;
            CK(cudaGetLastError());
            return DWTB200_OK;
        }
        launch_vol_xy(p, inverse, g.sm_count, g.st);
        p.src = v->buf[v->cur ^ 1];
        p.dst = v->buf[v->cur];
        launch_vol_z(p, inverse, g.sm_count, g.st);
        CK(cudaGetLastError());
        return DWTB200_OK;
    }
    // x, then y, then z in both directions (src/volume-dwt.c:727-770, 1115-1150); each pass is out of
    // place between the two buffers.  The reference's inverse skips an axis of size 1 (libdwt.c:17182).
    for (int axis = 0; axis < 3; axis++) {
        Axis3Params p;
        p.src = v->buf[v->cur];
        p.dst = v->buf[v->cur ^ 1];
        const int64_t sx = 1, sy = v->pitch, sz = v->slice;
        if (axis == 0) { p.n0 = v->ny; p.n1 = v->nz; p.N = v->nx; p.s_line0 = sy; p.s_line1 = sz; p.s_elem = sx; }
        else if (axis == 1) { p.n0 = v->nx; p.n1 = v->nz; p.N = v->ny; p.s_line0 = sx; p.s_line1 = sz; p.s_elem = sy; }
        else { p.n0 = v->nx; p.n1 = v->ny; p.N = v->nz; p.s_line0 = sx; p.s_line1 = sy; p.s_elem = sz; }
        p.d_line0 = p.s_line0;
        p.d_line1 = p.s_line1;
        p.d_elem = p.s_elem;
        launch_axis3(p, inverse, g.st);
        v->cur ^= 1;
    }
    CK(cudaGetLastError());
    return DWTB200_OK;
}
int dwtb200_volume_fwd3(dwtb200_volume *v)
{
    API_LOCK();
    NEED_DEV();
    if (!v) return fail(DWTB200_EINVAL, "volume_fwd3: null");
    if (v->nx < 5 || v->ny < 5 || v->nz < 5) return fail(DWTB200_EINVAL, "volume_fwd3: every size must be >= 5 (src/dwt-simple.c:2172)");
    return volume_axes(v, 0);
}
int dwtb200_volume_inv3(dwtb200_volume *v)
{
    API_LOCK();
    NEED_DEV();
    if (!v) return fail(DWTB200_EINVAL, "volume_inv3: null");
    return volume_axes(v, 1);
}

// the device volume of the *_host calls is kept between calls of the same shape (two planes of the volume's size: allocating and
// freeing 8.6 GB per call of a 1024^3 volume cost more than the transform and a good part of the copies); dwtb200_release_host_cache frees it
static dwtb200_volume *g_host_vol = nullptr;
static void release_host_volume()
{
    if (g_host_vol) dwtb200_volume_destroy(g_host_vol);
    g_host_vol = nullptr;
}
static dwtb200_volume *host_volume(int nx, int ny, int nz)
{
    if (g_host_vol && g_host_vol->nx == nx && g_host_vol->ny == ny && g_host_vol->nz == nz) return g_host_vol;
    if (g_host_vol) dwtb200_volume_destroy(g_host_vol);
    g_host_vol = dwtb200_volume_create(nx, ny, nz);
    return g_host_vol;
}
// Large volumes through the *_host calls: the z ranges of the one-pass kernel are launched one by one while the volume is still
// arriving, and every range's finished slices leave while the next range is uploaded (the 3-D transforms leave every coefficient
// where its sample was -- interleaved subbands -- so a download never lands on slices that have not been uploaded).  1024^3:
// upload, transform, download one after the other 154 ms; pipelined, both directions of the link at once.
static bool host3_pipeline_applies(const dwtb200_volume *v, size_t sx_src, size_t sx_dst)
{
    if (!g.pipeline || g.vol3 == 0 || g.force_generic || sx_src != sizeof(float) || sx_dst != sizeof(float)) return false;
    VolParams p;
    memset(&p, 0, sizeof p);
    p.nx = v->nx;
    p.ny = v->ny;
    p.nz = v->nz;
    return vol3_applies(p) && v->nz >= 64 && (size_t)v->nx * v->ny * v->nz * sizeof(float) >= ((size_t)64 << 20);
}
static int host3_pipelined(bool inverse, dwtb200_volume *v, const char *src, size_t ssy, size_t ssz, char *dst, size_t dsy, size_t dsz)
{
    if (!g_pipe.up) {
        CK(cudaStreamCreateWithFlags(&g_pipe.up, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&g_pipe.dn, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&g_pipe.dn2, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&g_pipe.dn3, cudaStreamNonBlocking));
    }
    const int nz = v->nz, units = inverse ? (nz >> 1) + 1 : (nz + 1) >> 1;
    const int want = std::min(16, std::max(1, units / 8)), pps = (units + want - 1) / want, nr = (units + pps - 1) / pps;
    VolParams p;
    memset(&p, 0, sizeof p);
    p.nx = v->nx;
    p.ny = v->ny;
    p.nz = nz;
    p.s_pitch = p.d_pitch = v->pitch;
    p.s_slice = p.d_slice = v->slice;
    p.src = v->buf[v->cur];
    p.dst = v->buf[v->cur ^ 1];
    // slices [z0, z1) between the host volume (rows hy, slices hz bytes apart) and a device plane
    auto slab = [&](char *host, size_t hy, size_t hz, float *dev, int z0, int z1, bool up, cudaStream_t on) -> cudaError_t {
        if (z1 <= z0) return cudaSuccess;
        char *hp = host + (size_t)z0 * hz;
        float *dp = dev + (size_t)z0 * v->slice;
        if (hz % hy == 0) {
            cudaMemcpy3DParms c;
            memset(&c, 0, sizeof c);
            const cudaPitchedPtr h = make_cudaPitchedPtr(hp, hy, v->nx, hz / hy), d = make_cudaPitchedPtr(dp, v->pitch * sizeof(float), v->nx, v->ny);
            c.srcPtr = up ? h : d;
            c.dstPtr = up ? d : h;
            c.extent = make_cudaExtent((size_t)v->nx * sizeof(float), v->ny, z1 - z0);
            c.kind = up ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost;
            return cudaMemcpy3DAsync(&c, on);
        }
        for (int z = z0; z < z1; z++, hp += hz, dp += v->slice) {
            const cudaError_t e = up ? cudaMemcpy2DAsync(dp, v->pitch * sizeof(float), hp, hy, (size_t)v->nx * sizeof(float), v->ny, cudaMemcpyHostToDevice, on)
                                     : cudaMemcpy2DAsync(hp, hy, dp, v->pitch * sizeof(float), (size_t)v->nx * sizeof(float), v->ny, cudaMemcpyDeviceToHost, on);
            if (e != cudaSuccess) return e;
        }
        return cudaSuccess;
    };
    // events: [0, nr) upload for range c done; [nr, 2 nr) kernel of range c done; 2 nr: previous work on g.st
    CK(cudaEventRecord(g_pipe.get(2 * (size_t)nr), g.st));
    CK(cudaStreamWaitEvent(g_pipe.up, g_pipe.get(2 * (size_t)nr), 0));
    CK(cudaStreamWaitEvent(g_pipe.dn, g_pipe.get(2 * (size_t)nr), 0));
    int zup = 0;
    for (int c = 0; c < nr; c++) {   // range c reads the slices up to 2 k1 + 2 (mirrored into the volume at its end)
        const int k1 = std::min((c + 1) * pps, units), need = (c == nr - 1) ? nz : std::min(nz, 2 * k1 + 4);
        CK(slab(const_cast<char *>(src), ssy, ssz, v->buf[v->cur], zup, need, true, g_pipe.up));
        zup = std::max(zup, need);
        CK(cudaEventRecord(g_pipe.get((size_t)c), g_pipe.up));
    }
    for (int c = 0; c < nr; c++) {
        const int k0 = c * pps, k1 = std::min((c + 1) * pps, units);
        CK(cudaStreamWaitEvent(g.st, g_pipe.get((size_t)c), 0));
        launch_vol3_ranges(p, inverse ? 1 : 0, g.vol3, pps, c, 1, g.st);
        CK(cudaGetLastError());
        CK(cudaEventRecord(g_pipe.get((size_t)(nr + c)), g.st));
        CK(cudaStreamWaitEvent(g_pipe.dn, g_pipe.get((size_t)(nr + c)), 0));
        // slices this range writes: forward pairs (2k, 2k+1), inverse pairs (2k-1, 2k); all of them below the slices uploaded so far
        const int z0 = inverse ? std::max(0, 2 * k0 - 1) : 2 * k0, z1 = std::min(nz, inverse ? 2 * k1 - 1 : 2 * k1);
        CK(slab(dst, dsy, dsz, v->buf[v->cur ^ 1], z0, (c == nr - 1) ? nz : z1, false, g_pipe.dn));
    }
    CK(cudaStreamSynchronize(g_pipe.dn));
    CK(cudaStreamSynchronize(g_pipe.up));
    CK(cudaStreamSynchronize(g.st));
    CK(cudaGetLastError());
    v->cur ^= 1;
    return DWTB200_OK;
}
int dwtb200_fwd3_host(const void *src, size_t ssx, size_t ssy, size_t ssz, void *dst, size_t dsx, size_t dsy, size_t dsz,
                      int nx, int ny, int nz)
{
    API_LOCK();
    NEED_DEV();
    if (!src || !dst) return fail(DWTB200_EINVAL, "fwd3_host: null argument");
    if (nx < 5 || ny < 5 || nz < 5) return fail(DWTB200_EINVAL, "fwd3_host: every size must be >= 5 (src/dwt-simple.c:2172)");
    dwtb200_volume *v = host_volume(nx, ny, nz);
    if (!v) return DWTB200_ENOMEM;
    if (host3_pipeline_applies(v, ssx, dsx)) return host3_pipelined(false, v, (const char *)src, ssy, ssz, (char *)dst, dsy, dsz);
    int r = dwtb200_volume_upload(v, src, ssx, ssy, ssz);
    if (!r) r = dwtb200_volume_fwd3(v);
    if (!r) r = dwtb200_volume_download(v, dst, dsx, dsy, dsz);
    return r;
}
int dwtb200_inv3_host(void *vol, size_t sx, size_t sy, size_t sz, int nx, int ny, int nz)
{
    API_LOCK();
    NEED_DEV();
    if (!vol) return fail(DWTB200_EINVAL, "inv3_host: null argument");
    dwtb200_volume *v = host_volume(nx, ny, nz);
    if (!v) return DWTB200_ENOMEM;
    if (host3_pipeline_applies(v, sx, sx)) return host3_pipelined(true, v, (const char *)vol, sy, sz, (char *)vol, sy, sz);
    int r = dwtb200_volume_upload(v, vol, sx, sy, sz);
    if (!r) r = dwtb200_volume_inv3(v);
    if (!r) r = dwtb200_volume_download(v, vol, sx, sy, sz);
    return r;
}

// volume_perftest_fwd97op_s (src/volume-dwt.c:2810) on the device: N runs of "fill, forward (timed with CUDA events),
// inverse, compare with the pattern"; *secs_per_voxel = minimum forward time / voxels; returns the number of runs whose
// round trip missed the reference's tolerance (dwt_util_compare2_s, 1e-3) in *errors
int dwtb200_perf3(int size, int N, double *secs_per_voxel, int *errors)
{
    API_LOCK();
    NEED_DEV();
    if (size < 5 || N < 1 || !secs_per_voxel) return fail(DWTB200_EINVAL, "perf3: bad arguments");
    dwtb200_volume *v = dwtb200_volume_create(size, size, size), *ref = dwtb200_volume_create(size, size, size);
    int r = (v && ref) ? DWTB200_OK : DWTB200_ENOMEM;
    double best = 1e30;
    int bad = 0;
    unsigned long long *d = nullptr;
    if (!r && cudaMalloc(&d, 16) != cudaSuccess) r = fail(DWTB200_ENOMEM, "perf3: cudaMalloc");
    if (!r) r = dwtb200_volume_fill(ref);
    for (int n = -1; n < N && !r; n++) {   // run -1 warms up
        r = dwtb200_volume_fill(v);
        if (!r) r = dwtb200_timer_start();
        if (!r) r = dwtb200_volume_fwd3(v);
        if (!r) {
            const double ms = dwtb200_timer_stop_ms();
            if (n >= 0 && ms >= 0 && ms * 1e-3 < best) best = ms * 1e-3;
        }
        if (!r) r = dwtb200_volume_inv3(v);
        if (!r) {
            unsigned long long h[2] = {0, 0};
            cudaMemsetAsync(d, 0, 16, g.st);
            launch_compare(4, v->buf[v->cur], ref->buf[ref->cur], v->pitch, v->slice, v->nx, v->ny, v->nz, 1, d, g.st);
            cudaMemcpyAsync(h, d, 16, cudaMemcpyDeviceToHost, g.st);
            if (cudaStreamSynchronize(g.st) != cudaSuccess) r = fail(DWTB200_ECUDA, "perf3: compare failed");
            double e;
            memcpy(&e, &h[1], 8);
            if (n >= 0 && !(e < 1e-3)) bad++;
        }
    }
    if (d) cudaFree(d);
    dwtb200_volume_destroy(v);
    dwtb200_volume_destroy(ref);
    *secs_per_voxel = best / ((double)size * size * size);
    if (errors) *errors = bad;
    return r;
}

// =====================================================================================================
// timing
// =====================================================================================================
int dwtb200_sync(void)
{
    API_LOCK();
    NEED_DEV();
    for (dwtb200_image *im : g.live) CK(cudaStreamSynchronize(im->st));
    CK(cudaStreamSynchronize(g.st0));
    return DWTB200_OK;
}
// the timer brackets ALL device work of the library: the start event waits for everything queued so far (library stream
// and every image's stream) and everything queued afterwards waits for it; the stop event waits for everything again
int dwtb200_timer_start(void)
{
    API_LOCK();
    NEED_DEV();
    for (dwtb200_image *im : g.live) wait_for_image(g.st0, im);
    CK(cudaEventRecord(g.e0, g.st0));
    for (dwtb200_image *im : g.live) CK(cudaStreamWaitEvent(im->st, g.e0, 0));
    return DWTB200_OK;
}
double dwtb200_timer_stop_ms(void)
{
    API_LOCK();
    if (g.dev < 0) return -1.0;
    for (dwtb200_image *im : g.live) wait_for_image(g.st0, im);
    if (cudaEventRecord(g.e1, g.st0) != cudaSuccess || cudaEventSynchronize(g.e1) != cudaSuccess) return -1.0;
    float ms = 0;
    if (cudaEventElapsedTime(&ms, g.e0, g.e1) != cudaSuccess) return -1.0;
    return (double)ms;
}
void *dwtb200_stream(void) { return (void *)g.st0; }

// device time of the calls made on ONE image between start and stop: CUDA events recorded on the image's own stream (the stream
// its kernels are launched on), no cross-stream joins in the timed region
int dwtb200_image_timer_start(dwtb200_image *im)
{
    API_LOCK();
    NEED_DEV();
    if (!im) return fail(DWTB200_EINVAL, "image_timer_start: null image");
    if (!im->t0) {
        CK(cudaEventCreate(&im->t0));
        CK(cudaEventCreate(&im->t1));
    }
    CK(cudaEventRecord(im->t0, im->st));
    return DWTB200_OK;
}
// A series of marks on the image's stream, recorded without synchronising: the host runs ahead of the device, so the interval
// between two marks around a call is the call's device time with no host latency in it.  dwtb200_image_timer_read synchronises and
// returns the intervals between consecutive marks (n marks -> n - 1 intervals), then forgets the marks.
int dwtb200_image_timer_mark(dwtb200_image *im)
{
    API_LOCK();
    NEED_DEV();
    if (!im) return fail(DWTB200_EINVAL, "image_timer_mark: null image");
    if (im->nmarks == im->marks.size()) {
        cudaEvent_t e;
        CK(cudaEventCreate(&e));
        im->marks.push_back(e);
    }
    CK(cudaEventRecord(im->marks[im->nmarks++], im->st));
    return DWTB200_OK;
}
int dwtb200_image_timer_read(dwtb200_image *im, double *ms, int capacity)
{
    API_LOCK();
    NEED_DEV();
    if (!im || !ms) return fail(DWTB200_EINVAL, "image_timer_read: null argument");
    int n = 0;
    if (im->nmarks) CK(cudaEventSynchronize(im->marks[im->nmarks - 1]));
    for (size_t i = 0; i + 1 < im->nmarks && n < capacity; i++) {
        float t = 0;
        CK(cudaEventElapsedTime(&t, im->marks[i], im->marks[i + 1]));
        ms[n++] = (double)t;
    }
    im->nmarks = 0;
    return n;
}
// everything queued later on `im` runs after everything queued so far on `other` (device-side ordering, no host wait)
int dwtb200_image_wait(dwtb200_image *im, dwtb200_image *other)
{
    API_LOCK();
    NEED_DEV();
    if (!im || !other) return fail(DWTB200_EINVAL, "image_wait: null image");
    wait_for_image(im->st, other);
    return DWTB200_OK;
}
double dwtb200_image_timer_stop_ms(dwtb200_image *im)
{
    API_LOCK();
    if (g.dev < 0 || !im || !im->t0) return -1.0;
    if (cudaEventRecord(im->t1, im->st) != cudaSuccess || cudaEventSynchronize(im->t1) != cudaSuccess) return -1.0;
    float ms = 0;
    if (cudaEventElapsedTime(&ms, im->t0, im->t1) != cudaSuccess) return -1.0;
    return (double)ms;
}

int dwtb200_flush_l2(size_t bytes)
{
    API_LOCK();
    NEED_DEV();
    if (bytes > g.flush_bytes) {
        if (g.flush) cudaFree(g.flush);
        g.flush = nullptr;
        g.flush_bytes = 0;
        CK(cudaMalloc(&g.flush, bytes));
        g.flush_bytes = bytes;
    }
    CK(cudaMemsetAsync(g.flush, 0x5a, bytes, g.st));
    return DWTB200_OK;
}

}  // extern "C"
