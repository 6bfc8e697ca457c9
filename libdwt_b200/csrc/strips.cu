// strips.cu -- ONE image partitioned into row strips over the GPUs of a box (BASELINE
// config 5a), behind the C ABI (dwtb200_strips_*, include/dwtb200.h).
//
// The reference has no multi-device code; the transform semantics are those of its level
// drivers (dwt_cdf97_2f_s / dwt_cdf97_2i_s, /root/reference/src/libdwt.c:12776-12924, 17040-17180: Mallat layout,
// level sizes ceil(size / 2^j)), and an image of this size cannot even be addressed by it (`y * stride_x` in int,
// src/inline.h:188).
//
// Scheme -- "exchange once, recompute the halo":
//   * one process per GPU; rank r owns image rows [R_r, R_{r+1}), R_r a multiple of A = 2^Jd (Jd = levels done
//     distributed), and HOLDS the extended strip [R_r - halo, R_{r+1} + halo) clipped to the image, halo = HALO * A
//     rows of the level-0 input (HALO = 4 for CDF 9/7, 2 for CDF 5/3: the lifting reach per level; summed over Jd
//     levels the error of the artificial strip borders travels less than HALO * 2^Jd input rows and never reaches an
//     owned row);
//   * forward: pull the halo rows from the two neighbours, run the ordinary Jd-level transform on the extended strip,
//     push the owned rows of LL_Jd into rank 0's `top` image, rank 0 runs the remaining J - Jd levels on it;
//   * inverse: rank 0 inverts the top, every rank pulls its (extended) rows of LL_Jd from rank 0 and the halo rows of
//     every distributed level's subbands from its neighbours, then inverts Jd levels locally.
//   The owned rows of every subband are bit-identical to the single-device transform (tests/test_gpu_strips.py).
//
// Transport: the ranks' planes are mapped into each other's address space with CUDA IPC (cudaIpcGetMemHandle /
// cudaIpcOpenMemHandle; handles travel through a POSIX shared-memory segment named by the caller's `session`), every
// transfer is ONE cudaMemcpy2DAsync between a local and a peer-mapped pointer (NVLink P2P through NVSwitch; never a
// batched-memcpy API).  Ordering between ranks is done ON THE DEVICE: monotonic sequence flags in a small control block
// per rank, written into the peer's block by a one-thread kernel (st.release.sys after __threadfence_system) and waited
// for by a one-warp kernel polling its own block (ld.acquire.sys).  No host synchronisation anywhere between exchange,
// level kernels, gather and rank 0's top: a call enqueues its work on the strip's streams and returns.
//
// Several ranks may also live in ONE process on one device (tests on a single GPU): the shared-memory record carries the
// raw pointers next to the IPC handles and a rank of the same process uses those.
#include <algorithm>
#include <atomic>
#include <cerrno>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "../../include/dwtb200.h"
#include "internal.h"
#include "lifting.cuh"

using namespace dwtb200;

namespace {

constexpr int MAXR = 16;            // ranks of one box
constexpr uint32_t SHM_MAGIC = 0x64773273u;

// ---- control block (device memory of every rank, mapped by every peer) -------------------------------------------
// All words hold the sequence number of a collective call (fwd / inv: 1, 2, 3, ...); they only grow.
struct Ctrl {
    uint32_t nb_ready[2];      // [0]: the rank above, [1]: the rank below -- its strip holds the data of call `seq`
    uint32_t done_by[MAXR];    // rank p has finished every read of MY memory that belongs to call `seq`
    uint32_t top_free;         // rank 0: its top image may be written for call `seq` (forward gather)
    uint32_t top_done;         // rank 0: its top image holds LL_Jd of call `seq` (inverse)
    uint32_t ll_ready[MAXR];   // (on rank 0) rank p's rows of LL_Jd have landed in the top image
    uint32_t error;            // a wait timed out (the peer died): nonzero, reported by dwtb200_strips_sync
    uint32_t pad[3];
};

struct FlagList {
    uint32_t *p[MAXR];
    int n;
};

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// everything this stream did before (kernels and copies, local or to a peer) is visible system-wide before the flags
__global__ void k_signal(FlagList f, uint32_t seq)
{
    __threadfence_system();
    if ((int)threadIdx.x < f.n) asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f.p[threadIdx.x]), "r"(seq) : "memory");
}
// lane i polls flag i of the rank's OWN control block until it reaches `seq`; gives up after `timeout_ns` (a dead peer
// must not hang the GPU) and records the failure
__global__ void k_wait(FlagList f, uint32_t seq, uint32_t *error, unsigned long long timeout_ns)
{
    if ((int)threadIdx.x < f.n) {
        unsigned long long t0;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        while ((int32_t)(ld_acquire_sys(f.p[threadIdx.x]) - seq) < 0) {
            __nanosleep(200);
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if (t - t0 > timeout_ns) {
                atomicExch(error, seq ? seq : 1u);
                break;
            }
        }
    }
    __syncwarp();
    __threadfence_system();
}

// Several rectangles copied by ONE launch (the inverse needs two per distributed level and neighbour: ten cudaMemcpy2DAsync calls
// per side kept the host busy for longer than the copies take).  The source is peer memory read over NVLink; 16-byte accesses
// where rows start and end on 16-byte boundaries.
constexpr int MAX_RECTS = 2 * 16;
struct RectList {
    int n, total_rows;
    struct R {
        char *dst;
        const char *src;
        int64_t dpitch, spitch;
        int rows, row_bytes, row0;   // row0: index of the rectangle's first row among all rows of the list
    } r[MAX_RECTS];
};
__global__ void __launch_bounds__(256) k_pull_rects(const RectList L)
{
    const int row = blockIdx.x;
    int i = 0;
    while (i + 1 < L.n && row >= L.r[i + 1].row0) i++;
    const RectList::R &q = L.r[i];
    const int y = row - q.row0;
    char *d = q.dst + (int64_t)y * q.dpitch;
    const char *s_ = q.src + (int64_t)y * q.spitch;
    const int nb = q.row_bytes;
    if ((((uintptr_t)d | (uintptr_t)s_ | (uintptr_t)nb) & 15) == 0) {
        const int4 *s4 = reinterpret_cast<const int4 *>(s_);
        int4 *d4 = reinterpret_cast<int4 *>(d);
        for (int k = blockIdx.y * blockDim.x + threadIdx.x; k < nb / 16; k += gridDim.y * blockDim.x) d4[k] = s4[k];
    } else {
        const int *s1 = reinterpret_cast<const int *>(s_);
        int *d1 = reinterpret_cast<int *>(d);
        for (int k = blockIdx.y * blockDim.x + threadIdx.x; k < nb / 4; k += gridDim.y * blockDim.x) d1[k] = s1[k];
    }
}

// bit-wise comparison of two rectangles with their own pitches: out[0] += number of differing samples
template <class U> __global__ void __launch_bounds__(256) k_cmp_rect(const U *a, int64_t pa, const U *b, int64_t pb, int nx, int ny,
                                                                     unsigned long long *out)
{
    const int x = blockIdx.x * 32 + threadIdx.x;
    unsigned cnt = 0;
    for (int y = blockIdx.y * 8 + threadIdx.y; y < ny; y += gridDim.y * 8)
        if (x < nx) cnt += a[(int64_t)y * pa + x] != b[(int64_t)y * pb + x];
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if (threadIdx.x == 0 && cnt) atomicAdd(out, (unsigned long long)cnt);
}

// ---- what a rank publishes for its peers ----------------------------------------------------------------------------
struct PeerRecord {
    std::atomic<uint32_t> ready;    // 1: the record is complete
    std::atomic<uint32_t> closed;   // 1: the rank has unmapped every peer pointer (the owners may free)
    int pid, dev;
    cudaIpcMemHandle_t plane[2], ctrl, top[2];
    uint64_t raw_plane[2], raw_ctrl, raw_top[2];
    int64_t pitch_bytes, top_pitch_bytes;
};
struct Shm {
    uint32_t magic, world;
    PeerRecord rank[MAXR];
};

inline int cdiv_pow2(int v, int j) { return (int)(((int64_t)v + ((int64_t)1 << j) - 1) >> j); }

}  // namespace

// ---- geometry (host arithmetic only; the Python model of the same plan is libdwt_b200/strips.py: StripPlan) ---------
extern "C" int dwtb200_strips_plan(int width, int height, int world, int levels_distributed, int halo_lines, int rank, dwtb200_strip_plan *o)
{
    if (!o || width < 1 || height < 1 || world < 1 || world > MAXR || levels_distributed < 0 || levels_distributed > 30 || halo_lines < 1 ||
        rank < 0 || rank >= world)
        return set_error(DWTB200_EINVAL, "strips_plan: bad arguments");
    const int Jd = levels_distributed;
    const int64_t A = (int64_t)1 << Jd;
    const int64_t units = cdiv_pow2(height, Jd);   // strips are whole multiples of A rows
    const int64_t per = (units + world - 1) / world;
    auto R = [&](int r) -> int { return r >= world ? height : (int)std::min<int64_t>(std::min<int64_t>(r * per, units) * A, height); };
    memset(o, 0, sizeof *o);
    o->halo = (int)std::min<int64_t>(halo_lines * A, 0x3fffffff);
    o->own0 = R(rank);
    o->own1 = R(rank + 1);
    o->ext0 = std::max(0, o->own0 - o->halo);
    o->ext1 = (int)std::min<int64_t>(height, (int64_t)o->own1 + o->halo);
    o->ll_w = cdiv_pow2(width, Jd);
    o->ll_h = cdiv_pow2(height, Jd);
    const bool last = o->own1 == height;
    o->ll_own0 = o->own0 >> Jd;
    o->ll_own1 = last ? cdiv_pow2(o->own1, Jd) : o->own1 >> Jd;
    o->ll_ext0 = o->ext0 >> Jd;
    o->ll_ext1 = o->ll_ext0 + cdiv_pow2(o->ext1 - o->ext0, Jd);
    o->neighbours_only = 1;   // every strip is non-empty and at least as tall as the halo: halos come from adjacent ranks only
    for (int r = 0; r < world; r++)
        if (R(r + 1) - R(r) < o->halo && world > 1) o->neighbours_only = 0;
    if (R(rank + 1) <= R(rank)) o->neighbours_only = 0;
    return DWTB200_OK;
}

// rows of distributed level j (0 <= j < Jd) in the level's OUTPUT resolution:
// out = { off, nly_g, nly_l, extL0, extL1, extH0, extH1, ownL0, ownL1, ownH0, ownH1 }
//   off    first global output row the rank holds;  nly_g / nly_l: number of L-type rows globally / locally (where the H-type rows start)
//   ext*   global ranges of L-type (LL/HL) and H-type (LH/HH) rows it holds, own*: the ones it owns
extern "C" int dwtb200_strips_band(int width, int height, int world, int levels_distributed, int halo_lines, int rank, int j, int *out)
{
    dwtb200_strip_plan p;
    const int rc = dwtb200_strips_plan(width, height, world, levels_distributed, halo_lines, rank, &p);
    if (rc) return rc;
    if (!out || j < 0 || j >= levels_distributed) return set_error(DWTB200_EINVAL, "strips_band: bad level");
    const int hloc = p.ext1 - p.ext0;
    const bool last = p.own1 == height;
    const int hg = cdiv_pow2(height, j), hl = cdiv_pow2(hloc, j);
    const int off = p.ext0 >> (j + 1);
    out[0] = off;
    out[1] = (hg + 1) >> 1;
    out[2] = (hl + 1) >> 1;
    out[3] = off;
    out[4] = off + ((hl + 1) >> 1);
    out[5] = off;
    out[6] = off + (hl >> 1);
    out[7] = p.own0 >> (j + 1);
    out[8] = last ? cdiv_pow2(p.own1, j + 1) : p.own1 >> (j + 1);
    out[9] = p.own0 >> (j + 1);
    out[10] = last ? (cdiv_pow2(p.own1, j) >> 1) : p.own1 >> (j + 1);
    return DWTB200_OK;
}

// =====================================================================================================================
struct dwtb200_strips {
    int kind = 0, W = 0, H = 0, J = 0, Jd = 0, rank = 0, world = 1, halo_lines = 4;
    size_t es = 4;
    dwtb200_strip_plan plan[MAXR];
    dwtb200_image *local = nullptr, *top = nullptr;   // top: rank 0 only
    Ctrl *ctrl = nullptr;                             // this rank's control block
    cudaStream_t up = nullptr, dn = nullptr;          // transfers from the rank above / below
    cudaEvent_t ev[6] = {};
    uint32_t seq = 0;
    int top_cur = 0;        // plane of rank 0's top image that holds / receives LL_Jd next (the same on every rank)
    int local_cur = 0;      // plane of every rank's strip that holds the data next
    // peers
    std::string shm_name;
    Shm *shm = nullptr;
    bool connected = false;
    bool direct = false;    // forward: level 0 reads the halo rows straight from the neighbours' planes (no halo copy)
    char *peer_plane[MAXR][2] = {};
    Ctrl *peer_ctrl[MAXR] = {};
    char *peer_top[2] = {nullptr, nullptr};
    int64_t peer_pitch[MAXR] = {}, top_pitch = 0;
    std::vector<void *> mapped;   // pointers obtained from cudaIpcOpenMemHandle
    unsigned long long timeout_ns = 20ull * 1000 * 1000 * 1000;
    unsigned long long nvlink_bytes = 0;   // bytes this rank moved from / to peers in the last call
};

namespace {

#define CKS(call)                                                                                                              \
    do {                                                                                                                       \
        cudaError_t e_ = (call);                                                                                               \
        if (e_ != cudaSuccess) return set_error(DWTB200_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

int signal(cudaStream_t st, std::initializer_list<uint32_t *> flags, uint32_t seq)
{
    FlagList f;
    f.n = 0;
    for (uint32_t *p : flags)
        if (p) f.p[f.n++] = p;
    if (!f.n) return 0;
    k_signal<<<1, 32, 0, st>>>(f, seq);
    CKS(cudaGetLastError());
    return 0;
}
int signal_list(cudaStream_t st, const FlagList &f, uint32_t seq)
{
    if (!f.n) return 0;
    k_signal<<<1, 32, 0, st>>>(f, seq);
    CKS(cudaGetLastError());
    return 0;
}
int wait_list(dwtb200_strips *s, cudaStream_t st, const FlagList &f, uint32_t seq)
{
    if (!f.n) return 0;
    k_wait<<<1, 32, 0, st>>>(f, seq, &s->ctrl->error, s->timeout_ns);
    CKS(cudaGetLastError());
    return 0;
}
int wait_one(dwtb200_strips *s, cudaStream_t st, uint32_t *flag, uint32_t seq)
{
    FlagList f;
    f.n = 1;
    f.p[0] = flag;
    return wait_list(s, st, f, seq);
}

int publish(dwtb200_strips *s)
{
    const bool creator = s->rank == 0;
    int fd = -1;
    if (creator) {
        shm_unlink(s->shm_name.c_str());
        fd = shm_open(s->shm_name.c_str(), O_CREAT | O_EXCL | O_RDWR, 0600);
        if (fd < 0) return set_error(DWTB200_EINVAL, "strips: shm_open(%s): %s", s->shm_name.c_str(), strerror(errno));
        if (ftruncate(fd, sizeof(Shm)) != 0) {
            close(fd);
            return set_error(DWTB200_ENOMEM, "strips: ftruncate: %s", strerror(errno));
        }
    } else {
        const auto t0 = std::chrono::steady_clock::now();
        for (;;) {
            fd = shm_open(s->shm_name.c_str(), O_RDWR, 0600);
            if (fd >= 0) {
                struct stat sb;
                if (fstat(fd, &sb) == 0 && (size_t)sb.st_size >= sizeof(Shm)) break;
                close(fd);
                fd = -1;
            }
            if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(120))
                return set_error(DWTB200_EINVAL, "strips: rank 0 never created the session %s", s->shm_name.c_str());
            std::this_thread::sleep_for(std::chrono::milliseconds(2));
        }
    }
    void *m = mmap(nullptr, sizeof(Shm), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (m == MAP_FAILED) return set_error(DWTB200_ENOMEM, "strips: mmap: %s", strerror(errno));
    s->shm = (Shm *)m;
    if (creator) {
        s->shm->world = (uint32_t)s->world;
        std::atomic_thread_fence(std::memory_order_release);
        ((volatile uint32_t *)&s->shm->magic)[0] = SHM_MAGIC;
    } else {
        const auto t0 = std::chrono::steady_clock::now();
        while (((volatile uint32_t *)&s->shm->magic)[0] != SHM_MAGIC) {
            if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(120)) return set_error(DWTB200_EINVAL, "strips: session never initialised");
            std::this_thread::sleep_for(std::chrono::milliseconds(1));
        }
        if (s->shm->world != (uint32_t)s->world) return set_error(DWTB200_EINVAL, "strips: session was created for %u ranks", s->shm->world);
    }
    PeerRecord &me = s->shm->rank[s->rank];
    const ImageView lv = image_view(s->local);
    me.pid = (int)getpid();
    me.dev = dwtb200_device();
    for (int i = 0; i < 2; i++) {
        CKS(cudaIpcGetMemHandle(&me.plane[i], lv.plane[i]));
        me.raw_plane[i] = (uint64_t)(uintptr_t)lv.plane[i];
    }
    CKS(cudaIpcGetMemHandle(&me.ctrl, s->ctrl));
    me.raw_ctrl = (uint64_t)(uintptr_t)s->ctrl;
    me.pitch_bytes = lv.pitch * (int64_t)lv.es;
    if (s->top) {
        const ImageView tv = image_view(s->top);
        for (int i = 0; i < 2; i++) {
            CKS(cudaIpcGetMemHandle(&me.top[i], tv.plane[i]));
            me.raw_top[i] = (uint64_t)(uintptr_t)tv.plane[i];
        }
        me.top_pitch_bytes = tv.pitch * (int64_t)tv.es;
    }
    me.closed.store(0);
    me.ready.store(1, std::memory_order_release);
    return 0;
}

int open_handle(dwtb200_strips *s, const cudaIpcMemHandle_t &h, void **out)
{
    CKS(cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess));
    s->mapped.push_back(*out);
    return 0;
}

int connect(dwtb200_strips *s, bool may_warm = false)
{
    if (s->connected) return 0;
    const int mypid = (int)getpid();
    const auto t0 = std::chrono::steady_clock::now();
    for (int p = 0; p < s->world; p++) {
        PeerRecord &r = s->shm->rank[p];
        while (!r.ready.load(std::memory_order_acquire)) {
            if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(120)) return set_error(DWTB200_EINVAL, "strips: rank %d never joined", p);
            std::this_thread::sleep_for(std::chrono::milliseconds(1));
        }
        const bool needed = p == s->rank || p == s->rank - 1 || p == s->rank + 1 || p == 0 || s->rank == 0;
        if (!needed) continue;
        if (r.pid == mypid) {   // a rank of this very process (single-GPU emulation, or this rank itself)
            for (int i = 0; i < 2; i++) s->peer_plane[p][i] = (char *)(uintptr_t)r.raw_plane[i];
            s->peer_ctrl[p] = (Ctrl *)(uintptr_t)r.raw_ctrl;
            if (p == 0)
                for (int i = 0; i < 2; i++) s->peer_top[i] = (char *)(uintptr_t)r.raw_top[i];
        } else {
            for (int i = 0; i < 2; i++) {
                void *q = nullptr;
                const int rc = open_handle(s, r.plane[i], &q);
                if (rc) return rc;
                s->peer_plane[p][i] = (char *)q;
            }
            void *q = nullptr;
            int rc = open_handle(s, r.ctrl, &q);
            if (rc) return rc;
            s->peer_ctrl[p] = (Ctrl *)q;
            if (p == 0)
                for (int i = 0; i < 2; i++) {
                    rc = open_handle(s, r.top[i], &q);
                    if (rc) return rc;
                    s->peer_top[i] = (char *)q;
                }
        }
        s->peer_pitch[p] = r.pitch_bytes;
        if (p == 0) s->top_pitch = r.top_pitch_bytes;
    }
    s->connected = true;
    // Forward level 0 can fetch the halo rows itself: its ring producer's bulk copies read them from the neighbours' planes over
    // NVLink while the rest of the strip streams from local HBM, instead of a copy phase (two 33 MB pulls per rank at 65536^2 x 8,
    // ~90 us of a 1.2 ms call) in front of the transform.  Only when level 0 is a ring level, and only from an explicit
    // dwtb200_strips_connect(): the graphs must be captured again with the neighbours' pointers in the kernel parameters, and
    // capturing transforms the planes (the caller's data would be lost inside an implicit connect).
    const char *e = getenv("DWTB200_STRIPS_DIRECT");
    const int r = s->rank, G = s->world;
    const ImageView lv = image_view(s->local);
    const int64_t mypitch = lv.pitch * (int64_t)lv.es;
    if (may_warm && G > 1 && !(e && !atoi(e)) && image_level0_is_ring(s->local, s->Jd) && (r == 0 || s->peer_pitch[r - 1] == mypitch) &&
        (r == G - 1 || s->peer_pitch[r + 1] == mypitch)) {
        const dwtb200_strip_plan &pl = s->plan[r];
        const void *up[2] = {nullptr, nullptr}, *dn[2] = {nullptr, nullptr};
        for (int i = 0; i < 2; i++) {
            if (r > 0) up[i] = s->peer_plane[r - 1][i];
            if (r < G - 1) dn[i] = s->peer_plane[r + 1][i];
        }
        image_set_row_sources(s->local, up, dn, r > 0 ? pl.own0 - pl.ext0 : 0, r < G - 1 ? pl.own1 - pl.ext0 : 0,
                              r > 0 ? pl.ext0 - s->plan[r - 1].ext0 : 0, r < G - 1 ? pl.own1 - s->plan[r + 1].ext0 : 0);
        bool ok = true;
        int j = s->Jd;
        for (int i = 0; ok && i < 2; i++) ok = dwtb200_image_fwd2(s->local, s->W, pl.ext1 - pl.ext0, &j, 0, 0) == 0;
        for (int i = 0; ok && i < 2; i++) ok = dwtb200_image_inv2(s->local, s->W, pl.ext1 - pl.ext0, s->Jd, 0, 0) == 0;
        ok = ok && cudaDeviceSynchronize() == cudaSuccess;
        if (!ok) return DWTB200_ECUDA;
        s->direct = true;
    }
    return 0;
}

// rows [r0, r1) x columns [c0, c1) (elements) of plane `dst` <- the same-sized rectangle of `src` starting at row sr0
int copy_rect(dwtb200_strips *s, cudaStream_t st, char *dst, int64_t dpitch, int dr0, const char *src, int64_t spitch, int sr0, int rows, int c0,
              int c1, bool remote)
{
    if (rows <= 0 || c1 <= c0) return 0;
    CKS(cudaMemcpy2DAsync(dst + (size_t)dr0 * dpitch + (size_t)c0 * s->es, (size_t)dpitch, src + (size_t)sr0 * spitch + (size_t)c0 * s->es,
                          (size_t)spitch, (size_t)(c1 - c0) * s->es, (size_t)rows, cudaMemcpyDeviceToDevice, st));
    if (remote) s->nvlink_bytes += (unsigned long long)(c1 - c0) * s->es * (unsigned long long)rows;
    return 0;
}

// ---- the start of every collective call ----------------------------------------------------------------------------
// (1) nothing of mine may be overwritten while a peer still reads it for the previous call: wait for their done_by flags;
// (2) then tell the neighbours that my strip holds the data of this call (stream order: after everything queued on it).
int begin_call(dwtb200_strips *s, cudaStream_t L, cudaStream_t T)
{
    const uint32_t seq = s->seq;
    const int r = s->rank;
    FlagList f;
    f.n = 0;
    if (seq > 1) {
        if (r > 0) f.p[f.n++] = &s->ctrl->done_by[r - 1];
        if (r < s->world - 1) f.p[f.n++] = &s->ctrl->done_by[r + 1];
        int rc = wait_list(s, L, f, seq - 1);
        if (rc) return rc;
        if (r == 0 && T) {   // the top image is read (inverse) and written (forward) by every rank
            f.n = 0;
            for (int p = 1; p < s->world; p++) f.p[f.n++] = &s->ctrl->done_by[p];
            rc = wait_list(s, T, f, seq - 1);
            if (rc) return rc;
        }
    }
    return signal(L, {r > 0 ? &s->peer_ctrl[r - 1]->nb_ready[1] : nullptr, r < s->world - 1 ? &s->peer_ctrl[r + 1]->nb_ready[0] : nullptr}, seq);
}

void band_of(const dwtb200_strips *s, int r, int j, int *b)
{
    dwtb200_strips_band(s->W, s->H, s->world, s->Jd, s->halo_lines, r, j, b);
}

}  // namespace

extern "C" {

dwtb200_strips *dwtb200_strips_create(int kind, int width, int height, int levels_distributed, int rank, int world, const char *session)
{
    std::lock_guard<std::recursive_mutex> lock(api_mutex());
    if (dwtb200_device() < 0 && dwtb200_init(-1)) return nullptr;
    if (kind < 0 || kind >= DWTB200_KIND_COUNT || width < 2 || height < 2 || world < 1 || world > MAXR || rank < 0 || rank >= world || !session ||
        !*session) {
        set_error(DWTB200_EINVAL, "strips_create: bad arguments");
        return nullptr;
    }
    {   // load the protocol kernels now: a lazy module load behind a spinning wait kernel could block
        cudaFuncAttributes a;
        if (cudaFuncGetAttributes(&a, k_signal) != cudaSuccess || cudaFuncGetAttributes(&a, k_wait) != cudaSuccess ||
            cudaFuncGetAttributes(&a, k_pull_rects) != cudaSuccess ||
            cudaFuncGetAttributes(&a, k_cmp_rect<uint32_t>) != cudaSuccess || cudaFuncGetAttributes(&a, k_cmp_rect<uint64_t>) != cudaSuccess) {
            set_error(DWTB200_ECUDA, "strips_create: %s", cudaGetErrorString(cudaGetLastError()));
            return nullptr;
        }
    }
    dwtb200_strips *s = new dwtb200_strips;
    s->kind = kind;
    s->W = width;
    s->H = height;
    s->rank = rank;
    s->world = world;
    s->es = (size_t)kind_elem_size(kind);
    s->halo_lines = kind_lifting_steps(kind) == 4 ? 4 : 2;
    s->J = dwtb200_clamp_j(-1, width, height, 0);
    int Jd = levels_distributed;
    if (Jd <= 0) {   // heuristic: hand over to rank 0 once LL is at most 2048^2 samples (its transform then costs about as much as the gather)
        Jd = 0;
        while (Jd < s->J && (int64_t)cdiv_pow2(width, Jd) * cdiv_pow2(height, Jd) > (int64_t)2048 * 2048) Jd++;
        if (Jd < 1) Jd = 1;
    }
    if (Jd > s->J) Jd = s->J;
    s->Jd = Jd;
    for (int p = 0; p < world; p++) dwtb200_strips_plan(width, height, world, Jd, s->halo_lines, p, &s->plan[p]);
    const dwtb200_strip_plan &pl = s->plan[rank];
    if (!pl.neighbours_only) {
        set_error(DWTB200_EINVAL, "strips_create: a strip of %d x %d over %d ranks with %d distributed levels is shorter than its halo (%d rows)", width,
                  height, world, Jd, pl.halo);
        delete s;
        return nullptr;
    }
    if (const char *e = getenv("DWTB200_STRIPS_TIMEOUT_S")) s->timeout_ns = (unsigned long long)atof(e) * 1000000000ull;
    bool ok = (s->local = dwtb200_image_create(kind, width, pl.ext1 - pl.ext0, 1)) != nullptr;
    if (ok && rank == 0) ok = (s->top = dwtb200_image_create(kind, pl.ll_w, pl.ll_h, 1)) != nullptr;
    ok = ok && cudaMalloc((void **)&s->ctrl, sizeof(Ctrl)) == cudaSuccess && cudaMemset(s->ctrl, 0, sizeof(Ctrl)) == cudaSuccess;
    ok = ok && cudaStreamCreateWithFlags(&s->up, cudaStreamNonBlocking) == cudaSuccess && cudaStreamCreateWithFlags(&s->dn, cudaStreamNonBlocking) == cudaSuccess;
    for (int i = 0; ok && i < 6; i++) ok = cudaEventCreateWithFlags(&s->ev[i], cudaEventDisableTiming) == cudaSuccess;
    if (ok) {
        // capture and instantiate the graphs of both directions for both planes NOW: a capture inside a collective call would synchronise the
        // stream behind a wait kernel (and, with several ranks in one process, never return)
        int j = Jd;
        for (int i = 0; ok && i < 2; i++) ok = dwtb200_image_fwd2(s->local, width, pl.ext1 - pl.ext0, &j, 0, 0) == 0;
        for (int i = 0; ok && i < 2; i++) ok = dwtb200_image_inv2(s->local, width, pl.ext1 - pl.ext0, Jd, 0, 0) == 0;
        if (ok && s->top && s->J > Jd) {
            int jt = s->J - Jd;
            for (int i = 0; ok && i < 2; i++) ok = dwtb200_image_fwd2(s->top, pl.ll_w, pl.ll_h, &jt, 0, 0) == 0;
            for (int i = 0; ok && i < 2; i++) ok = dwtb200_image_inv2(s->top, pl.ll_w, pl.ll_h, s->J - Jd, 0, 0) == 0;
        }
        ok = ok && cudaDeviceSynchronize() == cudaSuccess;
        if (ok) {
            s->local_cur = image_view(s->local).cur;
            s->top_cur = 0;
            if (s->top) s->top_cur = image_view(s->top).cur;
        }
    }
    if (!ok) {
        if (cudaPeekAtLastError() != cudaSuccess) set_error(DWTB200_ENOMEM, "strips_create: %s", cudaGetErrorString(cudaGetLastError()));
        dwtb200_strips_destroy(s);
        return nullptr;
    }
    s->shm_name = std::string("/") + session;
    if (publish(s)) {
        dwtb200_strips_destroy(s);
        return nullptr;
    }
    return s;
}

int dwtb200_strips_connect(dwtb200_strips *s)
{
    std::lock_guard<std::recursive_mutex> lock(api_mutex());
    if (!s) return set_error(DWTB200_EINVAL, "strips_connect: null");
    return connect(s, true);
}

void dwtb200_strips_destroy(dwtb200_strips *s)
{
    std::lock_guard<std::recursive_mutex> lock(api_mutex());
    if (!s) return;
    cudaDeviceSynchronize();
    for (void *p : s->mapped) cudaIpcCloseMemHandle(p);
    s->mapped.clear();
    if (s->shm) {
        // the owner of exported memory must not free it while another PROCESS still maps it
        PeerRecord &me = s->shm->rank[s->rank];
        me.closed.store(1, std::memory_order_release);
        const int mypid = (int)getpid();
        const auto t0 = std::chrono::steady_clock::now();
        for (int p = 0; p < s->world; p++) {
            PeerRecord &r = s->shm->rank[p];
            if (!r.ready.load() || r.pid == mypid) continue;
            while (!r.closed.load(std::memory_order_acquire) && std::chrono::steady_clock::now() - t0 < std::chrono::seconds(10))
                std::this_thread::sleep_for(std::chrono::milliseconds(1));
        }
        munmap(s->shm, sizeof(Shm));
        if (s->rank == 0) shm_unlink(s->shm_name.c_str());
    }
    for (cudaEvent_t e : s->ev)
        if (e) cudaEventDestroy(e);
    if (s->up) cudaStreamDestroy(s->up);
    if (s->dn) cudaStreamDestroy(s->dn);
    if (s->ctrl) cudaFree(s->ctrl);
    if (s->local) dwtb200_image_destroy(s->local);
    if (s->top) dwtb200_image_destroy(s->top);
    delete s;
}

dwtb200_image *dwtb200_strips_image(dwtb200_strips *s) { return s ? s->local : nullptr; }
dwtb200_image *dwtb200_strips_top(dwtb200_strips *s) { return s ? s->top : nullptr; }
int dwtb200_strips_levels(dwtb200_strips *s, int *j_total, int *j_distributed)
{
    if (!s) return set_error(DWTB200_EINVAL, "strips_levels: null");
    if (j_total) *j_total = s->J;
    if (j_distributed) *j_distributed = s->Jd;
    return DWTB200_OK;
}
int dwtb200_strips_get_plan(dwtb200_strips *s, dwtb200_strip_plan *out)
{
    if (!s || !out) return set_error(DWTB200_EINVAL, "strips_get_plan: null");
    *out = s->plan[s->rank];
    return DWTB200_OK;
}
unsigned long long dwtb200_strips_last_peer_bytes(dwtb200_strips *s) { return s ? s->nvlink_bytes : 0; }

// ---- forward --------------------------------------------------------------------------------------------------------
int dwtb200_strips_fwd2(dwtb200_strips *s, int *j_max_ptr)
{
    std::lock_guard<std::recursive_mutex> lock(api_mutex());
    if (!s) return set_error(DWTB200_EINVAL, "strips_fwd2: null");
    int rc = connect(s);
    if (rc) return rc;
    ImageView lv = image_view(s->local);
    if (lv.cur != s->local_cur) return set_error(DWTB200_EINVAL, "strips_fwd2: the strip image was transformed outside the strips calls");
    const int r = s->rank, G = s->world, cur = s->local_cur;
    const dwtb200_strip_plan &pl = s->plan[r];
    const int64_t pitch = lv.pitch * (int64_t)lv.es;
    cudaStream_t L = lv.st, T = s->top ? image_view(s->top).st : nullptr;
    s->seq++;
    const uint32_t seq = s->seq;
    s->nvlink_bytes = 0;
    rc = begin_call(s, L, T);
    if (rc) return rc;
    if (r == 0 && G > 1) {   // the top image may be written for this call once everything queued on its stream has run
        FlagList f;
        f.n = 0;
        for (int p = 1; p < G; p++) f.p[f.n++] = &s->peer_ctrl[p]->top_free;
        rc = signal_list(T, f, seq);
        if (rc) return rc;
    }
    if (s->direct) {
        // level 0 reads the halo rows from the neighbours' planes itself: just wait until their strips hold this call's data
        FlagList f;
        f.n = 0;
        if (r > 0) f.p[f.n++] = &s->ctrl->nb_ready[0];
        if (r < G - 1) f.p[f.n++] = &s->ctrl->nb_ready[1];
        rc = wait_list(s, L, f, seq);
        if (rc) return rc;
        s->nvlink_bytes += (unsigned long long)((pl.own0 - pl.ext0) + (pl.ext1 - pl.own1)) * (unsigned long long)s->W * s->es;
    }
    // halo rows of the level-0 input, pulled from the owned rows of the two neighbours on a stream each
    CKS(cudaEventRecord(s->ev[0], L));
    if (!s->direct && r > 0) {
        const dwtb200_strip_plan &nb = s->plan[r - 1];
        CKS(cudaStreamWaitEvent(s->up, s->ev[0], 0));
        rc = wait_one(s, s->up, &s->ctrl->nb_ready[0], seq);
        if (!rc) rc = copy_rect(s, s->up, (char *)lv.plane[cur], pitch, 0, s->peer_plane[r - 1][cur], s->peer_pitch[r - 1], pl.ext0 - nb.ext0,
                                pl.own0 - pl.ext0, 0, s->W, true);
        if (!rc) rc = signal(s->up, {&s->peer_ctrl[r - 1]->done_by[r]}, seq);
        if (rc) return rc;
        CKS(cudaEventRecord(s->ev[1], s->up));
        CKS(cudaStreamWaitEvent(L, s->ev[1], 0));
    }
    if (!s->direct && r < G - 1) {
        const dwtb200_strip_plan &nb = s->plan[r + 1];
        CKS(cudaStreamWaitEvent(s->dn, s->ev[0], 0));
        rc = wait_one(s, s->dn, &s->ctrl->nb_ready[1], seq);
        if (!rc) rc = copy_rect(s, s->dn, (char *)lv.plane[cur], pitch, pl.own1 - pl.ext0, s->peer_plane[r + 1][cur], s->peer_pitch[r + 1],
                                pl.own1 - nb.ext0, pl.ext1 - pl.own1, 0, s->W, true);
        if (!rc) rc = signal(s->dn, {&s->peer_ctrl[r + 1]->done_by[r]}, seq);
        if (rc) return rc;
        CKS(cudaEventRecord(s->ev[2], s->dn));
        CKS(cudaStreamWaitEvent(L, s->ev[2], 0));
    }
    // Jd levels on the extended strip (the cached CUDA graph of the ordinary transform)
    int jd = s->Jd;
    rc = dwtb200_image_fwd2(s->local, s->W, pl.ext1 - pl.ext0, &jd, 0, 0);
    if (rc) return rc;
    if (s->direct) {   // the neighbours' rows have been read
        rc = signal(L, {r > 0 ? &s->peer_ctrl[r - 1]->done_by[r] : nullptr, r < G - 1 ? &s->peer_ctrl[r + 1]->done_by[r] : nullptr}, seq);
        if (rc) return rc;
    }
    s->local_cur ^= 1;
    lv = image_view(s->local);
    // owned rows of LL_Jd -> rank 0's top image
    const char *mine = (const char *)lv.plane[lv.cur];
    const int lrow = pl.ll_own0 - pl.ll_ext0, nrows = pl.ll_own1 - pl.ll_own0;
    if (r == 0) {
        const ImageView tv = image_view(s->top);
        CKS(cudaEventRecord(s->ev[3], T));   // everything queued on the top image so far
        CKS(cudaStreamWaitEvent(L, s->ev[3], 0));
        rc = copy_rect(s, L, (char *)tv.plane[tv.cur], tv.pitch * (int64_t)tv.es, pl.ll_own0, mine, pitch, lrow, nrows, 0, pl.ll_w, false);
        if (rc) return rc;
        CKS(cudaEventRecord(s->ev[4], L));
        CKS(cudaStreamWaitEvent(T, s->ev[4], 0));
        FlagList f;
        f.n = 0;
        for (int p = 1; p < G; p++)
            if (s->plan[p].ll_own1 > s->plan[p].ll_own0) f.p[f.n++] = &s->ctrl->ll_ready[p];
        rc = wait_list(s, T, f, seq);
        if (rc) return rc;
        if (s->J > s->Jd) {
            int jt = s->J - s->Jd;
            rc = dwtb200_image_fwd2(s->top, pl.ll_w, pl.ll_h, &jt, 0, 0);
            if (rc) return rc;
            s->top_cur ^= 1;
        }
    } else {
        rc = wait_one(s, L, &s->ctrl->top_free, seq);
        if (!rc) rc = copy_rect(s, L, s->peer_top[s->top_cur], s->top_pitch, pl.ll_own0, mine, pitch, lrow, nrows, 0, pl.ll_w, true);
        if (!rc) rc = signal(L, {&s->peer_ctrl[0]->ll_ready[r], &s->peer_ctrl[0]->done_by[r]}, seq);
        if (rc) return rc;
        if (s->J > s->Jd) s->top_cur ^= 1;
    }
    if (j_max_ptr) *j_max_ptr = s->J;
    return DWTB200_OK;
}

// ---- inverse --------------------------------------------------------------------------------------------------------
int dwtb200_strips_inv2(dwtb200_strips *s, int j_max)
{
    std::lock_guard<std::recursive_mutex> lock(api_mutex());
    if (!s) return set_error(DWTB200_EINVAL, "strips_inv2: null");
    if (j_max != s->J && j_max >= 0) return set_error(DWTB200_EINVAL, "strips_inv2: only the full depth J = %d is distributed", s->J);
    int rc = connect(s);
    if (rc) return rc;
    ImageView lv = image_view(s->local);
    if (lv.cur != s->local_cur) return set_error(DWTB200_EINVAL, "strips_inv2: the strip image was transformed outside the strips calls");
    const int r = s->rank, G = s->world, cur = s->local_cur;
    const dwtb200_strip_plan &pl = s->plan[r];
    const int64_t pitch = lv.pitch * (int64_t)lv.es;
    char *plane = (char *)lv.plane[cur];
    cudaStream_t L = lv.st, T = s->top ? image_view(s->top).st : nullptr;
    s->seq++;
    const uint32_t seq = s->seq;
    s->nvlink_bytes = 0;
    rc = begin_call(s, L, T);
    if (rc) return rc;
    // rank 0: the top of the pyramid back to LL_Jd, then let everybody fetch its rows
    if (r == 0) {
        if (s->J > s->Jd) {
            rc = dwtb200_image_inv2(s->top, pl.ll_w, pl.ll_h, s->J - s->Jd, 0, 0);
            if (rc) return rc;
            s->top_cur ^= 1;
        }
        FlagList f;
        f.n = 0;
        for (int p = 1; p < G; p++) f.p[f.n++] = &s->peer_ctrl[p]->top_done;
        rc = signal_list(T, f, seq);
        if (rc) return rc;
        CKS(cudaEventRecord(s->ev[3], T));
    } else if (s->J > s->Jd) {
        s->top_cur ^= 1;
    }
    // halo rows of the subbands of every distributed level, from the neighbours (their Mallat strips are complete: nb_ready)
    CKS(cudaEventRecord(s->ev[0], L));
    for (int side = 0; side < 2; side++) {
        const int nb = side == 0 ? r - 1 : r + 1;
        if (nb < 0 || nb >= G) continue;
        cudaStream_t st = side == 0 ? s->up : s->dn;
        CKS(cudaStreamWaitEvent(st, s->ev[0], 0));
        rc = wait_one(s, st, &s->ctrl->nb_ready[side], seq);
        if (rc) return rc;
        RectList list;
        list.n = list.total_rows = 0;
        int widest = 0;
        for (int j = 0; j < s->Jd; j++) {
            int me[11], ot[11];
            band_of(s, r, j, me);
            band_of(s, nb, j, ot);
            const int w = cdiv_pow2(s->W, j), nlx = (w + 1) >> 1;
            // L-type rows (HL columns [nlx, w)) and H-type rows (LH | HH, columns [0, w)) the neighbour owns and I hold
            for (int type = 0; type < 2; type++) {
                const int e0 = me[3 + 2 * type], e1 = me[4 + 2 * type], o0 = ot[7 + 2 * type], o1 = ot[8 + 2 * type];
                const int g0 = std::max(o0, e0), g1 = std::min(o1, e1);
                const int c0 = type ? 0 : nlx;
                if (g1 <= g0 || w <= c0) continue;
                const int base_me = type ? me[2] : 0, base_ot = type ? ot[2] : 0;
                if (list.n == MAX_RECTS) {   // (more than 16 distributed levels: cannot happen below 2^20 rows per rank, kept for safety)
                    k_pull_rects<<<dim3(list.total_rows, std::max(1, std::min(16, widest / 16384))), 256, 0, st>>>(list);
                    list.n = list.total_rows = 0;
                }
                RectList::R &q = list.r[list.n++];
                q.dst = plane + (size_t)(base_me + g0 - me[0]) * pitch + (size_t)c0 * s->es;
                q.src = s->peer_plane[nb][cur] + (size_t)(base_ot + g0 - ot[0]) * s->peer_pitch[nb] + (size_t)c0 * s->es;
                q.dpitch = pitch;
                q.spitch = s->peer_pitch[nb];
                q.rows = g1 - g0;
                q.row_bytes = (int)((size_t)(w - c0) * s->es);
                q.row0 = list.total_rows;
                list.total_rows += q.rows;
                widest = std::max(widest, q.row_bytes);
                s->nvlink_bytes += (unsigned long long)q.rows * (unsigned long long)q.row_bytes;
            }
        }
        if (list.total_rows > 0) {
            const dim3 grid(list.total_rows, std::max(1, std::min(16, widest / 16384)));
            k_pull_rects<<<grid, 256, 0, st>>>(list);
            CKS(cudaGetLastError());
        }
        rc = signal(st, {&s->peer_ctrl[nb]->done_by[r]}, seq);
        if (rc) return rc;
        CKS(cudaEventRecord(s->ev[1 + side], st));
    }
    // my (extended) rows of LL_Jd from rank 0's top image
    const int nrows = pl.ll_ext1 - pl.ll_ext0;
    if (r == 0) {
        const ImageView tv = image_view(s->top);
        CKS(cudaStreamWaitEvent(L, s->ev[3], 0));
        rc = copy_rect(s, L, plane, pitch, 0, (const char *)tv.plane[tv.cur], tv.pitch * (int64_t)tv.es, pl.ll_ext0, nrows, 0, pl.ll_w, false);
        if (rc) return rc;
    } else {
        rc = wait_one(s, L, &s->ctrl->top_done, seq);
        if (!rc) rc = copy_rect(s, L, plane, pitch, 0, s->peer_top[s->top_cur], s->top_pitch, pl.ll_ext0, nrows, 0, pl.ll_w, true);
        if (!rc) rc = signal(L, {&s->peer_ctrl[0]->done_by[r]}, seq);
        if (rc) return rc;
    }
    if (r > 0) CKS(cudaStreamWaitEvent(L, s->ev[1], 0));
    if (r < G - 1) CKS(cudaStreamWaitEvent(L, s->ev[2], 0));
    rc = dwtb200_image_inv2(s->local, s->W, pl.ext1 - pl.ext0, s->Jd, 0, 0);
    if (rc) return rc;
    s->local_cur ^= 1;
    return DWTB200_OK;
}

// waits for this rank's part of the queued calls; reports a timed-out peer wait
int dwtb200_strips_sync(dwtb200_strips *s)
{
    std::lock_guard<std::recursive_mutex> lock(api_mutex());
    if (!s) return set_error(DWTB200_EINVAL, "strips_sync: null");
    CKS(cudaStreamSynchronize(image_view(s->local).st));
    CKS(cudaStreamSynchronize(s->up));
    CKS(cudaStreamSynchronize(s->dn));
    if (s->top) CKS(cudaStreamSynchronize(image_view(s->top).st));
    uint32_t err = 0;
    CKS(cudaMemcpy(&err, &s->ctrl->error, sizeof err, cudaMemcpyDeviceToHost));
    if (err) return set_error(DWTB200_ECUDA, "strips: rank %d gave up waiting for a peer in call %u (timeout)", s->rank, err);
    return DWTB200_OK;
}

// ---- verification helpers ------------------------------------------------------------------------------------------------
// Number of samples in which the rows this rank OWNS differ (bit-wise) from a single-device image of the whole picture living on
// this device: mallat != 0: `full` holds the forward transform (Mallat layout) -- every distributed level's owned HL and LH | HH
// rows are compared, and on rank 0 the whole top image against the top-left LL_Jd block; mallat == 0: `full` holds samples (the input,
// or the inverse transform) and the owned rows of the strip are compared.
int64_t dwtb200_strips_compare_owned(dwtb200_strips *s, dwtb200_image *full, int mallat)
{
    std::lock_guard<std::recursive_mutex> lock(api_mutex());
    if (!s || !full) {
        set_error(DWTB200_EINVAL, "strips_compare_owned: null");
        return -1;
    }
    const ImageView lv = image_view(s->local), fv = image_view(full);
    if (fv.ox != s->W || fv.oy != s->H || fv.kind != s->kind) {
        set_error(DWTB200_EINVAL, "strips_compare_owned: the single-device image must be %d x %d of the same kind", s->W, s->H);
        return -1;
    }
    if (dwtb200_strips_sync(s)) return -1;
    cudaStreamSynchronize(fv.st);
    unsigned long long *d = nullptr, h = 0;
    if (cudaMalloc(&d, 8) != cudaSuccess || cudaMemset(d, 0, 8) != cudaSuccess) return -1;
    const dwtb200_strip_plan &pl = s->plan[s->rank];
    const size_t es = s->es;
    auto cmp = [&](const char *a, int64_t pa, const char *b, int64_t pb, int nx, int ny) {
        if (nx <= 0 || ny <= 0) return;
        const dim3 blk(32, 8), grid((nx + 31) / 32, std::min((ny + 7) / 8, 32768));
        if (es == 8) k_cmp_rect<uint64_t><<<grid, blk>>>((const uint64_t *)a, pa, (const uint64_t *)b, pb, nx, ny, d);
        else k_cmp_rect<uint32_t><<<grid, blk>>>((const uint32_t *)a, pa, (const uint32_t *)b, pb, nx, ny, d);
    };
    const char *lp = (const char *)lv.plane[lv.cur], *fp = (const char *)fv.plane[fv.cur];
    const int64_t lpitch = lv.pitch, fpitch = fv.pitch;
    auto at = [&](const char *base, int64_t pitch, int row, int col) { return base + ((size_t)row * pitch + col) * es; };
    if (!mallat) {
        cmp(at(lp, lpitch, pl.own0 - pl.ext0, 0), lpitch, at(fp, fpitch, pl.own0, 0), fpitch, s->W, pl.own1 - pl.own0);
    } else {
        for (int j = 0; j < s->Jd; j++) {
            int b[11];
            band_of(s, s->rank, j, b);
            const int w = cdiv_pow2(s->W, j), nlx = (w + 1) >> 1;
            cmp(at(lp, lpitch, b[7] - b[0], nlx), lpitch, at(fp, fpitch, b[7], nlx), fpitch, w - nlx, b[8] - b[7]);                     // HL
            cmp(at(lp, lpitch, b[2] + b[9] - b[0], 0), lpitch, at(fp, fpitch, b[1] + b[9], 0), fpitch, w, b[10] - b[9]);                // LH | HH
        }
        if (s->top) {
            const ImageView tv = image_view(s->top);
            cudaStreamSynchronize(tv.st);
            cmp((const char *)tv.plane[tv.cur], tv.pitch, fp, fpitch, pl.ll_w, pl.ll_h);
        }
    }
    const cudaError_t e = cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) {
        set_error(DWTB200_ECUDA, "strips_compare_owned: %s", cudaGetErrorString(e));
        return -1;
    }
    return (int64_t)h;
}

// The owned rows of this rank's result written into a host image of the WHOLE picture (byte stride_x between rows): mallat != 0: the
// owned subband rows of every distributed level at their Mallat positions and, on rank 0, the top of the pyramid in the top-left corner;
// mallat == 0: the owned image rows.  With every rank writing into the same array (or gathering), the array is the complete result.
int dwtb200_strips_download_owned(dwtb200_strips *s, void *host_full, int64_t stride_x, int mallat)
{
    std::lock_guard<std::recursive_mutex> lock(api_mutex());
    if (!s || !host_full || stride_x < (int64_t)(s->W * s->es)) return set_error(DWTB200_EINVAL, "strips_download_owned: bad arguments");
    int rc = dwtb200_strips_sync(s);
    if (rc) return rc;
    const ImageView lv = image_view(s->local);
    const dwtb200_strip_plan &pl = s->plan[s->rank];
    const size_t es = s->es, lpitch = (size_t)lv.pitch * es;
    const char *lp = (const char *)lv.plane[lv.cur];
    char *hp = (char *)host_full;
    auto get = [&](int hrow, int lrow, int col, int ncol, int nrow) -> cudaError_t {
        if (ncol <= 0 || nrow <= 0) return cudaSuccess;
        return cudaMemcpy2D(hp + (size_t)hrow * stride_x + (size_t)col * es, (size_t)stride_x, lp + (size_t)lrow * lpitch + (size_t)col * es, lpitch,
                            (size_t)ncol * es, nrow, cudaMemcpyDeviceToHost);
    };
    if (!mallat) {
        CKS(get(pl.own0, pl.own0 - pl.ext0, 0, s->W, pl.own1 - pl.own0));
        return DWTB200_OK;
    }
    for (int j = 0; j < s->Jd; j++) {
        int b[11];
        band_of(s, s->rank, j, b);
        const int w = cdiv_pow2(s->W, j), nlx = (w + 1) >> 1;
        CKS(get(b[7], b[7] - b[0], nlx, w - nlx, b[8] - b[7]));
        CKS(get(b[1] + b[9], b[2] + b[9] - b[0], 0, w, b[10] - b[9]));
    }
    if (s->top) {
        const ImageView tv = image_view(s->top);
        CKS(cudaMemcpy2D(hp, (size_t)stride_x, tv.plane[tv.cur], (size_t)tv.pitch * es, (size_t)pl.ll_w * es, pl.ll_h, cudaMemcpyDeviceToHost));
    }
    return DWTB200_OK;
}

}  // extern "C"
