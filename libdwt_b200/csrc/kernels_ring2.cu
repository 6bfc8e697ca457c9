// kernels_ring2.cu -- second generation of the bulk-copy ring kernels (sm_100a): 256-column warp windows, 8 consumer
// warps per CTA that take turns as the producer.
//
// kernels_ring.cu gives a warp a window of 256 staged columns of which lanes 0 and 31 are pure halo lanes (their only job
// is to hand the row-lifted neighbour samples to lanes 1 and 30), so a warp EMITS 240 columns and a power-of-two row never
// splits evenly: 4096 columns are 17.07 -> 18 column groups in three bands of 6 of the 7 consumer warps, and a level of that
// width ran at 0.76 of the HBM roofline where an 8192-wide one reaches 0.92 (profiles/ncu_ring_pyramid_r1.txt).  Here every
// lane emits its 8 columns:
//   * the CTA stages whole band rows (one bulk copy per row: CW windows plus 4 samples either side), so the few raw samples a
//     warp's edge lanes miss lie in shared memory next to the warp's own window;
//   * lane 0 / lane 31 evaluate the row-lifted samples just outside the window from those raw samples (window evaluation, the
//     same operations in the same order as their owner computes them: bit-identical) -- four extra lifting steps per row, issued
//     once for both edge lanes -- and feed them into the warp's shuffle chain where the neighbour lane's value used to come from;
//   * rows must be a whole number of 256-column windows (the power-of-two widths this kernel exists for; every other width stays
//     with kernels_ring.cu): no partially filled lanes and no border path in the loop, only the mirror images the first and the
//     last lane of a row need;
//   * 8 consumer warps x 256 columns: 2048 / 4096 / 8192 columns are exactly 1 / 2 / 4 bands of 8 warps.  A ninth (producer) warp
//     would leave only 96 registers per thread (two CTAs of 9 warps put 5 warps on one of the SM's four register-file partitions),
//     so the producer duty rotates: item q of the ring is issued by lane 0 of warp q mod 8, just before that warp waits for item
//     q - 5 itself.  (Measured and dropped: ALL copies issued by warp 0 between its own iterations: 270 instead of 100 us for
//     level 0 of 8192^2 -- a freed slot is refilled only when warp 0 comes round, and every warp is gated by the one that produces.)
// Arithmetic, mirror rule and pass order are those of kernels_ring.cu / kernels_stream.cu (see there for the
// /root/reference/src/libdwt.c lines each pass replaces: rows :10744 / :11530, drivers :12837-12893, :17098-17154).
// Inverse levels of the integer wavelets lift columns first (:18178), i.e. rows are lifted on column-lifted register values that
// have no copy in shared memory: they stay with kernels_ring.cu.
#include <type_traits>
#include "chain.cuh"
#include "ring_common.cuh"
#include "stream_common.cuh"

namespace dwtb200 {

// CW = warps (windows) per CTA: 8 for rows of at least 8 windows, 4 / 2 / 1 for the narrower levels of a batch of frames (a level of
// 1024 / 512 / 256 columns is 4 / 2 / 1 windows wide: an 8-warp CTA would run with half, a quarter, an eighth of its warps and the
// SM with as little of its occupancy); 16 / CW CTAs per SM keep 16 warps and ~200 KB of ring per SM in every shape.
template <class T, int CW_ = 8> struct R2 {
    static constexpr int ES = (int)sizeof(T), VPL = 32 / ES, HV = VPL / 2, OUTW = 32 * VPL;
    static constexpr int CW = CW_, NCTA = 16 / CW_, THREADS = CW * 32;   // all warps are consumers; they take turns as the producer
    static constexpr int HPAD = 4;                   // raw samples staged beyond either end of the band (the lifting reach)
    static constexpr int ROWB = CW * 1024 + 64;      // bytes of a staged band row (>= (CW * OUTW + 2 * HPAD) * ES)
    static constexpr int SLOTB = 2 * ROWB;           // a row pair; inverse: four subband segments of SEGB bytes
    static constexpr int HPS = 16 / ES;              // subband samples staged beyond either end of a segment (>= 2, 16 bytes)
    static constexpr int SEGB = ROWB / 2;            // >= (CW * OUTW / 2 + 2 * HPS) * ES
    static constexpr int DATA = RING_SLOTS * SLOTB;
    static constexpr int SMEM = DATA + 2 * RING_SLOTS * 8;
    static_assert(DATA % 128 == 0, "barriers must stay 8-byte aligned");
    static_assert((CW * OUTW + 2 * HPAD) * ES <= ROWB && (CW * OUTW / 2 + 2 * HPS) * ES <= SEGB, "staging");
};

template <class T> __device__ __forceinline__ void lds_pair(uint32_t addr, T &a, T &b)
{
    if constexpr (sizeof(T) == 4) {
        uint32_t x, y;
        asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(x), "=r"(y) : "r"(addr));
        a = *reinterpret_cast<T *>(&x);
        b = *reinterpret_cast<T *>(&y);
    } else {
        unsigned long long x, y;
        asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(x), "=l"(y) : "r"(addr));
        a = *reinterpret_cast<T *>(&x);
        b = *reinterpret_cast<T *>(&y);
    }
}
template <class T> __device__ __forceinline__ void lds_quad(uint32_t addr, T (&v)[4])
{
    if constexpr (sizeof(T) == 4) {
        lds_vec<T, 4>(addr, v);
    } else {
        lds_pair<T>(addr, v[0], v[1]);
        lds_pair<T>(addr + 16, v[2], v[3]);
    }
}

// ---- row lifting with the window's outside neighbours supplied by the edge lanes ---------------------------------------------
// forward: v = the lane's VPL raw samples (v[0] at an even column); xh = the 4 raw samples left of the window (lane 0: x[-4..-1]) or
// right of it (lane 31: x[VPL .. VPL+3], relative to the lane); other lanes: anything.
template <class WV, int VPL> __device__ __forceinline__ void hfwd_edge(typename WV::T (&v)[VPL], const typename WV::T (&xh)[4], int lane)
{
    using T = typename WV::T;
    const bool l0 = lane == 0, l31 = lane == 31;
    if constexpr (WV::NS == 4) {
        // d1 = step 0 (odd), s1 = step 1 (even), d2 = step 2 (odd), s2 = step 3 (even); indices relative to the lane's v[0]
        const T eA = WV::template f<0>(xh[1], xh[0], xh[2]);   // lane 0: d1[-3]      lane 31: d1[VPL+1]
        const T eB = WV::template f<0>(xh[3], xh[2], v[0]);    // lane 0: d1[-1]
        {   // step 0, odd samples
            T nxt = __shfl_down_sync(FULL, v[0], 1);
            if (l31) nxt = xh[0];
#pragma unroll
            for (int i = 1; i < VPL; i += 2) v[i] = WV::template f<0>(v[i], v[i - 1], (i + 1 < VPL) ? v[(i + 1) % VPL] : nxt);
        }
        const T eC = WV::template f<1>(l31 ? xh[0] : xh[2], l31 ? v[VPL - 1] : eA, l31 ? eA : eB);   // lane 0: s1[-2]   lane 31: s1[VPL]
        {   // step 1, even samples
            T prv = __shfl_up_sync(FULL, v[VPL - 1], 1);
            if (l0) prv = eB;
#pragma unroll
            for (int i = 0; i < VPL; i += 2) v[i] = WV::template f<1>(v[i], i ? v[(i + VPL - 1) % VPL] : prv, v[i + 1]);
        }
        const T eD = WV::template f<2>(eB, eC, v[0]);          // lane 0: d2[-1]
        {   // step 2, odd samples
            T nxt = __shfl_down_sync(FULL, v[0], 1);
            if (l31) nxt = eC;
#pragma unroll
            for (int i = 1; i < VPL; i += 2) v[i] = WV::template f<2>(v[i], v[i - 1], (i + 1 < VPL) ? v[(i + 1) % VPL] : nxt);
        }
        {   // step 3, even samples
            T prv = __shfl_up_sync(FULL, v[VPL - 1], 1);
            if (l0) prv = eD;
#pragma unroll
            for (int i = 0; i < VPL; i += 2) v[i] = WV::template f<3>(v[i], i ? v[(i + VPL - 1) % VPL] : prv, v[i + 1]);
        }
    } else {
        const T eB = WV::template f<0>(xh[3], xh[2], v[0]);    // lane 0: d1[-1]
        {
            T nxt = __shfl_down_sync(FULL, v[0], 1);
            if (l31) nxt = xh[0];
#pragma unroll
            for (int i = 1; i < VPL; i += 2) v[i] = WV::template f<0>(v[i], v[i - 1], (i + 1 < VPL) ? v[(i + 1) % VPL] : nxt);
        }
        {
            T prv = __shfl_up_sync(FULL, v[VPL - 1], 1);
            if (l0) prv = eB;
#pragma unroll
            for (int i = 0; i < VPL; i += 2) v[i] = WV::template f<1>(v[i], i ? v[(i + VPL - 1) % VPL] : prv, v[i + 1]);
        }
    }
#pragma unroll
    for (int i = 0; i < VPL; i += 2) {
        v[i] = WV::fse(v[i]);
        v[i + 1] = WV::fso(v[i + 1]);
    }
}
// inverse: v = the lane's VPL interleaved coefficients (v[0] an L coefficient); yh = the 4 interleaved coefficients left of the
// window (lane 0: y[-4..-1]) or right of it (lane 31: y[VPL .. VPL+3]); yh[0], yh[2] are L, yh[1], yh[3] H coefficients either way.
template <class WV, int VPL> __device__ __forceinline__ void hinv_edge(typename WV::T (&v)[VPL], typename WV::T (&yh)[4], int lane)
{
    using T = typename WV::T;
    const bool l0 = lane == 0, l31 = lane == 31;
#pragma unroll
    for (int i = 0; i < VPL; i += 2) {
        v[i] = WV::ise(v[i]);
        v[i + 1] = WV::iso(v[i + 1]);
    }
    yh[0] = WV::ise(yh[0]);
    yh[1] = WV::iso(yh[1]);
    yh[2] = WV::ise(yh[2]);
    yh[3] = WV::iso(yh[3]);
    if constexpr (WV::NS == 4) {
        // e1 = step 0 (even), o1 = step 1 (odd), e2 = step 2 (even), o2 = step 3 (odd)
        const T eA = WV::template i<0>(yh[2], yh[1], yh[3]);       // lane 0: e1[-2]      lane 31: e1[VPL+2]
        const T eB = WV::template i<0>(yh[0], v[VPL - 1], yh[1]);  // lane 31: e1[VPL]
        {   // step 0, even samples
            T prv = __shfl_up_sync(FULL, v[VPL - 1], 1);
            if (l0) prv = yh[3];
#pragma unroll
            for (int i = 0; i < VPL; i += 2) v[i] = WV::template i<0>(v[i], i ? v[(i + VPL - 1) % VPL] : prv, v[i + 1]);
        }
        const T eC = WV::template i<1>(l31 ? yh[1] : yh[3], l31 ? eB : eA, l31 ? eA : v[0]);   // lane 0: o1[-1]   lane 31: o1[VPL+1]
        {   // step 1, odd samples
            T nxt = __shfl_down_sync(FULL, v[0], 1);
            if (l31) nxt = eB;
#pragma unroll
            for (int i = 1; i < VPL; i += 2) v[i] = WV::template i<1>(v[i], v[i - 1], (i + 1 < VPL) ? v[(i + 1) % VPL] : nxt);
        }
        const T eD = WV::template i<2>(eB, v[VPL - 1], eC);        // lane 31: e2[VPL]
        {   // step 2, even samples
            T prv = __shfl_up_sync(FULL, v[VPL - 1], 1);
            if (l0) prv = eC;
#pragma unroll
            for (int i = 0; i < VPL; i += 2) v[i] = WV::template i<2>(v[i], i ? v[(i + VPL - 1) % VPL] : prv, v[i + 1]);
        }
        {   // step 3, odd samples
            T nxt = __shfl_down_sync(FULL, v[0], 1);
            if (l31) nxt = eD;
#pragma unroll
            for (int i = 1; i < VPL; i += 2) v[i] = WV::template i<3>(v[i], v[i - 1], (i + 1 < VPL) ? v[(i + 1) % VPL] : nxt);
        }
    } else {
        const T eB = WV::template i<0>(yh[0], v[VPL - 1], yh[1]);  // lane 31: e1[VPL]
        {
            T prv = __shfl_up_sync(FULL, v[VPL - 1], 1);
            if (l0) prv = yh[3];
#pragma unroll
            for (int i = 0; i < VPL; i += 2) v[i] = WV::template i<0>(v[i], i ? v[(i + VPL - 1) % VPL] : prv, v[i + 1]);
        }
        {
            T nxt = __shfl_down_sync(FULL, v[0], 1);
            if (l31) nxt = eB;
#pragma unroll
            for (int i = 1; i < VPL; i += 2) v[i] = WV::template i<1>(v[i], v[i - 1], (i + 1 < VPL) ? v[(i + 1) % VPL] : nxt);
        }
    }
}

// =====================================================================================================
// forward level
// =====================================================================================================
// CTA (band, strip): band = p.bw adjacent column groups of 32 * VPL columns (one warp each), strip = p.pps row pairs.
template <class WV, int CW> __global__ void __launch_bounds__(R2<typename WV::T, CW>::THREADS, R2<typename WV::T, CW>::NCTA) k_fwd_ring2(const LevelParams p)
{
    using T = typename WV::T;
    using C = R2<T, CW>;
    constexpr int VPL = C::VPL, HV = C::HV, OUTW = C::OUTW, ES = C::ES, HPAD = C::HPAD;
    extern __shared__ __align__(128) unsigned char ring_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t ring0 = smem_u32(ring_smem);
    const uint32_t full = ring0 + C::DATA, empty = full + 8 * RING_SLOTS;
    const int band = blockIdx.x % p.nbands, strip = blockIdx.x / p.nbands + p.strip0;
    const int cg0 = band * p.bw, nact = min(p.bw, p.ncg - cg0);   // active warps of this CTA
    if (threadIdx.x == 0) {
        for (int i = 0; i < RING_SLOTS; i++) {
            mbar_init(full + 8 * i, 1);
            mbar_init(empty + 8 * i, nact);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint32_t gen = chain_begin(p.chain);

    constexpr int WARM = WV::NS / 2 + (WV::NS == 4 ? 1 : 0);   // warm-up iterations: 3 (9/7) or 1 (5/3)
    constexpr int DELAY = WV::NS / 2 - 1;                      // iteration m emits pair m - DELAY
    const int W = p.W, H = p.H;
    const int k0 = strip * p.pps, k1 = min(k0 + p.pps, p.nLy);
    const int m0 = k0 + DELAY - WARM;   // first iteration
    const int m1 = k1 - 1 + DELAY;      // last iteration (inclusive)
    const int xs0 = cg0 * OUTW - HPAD;  // first column of the CTA's staged rows
    const int stw = nact * OUTW + 2 * HPAD;   // staged columns

    if (warp >= nact) return;

    // ---------------- producer duty: item q is issued by lane 0 of warp q % nact ----------------
    // Item 0 is the single row a strip starts with, item q >= 1 the row pair of iteration m0 + q - 1; slot and barrier phase follow
    // from q alone.  Before a warp waits for item `it` it issues those of its items up to it + RING_SLOTS - 1 whose slot is free and
    // (chained level) whose rows the previous level has written -- WITHOUT waiting: a consumer that blocked there would stop
    // draining the ring for everybody -- and only for an item it is about to wait for itself does it wait.  Every warp reaches
    // every item, so every item is issued by then at the latest, and what a waiting warp waits for never depends on itself.
    const int nitems = m1 - m0 + 2;
    ChainWindow win;
    int nq = warp;   // the next item this warp has to issue
    // issue item q unless its ring slot is still being read or (chained level) its rows are not written yet; `must`: wait for both
    auto try_issue = [&](int q, bool must) -> bool {
        const int slot = q % RING_SLOTS, use = q / RING_SLOTS;
        if (use && !mbar_test(empty + 8 * slot, (use & 1) ^ 1)) {
            if (!must) return false;
            mbar_wait(empty + 8 * slot, (use & 1) ^ 1);
        }
        const int m = m0 + q - 1;
        const int ra = reflect(q ? 2 * m + 1 : 2 * m0, H), rb = reflect(2 * m + 2, H);   // item 0: the single row 2 m0
        if (p.chain.in != nullptr) {
            if (must) {
                win.need_row(p.chain, gen, blockIdx.y, ra);
                if (q) win.need_row(p.chain, gen, blockIdx.y, rb);
            } else if (!win.try_row(p.chain, gen, blockIdx.y, ra) || (q && !win.try_row(p.chain, gen, blockIdx.y, rb))) {
                return false;
            }
        }
        const int c0 = max(xs0, 0), c1 = (int)min((int64_t)xs0 + stw, p.src_pitch);   // clipped to the plane
        const uint32_t bytes = (uint32_t)(c1 - c0) * ES;
        const T *src = (const T *)p.src + (int64_t)blockIdx.y * p.src_frame;
        const uint32_t d = ring0 + (c0 - xs0) * ES + slot * C::SLOTB, fb = full + 8 * slot;
        if (q == 0) {
            mbar_expect_tx(fb, bytes);
            bulk_g2s(d + C::ROWB, level_src_row(p, src, ra) + c0, bytes, fb);
        } else {
            mbar_expect_tx(fb, 2 * bytes);
            bulk_g2s(d, level_src_row(p, src, ra) + c0, bytes, fb);
            bulk_g2s(d + C::ROWB, level_src_row(p, src, rb) + c0, bytes, fb);
        }
        return true;
    };
    // before waiting for item `it`: this warp's items up to it + RING_SLOTS - 1 that can be issued, and item `it` itself whatever it takes
    auto duty = [&](int it) {
        if (nq < nitems && nq <= it + RING_SLOTS - 1) {
            if (lane == 0)
                while (nq < nitems && nq <= it + RING_SLOTS - 1 && try_issue(nq, nq <= it)) nq += nact;
            nq = __shfl_sync(FULL, nq, 0);
        }
    };
    duty(0);

    // ---------------- consumers ----------------
    const int cg = cg0 + warp;
    const int xl = cg * OUTW + lane * VPL;                       // first column held by this lane (even)
    const uint32_t wofs = (uint32_t)warp * (OUTW * ES);          // this warp's window inside a staged row (after the left pad)
    // The row is a whole number of windows (W % OUTW == 0: the host sends every other width to kernels_ring.cu), so every lane's own
    // columns exist; the only samples outside the row are the outside neighbours of the row's first and last lane, which are the
    // whole-sample mirror images x[-1-i] = x[1+i] and x[W+i] = x[W-2-i]: read backwards from `hofs`.
    const bool edge = lane == 0 || lane == 31;
    const bool hmir = lane == 0 ? cg == 0 : (cg + 1) * OUTW >= W;
    const uint32_t hofs = hmir ? (lane == 0 ? (uint32_t)((HPAD + 4) * ES) : (uint32_t)((W - 2 - xs0) * ES))   // x[4] resp. x[W-2] ...
                               : (lane == 31 ? wofs + HPAD * ES + OUTW * ES : wofs);                          // ... or where the neighbours lie
    auto read_own = [&](uint32_t row, T(&v)[VPL]) { lds_vec<T, VPL>(row + HPAD * ES + wofs + lane * 32, v); };
    auto read_out = [&](uint32_t row, T(&xh)[4]) {
        if (edge) {
            if (!hmir) {
                lds_quad<T>(row + hofs, xh);
            } else {   // lane 0: x[-4 .. -1] = x[4], x[3], x[2], x[1]; lane 31: x[W .. W+3] = x[W-2], x[W-3], x[W-4], x[W-5]
#pragma unroll
                for (int i = 0; i < 4; i++) xh[i] = lds_one<T>(row + hofs - i * ES);
            }
        }
    };

    // destinations: even columns -> L half (ll / lh), odd columns -> H half (hl / hh)
    const int cb = xl >> 1;
    constexpr bool live = true, whole = true;   // W % OUTW == 0
    T *ll = (T *)p.ll + (int64_t)blockIdx.y * p.ll_frame + cb;
    T *hl = (T *)p.hl + (int64_t)blockIdx.y * p.sub_frame + cb;
    T *lh = (T *)p.lh + (int64_t)blockIdx.y * p.sub_frame + cb;
    T *hh = (T *)p.hh + (int64_t)blockIdx.y * p.sub_frame + cb;
    const bool vec_sub = whole && p.sub_aligned;
    auto put = [&](T *base, int64_t pitch, int row, const T(&o)[HV], int limit, bool vec) {
        T *q = base + (int64_t)row * pitch;
        if (vec) {
            st_vec<T, HV>(q, o);
        } else {
#pragma unroll
            for (int i = 0; i < HV; i++)
                if (cb + i < limit) q[i] = o[i];
        }
    };

    T st[WV::NS][VPL];   // NS==4: xe, d1, s1, d2     NS==2: xe, d1
    T a[VPL], b[VPL], xa[4];
#pragma unroll
    for (int i = 0; i < 4; i++) xa[i] = T(0);

    RingState rs;
    mbar_wait(full + 8 * rs.slot, rs.phase);
    read_own(ring0 + rs.slot * C::SLOTB + C::ROWB, st[0]);
    read_out(ring0 + rs.slot * C::SLOTB + C::ROWB, xa);
    __syncwarp();
    if (lane == 0) mbar_arrive(empty + 8 * rs.slot);
    rs.next();
    hfwd_edge<WV, VPL>(st[0], xa, lane);
#pragma unroll
    for (int s = 1; s < WV::NS; s++)
#pragma unroll
        for (int i = 0; i < VPL; i++) st[s][i] = T(0);

    for (int m = m0; m <= m1; m++) {
        duty(m - m0 + 1);   // this iteration consumes item m - m0 + 1
        mbar_wait(full + 8 * rs.slot, rs.phase);
        const uint32_t base = ring0 + rs.slot * C::SLOTB;
        read_own(base, a);
        read_own(base + C::ROWB, b);
        read_out(base, xa);   // the edge lanes' outside neighbours one row at a time: both rows at once would cost 4 more registers
        hfwd_edge<WV, VPL>(a, xa, lane);
        read_out(base + C::ROWB, xa);
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + 8 * rs.slot);   // the slot may be refilled once every warp has read it
        rs.next();
        hfwd_edge<WV, VPL>(b, xa, lane);
        T oL[VPL], oH[VPL];   // column-lifted low / high outputs of this iteration
        vfwd<WV, VPL>(a, b, st, oL, oH);
        const int kk = m - DELAY;
        if (kk >= k0 && live) {
            T o[HV];
#pragma unroll
            for (int i = 0; i < HV; i++) o[i] = oL[2 * i];
            put(ll, p.ll_pitch, kk, o, p.nLx, whole);
#pragma unroll
            for (int i = 0; i < HV; i++) o[i] = oL[2 * i + 1];
            put(hl, p.sub_pitch, kk, o, p.nHx, vec_sub);
            if (kk < p.nHy) {
#pragma unroll
                for (int i = 0; i < HV; i++) o[i] = oH[2 * i];
                put(lh, p.sub_pitch, kk, o, p.nLx, whole);
#pragma unroll
                for (int i = 0; i < HV; i++) o[i] = oH[2 * i + 1];
                put(hh, p.sub_pitch, kk, o, p.nHx, vec_sub);
            }
        }
    }
    if (p.chain.gen) {   // the strip's LL rows are written: tell the next level (active warps only: named barrier)
        asm volatile("bar.sync 1, %0;" ::"r"(nact * 32) : "memory");
        if (threadIdx.x == 0) chain_signal(p.chain, blockIdx.y, strip);
    }
}

// =====================================================================================================
// inverse level (rows first: the float and double wavelets)
// =====================================================================================================
// A slot holds the four subband row segments one iteration consumes: [LL | HL] of coefficient row 2k and [LH | HH] of row 2k+1.
// Needs 16-byte aligned HL / HH column origins (p.sub_aligned); the host falls back to k_inv_level otherwise.
template <class WV, int CW> __global__ void __launch_bounds__(R2<typename WV::T, CW>::THREADS, R2<typename WV::T, CW>::NCTA) k_inv_ring2(const LevelParams p)
{
    using T = typename WV::T;
    using C = R2<T, CW>;
    static_assert(!WV::INV_COLS_FIRST, "columns-first inverses lift rows on register values: kernels_ring.cu");
    constexpr int VPL = C::VPL, HV = C::HV, OUTW = C::OUTW, ES = C::ES, HPS = C::HPS, SEGB = C::SEGB, SW = OUTW / 2;
    extern __shared__ __align__(128) unsigned char ring_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t ring0 = smem_u32(ring_smem);
    const uint32_t full = ring0 + C::DATA, empty = full + 8 * RING_SLOTS;
    const int band = blockIdx.x % p.nbands, strip = blockIdx.x / p.nbands + p.strip0;
    const int cg0 = band * p.bw, nact = min(p.bw, p.ncg - cg0);
    if (threadIdx.x == 0) {
        for (int i = 0; i < RING_SLOTS; i++) {
            mbar_init(full + 8 * i, 1);
            mbar_init(empty + 8 * i, nact);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint32_t gen = chain_begin(p.chain);

    constexpr int DELAY = WV::NS / 2 - 1;   // iteration k emits rows 2(k-DELAY)-1 and 2(k-DELAY)
    constexpr int WARM = WV::NS;            // warm-up iterations
    const int W = p.W, H = p.H;
    const int q0 = strip * p.pps, q1 = min(q0 + p.pps, (H >> 1) + 1);   // emitted q = k - DELAY in [q0, q1)
    const int ka = q0 + DELAY - WARM, kb = q1 - 1 + DELAY;
    const int cs0 = cg0 * SW - HPS;          // first subband column of the CTA's staged segments
    const int stw = nact * SW + 2 * HPS;     // staged subband columns

    if (warp >= nact) return;

    // ---------------- producer duty: item q (iteration ka + q) is issued by lane 0 of warp q % nact (see k_fwd_ring2) ----------------
    const int nitems = kb - ka + 1;
    ChainWindow win;
    int nq = warp;
    auto try_issue = [&](int q, bool must) -> bool {
        const int slot = q % RING_SLOTS, use = q / RING_SLOTS;
        if (use && !mbar_test(empty + 8 * slot, (use & 1) ^ 1)) {
            if (!must) return false;
            mbar_wait(empty + 8 * slot, (use & 1) ^ 1);
        }
        const int k = ka + q;
        const int ra = reflect(2 * k, H) >> 1, rb = reflect(2 * k + 1, H) >> 1;
        if (p.chain.in != nullptr) {   // only the LL band is produced inside this transform
            if (must) win.need_row(p.chain, gen, blockIdx.y, ra);
            else if (!win.try_row(p.chain, gen, blockIdx.y, ra)) return false;
        }
        const int c0 = max(cs0, 0);
        const int64_t e = (int64_t)cs0 + stw;
        const uint32_t nll = (uint32_t)((int)min(e, p.ll_pitch) - c0) * ES;         // clipped to the pitched rows
        const uint32_t nh = (uint32_t)((int)min(e, (int64_t)p.h_room) - c0) * ES;
        const uint32_t nlh = (uint32_t)((int)min(e, p.sub_pitch) - c0) * ES;
        const int64_t fo = (int64_t)blockIdx.y * p.sub_frame + c0;
        const uint32_t d = ring0 + (c0 - cs0) * ES + slot * C::SLOTB, fb = full + 8 * slot;
        mbar_expect_tx(fb, nll + nh + nlh + nh);
        bulk_g2s(d, (const T *)p.ll + (int64_t)blockIdx.y * p.ll_frame + c0 + (int64_t)ra * p.ll_pitch, nll, fb);
        bulk_g2s(d + SEGB, (const T *)p.hl + fo + (int64_t)ra * p.sub_pitch, nh, fb);
        bulk_g2s(d + 2 * SEGB, (const T *)p.lh + fo + (int64_t)rb * p.sub_pitch, nlh, fb);
        bulk_g2s(d + 3 * SEGB, (const T *)p.hh + fo + (int64_t)rb * p.sub_pitch, nh, fb);
        return true;
    };
    auto duty = [&](int it) {
        if (nq < nitems && nq <= it + RING_SLOTS - 1) {
            if (lane == 0)
                while (nq < nitems && nq <= it + RING_SLOTS - 1 && try_issue(nq, nq <= it)) nq += nact;
            nq = __shfl_sync(FULL, nq, 0);
        }
    };

    // ---------------- consumers ----------------
    const int cg = cg0 + warp;
    const int xl = cg * OUTW + lane * VPL;   // first OUTPUT column of this lane (even)
    const uint32_t wofs = (uint32_t)(HPS * ES) + (uint32_t)warp * (SW * ES) + (uint32_t)lane * (HV * ES);   // the lane's HV subband columns in a segment
    const uint32_t hofs = lane == 31 ? wofs + HV * ES : wofs - 2 * ES;                                      // the two subband columns outside the window
    T *dst = (T *)p.dst + (int64_t)blockIdx.y * p.dst_frame;
    // W % OUTW == 0 (see k_fwd_ring2): only the outside neighbours of the row's first and last lane are mirror images,
    // y[-1-i] = y[1+i] and y[W+i] = y[W-2-i] (W even: y[W-1] is the last H coefficient)
    const bool edge = lane == 0 || lane == 31;
    const bool hmir = lane == 0 ? cg == 0 : (cg + 1) * OUTW >= W;
    // one interleaved row: even positions from the L segment `lo`, odd positions from the H segment `hi`
    auto read_own = [&](uint32_t lo, uint32_t hi, T(&v)[VPL]) {
        T l[HV], h[HV];
        lds_vec<T, HV>(lo + wofs, l);
        lds_vec<T, HV>(hi + wofs, h);
#pragma unroll
        for (int i = 0; i < HV; i++) {
            v[2 * i] = l[i];
            v[2 * i + 1] = h[i];
        }
    };
    auto read_out = [&](uint32_t lo, uint32_t hi, T(&yh)[4]) {
        if (edge) {
            if (!hmir) {
                lds_pair<T>(lo + hofs, yh[0], yh[2]);
                lds_pair<T>(hi + hofs, yh[1], yh[3]);
            } else if (lane == 0) {   // y[-4 .. -1] = y[4], y[3], y[2], y[1] = L[2], H[1], L[1], H[0]
                const uint32_t o = (uint32_t)(HPS * ES);
                yh[0] = lds_one<T>(lo + o + 2 * ES);
                yh[1] = lds_one<T>(hi + o + ES);
                yh[2] = lds_one<T>(lo + o + ES);
                yh[3] = lds_one<T>(hi + o);
            } else {                  // y[W .. W+3] = y[W-2], y[W-3], y[W-4], y[W-5] = L[n-1], H[n-2], L[n-2], H[n-3], n = W / 2
                const uint32_t o = (uint32_t)((W / 2 - cs0) * ES);
                yh[0] = lds_one<T>(lo + o - ES);
                yh[1] = lds_one<T>(hi + o - 2 * ES);
                yh[2] = lds_one<T>(lo + o - 2 * ES);
                yh[3] = lds_one<T>(hi + o - 3 * ES);
            }
        }
    };
    constexpr bool live = true, whole = true;   // W % OUTW == 0
    auto put = [&](int row, const T(&o)[VPL]) {
        if (row < 0 || row >= H || !live) return;
        T *q = dst + (int64_t)row * p.dst_pitch + xl;
        if (whole) {
            st_vec<T, VPL>(q, o);
        } else {
#pragma unroll
            for (int i = 0; i < VPL; i++)
                if (xl + i < W) q[i] = o[i];
        }
    };

    T st[WV::NS][VPL];   // NS==4: d2p, s1p, d1p, xep      NS==2: cp, xep
    T a[VPL], b[VPL], ya[4];
#pragma unroll
    for (int s = 0; s < WV::NS; s++)
#pragma unroll
        for (int i = 0; i < VPL; i++) st[s][i] = T(0);
#pragma unroll
    for (int i = 0; i < 4; i++) ya[i] = T(0);

    RingState rs;
    for (int k = ka; k <= kb; k++) {
        duty(k - ka);   // this iteration consumes item k - ka
        mbar_wait(full + 8 * rs.slot, rs.phase);
        const uint32_t base = ring0 + rs.slot * C::SLOTB;
        read_own(base, base + SEGB, a);
        read_own(base + 2 * SEGB, base + 3 * SEGB, b);
        read_out(base, base + SEGB, ya);
        hinv_edge<WV, VPL>(a, ya, lane);   // rows first (float / double): libdwt.c:17098 then 17127
        read_out(base + 2 * SEGB, base + 3 * SEGB, ya);
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + 8 * rs.slot);
        rs.next();
        hinv_edge<WV, VPL>(b, ya, lane);
        T oO[VPL], oE[VPL];   // output rows 2q-1 (odd) and 2q (even)
        vinv<WV, VPL>(a, b, st, oO, oE);
        const int q = k - DELAY;
        if (q >= q0) {   // warp-uniform
            put(2 * q - 1, oO);
            put(2 * q, oE);
        }
    }
    if (p.chain.gen) {
        asm volatile("bar.sync 1, %0;" ::"r"(nact * 32) : "memory");
        if (threadIdx.x == 0) chain_signal(p.chain, blockIdx.y, strip);
    }
}

// ---- launchers -------------------------------------------------------------------------------
template <class K> static cudaError_t prep2(K kern, int smem)
{
    cudaFuncAttributes a;
    cudaError_t e = cudaFuncGetAttributes(&a, kern);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    // two CTAs of 97 KB per SM need the largest shared-memory carve-out (228 KB); left to its default the driver sizes the carve-out
    // for ONE such CTA and the kernel runs at half its occupancy (measured: 270 instead of 100 us for level 0 of 8192^2)
    if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
    return e;
}
template <class F> static void dispatch_cw(int cw, F &&f)
{
    if (cw <= 1) f(std::integral_constant<int, 1>{});
    else if (cw == 2) f(std::integral_constant<int, 2>{});
    else if (cw <= 4) f(std::integral_constant<int, 4>{});
    else f(std::integral_constant<int, 8>{});
}
cudaError_t preload_ring2()
{
    cudaError_t e = cudaSuccess;
    for (int kind = 0; kind < K_COUNT; kind++)
        dispatch_kind(kind, [&](auto wv) {
            using WV = decltype(wv);
            if constexpr (sizeof(typename WV::T) == 4) {   // the kinds ring2_width_ok admits
                for (int cw = 1; cw <= 8; cw *= 2)
                    dispatch_cw(cw, [&](auto c) {
                        constexpr int CW = decltype(c)::value;
                        using C = R2<typename WV::T, CW>;
                        if (e == cudaSuccess) e = prep2(k_fwd_ring2<WV, CW>, C::SMEM);
                        if constexpr (!WV::INV_COLS_FIRST) {
                            if (e == cudaSuccess) e = prep2(k_inv_ring2<WV, CW>, C::SMEM);
                        }
                    });
            }
        });
    return e;
}
bool ring2_inverse_ok(int kind) { return !(kind == K_CDF53_I32 || kind == K_CDF97_I32); }
// the kernels take rows that are a whole number of warp windows (2048, 4096, 8192 ... samples): no partial lanes, no border path
// ... of 4-byte samples: measured (profiles/ring_gen2_vs_gen1_r2.txt), the double-precision levels are 2 % slower than with
// kernels_ring.cu (a window is only 128 columns wide there, and the edge evaluation costs as much as for 256)
bool ring2_width_ok(int kind, int W) { return kind_elem_size(kind) == 4 && W % ring2_out_width(kind) == 0; }
int ring2_out_width(int kind) { return 32 * (32 / kind_elem_size(kind)); }

void launch_fwd_ring2(int kind, const LevelParams &p, int frames, int cw, cudaStream_t st)
{
    dispatch_kind(kind, [&](auto wv) {
        using WV = decltype(wv);
        if constexpr (sizeof(typename WV::T) == 4)
            dispatch_cw(cw, [&](auto c) {
                constexpr int CW = decltype(c)::value;
                using C = R2<typename WV::T, CW>;
                launch_pdl(k_fwd_ring2<WV, CW>, dim3(p.nbands * p.nstrips, frames), dim3(C::THREADS), (size_t)C::SMEM, st, p.chain.pdl, p);
            });
    });
}
void launch_inv_ring2(int kind, const LevelParams &p, int frames, int cw, cudaStream_t st)
{
    dispatch_kind(kind, [&](auto wv) {
        using WV = decltype(wv);
        if constexpr (sizeof(typename WV::T) == 4 && !WV::INV_COLS_FIRST)
            dispatch_cw(cw, [&](auto c) {
                constexpr int CW = decltype(c)::value;
                using C = R2<typename WV::T, CW>;
                launch_pdl(k_inv_ring2<WV, CW>, dim3(p.nbands * p.nstrips, frames), dim3(C::THREADS), (size_t)C::SMEM, st, p.chain.pdl, p);
            });
    });
}
// the CTA shapes of this generation as ring cfg numbers: RING_CFG_V2 (8 warps) and RING_CFG_V2 + 2, + 3, + 4 (4, 2, 1 warps)
int ring2_cfg_for(int ncg) { return ncg >= 5 ? RING_CFG_V2 : ncg >= 3 ? RING_CFG_V2 + 2 : ncg == 2 ? RING_CFG_V2 + 3 : RING_CFG_V2 + 4; }
int ring2_cfg_warps(int cfg) { return cfg == RING_CFG_V2 ? 8 : cfg == RING_CFG_V2 + 2 ? 4 : cfg == RING_CFG_V2 + 3 ? 2 : 1; }
bool ring2_cfg(int cfg) { return cfg == RING_CFG_V2 || (cfg >= RING_CFG_V2 + 2 && cfg <= RING_CFG_V2 + 4); }

}  // namespace dwtb200
