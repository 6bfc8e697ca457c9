// kernels_ring.cu -- the streaming level kernels with their input staged through a deep shared-memory
// ring by the bulk-copy engine (sm_100a: cp.async.bulk + mbarrier, SASS UBLKCP).
//
// Why: the register double buffer of kernels_stream.cu keeps at most 2 KB per warp (32 KB per SM) in
// flight, and measurements (profiles/dbg_level0.py: the kernel runs equally long WITHOUT its lifting
// arithmetic) show level 0 is bound by exactly that -- memory latency x bytes in flight -- not by
// instruction issue or DRAM.  Here a dedicated producer warp per CTA keeps RING_SLOTS row pairs
// (12 KB) per consumer warp in flight through the asynchronous copy engine (192 KB per SM, no registers,
// no L1 allocation), and the consumer warps run the same fused row + column lifting as the register
// kernels, reading each row pair from shared memory once it has landed.
//
//   CTA = RING_CW consumer warps + 1 producer warp, 2 CTAs per SM.
//   consumer warp  = one (column group, row strip) exactly as in kernels_stream.cu: 30*VPL output
//                    columns, lanes 0 / 31 are halo lanes, column lifting carried in registers from row
//                    pair to row pair, four subband rows stored per iteration with 16-byte stores.
//   producer warp  = lane l feeds consumer warp l: for every row pair one expect_tx + two 1 KB bulk
//                    copies (rows mirrored at the top / bottom image border) into the next free slot;
//                    slots are recycled through a full / empty mbarrier pair each.
//   left / right image borders: the copy is clipped to the plane and the (warp-uniform) border path
//                    reads whole-sample-mirrored columns from the staged row.
//
// Same reference semantics and bit-identical results as kernels_stream.cu (see there for the
// /root/reference/src/libdwt.c lines each pass replaces).
#include <type_traits>
#include "chain.cuh"
#include "ring_common.cuh"
#include "stream_common.cuh"

namespace dwtb200 {

// CTA shapes (consumer warps per CTA, CTAs per SM).  The register file gives 128 registers per thread to
// 16 warps per SM, 96 to 20: with 8 + 1 warps twice per SM ptxas has to spill the subband pointers into
// the loop (measured: 104 us, long-scoreboard stalls on the LDLs), so the consumers get 7 warps.
template <int CW_, int NCTA_> struct RingCfg {
    static constexpr int CW = CW_, NCTA = NCTA_;
    static constexpr int THREADS = (CW + 1) * 32;
    static constexpr int ROWB = CW * 960 + 64;           // bytes of one staged row: CW windows of 30 lanes + two halo lanes
    static constexpr int SLOTB = 2 * ROWB;               // a row pair (inverse: four half-width subband segments)
    static constexpr int DATA = (RING_SLOTS * SLOTB + 127) / 128 * 128;
    static constexpr int SMEM = DATA + 2 * RING_SLOTS * 8;
    // inverse level reading the interleaved layout: five segments per slot (LL of row 2k from its dense band, the interleaved
    // rows 2k and 2k+1 at full width)
    static constexpr int SLOTB_IL = 5 * (SLOTB / 4);
    static constexpr int DATA_IL = (RING_SLOTS * SLOTB_IL + 127) / 128 * 128;
    static constexpr int SMEM_IL = DATA_IL + 2 * RING_SLOTS * 8;
};

// =====================================================================================================
// forward level
// =====================================================================================================
// CTA (band, strip): band = p.bw adjacent column groups (one consumer warp each), strip = p.pps row pairs.
template <class WV, int VPL, class CFG, bool IL = false> __global__ void __launch_bounds__(CFG::THREADS, CFG::NCTA) k_fwd_ring(const LevelParams p)
{
    using T = typename WV::T;
    static_assert(VPL * sizeof(T) == 32, "a lane holds 32 bytes of a row");
    constexpr int OUTW = 30 * VPL, HV = VPL / 2, ES = (int)sizeof(T);
    extern __shared__ __align__(128) unsigned char ring_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t ring0 = smem_u32(ring_smem);
    const uint32_t full = ring0 + CFG::DATA, empty = full + 8 * RING_SLOTS;
    const int band = blockIdx.x % p.nbands, strip = blockIdx.x / p.nbands + p.strip0;
    const int cg0 = band * p.bw, nact = min(p.bw, p.ncg - cg0);   // active consumer warps of this CTA
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
    const int xs0 = cg0 * OUTW - VPL;   // first column of the CTA's staged rows

    if (warp == CFG::CW) {
        // ---------------- producer: one thread feeds the CTA's ring ----------------
        if (lane != 0) return;
        const int c0 = max(xs0, 0), c1 = (int)min((int64_t)xs0 + nact * OUTW + 2 * VPL, p.src_pitch);   // clipped to the plane
        const uint32_t bytes = (uint32_t)(c1 - c0) * ES;
        const T *src = (const T *)p.src + (int64_t)blockIdx.y * p.src_frame;
        const uint32_t dst0 = ring0 + (c0 - xs0) * ES;
        const int nitems = m1 - m0 + 2;   // one single row, then one row pair per iteration
        RingState rs;
        ChainWindow win;
        const bool dep = p.chain.in != nullptr;
        for (int q = 0; q < nitems; q++) {
            if (q >= RING_SLOTS) mbar_wait(empty + 8 * rs.slot, rs.phase ^ 1);
            const uint32_t d = dst0 + rs.slot * CFG::SLOTB, fb = full + 8 * rs.slot;
            if (q == 0) {
                const int r = reflect(2 * m0, H);
                if (dep) win.need_row(p.chain, gen, blockIdx.y, r);
                mbar_expect_tx(fb, bytes);
                bulk_g2s(d + CFG::ROWB, level_src_row(p, src, r) + c0, bytes, fb);
            } else {
                const int m = m0 + q - 1;
                const int ra = reflect(2 * m + 1, H), rb = reflect(2 * m + 2, H);
                if (dep) {
                    win.need_row(p.chain, gen, blockIdx.y, ra);
                    win.need_row(p.chain, gen, blockIdx.y, rb);
                }
                mbar_expect_tx(fb, 2 * bytes);
                bulk_g2s(d, level_src_row(p, src, ra) + c0, bytes, fb);
                bulk_g2s(d + CFG::ROWB, level_src_row(p, src, rb) + c0, bytes, fb);
            }
            rs.next();
        }
        return;
    }

    // ---------------- consumers ----------------
    if (warp >= nact) return;
    const int cg = cg0 + warp;
    const int xl = cg * OUTW - VPL + lane * VPL;   // first column held by this lane (even)
    const uint32_t wofs = (uint32_t)warp * (OUTW * ES);   // this warp's window inside a staged row

    const bool fast = __all_sync(FULL, xl >= 0 && xl + VPL <= W);
    int cx[VPL];   // border path: byte offsets of the mirrored columns inside the staged row
    if (!fast) {
#pragma unroll
        for (int i = 0; i < VPL; i++) cx[i] = min(max(reflect(xl + i, W) - xs0, 0), nact * OUTW + 2 * VPL - 1) * ES;
    }
    auto read = [&](uint32_t row, T(&v)[VPL]) {
        if (fast) {
            lds_vec<T, VPL>(row + wofs + lane * 32, v);
        } else {
#pragma unroll
            for (int i = 0; i < VPL; i++) v[i] = lds_one<T>(row + cx[i]);
        }
    };

    // destinations: even columns -> L half (ll / lh), odd columns -> H half (hl / hh)
    const int cb = xl >> 1;
    const bool producer = lane >= 1 && lane <= 30 && xl < W;
    const bool whole = xl + VPL <= W;   // all VPL/2 L and H columns exist
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
    T a[VPL], b[VPL];

    RingState rs;
    mbar_wait(full + 8 * rs.slot, rs.phase);
    read(ring0 + rs.slot * CFG::SLOTB + CFG::ROWB, st[0]);
    __syncwarp();
    if (lane == 0) mbar_arrive(empty + 8 * rs.slot);
    rs.next();
    hfwd<WV, VPL>(st[0]);
#pragma unroll
    for (int s = 1; s < WV::NS; s++)
#pragma unroll
        for (int i = 0; i < VPL; i++) st[s][i] = T(0);

    for (int m = m0; m <= m1; m++) {
        mbar_wait(full + 8 * rs.slot, rs.phase);
        const uint32_t base = ring0 + rs.slot * CFG::SLOTB;
        read(base, a);
        read(base + CFG::ROWB, b);
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + 8 * rs.slot);   // the slot may be refilled once every warp has read it
        rs.next();
        hfwd<WV, VPL>(a);
        hfwd<WV, VPL>(b);
        T oL[VPL], oH[VPL];   // column-lifted low / high outputs of this iteration
        vfwd<WV, VPL>(a, b, st, oL, oH);
        const int kk = m - DELAY;
        if constexpr (IL) {   // interleaved layout: the lane's columns as they are, rows 2 kk and 2 kk + 1; LL also goes to `ll`
            if (kk >= k0 && producer) {
                T o[HV];
#pragma unroll
                for (int i = 0; i < HV; i++) o[i] = oL[2 * i];
                put(ll, p.ll_pitch, kk, o, p.nLx, whole);
                T *q = (T *)p.il + (int64_t)blockIdx.y * p.il_frame + (int64_t)(2 * kk) * p.il_pitch + xl;
                if (whole) {
                    st_vec<T, VPL>(q, oL);
                    if (kk < p.nHy) st_vec<T, VPL>(q + p.il_pitch, oH);
                } else {
#pragma unroll
                    for (int i = 0; i < VPL; i++)
                        if (xl + i < W) {
                            q[i] = oL[i];
                            if (kk < p.nHy) q[p.il_pitch + i] = oH[i];
                        }
                }
            }
            continue;
        }
        if (kk >= k0 && producer) {
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
    if (p.chain.gen) {   // the strip's LL rows are written: tell the next level (consumer warps only: named barrier)
        asm volatile("bar.sync 1, %0;" ::"r"(nact * 32) : "memory");
        if (threadIdx.x == 0) chain_signal(p.chain, blockIdx.y, strip);
    }
}

// =====================================================================================================
// inverse level
// =====================================================================================================
// A slot holds the four subband row segments one iteration consumes: [LL | HL] of coefficient row 2k and
// [LH | HH] of row 2k+1, half a staged row each.  Needs 16-byte aligned HL / HH column origins
// (p.sub_aligned); the host falls back to k_inv_level otherwise.
template <class WV, int VPL, class CFG, bool IL = false> __global__ void __launch_bounds__(CFG::THREADS, CFG::NCTA) k_inv_ring(const LevelParams p)
{
    using T = typename WV::T;
    static_assert(VPL * sizeof(T) == 32, "a lane holds 32 bytes of an output row");
    constexpr int OUTW = 30 * VPL, HV = VPL / 2, ES = (int)sizeof(T), SEGB = CFG::SLOTB / 4;
    constexpr int SLOTB = IL ? CFG::SLOTB_IL : CFG::SLOTB, DATA = IL ? CFG::DATA_IL : CFG::DATA;
    extern __shared__ __align__(128) unsigned char ring_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t ring0 = smem_u32(ring_smem);
    const uint32_t full = ring0 + DATA, empty = full + 8 * RING_SLOTS;
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
    const int cs0 = cg0 * (OUTW / 2) - HV;   // first subband column of the CTA's staged segments

    if (warp == CFG::CW) {
        // ---------------- producer ----------------
        if (lane != 0) return;
        const int c0 = max(cs0, 0);
        const int64_t e = (int64_t)cs0 + nact * (OUTW / 2) + 2 * HV;
        const uint32_t nll = (uint32_t)((int)min(e, p.ll_pitch) - c0) * ES;         // clipped to the pitched rows
        const uint32_t nh = (uint32_t)((int)min(e, (int64_t)p.h_room) - c0) * ES;
        const uint32_t nlh = (uint32_t)((int)min(e, p.sub_pitch) - c0) * ES;
        const T *ll = (const T *)p.ll + (int64_t)blockIdx.y * p.ll_frame + c0;
        const T *hl = (const T *)p.hl + (int64_t)blockIdx.y * p.sub_frame + c0;
        const T *lh = (const T *)p.lh + (int64_t)blockIdx.y * p.sub_frame + c0;
        const T *hh = (const T *)p.hh + (int64_t)blockIdx.y * p.sub_frame + c0;
        const uint32_t dst0 = ring0 + (c0 - cs0) * ES;
        const int nitems = kb - ka + 1;
        RingState rs;
        ChainWindow win;
        const bool dep = p.chain.in != nullptr;
        if constexpr (IL) {   // interleaved layout: LL of row 2k from its dense band, rows 2k and 2k+1 as they lie (full staged rows)
            const int xs0 = 2 * cs0, x0 = max(xs0, 0);
            const uint32_t nb = (uint32_t)((int)min((int64_t)xs0 + nact * OUTW + 2 * VPL, p.il_pitch) - x0) * ES;
            const T *il = (const T *)p.il + (int64_t)blockIdx.y * p.il_frame + x0;
            const uint32_t d0 = ring0 + SEGB + (x0 - xs0) * ES;
            for (int q = 0; q < nitems; q++) {
                if (q >= RING_SLOTS) mbar_wait(empty + 8 * rs.slot, rs.phase ^ 1);
                const uint32_t fb = full + 8 * rs.slot;
                const int k = ka + q;
                mbar_expect_tx(fb, nll + 2 * nb);
                bulk_g2s(dst0 + rs.slot * SLOTB, ll + (int64_t)(reflect(2 * k, H) >> 1) * p.ll_pitch, nll, fb);
                bulk_g2s(d0 + rs.slot * SLOTB, il + (int64_t)reflect(2 * k, H) * p.il_pitch, nb, fb);
                bulk_g2s(d0 + rs.slot * SLOTB + 2 * SEGB, il + (int64_t)reflect(2 * k + 1, H) * p.il_pitch, nb, fb);
                rs.next();
            }
        } else {
            for (int q = 0; q < nitems; q++) {
                if (q >= RING_SLOTS) mbar_wait(empty + 8 * rs.slot, rs.phase ^ 1);
                const uint32_t d = dst0 + rs.slot * CFG::SLOTB, fb = full + 8 * rs.slot;
                const int k = ka + q;
                const int ra = reflect(2 * k, H) >> 1, rb = reflect(2 * k + 1, H) >> 1;
                if (dep) win.need_row(p.chain, gen, blockIdx.y, ra);   // only the LL band is produced inside this transform
                mbar_expect_tx(fb, nll + nh + nlh + nh);
                bulk_g2s(d, ll + (int64_t)ra * p.ll_pitch, nll, fb);
                bulk_g2s(d + SEGB, hl + (int64_t)ra * p.sub_pitch, nh, fb);
                bulk_g2s(d + 2 * SEGB, lh + (int64_t)rb * p.sub_pitch, nlh, fb);
                bulk_g2s(d + 3 * SEGB, hh + (int64_t)rb * p.sub_pitch, nh, fb);
                rs.next();
            }
        }
        return;
    }

    // ---------------- consumers ----------------
    if (warp >= nact) return;
    const int cg = cg0 + warp;
    const int xl = cg * OUTW - VPL + lane * VPL;   // first OUTPUT column of this lane (even)
    const uint32_t wofs = (uint32_t)warp * (OUTW / 2 * ES);
    T *dst = (T *)p.dst + (int64_t)blockIdx.y * p.dst_frame;

    const bool fast = __all_sync(FULL, xl >= 0 && xl + VPL <= W);
    int cL[HV], cH[HV];   // border path: byte offsets of the mirrored subband columns inside a segment
    int cE[IL ? HV : 1];  // interleaved layout: byte offsets of the mirrored EVEN columns inside a staged interleaved row (cH: the odd ones)
    if (!fast) {
        const int hi = nact * (OUTW / 2) + 2 * HV - 1;
#pragma unroll
        for (int i = 0; i < HV; i++) {
            cL[i] = min(max((reflect(xl + 2 * i, W) >> 1) - cs0, 0), hi) * ES;
            if constexpr (IL) {
                cE[i] = min(max(reflect(xl + 2 * i, W) - 2 * cs0, 0), 2 * hi + 1) * ES;
                cH[i] = min(max(reflect(xl + 2 * i + 1, W) - 2 * cs0, 0), 2 * hi + 1) * ES;
            } else {
                cH[i] = min(max((reflect(xl + 2 * i + 1, W) >> 1) - cs0, 0), hi) * ES;
            }
        }
    }
    // one interleaved row: even positions from the L segment, odd positions from the H segment
    // (interleaved layout: `hi` is a staged interleaved row, of which only the odd columns are used)
    auto read = [&](uint32_t lo, uint32_t hi, T(&v)[VPL]) {
        T l[HV], h[HV];
        if (fast) {
            lds_vec<T, HV>(lo + wofs + lane * 16, l);
            if constexpr (IL) {
                T r[VPL];
                lds_vec<T, VPL>(hi + 2 * wofs + lane * 32, r);
#pragma unroll
                for (int i = 0; i < HV; i++) h[i] = r[2 * i + 1];
            } else {
                lds_vec<T, HV>(hi + wofs + lane * 16, h);
            }
        } else {
#pragma unroll
            for (int i = 0; i < HV; i++) {
                l[i] = lds_one<T>(lo + cL[i]);
                h[i] = lds_one<T>(hi + cH[i]);
            }
        }
#pragma unroll
        for (int i = 0; i < HV; i++) {
            v[2 * i] = l[i];
            v[2 * i + 1] = h[i];
        }
    };
    // interleaved layout: a staged interleaved row used as it is (coefficient row 2k+1: LH | HH)
    auto read_row = [&](uint32_t row, T(&v)[VPL]) {
        if (fast) {
            lds_vec<T, VPL>(row + 2 * wofs + lane * 32, v);
        } else {
#pragma unroll
            for (int i = 0; i < HV; i++) {
                v[2 * i] = lds_one<T>(row + cE[IL ? i : 0]);
                v[2 * i + 1] = lds_one<T>(row + cH[i]);
            }
        }
    };
    const bool producer = lane >= 1 && lane <= 30 && xl < W;
    const bool whole = xl + VPL <= W;
    auto put = [&](int row, const T(&o)[VPL]) {
        if (row < 0 || row >= H || !producer) return;
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
    T a[VPL], b[VPL];
#pragma unroll
    for (int s = 0; s < WV::NS; s++)
#pragma unroll
        for (int i = 0; i < VPL; i++) st[s][i] = T(0);

    RingState rs;
    for (int k = ka; k <= kb; k++) {
        mbar_wait(full + 8 * rs.slot, rs.phase);
        const uint32_t base = ring0 + rs.slot * SLOTB;
        read(base, base + SEGB, a);
        if constexpr (IL) read_row(base + 3 * SEGB, b);
        else read(base + 2 * SEGB, base + 3 * SEGB, b);
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + 8 * rs.slot);
        rs.next();
        if constexpr (!WV::INV_COLS_FIRST) {   // rows first (float / double): libdwt.c:17098 then 17127
            hinv<WV, VPL>(a);
            hinv<WV, VPL>(b);
        }
        T oO[VPL], oE[VPL];   // output rows 2q-1 (odd) and 2q (even)
        vinv<WV, VPL>(a, b, st, oO, oE);
        const int q = k - DELAY;
        if (q >= q0) {   // warp-uniform
            if constexpr (WV::INV_COLS_FIRST) {   // columns first (int): libdwt.c:18178 then 18187
                hinv<WV, VPL>(oO);
                hinv<WV, VPL>(oE);
            }
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
template <class K> static cudaError_t prep(K kern, int smem)
{
    cudaFuncAttributes a;
    cudaError_t e = cudaFuncGetAttributes(&a, kern);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    return e;
}
// cfg: 0 = 7 consumer warps x 2 CTAs per SM, 1 = 15 x 1, 2 = 8 x 2 (96 registers), 3 = 5 x 3 (rows of 9-10 column groups:
// two bands of 5 keep 15 instead of 10 consumer warps per SM busy)
constexpr int RING_NCFG = 4;
// kinds with an interleaved in-place family in the reference: float 9/7 and float 5/3
template <class WV> constexpr bool ring_il() { return std::is_same<WV, W97F>::value || std::is_same<WV, W53F>::value; }
bool ring_interleaved_ok(int kind) { return kind == K_CDF97_F32 || kind == K_CDF53_F32; }
// CTA shapes whose five-segment inverse ring still fits the SM's shared memory NCTA times (all but 8 x 2)
bool ring_interleaved_cfg_ok(int cfg) { return cfg != 2 && !ring2_cfg(cfg); }
template <class F> static void dispatch_cfg(int cfg, F &&f)
{
    if (cfg == 1) f(RingCfg<15, 1>{});
    else if (cfg == 2) f(RingCfg<8, 2>{});
    else if (cfg == 3) f(RingCfg<5, 3>{});
    else f(RingCfg<7, 2>{});
}
cudaError_t preload_ring()
{
    cudaError_t e = cudaSuccess;
    for (int kind = 0; kind < K_COUNT; kind++)
        dispatch_kind(kind, [&](auto wv) {
            using WV = decltype(wv);
            constexpr int V = 32 / (int)sizeof(typename WV::T);
            for (int cfg = 0; cfg < RING_NCFG; cfg++)
                dispatch_cfg(cfg, [&](auto c) {
                    using CFG = decltype(c);
                    if (e == cudaSuccess) e = prep(k_fwd_ring<WV, V, CFG>, CFG::SMEM);
                    if (e == cudaSuccess) e = prep(k_inv_ring<WV, V, CFG>, CFG::SMEM);
                    if constexpr (ring_il<WV>()) {
                        if (e == cudaSuccess) e = prep(k_fwd_ring<WV, V, CFG, true>, CFG::SMEM);
                        if (e == cudaSuccess && CFG::SMEM_IL * CFG::NCTA <= 226 * 1024) e = prep(k_inv_ring<WV, V, CFG, true>, CFG::SMEM_IL);
                    }
                });
        });
    return e;
}
int ring_warps_per_sm(int cfg) { return ring_cta_warps(cfg) * ring_ctas_per_sm(cfg); }

void launch_fwd_ring(int kind, const LevelParams &p, int frames, int cfg, cudaStream_t st)
{
    if (ring2_cfg(cfg) && !p.il) return launch_fwd_ring2(kind, p, frames, ring2_cfg_warps(cfg), st);
    dispatch_kind(kind, [&](auto wv) {
        using WV = decltype(wv);
        constexpr int V = 32 / (int)sizeof(typename WV::T);
        dispatch_cfg(cfg, [&](auto c) {
            using CFG = decltype(c);
            const dim3 grid(p.nbands * p.nstrips, frames);
            if constexpr (ring_il<WV>()) {
                if (p.il) {
                    launch_pdl(k_fwd_ring<WV, V, CFG, true>, grid, dim3(CFG::THREADS), (size_t)CFG::SMEM, st, p.chain.pdl, p);
                    return;
                }
            }
            launch_pdl(k_fwd_ring<WV, V, CFG>, grid, dim3(CFG::THREADS), (size_t)CFG::SMEM, st, p.chain.pdl, p);
        });
    });
}

void launch_inv_ring(int kind, const LevelParams &p, int frames, int cfg, cudaStream_t st)
{
    if (ring2_cfg(cfg) && !p.il && ring2_inverse_ok(kind)) return launch_inv_ring2(kind, p, frames, ring2_cfg_warps(cfg), st);
    dispatch_kind(kind, [&](auto wv) {
        using WV = decltype(wv);
        constexpr int V = 32 / (int)sizeof(typename WV::T);
        dispatch_cfg(cfg, [&](auto c) {
            using CFG = decltype(c);
            const dim3 grid(p.nbands * p.nstrips, frames);
            if constexpr (ring_il<WV>()) {
                if (p.il) {
                    launch_pdl(k_inv_ring<WV, V, CFG, true>, grid, dim3(CFG::THREADS), (size_t)CFG::SMEM_IL, st, p.chain.pdl, p);
                    return;
                }
            }
            launch_pdl(k_inv_ring<WV, V, CFG>, grid, dim3(CFG::THREADS), (size_t)CFG::SMEM, st, p.chain.pdl, p);
        });
    });
}
int ring_cta_warps(int cfg) { return ring2_cfg(cfg) ? ring2_cfg_warps(cfg) : cfg == 1 ? 15 : cfg == 2 ? 8 : cfg == 3 ? 5 : 7; }
int ring_ctas_per_sm(int cfg) { return ring2_cfg(cfg) ? 16 / ring2_cfg_warps(cfg) : cfg == 1 ? 1 : cfg == 3 ? 3 : 2; }

}  // namespace dwtb200
