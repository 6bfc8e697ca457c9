"""Row-strip partition of ONE image over several GPUs (BASELINE config 5a).

The reference has no multi-device code; this is the B200 design for an image too large (or too slow) for one
GPU, behind the same transform semantics (dwt_cdf97_2f_s / dwt_cdf53_2f_i ..., Mallat layout).

Scheme ("exchange once, recompute the halo"):
  * rank r owns image rows [R_r, R_{r+1}), R_r a multiple of A = 2^Jd (Jd = levels done distributed);
  * before the transform every rank receives HALO = 4*A rows of the level-0 INPUT from each neighbour
    (one exchange per transform instead of one per level: 4 level-resolution rows of lifting reach per
    level, summed over the levels, stay below 4*A input rows), so its extended strip is
    [R_r - HALO, R_{r+1} + HALO) clipped to the image;
  * it runs the ordinary Jd-level transform on the extended strip as if it were an image: the kernels'
    mirror at the two artificial borders is wrong, but the error travels at most 4 rows per level at
    that level's resolution and never reaches an owned row; at the true image top and bottom the strip
    border IS the image border and the mirror is the reference's;
  * the owned rows of every subband are exact (bit-identical to the single-device transform); the
    LL_Jd strips are gathered on rank 0, which runs the remaining levels as an ordinary transform.
The inverse is the mirror image: rank 0 inverts the top levels, the LL_Jd band and the subband rows of
every distributed level are exchanged with the same halo, every rank inverts Jd levels locally.

The exchange goes through a `Comm` object: `DistComm` (torch.distributed send/recv: NCCL over NVLink on
GPUs, gloo in the CPU tests) or `LocalComm` (all ranks emulated in one process).  The arithmetic goes
through an `Engine`: `DeviceEngine` (libdwtb200) in the product, an oracle-backed engine only in tests.
"""
import numpy as np


def ceil_div_pow2(v, j):
    return (v + (1 << j) - 1) >> j


class StripPlan:
    """Geometry of the partition: which rows each rank owns / holds at every level."""

    def __init__(self, width, height, world, levels_distributed, halo_lines=4):
        self.W, self.H, self.G, self.Jd = width, height, world, levels_distributed
        self.A = 1 << levels_distributed
        self.halo = halo_lines * self.A
        units = ceil_div_pow2(height, levels_distributed)       # strips are whole multiples of A rows
        per = -(-units // world)
        self.R = [min(min(r * per, units) * self.A, height) for r in range(world)] + [height]

    def owned(self, r):
        return self.R[r], self.R[r + 1]

    def extended(self, r):
        a, b = self.owned(r)
        return max(0, a - self.halo), min(self.H, b + self.halo)

    def local_height(self, r):
        a, b = self.extended(r)
        return b - a

    def neighbours_only(self):
        """True when every halo comes from the adjacent rank (every non-empty strip has >= halo rows)."""
        return all(b - a >= self.halo for a, b in (self.owned(r) for r in range(self.G)) if b > a)

    def bands(self, r, j):
        """Row geometry of rank r at distributed level j (0 <= j < Jd), rows counted in the level's OUTPUT
        resolution: dict with the global ranges it holds (ext) and owns (own) of the L-type rows (LL/HL) and of
        the H-type rows (LH/HH), the local row where the H-type rows start, and the global heights."""
        a, b = self.owned(r)
        ea, eb = self.extended(r)
        hloc = eb - ea
        last = b == self.H
        hg, hl = ceil_div_pow2(self.H, j), ceil_div_pow2(hloc, j)
        off = ea >> (j + 1)
        return {
            "off": off, "nly_g": (hg + 1) >> 1, "nly_l": (hl + 1) >> 1,
            "extL": (off, off + ((hl + 1) >> 1)), "extH": (off, off + (hl >> 1)),
            "ownL": (a >> (j + 1), ceil_div_pow2(b, j + 1) if last else b >> (j + 1)),
            "ownH": (a >> (j + 1), (ceil_div_pow2(b, j) >> 1) if last else b >> (j + 1)),
        }

    def ll_rows(self, r):
        """(global rows of LL_Jd rank r owns, global rows it holds, local offset)."""
        a, b = self.owned(r)
        ea, eb = self.extended(r)
        last = b == self.H
        off = ea >> self.Jd
        return (a >> self.Jd, ceil_div_pow2(b, self.Jd) if last else b >> self.Jd), (off, off + ceil_div_pow2(eb - ea, self.Jd)), off


class NumpyEngine:
    """Transforms numpy arrays in place through callables (fwd2(img, j_max) -> J, inv2(img, j_max))."""

    def __init__(self, fwd2, inv2):
        self.fwd2, self.inv2 = fwd2, inv2


# ======================================================================================================
# single-process reference implementation of the scheme on numpy arrays: used by the CPU tests (with an
# oracle-backed engine) and by the single-GPU emulation test (with the device engine).  The multi-process
# driver below does the same steps with the exchanges going through torch.distributed.
# ======================================================================================================
def forward_strips_local(image, world, levels_distributed, engine, j_max=-1, halo_lines=4):
    """Returns the Mallat-layout forward transform of `image` computed strip-wise by `world` emulated ranks."""
    H, W = image.shape
    J = _full_depth(W, H) if j_max < 0 else min(j_max, _full_depth(W, H))
    Jd = min(levels_distributed, J)
    plan = StripPlan(W, H, world, Jd, halo_lines)
    out = np.empty_like(image)
    ll = np.empty((ceil_div_pow2(H, Jd), ceil_div_pow2(W, Jd)), dtype=image.dtype)
    for r in range(world):
        a, b = plan.owned(r)
        if a >= b:
            continue
        ea, eb = plan.extended(r)
        local = image[ea:eb].copy()                          # owned rows + halo rows "received" from the neighbours
        engine.fwd2(local, Jd)
        _scatter_owned(plan, r, local, out, ll)
    # rank 0: remaining levels on the gathered LL_Jd band
    if J > Jd:
        engine.fwd2(ll, J - Jd)
    out[:ll.shape[0], :ll.shape[1]] = ll
    return out, J


def inverse_strips_local(coeffs, world, levels_distributed, engine, J, halo_lines=4):
    H, W = coeffs.shape
    Jd = min(levels_distributed, J)
    plan = StripPlan(W, H, world, Jd, halo_lines)
    h_d, w_d = ceil_div_pow2(H, Jd), ceil_div_pow2(W, Jd)
    ll = np.ascontiguousarray(coeffs[:h_d, :w_d])
    if J > Jd:
        engine.inv2(ll, J - Jd)
    full = coeffs.copy()
    full[:h_d, :w_d] = ll
    out = np.empty_like(coeffs)
    for r in range(world):
        a, b = plan.owned(r)
        if a >= b:
            continue
        ea, eb = plan.extended(r)
        local = _gather_extended(plan, r, full)
        engine.inv2(local, Jd)
        out[a:b] = local[a - ea:b - ea]
    return out


def _full_depth(W, H):
    m, j = min(W, H), 0
    while (1 << j) < m:
        j += 1
    return j


def _scatter_owned(plan, r, local, out, ll):
    """Copy the subband rows rank r owns from its local Mallat image into the global one."""
    W, Jd = plan.W, plan.Jd
    a, b = plan.owned(r)
    ea, eb = plan.extended(r)
    hloc = eb - ea
    last = b == plan.H
    for j in range(Jd):
        w = ceil_div_pow2(W, j)
        nlx = (w + 1) >> 1
        # global / local geometry of level j
        hg, hl_ = ceil_div_pow2(plan.H, j), ceil_div_pow2(hloc, j)
        nly_g, nly_l = (hg + 1) >> 1, (hl_ + 1) >> 1
        lo = a >> (j + 1)
        hiL = ceil_div_pow2(b, j + 1) if last else b >> (j + 1)      # L-type output rows of level j owned
        hiH = (ceil_div_pow2(b, j) >> 1) if last else b >> (j + 1)   # H-type output rows owned
        off = ea >> (j + 1)
        out[lo:hiL, nlx:w] = local[lo - off:hiL - off, nlx:w]                                   # HL
        out[nly_g + lo:nly_g + hiH, 0:w] = local[nly_l + lo - off:nly_l + hiH - off, 0:w]       # LH | HH
    lo = a >> Jd
    hi = ceil_div_pow2(b, Jd) if last else b >> Jd
    off = ea >> Jd
    ll[lo:hi, :] = local[lo - off:hi - off, :ll.shape[1]]


def _gather_extended(plan, r, full):
    """Build rank r's local Mallat image (extended strip, Jd levels) from the global coefficient image."""
    W, Jd = plan.W, plan.Jd
    ea, eb = plan.extended(r)
    hloc = eb - ea
    last = eb == plan.H
    local = np.zeros((hloc, W), dtype=full.dtype)
    for j in range(Jd):
        w = ceil_div_pow2(W, j)
        nlx = (w + 1) >> 1
        hg, hl_ = ceil_div_pow2(plan.H, j), ceil_div_pow2(hloc, j)
        nly_g, nly_l = (hg + 1) >> 1, (hl_ + 1) >> 1
        lo = ea >> (j + 1)
        nL, nH = nly_l, hl_ >> 1
        local[0:nL, nlx:w] = full[lo:lo + nL, nlx:w]
        local[nly_l:nly_l + nH, 0:w] = full[nly_g + lo:nly_g + lo + nH, 0:w]
    h_d = ceil_div_pow2(hloc, Jd)
    w_d = ceil_div_pow2(W, Jd)
    lo = ea >> Jd
    local[0:h_d, 0:w_d] = full[lo:lo + h_d, 0:w_d]
    return local


# ======================================================================================================
# one process per GPU: torch.distributed for the exchanges (NCCL -> NVLink P2P on GPUs, gloo on CPU)
# ======================================================================================================
class DistStrips:
    """Row-strip transform of one image distributed over the ranks of a torch.distributed group.

    make_engine(width, height) returns an object with
        view()      -> 2-D torch tensor [height, width] aliasing the engine's CURRENT data (re-fetch after a transform)
        fwd2(J), inv2(J)   in-place J-level transform with the reference's semantics
    (`DeviceStripEngine` below for GPUs; the CPU tests plug an oracle-backed one)."""

    def __init__(self, width, height, levels_distributed, make_engine, dist):
        self.dist = dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.W, self.H = width, height
        self.J = _full_depth(width, height)
        self.Jd = min(levels_distributed, self.J)
        self.plan = StripPlan(width, height, self.world, self.Jd)
        if not self.plan.neighbours_only():
            raise ValueError("strips shorter than the halo: use fewer ranks or fewer distributed levels")
        self.a, self.b = self.plan.owned(self.rank)
        self.ea, self.eb = self.plan.extended(self.rank)
        self.local = make_engine(width, self.eb - self.ea)
        self.h_d, self.w_d = ceil_div_pow2(height, self.Jd), ceil_div_pow2(width, self.Jd)
        self.top = make_engine(self.w_d, self.h_d) if self.rank == 0 else None   # LL_Jd and everything above it

    # sends / receives are queued on an "exchange" = (ops, pending copies); non-contiguous views (pitched
    # planes, column ranges) are staged through dense temporaries
    def _send(self, ex, tensor, peer):
        ex[0].append(self.dist.P2POp(self.dist.isend, tensor if tensor.is_contiguous() else tensor.contiguous(), peer))

    def _recv(self, ex, tensor, peer):
        if tensor.is_contiguous():
            ex[0].append(self.dist.P2POp(self.dist.irecv, tensor, peer))
        else:
            buf = tensor.new_empty(tuple(tensor.shape))
            ex[1].append((buf, tensor))
            ex[0].append(self.dist.P2POp(self.dist.irecv, buf, peer))

    def _run(self, ex):
        if ex[0]:
            for req in self.dist.batch_isend_irecv(ex[0]):
                req.wait()
        for buf, dst in ex[1]:
            dst.copy_(buf)

    def owned_view(self):
        """The rank's owned rows of the image (before forward / after inverse)."""
        return self.local.view()[self.a - self.ea:self.b - self.ea]

    # ---- forward ------------------------------------------------------------------------------------
    def exchange_input_halo(self):
        p, r, dist = self.plan, self.rank, self.dist
        v = self.local.view()
        a, b, ea, eb = self.a, self.b, self.ea, self.eb
        ex = ([], [])
        if r > 0:
            pa, pb = p.owned(r - 1)
            peb = p.extended(r - 1)[1]
            self._recv(ex, v[0:a - ea], r - 1)                        # rows [ea, a) from above
            self._send(ex, v[a - ea:a - ea + (peb - pb)], r - 1)      # its bottom halo = my first rows
        if r < self.world - 1:
            nea = p.extended(r + 1)[0]
            self._recv(ex, v[b - ea:eb - ea], r + 1)
            self._send(ex, v[nea - ea:b - ea], r + 1)
        self._run(ex)

    def forward(self):
        """Owned rows of local.view() hold the image; afterwards local.view() holds the rank's Mallat strip and,
        on rank 0, top.view() the top of the pyramid (the Mallat image of LL_Jd)."""
        dist, p, r = self.dist, self.plan, self.rank
        self.exchange_input_halo()
        self.local.fwd2(self.Jd)
        v = self.local.view()
        own, _, off = p.ll_rows(r)
        mine = v[own[0] - off:own[1] - off, :self.w_d]
        ex = ([], [])
        if r == 0:
            t = self.top.view()
            t[own[0]:own[1]].copy_(mine)
            for s in range(1, self.world):
                so, _, _ = p.ll_rows(s)
                if so[1] > so[0]:
                    self._recv(ex, t[so[0]:so[1]], s)
        elif own[1] > own[0]:
            self._send(ex, mine, 0)
        self._run(ex)
        if r == 0 and self.J > self.Jd:
            self.top.fwd2(self.J - self.Jd)

    # ---- inverse ------------------------------------------------------------------------------------
    def inverse(self):
        dist, p, r = self.dist, self.plan, self.rank
        if r == 0 and self.J > self.Jd:
            self.top.inv2(self.J - self.Jd)
        # LL_Jd: every rank gets the rows of its extended strip
        v = self.local.view()
        _, ext, off = p.ll_rows(r)
        ex = ([], [])
        if r == 0:
            t = self.top.view()
            v[0:ext[1] - ext[0], :self.w_d].copy_(t[ext[0]:ext[1]])
            for s in range(1, self.world):
                _, se, _ = p.ll_rows(s)
                if se[1] > se[0]:
                    self._send(ex, t[se[0]:se[1]], s)
        elif ext[1] > ext[0]:
            self._recv(ex, v[0:ext[1] - ext[0], :self.w_d], 0)
        self._run(ex)
        # subband halo rows of every distributed level, from the adjacent ranks
        ex = ([], [])
        for j in range(self.Jd):
            w = ceil_div_pow2(self.W, j)
            nlx = (w + 1) >> 1
            me = p.bands(r, j)
            for nb in (r - 1, r + 1):
                if nb < 0 or nb >= self.world:
                    continue
                ot = p.bands(nb, j)
                for kind, cols, base_me in (("L", slice(nlx, w), 0), ("H", slice(0, w), me["nly_l"])):
                    # what I own and the neighbour holds -> send; what the neighbour owns and I hold -> receive
                    s0, s1 = max(me["own" + kind][0], ot["ext" + kind][0]), min(me["own" + kind][1], ot["ext" + kind][1])
                    if s1 > s0:
                        self._send(ex, v[base_me + s0 - me["off"]:base_me + s1 - me["off"], cols], nb)
                    g0, g1 = max(ot["own" + kind][0], me["ext" + kind][0]), min(ot["own" + kind][1], me["ext" + kind][1])
                    if g1 > g0:
                        self._recv(ex, v[base_me + g0 - me["off"]:base_me + g1 - me["off"], cols], nb)
        self._run(ex)
        self.local.inv2(self.Jd)

    # ---- whole-image views for tests --------------------------------------------------------------------
    def gather_mallat(self):
        """The complete Mallat-layout coefficient image on rank 0 (numpy), None elsewhere.  Test helper."""
        import numpy as _np
        objs = [None] * self.world
        v = self.local.view().cpu().numpy()
        self.dist.all_gather_object(objs, v)
        if self.rank != 0:
            return None
        out = _np.empty((self.H, self.W), dtype=v.dtype)
        ll = _np.empty((self.h_d, self.w_d), dtype=v.dtype)
        for s in range(self.world):
            if self.plan.owned(s)[1] > self.plan.owned(s)[0]:
                _scatter_owned(self.plan, s, objs[s], out, ll)
        t = self.top.view().cpu().numpy()
        out[:self.h_d, :self.w_d] = t
        return out


class DeviceStripEngine:
    """A strip living in a dwtb200_image; view() aliases the image's current plane as a torch CUDA tensor."""

    def __init__(self, dwt, torch, kind, width, height):
        self.dwt, self.torch = dwt, torch
        self.img = dwt.DeviceImage(kind, width, height)
        self.width, self.height = width, height
        self.dtype = {dwt.CDF97_F32: torch.float32, dwt.CDF97_F64: torch.float64, dwt.CDF53_I32: torch.int32}[kind]
        self.typestr = {dwt.CDF97_F32: "<f4", dwt.CDF97_F64: "<f8", dwt.CDF53_I32: "<i4"}[kind]

    def view(self):
        ptr, pitch, _ = self.img.devptr()
        self.img.L.check(self.img.L.c.dwtb200_sync())

        class _Alias:
            pass
        al = _Alias()
        es = self.torch.empty((), dtype=self.dtype).element_size()
        al.__cuda_array_interface__ = {"shape": (self.height, pitch // es), "typestr": self.typestr, "data": (ptr, False), "version": 3}
        return self.torch.as_tensor(al, device="cuda")[:, :self.width]

    def fwd2(self, J):
        self.torch.cuda.synchronize()
        assert self.img.fwd2(J) == J
        self.img.L.check(self.img.L.c.dwtb200_sync())

    def inv2(self, J):
        self.torch.cuda.synchronize()
        self.img.inv2(J)
        self.img.L.check(self.img.L.c.dwtb200_sync())
