"""Fluid model of the pipelined host path (forward): one H2D engine (pieces in order), one D2H engine (ready items, FIFO by issue),
rates depend on whether the other direction is busy.  Units: image rows of full width; H = 8192."""
import sys
H = 8192.0
MB_PER_ROW = 8192 * 4 / 1e6
UP_ALONE, UP_BOTH, DN_ALONE, DN_BOTH = 54.0, 43.5, 56.0, 42.5   # GB/s = MB/ms

def simulate(pieces, downloads, final_after=0.15):
    """pieces: list of (rows) uploaded in order, each a dict(name, size_rows, set of row-ids covered as (r0, r1));
    downloads: list of dict(size_rows, need_pieces: index of last piece needed (inputs + landing), extra_delay)"""
    t = 0.0
    dt = 0.002
    up_i, up_left = 0, pieces[0]['size'] * MB_PER_ROW
    up_done_t = [None] * len(pieces)
    dn_left = [d['size'] * MB_PER_ROW for d in downloads]
    dn_done = [False] * len(downloads)
    cur = None
    busy_dn = 0.0
    while True:
        up_busy = up_i < len(pieces)
        # choose download
        if cur is None:
            for i, d in enumerate(downloads):
                if dn_done[i]:
                    continue
                k = d['need']
                ready = up_done_t[k] is not None and t >= up_done_t[k] + d.get('delay', 0.03)
                if ready:
                    cur = i
                    break
        dn_busy = cur is not None
        if not up_busy and not dn_busy and all(dn_done):
            break
        if up_busy:
            rate = UP_BOTH if dn_busy else UP_ALONE
            up_left -= rate * dt
            if up_left <= 0:
                up_done_t[up_i] = t
                up_i += 1
                if up_i < len(pieces):
                    up_left = pieces[up_i]['size'] * MB_PER_ROW
        if dn_busy:
            rate = DN_BOTH if up_busy else DN_ALONE
            dn_left[cur] -= rate * dt
            busy_dn += dt
            if dn_left[cur] <= 0:
                dn_done[cur] = True
                cur = None
        t += dt
        if t > 50:
            raise RuntimeError('stuck')
    return t, up_done_t[-1], busy_dn

def piece_index_covering(pieces, row):
    for i, p in enumerate(pieces):
        if p['r0'] <= row < p['r1']:
            return i
    raise KeyError(row)

def build(order, levels, nch=16):
    """order: 'seq' | 'geo'; levels: how many levels are pipelined (1 or 2 or 3)"""
    # upload plan: list of row ranges
    ranges = []
    if order == 'seq':
        s = H / nch
        for c in range(nch):
            ranges.append((c * s, (c + 1) * s))
    else:
        steps = nch // 2
        runs = []
        lo, ln = 0.0, H / 2
        while ln >= 64 * steps / steps and len(runs) < (6 if order == 'geo' else 2):
            runs.append((lo, ln))
            lo += ln
            ln /= 2
        # the last run takes everything left
        runs[-1] = (runs[-1][0], H - runs[-1][0])
        for i in range(steps):
            for (lo, ln) in runs:
                ranges.append((lo + ln * i / steps, lo + ln * (i + 1) / steps))
    pieces = [dict(r0=a, r1=b, size=b - a) for a, b in ranges]
    downloads = []
    # outputs of the level-0 run on each piece (issued in piece order)
    def add(size_rows, inputs_last_piece, dest_r0, dest_r1, cls):
        k = max(inputs_last_piece, piece_index_covering(pieces, dest_r0), piece_index_covering(pieces, max(dest_r0, dest_r1 - 1e-6)))
        downloads.append(dict(size=size_rows, need=k, cls=cls))
    for i, (a, b) in enumerate(ranges):
        n = b - a
        # level 0: HL rows [a/2, b/2) half width; LH|HH rows H/2 + [a/2, b/2) full width
        add(n / 2 * 0.5, i, a / 2, b / 2, 'HL0')
        add(n / 2 * 1.0, i, H / 2 + a / 2, H / 2 + b / 2, 'LHHH0')
        if levels >= 1:   # level 1 on the LL rows [a/2, b/2): HL1 rows [a/4, b/4) quarter width; LH1|HH1 rows H/4 + [a/4, b/4) half width
            add(n / 4 * 0.25, i, a / 4, b / 4, 'HL1')
            add(n / 4 * 0.5, i, H / 4 + a / 4, H / 4 + b / 4, 'LHHH1')
        if levels >= 2:
            add(n / 8 * 0.125, i, a / 8, b / 8, 'HL2')
            add(n / 8 * 0.25, i, H / 8 + a / 8, H / 8 + b / 8, 'LHHH2')
    # the rest of the pyramid: the LL quadrant of the deepest pipelined level, after everything
    frac = 0.25 ** (levels + 1) if levels >= 0 else 0.25
    downloads.append(dict(size=H * (0.25 ** (levels + 1)) , need=len(pieces) - 1, delay=0.2, cls='LL'))
    return pieces, downloads

for order in ('seq', '2:1', 'geo'):
    for levels in (0, 1, 2):
        p, d = build(order, levels)
        tot = sum(x['size'] for x in d)
        t, tup, busy = simulate(p, d)
        print(f"{order:4s} levels pipelined {levels}: total {t:.2f} ms  upload done {tup:.2f} ms  (download rows {tot:.0f} of {H:.0f})")
