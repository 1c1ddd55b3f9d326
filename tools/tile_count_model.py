"""Host model of the word-count COVER tile kernel (cover_tiled.cuh: k_cover_tile_wc), used to validate the algorithm against a
plain sweep before it was written in CUDA, and to size its tables (how many 32-position words need exact resolution).

Per tile of W positions (W / 32 words): count starts / stops' per word; the depth at every word boundary is exact (prefix sum);
a word whose depth bounds [d0 - nE, d0 + nS] lie inside [min, max] holds no transition (IN), one whose bounds lie on one side
outside holds none either (OUT); the others are resolved position by position.  AccIndex (max depth of a run) is exact because a
word is also resolved when its upper bound d0 + nS exceeds the largest boundary depth of its segment of consecutive IN words.

usage: python tools/tile_count_model.py [n_regions genome_len depth_min depth_max seed]
"""
import sys

import numpy as np


def sweep(starts, stops, mn, mx):
    """Reference: maximal runs with mn <= depth <= mx and their max depth, from the event list (coincident events summed)."""
    pos = np.concatenate([starts, stops])
    dl = np.concatenate([np.ones(len(starts), np.int64), -np.ones(len(stops), np.int64)])
    o = np.argsort(pos, kind="stable")
    pos, dl = pos[o], dl[o]
    up, idx = np.unique(pos, return_index=True)
    net = np.add.reduceat(dl, idx)
    depth = np.cumsum(net)
    runs = []
    cur = None
    for p, d in zip(up, depth):
        inr = mn <= d <= mx
        if inr:
            if cur is None:
                cur = [int(p), None, int(d)]
            else:
                cur[2] = max(cur[2], int(d))
        elif cur is not None:
            cur[1] = int(p)
            runs.append(tuple(cur))
            cur = None
    assert cur is None
    return runs


def tile_model(starts, stops, mn, mx, logw=16):
    """Runs over the whole coordinate range through the per-tile word-count algorithm; returns (runs, stats)."""
    W = 1 << logw
    nW = W // 32
    n_tiles = int(max(stops.max(), starts.max()) // W) + 2
    s_tile = starts // W
    e_tile = stops // W
    pieces = []          # (start, end, acc, cont0, reach_end)
    stats = dict(noisy_thr=0, noisy_max=0, words=0, tiles=0, max_noisy_per_tile=0, max_segments=0)
    depth_in = 0
    order_s = np.argsort(s_tile, kind="stable")
    order_e = np.argsort(e_tile, kind="stable")
    s_sorted, e_sorted = s_tile[order_s], e_tile[order_e]
    for t in range(n_tiles):
        lo = t * W
        a0, a1 = np.searchsorted(s_sorted, t), np.searchsorted(s_sorted, t + 1)
        b0, b1 = np.searchsorted(e_sorted, t), np.searchsorted(e_sorted, t + 1)
        qs = starts[order_s[a0:a1]] - lo
        qe = stops[order_e[b0:b1]] - lo
        if len(qs) == 0 and len(qe) == 0 and depth_in == 0:
            continue
        stats["tiles"] += 1
        nS = np.bincount(qs >> 5, minlength=nW)
        nE = np.bincount(qe >> 5, minlength=nW)
        net = nS - nE
        dend = depth_in + np.cumsum(net)
        d0 = dend - net
        lo_w, hi_w = d0 - nE, d0 + nS
        IN = (lo_w >= mn) & (hi_w <= mx)
        OUT = (hi_w < mn) | (lo_w > mx)
        thr = ~IN & ~OUT
        # segments of consecutive IN words and the largest boundary depth of each
        segstart = IN & ~np.concatenate([[False], IN[:-1]])
        seg = np.cumsum(segstart) - 1
        nseg = int(segstart.sum())
        segmax = np.full(max(nseg, 1), -1, np.int64)
        if nseg:
            np.maximum.at(segmax, seg[IN], np.maximum(d0[IN], dend[IN]))
        mxn = IN & (hi_w > np.where(IN, segmax[np.maximum(seg, 0)], 0))
        noisy = thr | mxn
        stats["noisy_thr"] += int(thr.sum()); stats["noisy_max"] += int(mxn.sum()); stats["words"] += nW
        stats["max_noisy_per_tile"] = max(stats["max_noisy_per_tile"], int(noisy.sum()))
        stats["max_segments"] = max(stats["max_segments"], nseg)
        # exact resolution of the noisy words: net per position
        npos = {}
        for w in np.nonzero(noisy)[0]:
            npos[int(w)] = np.zeros(32, np.int64)
        for q in qs:
            w = int(q) >> 5
            if w in npos:
                npos[w][int(q) & 31] += 1
        for q in qe:
            w = int(q) >> 5
            if w in npos:
                npos[w][int(q) & 31] -= 1
        # walk the words in order (the kernel does this with scans; the model keeps it sequential)
        cont0 = mn <= depth_in <= mx
        cur = None
        if cont0:
            cur = [lo, None, depth_in, True, False]
        seg_done = set()
        for w in range(nW):
            if noisy[w]:
                d = int(d0[w])
                for p in range(32):
                    if npos[w][p] == 0:
                        continue
                    d += int(npos[w][p])
                    inr = mn <= d <= mx
                    x = lo + 32 * w + p
                    if inr:
                        if cur is None:
                            cur = [x, None, d, False, False]
                        else:
                            cur[2] = max(cur[2], d)
                    elif cur is not None:
                        cur[1] = x
                        pieces.append(tuple(cur)); cur = None
                assert d == dend[w]
            elif IN[w]:
                assert cur is not None
                if seg[w] not in seg_done:
                    seg_done.add(int(seg[w]))
                    cur[2] = max(cur[2], int(segmax[seg[w]]))
            else:
                assert cur is None
        if cur is not None:
            cur[1] = lo + W
            cur[4] = True
            pieces.append(tuple(cur))
        depth_in = int(dend[-1])
    # stitch pieces that touch at tile edges
    runs = []
    for (s, e, acc, c0, re) in pieces:
        if runs and c0 and runs[-1][3] and runs[-1][1] == s:
            runs[-1] = (runs[-1][0], e, max(runs[-1][2], acc), re)
        else:
            runs.append((s, e, acc, re))
    return [(s, e, a) for (s, e, a, _r) in runs], stats


def main():
    a = sys.argv[1:]
    n = int(a[0]) if len(a) > 0 else 200000
    glen = int(a[1]) if len(a) > 1 else 6_000_000
    mn = int(a[2]) if len(a) > 2 else 2
    mx = int(a[3]) if len(a) > 3 else 2 ** 31 - 1
    seed = int(a[4]) if len(a) > 4 else 1
    rng = np.random.Generator(np.random.PCG64(seed))
    ln = np.clip(np.rint(rng.lognormal(np.log(300.0), 0.6, n)), 50, 5000).astype(np.int64)
    st = (rng.random(n) * (glen - ln)).astype(np.int64)
    en = st + ln
    en = en + (en % 2000 == 0)
    want = sweep(st, en, mn, mx)
    got, stats = tile_model(st, en, mn, mx)
    print("runs", len(want), len(got), "equal", want == got)
    print(stats, "noisy per tile avg", (stats["noisy_thr"] + stats["noisy_max"]) / max(1, stats["tiles"]))


if __name__ == "__main__":
    main()
