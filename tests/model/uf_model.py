"""Serial Python model of the lock-free keyed union-find used by the CUDA extract kernels
(design validation only; compared against the oracle's canonical node set in tests).

key(p) = level(p) << 26 | pixel index.  par[p] holds the key of SOME ancestor of p (or INF);
link() inserts an edge by an atomicMin-style update and re-links whatever it displaced, so the
final forest is independent of the order in which edges are processed.
"""
import numpy as np

INF = 0xFFFFFFFF
SH = 26
MASK = (1 << SH) - 1


def build(levels, hi=32, rng=None):
    h, w = levels.shape
    lv = levels.ravel().astype(np.int64)
    n = h * w
    par = [INF] * n

    def key(p):
        return (int(lv[p]) << SH) | p

    def find(k):
        while True:
            p = par[k & MASK]
            if p == INF or (p >> SH) != (k >> SH):
                return k
            k = p

    def link(x, y):
        while True:
            x = find(x); y = find(y)
            if x == y:
                return
            if x > y:
                x, y = y, x
            old = par[x & MASK]
            if y < old:
                par[x & MASK] = y
            if old == y or old == INF:
                return
            if old < y:
                x = old
            else:
                x, y = y, old

    edges = []
    for y in range(h):
        for x in range(w):
            p = y * w + x
            if lv[p] >= hi:
                continue
            if x + 1 < w and lv[p + 1] < hi:
                edges.append((p, p + 1))
            if y + 1 < h and lv[p + w] < hi:
                edges.append((p, p + w))
    if rng is not None:
        rng.shuffle(edges)
    for a, b in edges:
        link(key(a), key(b))
    return par, key, find


def nodes_from_forest(levels, min_area, hi=32, rng=None):
    h, w = levels.shape
    n = h * w
    lv = levels.ravel()
    par, key, find = build(levels, hi, rng)
    # direct attributes per level root
    cnt = {}; bb = {}
    for p in range(n):
        if lv[p] >= hi:
            continue
        r = find(key(p))
        y, x = divmod(p, w)
        cnt[r] = cnt.get(r, 0) + 1
        b = bb.get(r)
        bb[r] = (x, y, x, y) if b is None else (min(b[0], x), min(b[1], y), max(b[2], x), max(b[3], y))
    parent = {}
    for r in cnt:
        pr = par[r & MASK]
        parent[r] = None if pr == INF else find(pr)
    # refit bottom-up in key order (children have smaller keys than parents)
    nn = {r: 1 for r in cnt}
    for r in sorted(cnt):
        q = parent[r]
        if q is None:
            continue
        assert q > r and (q >> SH) > (r >> SH)
        cnt[q] += cnt[r]; nn[q] += nn[r]
        a, b = bb[q], bb[r]
        bb[q] = (min(a[0], b[0]), min(a[1], b[1]), max(a[2], b[2]), max(a[3], b[3]))
    # reach tree
    start = None
    for s in (0, 1, w):
        if s < n and lv[s] < hi:
            start = s; break
    if start is None:
        return [(int(lv[0]), 2, 0, 0, 1, 1)]
    t = find(key(start))
    while parent[t] is not None:
        t = parent[t]
    out = []
    for r in cnt:
        q = r
        while parent[q] is not None:
            q = parent[q]
        if q != t:
            continue
        area = cnt[r] + nn[r]
        if area > min_area or r == t:
            b = bb[r]
            out.append((r >> SH, area, b[0], b[1], b[2] - b[0] + 1, b[3] - b[1] + 1))
    return out
