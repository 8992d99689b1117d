"""Serial Python model of the WHOLE tiled component-tree pipeline of the CUDA path (design validation only):

  per tile   k_tile_build : keyed union-find on the in-tile edges, own-level pixel count / bbox per tile-local node,
                            the seam-aware BORDER rule (phase D2), interior subtrees folded on chip (D3), emission of
                            BORDER nodes + kept interior nodes (E), one seam record per tile-side position
  global     k_seam_link_rec (keyed union of the records across every seam), k_fold (aliases hand their totals to the
                            final node of their level, final nodes resolve their final parent), k_refit (bottom-up
                            totals), k_emit_kept (the flood's start rule, area > MIN_AREA or root)

tests/test_tiled_model_cpu.py compares its node set with the oracle's canonical node set on seeded planes (walls, noise,
flat, checker ...) for several tile shapes -- an interpreter-speed proof of the decomposition, in particular of the rule
"a side pixel p with outside neighbour q only drags in A(p), the highest ancestor-or-self of p's node with level <=
max(level p, level q), and A(p) stands for p in the seam record".
`seam_aware=False` gives the older rule (p's own node and all its ancestors) for comparison; `stats` returns how many
tile-local nodes went global under either rule."""
import numpy as np

INF = 0xFFFFFFFFFFFF
SH = 32                      # global key = level << 32 | pixel index (the kernels use 26 bits; the model is not size-limited)
MASK = (1 << SH) - 1
WALL = 255


def _find(par, k):
    while True:
        p = par.get(k & MASK, INF)
        if p == INF or (p >> SH) != (k >> SH):
            return k
        k = p


def _link(par, x, y):
    """insert the edge x--y into the keyed forest (same procedure as link_s / link_g of er_extract.cu)"""
    while True:
        x = _find(par, x); y = _find(par, y)
        if x == y:
            return
        if x > y:
            x, y = y, x
        old = par.get(x & MASK, INF)
        if y < old:
            par[x & MASK] = y
        if old == y or old == INF:
            return
        if old < y:
            x = old
        else:
            x, y = y, old


def _supertile_merge(lv, W, H, TH, TW, SY, SX, min_area, gpar, attr, recs, stats):
    """Round-2 design, validated here first: unite the BORDER graphs of SY x SX tiles before the global kernels see them.
    Per super-tile: the keyed union on its INTERIOR seams, aliases folded into their final node, the seam-aware BORDER rule
    re-evaluated against the super-tile's OUTER ring only (in the merged forest), everything that became interior folded
    bottom-up into its parent and dropped (or kept as a complete node if the reference keeps it)."""
    SH_, SW_ = SY * TH, SX * TW
    level_of = lambda k: k >> SH
    for Y0 in range(0, H, SH_):
        for X0 in range(0, W, SW_):
            Y1, X1 = min(H, Y0 + SH_), min(W, X0 + SW_)
            inside = lambda k: Y0 <= (k & MASK) // W < Y1 and X0 <= (k & MASK) % W < X1
            # 1. interior seams
            for (y, x, side), ra in list(recs.items()):
                if not (Y0 <= y < Y1 and X0 <= x < X1) or ra is None:
                    continue
                if side == "R" and x + 1 < X1:
                    rb = recs.get((y, x + 1, "L"))
                elif side == "B" and y + 1 < Y1:
                    rb = recs.get((y + 1, x, "T"))
                else:
                    continue
                if rb is not None:
                    _link(gpar, ra, rb)
            # 2. aliases hand their totals to the final node of their level and disappear
            members = [g for g in attr if inside(g) and not attr[g][6]]
            for g in members:
                f = _find(gpar, g)
                if f != g:
                    a, b = attr.pop(g), attr[f]
                    b[0] += a[0]; b[1] += a[1] - 1
                    b[2] = min(b[2], a[2]); b[3] = min(b[3], a[3]); b[4] = max(b[4], a[4]); b[5] = max(b[5], a[5])
            members = [g for g in members if g in attr]
            parent = {}
            for g in members:
                pk = gpar.get(g & MASK, INF)
                parent[g] = None if pk == INF else _find(gpar, pk)
                if parent[g] is not None:
                    gpar[g & MASK] = parent[g]
            # 3. BORDER against the outer ring, in the merged forest
            border = set()

            def mark(r):
                while r is not None and r not in border:
                    border.add(r); r = parent[r]

            for (y, x, side), r in list(recs.items()):
                if not (Y0 <= y < Y1 and X0 <= x < X1) or r is None:
                    continue
                dy, dx = {"T": (-1, 0), "B": (1, 0), "L": (0, -1), "R": (0, 1)}[side]
                qy, qx = y + dy, x + dx
                if Y0 <= qy < Y1 and X0 <= qx < X1:
                    recs[(y, x, side)] = None                    # interior seam: done
                    continue
                M = max(int(lv[y, x]), int(lv[qy, qx]))
                a = _find(gpar, r)
                while parent[a] is not None and level_of(parent[a]) <= M:
                    a = parent[a]
                recs[(y, x, side)] = a
                mark(a)
            for (gy, gx) in ((0, 0), (0, 1), (1, 0)):
                if Y0 <= gy < Y1 and X0 <= gx < X1 and gx < W and gy < H and lv[gy, gx] != WALL:
                    mark(_find(gpar, _find(gpar, (int(lv[gy, gx]) << SH) | (gy * W + gx))))
            # 4. what became interior folds into its parent, bottom-up, and leaves the global forest unless the reference keeps it
            for g in sorted(members):
                if g in border:
                    continue
                a = attr[g]
                if parent[g] is not None:
                    b = attr[parent[g]]
                    b[0] += a[0]; b[1] += a[1]
                    b[2] = min(b[2], a[2]); b[3] = min(b[3], a[3]); b[4] = max(b[4], a[4]); b[5] = max(b[5], a[5])
                if a[0] + a[1] > min_area:
                    a[6] = True
                else:
                    del attr[g]
    if stats is not None:
        stats["global_nodes_after_supertile"] = sum(1 for a in attr.values() if not a[6])


def tiled_nodes(levels, min_area, TH=32, TW=64, hi=32, seam_aware=True, stats=None, supertile=None):
    """levels: int array [H, W] of quantised levels (>= hi = wall).  Returns the kept nodes as tuples
    (level, area, x, y, w, h), the same columns as the oracle's canonical dump."""
    H, W = levels.shape
    lv = np.where(levels >= hi, WALL, levels).astype(np.int64)
    gkey = lambda L, y, x: (int(L) << SH) | (y * W + x)
    gpar = {}                 # global forest: pixel index -> key of an ancestor
    attr = {}                 # global node (by key) -> [cnt, nn, x0, y0, x1, y1, complete]
    recs = {}                 # (y, x, side) -> key of the node that stands for that side position, or None
    n_local = n_global = 0

    for Y0 in range(0, H, TH):
        for X0 in range(0, W, TW):
            rows, cols = min(TH, H - Y0), min(TW, W - X0)
            t = lv[Y0:Y0 + rows, X0:X0 + cols]
            # ---- phases A/B: in-tile keyed union-find (tile-local keys are global keys here: simpler, same forest) ----
            par = {}
            for y in range(rows):
                for x in range(cols):
                    if t[y, x] == WALL:
                        continue
                    if x + 1 < cols and t[y, x + 1] != WALL:
                        _link(par, gkey(t[y, x], Y0 + y, X0 + x), gkey(t[y, x + 1], Y0 + y, X0 + x + 1))
                    if y + 1 < rows and t[y + 1, x] != WALL:
                        _link(par, gkey(t[y, x], Y0 + y, X0 + x), gkey(t[y + 1, x], Y0 + y + 1, X0 + x))
            root_of = lambda y, x: _find(par, gkey(t[y, x], Y0 + y, X0 + x))
            # ---- phase D: own-level pixels and bbox per tile-local node; phase C: parent's level root ----
            loc = {}
            for y in range(rows):
                for x in range(cols):
                    if t[y, x] == WALL:
                        continue
                    a = loc.setdefault(root_of(y, x), [0, 1, X0 + x, Y0 + y, X0 + x, Y0 + y])
                    a[0] += 1
                    a[2] = min(a[2], X0 + x); a[3] = min(a[3], Y0 + y); a[4] = max(a[4], X0 + x); a[5] = max(a[5], Y0 + y)
            up = {}
            for r in loc:
                p = par.get(r & MASK, INF)
                up[r] = None if p == INF else _find(par, p)
            n_local += len(loc)
            # ---- phase D2: BORDER ----
            border = set()

            def mark(r):
                while r is not None and r not in border:
                    border.add(r); r = up[r]

            sides = []
            if Y0 > 0:
                sides += [(0, x, -1, 0, "T") for x in range(cols)]
            if Y0 + rows < H:
                sides += [(rows - 1, x, 1, 0, "B") for x in range(cols)]
            if X0 > 0:
                sides += [(y, 0, 0, -1, "L") for y in range(rows)]
            if X0 + cols < W:
                sides += [(y, cols - 1, 0, 1, "R") for y in range(rows)]
            for (y, x, dy, dx, side) in sides:
                recs[(Y0 + y, X0 + x, side)] = None
                if t[y, x] == WALL:
                    continue
                Lq = lv[Y0 + y + dy, X0 + x + dx]
                r = root_of(y, x)
                if seam_aware:
                    if Lq == WALL:
                        continue
                    M = max(int(t[y, x]), int(Lq))
                    while up[r] is not None and (up[r] >> SH) <= M:
                        r = up[r]
                recs[(Y0 + y, X0 + x, side)] = r
                mark(r)
            for (gy, gx) in ((0, 0), (0, 1), (1, 0)):        # the flood's start candidates: pixels 0, 1, W
                y, x = gy - Y0, gx - X0
                if 0 <= y < rows and 0 <= x < cols and gx < W and gy < H and t[y, x] != WALL:
                    mark(root_of(y, x))
            # ---- phase D3: fold interior subtrees bottom-up ----
            for r in sorted(loc):                                # ascending key = ascending level
                if r in border or up[r] is None:
                    continue
                a, b = loc[up[r]], loc[r]
                a[0] += b[0]; a[1] += b[1]
                a[2] = min(a[2], b[2]); a[3] = min(a[3], b[3]); a[4] = max(a[4], b[4]); a[5] = max(a[5], b[5])
            # ---- phase E: emit ----
            for r, a in loc.items():
                is_border = r in border
                if not (is_border or a[0] + a[1] > min_area):
                    continue
                attr[r] = a + [not is_border]
                gpar[r & MASK] = INF if up[r] is None else up[r]
                n_global += 1
            # start candidates that are not level roots publish their root (looked up by pixel in k_emit_kept)
            for (gy, gx) in ((0, 0), (0, 1), (1, 0)):
                y, x = gy - Y0, gx - X0
                if 0 <= y < rows and 0 <= x < cols and gx < W and gy < H and t[y, x] != WALL:
                    r = root_of(y, x)
                    if (r & MASK) != gy * W + gx:
                        gpar[gy * W + gx] = r
    if stats is not None:
        stats["local_nodes"] = n_local; stats["global_nodes"] = n_global
        stats["global_border_nodes"] = sum(1 for a in attr.values() if not a[6])
    if supertile is not None:
        assert seam_aware
        _supertile_merge(lv, W, H, TH, TW, supertile[0], supertile[1], min_area, gpar, attr, recs, stats)

    # ---- k_seam_link_rec: unite the records across every seam ----
    for (y, x, side), ra in list(recs.items()):
        if side == "R":
            rb = recs.get((y, x + 1, "L"))
        elif side == "B":
            rb = recs.get((y + 1, x, "T"))
        else:
            continue
        if ra is not None and rb is not None:
            _link(gpar, ra, rb)

    # ---- k_fold ----
    final = {}
    for g, a in attr.items():
        if a[6]:
            continue                                             # complete interior node: totals final, only its parent may move
        f = _find(gpar, g)
        if f != g:                                               # alias: merged into a same-level node of another tile
            b = attr[f]
            b[0] += a[0]; b[1] += a[1] - 1
            b[2] = min(b[2], a[2]); b[3] = min(b[3], a[3]); b[4] = max(b[4], a[4]); b[5] = max(b[5], a[5])
            a[1] = 0
    parent = {}
    for g, a in attr.items():
        if a[1] == 0:
            continue
        pk = gpar.get(g & MASK, INF)
        parent[g] = None if pk == INF else _find(gpar, pk)
        final[g] = a
    # ---- k_refit: bottom-up totals.  Complete (interior, kept) nodes were already folded into their parent on chip and
    # never start or continue a chain (pend == NODE_COMPLETE in the kernel) ----
    for g in sorted(final):
        q = parent[g]
        if q is None or final[g][6]:
            continue
        assert q > g and (q >> SH) > (g >> SH), "a parent must sit on a strictly higher level"
        a, b = final[q], final[g]
        a[0] += b[0]; a[1] += b[1]
        a[2] = min(a[2], b[2]); a[3] = min(a[3], b[3]); a[4] = max(a[4], b[4]); a[5] = max(a[5], b[5])
    # ---- k_emit_kept: the flood's start rule (src/ER.cpp:267-341), area > MIN_AREA or the root ----
    start = None
    for s in (0, 1, W):
        if s < H * W and lv.ravel()[s] != WALL:
            start = s
            break
    if start is None:
        return [(int(levels.ravel()[0]), 2, 0, 0, 1, 1)]
    tkey = _find(gpar, (int(lv.ravel()[start]) << SH) | start)
    while parent[tkey] is not None:
        tkey = parent[tkey]
    out = []
    for g, a in final.items():
        q = g
        while parent[q] is not None:
            q = parent[q]
        if q != tkey:
            continue
        area = a[0] + a[1]
        if area > min_area or g == tkey:
            out.append((g >> SH, area, a[2], a[3], a[4] - a[2] + 1, a[5] - a[3] + 1))
    return out
