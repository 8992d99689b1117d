"""A/B parity of tile configurations: every config must give byte-identical nodes / pool / labels on real frames and on the
synthetic edge-case planes (walls, all-wall, flat, checker, noise ...).  Usage: python tools/ab_tile_parity.py 0 5 6"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scene-text-recognition_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import ertext
from conftest import make_plane
cfgs = [int(a) for a in sys.argv[1:]] or [0, 5]
g = np.load(os.path.join(ROOT, "tests", "golden", "frames.npz"))["frames"]
e = ertext.ErText()
def sig(r):
    return [(p.nodes.tobytes(), p.pool.tobytes(), p.label.tobytes()) for p in r.planes], r.status
cases = [("golden", lambda: e.detect_classify(g))]
for kind in ("noise", "smooth", "blobs", "walls", "wall0", "wall01", "allwall", "flat", "checker", "ramp"):
    for (h, w) in ((37, 53), (64, 96), (90, 70), (200, 333)):
        for ma in (0, 120):
            cases.append(("%s %dx%d ma%d" % (kind, w, h, ma), (lambda kind=kind, h=h, w=w, ma=ma: (e.set_min_area(ma), e.planes_detect(np.stack([make_plane(1, h, w, kind), make_plane(2, h, w, kind)])))[1])))
bad = 0
for name, fn in cases:
    ref = None
    for c in cfgs:
        e.set_tile_config(c)
        s = sig(fn())
        if ref is None:
            ref = s
        elif s != ref:
            bad += 1
            print("MISMATCH", name, "cfg", c, "status", s[1], ref[1], flush=True)
e.set_min_area(120)
print("cases", len(cases), "configs", cfgs, "mismatches", bad)
