"""A/B of the tile-build kernels (ert_set_tile_config): outputs must be byte-identical between configurations on a
ladder of planes (ert_set_tile_config 0..3 = variants of k_tile_build2; the default is oracle-validated by the test-suite), then the bench workload
(8 synthetic 1080p frames = 48 planes) is timed per configuration with the library's CUDA events around the tile
kernel, plus the per-phase cycle sums of ert_debug_phase_cycles.   python tools/tile_ab.py [--cfgs 0,1] [--quick]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scene-text-recognition_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ertext  # noqa: E402
from conftest import make_plane  # noqa: E402

KINDS = ["noise", "smooth", "blobs", "walls", "wall0", "wall01", "allwall", "flat", "checker", "ramp"]
SIZES = [(1, 1), (1, 9), (7, 1), (31, 33), (32, 64), (33, 65), (64, 128), (100, 130), (257, 191), (480, 640)]


def same(a, b):
    if a.status != b.status or len(a.planes) != len(b.planes):
        return False
    for p, q in zip(a.planes, b.planes):
        if p.nodes.tobytes() != q.nodes.tobytes() or p.pool.tobytes() != q.pool.tobytes() or p.label.tobytes() != q.label.tobytes():
            return False
    return True


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfgs", default="0,1,2,3")
    ap.add_argument("--base", type=int, default=0)
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--iters", type=int, default=12)
    a = ap.parse_args()
    cfgs = [int(x) for x in a.cfgs.split(",")]
    e = ertext.ErText(device=0)
    bad = 0
    for kind in KINDS:
        for si, (h, w) in enumerate(SIZES):
            img = make_plane(si, h, w, kind)
            for ma in (3, 120):
                e.set_min_area(ma)
                e.set_tile_config(a.base)
                ref = e.planes_detect(img, upto=ertext.STAGE_NMS)
                if ref.status:      # kept-capacity overflow (tiny MIN_AREA on a large noisy plane): the kept SET is then arbitrary
                    continue
                for c in cfgs:
                    if c == a.base:
                        continue
                    e.set_tile_config(c)
                    got = e.planes_detect(img, upto=ertext.STAGE_NMS)
                    if not same(ref, got):
                        bad += 1
                        if bad <= 12:
                            rn, gn = ref.planes[0].nodes, got.planes[0].nodes
                            print("MISMATCH cfg %d kind %s %dx%d min_area %d: status %d/%d nodes %d/%d pool %d/%d" % (
                                c, kind, h, w, ma, ref.status, got.status, len(rn), len(gn), len(ref.planes[0].pool), len(got.planes[0].pool)))
                            rs, gs = set(map(tuple, rn[:, :6])), set(map(tuple, gn[:, :6]))
                            print("   only-base:", sorted(rs - gs)[:5], " only-new:", sorted(gs - rs)[:5])
    e.set_min_area(120)
    print("ladder mismatches:", bad)
    if not a.quick:
        for kind in ("blobs", "noise"):
            img = make_plane(42, 1080, 1920, kind)
            e.set_tile_config(a.base)
            ref = e.planes_detect(img, upto=ertext.STAGE_NMS)
            for c in cfgs:
                if c != a.base:
                    e.set_tile_config(c)
                    ok = same(ref, e.planes_detect(img, upto=ertext.STAGE_NMS))
                    print("1080p %s cfg %d identical: %s" % (kind, c, ok))
                    bad += 0 if ok else 1
    # timing on the bench workload
    from ertext import synth
    frames = synth.s_text_batch(1234, 8, 1920, 1080)
    out = {}
    res0 = None
    for c in cfgs:
        e.set_tile_config(c)
        e.phase_cycles(True)
        tile, total, ext = [], [], []
        for i in range(a.iters):
            r = e.detect_classify(frames)
            if i >= 2:
                tile.append(r.stage_ms[6]); total.append(r.stage_ms[5] - r.stage_ms[3]); ext.append(r.stage_ms[0])
        cyc = e.phase_cycles(True)
        if res0 is None:
            res0 = r
        else:
            ok = same(res0, r)
            print("bench frames cfg %d identical to cfg %d: %s" % (c, cfgs[0], ok))
            bad += 0 if ok else 1
        e.phase_cycles(False)
        s = float(sum(cyc)) or 1.0
        out[c] = {"tile_ms": float(np.median(tile)), "extract_ms": float(np.median(ext)), "device_ms_without_h2d": float(np.median(total)),
                  "phase_share_pct": [round(100.0 * v / s, 1) for v in cyc[:10]]}
        print("cfg %d: tile %.3f ms  extract %.3f ms  batch (no h2d) %.3f ms  phases%% %s" % (
            c, out[c]["tile_ms"], out[c]["extract_ms"], out[c]["device_ms_without_h2d"], out[c]["phase_share_pct"]))
    print(json.dumps(out))
    e.close()
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
