"""First-contact GPU diagnostics: run the CUDA path against the oracle on a ladder of cases and
print where (if anywhere) they diverge.  Not a test -- a debugging aid for `gpurun` sessions
(the pytest -m gpu suite is the gate).  Usage: python tools/gpu_check.py [--quick]
"""
import os, sys, time, json
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scene-text-recognition_b200"))
import ertext  # noqa: E402
from oracle.refbind import PortOracle  # noqa: E402


def blur(img, k):
    a = img.astype(np.float32)
    for _ in range(k):
        a = (a + np.roll(a, 1, 0) + np.roll(a, -1, 0) + np.roll(a, 1, 1) + np.roll(a, -1, 1)) / 5.0
    return np.clip(a, 0, 255).astype(np.uint8)


def make_plane(rng, h, w, kind):
    img = rng.randint(0, 256, (h, w)).astype(np.uint8)
    if kind == "noise":
        return img
    if kind == "smooth":
        b = blur(img, 6)
        return np.clip((b.astype(int) - 128) * 6 + 128, 0, 255).astype(np.uint8)
    if kind == "walls":
        b = blur(img, 3)
        b = np.clip((b.astype(int) - 128) * 5 + 128, 0, 255).astype(np.uint8)
        b[h // 3, : w - 3] = 255
        b[:, w // 2] = 255
        b[0, 0] = 255
        return b
    if kind == "wall0":
        b = blur(img, 2)
        b[0, 0] = 255; b[0, 1] = 255
        return b
    if kind == "allwall":
        b = blur(img, 2)
        b[0, 0] = 255; b[0, 1] = 254; b[1, 0] = 253
        return b
    if kind == "flat":
        return np.full((h, w), 77, np.uint8)
    raise ValueError(kind)


def cmp_plane(tag, got, exp, verbose=True):
    ok = True
    gn, en = got.nodes, exp["nodes"]
    if gn.shape != en.shape or not (gn == en).all():
        ok = False
        gs = sorted(map(tuple, gn[:, :6])); es = sorted(map(tuple, en[:, :6]))
        if verbose:
            print("  [%s] NODE MISMATCH: got %d exp %d ; multiset equal=%s" % (tag, len(gn), len(en), gs == es))
            if gs != es:
                sg, se = set(gs), set(es)
                print("     only-got:", sorted(sg - se)[:6], " only-exp:", sorted(se - sg)[:6])
            else:
                bad = np.nonzero((gn != en).any(1))[0][:4]
                for b in bad:
                    print("     row", b, "got", gn[b], "exp", en[b])
    gp = [tuple(gn[i][:6]) for i in got.pool] if len(got.pool) else []
    ep = [tuple(en[i][:6]) for i in exp["pool"]] if len(exp["pool"]) else []
    if gp != ep:
        ok = False
        if verbose:
            print("  [%s] POOL MISMATCH: got %d exp %d setdiff %d" % (tag, len(gp), len(ep), len(set(gp) ^ set(ep))))
    elif "label" in exp and len(ep):
        if not (got.label == exp["label"]).all():
            ok = False
            if verbose:
                print("  [%s] LABEL MISMATCH" % tag, got.label[:10], exp["label"][:10])
        if not (got.strong_score == exp["strong_score"]).all() or not (got.weak_score == exp["weak_score"]).all():
            ok = False
            if verbose:
                d = np.nonzero(got.strong_score != exp["strong_score"])[0][:3]
                print("  [%s] SCORE MISMATCH idx" % tag, d, got.strong_score[d], exp["strong_score"][d])
    return ok


def main():
    quick = "--quick" in sys.argv
    t0 = time.time()
    e = ertext.ErText()
    port = PortOracle()
    print("ctx created in %.2fs" % (time.time() - t0))
    rng = np.random.RandomState(3)
    cases = [(40, 50, "noise"), (32, 64, "smooth"), (33, 65, "smooth"), (64, 128, "walls"), (70, 90, "wall0"), (50, 60, "allwall"),
             (48, 48, "flat"), (100, 130, "noise"), (200, 300, "smooth"), (480, 640, "smooth"), (480, 640, "noise")]
    if quick:
        cases = cases[:6]
    summary = {}
    for mode in (0, 1):
        e.set_tile_local_union(mode)
        nok = 0
        for (h, w, kind) in cases:
            img = make_plane(rng, h, w, kind)
            for min_area in ((120, 3) if h * w < 40000 else (120,)):
                e.set_min_area(min_area)
                port.params["min_area"] = min_area
                exp = port.plane(img, classify=True, scores=True, canonical_order=True)
                try:
                    got = e.planes_detect(img).planes[0]
                except Exception as ex:  # noqa: BLE001
                    print("  EXC", (h, w, kind, min_area), ex)
                    continue
                tag = "lu%d %dx%d %s ma%d" % (mode, h, w, kind, min_area)
                ok = cmp_plane(tag, got, exp)
                nok += ok
                if ok:
                    print("  ok  [%s] nodes %d pool %d" % (tag, len(got.nodes), len(got.pool)))
        summary["local_union_%d" % mode] = nok
    e.set_tile_local_union(1); e.set_min_area(120); port.params["min_area"] = 120
    # golden frames, BGR path
    g = np.load(os.path.join(ROOT, "tests", "golden", "frames.npz"))["frames"]
    res = e.detect_classify(g)
    print("frames: status", res.status, "stage_ms", ["%.3f" % v for v in res.stage_ms], "launches", e.launch_count())
    nbad = 0
    for f in range(g.shape[0]):
        ch = port.channels(g[f])
        for k in range(6):
            exp = port.plane(ch[k], scores=True, canonical_order=True)
            nbad += not cmp_plane("frame%d plane%d" % (f, k), res.planes[f * 6 + k], exp)
    summary["golden_frame_planes_bad"] = nbad
    # timing at 1080p
    if not quick:
        big = np.stack([np.kron(g[i % 3], np.ones((3, 3, 1), np.uint8))[:1080, :1920] for i in range(4)])
        for it in range(3):
            t = time.time(); r = e.detect_classify(big); dt = time.time() - t
            print("1080p x4 frames: wall %.1f ms stage_ms %s kept %d pool %d status %d" % (
                dt * 1e3, ["%.2f" % v for v in r.stage_ms], sum(len(p.nodes) for p in r.planes), sum(len(p.pool) for p in r.planes), r.status))
    print("SUMMARY", json.dumps(summary))


if __name__ == "__main__":
    main()
