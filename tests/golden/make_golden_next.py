"""Golden fixtures for the rows after the detect path (er_track, OCR::chain_run), generated from the REFERENCE's own
code (oracle/_ref: src/ER.cpp:532-609,1391-1437 and src/OCR.cpp:67-250,254-360,394-430 compiled verbatim).

Run here (where /root/reference exists):  python tests/golden/make_golden_next.py
Output (committed): tests/golden/ref_next.npz
  f{i}_strong / f{i}_weak      rows (ch, x, y, w, h, area): classify's strong[ch] / weak[ch] of golden frame i, channel-major
  f{i}_tracked                 er_track's all_er as (kind 0 strong / 1 weak, row)
  f{i}_strong_color / _weak_color / _strong_center / _weak_center
  s{j}_*                       the same for seeded synthetic region lists on frame 1 (dense matches, empty lists)
  ocr_rows                     (frame, ch, x, y, w, h) of the regions sent through OCR::chain_run
  ocr_slope, ocr_value         chain_run(channel(bound), 0, slope) return values
  ocr_img, ocr_feat            the 30x30 image extract_feature received and the 1800 feature bytes (value * 255)
"""
import os, sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.refbind import RefOracle  # noqa: E402


def frame_lists(ref, ch):
    S, Wk = [], []
    for k in range(6):
        r = ref.plane(ch[k])
        for pi, lab in zip(r["pool"], r["label"]):
            n = r["nodes"][pi]
            row = (k, n[2], n[3], n[4], n[5], n[1])
            if lab == 2:
                S.append(row)
            elif lab == 1:
                Wk.append(row)
    return np.array(S, np.int32).reshape(-1, 6), np.array(Wk, np.int32).reshape(-1, 6)


def synth_lists(rng, case, W, H):
    """Region lists that stress the greedy growth: clusters of similar boxes so that chains of matches form."""
    def boxes(n, cx, cy, s):
        return sorted(boxes_(n, cx, cy, s), key=lambda r: r[0])       # channel-major like classify's output (stable)

    def boxes_(n, cx, cy, s):
        out = []
        for _ in range(n):
            w = int(np.clip(rng.normal(s, s * 0.25), 4, W // 2)); h = int(np.clip(rng.normal(s * 1.3, s * 0.3), 4, H // 2))
            x = int(np.clip(cx + rng.normal(0, 3 * s), 0, W - w)); y = int(np.clip(cy + rng.normal(0, s), 0, H - h))
            out.append((rng.randint(0, 6), x, y, w, h, int(w * h * rng.uniform(0.3, 0.9)) + 121))
        return out
    if case == 0:      # no strong seeds: nothing is tracked
        return np.zeros((0, 6), np.int32), np.array(boxes(20, 300, 200, 24), np.int32)
    if case == 1:      # no weak regions
        return np.array(boxes(12, 300, 200, 24), np.int32), np.zeros((0, 6), np.int32)
    if case == 2:      # both empty
        return np.zeros((0, 6), np.int32), np.zeros((0, 6), np.int32)
    S = boxes(6 + 4 * case, rng.randint(100, 500), rng.randint(100, 380), 14 + 3 * case)
    Wk = boxes(40 + 30 * case, rng.randint(100, 500), rng.randint(100, 380), 14 + 3 * case)
    return np.array(S, np.int32), np.array(Wk, np.int32)


def main():
    frames = np.load(os.path.join(HERE, "frames.npz"))["frames"]
    ref = RefOracle(with_svm=True)
    out = {}
    ocr_rows, ocr_slope, ocr_value, ocr_img, ocr_feat = [], [], [], [], []
    slopes = [0.0, 0.0, 0.05, -0.12, 0.3, -0.4, 0.009, 0.0, 0.011, -0.7]
    for f in range(frames.shape[0]):
        ch = ref.channels(frames[f])
        ycc = np.stack([ch[0], ch[1], ch[2]], axis=-1)
        S, Wk = frame_lists(ref, ch)
        r = ref.er_track(ch, ycc, S, Wk)
        out["f%d_strong" % f] = S; out["f%d_weak" % f] = Wk
        for k, v in r.items():
            out["f%d_%s" % (f, k)] = v
        for i, (kind, idx) in enumerate(r["tracked"][:24]):
            row = (S if kind == 0 else Wk)[idx]
            k, x, y, w, h, _a = [int(v) for v in row]
            sl = slopes[(i + f) % len(slopes)]
            crop = ch[k][y:y + h, x:x + w]
            img, feat = ref.ocr_features(crop, sl)
            ocr_rows.append((f, k, x, y, w, h)); ocr_slope.append(sl); ocr_value.append(ref.chain_run(crop, 0, sl))
            ocr_img.append(img); ocr_feat.append(feat)
    # odd shapes on frame 1: exact-2x (INTER_AREA path), tiny, long, large
    ch = ref.channels(frames[1])
    for (k, x, y, w, h, sl) in [(0, 100, 100, 60, 60, 0.0), (3, 40, 30, 60, 60, 0.2), (1, 5, 7, 3, 5, 0.0), (4, 200, 10, 9, 80, 0.0),
                                (2, 10, 300, 300, 40, 0.0), (5, 20, 20, 500, 400, -0.08), (0, 320, 240, 30, 30, 0.0), (0, 0, 0, 640, 480, 0.0),
                                (3, 600, 440, 40, 40, 1.0), (0, 50, 60, 15, 30, -1.5)]:
        crop = ch[k][y:y + h, x:x + w]
        img, feat = ref.ocr_features(crop, sl)
        ocr_rows.append((1, k, x, y, w, h)); ocr_slope.append(sl); ocr_value.append(ref.chain_run(crop, 0, sl))
        ocr_img.append(img); ocr_feat.append(feat)
    out.update(ocr_rows=np.array(ocr_rows, np.int32), ocr_slope=np.array(ocr_slope), ocr_value=np.array(ocr_value),
               ocr_img=np.stack(ocr_img), ocr_feat=np.stack(ocr_feat))
    # synthetic lists for er_track on frame 1 (the NaN-colour case -- empty OTSU mask -- is tested live in tests/test_gpu_next.py)
    rng = np.random.RandomState(11)
    ycc = np.stack([ch[0], ch[1], ch[2]], axis=-1)
    for case in range(6):
        S, Wk = synth_lists(rng, case, 640, 480)
        r = ref.er_track(ch, ycc, S, Wk)
        out["s%d_strong" % case] = S; out["s%d_weak" % case] = Wk
        for k, v in r.items():
            out["s%d_%s" % (case, k)] = v
    np.savez_compressed(os.path.join(HERE, "ref_next.npz"), **out)
    print("tracked per frame:", [len(out["f%d_tracked" % f]) for f in range(frames.shape[0])],
          "synthetic:", [len(out["s%d_tracked" % c]) for c in range(6)], "ocr cases:", len(ocr_rows))


if __name__ == "__main__":
    main()
