"""CPU: the oracle for the rows after the detect path (er_track, OCR::chain_run) is pinned.
(1) the OpenCV primitives the reference calls there (threshold OTSU, findContours, GaussianBlur, normalize), restated in
    oracle/cvshim, are checked against python cv2; (2) the reference's own code on top of them (oracle/_ref) reproduces the
    committed golden vectors; (3) chain_run == its visible stages + svm_predict_probability."""
import os
import numpy as np
import pytest
from conftest import GOLDEN

TABLE = "0123456789ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz&()"


@pytest.fixture(scope="module")
def golden_next():
    return np.load(os.path.join(GOLDEN, "ref_next.npz"))


def _images(rng, t):
    h, w = rng.randint(4, 48), rng.randint(4, 48)
    kind = t % 4
    if kind == 0:
        img = rng.randint(0, 256, (h, w))
    elif kind == 1:
        img = np.clip(rng.normal(80, 20, (h, w)), 0, 255); img[rng.rand(h, w) < 0.3] = rng.randint(150, 255)
    elif kind == 2:
        img = (rng.rand(h, w) < rng.uniform(0.1, 0.9)) * 255
    else:
        img = np.full((h, w), rng.randint(0, 256))
    return np.ascontiguousarray(img).astype(np.uint8)


def test_cvshim_otsu_contours_blur_normalize_vs_cv2(ref):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.RandomState(3)
    for t in range(600):
        img = _images(rng, t)
        tcv, dcv = cv2.threshold(img, 128, 255, cv2.THRESH_OTSU)
        tm, dm = ref.prim_threshold_otsu(img)
        assert int(tcv) == tm and (dcv == dm).all(), t
        cs, _ = cv2.findContours(dm.copy(), cv2.RETR_LIST, cv2.CHAIN_APPROX_NONE)
        mine = ref.prim_find_contours(dm)
        a = sorted(tuple(map(tuple, c.reshape(-1, 2))) for c in cs)
        b = sorted(tuple(map(tuple, c)) for c in mine)
        assert a == b, t
        g = cv2.GaussianBlur(dm, (7, 7), 0)
        assert (g == ref.prim_gaussian7(dm)).all(), t
        f = g.copy()
        cv2.normalize(f, f, 0, 255, cv2.NORM_MINMAX, cv2.CV_8U)
        assert (f == ref.prim_normalize_minmax(g)).all(), t


def test_cvshim_normalize_on_feature_maps_vs_cv2(ref):
    """normalize's float rounding on the images extract_feature really feeds it: blurred sparse 30x30 bitmaps, in place."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.RandomState(4)
    for t in range(400):
        img = np.zeros((30, 30), np.uint8)
        img[rng.rand(30, 30) < rng.uniform(0.01, 0.25)] = 255
        g = cv2.GaussianBlur(img, (7, 7), 0)
        f = g.copy()
        cv2.normalize(f, f, 0, 255, cv2.NORM_MINMAX, cv2.CV_8U)
        assert (f == ref.prim_normalize_minmax(g)).all(), t
        assert (cv2.resize(f, (15, 15)) == ref.resize(f, 15, 15)).all()


def test_reference_reproduces_golden_track(ref, golden_frames, golden_next):
    g = golden_next
    for tag, f in [("f0", 0), ("f1", 1), ("f2", 2)] + [("s%d" % c, 1) for c in range(6)]:
        ch = ref.channels(golden_frames[f])
        ycc = np.stack([ch[0], ch[1], ch[2]], axis=-1)
        r = ref.er_track(ch, ycc, g[tag + "_strong"], g[tag + "_weak"])
        assert (r["tracked"] == g[tag + "_tracked"]).all(), tag
        for k in ("strong_color", "weak_color"):
            assert np.array_equal(r[k], g[tag + "_" + k], equal_nan=True), (tag, k)
        for k in ("strong_center", "weak_center"):
            assert (r[k] == g[tag + "_" + k]).all(), (tag, k)


def test_golden_frame_lists_are_classify_output(ref, golden_frames, golden_planes, golden_next):
    """the strong / weak rows in the fixture are exactly what the detect path's golden vectors hold"""
    for f in range(3):
        S, Wk = [], []
        for k in range(6):
            nodes = golden_planes["f%d_p%d_nodes" % (f, k)]
            for pi, lab in zip(golden_planes["f%d_p%d_pool" % (f, k)], golden_planes["f%d_p%d_label" % (f, k)]):
                n = nodes[pi]
                (S if lab == 2 else Wk if lab == 1 else []).append((k, n[2], n[3], n[4], n[5], n[1]))
        assert (np.array(S, np.int32).reshape(-1, 6) == golden_next["f%d_strong" % f]).all()
        assert (np.array(Wk, np.int32).reshape(-1, 6) == golden_next["f%d_weak" % f]).all()


def test_calc_color_reads_colour_from_the_image_origin(ref, golden_frames):
    """the reference's quirk (src/ER.cpp:1404): colour rows/cols are counted from the image origin, not the bound"""
    ch = ref.channels(golden_frames[1])
    ycc = np.stack([ch[0], ch[1], ch[2]], axis=-1)
    x, y, w, h = 200, 150, 40, 50
    got = ref.calc_color(ch[0], ycc, [(x, y, w, h)])[0]
    import cv2
    crop = 255 - ch[0][y:y + h, x:x + w]
    _t, mask = cv2.threshold(crop, 128, 255, cv2.THRESH_OTSU)
    exp = [ycc[:h, :w, c][mask != 0].astype(np.float64).sum() / (mask != 0).sum() for c in range(3)]
    assert np.allclose(got, exp, rtol=0, atol=0)


def test_reference_reproduces_golden_ocr(ref, golden_frames, golden_next):
    g = golden_next
    chans = [ref.channels(golden_frames[f]) for f in range(3)]
    for i, (f, k, x, y, w, h) in enumerate(g["ocr_rows"]):
        crop = chans[f][k][y:y + h, x:x + w]
        sl = float(g["ocr_slope"][i])
        img, feat = ref.ocr_features(crop, sl)
        assert (img == g["ocr_img"][i]).all() and (feat == g["ocr_feat"][i]).all(), i
        v = ref.chain_run(crop, 0, sl)
        assert v == g["ocr_value"][i], i
        # chain_run == svm_predict_probability on the visible feature vector
        lab, prob = ref.svm_predict_probability(feat[None].astype(np.float64) / 255.0)
        assert ord(TABLE[int(lab[0])]) + prob[0, int(lab[0])] == v, i


def _cv2_chain_run_features(cv2, crop):
    """OCR::chain_run's pre-processing + extract_feature (src/OCR.cpp:72-84, 144-218, slope 0) written with the REAL OpenCV
    functions of python cv2 -- an end-to-end check of the reference-backed oracle (reference code + oracle/cvshim)."""
    import math
    _t, img = cv2.threshold(255 - crop, 128, 255, cv2.THRESH_OTSU)
    rows, cols = img.shape
    L = 30
    R1 = rows / cols if cols > rows else cols / rows
    minor = int(L * math.pow(R1, 0.5))
    size = (L, minor) if cols > rows else (minor, L)            # cv::Size(width, height)
    tmp = cv2.resize(img, size)
    dst = np.zeros((L, L), np.uint8)
    if tmp.shape[1] > tmp.shape[0]:
        off = int(round((L - tmp.shape[0]) // 2))
        dst[off:off + tmp.shape[0], :tmp.shape[1]] = tmp
    else:
        off = int(round((L - tmp.shape[1]) // 2))
        dst[:tmp.shape[0], off:off + tmp.shape[1]] = tmp
    f = [np.zeros((L, L), np.uint8) for _ in range(8)]
    contours, _h = cv2.findContours(dst.copy(), cv2.RETR_LIST, cv2.CHAIN_APPROX_NONE)

    def direction(p1, p2):                                       # OCR::chain_code_direction(next, current), src/OCR.cpp:602-622
        if p1[0] < p2[0] and p1[1] == p2[1]: return 0
        if p1[0] < p2[0] and p1[1] < p2[1]: return 1
        if p1[0] == p2[0] and p1[1] < p2[1]: return 2
        if p1[0] > p2[0] and p1[1] < p2[1]: return 3
        if p1[0] > p2[0] and p1[1] == p2[1]: return 4
        if p1[0] > p2[0] and p1[1] > p2[1]: return 5
        if p1[0] == p2[0] and p1[1] > p2[1]: return 6
        return 7
    for c in contours:
        pts = [tuple(p) for p in c.reshape(-1, 2)]
        if len(pts) == 1:
            continue
        pts.append(pts[0])
        for j in range(len(pts) - 1):
            f[direction(pts[j + 1], pts[j])][pts[j][1], pts[j][0]] = 255
    feat = []
    for k in range(8):
        g = cv2.GaussianBlur(f[k], (7, 7), 0)
        cv2.normalize(g, g, 0, 255, cv2.NORM_MINMAX, cv2.CV_8U)
        feat.append(cv2.resize(g, (15, 15)).ravel())
    return dst, np.concatenate(feat)


def test_ocr_oracle_matches_a_cv2_pipeline_end_to_end(ref, golden_frames):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.RandomState(12)
    ch = ref.channels(golden_frames[1])
    n = 0
    for t in range(220):
        k = rng.randint(0, 6)
        w, h = rng.randint(4, 160), rng.randint(4, 160)
        if t % 9 == 0:
            w = h = 2 * rng.randint(6, 31)                       # hits the exact-2x INTER_AREA switch when it reaches 60
        x, y = rng.randint(0, 640 - w), rng.randint(0, 480 - h)
        crop = np.ascontiguousarray(ch[k][y:y + h, x:x + w])
        if int(30 * (min(w, h) / max(w, h)) ** 0.5) < 1:
            continue
        img, feat = ref.ocr_features(crop, 0.0)
        eimg, efeat = _cv2_chain_run_features(cv2, crop)
        assert (img == eimg).all(), (t, w, h)
        assert (feat == efeat).all(), (t, w, h)
        n += 1
    assert n > 200


def test_calc_color_oracle_matches_a_cv2_pipeline(ref, golden_frames):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.RandomState(13)
    ch = ref.channels(golden_frames[2])
    ycc = cv2.cvtColor(golden_frames[2], cv2.COLOR_BGR2YCrCb)
    for t in range(200):
        k = rng.randint(0, 6)
        w, h = rng.randint(1, 200), rng.randint(1, 200)
        x, y = rng.randint(0, 640 - w), rng.randint(0, 480 - h)
        _t, mask = cv2.threshold(255 - ch[k][y:y + h, x:x + w], 128, 255, cv2.THRESH_OTSU)
        m = mask != 0
        got = ref.calc_color(ch[k], ycc, [(x, y, w, h)])[0]
        if m.sum() == 0:
            assert np.isnan(got).all()
        else:
            exp = [ycc[:h, :w, c][m].astype(np.float64).sum() / m.sum() for c in range(3)]      # rows / cols from the image origin
            assert (got == np.array(exp)).all(), (t, x, y, w, h)
