"""GPU parity for the rows after the detect path: ERFilter::er_track + calc_color and OCR::chain_run
(pre-processing, extract_feature, SVM), through the C ABI, against the committed golden vectors (produced by the
reference's own code) and, where oracle/_ref is present, against that code live."""
import os
import numpy as np
import pytest
from conftest import GOLDEN

pytestmark = pytest.mark.gpu

TABLE = "0123456789ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz&()"


@pytest.fixture(scope="module")
def golden_next():
    return np.load(os.path.join(GOLDEN, "ref_next.npz"))


def _rows(ft, lo, hi):
    c = ft.cand[lo:hi]
    return np.stack([c["plane"], c["x"], c["y"], c["w"], c["h"], c["area"]], axis=1).astype(np.int32).reshape(-1, 6)


def _check_track(ft, exp, tag):
    ns = ft.n_strong
    S, Wk = _rows(ft, 0, ns), _rows(ft, ns, len(ft.cand))
    assert (S == exp["strong"]).all() and (Wk == exp["weak"]).all(), tag
    col = np.stack([ft.cand["color1"], ft.cand["color2"], ft.cand["color3"]], axis=1).reshape(-1, 3)
    assert np.array_equal(col[:ns], exp["strong_color"], equal_nan=True), tag       # bit-exact doubles
    assert np.array_equal(col[ns:], exp["weak_color"], equal_nan=True), tag
    cen = np.stack([ft.cand["center_x"], ft.cand["center_y"]], axis=1).reshape(-1, 2)
    assert (cen[:ns] == exp["strong_center"]).all() and (cen[ns:] == exp["weak_center"]).all(), tag
    want = np.array([idx if kind == 0 else ns + idx for kind, idx in exp["tracked"]], np.int32)
    assert len(ft.tracked) == len(want) and (ft.tracked == want).all(), tag           # same regions, same order


def _check_track_sets(ft, exp, tag):
    """Order-free comparison against the golden vectors.  The GPU path visits NMS siblings in the canonical order
    (DESIGN.md 3), so pool order inside a plane -- hence the order of strong[ch] / weak[ch] and of all_er -- may differ
    from the reference's while the SETS are equal: per region colours/centres and the set of tracked regions (the
    growth is a reachability closure from the strong seeds, independent of visiting order)."""
    ns = ft.n_strong
    def key(rows):
        return [tuple(int(v) for v in r) for r in rows]
    S, Wk = key(_rows(ft, 0, ns)), key(_rows(ft, ns, len(ft.cand)))
    assert sorted(S) == sorted(key(exp["strong"])) and sorted(Wk) == sorted(key(exp["weak"])), tag
    col = np.stack([ft.cand["color1"], ft.cand["color2"], ft.cand["color3"]], axis=1).reshape(-1, 3)
    got = {("s" if i < ns else "w",) + (S + Wk)[i]: tuple(col[i]) for i in range(len(ft.cand))}
    want = {("s",) + k: tuple(c) for k, c in zip(key(exp["strong"]), exp["strong_color"])}
    want.update({("w",) + k: tuple(c) for k, c in zip(key(exp["weak"]), exp["weak_color"])})
    assert got == want, tag
    got_t = sorted(("s" if i < ns else "w",) + (S + Wk)[i] for i in ft.tracked)
    want_t = sorted((("s",) + key(exp["strong"])[idx]) if kind == 0 else (("w",) + key(exp["weak"])[idx]) for kind, idx in exp["tracked"])
    assert got_t == want_t, tag


def _live(ref, bgr, ft):
    """the reference's er_track fed with the GPU's own strong / weak lists (same order) -> exact expectation"""
    ns = ft.n_strong
    S, Wk = _rows(ft, 0, ns), _rows(ft, ns, len(ft.cand))
    ch = ref.channels(bgr)
    exp = ref.er_track(ch, np.stack([ch[0], ch[1], ch[2]], axis=-1), S, Wk)
    exp.update(strong=S, weak=Wk)
    return exp


def _gold(g, tag):
    return {k: g["%s_%s" % (tag, k)] for k in ("strong", "weak", "tracked", "strong_color", "weak_color", "strong_center", "weak_center")}


def test_er_track_batch_matches_golden(ert, ref, golden_frames, golden_next):
    """detect_classify on the golden frames, then er_track on the device-resident batch == the reference's er_track."""
    import ertext
    res = ert.detect_classify(golden_frames)
    assert res.status == 0
    tracks, ms = ert.er_track()
    assert len(tracks) == 3 and ms > 0
    for f in range(3):
        _check_track_sets(tracks[f], _gold(golden_next, "f%d" % f), "f%d" % f)
        _check_track(tracks[f], _live(ref, golden_frames[f], tracks[f]), "f%d/live" % f)     # exact order, exact doubles
        # cand rows point back into the batch result
        for c in tracks[f].cand:
            pl = res.planes[f * 6 + c["plane"]]
            assert pl.pool[c["pool_index"]] == c["node"] and pl.label[c["pool_index"]] == c["label"]
            assert tuple(pl.nodes[c["node"]][:6]) == (c["level"], c["area"], c["x"], c["y"], c["w"], c["h"])
    # the same in one stream submission (upto = ERT_STAGE_TRACK)
    res2 = ert.detect_classify(golden_frames, upto=ertext.STAGE_TRACK)
    tracks2, _ = ert.er_track()
    for f in range(3):
        _check_track(tracks2[f], _live(ref, golden_frames[f], tracks2[f]), "f%d/fused" % f)
        assert (tracks2[f].tracked == tracks[f].tracked).all()


def test_er_track_regions_matches_golden(ert, golden_frames, golden_next):
    for c in range(6):
        ft = ert.er_track_regions(golden_frames[1], golden_next["s%d_strong" % c], golden_next["s%d_weak" % c])
        _check_track(ft, _gold(golden_next, "s%d" % c), "s%d" % c)


def test_er_track_live_random_and_nan(ert, ref):
    """seeded frames and region lists (incl. saturated planes whose OTSU mask is empty -> NaN colours) vs the reference live"""
    rng = np.random.RandomState(5)
    H, W = 240, 320
    for t in range(6):
        if t == 0:
            bgr = np.full((H, W, 3), 255, np.uint8)          # Y = 255: empty mask on channel 0, full mask on channel 3
        elif t == 1:
            bgr = np.zeros((H, W, 3), np.uint8)
        else:
            base = rng.randint(0, 256, (H // 8, W // 8, 3)).astype(np.uint8)
            bgr = np.kron(base, np.ones((8, 8, 1), np.uint8)) + rng.randint(0, 12, (H, W, 3)).astype(np.uint8)
            bgr = np.ascontiguousarray(bgr.astype(np.uint8))
        def boxes(n):
            out = []
            for _ in range(n):
                w, h = rng.randint(1, 60), rng.randint(1, 80)
                out.append((rng.randint(0, 6), rng.randint(0, W - w + 1), rng.randint(0, H - h + 1), w, h, rng.randint(121, 4000)))
            out.sort(key=lambda r: r[0])
            return np.array(out, np.int32).reshape(-1, 6)
        S, Wk = boxes(rng.randint(0, 30)), boxes(rng.randint(0, 90))
        ch = ref.channels(bgr)
        exp = ref.er_track(ch, np.stack([ch[0], ch[1], ch[2]], axis=-1), S, Wk)
        exp.update(strong=S, weak=Wk)
        _check_track(ert.er_track_regions(bgr, S, Wk), exp, "live%d" % t)


def test_ocr_features_match_golden(ert, port, golden_frames, golden_next):
    g = golden_next
    chans = [port.channels(golden_frames[f]) for f in range(3)]
    rows = g["ocr_rows"]
    for f in range(3):
        for k in range(6):
            sel = np.where((rows[:, 0] == f) & (rows[:, 1] == k))[0]
            if not len(sel):
                continue
            r = ert.ocr_features_plane(chans[f][k], rows[sel][:, 2:6], g["ocr_slope"][sel])
            assert (r.img == g["ocr_img"][sel]).all(), (f, k)
            assert (r.feat == g["ocr_feat"][sel]).all(), (f, k)        # bit-exact feature bytes


def test_ocr_chain_run_matches_golden(ert, port, golden_frames, golden_next):
    g = golden_next
    rows = g["ocr_rows"]
    # device-resident batch form: regions address (frame, channel) of the batch processed last
    ert.detect_classify(golden_frames)
    r = ert.ocr_chain_run_batch(rows[:, 0], rows[:, 1], rows[:, 2:6], g["ocr_slope"])
    assert (r.feat == g["ocr_feat"]).all()
    v = g["ocr_value"]
    assert (np.floor(r.value) == np.floor(v)).all()                   # er->letter
    assert np.allclose(r.value - np.floor(r.value), v - np.floor(v), rtol=1e-4, atol=1e-9)      # er->prob, 1e-4 relative
    assert [TABLE[l] for l in r.label] == [chr(int(x)) for x in np.floor(v)]
    assert np.allclose(r.prob.sum(axis=1), 1.0, atol=1e-9)
    # plane form gives the same numbers
    chans = port.channels(golden_frames[1])
    sel = np.where((rows[:, 0] == 1) & (rows[:, 1] == 0))[0]
    rp = ert.ocr_chain_run_plane(chans[0], rows[sel][:, 2:6], g["ocr_slope"][sel])
    assert (rp.value == r.value[sel]).all() and (rp.feat == r.feat[sel]).all()


def test_ocr_features_live_random(ert, ref):
    """seeded planes, ragged rectangles (1-pixel sides, exact-2x sizes, full plane) and slopes vs the reference live"""
    rng = np.random.RandomState(9)
    H, W = 200, 260
    base = rng.randint(0, 256, (H // 10, W // 10)).astype(np.uint8)
    plane = np.kron(base, np.ones((10, 10), np.uint8))
    plane = np.clip(plane.astype(int) + rng.randint(-10, 10, (H, W)), 0, 255).astype(np.uint8)
    rects, slopes = [], []
    for t in range(120):
        w, h = rng.randint(2, 120), rng.randint(2, 120)
        if t % 10 == 0:
            w = h = 2 * rng.randint(8, 31)
        if t % 17 == 0:
            w, h = W, H
        rects.append((rng.randint(0, W - w + 1), rng.randint(0, H - h + 1), w, h))
        slopes.append(0.0 if t % 3 == 0 else float(rng.uniform(-1.2, 1.2)))
    keep_r, keep_s, exp_img, exp_feat = [], [], [], []
    for (x, y, w, h), sl in zip(rects, slopes):
        try:
            img, feat = ref.ocr_features(plane[y:y + h, x:x + w], sl)
        except Exception:
            continue
        keep_r.append((x, y, w, h)); keep_s.append(sl); exp_img.append(img); exp_feat.append(feat)
    assert len(keep_r) > 60
    r = ert.ocr_features_plane(plane, np.array(keep_r, np.int32), np.array(keep_s))
    bad = [i for i in range(len(keep_r)) if not ((r.img[i] == exp_img[i]).all() and (r.feat[i] == exp_feat[i]).all())]
    assert not bad, [(keep_r[i], keep_s[i]) for i in bad[:5]]


def test_ocr_rejects_bad_regions(ert, golden_frames, port):
    import ertext
    plane = port.channels(golden_frames[0])[0]
    with pytest.raises(ertext.ErtError):
        ert.ocr_features_plane(plane, [(630, 470, 20, 20)])           # outside the plane
    with pytest.raises(ertext.ErtError):
        ert.ocr_features_plane(np.zeros((4, 1000), np.uint8), [(0, 0, 1000, 1)])   # ARAN collapses to 30x0 (cv::resize would throw)
    with pytest.raises(ertext.ErtError):
        ert.er_track_regions(golden_frames[0], [(3, 0, 0, 5, 5, 130), (1, 0, 0, 5, 5, 130)], np.zeros((0, 6), np.int32))   # not channel-major
    r = ert.ocr_features_plane(plane, np.zeros((0, 4), np.int32))     # empty batch
    assert r.feat.shape == (0, 1800)


def test_er_track_and_ocr_at_bench_size(ert, ref):
    """BASELINE's full frame size: two synthetic 1080p S-text frames through detect -> er_track (fused submission) ->
    chain_run on every tracked region, each checked against the reference's own code on the same inputs."""
    import ertext
    from ertext import synth
    frames = synth.s_text_batch(4321, 2, 1920, 1080)
    res = ert.detect_classify(frames, upto=ertext.STAGE_TRACK)
    assert res.status == 0
    tracks, _ = ert.er_track()
    fr, pl, rc, exp_feat, exp_val = [], [], [], [], []
    for f in range(2):
        ft = tracks[f]
        assert len(ft.cand) > 20 and len(ft.tracked) >= ft.n_strong
        _check_track(ft, _live(ref, frames[f], ft), "1080p/%d" % f)
        ch = ref.channels(frames[f])
        for j, i in enumerate(ft.tracked[:40]):
            c = ft.cand[i]
            k, x, y, w, h = int(c["plane"]), int(c["x"]), int(c["y"]), int(c["w"]), int(c["h"])
            sl = [0.0, 0.07, -0.2][j % 3]
            crop = ch[k][y:y + h, x:x + w]
            fr.append(f); pl.append(k); rc.append((x, y, w, h))
            exp_feat.append(ref.ocr_features(crop, sl)[1]); exp_val.append(ref.chain_run(crop, 0, sl))
    sl = np.array([[0.0, 0.07, -0.2][j % 3] for f in range(2) for j in range(min(40, len(tracks[f].tracked)))])
    o = ert.ocr_chain_run_batch(np.array(fr, np.int32), np.array(pl, np.int32), np.array(rc, np.int32), sl)
    assert (o.feat == np.stack(exp_feat)).all()
    ev = np.array(exp_val)
    assert (np.floor(o.value) == np.floor(ev)).all()
    assert np.allclose(o.value - np.floor(o.value), ev - np.floor(ev), rtol=1e-4, atol=1e-9)


def test_next_rows_properties_at_full_size(ert):
    """Size-independent properties on 1080p frames (no oracle needed): results do not depend on where a frame sits in the
    batch, on the order / composition of an OCR batch, on presenting a crop alone or inside its plane; the tracked SET is
    the closure from the strong seeds, so it is invariant under any reordering of the weak list."""
    import ertext
    from ertext import synth
    frames = synth.s_text_batch(777, 3, 1920, 1080)
    ert.detect_classify(frames, upto=ertext.STAGE_TRACK)
    tr_a, _ = ert.er_track()
    perm = [2, 0, 1]
    ert.detect_classify(frames[perm], upto=ertext.STAGE_TRACK)
    tr_b, _ = ert.er_track()
    for i, f in enumerate(perm):                                      # frame position in the batch is irrelevant
        assert tr_b[i].cand.tobytes() == tr_a[f].cand.tobytes() and (tr_b[i].tracked == tr_a[f].tracked).all()
    # er_track on caller lists: shuffle the weak rows inside each channel -> same tracked set, same colours per region
    ft = tr_a[0]
    ns = ft.n_strong
    rows = np.stack([ft.cand["plane"], ft.cand["x"], ft.cand["y"], ft.cand["w"], ft.cand["h"], ft.cand["area"]], axis=1).astype(np.int32)
    S, Wk = rows[:ns], rows[ns:]
    rng = np.random.RandomState(1)
    order = np.lexsort((rng.rand(len(Wk)), Wk[:, 0]))                 # random inside a channel, channel-major overall
    ft2 = ert.er_track_regions(frames[0], S, Wk[order])
    key = lambda c: (int(c["plane"]), int(c["x"]), int(c["y"]), int(c["w"]), int(c["h"]), int(c["area"]), int(c["label"]))
    assert sorted(key(ft.cand[i]) for i in ft.tracked) == sorted(key(ft2.cand[i]) for i in ft2.tracked)
    col = lambda t: {key(c): (c["color1"], c["color2"], c["color3"]) for c in t.cand}
    assert col(ft) == col(ft2)
    # OCR: batch order / composition, and crop-alone vs crop-in-plane
    ert.detect_classify(frames)
    c = ft.cand[ft.tracked][:48]
    fr = np.zeros(len(c), np.int32)
    rc = np.stack([c["x"], c["y"], c["w"], c["h"]], axis=1).astype(np.int32)
    sl = np.where(np.arange(len(c)) % 3 == 1, 0.12, 0.0)
    o1 = ert.ocr_chain_run_batch(fr, c["plane"], rc, sl)
    p2 = rng.permutation(len(c))
    o2 = ert.ocr_chain_run_batch(fr[p2], c["plane"][p2], rc[p2], sl[p2])
    assert (o2.feat == o1.feat[p2]).all() and (o2.value == o1.value[p2]).all()
    o3 = ert.ocr_chain_run_batch(fr[:5], c["plane"][:5], rc[:5], sl[:5])
    assert (o3.feat == o1.feat[:5]).all() and np.allclose(o3.value, o1.value[:5], rtol=1e-12, atol=0)
    planes = ert.compute_channels(frames[0])
    for i in range(6):
        k, (x, y, w, h) = int(c["plane"][i]), rc[i]
        alone = np.ascontiguousarray(planes[k][y:y + h, x:x + w])
        oa = ert.ocr_features_plane(alone, [(0, 0, w, h)], [sl[i]])
        assert (oa.feat[0] == o1.feat[i]).all() and (oa.img[0] == o1.img[i]).all()
