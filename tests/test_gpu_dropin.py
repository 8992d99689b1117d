"""GPU: the drop-in proof for SURVEY 8(b).  oracle/_ref/dropin_demo (oracle/build_dropin_test.sh) is the REFERENCE's own
ERFilter::text_detect (src/ER.cpp:33-111), the per-frame block of video_mode (src/utils.cpp:113-141) and OCR::chain_run
(src/OCR.cpp:67-140), extracted by line range and compiled unmodified against the reference's own headers, linked with
this repository's drop-in definitions (host/dropin/erfilter_dropin.cpp -> libertext.so, libertext_svm.so) instead of the
reference's hot-path code.  What those callers produce must equal what the reference's own code produced
(tests/golden/ref_next.npz, ref_planes.npz: generated from oracle/_ref)."""
import os
import subprocess
import numpy as np
import pytest
from conftest import ROOT, GOLDEN

pytestmark = pytest.mark.gpu
DEMO = os.path.join(ROOT, "oracle", "_ref", "dropin_demo")


@pytest.fixture(scope="module")
def demo_lines(tmp_path_factory):
    if not os.path.exists(DEMO):
        pytest.skip("oracle/_ref/dropin_demo not built (needs /root/reference at build time)")
    import ertext
    d = tmp_path_factory.mktemp("dropin")
    frames = np.load(os.path.join(GOLDEN, "frames.npz"))["frames"]
    with open(d / "frames.bin", "wb") as f:
        np.array([frames.shape[0], frames.shape[1], frames.shape[2]], np.int32).tofile(f)
        frames.tofile(f)
    g = np.load(os.path.join(GOLDEN, "ref_next.npz"))
    with open(d / "ocr_rows.txt", "w") as f:
        for row, sl in zip(g["ocr_rows"], g["ocr_slope"]):
            f.write("%d %d %d %d %d %d %.17g\n" % (*[int(v) for v in row], sl))
    cdir = os.path.join(ROOT, "assets", "classifier")
    out = subprocess.run([DEMO, str(d / "frames.bin"), os.path.join(cdir, "strong.classifier"), os.path.join(cdir, "weak.classifier"),
                          ertext.svm_model_path(), str(d / "ocr_rows.txt")], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = {}
    for ln in out.stdout.splitlines():
        t = ln.split()
        lines.setdefault(t[0], []).append(t[1:])
    return lines, g


def _rows(lines, tag, frame):
    return sorted(tuple(int(v) for v in t[1:7]) for t in lines.get(tag, []) if int(t[0]) == frame)


def test_reference_text_detect_and_video_loop_on_the_dropin(demo_lines):
    lines, g = demo_lines
    assert lines["num_iter"][0] == ["2660", "1354"]                      # CascadeBoost::get_num_iter of the reference's own loader
    nframes = 3
    for f in range(nframes):
        S = sorted(tuple(int(v) for v in r) for r in g["f%d_strong" % f])
        W = sorted(tuple(int(v) for v in r) for r in g["f%d_weak" % f])
        for pre in ("", "v"):                                            # text_detect, then video_mode's block
            assert _rows(lines, pre + "strong", f) == S, (pre, f)
            assert _rows(lines, pre + "weak", f) == W, (pre, f)
            exp_tr = sorted(tuple(int(v) for v in (g["f%d_strong" % f] if k == 0 else g["f%d_weak" % f])[i]) for k, i in g["f%d_tracked" % f])
            got = [t for t in lines.get(pre + "tracked", []) if int(t[0]) == f]
            assert sorted(tuple(int(v) for v in t[1:7]) for t in got) == exp_tr, (pre, f)
            # colours and centres are bit-identical doubles
            col = {tuple(int(v) for v in r): (tuple(c), tuple(int(x) for x in ctr)) for r, c, ctr in
                   zip(list(g["f%d_strong" % f]) + list(g["f%d_weak" % f]), list(g["f%d_strong_color" % f]) + list(g["f%d_weak_color" % f]),
                       list(g["f%d_strong_center" % f]) + list(g["f%d_weak_center" % f]))}
            for t in got:
                key = tuple(int(v) for v in t[1:7])
                c, ctr = col[key]
                assert (int(t[7]), int(t[8])) == ctr
                assert tuple(float(v) for v in t[9:12]) == c, (pre, f, key)
        assert any(int(t[0]) == f and t[1] == "7" for t in lines["times"])          # vector<double> times(7)
        assert ["%d" % f, "6"] in lines["vchannel_vec"]
    # pool of every plane: the same SET as the reference's (order: canonical siblings)
    p = np.load(os.path.join(GOLDEN, "ref_planes.npz"))
    for f in range(nframes):
        exp = sorted((k, int(n[2]), int(n[3]), int(n[4]), int(n[5]), int(n[1])) for k in range(6) for n in p["f%d_p%d_nodes" % (f, k)][p["f%d_p%d_pool" % (f, k)]])
        got = _rows(lines, "pool", f)
        assert len(set(exp) ^ set(got)) <= 2, (f, sorted(set(exp) ^ set(got)))


def test_reference_lbp_and_predict_signatures(demo_lines):
    lines, _ = demo_lines
    n, s, diff = lines["lbp"][0]
    assert (int(n), float(s), int(diff)) == (1024, 576.0, 0)             # make_LBP_hist == histogram of calc_LBP's image, 4 blocks x 144
    sp, wp = (float(v) for v in lines["predict"][0])
    assert sp == -1.7976931348623157e308 or sp > -1e300                  # a score or the reference's -DBL_MAX sentinel


def test_reference_ocr_chain_run_on_the_svm_shim(demo_lines):
    lines, g = demo_lines
    got = np.array([float(t[6]) for t in lines["ocr"]])
    exp = g["ocr_value"]
    assert got.shape == exp.shape
    assert (np.floor(got) == np.floor(exp)).all()                        # letters
    assert np.allclose(got - np.floor(got), exp - np.floor(exp), rtol=1e-4, atol=1e-12)   # probabilities (bar: 1e-4 relative)
