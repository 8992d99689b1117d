"""JPEG bitstreams in, decoded on the device (nvJPEG), then the hot path.  Parity is stated on the decoded
pixels (ert_jpeg_fetch_frames): nvJPEG and libjpeg differ in the last bit of the IDCT / upsampling, the path after the
decode is bit-exact against the oracle on the SAME pixels."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scene-text-recognition_b200"))

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ert():
    import ertext
    e = ertext.ErText(device=0)
    yield e
    e.close()


def _encode(frames, quality=90):
    cv2 = pytest.importorskip("cv2")
    out = []
    for f in frames:
        ok, buf = cv2.imencode(".jpg", f, [int(cv2.IMWRITE_JPEG_QUALITY), quality])
        assert ok
        out.append(buf.tobytes())
    return out


def test_jpeg_batch_decodes_on_device_and_matches_oracle_on_decoded_pixels(ert):
    cv2 = pytest.importorskip("cv2")
    from oracle.refbind import PortOracle
    frames = np.load(os.path.join(ROOT, "tests", "golden", "frames.npz"))["frames"]
    jpegs = _encode(frames)
    f, h, w, _ = frames.shape
    ert.enqueue_jpeg(jpegs, w, h)
    res = ert.fetch()
    assert res.status == 0
    dec = ert.jpeg_fetch_frames()
    assert dec.shape == frames.shape
    assert ert.jpeg_backend_name() in ("hardware", "gpu_hybrid", "default", "hybrid")
    # the decoder is a JPEG decoder: close to libjpeg's output (not identical: IDCT / upsampling rounding)
    for i in range(f):
        ref = cv2.imdecode(np.frombuffer(jpegs[i], np.uint8), cv2.IMREAD_COLOR)
        d = np.abs(dec[i].astype(np.int32) - ref.astype(np.int32))
        assert d.mean() < 2.5, (d.max(), d.mean())
    # the path behind the decode: bit-exact against the oracle on the decoded pixels
    port = PortOracle()
    for i in range(f):
        ch = port.channels(dec[i])
        for k in range(6):
            exp = port.plane(ch[k], scores=True, canonical_order=True)
            got = res.planes[i * 6 + k]
            assert got.nodes.shape == exp["nodes"].shape and (got.nodes == exp["nodes"]).all(), (i, k)
            assert (got.pool == exp["pool"]).all() and (got.label == exp["label"]).all(), (i, k)
    # and the same result as handing the decoded pixels in through the host entry point
    res2 = ert.detect_classify(dec)
    for a, b in zip(res.planes, res2.planes):
        assert (a.nodes == b.nodes).all() and (a.pool == b.pool).all() and (a.label == b.label).all()


def test_jpeg_errors_are_loud(ert):
    frames = np.load(os.path.join(ROOT, "tests", "golden", "frames.npz"))["frames"]
    jpegs = _encode(frames[:1])
    import ertext
    with pytest.raises(ertext.ErtError):
        ert.enqueue_jpeg(jpegs, 320, 240)              # declared size does not match the bitstream
    with pytest.raises(ertext.ErtError):
        ert.enqueue_jpeg([b"not a jpeg at all"], 640, 480)


def test_cpp_video_loop_on_jpeg_input_accumulates_two_frames(ert, tmp_path):
    """host/FramePipeline.hpp: push_jpeg (device decode) -> per-frame regions in order -> FrameAccumulator(2) = video_mode's
    tracked_vec of two consecutive frames (src/utils.cpp:106-144), VideoTimes = avg_time[7].  The accumulated tracked lists
    equal what the binding returns for the same bitstreams."""
    import subprocess
    import ertext
    frames = np.load(os.path.join(ROOT, "tests", "golden", "frames.npz"))["frames"]
    jpegs = _encode(frames)
    PKG = os.path.join(ROOT, "scene-text-recognition_b200")
    exe = str(tmp_path / "video_demo")
    subprocess.check_call(["g++", "-std=c++11", "-O1", os.path.join(ROOT, "tests", "cpp", "video_demo.cpp"), "-o", exe,
                           "-L", PKG, "-l:libertext.so", "-Wl,-rpath," + PKG])
    (tmp_path / "frames.jpgs").write_bytes(b"".join(jpegs))
    (tmp_path / "sizes.txt").write_text("\n".join(str(len(j)) for j in jpegs))
    assets = os.path.join(ROOT, "assets", "classifier")
    n_total = 14
    out = subprocess.run([exe, str(tmp_path / "frames.jpgs"), str(tmp_path / "sizes.txt"), "640", "480", str(n_total), "4", "2",
                          os.path.join(assets, "strong.classifier"), os.path.join(assets, "weak.classifier")],
                         capture_output=True, text=True, timeout=180)
    assert out.returncode == 0, out.stderr
    lines = [l.split() for l in out.stdout.strip().splitlines()]
    ert.enqueue_jpeg(jpegs, 640, 480, upto=ertext.STAGE_TRACK)
    ert.fetch()
    tracks, _ = ert.er_track()
    per_frame = [[tracks[f].cand[i] for i in tracks[f].tracked] for f in range(3)]
    groups = [l for l in lines if l[0] == "G"]
    assert len(groups) == n_total // 2
    for g in groups:
        gi = int(g[1])
        assert int(g[3]) == 2 * gi + 1                                  # the middle frame (n == frame_count / 2) of the group
        h = 1469598103934665603
        n = 0
        for fidx in (2 * gi, 2 * gi + 1):
            for c in per_frame[fidx % 3]:
                n += 1
                for v in (c["plane"], c["x"], c["y"], c["center_x"], c["center_y"]):
                    h = ((h ^ (int(v) & 0xffffffff)) * 1099511628211) & 0xffffffffffffffff
        assert int(g[5]) == n and int(g[7]) == h, g
    t = [l for l in lines if l[0] == "TIMES"][0]
    vals = {t[i]: float(t[i + 1]) for i in range(1, 15, 2)}
    assert int(t[-1]) == n_total and vals["extract"] > 0 and vals["nms"] > 0 and vals["classify"] > 0 and vals["track"] > 0 and vals["total"] > 0
