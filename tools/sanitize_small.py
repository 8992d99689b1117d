"""Small end-to-end workload for compute-sanitizer (memcheck / initcheck): every entry point once, small inputs."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scene-text-recognition_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import ertext
from conftest import make_plane
e = ertext.ErText(load_svm=True)
g = np.load(os.path.join(ROOT, "tests", "golden", "frames.npz"))["frames"]
r = e.detect_classify(g[:1, :200, :330])
print("bgr", sum(len(p.nodes) for p in r.planes), r.status)
for kind in ("noise", "walls", "allwall", "flat"):
    r = e.planes_detect(np.stack([make_plane(1, 70, 131, kind), make_plane(2, 70, 131, "smooth")]))
    print(kind, [len(p.nodes) for p in r.planes], r.status)
e.set_seam_list(0); r = e.planes_detect(make_plane(3, 40, 70, "smooth")); e.set_seam_list(1)
pl = make_plane(4, 90, 120, "blobs")
print(e.classify_regions(pl, np.array([[3, 4, 40, 50], [10, 10, 26, 52]], np.int32))[0])
print(e.lbp_hist(pl, np.array([[0, 0, 30, 30]], np.int32)).sum())
print(e.nms_nodes(r.planes[0].nodes, 70, 40))
x = np.load(os.path.join(ROOT, "tests", "golden", "ref_svm.npz"))["x_u8"][:3]
print(e.svm_predict_probability(x)[0], e.svm_predict_probability(x.astype(np.float64) / 255.0)[0])
print(e.cascade_predict(0, np.zeros((2, 1024)))[:2], e.compute_channels(g[0, :50, :60]).shape)
# rows after the path: er_track (fused, stand-alone, caller regions) and OCR::chain_run (batch / plane, with rotation)
r = e.detect_classify(g[1:2, :240, :400], upto=ertext.STAGE_TRACK)
tr, _ = e.er_track()
print("track", len(tr[0].cand), len(tr[0].tracked))
r = e.detect_classify(g[1:2, :240, :400]); tr, _ = e.er_track()
c = tr[0].cand
if len(c):
    o = e.ocr_chain_run_batch(np.zeros(len(c), np.int32), c["plane"], np.stack([c["x"], c["y"], c["w"], c["h"]], axis=1),
                              np.where(np.arange(len(c)) % 2 == 0, 0.0, 0.3))
    print("ocr batch", o.label[:4], o.feat.sum())
ft = e.er_track_regions(g[1, :200, :300], [(0, 10, 10, 30, 40, 500), (3, 50, 60, 20, 25, 300)], [(1, 12, 14, 28, 38, 450), (4, 100, 20, 1, 1, 130)])
print("track regions", ft.tracked)
o = e.ocr_chain_run_plane(pl, [(0, 0, 120, 90), (5, 5, 60, 60), (7, 9, 3, 40), (20, 20, 40, 30)], [0.0, 0.0, 0.0, -0.5])
print("ocr plane", o.label, o.img.sum())
print(e.ocr_features_plane(pl, [(1, 1, 2, 2)]).feat.sum())
# pyramid level, SVM batch with a partial tile, JPEG ingest (when an encoder is importable)
x = np.load(os.path.join(ROOT, "tests", "golden", "ref_svm.npz"))["x_u8"]
print("svm 24", e.svm_predict_probability(np.tile(x, (8, 1))[:150])[0][:3])
try:
    import cv2
    jp = [cv2.imencode(".jpg", g[0, :160, :240])[1].tobytes()]
    e.enqueue_jpeg(jp, 240, 160); rj = e.fetch()
    print("jpeg", e.jpeg_backend_name(), sum(len(p.nodes) for p in rj.planes))
except ImportError:
    print("jpeg skipped (no cv2)")
e.close()
print("done")
