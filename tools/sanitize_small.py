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
e.set_tile_local_union(0); r = e.planes_detect(make_plane(3, 40, 70, "smooth")); e.set_tile_local_union(1)
pl = make_plane(4, 90, 120, "blobs")
print(e.classify_regions(pl, np.array([[3, 4, 40, 50], [10, 10, 26, 52]], np.int32))[0])
print(e.lbp_hist(pl, np.array([[0, 0, 30, 30]], np.int32)).sum())
print(e.nms_nodes(r.planes[0].nodes, 70, 40))
x = np.load(os.path.join(ROOT, "tests", "golden", "ref_svm.npz"))["x_u8"][:3]
print(e.svm_predict_probability(x)[0], e.svm_predict_probability(x.astype(np.float64) / 255.0)[0])
print(e.cascade_predict(0, np.zeros((2, 1024)))[:2], e.compute_channels(g[0, :50, :60]).shape)
e.close()
print("done")
