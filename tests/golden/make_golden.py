"""Generate the golden fixtures from the REFERENCE's own code (oracle/_ref, built from /root/reference).

Run here (where /root/reference exists):  python tests/golden/make_golden.py
Outputs (committed):
  tests/golden/frames.npz      3 ICDAR test frames (640x480 BGR, decoded with cv2) used as inputs
  tests/golden/ref_planes.npz  per frame x 6 planes: kept nodes (DFS pre-order, reference child order),
                               pool indices, labels, strong / weak cascade scores -- reference outputs
  tests/golden/ref_feats.npz   LBP histograms + cascade scores for crops; ARAN patches; resize cases
  tests/golden/ref_svm.npz     svm_predict_probability outputs for seeded feature vectors
"""
import os, sys
import numpy as np
import cv2

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.refbind import RefOracle  # noqa: E402

REF_IMG = "/root/reference/res/ICDAR2015_test"
names = ["img_6.jpg", "img_105.jpg", "img_28.jpg"]
frames = []
for n in names:
    im = cv2.imread(os.path.join(REF_IMG, n))
    if im.shape[:2] != (480, 640):
        im = cv2.resize(im, (640, 480))
    frames.append(im)
frames = np.stack(frames)
np.savez_compressed(os.path.join(HERE, "frames.npz"), frames=frames, names=np.array(names))

ref = RefOracle(with_svm=True)
out = {}
for f in range(frames.shape[0]):
    ch = ref.channels(frames[f])
    for k in range(6):
        r = ref.plane(ch[k], scores=True)
        for key in ("nodes", "pool", "label", "strong_score", "weak_score"):
            out["f%d_p%d_%s" % (f, k, key)] = r[key]
np.savez_compressed(os.path.join(HERE, "ref_planes.npz"), **out)

rng = np.random.RandomState(7)
ch = ref.channels(frames[1])
rects, hists, sscore, wscore, patches = [], [], [], [], []
for i in range(96):
    k = rng.randint(0, 6)
    w = rng.randint(4, 220); h = rng.randint(max(4, w // 2 + 1), min(400, w * 9))
    x = rng.randint(0, 640 - w) if w < 640 else 0
    h = min(h, 479); y = rng.randint(0, 480 - h)
    crop = ch[k][y:y + h, x:x + w]
    hist = ref.lbp_hist(crop)
    rects.append((k, x, y, w, h)); hists.append(hist); patches.append(ref.aran(crop, 26))
hists = np.array(hists)
sscore = ref.cascade_predict(0, hists); wscore = ref.cascade_predict(1, hists)
rs_cases, rs_out = [], []
for i in range(64):
    sw, sh = rng.randint(3, 90), rng.randint(3, 90)
    dw, dh = rng.randint(1, 27), rng.randint(1, 27)
    if i % 8 == 0:
        dw, dh = rng.randint(2, 20), rng.randint(2, 20); sw, sh = 2 * dw, 2 * dh
    src = rng.randint(0, 256, (sh, sw)).astype(np.uint8)
    rs_cases.append(src.ravel()); rs_out.append(ref.resize(src, dw, dh).ravel())
    rects.append((-1, sw, sh, dw, dh))
np.savez_compressed(os.path.join(HERE, "ref_feats.npz"), rects=np.array(rects[:96], np.int32), hists=hists.astype(np.uint8),
                    strong=sscore, weak=wscore, patches=np.array(patches),
                    rs_shapes=np.array(rects[96:], np.int32)[:, 1:], rs_src=np.concatenate(rs_cases), rs_out=np.concatenate(rs_out))

# SVM: seeded dense feature vectors with ~32% non-zeros, values k/255
x = np.zeros((24, 1800))
for i in range(24):
    nz = rng.rand(1800) < (0.05 + 0.5 * rng.rand())
    x[i, nz] = rng.randint(1, 256, nz.sum()) / 255.0
label, prob = ref.svm_predict_probability(x)
np.savez_compressed(os.path.join(HERE, "ref_svm.npz"), x_u8=np.rint(x * 255).astype(np.uint8), label=label, prob=prob)
print("golden fixtures written:", [f for f in os.listdir(HERE) if f.endswith(".npz")])

# larger real frames for the GPU parity tests, kept as their JPEG bytes (inputs only: the expectation is computed live by the C
# restatement on whatever cv2 decodes, and that restatement is itself pinned to the reference above):
# ICDAR img_123 (1280x960), img_1 (960x1280 portrait), img_100 (827x959, odd size)
big = {k: np.frombuffer(open(os.path.join(REF_IMG, n), "rb").read(), np.uint8) for k, n in (("landscape", "img_123.jpg"), ("portrait", "img_1.jpg"), ("odd", "img_100.jpg"))}
np.savez(os.path.join(HERE, "frames_large.npz"), names=np.array(["img_123.jpg", "img_1.jpg", "img_100.jpg"]), **big)
