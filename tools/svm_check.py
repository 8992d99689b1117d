import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scene-text-recognition_b200"))
import ertext
from ertext import synth
g = np.load(os.path.join(ROOT, "tests", "golden", "ref_svm.npz"))
e = ertext.ErText(load_svm=True)
for tc in (1, 2, 0):
    e.set_svm_tensor_cores(tc)
    label, prob = e.svm_predict_probability(g["x_u8"])
    rel = np.abs(prob - g["prob"]) / np.maximum(np.abs(g["prob"]), 1e-300)
    print("tensor_cores", tc, "labels equal", bool((label == g["label"]).all()), "max rel prob err %.3e" % rel.max(), "max abs %.3e" % np.abs(prob - g["prob"]).max(), flush=True)
x = synth.svm_features_u8(7, 300)
e.set_svm_tensor_cores(1); l1, p1 = e.svm_predict_probability(x)
e.set_svm_tensor_cores(0); l0, p0 = e.svm_predict_probability(x)
print("300 vectors: labels equal", bool((l1 == l0).all()), "max rel diff tc vs fp64 %.3e" % (np.abs(p1 - p0) / np.maximum(p0, 1e-300)).max(), flush=True)
for n in (1000, 16000, 64000):
    xx = np.tile(x, (n // 300 + 1, 1))[:n]
    for tc in (1, 2, 0):
        e.set_svm_tensor_cores(tc)
        ms = e.bench_svm_u8(xx, 2)
        print("n", n, "tc", tc, "ms %.3f" % ms, "regions/s %.0f" % (n / ms * 1e3), flush=True)
