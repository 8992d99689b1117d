"""One scoring pass of n vectors (default 16000) for ncu: `ncu --set full --import-source on -k regex:k_svm_ -s 4 -c 4 python tools/svm_profile.py`
(the first pass is the warm-up inside ert_bench_svm_u8; -s 4 skips its four kernels)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scene-text-recognition_b200"))
import ertext
from ertext import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16000
x = synth.svm_features_u8(7, 300)
xx = np.tile(x, (n // 300 + 1, 1))[:n]
e = ertext.ErText(load_svm=True)
print("ms", e.bench_svm_u8(xx, 1))
