"""BASELINE config 5: ER-candidate sweep 1k..500k regions -- AdaBoost cascades and SVM batch scoring.
Device-timed (CUDA events inside the library, inputs resident), next to the reference's own CPU code on a
bounded sample.  Writes JSON lines to stdout.  Usage: python tools/sweep_classifiers.py [--max N]"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scene-text-recognition_b200"))
import ertext
from ertext import synth
from oracle.refbind import RefOracle, PortOracle

maxn = int(sys.argv[sys.argv.index("--max") + 1]) if "--max" in sys.argv else 500000
sizes = [n for n in (1000, 4000, 16000, 64000, 256000, 500000) if n <= maxn]
e = ertext.ErText(load_svm=True)
port = PortOracle()
cores = len(os.sched_getaffinity(0))

# realistic histograms: LBP features of random crops of a synthetic frame (stage pass rates as in the pipeline)
frame = synth.s_text_frame(1234)
plane = port.channels(frame)[0]
rng = np.random.RandomState(5)
rects = []
for _ in range(4096):
    w = rng.randint(13, 200); h = rng.randint(max(13, w // 2 + 1), min(400, w * 5))
    rects.append((rng.randint(0, 1920 - w), rng.randint(0, 1080 - h), w, h))
rects = np.array(rects, np.int32)
_, _, _, base_hist = e.classify_regions(plane, rects, want_hist=True)
lab, ss, ws = e.cascade_classify_u8(base_hist)
pass_rate = {"strong": float((lab == 2).mean()), "weak": float((lab == 1).mean())}

try:
    ref = RefOracle(with_svm=True)
except Exception:
    ref = None
cpu = {}
if ref is not None:
    fv = base_hist[:2048].astype(np.float64)
    t = time.time(); ref.cascade_predict(0, fv); ref.cascade_predict(1, fv); dt = time.time() - t
    cpu["cascade_regions_per_s_1thread"] = 2048 / dt
    xs = synth.svm_features_u8(9, 2 * cores).astype(np.float64) / 255.0
    t = time.time(); ref.svm_predict_probability(xs, nthreads=cores); dt = time.time() - t
    cpu["svm_regions_per_s_%dthreads" % cores] = len(xs) / dt
    t = time.time(); ref.svm_predict_probability(xs[:4], nthreads=1); dt = time.time() - t
    cpu["svm_regions_per_s_1thread"] = 4 / dt
print(json.dumps({"cpu_reference": cpu, "cores": cores, "cascade_pass_rate": pass_rate}), flush=True)

for n in sizes:
    hist = np.tile(base_hist, (n // len(base_hist) + 1, 1))[:n]
    ms = e.bench_cascade_u8(hist, 5 if n <= 64000 else 2)
    hist2 = synth.lbp_like_hist_u8(3, min(n, 16000)); hist2 = np.tile(hist2, (n // len(hist2) + 1, 1))[:n]
    ms2 = e.bench_cascade_u8(hist2, 5 if n <= 64000 else 2)
    print(json.dumps({"kernel": "k_cascade<u8>", "n": n, "ms": ms, "regions_per_s": n / ms * 1e3, "ms_all_reject": ms2,
                      "regions_per_s_all_reject": n / ms2 * 1e3, "bytes_per_region": 1024,
                      "achieved_GBps": n * 1024 / ms / 1e6}), flush=True)
for n in sizes:
    x = synth.svm_features_u8(7, min(n, 8000)); x = np.tile(x, (n // len(x) + 1, 1))[:n]
    ms = e.bench_svm_u8(x, 2 if n <= 64000 else 1)
    ops = 2.0 * 1800 * 1910 * n
    print(json.dumps({"kernel": "k_svm_kvalue<u8>+k_svm_prob", "n": n, "ms": ms, "regions_per_s": n / ms * 1e3,
                      "distance_gemm_TFLOPs_equiv": ops / ms / 1e9}), flush=True)
