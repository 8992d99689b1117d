"""BASELINE.json configs 1, 2 and 4 as latency measurements through the host-buffer C-ABI call (wall clock around
ert_detect_classify / ert_planes_detect incl. H2D and result D2H), next to the reference's CPU code (one thread and all
threads) on the same inputs.  JSON lines on stdout."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scene-text-recognition_b200"))
import ertext
from ertext import synth
from oracle.refbind import RefOracle, PortOracle
import cv2

e = ertext.ErText()
port = PortOracle()
try:
    ref = RefOracle()
except Exception:
    ref = None
cores = len(os.sched_getaffinity(0))


def timeit(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(n):
        t = time.perf_counter(); fn(); ts.append(time.perf_counter() - t)
    return float(np.median(ts)) * 1e3


g = np.load(os.path.join(ROOT, "tests", "golden", "frames.npz"))["frames"]
# config 1: single 640x480 frame
f1 = g[0]
ms = timeit(lambda: e.detect_classify(f1))
r = e.detect_classify(f1)
out = {"config": 1, "what": "640x480 ICDAR frame (img_6), 6 planes, host buffers in/out", "gpu_ms": ms, "device_stage_ms": r.stage_ms[:6]}
if ref:
    out["ref_ms_1thread"] = ref.detect_frames(f1[None], mode=1, nthreads=1)[0] * 1e3
    out["ref_ms_6threads_reference_omp"] = ref.detect_frames(f1[None], mode=0, nthreads=6)[0] * 1e3
print(json.dumps(out), flush=True)
# config 2: single 1080p frame
f2 = synth.s_text_frame(1234)
ms = timeit(lambda: e.detect_classify(f2))
r = e.detect_classify(f2)
out = {"config": 2, "what": "1920x1080 S-text frame, 6 planes, host buffers in/out", "gpu_ms": ms, "device_stage_ms": r.stage_ms[:6]}
if ref:
    out["ref_ms_1thread"] = ref.detect_frames(f2[None], mode=1, nthreads=1)[0] * 1e3
    out["ref_ms_6threads_reference_omp"] = ref.detect_frames(f2[None], mode=0, nthreads=6)[0] * 1e3
print(json.dumps(out), flush=True)
# config 4: 4K, planes Y/Cr/Cb, scales 1, 1/2, 1/4, 1/8
f4 = synth.s_text_frame(77, 3840, 2160, n_glyphs=300)
planes = port.channels(f4)[:3]
levels = [planes] + [np.stack([cv2.resize(p, (3840 // s, 2160 // s), interpolation=cv2.INTER_LINEAR) for p in planes]) for s in (2, 4, 8)]
def run4():
    return [e.planes_detect(l) for l in levels]
ms = timeit(run4, n=5, warm=2)
rs = run4()
out = {"config": 4, "what": "3840x2160 S-text, 3 planes x scales 1,1/2,1/4,1/8 (11.0 MP per plane pyramid), host planes in", "gpu_ms": ms,
       "device_ms_per_scale": [x.stage_ms[5] for x in rs], "kept_nodes": [sum(len(p.nodes) for p in x.planes) for x in rs]}
# the same with one context per scale, enqueued asynchronously: the four levels overlap on the device
ctxs = [ertext.ErText() for _ in levels]
def run4_async():
    for c_, l in zip(ctxs, levels):
        c_.enqueue_planes(l)
    return [c_.fetch() for c_ in ctxs]
ms_async = timeit(run4_async, n=5, warm=2)
ra = run4_async()
assert all((a.planes[k].nodes == b.planes[k].nodes).all() and (a.planes[k].pool == b.planes[k].pool).all() for a, b in zip(ra, rs) for k in range(3))
out["gpu_ms_async_one_context_per_scale"] = ms_async
if ref:
    t = time.perf_counter()
    for l in levels:
        for k in range(3):
            ref.plane(l[k])
    out["ref_ms_1thread"] = (time.perf_counter() - t) * 1e3
print(json.dumps(out), flush=True)
