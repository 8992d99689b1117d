"""JPEG ingest timing: 8 x 1080p per step, nvJPEG decode on the device + the hot path, depth contexts round-robin."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scene-text-recognition_b200"))
import cv2
import ertext
from ertext import synth
W, H, F = 1920, 1080, 8
frames = [synth.s_text_frame(1234 + i, W, H) for i in range(F)]
jpegs = [cv2.imencode(".jpg", f, [int(cv2.IMWRITE_JPEG_QUALITY), 90])[1].tobytes() for f in frames]
print("jpeg bytes per frame", sum(map(len, jpegs)) // F)
import threading
def run(backend, depth, fpb, steps_per_thread):
    batch = (jpegs * ((fpb + F - 1) // F))[:fpb]
    ctxs = [ertext.ErText(device=0) for _ in range(depth)]
    try:
        for c in ctxs:
            c.set_jpeg_backend(backend, 4)
            c.enqueue_jpeg(batch, W, H); c.fetch()
        def worker(c):
            for _ in range(steps_per_thread):
                c.enqueue_jpeg(batch, W, H); c.fetch()
        th = [threading.Thread(target=worker, args=(c,)) for c in ctxs]
        t0 = time.perf_counter()
        for t in th: t.start()
        for t in th: t.join()
        dt = time.perf_counter() - t0
        print("backend", backend, ctxs[0].jpeg_backend_name(), "threads", depth, "frames/batch", fpb, "fps %.0f" % (depth * steps_per_thread * fpb / dt),
              "decode_ms %.3f" % ctxs[0].jpeg_decode_ms(), flush=True)
    except Exception as ex:
        print("backend", backend, "threads", depth, "fpb", fpb, "failed:", ex, flush=True)
    for c in ctxs: c.close()

for depth in (1, 4, 8, 16):
    run(-1, depth, 8, 6)
run(2, 1, 64, 3)
run(2, 2, 64, 3)
run(3, 1, 8, 2)
