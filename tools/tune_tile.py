"""Tile-shape sweep for k_tile_build on 8 synthetic 1080p frames; checks every config gives identical results."""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scene-text-recognition_b200"))
import ertext
from ertext import synth

frames = synth.s_text_batch(1234, 8)
e = ertext.ErText()
ref = None
names = ["64x32x256", "64x64x256", "64x64x512", "128x32x256", "128x32x512", "128x64x512", "64x32x128", "32x32x128"]
only = [int(a) for a in sys.argv[1:]] or list(range(len(names)))
for cfg in only:
    e.set_tile_config(cfg)
    best = None
    for it in range(4):
        r = e.detect_classify(frames)
        t = r.stage_ms
        if best is None or t[6] < best[6]:
            best = t
    sig = [(p.nodes.tobytes(), p.pool.tobytes(), p.label.tobytes()) for p in r.planes]
    if ref is None:
        ref = sig
    same = sig == ref
    print("cfg %d %-11s tile %.3f ms  extract %.3f  nms %.3f  classify %.3f  total %.3f  status %d  same_as_first %s" % (
        cfg, names[cfg], best[6], best[0], best[1], best[2], best[5], r.status, same), flush=True)
