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
names = ["64x32x512", "64x32x256", "128x32x512", "32x32x256", "64x32x512-queue"]
only = [int(a) for a in sys.argv[1:]] or list(range(len(names)))
e.phase_cycles(True)
for cfg in only:
    e.set_tile_config(cfg)
    e.detect_classify(frames); e.phase_cycles(True)
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
    pc = e.phase_cycles(True); tot = float(sum(pc)) or 1.0
    print("   phases %: setup %.1f A %.1f B1 %.1f B2 %.1f init %.1f D %.1f C %.1f D2 %.1f D3 %.1f E %.1f" % tuple(100 * v / tot for v in pc[:9] + [0]) if False else "   phases %%: " + " ".join("%s %.1f" % (n, 100 * v / tot) for n, v in zip(["setup+TMA", "A", "B1", "B2", "-", "C(roots)", "D2", "D3", "E", "init+D"], pc[:10])))
    print("cfg %d %-11s tile %.3f ms  extract %.3f  nms %.3f  classify %.3f  total %.3f  status %d  same_as_first %s" % (
        cfg, names[cfg], best[6], best[0], best[1], best[2], best[5], r.status, same), flush=True)
