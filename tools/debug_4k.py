import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scene-text-recognition_b200")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import ertext
from ertext import synth
from oracle.refbind import PortOracle
from gpu_check import cmp_plane
e = ertext.ErText(); port = PortOracle()
frame = synth.s_text_frame(77, 3840, 2160, n_glyphs=300)
planes = port.channels(frame)[:3]
exp = port.plane(planes[2], scores=True, canonical_order=True)
expu = port.plane(planes[2], scores=True, canonical_order=False)
print("oracle nodes", exp["nodes"].shape, "pool", len(exp["pool"]))
sigs = []
for it in range(4):
    for lu in (1, 0):
        e.set_tile_local_union(lu)
        got = e.planes_detect(planes[2]).planes[0]
        ok = cmp_plane("it%d lu%d" % (it, lu), got, exp)
        sigs.append(got.nodes.tobytes())
        gs = sorted(map(tuple, got.nodes[:, :6])); es = sorted(map(tuple, exp["nodes"][:, :6]))
        print("it", it, "lu", lu, "ok", ok, "multiset equal", gs == es)
        if not ok and gs == es:
            bad = np.nonzero((got.nodes != exp["nodes"]).any(1))[0]
            print("  differing rows", bad[:10])
            for b in bad[:6]:
                print("   got", got.nodes[b], "exp", exp["nodes"][b])
print("deterministic across runs:", len(set(sigs)) == 1)
