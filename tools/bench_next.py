"""Measurement of the rows after the detect path on the bench workload: 8 synthetic 1080p S-text frames.
  er_track  : device ms per batch (CUDA events around gather + calc_color + track + emit), regions per batch,
              next to the reference's er_track on one host thread (oracle/_ref, same region lists)
  chain_run : OCR feature path + SVM, regions/s device-timed for N regions (the batch's tracked regions, tiled),
              next to the reference's OCR::chain_run on one host thread (bounded sample)
Writes JSON lines to stdout.  Usage: python tools/bench_next.py"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scene-text-recognition_b200"))
import ertext
from ertext import synth
from oracle.refbind import RefOracle

F, W, H = 8, 1920, 1080
frames = synth.s_text_batch(1234, F, W, H)
e = ertext.ErText(load_svm=True)
res = e.detect_classify(frames, upto=ertext.STAGE_TRACK)
ms_list = []
for _ in range(6):
    e.detect_classify(frames, upto=ertext.STAGE_TRACK)
    tracks, ms = e.er_track()
    ms_list.append(ms)
ncand = sum(len(t.cand) for t in tracks); ntr = sum(len(t.tracked) for t in tracks)
# algorithmic bytes of calc_color: each candidate's bound is read twice from its channel (histogram, mask) and the three
# colour planes once under the mask
px = int(sum(int(c["w"]) * int(c["h"]) for t in tracks for c in t.cand))
line = {"row": "er_track", "frames": F, "candidates_per_batch": ncand, "tracked_per_batch": ntr, "device_ms_per_batch": float(np.median(ms_list[1:])),
        "bound_pixels": px, "algorithmic_bytes": 5 * px, "achieved_GBps": 5 * px / (np.median(ms_list[1:]) * 1e-3) / 1e9}
try:
    ref = RefOracle(with_svm=True)
except Exception:
    ref = None
if ref is not None:
    t_cpu = 0.0
    for f in range(F):
        ch = ref.channels(frames[f]); ycc = np.stack([ch[0], ch[1], ch[2]], axis=-1)
        ns = tracks[f].n_strong; c = tracks[f].cand
        rows = np.stack([c["plane"], c["x"], c["y"], c["w"], c["h"], c["area"]], axis=1).astype(np.int32).reshape(-1, 6)
        t0 = time.perf_counter(); r = ref.er_track(ch, ycc, rows[:ns], rows[ns:]); t_cpu += time.perf_counter() - t0
        want = np.array([i if k == 0 else ns + i for k, i in r["tracked"]], np.int32)
        assert (want == tracks[f].tracked).all()
    line["reference_cpu_ms_per_batch_1thread"] = t_cpu * 1e3
print(json.dumps(line), flush=True)

# OCR::chain_run on the tracked regions of the batch (device-resident planes)
fr, pl, rc = [], [], []
for f, t in enumerate(tracks):
    for i in t.tracked:
        c = t.cand[i]
        fr.append(f); pl.append(int(c["plane"])); rc.append((int(c["x"]), int(c["y"]), int(c["w"]), int(c["h"])))
fr, pl, rc = np.array(fr, np.int32), np.array(pl, np.int32), np.array(rc, np.int32).reshape(-1, 4)
base_n = len(fr)
for n in (base_n, 1024, 4096, 16384):
    rep = (n + base_n - 1) // base_n
    f2, p2, r2 = np.tile(fr, rep)[:n], np.tile(pl, rep)[:n], np.tile(rc, (rep, 1))[:n]
    sl = np.where(np.arange(n) % 4 == 3, 0.15, 0.0)       # every 4th region takes the rotate_mat route
    ms = []
    for _ in range(4):
        r = e.ocr_chain_run_batch(f2, p2, r2, sl)
        ms.append(r.ocr_ms)
    out = {"row": "chain_run", "n": int(n), "device_ms": float(np.median(ms[1:])), "regions_per_s": n / (np.median(ms[1:]) * 1e-3),
           "note": "feature kernel + SVM (tcgen05 distances), planes resident; every 4th region rotated"}
    if ref is not None and n == base_n:
        chans = [ref.channels(frames[f]) for f in range(F)]
        m = min(n, 48)
        t0 = time.perf_counter()
        vals = [ref.chain_run(chans[f2[i]][p2[i]][r2[i][1]:r2[i][1] + r2[i][3], r2[i][0]:r2[i][0] + r2[i][2]], 0, float(sl[i])) for i in range(m)]
        dt = time.perf_counter() - t0
        out["reference_cpu_regions_per_s_1thread"] = m / dt
        assert (np.floor(vals) == np.floor(r.value[:m])).all()
        out["max_rel_prob_err_vs_reference"] = float(np.max(np.abs((np.array(vals) - np.floor(vals)) - (r.value[:m] - np.floor(r.value[:m]))) / (np.array(vals) - np.floor(vals))))
    print(json.dumps(out), flush=True)
