#!/usr/bin/env python
"""bench.py -- 1080p frames/s of the ER detect+classify path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU code (oracle/_ref)

A "step" is one pass of the hot path (compute_channels -> er_tree_extract -> non_maximum_supression ->
classify, 6 planes per frame) over one batch of `--frames-per-gpu` synthetic 1080p S-text frames per GPU.
  value : frames/s with the BGR frames already resident in HBM (device-timed with CUDA events)
  e2e   : frames/s through the host-buffer C-ABI call (pinned host frames -> H2D -> kernels -> result D2H
          inside the timed region), several contexts (streams) used round-robin so copies and the narrow
          kernels of one batch overlap the tile kernel of another
Weak scaling: every rank processes its own `frames-per-gpu` frames per step (frames are independent units,
no collective on the compute path); for N > 1 the per-step region records are all-gathered over NCCL.
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "scene-text-recognition_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

_JSON_OUT = None


def emit(line):
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


METRIC = "1080p frames/sec (ER detect+classify)"
UNIT = "frames/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames-per-gpu", type=int, default=8)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--input-batches", type=int, default=4, help="distinct input batches rotated through (defeats L2 reuse)")
    ap.add_argument("--cpu-sample-frames", type=int, default=32)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-next-rows", action="store_true", help="skip the er_track / chain_run leg (SURVEY 8f rows, outside the timed region)")
    ap.add_argument("--no-tile-fifo", action="store_true", help="A/B: do not chain the tile kernels of the contexts in submission order")
    ap.add_argument("--no-stream-split", action="store_true", help="A/B: run every stage of a batch on ONE stream (no high-priority post stream)")
    ap.add_argument("--contexts", type=int, default=5, help="contexts / streams used round-robin (copy/compute overlap)")
    return ap.parse_args()


def make_frames(first_seed, n, w, h):
    from ertext import synth
    return synth.s_text_batch(first_seed, n, w, h)


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


class ClockSampler:
    """nvidia-smi sampled every 200 ms DURING the timed region (B200_PROFILING.md clocks line)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.path = "/tmp/ert_clocks_%d_%d.csv" % (os.getpid(), gpu_index)
        self.proc = None
        self.idx = gpu_index

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.remove(self.path)
        except OSError:
            pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic_bytes():
    """dram bytes per launch of the tile-build kernel from the committed ncu capture, if any (profiles/*.json)."""
    p = os.path.join(ROOT, "profiles", "tile_build_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get("dram_bytes_per_launch")
        except Exception:
            return None
    return None


# ---------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU code (oracle/_ref), all host threads, bounded sample per step
# ---------------------------------------------------------------------------------------------
def run_reference(a, rank):
    if rank != 0:
        return
    from oracle.refbind import RefOracle, PortOracle
    cores = host_cores()
    try:
        ref = RefOracle()
        kind = "reference"
    except (FileNotFoundError, OSError):
        ref = None
        kind = "port"
    fpg = a.frames_per_gpu
    base = make_frames(1234, fpg, a.width, a.height)
    # a step of the CPU arm covers at least one frame per host thread so that every core has work
    # (same frames, repeated; the metric is frames/s either way)
    n_step = max(fpg, min(cores, 16 * fpg))
    frames = np.stack([base[i % fpg] for i in range(n_step)])
    fpg = n_step
    if ref is None:
        emit({"impl": "reference", "unavailable": "oracle/_ref/libref_oracle.so not present and the C port has no frame driver"})
        return
    for _ in range(a.warmup):
        ref.detect_frames(frames[: max(1, min(fpg, cores))], mode=1, nthreads=cores)
    t = 0.0
    counts = None
    for _ in range(a.steps):
        sec, counts, _st = ref.detect_frames(frames, mode=1, nthreads=cores)
        t += sec
    fps = fpg * a.steps / t
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": 1e3 * t / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/int32 (+f64 cascade sums)",
        "data": "synthetic",
        "config": {"workload": "%dx%d S-text frames (the same seeds as this repo's arm), 6 planes native scale, %d frames per step (>= one per host thread)" % (a.width, a.height, fpg),
                   "threads": cores, "mode": "throughput-fair: frames over all host threads, planes sequential per frame"},
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": "%d steps x %d frames %dx%d, ERFilter per-channel loop (src/ER.cpp:50-60) verbatim" % (a.steps, fpg, a.width, a.height)},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "regions_per_frame": float(counts[:, 1].mean()) if counts is not None else None,
    }
    emit(line)


# ---------------------------------------------------------------------------------------------
# this repo's arm
# ---------------------------------------------------------------------------------------------
def run_ours(a, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    import ertext
    from ertext import dist as edist

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    W, H, fpg, NB = a.width, a.height, a.frames_per_gpu, a.input_batches
    # global frame g of batch b has seed 1234 + 1000*b + g ; rank r owns g = r, r+world, ...
    my_ids = [rank + world * i for i in range(fpg)]
    host_batches = []
    for b in range(NB):
        fr = np.stack([make_frames(1234 + 1000 * b + g, 1, W, H)[0] for g in my_ids])
        host_batches.append(torch.from_numpy(fr).pin_memory())
    dev_batches = [hb.to(dev) for hb in host_batches]
    torch.cuda.synchronize()

    NC = max(1, a.contexts)
    ctxs = [ertext.ErText(device=local_rank) for _ in range(NC)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(NC)]
    for c, s in zip(ctxs, streams):
        c.set_stream(s.cuda_stream)
        if a.no_tile_fifo:
            c.set_tile_fifo(False)
        if a.no_stream_split:
            c.set_stream_split(False)

    gatherer = edist.RegionGatherer(dev) if world > 1 else None
    stats = {"tile_ms": [], "extract_ms": [], "nms_ms": [], "classify_ms": [], "launches": 0, "d2h": 0, "regions": 0, "kept": 0, "steps": 0}

    def collect(c, record, do_gather):
        r = c.fetch()
        if r.status:
            raise RuntimeError("device status %d (%s)" % (r.status, c.L.ert_status_string(r.status).decode()))
        if record:
            stats["tile_ms"].append(r.stage_ms[6]); stats["extract_ms"].append(r.stage_ms[0])
            stats["nms_ms"].append(r.stage_ms[1]); stats["classify_ms"].append(r.stage_ms[2])
            stats["launches"] += c.launch_count()
            nk = sum(len(p.nodes) for p in r.planes); npool = sum(len(p.pool) for p in r.planes)
            stats["d2h"] += nk * 32 + npool * 24 + 2 * 4 * (len(r.planes) + 1) + 4
            stats["regions"] += npool; stats["kept"] += nk; stats["steps"] += 1
        if do_gather:
            gatherer.submit(edist.pack_records(r, my_ids))   # asynchronous NCCL all_gather, collected 2 steps later
        return r

    def run_loop(n_steps, resident, record):
        pending = [False] * NC
        for i in range(n_steps):
            k = i % NC
            if pending[k]:
                collect(ctxs[k], record, world > 1)
            if resident:
                ctxs[k].enqueue_device(dev_batches[i % NB].data_ptr(), fpg, W, H, W * 3)
            else:
                ctxs[k].enqueue_host(host_batches[i % NB].data_ptr(), fpg, W, H, W * 3)
            pending[k] = True
        for j in range(NC):
            k = (n_steps + j) % NC
            if pending[k]:
                collect(ctxs[k], record, world > 1)
        if world > 1:
            gathered = gatherer.drain()          # the final region gather completes inside the timed region
            if rank == 0 and record:
                stats["gathered_rows"] = stats.get("gathered_rows", 0) + int(sum(len(g) for g in gathered))

    def timed(resident):
        # every context allocates its workspace on first use: W warm-up steps, but at least one per context
        run_loop(max(a.warmup, NC), resident, False)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        start = torch.cuda.Event(enable_timing=True)
        ends = [torch.cuda.Event(enable_timing=True) for _ in range(NC)]
        start.record(streams[0])
        for s_ in streams[1:]:
            s_.wait_event(start)
        t0 = time.perf_counter()
        run_loop(a.steps, resident, True)
        for e, s in zip(ends, streams):
            e.record(s)
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        if world > 1:
            dist.barrier()
        ms = max(start.elapsed_time(e) for e in ends)
        # the host-side result collection of the last batches happens after the last kernel: the step ends when the
        # result is in host memory, so take the larger of the device bracket and the host bracket
        ms = max(ms, wall) if not resident else ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_res = timed(True)
    res_stats = dict(stats)
    for k in ("tile_ms", "extract_ms", "nms_ms", "classify_ms"):
        res_stats[k] = list(stats[k])
    for k in stats:
        stats[k] = [] if isinstance(stats[k], list) else 0
    ms_e2e = timed(False)
    clocks = sampler.stop() if rank == 0 else None

    # the dominant kernel timed ALONE (one context, nothing else in flight): CUDA events recorded by the library
    # around the k_tile_build launch on its launching stream; this is the roofline numerator's time base
    solo_tile, solo_total = [], []
    for i in range(6):
        ctxs[0].enqueue_device(dev_batches[i % NB].data_ptr(), fpg, W, H, W * 3)
        r = ctxs[0].fetch()
        if i:
            solo_tile.append(r.stage_ms[6]); solo_total.append(r.stage_ms[5])

    frames_total = fpg * world * a.steps
    value = frames_total / (ms_res * 1e-3)
    e2e = frames_total / (ms_e2e * 1e-3)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peaks()
    tile_ms = float(np.median(solo_tile))
    tile_ms_in_flight = float(np.mean(res_stats["tile_ms"]))
    alg_bytes = W * H * 6 * fpg                      # one u8 read per pixel per plane (B_extract = W*H per plane)
    achieved = alg_bytes / (tile_ms * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms_res / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8/int32 (+f64 cascade sums)", "data": "synthetic",
        "config": {"workload": "%dx%d S-text synthetic frames (seeds 1234+), 6 planes native scale, %d frames per GPU per step" % (W, H, fpg),
                   "global_frames_per_step": fpg * world, "parallelism": "dp%d (frames sharded, no data-path collective%s)" % (world, "; NCCL all_gather of region records per step" if world > 1 else ""),
                   "l2": "inputs rotate over %d distinct batches (%d x %.1f MB > 126 MB L2); all workspaces rewritten every step" % (NB, NB, fpg * W * H * 3 / 1e6),
                   "pipelining": "%d contexts / streams used round-robin" % NC, "params": "THRESH_STEP 8, MIN_AREA 120, MAX_AREA 900000, STABILITY_T 2, OVERLAP 0.7"},
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": fpg * W * H * 3, "d2h_bytes_per_step": int(stats["d2h"] / max(stats["steps"], 1)),
                "ms_per_step": ms_e2e / a.steps},
        "gpu_launches": int(res_stats["launches"]),
        "roofline": {"bound": "hbm", "kernel": "k_tile_build<64,32,512>", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": ncu_traffic_bytes(), "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": tile_ms, "peak_source": peak_src,
                     "share_of_step": tile_ms / float(np.median(solo_total)), "kernel_ms_with_3_batches_in_flight": tile_ms_in_flight,
                     "timing": "CUDA events around the launch on its stream, batch processed alone (median of 5)"},
        "stage_ms_per_batch": {"extract": float(np.mean(res_stats["extract_ms"])), "tile_build": tile_ms, "nms": float(np.mean(res_stats["nms_ms"])),
                               "classify": float(np.mean(res_stats["classify_ms"]))},
        "clocks": clocks,
        "regions_per_frame": res_stats["regions"] / max(res_stats["steps"] * fpg, 1),
        "kept_nodes_per_frame": res_stats["kept"] / max(res_stats["steps"] * fpg, 1),
    }
    if world == 1 and not a.no_next_rows:
        line["next_rows"] = next_rows(a, ctxs[0], dev_batches, fpg, W, H)
    if world == 1 and not a.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(a, host_batches)
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def next_rows(a, ctx, dev_batches, fpg, W, H):
    """The two steps after the path, measured OUTSIDE the timed region on one context, device events:
    the same batch with er_track fused into the submission (upto = ERT_STAGE_TRACK), then OCR::chain_run
    (feature kernel + SVM) on every tracked region of the batch, planes still resident."""
    import ertext
    try:
        ctx.load_svm(ertext.svm_model_path())
        tr_ms, ocr_ms, tot_ms, n_tr = [], [], [], 0
        for i in range(5):
            ctx.enqueue_device(dev_batches[i % len(dev_batches)].data_ptr(), fpg, W, H, W * 3, upto=ertext.STAGE_TRACK)
            r = ctx.fetch()
            tracks, ms = ctx.er_track()
            fr, pl, rc = [], [], []
            for f, t in enumerate(tracks):
                c = t.cand[t.tracked]
                fr += [f] * len(c); pl += c["plane"].tolist(); rc += list(zip(c["x"].tolist(), c["y"].tolist(), c["w"].tolist(), c["h"].tolist()))
            o = ctx.ocr_chain_run_batch(np.array(fr, np.int32), np.array(pl, np.int32), np.array(rc, np.int32).reshape(-1, 4))
            if i:
                tr_ms.append(ms); ocr_ms.append(o.ocr_ms); tot_ms.append(r.stage_ms[5] + ms + o.ocr_ms); n_tr = len(fr)
        return {"er_track_ms_per_batch": float(np.median(tr_ms)), "chain_run_ms_per_batch": float(np.median(ocr_ms)),
                "tracked_regions_per_batch": n_tr, "serial_ms_per_batch_detect_to_letters": float(np.median(tot_ms)),
                "frames_per_s_detect_to_letters_one_context": fpg / (float(np.median(tot_ms)) * 1e-3),
                "note": "device time, one context, no overlap; er_grouping (CPU, between the two) not included, slope 0"}
    except Exception as ex:    # never lose the headline line to the extra leg
        return {"error": str(ex)}


def cpu_baseline(a, host_batches):
    """The reference's own CPU code (oracle/_ref) on a bounded sample of the same frames, all host cores."""
    from oracle.refbind import RefOracle
    cores = host_cores()
    try:
        ref = RefOracle()
    except (FileNotFoundError, OSError) as ex:
        return {"value": None, "unit": UNIT, "cores": cores, "kind": "reference", "sample": "unavailable: %s" % ex}
    frames = np.concatenate([hb.numpy() for hb in host_batches])[: a.cpu_sample_frames]
    ref.detect_frames(frames[: min(len(frames), cores)], mode=1, nthreads=cores)   # warm-up
    sec, counts, st = ref.detect_frames(frames, mode=1, nthreads=cores)
    sec1, _, _ = ref.detect_frames(frames[:2], mode=1, nthreads=1)
    return {"value": len(frames) / sec, "unit": UNIT, "cores": cores, "kind": "reference",
            "sample": "%d of the benchmark's frames, frames spread over %d threads (1 thread: %.2f fps); stage share extract/nms/classify = %.0f/%.0f/%.0f %%" % (
                len(frames), cores, 2 / sec1, 100 * st[0] / st.sum(), 100 * st[1] / st.sum(), 100 * st[2] / st.sum())}


def main():
    # keep stdout to the ONE JSON line: libraries that write to fd 1 (NCCL prints its version banner there)
    # are pointed at stderr; the JSON line goes to a private duplicate of the original stdout
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if a.impl == "reference":
        run_reference(a, rank)
        return
    if world != a.gpus and world == 1 and a.gpus > 1:
        emit({"error": "launch with torchrun --nproc-per-node %d for --gpus %d" % (a.gpus, a.gpus)})
        sys.exit(2)
    run_ours(a, rank, local_rank, world)


if __name__ == "__main__":
    main()
