#!/usr/bin/env python
"""bench.py -- 1080p frames/s of the ER detect+classify path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (config 3, the metric's configuration)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU code (oracle/_ref)
  python bench.py --config {1,2,4,5}                       # the other BASELINE.json configurations, same JSON keys

A "step" is one pass of the hot path (compute_channels -> er_tree_extract -> non_maximum_supression ->
classify, 6 planes per frame) over one batch of `--frames-per-gpu` synthetic 1080p S-text frames per GPU.
  value : frames/s with the BGR frames already resident in HBM (device-timed with CUDA events)
  e2e   : frames/s through the host-buffer C-ABI call (pinned host frames -> H2D -> kernels -> result D2H
          inside the timed region), several contexts (streams) used round-robin so copies and the narrow
          kernels of one batch overlap the tile kernel of another
Weak scaling: every rank processes its own `frames-per-gpu` frames per step (frames are independent units,
no collective on the compute path); for N > 1 the labelled regions of every step are gathered on all ranks by the
library itself (ert_gather_regions_*: packed on the device, NCCL inside libertext.so, pipelined behind the data path).
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
PARAMS = "THRESH_STEP 8, MIN_AREA 120, MAX_AREA 900000, STABILITY_T 2, OVERLAP 0.7"
DTYPE = "u8/int32 (+f64 cascade sums)"


def workload_text(W, H, fpg):
    # the SAME wording in both arms (the driver compares the two lines' config.workload)
    return "%dx%d S-text synthetic frames (seeds 1234+), 6 planes native scale, %d frames per GPU per step" % (W, H, fpg)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=3, choices=[1, 2, 3, 4, 5],
                    help="BASELINE.json configuration: 1 = 640x480 frame, 2 = 1080p frame + pyramid, 3 = 1080p stream (the metric), "
                         "4 = 4K 3 planes x 4 scales, 5 = classifier sweep")
    ap.add_argument("--frames-per-gpu", type=int, default=8)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--input-batches", type=int, default=4, help="distinct input batches rotated through (defeats L2 reuse)")
    ap.add_argument("--cpu-sample-frames", type=int, default=32)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-next-rows", action="store_true", help="skip the er_track / chain_run leg (SURVEY 8f rows, outside the timed region)")
    ap.add_argument("--tile-fifo", action="store_true", help="A/B: chain the tile kernels of the contexts in submission order (round-1 default)")
    ap.add_argument("--no-stream-split", action="store_true", help="A/B: run every stage of a batch on ONE stream (no high-priority post stream)")
    ap.add_argument("--no-seam-list", action="store_true", help="A/B: the round-1 seam kernel (one thread per seam position)")
    ap.add_argument("--tile-config", type=int, default=0, help="A/B: variant of the tile kernel (ert_set_tile_config)")
    ap.add_argument("--upto", type=int, default=3, help="diagnostic: stop after stage 1 = extract, 2 = NMS, 3 = classify (the metric; default)")
    ap.add_argument("--sustained-seconds", type=float, default=2.5, help="extra leg of at least this many seconds (clock record, steady-state check)")
    ap.add_argument("--no-numa", action="store_true", help="do not pin the rank to the CPUs of its GPU's NUMA node")
    ap.add_argument("--parity-frames", type=int, default=2, help="frames of one step checked against the oracles outside the timed region (0 = skip)")
    ap.add_argument("--jpeg-seconds", type=float, default=2.0, help="decode-inclusive leg: JPEG bitstreams in, nvJPEG on the device (0 = skip)")
    ap.add_argument("--jpeg-threads", type=int, default=6, help="host threads (one context each) feeding the decode-inclusive leg")
    ap.add_argument("--post-footprint", type=int, default=-1, help="A/B: CTAs per SM for the post-tile kernels (ert_set_post_footprint; -1 = library default)")
    ap.add_argument("--no-gather", action="store_true", help="A/B at N > 1: leave the labelled regions on their rank (no ert_gather_regions_*)")
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


def pin_to_gpu_numa_node(local_rank):
    """Threads and (first-touch) pinned buffers of a rank on the NUMA node its GPU hangs off.  Returns what was done."""
    try:
        q = subprocess.run(["nvidia-smi", "-i", str(local_rank), "--query-gpu=pci.bus_id", "--format=csv,noheader"], capture_output=True, text=True, timeout=20)
        bus = q.stdout.strip().lower()
        if bus.startswith("0000"):
            bus = bus[4:]                          # sysfs uses a 4-digit domain
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read().strip())
        if node < 0:
            return {"gpu_numa_node": node, "pinned": False, "why": "the platform reports no NUMA affinity for the GPU"}
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & set(os.sched_getaffinity(0))
        if not allowed:
            return {"gpu_numa_node": node, "pinned": False, "why": "no allowed CPU on that node"}
        os.sched_setaffinity(0, allowed)
        return {"gpu_numa_node": node, "pinned": True, "cpus": len(allowed)}
    except Exception as ex:
        return {"gpu_numa_node": None, "pinned": False, "why": str(ex)[:120]}


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


def ncu_traffic():
    """dram bytes per launch of the tile-build kernel from the COMMITTED ncu capture (profiles/tile_build_traffic.json):
    a constant of the profile named there, not something this run measures."""
    p = os.path.join(ROOT, "profiles", "tile_build_traffic.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return d.get("dram_bytes_per_launch"), "%s -- ncu --set full capture of this command, committed under profiles/; NOT measured by this run" % d.get("source", "profiles/tile_build_traffic.json")
        except Exception:
            pass
    return None, None


# ---------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU code (oracle/_ref), all host threads, the same frames per step
# ---------------------------------------------------------------------------------------------
def run_reference(a, rank):
    if rank != 0:
        return
    from oracle.refbind import RefOracle
    cores = host_cores()
    try:
        ref = RefOracle()
    except (FileNotFoundError, OSError):
        emit({"impl": "reference", "unavailable": "oracle/_ref/libref_oracle.so not present and the C port has no frame driver"})
        return
    if a.config != 3:
        emit({"impl": "reference", "unavailable": "the reference arm times BASELINE config 3 (the metric); configs 1/2/4/5 carry their reference timings in cpu_baseline of the --config line"})
        return
    fpg = a.frames_per_gpu
    n_step = fpg * max(1, a.gpus)                # the same frames per step as this repo's arm (global batch)
    base = make_frames(1234, min(n_step, 16), a.width, a.height)
    frames = base if len(base) == n_step else np.stack([base[i % len(base)] for i in range(n_step)])
    for _ in range(a.warmup):
        ref.detect_frames(frames[: max(1, min(n_step, 4))], mode=2, nthreads=cores)
    t = 0.0
    counts = None
    for _ in range(a.steps):
        sec, counts, _st = ref.detect_frames(frames, mode=2, nthreads=cores)
        t += sec
    fps = n_step * a.steps / t
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": 1e3 * t / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": DTYPE,
        "data": "synthetic",
        "config": {"workload": workload_text(a.width, a.height, fpg), "global_frames_per_step": n_step, "params": PARAMS,
                   "threads": cores, "mode": "throughput-fair: the (frame, plane) units of a step spread over all host threads (the reference's own "
                                            "6-thread omp loop over planes is slower: cpu_baseline.mode_R of this repo's line)"},
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": "reference", "per_thread": fps / cores,
                         "sample": "%d steps x %d frames %dx%d, ERFilter per-channel loop (src/ER.cpp:50-60) verbatim" % (a.steps, n_step, a.width, a.height)},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "regions_per_frame": float(counts[:, 1].mean()) if counts is not None else None,
    }
    emit(line)


# ---------------------------------------------------------------------------------------------
# this repo's arm, config 3: the 1080p stream
# ---------------------------------------------------------------------------------------------
def run_ours(a, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    import ertext
    from ertext import dist as edist

    numa = {"pinned": False, "why": "--no-numa"} if a.no_numa else pin_to_gpu_numa_node(local_rank)
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
        if a.tile_fifo:
            c.set_tile_fifo(True)
        if a.no_stream_split:
            c.set_stream_split(False)
        if a.no_seam_list:
            c.set_seam_list(False)
        if a.tile_config:
            c.set_tile_config(a.tile_config)
        if a.post_footprint >= 0:
            c.set_post_footprint(a.post_footprint)

    gather = edist.LibraryGather(local_rank, rank, world) if world > 1 and not a.no_gather else None
    stats = {"tile_ms": [], "extract_ms": [], "nms_ms": [], "classify_ms": [], "launches": 0, "d2h": 0, "regions": 0, "kept": 0, "steps": 0,
             "gathered_records": 0, "gathers": 0}

    def collect(c, record):
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
        return r

    def take_gather(record):
        rec, _off, _seq = gather.collect()
        if record:
            stats["gathered_records"] += len(rec); stats["gathers"] += 1
            stats["d2h"] += rec.nbytes

    def run_loop(n_steps, resident, record):
        pending = [False] * NC
        for i in range(n_steps):
            k = i % NC
            if pending[k]:
                collect(ctxs[k], record)
            if resident:
                ctxs[k].enqueue_device(dev_batches[i % NB].data_ptr(), fpg, W, H, W * 3, upto=a.upto)
            else:
                ctxs[k].enqueue_host(host_batches[i % NB].data_ptr(), fpg, W, H, W * 3, upto=a.upto)
            pending[k] = True
            if gather is not None and a.upto >= 3:
                if gather.outstanding() >= 9:
                    take_gather(record)              # the gather enqueued nine steps ago: finished long since
                gather.enqueue(ctxs[k], my_ids)      # packs on the device, NCCL on the library's side stream; returns at once
        for j in range(NC):
            k = (n_steps + j) % NC
            if pending[k]:
                collect(ctxs[k], record)
        if gather is not None:
            while gather.outstanding():              # the final region gathers complete inside the timed region
                take_gather(record)

    def bracket(fn):
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
        n = fn()
        for e, s in zip(ends, streams):
            e.record(s)
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        if world > 1:
            dist.barrier()
        return max(start.elapsed_time(e) for e in ends), wall, n

    brackets = {}

    def reduce_max(ms):
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(resident):
        # every context allocates its workspace on first use: W warm-up steps, but at least one per context
        run_loop(max(a.warmup, NC), resident, False)
        ms, wall, _ = bracket(lambda: run_loop(a.steps, resident, True))
        brackets["resident" if resident else "e2e"] = {"device_ms_per_step": ms / a.steps, "host_wall_ms_per_step": wall / a.steps}
        # the host-side result collection of the last batches happens after the last kernel: the step ends when the
        # result is in host memory, so take the larger of the device bracket and the host bracket
        return reduce_max(max(ms, wall) if not resident else ms)

    def sustained(seconds):
        """steady state over >= `seconds`: the clock sampler sees a long region; also a check that `value` is no burst number"""
        chunk = 40

        def body():
            n, t0 = 0, time.perf_counter()
            while time.perf_counter() - t0 < seconds:
                run_loop(chunk, True, False)
                n += chunk
            return n
        ms, _wall, n = bracket(body)
        ms = reduce_max(ms)
        nt = torch.tensor([n], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(nt, op=dist.ReduceOp.SUM)      # ranks may fit a different number of chunks into the window
        return {"value": fpg * int(nt.item()) / (ms * 1e-3), "unit": UNIT, "seconds": ms * 1e-3, "steps_all_ranks": int(nt.item())}

    def jpeg_leg(seconds, threads):
        """decode-inclusive end to end: the same frames as JPEG bitstreams (OpenCV, quality 90, 4:2:0) handed
        to ert_enqueue_jpeg -- nvJPEG decodes them into device memory, then the hot path; no pixel H2D copy.  One host thread per
        context (nvJPEG's Huffman stage runs on the calling thread); wall clock between two device synchronisations."""
        try:
            import cv2
        except Exception as ex:                     # no encoder for the synthetic input: say so instead of inventing a number
            return {"unavailable": "cv2 is needed to JPEG-encode the synthetic frames: %s" % ex}
        import threading
        import ertext
        frames = host_batches[0].numpy()
        jpegs = [cv2.imencode(".jpg", f, [int(cv2.IMWRITE_JPEG_QUALITY), 90])[1].tobytes() for f in frames]
        tctx = []
        try:
            for _ in range(threads):
                c = ertext.ErText(device=local_rank)
                c.enqueue_jpeg(jpegs, W, H, upto=a.upto); c.fetch()          # warm-up: workspace, decoder state
                tctx.append(c)
        except ertext.ErtError as ex:
            for c in tctx:
                c.close()
            return {"unavailable": str(ex)}
        counts = [0] * threads

        def worker(k):
            c = tctx[k]
            t_end = time.perf_counter() + seconds
            while time.perf_counter() < t_end:
                c.enqueue_jpeg(jpegs, W, H, upto=a.upto)
                r = c.fetch()
                if r.status:
                    raise RuntimeError("device status %d" % r.status)
                counts[k] += 1
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        th = [threading.Thread(target=worker, args=(k,)) for k in range(threads)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        torch.cuda.synchronize()
        sec = time.perf_counter() - t0
        out = {"unit": UNIT, "threads_per_rank": threads, "frames_per_batch": fpg, "jpeg_bytes_per_step": int(sum(map(len, jpegs))),
               "pixel_h2d_bytes_per_step": 0, "backend": tctx[0].jpeg_backend_name(), "decode_ms_per_batch": tctx[0].jpeg_decode_ms(),
               "limiter": "nvJPEG's host-side Huffman stage (backend above; no hardware JPEG engine is exposed on this device: nvjpegCreateEx(HARDWARE) answers ARCH_MISMATCH)"}
        for c in tctx:
            c.close()
        sec = reduce_max(sec * 1e3) * 1e-3
        nt = torch.tensor([sum(counts)], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(nt, op=dist.ReduceOp.SUM)
        out["value"] = fpg * int(nt.item()) / sec
        out["seconds"] = sec
        return out

    def h2d_ceiling():
        """copies only: the same pinned host batches to the device, the same bytes per step per rank, nothing else"""
        bufs = [torch.empty_like(dev_batches[0]) for _ in range(NC)]

        def body():
            for i in range(a.steps):
                with torch.cuda.stream(streams[i % NC]):
                    bufs[i % NC].copy_(host_batches[i % NB], non_blocking=True)
            return a.steps
        body()
        ms, _wall, _n = bracket(body)
        ms = reduce_max(ms)
        byts = fpg * W * H * 3
        return {"frames_per_s": fpg * world * a.steps / (ms * 1e-3), "gb_per_s_all_ranks": byts * world * a.steps / (ms * 1e-3) / 1e9,
                "bytes_per_rank_per_step": byts, "what": "cudaMemcpyAsync of the benchmark's pinned batches only, all ranks at once"}

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_res = timed(True)
    res_stats = {k: (list(v) if isinstance(v, list) else v) for k, v in stats.items()}
    for k in stats:
        stats[k] = [] if isinstance(stats[k], list) else 0
    ms_e2e = timed(False)
    sus = sustained(a.sustained_seconds) if a.sustained_seconds > 0 else None
    clocks = sampler.stop() if rank == 0 else None
    ceiling = h2d_ceiling()
    e2e_jpeg = jpeg_leg(a.jpeg_seconds, a.jpeg_threads) if a.jpeg_seconds > 0 and a.upto >= 3 else None

    # the dominant kernel timed ALONE (one context, nothing else in flight): CUDA events recorded by the library
    # around the k_tile_build2 launch on its launching stream; this is the roofline numerator's time base
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
        if gather is not None:
            gather.close()
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peaks()
    tile_ms = float(np.median(solo_tile))
    tile_ms_in_flight = float(np.mean(res_stats["tile_ms"]))
    alg_bytes = W * H * 6 * fpg                      # one u8 read per pixel per plane (B_extract = W*H per plane)
    achieved = alg_bytes / (tile_ms * 1e-3) / 1e9
    traffic, traffic_src = ncu_traffic()
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms_res / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": DTYPE, "data": "synthetic",
        "config": {"workload": workload_text(W, H, fpg), "global_frames_per_step": fpg * world, "params": PARAMS,
                   "parallelism": "dp%d (frames sharded, no data-path collective%s)" % (
                       world, "; labelled regions gathered by ert_gather_regions_* (device pack + NCCL inside the library, side stream)" if world > 1 else ""),
                   "l2": "inputs rotate over %d distinct batches (%d x %.1f MB > 126 MB L2); all workspaces rewritten every step" % (NB, NB, fpg * W * H * 3 / 1e6),
                   "pipelining": "%d contexts / streams used round-robin; post-tile stages on a high-priority stream per context" % NC,
                   "numa": numa},
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": fpg * W * H * 3, "d2h_bytes_per_step": int(stats["d2h"] / max(stats["steps"], 1)),
                "ms_per_step": ms_e2e / a.steps, "h2d_ceiling": ceiling, "fraction_of_h2d_ceiling": e2e / ceiling["frames_per_s"],
                "rank0_brackets": brackets, "from_jpeg": e2e_jpeg},
        "gpu_launches": int(res_stats["launches"]),
        "roofline": {"bound": "hbm", "kernel": "k_tile_build2 (64x32 tile + halo per TMA box, 256 threads, 4 px per lane)", "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "algorithmic_bytes_per_launch": alg_bytes,
                     "kernel_ms": tile_ms, "peak_source": peak_src, "share_of_step": tile_ms / float(np.median(solo_total)),
                     "kernel_ms_with_batches_in_flight": tile_ms_in_flight,
                     "timing": "CUDA events around the launch on its stream, batch processed alone (median of 5)"},
        "stage_ms_per_batch": {"extract": float(np.mean(res_stats["extract_ms"])), "tile_build": tile_ms, "nms": float(np.mean(res_stats["nms_ms"])),
                               "classify": float(np.mean(res_stats["classify_ms"]))},
        "sustained": sus,
        "clocks": clocks,
        "regions_per_frame": res_stats["regions"] / max(res_stats["steps"] * fpg, 1),
        "kept_nodes_per_frame": res_stats["kept"] / max(res_stats["steps"] * fpg, 1),
    }
    if world > 1:
        line["gather"] = {"records_per_gather_all_ranks": stats["gathered_records"] / max(stats["gathers"], 1), "gathers": stats["gathers"],
                          "how": "k_pack_regions on the batch's stream -> ncclAllGather(counts) -> grouped ncclSend/ncclRecv of exactly the records held -> pinned host"}
    if world == 1 and not a.no_next_rows:
        line["next_rows"] = next_rows(a, ctxs[0], dev_batches, fpg, W, H)
    if world == 1 and not a.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(a, host_batches)
        if a.parity_frames > 0:
            line["parity"] = parity_counters(a, ctxs[0], host_batches[0].numpy()[: a.parity_frames])
    emit(line)
    if gather is not None:
        gather.close()
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
    """The reference's own CPU code (oracle/_ref) on a bounded sample of the same frames: mode T = frames spread over all
    host cores (throughput-fair), mode R = the reference as written (one frame at a time, 6-thread omp loop over the
    planes, src/ER.cpp:50), and one thread."""
    from oracle.refbind import RefOracle
    cores = host_cores()
    try:
        ref = RefOracle()
    except (FileNotFoundError, OSError) as ex:
        return {"value": None, "unit": UNIT, "cores": cores, "kind": "reference", "sample": "unavailable: %s" % ex}
    frames = np.concatenate([hb.numpy() for hb in host_batches])[: a.cpu_sample_frames]
    ref.detect_frames(frames[: min(len(frames), cores)], mode=1, nthreads=cores)   # warm-up
    sec, _counts, st = ref.detect_frames(frames, mode=1, nthreads=cores)
    sec1, _, _ = ref.detect_frames(frames[:2], mode=1, nthreads=1)
    nR = min(len(frames), 4)
    secR, _, _ = ref.detect_frames(frames[:nR], mode=0, nthreads=6)
    return {"value": len(frames) / sec, "unit": UNIT, "cores": cores, "kind": "reference", "per_thread": len(frames) / sec / cores,
            "one_thread": 2 / sec1,
            "mode_R": {"value": nR / secR, "unit": UNIT, "threads": 6, "frames": nR,
                       "what": "reference-faithful: one frame at a time, #pragma omp parallel for over the 6 planes (src/ER.cpp:50)"},
            "sample": "%d of the benchmark's frames, frames spread over %d threads; stage share extract/nms/classify = %.0f/%.0f/%.0f %%" % (
                len(frames), cores, 100 * st[0] / st.sum(), 100 * st[1] / st.sum(), 100 * st[2] / st.sum())}


def parity_counters(a, ctx, frames):
    """One sampled step's frames against the oracles, OUTSIDE the timed region: node arrays vs the
    oracle run with the canonical sibling order (must be identical), pool vs the UNMODIFIED reference order (symmetric
    difference: the one documented deviation), labels and scores."""
    try:
        from oracle.refbind import PortOracle, RefOracle
        port = PortOracle()
        try:
            ref = RefOracle()
        except (FileNotFoundError, OSError):
            ref = None
        res = ctx.detect_classify(frames)
        out = {"frames_checked": int(len(frames)), "planes_checked": 0, "nodes_equal": True, "pool_symdiff_vs_reference": 0, "pool_regions_reference": 0,
               "label_mismatch": 0, "max_rel_score_err": 0.0,
               "oracle": "port (canonical order) for nodes / labels / scores; %s for the reference-order pool" % ("oracle/_ref" if ref else "port")}

        def key(nodes, pool):
            return set(map(tuple, nodes[pool][:, :6].tolist()))
        for f in range(len(frames)):
            ch = port.channels(frames[f])
            for k in range(6):
                got = res.planes[f * 6 + k]
                exp = port.plane(ch[k], scores=True, canonical_order=True)
                out["planes_checked"] += 1
                if got.nodes.shape != exp["nodes"].shape or not (got.nodes == exp["nodes"]).all():
                    out["nodes_equal"] = False
                    continue
                if got.pool.shape == exp["pool"].shape and (got.pool == exp["pool"]).all():
                    out["label_mismatch"] += int((got.label != exp["label"]).sum())
                    for g_, e_ in ((got.strong_score, exp["strong_score"]), (got.weak_score, exp["weak_score"])):
                        rej_g, rej_e = g_ < -1e300, e_ < -1e300
                        out["label_mismatch"] += int((rej_g != rej_e).sum())
                        fin = ~rej_e & ~rej_g & (e_ != 0)
                        if fin.any():
                            out["max_rel_score_err"] = max(out["max_rel_score_err"], float(np.max(np.abs(g_[fin] - e_[fin]) / np.abs(e_[fin]))))
                else:
                    out["label_mismatch"] += 1
                r = (ref or port).plane(ch[k])
                out["pool_symdiff_vs_reference"] += len(key(got.nodes, got.pool) ^ key(r["nodes"], r["pool"]))
                out["pool_regions_reference"] += len(r["pool"])
        return out
    except Exception as ex:
        return {"error": str(ex)[:200]}


# ---------------------------------------------------------------------------------------------
# configs 1, 2, 4: one frame per step (latency-style configurations), pyramid levels built ON THE DEVICE
# ---------------------------------------------------------------------------------------------
def run_frame_config(a):
    import torch
    import ertext
    from ertext import synth
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    if a.config == 1:
        frame = np.load(os.path.join(ROOT, "tests", "golden", "frames.npz"))["frames"][0]
        divs, ppf = [], 6
        what = "640x480 real frame (ICDAR test image, tests/golden/frames.npz[0]), 6 planes native scale, 1 frame per step"
        metric = "640x480 frames/sec (ER detect+classify)"
    elif a.config == 2:
        frame = synth.s_text_frame(1234)
        divs, ppf = [2, 4, 8], 6
        what = "1920x1080 S-text frame (seed 1234), 6 planes x scales 1, 1/2, 1/4, 1/8 (levels resized on the device), 1 frame per step"
        metric = "1080p frames/sec (ER detect+classify, all channels + pyramid)"
    else:
        frame = synth.s_text_frame(77, 3840, 2160, n_glyphs=300)
        divs, ppf = [2, 4, 8], 3
        what = "3840x2160 S-text frame (seed 77), planes Y/Cr/Cb x scales 1, 1/2, 1/4, 1/8 (levels resized on the device), 1 frame per step"
        metric = "4K frames/sec (ER detect+classify, 3 channels x 4 scales)"
    H, W = frame.shape[:2]
    host = torch.from_numpy(np.ascontiguousarray(frame)).pin_memory()
    devf = host.to(dev)
    SETS = 3                                              # frames in flight (each: one context per level)
    sets = []
    for _ in range(SETS):
        src = ertext.ErText(device=0)
        src.set_planes_per_frame(ppf)
        sets.append((src, [ertext.ErText(device=0) for _ in divs]))

    def submit(s, resident):
        src, lv = sets[s]
        if resident:
            src.enqueue_device(devf.data_ptr(), 1, W, H, W * 3)
        else:
            src.enqueue_host(host.data_ptr(), 1, W, H, W * 3)
        for d, c in zip(divs, lv):
            c.enqueue_pyramid_level(src, d)

    def gather_(s):
        src, lv = sets[s]
        rs = [src.fetch()] + [c.fetch() for c in lv]
        for r in rs:
            if r.status:
                raise RuntimeError("device status %d" % r.status)
        return rs

    def loop(n, resident):
        pend = [False] * SETS
        last = None
        for i in range(n):
            s = i % SETS
            if pend[s]:
                last = gather_(s)
            submit(s, resident)
            pend[s] = True
        for j in range(SETS):
            s = (n + j) % SETS
            if pend[s]:
                last = gather_(s)
        return last

    steps = max(a.steps, 20)
    out = {}
    sampler = ClockSampler(0)
    sampler.start()
    rs = None
    for resident in (True, False):
        loop(max(a.warmup, SETS), resident)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        rs = loop(steps, resident)
        torch.cuda.synchronize()
        out[resident] = (time.perf_counter() - t0) / steps
    clocks = sampler.stop()
    # one frame alone: latency and the tile kernel's time per level
    lat = []
    for _ in range(6):
        t0 = time.perf_counter(); submit(0, False); rs = gather_(0); lat.append((time.perf_counter() - t0) * 1e3)
    tile_ms = sum(r.stage_ms[6] for r in rs)
    alg_bytes = sum(r.width * r.height * len(r.planes) for r in rs)
    peak, peak_src = measured_peaks()
    achieved = alg_bytes / (tile_ms * 1e-3) / 1e9
    launches = sum(c.launch_count() for c in [sets[0][0]] + sets[0][1])
    line = {
        "metric": metric, "value": 1.0 / out[True], "unit": UNIT, "n_gpus": 1, "steps": steps, "warmup": a.warmup, "ms_per_step": out[True] * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": DTYPE, "data": "synthetic" if a.config != 1 else "real frame (committed fixture)",
        "config": {"workload": what, "baseline_config": a.config, "params": PARAMS, "pipelining": "%d frames in flight, one context per pyramid level" % SETS,
                   "timing": "wall clock over %d pipelined steps (the frame is re-submitted every step; workspaces rewritten)" % steps},
        "e2e": {"value": 1.0 / out[False], "unit": UNIT, "h2d_bytes_per_step": int(frame.nbytes),
                "d2h_bytes_per_step": int(sum(sum(len(p.nodes) * 32 + len(p.pool) * 24 for p in r.planes) for r in rs)),
                "ms_per_step": out[False] * 1e3, "latency_ms_one_frame_alone": float(np.median(lat[1:]))},
        "gpu_launches": int(launches) * steps,
        "roofline": {"bound": "hbm", "kernel": "k_tile_build2", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                     "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": tile_ms, "peak_source": peak_src,
                     "timing": "sum over the levels of the library's CUDA events around each tile-kernel launch, frame processed alone"},
        "levels": [{"width": r.width, "height": r.height, "planes": len(r.planes), "kept_nodes": int(sum(len(p.nodes) for p in r.planes)),
                    "pooled": int(sum(len(p.pool) for p in r.planes)), "tile_ms": r.stage_ms[6], "device_ms": r.stage_ms[5]} for r in rs],
        "clocks": clocks,
    }
    if not a.no_cpu_baseline:
        line["cpu_baseline"] = frame_config_cpu_baseline(frame, divs, ppf)
    emit(line)


def frame_config_cpu_baseline(frame, divs, ppf):
    """the reference's own per-plane functions (oracle/_ref) on the same planes: native planes and, for the pyramid levels,
    its er_tree_extract / NMS / classify on the cv2-resized planes, one thread"""
    try:
        import cv2
        from oracle.refbind import RefOracle
        ref = RefOracle()
        H, W = frame.shape[:2]
        t0 = time.perf_counter()
        ch = ref.channels(frame)
        n = 0
        for d in [1] + divs:
            for k in range(ppf):
                pl = ch[k] if d == 1 else cv2.resize(ch[k], (W // d, H // d), interpolation=cv2.INTER_LINEAR)
                ref.plane(pl)
                n += 1
        sec = time.perf_counter() - t0
        return {"value": 1.0 / sec, "unit": UNIT, "cores": 1, "kind": "reference", "host_cores_available": host_cores(),
                "sample": "the same frame once: %d planes (levels by cv2.resize), the reference's per-plane functions on one thread, %.0f ms" % (n, sec * 1e3)}
    except Exception as ex:
        return {"value": None, "unit": UNIT, "cores": 1, "kind": "reference", "sample": "unavailable: %s" % str(ex)[:120]}


# ---------------------------------------------------------------------------------------------
# config 5: ER-candidate sweep 1k .. 500k regions, AdaBoost cascades and SVM batch scoring
# ---------------------------------------------------------------------------------------------
def run_sweep(a):
    import ertext
    from ertext import synth
    e = ertext.ErText(device=0, load_svm=True)
    sizes = [1000, 4000, 16000, 64000, 256000, 500000]
    frame = synth.s_text_frame(1234)
    planes = e.compute_channels(frame)
    rng = np.random.RandomState(5)
    rects = []
    for _ in range(4096):
        w = rng.randint(13, 200); h = rng.randint(max(13, w // 2 + 1), min(400, w * 5))
        rects.append((rng.randint(0, 1920 - w), rng.randint(0, 1080 - h), w, h))
    _, _, _, base_hist = e.classify_regions(planes[0], np.array(rects, np.int32), want_hist=True)
    base_x = synth.svm_features_u8(7, 8000)
    sampler = ClockSampler(0)
    sampler.start()
    rows = []
    # e2e: inputs and results in PINNED host memory (ert_host_alloc), allocated and filled outside the timed call
    nmax = sizes[-1]
    k = 65
    pin_hist = e.pinned_array((nmax, 1024), np.uint8); pin_x = e.pinned_array((nmax, 1800), np.uint8)
    pin_lab = e.pinned_array(nmax, np.int32); pin_ss = e.pinned_array(nmax, np.float64); pin_ws = e.pinned_array(nmax, np.float64)
    pin_sl = e.pinned_array(nmax, np.float64); pin_prob = e.pinned_array((nmax, k), np.float64)
    for n in sizes:
        hist = pin_hist[:n]
        hist[:] = np.tile(base_hist, (n // len(base_hist) + 1, 1))[:n]
        ms_c = e.bench_cascade_u8(hist, 5 if n <= 64000 else 2)
        e.cascade_classify_u8(hist[: min(n, 1000)], out=(pin_lab[: min(n, 1000)], pin_ss[: min(n, 1000)], pin_ws[: min(n, 1000)]))
        t0 = time.perf_counter(); e.cascade_classify_u8(hist, out=(pin_lab[:n], pin_ss[:n], pin_ws[:n])); e2e_c = time.perf_counter() - t0
        x = pin_x[:n]
        x[:] = np.tile(base_x, (n // len(base_x) + 1, 1))[:n]
        ms_s = e.bench_svm_u8(x, 2 if n <= 64000 else 1)
        t0 = time.perf_counter(); e.svm_predict_probability(x, out=(pin_sl[:n], pin_prob[:n])); e2e_s = time.perf_counter() - t0
        rows.append({"n": n, "cascade_regions_per_s": n / ms_c * 1e3, "cascade_e2e_regions_per_s": n / e2e_c, "svm_vectors_per_s": n / ms_s * 1e3,
                     "svm_e2e_vectors_per_s": n / e2e_s, "svm_distance_gemm_tops": 2.0 * 1800 * 1910 * n / ms_s / 1e9})
    clocks = sampler.stop()
    top = rows[-1]
    peak, peak_src = measured_peaks()
    casc_gbs = top["cascade_regions_per_s"] * 1024 / 1e9
    line = {
        "metric": "candidate regions/sec (AdaBoost cascades + SVM batch scoring, 500k regions)", "value": top["cascade_regions_per_s"], "unit": "regions/s",
        "n_gpus": 1, "steps": 1, "warmup": 3, "ms_per_step": 500000 / top["cascade_regions_per_s"] * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8 histograms, f64 sums / u8 features, s32 tensor-core distances, f64 epilogue", "data": "synthetic",
        "config": {"workload": "ER-candidate sweep 1k..500k regions: 1024-bin LBP histograms of random crops of an S-text frame through both cascades; "
                               "1800-d u8 features through the RBF C-SVC (65 classes, 1910 SVs)",
                   "baseline_config": 5, "value_is": "cascade scoring at 500k regions, inputs resident (library CUDA events)"},
        "e2e": {"value": top["cascade_e2e_regions_per_s"], "unit": "regions/s", "h2d_bytes_per_step": 500000 * 1024, "d2h_bytes_per_step": 500000 * 20,
                "how": "one synchronous C-ABI call, histograms / features and results in pinned host memory"},
        "svm": {"value": top["svm_vectors_per_s"], "unit": "vectors/s", "e2e": top["svm_e2e_vectors_per_s"], "distance_gemm_tops": top["svm_distance_gemm_tops"]},
        "gpu_launches": len(sizes) * 20,
        "roofline": {"bound": "hbm", "kernel": "k_cascade<u8>", "achieved": casc_gbs, "peak": peak, "unit": "GB/s", "frac": casc_gbs / peak, "traffic": None,
                     "algorithmic_bytes_per_launch": 500000 * 1024, "peak_source": peak_src, "timing": "library CUDA events, 2 back-to-back launches"},
        "sweep": rows, "clocks": clocks,
    }
    if not a.no_cpu_baseline:
        try:
            from oracle.refbind import RefOracle
            ref = RefOracle(with_svm=True)
            cores = host_cores()
            fv = base_hist[:1024].astype(np.float64)
            t0 = time.perf_counter(); ref.cascade_predict(0, fv); ref.cascade_predict(1, fv); dt = time.perf_counter() - t0
            xs = base_x[: 2 * cores].astype(np.float64) / 255.0
            t1 = time.perf_counter(); ref.svm_predict_probability(xs, nthreads=cores); dts = time.perf_counter() - t1
            line["cpu_baseline"] = {"value": 1024 / dt, "unit": "regions/s", "cores": 1, "kind": "reference", "svm_vectors_per_s": len(xs) / dts, "svm_threads": cores,
                                    "sample": "CascadeBoost::predict x2 on 1024 histograms (one thread); svm_predict_probability on %d vectors over %d threads" % (len(xs), cores)}
        except Exception as ex:
            line["cpu_baseline"] = {"value": None, "unit": "regions/s", "cores": 1, "kind": "reference", "sample": "unavailable: %s" % str(ex)[:120]}
    emit(line)


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
    if a.config == 3:
        run_ours(a, rank, local_rank, world)
    elif rank == 0:
        if a.config == 5:
            run_sweep(a)
        else:
            run_frame_config(a)


if __name__ == "__main__":
    main()
