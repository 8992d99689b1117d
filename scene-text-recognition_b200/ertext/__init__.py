"""ertext -- thin ctypes binding of libertext.so (the C ABI in include/ertext.h).

Used by tests/, bench.py and __graft_entry__.py.  There is no Python or CPU implementation of
the path behind this module: if libertext.so is missing or no CUDA device is present, it raises.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PKG_ROOT = os.path.dirname(_HERE)
REPO_ROOT = os.path.dirname(PKG_ROOT)
LIB_PATH = os.path.join(PKG_ROOT, "libertext.so")
ASSETS = os.path.join(REPO_ROOT, "assets", "classifier")

STAGE_EXTRACT, STAGE_NMS, STAGE_CLASSIFY, STAGE_TRACK = 1, 2, 3, 4
LABEL_NONE, LABEL_WEAK, LABEL_STRONG = 0, 1, 2
NEG_DBL_MAX = -1.7976931348623157e308

_u8p = C.POINTER(C.c_uint8)
_i32p = C.POINTER(C.c_int32)
_f64p = C.POINTER(C.c_double)


class ErtParams(C.Structure):
    _fields_ = [("thresh_step", C.c_int), ("min_area", C.c_int), ("max_area", C.c_int), ("stability_t", C.c_int),
                ("overlap_coef", C.c_double), ("min_ocr_prob", C.c_double)]


class ErtResult(C.Structure):
    _fields_ = [("n_planes", C.c_int32), ("width", C.c_int32), ("height", C.c_int32),
                ("node_offset", _i32p), ("nodes", _i32p), ("pool_offset", _i32p), ("pool_node", _i32p),
                ("pool_label", _i32p), ("pool_strong_score", _f64p), ("pool_weak_score", _f64p),
                ("pool_hist", _u8p), ("status", C.c_uint32), ("stage_ms", C.c_double * 8),
                ("plane_order_sensitive", _i32p), ("order_sensitive_total", C.c_int32)]


class ErtTracked(C.Structure):
    _fields_ = [("plane", C.c_int32), ("pool_index", C.c_int32), ("node", C.c_int32), ("label", C.c_int32), ("level", C.c_int32),
                ("area", C.c_int32), ("x", C.c_int32), ("y", C.c_int32), ("w", C.c_int32), ("h", C.c_int32),
                ("center_x", C.c_int32), ("center_y", C.c_int32), ("color1", C.c_double), ("color2", C.c_double), ("color3", C.c_double)]


TRACKED_DTYPE = np.dtype([("plane", "<i4"), ("pool_index", "<i4"), ("node", "<i4"), ("label", "<i4"), ("level", "<i4"), ("area", "<i4"),
                          ("x", "<i4"), ("y", "<i4"), ("w", "<i4"), ("h", "<i4"), ("center_x", "<i4"), ("center_y", "<i4"),
                          ("color1", "<f8"), ("color2", "<f8"), ("color3", "<f8")])
assert TRACKED_DTYPE.itemsize == C.sizeof(ErtTracked) == 72


class ErtTrackResult(C.Structure):
    _fields_ = [("n_frames", C.c_int32), ("cand_offset", _i32p), ("n_strong", _i32p), ("cand", C.POINTER(ErtTracked)),
                ("track_offset", _i32p), ("tracked", _i32p), ("track_ms", C.c_double)]


class ErtOcrRegion(C.Structure):
    _fields_ = [("frame", C.c_int32), ("plane", C.c_int32), ("x", C.c_int32), ("y", C.c_int32), ("w", C.c_int32), ("h", C.c_int32),
                ("slope", C.c_double)]


OCR_REGION_DTYPE = np.dtype([("frame", "<i4"), ("plane", "<i4"), ("x", "<i4"), ("y", "<i4"), ("w", "<i4"), ("h", "<i4"), ("slope", "<f8")])
assert OCR_REGION_DTYPE.itemsize == C.sizeof(ErtOcrRegion) == 32


class ErtOcrResult(C.Structure):
    _fields_ = [("n", C.c_int32), ("nr_class", C.c_int32), ("value", _f64p), ("label", _i32p), ("prob_all", _f64p), ("feat", _u8p),
                ("img", _u8p), ("ocr_ms", C.c_double)]


class FrameTrack:
    """er_track of one frame: cand = structured array (TRACKED_DTYPE) of every strong then every weak region,
    n_strong, tracked = indices into cand in the reference's all_er order."""
    __slots__ = ("cand", "n_strong", "tracked")

    def __init__(self, cand, n_strong, tracked):
        self.cand, self.n_strong, self.tracked = cand, n_strong, tracked


class OcrResult:
    __slots__ = ("value", "label", "prob", "feat", "img", "ocr_ms")

    def __init__(self, value, label, prob, feat, img, ms):
        self.value, self.label, self.prob, self.feat, self.img, self.ocr_ms = value, label, prob, feat, img, ms


EXPORTS = [
    "ert_enqueue_jpeg", "ert_jpeg_fetch_frames", "ert_set_jpeg_backend", "ert_jpeg_backend_name", "ert_jpeg_decode_ms",
    "ert_set_tile_fifo", "ert_set_nms_sequential", "ert_host_alloc", "ert_host_free", "ert_batch_done", "ert_er_track", "ert_er_track_regions", "ert_ocr_chain_run_batch", "ert_ocr_chain_run_plane", "ert_ocr_features_plane",
    "ert_abi_version", "ert_last_error", "ert_status_string", "ert_create", "ert_destroy", "ert_set_thresh_step",
    "ert_set_min_area", "ert_set_return_hist", "ert_set_tile_config", "ert_set_node_capacity", "ert_debug_phase_cycles", "ert_set_capacity", "ert_load_cascade",
    "ert_load_svm", "ert_svm_nr_class", "ert_set_svm_tensor_cores", "ert_svm_dims", "ert_detect_classify", "ert_enqueue_host", "ert_detect_classify_device",
    "ert_fetch_result", "ert_compute_channels", "ert_planes_detect", "ert_enqueue_planes", "ert_nms_nodes", "ert_classify_regions", "ert_lbp_hist",
    "ert_cascade_predict_batch", "ert_cascade_classify_u8", "ert_svm_predict_probability_batch",
    "ert_svm_predict_probability_batch_u8", "ert_set_stream", "ert_get_stream", "ert_last_launch_count",
    "ert_bench_cascade_u8", "ert_bench_svm_u8",
    "ert_set_params", "ert_set_cascade", "ert_cascade_stage_info", "ert_svm_total_sv", "ert_svm_labels", "ert_svm_gamma", "ert_calc_lbp",
    "ert_er_track_regions_ycc", "ert_set_stream_split", "ert_set_seam_list", "ert_set_post_footprint", "ert_enqueue_pyramid_level", "ert_set_planes_per_frame", "ert_set_svm_legacy_prob",
    "ert_dist_unique_id", "ert_dist_create", "ert_dist_destroy", "ert_gather_regions_enqueue", "ert_gather_regions_collect", "ert_gather_regions_outstanding",
]

_lib = None


def load_library():
    """dlopen libertext.so (built in tree by `make -C scene-text-recognition_b200`) and declare prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("libertext.so is not built (%s): run __graft_entry__.build(); there is no fallback path" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    L.ert_last_error.restype = C.c_char_p
    L.ert_status_string.restype = C.c_char_p
    L.ert_status_string.argtypes = [C.c_uint32]
    L.ert_create.restype = C.c_void_p
    L.ert_create.argtypes = [C.POINTER(ErtParams), C.c_int]
    L.ert_destroy.argtypes = [C.c_void_p]
    for f in ("ert_set_thresh_step", "ert_set_min_area", "ert_set_return_hist", "ert_set_tile_config", "ert_set_node_capacity", "ert_set_tile_fifo", "ert_set_nms_sequential", "ert_set_stream_split", "ert_set_seam_list", "ert_set_post_footprint"):
        getattr(L, f).argtypes = [C.c_void_p, C.c_int]
    L.ert_set_capacity.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.ert_enqueue_pyramid_level.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    L.ert_set_planes_per_frame.argtypes = [C.c_void_p, C.c_int]
    L.ert_debug_phase_cycles.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_uint64)]
    L.ert_load_cascade.argtypes = [C.c_void_p, C.c_int, C.c_char_p]
    L.ert_load_svm.argtypes = [C.c_void_p, C.c_char_p]
    L.ert_svm_nr_class.argtypes = [C.c_void_p]
    L.ert_set_svm_tensor_cores.argtypes = [C.c_void_p, C.c_int]
    L.ert_set_svm_legacy_prob.argtypes = [C.c_void_p, C.c_int]
    L.ert_svm_dims.argtypes = [C.c_void_p]
    RP = C.POINTER(C.POINTER(ErtResult))
    L.ert_detect_classify.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, RP]
    L.ert_detect_classify_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    L.ert_enqueue_host.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    L.ert_enqueue_jpeg.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.c_int, C.c_int, C.c_int, C.c_int]
    L.ert_jpeg_fetch_frames.argtypes = [C.c_void_p, C.c_void_p]
    L.ert_set_jpeg_backend.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.ert_jpeg_backend_name.argtypes = [C.c_void_p]; L.ert_jpeg_backend_name.restype = C.c_char_p
    L.ert_jpeg_decode_ms.argtypes = [C.c_void_p]; L.ert_jpeg_decode_ms.restype = C.c_double
    L.ert_fetch_result.argtypes = [C.c_void_p, RP]
    L.ert_compute_channels.argtypes = [C.c_void_p, _u8p, C.c_int, C.c_int, C.c_int, _u8p]
    L.ert_planes_detect.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_int, RP]
    L.ert_enqueue_planes.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_int]
    L.ert_nms_nodes.argtypes = [C.c_void_p, _i32p, C.c_int, C.c_int, C.c_int, _i32p, C.c_int, C.POINTER(C.c_int)]
    L.ert_classify_regions.argtypes = [C.c_void_p, _u8p, C.c_int, C.c_int, C.c_int, _i32p, C.c_int, _i32p, _f64p, _f64p, _u8p]
    L.ert_lbp_hist.argtypes = [C.c_void_p, _u8p, C.c_int, C.c_int, C.c_int, _i32p, C.c_int, _f64p]
    L.ert_cascade_predict_batch.argtypes = [C.c_void_p, C.c_int, _f64p, C.c_int, C.c_int, _f64p]
    L.ert_cascade_classify_u8.argtypes = [C.c_void_p, _u8p, C.c_int, _i32p, _f64p, _f64p]
    L.ert_svm_predict_probability_batch.argtypes = [C.c_void_p, _f64p, C.c_int, _f64p, _f64p]
    L.ert_svm_predict_probability_batch_u8.argtypes = [C.c_void_p, _u8p, C.c_int, _f64p, _f64p]
    L.ert_set_stream.argtypes = [C.c_void_p, C.c_uint64]
    L.ert_get_stream.restype = C.c_uint64
    L.ert_get_stream.argtypes = [C.c_void_p]
    L.ert_last_launch_count.argtypes = [C.c_void_p]
    L.ert_bench_cascade_u8.argtypes = [C.c_void_p, _u8p, C.c_int, C.c_int, _f64p]
    L.ert_bench_svm_u8.argtypes = [C.c_void_p, _u8p, C.c_int, C.c_int, _f64p]
    L.ert_batch_done.argtypes = [C.c_void_p]
    L.ert_host_alloc.restype = C.c_void_p
    L.ert_host_alloc.argtypes = [C.c_size_t]
    L.ert_host_free.argtypes = [C.c_void_p]
    TP = C.POINTER(C.POINTER(ErtTrackResult))
    L.ert_er_track.argtypes = [C.c_void_p, TP]
    L.ert_er_track_regions.argtypes = [C.c_void_p, _u8p, C.c_int, C.c_int, C.c_int, _i32p, C.c_int, _i32p, C.c_int, TP]
    OP = C.POINTER(C.POINTER(ErtOcrResult))
    L.ert_ocr_chain_run_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int, OP]
    L.ert_ocr_chain_run_plane.argtypes = [C.c_void_p, _u8p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, OP]
    L.ert_ocr_features_plane.argtypes = [C.c_void_p, _u8p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, OP]
    _lib = L
    return L


def _ptr(a, t):
    return a.ctypes.data_as(t)


class ErtError(RuntimeError):
    pass


def svm_model_path():
    """assets/classifier/OCR.model is shipped xz-compressed; unpack once."""
    import lzma
    dst = os.path.join(ASSETS, "OCR.model")
    if not os.path.exists(dst):
        with lzma.open(dst + ".xz", "rb") as f, open(dst + ".tmp", "wb") as g:
            g.write(f.read())
        os.replace(dst + ".tmp", dst)
    return dst


class PlaneResult:
    """Kept nodes (DFS pre-order: level, area, x, y, w, h, parent, n_children), pool and labels of one plane."""
    __slots__ = ("nodes", "pool", "label", "strong_score", "weak_score", "hist")

    def __init__(self, nodes, pool, label, ss, ws, hist):
        self.nodes, self.pool, self.label, self.strong_score, self.weak_score, self.hist = nodes, pool, label, ss, ws, hist


class BatchResult:
    def __init__(self, planes, status, stage_ms, width, height, flat=None):
        self.planes, self.status, self.stage_ms, self.width, self.height = planes, status, stage_ms, width, height
        self.flat = flat   # (node_offset, nodes, pool_offset, pool_node, pool_label): the contiguous arrays of the C result


class ErText:
    """Context object = `new ERFilter(...)` + the two CascadeBoost objects (+ the OCR SVM model)."""

    def __init__(self, device=0, thresh_step=8, min_area=120, max_area=900000, stability_t=2, overlap_coef=0.7,
                 min_ocr_prob=0.15, load_cascades=True, load_svm=False):
        self.L = load_library()
        self._pinned = []
        prm = ErtParams(thresh_step, min_area, max_area, stability_t, overlap_coef, min_ocr_prob)
        self.ctx = self.L.ert_create(C.byref(prm), device)
        if not self.ctx:
            raise ErtError(self.L.ert_last_error().decode())
        if load_cascades:
            self.load_cascade(0, os.path.join(ASSETS, "strong.classifier"))
            self.load_cascade(1, os.path.join(ASSETS, "weak.classifier"))
        if load_svm:
            self.load_svm(svm_model_path())

    def _check(self, rc):
        if rc < 0:
            raise ErtError(self.L.ert_last_error().decode())
        return rc

    def close(self):
        if getattr(self, "ctx", None):
            self.L.ert_destroy(self.ctx)
            self.ctx = None
        for p in getattr(self, "_pinned", []):
            self.L.ert_host_free(p)
        self._pinned = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def load_cascade(self, which, path):
        return self._check(self.L.ert_load_cascade(self.ctx, which, path.encode()))

    def load_svm(self, path):
        return self._check(self.L.ert_load_svm(self.ctx, path.encode()))

    def set_thresh_step(self, t):
        self._check(self.L.ert_set_thresh_step(self.ctx, t))

    def set_min_area(self, m):
        self._check(self.L.ert_set_min_area(self.ctx, m))

    def set_return_hist(self, on):
        self._check(self.L.ert_set_return_hist(self.ctx, int(on)))

    def set_node_capacity(self, slots_per_plane):
        self._check(self.L.ert_set_node_capacity(self.ctx, int(slots_per_plane)))

    def set_nms_sequential(self, on):
        self._check(self.L.ert_set_nms_sequential(self.ctx, int(on)))

    def set_tile_fifo(self, on):
        self._check(self.L.ert_set_tile_fifo(self.ctx, int(on)))

    def set_stream_split(self, on):
        self._check(self.L.ert_set_stream_split(self.ctx, int(on)))

    def set_planes_per_frame(self, n):
        self._check(self.L.ert_set_planes_per_frame(self.ctx, int(n)))

    def enqueue_pyramid_level(self, src, div, upto=STAGE_CLASSIFY):
        """this context <- the planes of `src`'s batch in flight, resized by 1/div on the device; collect with fetch()"""
        self._check(self.L.ert_enqueue_pyramid_level(self.ctx, src.ctx, int(div), upto))

    def set_post_footprint(self, ctas_per_sm):
        self._check(self.L.ert_set_post_footprint(self.ctx, int(ctas_per_sm)))

    def set_seam_list(self, on):
        self._check(self.L.ert_set_seam_list(self.ctx, int(on)))

    def set_tile_config(self, i):
        self._check(self.L.ert_set_tile_config(self.ctx, i))

    def phase_cycles(self, enable=True):
        out = (C.c_uint64 * 16)()
        self._check(self.L.ert_debug_phase_cycles(self.ctx, int(enable), out))
        return list(out)

    def set_capacity(self, kept, pool):
        self._check(self.L.ert_set_capacity(self.ctx, kept, pool))

    def set_stream(self, s):
        self._check(self.L.ert_set_stream(self.ctx, s))

    def launch_count(self):
        return self.L.ert_last_launch_count(self.ctx)

    # ---- result unpacking ----------------------------------------------------------------------
    def _unpack(self, rp):
        r = rp.contents
        P = r.n_planes
        noff = np.ctypeslib.as_array(r.node_offset, (P + 1,)).copy()
        poff = np.ctypeslib.as_array(r.pool_offset, (P + 1,)).copy()
        nt, pt = int(noff[-1]), int(poff[-1])
        nodes = np.ctypeslib.as_array(r.nodes, (max(nt, 1) * 8,))[: nt * 8].reshape(nt, 8).copy()
        if pt:
            pool = np.ctypeslib.as_array(r.pool_node, (pt,)).copy()
            label = np.ctypeslib.as_array(r.pool_label, (pt,)).copy()
            ss = np.ctypeslib.as_array(r.pool_strong_score, (pt,)).copy()
            ws = np.ctypeslib.as_array(r.pool_weak_score, (pt,)).copy()
            hist = np.ctypeslib.as_array(r.pool_hist, (pt * 1024,)).reshape(pt, 1024).copy() if r.pool_hist else None
        else:
            pool = np.zeros(0, np.int32); label = np.zeros(0, np.int32); ss = np.zeros(0); ws = np.zeros(0); hist = None
        planes = []
        for p in range(P):
            a, b = int(poff[p]), int(poff[p + 1])
            planes.append(PlaneResult(nodes[noff[p]:noff[p + 1]], pool[a:b], label[a:b], ss[a:b], ws[a:b],
                                      hist[a:b] if hist is not None else None))
        br = BatchResult(planes, int(r.status), list(r.stage_ms), r.width, r.height, (noff, nodes, poff, pool, label))
        br.order_sensitive = np.ctypeslib.as_array(r.plane_order_sensitive, shape=(r.n_planes,)).copy() if r.plane_order_sensitive else None
        br.order_sensitive_total = int(r.order_sensitive_total)
        return br

    # ---- the batched hot path ------------------------------------------------------------------
    def detect_classify(self, bgr, upto=STAGE_CLASSIFY):
        """bgr: uint8 [F,H,W,3] (or [H,W,3]) in host memory -> BatchResult with 6*F planes."""
        bgr = np.ascontiguousarray(bgr, dtype=np.uint8)
        if bgr.ndim == 3:
            bgr = bgr[None]
        f, h, w, c = bgr.shape
        assert c == 3
        rp = C.POINTER(ErtResult)()
        self._check(self.L.ert_detect_classify(self.ctx, bgr.ctypes.data, f, w, h, w * 3, upto, C.byref(rp)))
        return self._unpack(rp)

    def detect_classify_ptr(self, host_ptr, f, w, h, stride, upto=STAGE_CLASSIFY):
        """Raw host pointer (e.g. pinned torch tensor .data_ptr())."""
        rp = C.POINTER(ErtResult)()
        self._check(self.L.ert_detect_classify(self.ctx, host_ptr, f, w, h, stride, upto, C.byref(rp)))
        return self._unpack(rp)

    def enqueue_host(self, host_ptr, f, w, h, stride, upto=STAGE_CLASSIFY):
        self._check(self.L.ert_enqueue_host(self.ctx, host_ptr, f, w, h, stride, upto))

    def enqueue_host_array(self, bgr, upto=STAGE_CLASSIFY):
        """enqueue one frame [H, W, 3] or a batch [F, H, W, 3] held in a numpy array (kept alive until fetch)"""
        a = np.ascontiguousarray(bgr, dtype=np.uint8)
        if a.ndim == 3:
            a = a[None]
        self._keep = a
        f, h, w, _ = a.shape
        self._check(self.L.ert_enqueue_host(self.ctx, a.ctypes.data, f, w, h, w * 3, upto))

    def enqueue_jpeg(self, jpegs, w, h, upto=STAGE_CLASSIFY):
        """enqueue a batch of JPEG bitstreams (bytes-like, each w x h): decoded on the device by nvJPEG, no pixel H2D copy"""
        bufs = [np.frombuffer(j, dtype=np.uint8) for j in jpegs]
        self._keep = bufs
        n = len(bufs)
        ptrs = (C.c_void_p * n)(*[b.ctypes.data for b in bufs])
        sizes = (C.c_size_t * n)(*[b.size for b in bufs])
        self._jpeg_shape = (n, h, w, 3)
        self._check(self.L.ert_enqueue_jpeg(self.ctx, ptrs, sizes, n, w, h, upto))

    def jpeg_fetch_frames(self):
        out = np.empty(self._jpeg_shape, dtype=np.uint8)
        self._check(self.L.ert_jpeg_fetch_frames(self.ctx, out.ctypes.data))
        return out

    def set_jpeg_backend(self, backend=-1, cpu_threads=4):
        self._check(self.L.ert_set_jpeg_backend(self.ctx, backend, cpu_threads))

    def jpeg_backend_name(self):
        r = self.L.ert_jpeg_backend_name(self.ctx)
        return r.decode() if r else None

    def jpeg_decode_ms(self):
        return self.L.ert_jpeg_decode_ms(self.ctx)

    def enqueue_device(self, dev_ptr, f, w, h, stride, upto=STAGE_CLASSIFY):
        self._check(self.L.ert_detect_classify_device(self.ctx, dev_ptr, f, w, h, stride, upto))

    def fetch(self):
        rp = C.POINTER(ErtResult)()
        self._check(self.L.ert_fetch_result(self.ctx, C.byref(rp)))
        return self._unpack(rp)

    def compute_channels(self, bgr):
        """ERFilter::compute_channels: uint8 [H,W,3] BGR -> uint8 [6,H,W] (Y, Cr, Cb and their inverses)."""
        bgr = np.ascontiguousarray(bgr, dtype=np.uint8)
        h, w, _ = bgr.shape
        out = np.zeros((6, h, w), np.uint8)
        self._check(self.L.ert_compute_channels(self.ctx, _ptr(bgr, _u8p), w, h, w * 3, _ptr(out, _u8p)))
        return out

    def planes_detect(self, planes, upto=STAGE_CLASSIFY):
        """planes: uint8 [P,H,W] (or [H,W]) single-channel images."""
        planes = np.ascontiguousarray(planes, dtype=np.uint8)
        if planes.ndim == 2:
            planes = planes[None]
        p, h, w = planes.shape
        rp = C.POINTER(ErtResult)()
        self._check(self.L.ert_planes_detect(self.ctx, planes.ctypes.data, p, w, h, w, w * h, upto, C.byref(rp)))
        return self._unpack(rp)

    def enqueue_planes(self, planes, upto=STAGE_CLASSIFY):
        """asynchronous planes_detect: returns at once, collect with fetch(); `planes` must stay alive until then"""
        assert planes.dtype == np.uint8 and planes.flags["C_CONTIGUOUS"] and planes.ndim == 3
        p, h, w = planes.shape
        self._check(self.L.ert_enqueue_planes(self.ctx, planes.ctypes.data, p, w, h, w, w * h, upto))

    # ---- stage entry points --------------------------------------------------------------------
    def nms_nodes(self, nodes, w, h):
        nodes = np.ascontiguousarray(nodes, dtype=np.int32)
        n = nodes.shape[0]
        pool = np.zeros(max(n, 1), np.int32)
        npool = C.c_int(0)
        self._check(self.L.ert_nms_nodes(self.ctx, _ptr(nodes, _i32p), n, w, h, _ptr(pool, _i32p), pool.size, C.byref(npool)))
        return pool[: npool.value].copy()

    def classify_regions(self, plane, rects, want_hist=False):
        plane = np.ascontiguousarray(plane, dtype=np.uint8)
        rects = np.ascontiguousarray(rects, dtype=np.int32).reshape(-1, 4)
        h, w = plane.shape
        n = rects.shape[0]
        label = np.zeros(n, np.int32); ss = np.zeros(n); ws = np.zeros(n)
        hist = np.zeros((n, 1024), np.uint8) if want_hist else None
        self._check(self.L.ert_classify_regions(self.ctx, _ptr(plane, _u8p), w, h, w, _ptr(rects, _i32p), n, _ptr(label, _i32p),
                                                _ptr(ss, _f64p), _ptr(ws, _f64p), _ptr(hist, _u8p) if want_hist else None))
        return label, ss, ws, hist

    def lbp_hist(self, plane, rects):
        plane = np.ascontiguousarray(plane, dtype=np.uint8)
        rects = np.ascontiguousarray(rects, dtype=np.int32).reshape(-1, 4)
        h, w = plane.shape
        n = rects.shape[0]
        hist = np.zeros((n, 1024), np.float64)
        self._check(self.L.ert_lbp_hist(self.ctx, _ptr(plane, _u8p), w, h, w, _ptr(rects, _i32p), n, _ptr(hist, _f64p)))
        return hist

    def cascade_predict(self, which, fv):
        fv = np.ascontiguousarray(fv, dtype=np.float64)
        n, d = fv.shape
        out = np.zeros(n)
        self._check(self.L.ert_cascade_predict_batch(self.ctx, which, _ptr(fv, _f64p), n, d, _ptr(out, _f64p)))
        return out

    def pinned_array(self, shape, dtype):
        """numpy array over page-locked host memory (ert_host_alloc); freed with the object"""
        shape = tuple(int(v) for v in (shape if isinstance(shape, (tuple, list)) else (shape,)))
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = self.L.ert_host_alloc(max(nbytes, 1))
        if not p:
            raise ErtError("ert_host_alloc(%d) failed: %s" % (nbytes, self.L.ert_last_error().decode()))
        buf = (C.c_uint8 * max(nbytes, 1)).from_address(p)
        arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
        self._pinned.append(p)
        return arr

    def cascade_classify_u8(self, hist, out=None):
        hist = np.ascontiguousarray(hist, dtype=np.uint8)
        n = hist.shape[0]
        label, ss, ws = out if out is not None else (np.zeros(n, np.int32), np.zeros(n), np.zeros(n))
        self._check(self.L.ert_cascade_classify_u8(self.ctx, _ptr(hist, _u8p), n, _ptr(label, _i32p), _ptr(ss, _f64p), _ptr(ws, _f64p)))
        return label, ss, ws

    def svm_predict_probability(self, x, out=None):
        k = self.L.ert_svm_nr_class(self.ctx)
        if k < 0:
            raise ErtError("svm model is not loaded")
        if x.dtype == np.uint8:
            x = np.ascontiguousarray(x)
            n = x.shape[0]
            label, prob = out if out is not None else (np.zeros(n), np.zeros((n, k)))
            self._check(self.L.ert_svm_predict_probability_batch_u8(self.ctx, _ptr(x, _u8p), n, _ptr(label, _f64p), _ptr(prob, _f64p)))
        else:
            x = np.ascontiguousarray(x, dtype=np.float64)
            n = x.shape[0]
            label = np.zeros(n); prob = np.zeros((n, k))
            self._check(self.L.ert_svm_predict_probability_batch(self.ctx, _ptr(x, _f64p), n, _ptr(label, _f64p), _ptr(prob, _f64p)))
        return label, prob

    # ---- after the detect path: er_track, OCR::chain_run ---------------------------------------
    def _unpack_track(self, tp):
        r = tp.contents
        F = r.n_frames
        coff = np.ctypeslib.as_array(r.cand_offset, (F + 1,)).copy()
        toff = np.ctypeslib.as_array(r.track_offset, (F + 1,)).copy()
        ns = np.ctypeslib.as_array(r.n_strong, (F,)).copy()
        nc, nt = int(coff[-1]), int(toff[-1])
        if nc:
            raw = np.ctypeslib.as_array(C.cast(r.cand, _u8p), (nc * 72,)).copy()
            cand = raw.view(TRACKED_DTYPE)
        else:
            cand = np.zeros(0, TRACKED_DTYPE)
        tracked = np.ctypeslib.as_array(r.tracked, (nt,)).copy() if nt else np.zeros(0, np.int32)
        out = [FrameTrack(cand[coff[f]:coff[f + 1]], int(ns[f]), tracked[toff[f]:toff[f + 1]]) for f in range(F)]
        return out, r.track_ms

    def er_track(self):
        """ERFilter::er_track on the BGR batch this context processed last -> ([FrameTrack per frame], device ms)."""
        tp = C.POINTER(ErtTrackResult)()
        self._check(self.L.ert_er_track(self.ctx, C.byref(tp)))
        return self._unpack_track(tp)

    def er_track_regions(self, bgr, strong, weak):
        """er_track on caller regions of one BGR frame; strong / weak = [n,6] rows (ch, x, y, w, h, area), channel-major."""
        bgr = np.ascontiguousarray(bgr, dtype=np.uint8)
        strong = np.ascontiguousarray(strong, dtype=np.int32).reshape(-1, 6)
        weak = np.ascontiguousarray(weak, dtype=np.int32).reshape(-1, 6)
        h, w, _ = bgr.shape
        tp = C.POINTER(ErtTrackResult)()
        self._check(self.L.ert_er_track_regions(self.ctx, _ptr(bgr, _u8p), w, h, w * 3, _ptr(strong, _i32p), len(strong),
                                                _ptr(weak, _i32p), len(weak), C.byref(tp)))
        return self._unpack_track(tp)[0][0]

    @staticmethod
    def _regions(rects, slopes=None, frames=None, planes=None):
        rects = np.asarray(rects, dtype=np.int32).reshape(-1, 4)
        reg = np.zeros(len(rects), OCR_REGION_DTYPE)
        reg["x"], reg["y"], reg["w"], reg["h"] = rects[:, 0], rects[:, 1], rects[:, 2], rects[:, 3]
        if slopes is not None:
            reg["slope"] = slopes
        if frames is not None:
            reg["frame"] = frames
        if planes is not None:
            reg["plane"] = planes
        return reg

    def _unpack_ocr(self, op):
        r = op.contents
        n, k = r.n, r.nr_class
        feat = np.ctypeslib.as_array(r.feat, (n * 1800,)).reshape(n, 1800).copy() if n else np.zeros((0, 1800), np.uint8)
        img = np.ctypeslib.as_array(r.img, (n * 900,)).reshape(n, 30, 30).copy() if n else np.zeros((0, 30, 30), np.uint8)
        if n and r.value:
            value = np.ctypeslib.as_array(r.value, (n,)).copy()
            label = np.ctypeslib.as_array(r.label, (n,)).copy()
            prob = np.ctypeslib.as_array(r.prob_all, (n * k,)).reshape(n, k).copy()
        else:
            value = label = prob = None
        return OcrResult(value, label, prob, feat, img, r.ocr_ms)

    def ocr_chain_run_plane(self, plane, rects, slopes=None):
        """OCR::chain_run on regions of one single-channel image: rects [n,4] (x,y,w,h), slopes [n] or None."""
        plane = np.ascontiguousarray(plane, dtype=np.uint8)
        h, w = plane.shape
        reg = self._regions(rects, slopes)
        op = C.POINTER(ErtOcrResult)()
        self._check(self.L.ert_ocr_chain_run_plane(self.ctx, _ptr(plane, _u8p), w, h, w, reg.ctypes.data, len(reg), C.byref(op)))
        return self._unpack_ocr(op)

    def ocr_features_plane(self, plane, rects, slopes=None):
        plane = np.ascontiguousarray(plane, dtype=np.uint8)
        h, w = plane.shape
        reg = self._regions(rects, slopes)
        op = C.POINTER(ErtOcrResult)()
        self._check(self.L.ert_ocr_features_plane(self.ctx, _ptr(plane, _u8p), w, h, w, reg.ctypes.data, len(reg), C.byref(op)))
        return self._unpack_ocr(op)

    def ocr_chain_run_batch(self, frames, planes, rects, slopes=None):
        """chain_run on regions of the last BGR batch (device-resident planes): frames [n], planes [n] (channel 0..5), rects [n,4]."""
        reg = self._regions(rects, slopes, frames, planes)
        op = C.POINTER(ErtOcrResult)()
        self._check(self.L.ert_ocr_chain_run_batch(self.ctx, reg.ctypes.data, len(reg), C.byref(op)))
        return self._unpack_ocr(op)

    def set_svm_legacy_prob(self, on):
        self._check(self.L.ert_set_svm_legacy_prob(self.ctx, int(on)))

    def set_svm_tensor_cores(self, on):
        self._check(self.L.ert_set_svm_tensor_cores(self.ctx, int(on)))

    def svm_dims(self):
        return self.L.ert_svm_dims(self.ctx)

    def bench_cascade_u8(self, hist, iters):
        hist = np.ascontiguousarray(hist, dtype=np.uint8)
        ms = C.c_double(0)
        self._check(self.L.ert_bench_cascade_u8(self.ctx, _ptr(hist, _u8p), hist.shape[0], iters, C.byref(ms)))
        return ms.value

    def bench_svm_u8(self, x, iters):
        x = np.ascontiguousarray(x, dtype=np.uint8)
        ms = C.c_double(0)
        self._check(self.L.ert_bench_svm_u8(self.ctx, _ptr(x, _u8p), x.shape[0], iters, C.byref(ms)))
        return ms.value
