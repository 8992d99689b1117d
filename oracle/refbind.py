"""ctypes binding of the oracles -- TEST INFRASTRUCTURE (the checker), never the product path.

Two shared objects live under oracle/:
  * oracle/_ref/libref_oracle.so  -- the reference's OWN code (built by oracle/build_ref.sh from
    /root/reference, git-ignored, travels to the GPU box as a prebuilt file)      kind="reference"
  * oracle/libert_port.so         -- the C restatement oracle/er_port.c            kind="port"
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes as C
import os
import lzma
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
ASSETS = os.path.join(ROOT, "assets", "classifier")

_u8p = C.POINTER(C.c_uint8)
_i32p = C.POINTER(C.c_int32)
_f64p = C.POINTER(C.c_double)


def _p(a, t):
    return a.ctypes.data_as(t)


def svm_model_path():
    """OCR.model is shipped xz-compressed (16 MB of text); unpack once next to it."""
    dst = os.path.join(ASSETS, "OCR.model")
    if not os.path.exists(dst):
        with lzma.open(dst + ".xz", "rb") as f, open(dst + ".tmp", "wb") as g:
            g.write(f.read())
        os.replace(dst + ".tmp", dst)
    return dst


DEFAULTS = dict(thresh_step=8, min_area=120, max_area=900000, stability_t=2, overlap_coef=0.7)


class RefOracle:
    """The reference's own ERFilter / CascadeBoost / libsvm code behind a C wrapper (oracle/ref_capi.cpp)."""

    kind = "reference"

    def __init__(self, with_svm=False, **params):
        path = os.path.join(HERE, "_ref", "libref_oracle.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path + " (run oracle/build_ref.sh where /root/reference exists)")
        L = self.L = C.CDLL(path)
        p = dict(DEFAULTS); p.update(params); self.params = p
        L.ref_create.restype = C.c_void_p
        L.ref_create.argtypes = [C.c_int] * 4 + [C.c_double, C.c_char_p, C.c_char_p, C.c_char_p]
        L.ref_tree_extract.restype = C.c_void_p
        L.ref_tree_extract.argtypes = [C.c_void_p, _u8p, C.c_int, C.c_int, C.c_int]
        L.ref_tree_size.argtypes = [C.c_void_p]
        L.ref_tree_dump.argtypes = [C.c_void_p, _i32p]
        L.ref_nms.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_pool_indices.argtypes = [C.c_void_p, _i32p]
        L.ref_classify.argtypes = [C.c_void_p, C.c_void_p, _i32p, _f64p, _f64p]
        L.ref_tree_free.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_lbp_hist.argtypes = [C.c_void_p, _u8p, C.c_int, C.c_int, C.c_int, _f64p]
        L.ref_aran.argtypes = [C.c_void_p, _u8p, C.c_int, C.c_int, C.c_int, C.c_int, _u8p]
        L.ref_resize.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _u8p]
        L.ref_divide.argtypes = [_u8p, C.c_int, C.c_int, _u8p]
        L.ref_cascade_predict.argtypes = [C.c_void_p, C.c_int, _f64p, C.c_int, C.c_int, _f64p]
        L.ref_cascade_num_stumps.argtypes = [C.c_void_p, C.c_int]
        L.ref_svm_nr_class.argtypes = [C.c_void_p]
        L.ref_svm_total_sv.argtypes = [C.c_void_p]
        L.ref_svm_labels.argtypes = [C.c_void_p, _i32p]
        L.ref_svm_predict_probability.argtypes = [C.c_void_p, _f64p, C.c_int, C.c_int, C.c_int, _f64p, _f64p]
        L.ref_detect_frames.restype = C.c_double
        L.ref_detect_frames.argtypes = [C.c_void_p, _u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _i32p, _f64p]
        L.ref_channels.argtypes = [_u8p, C.c_int, C.c_int, _u8p]
        L.ref_destroy.argtypes = [C.c_void_p]
        # rows after the detect path (oracle/ref_capi_next.cpp)
        L.ref_prim_threshold_otsu.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, _u8p]
        L.ref_prim_find_contours.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, _i32p, C.c_int, _i32p, C.c_int]
        L.ref_prim_gaussian7.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, _u8p]
        L.ref_prim_normalize_minmax.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, _u8p]
        L.ref_calc_color.argtypes = [_u8p, _u8p, C.c_int, C.c_int, _i32p, C.c_int, _f64p]
        L.ref_er_track.argtypes = [C.c_void_p, _u8p, _u8p, C.c_int, C.c_int, _i32p, C.c_int, _i32p, C.c_int, _i32p, _f64p, _f64p, _i32p, _i32p]
        L.ref_er_grouping.argtypes = [C.c_void_p, _f64p, C.c_int, C.c_int, C.c_int, C.c_int, _i32p, _i32p, _i32p, _i32p, _i32p, C.c_int, _f64p, C.c_int]
        L.ref_chain_run.restype = C.c_double
        L.ref_chain_run.argtypes = [C.c_void_p, _u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double]
        L.ref_ocr_features.argtypes = [C.c_void_p, _u8p, C.c_int, C.c_int, C.c_int, C.c_double, _u8p, _u8p]
        L.ref_rotate_mat.argtypes = [C.c_void_p, _u8p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, _u8p, C.c_int]
        svm = svm_model_path().encode() if with_svm else None
        self.ctx = L.ref_create(p["thresh_step"], p["min_area"], p["max_area"], p["stability_t"], p["overlap_coef"],
                                os.path.join(ASSETS, "strong.classifier").encode(),
                                os.path.join(ASSETS, "weak.classifier").encode(), svm)

    def close(self):
        if self.ctx:
            self.L.ref_destroy(self.ctx)
            self.ctx = None

    # -- stage functions -------------------------------------------------------------------
    def plane(self, plane, classify=True, scores=False):
        """er_tree_extract -> nms -> classify on one u8 plane.
        Returns dict(nodes[n,8]=level,area,x,y,w,h,parent,nchild (DFS preorder, reference child order),
                     pool[m] node indices, label[m], strong_score, weak_score)."""
        plane = np.ascontiguousarray(plane, dtype=np.uint8)
        h, w = plane.shape
        t = self.L.ref_tree_extract(self.ctx, _p(plane, _u8p), w, h, w)
        n = self.L.ref_tree_size(t)
        nodes = np.zeros((n, 8), np.int32)
        self.L.ref_tree_dump(t, _p(nodes, _i32p))
        m = self.L.ref_nms(self.ctx, t)
        pool = np.zeros(m, np.int32)
        self.L.ref_pool_indices(t, _p(pool, _i32p))
        out = dict(nodes=nodes, pool=pool)
        if classify:
            label = np.zeros(m, np.int32)
            ss = np.zeros(m, np.float64); ws = np.zeros(m, np.float64)
            self.L.ref_classify(self.ctx, t, _p(label, _i32p), _p(ss, _f64p) if scores else None,
                                _p(ws, _f64p) if scores else None)
            out.update(label=label, strong_score=ss, weak_score=ws)
        self.L.ref_tree_free(self.ctx, t)
        return out

    def channels(self, bgr):
        bgr = np.ascontiguousarray(bgr, dtype=np.uint8)
        h, w, _ = bgr.shape
        out = np.zeros((6, h, w), np.uint8)
        self.L.ref_channels(_p(bgr, _u8p), w, h, _p(out, _u8p))
        return out

    def lbp_hist(self, crop):
        crop = np.ascontiguousarray(crop, dtype=np.uint8)
        h, w = crop.shape
        out = np.zeros(1024, np.float64)
        self.L.ref_lbp_hist(self.ctx, _p(crop, _u8p), w, h, w, _p(out, _f64p))
        return out

    def aran(self, crop, L=26):
        crop = np.ascontiguousarray(crop, dtype=np.uint8)
        h, w = crop.shape
        out = np.zeros((L, L), np.uint8)
        self.L.ref_aran(self.ctx, _p(crop, _u8p), w, h, w, L, _p(out, _u8p))
        return out

    def resize(self, src, dw, dh):
        src = np.ascontiguousarray(src, dtype=np.uint8)
        h, w = src.shape
        out = np.zeros((dh, dw), np.uint8)
        self.L.ref_resize(_p(src, _u8p), w, h, w, dw, dh, _p(out, _u8p))
        return out

    def divide(self, vals, step):
        vals = np.ascontiguousarray(vals, dtype=np.uint8)
        out = np.zeros_like(vals)
        self.L.ref_divide(_p(vals, _u8p), vals.size, step, _p(out, _u8p))
        return out

    def cascade_predict(self, which, fv):
        fv = np.ascontiguousarray(fv, dtype=np.float64)
        n, d = fv.shape
        out = np.zeros(n, np.float64)
        self.L.ref_cascade_predict(self.ctx, which, _p(fv, _f64p), n, d, _p(out, _f64p))
        return out

    def svm_predict_probability(self, x, nthreads=1):
        x = np.ascontiguousarray(x, dtype=np.float64)
        n, d = x.shape
        k = self.L.ref_svm_nr_class(self.ctx)
        label = np.zeros(n, np.float64); prob = np.zeros((n, k), np.float64)
        self.L.ref_svm_predict_probability(self.ctx, _p(x, _f64p), n, d, nthreads, _p(label, _f64p), _p(prob, _f64p))
        return label, prob

    # -- rows after the detect path: er_track / OCR::chain_run ---------------------------------
    def prim_threshold_otsu(self, img):
        img = np.ascontiguousarray(img, dtype=np.uint8)
        h, w = img.shape
        out = np.zeros_like(img)
        t = self.L.ref_prim_threshold_otsu(_p(img, _u8p), w, h, w, _p(out, _u8p))
        return t, out

    def prim_find_contours(self, img):
        img = np.ascontiguousarray(img, dtype=np.uint8)
        h, w = img.shape
        cap = 8 * h * w + 16
        pts = np.zeros((cap, 2), np.int32); sizes = np.zeros(h * w + 4, np.int32)
        n = self.L.ref_prim_find_contours(_p(img, _u8p), w, h, w, _p(pts, _i32p), cap, _p(sizes, _i32p), sizes.size)
        out, o = [], 0
        for k in range(n):
            out.append(pts[o:o + sizes[k]].copy()); o += sizes[k]
        return out

    def prim_gaussian7(self, img):
        img = np.ascontiguousarray(img, dtype=np.uint8)
        h, w = img.shape
        out = np.zeros_like(img)
        self.L.ref_prim_gaussian7(_p(img, _u8p), w, h, w, _p(out, _u8p))
        return out

    def prim_normalize_minmax(self, img):
        img = np.ascontiguousarray(img, dtype=np.uint8)
        h, w = img.shape
        out = np.zeros_like(img)
        self.L.ref_prim_normalize_minmax(_p(img, _u8p), w, h, w, _p(out, _u8p))
        return out

    def calc_color(self, plane, ycrcb, rects):
        """calc_color (src/ER.cpp:1391-1437): plane [H,W] u8 (the ER's channel), ycrcb [H,W,3], rects [n,4] -> [n,3] f64."""
        plane = np.ascontiguousarray(plane, dtype=np.uint8)
        ycrcb = np.ascontiguousarray(ycrcb, dtype=np.uint8)
        rects = np.ascontiguousarray(rects, dtype=np.int32).reshape(-1, 4)
        h, w = plane.shape
        out = np.zeros((len(rects), 3), np.float64)
        self.L.ref_calc_color(_p(plane, _u8p), _p(ycrcb, _u8p), w, h, _p(rects, _i32p), len(rects), _p(out, _f64p))
        return out

    def er_track(self, planes6, ycrcb, strong, weak):
        """ERFilter::er_track (src/ER.cpp:532-609).  strong / weak = [n,6] rows (ch,x,y,w,h,area), channel-major.
        Returns dict(tracked [m,2] = (kind 0 strong / 1 weak, row), strong_color, weak_color, strong_center, weak_center)."""
        planes6 = np.ascontiguousarray(planes6, dtype=np.uint8)
        ycrcb = np.ascontiguousarray(ycrcb, dtype=np.uint8)
        strong = np.ascontiguousarray(strong, dtype=np.int32).reshape(-1, 6)
        weak = np.ascontiguousarray(weak, dtype=np.int32).reshape(-1, 6)
        _, h, w = planes6.shape
        ns, nw = len(strong), len(weak)
        tr = np.zeros((ns + nw + 1, 2), np.int32)
        sc = np.zeros((ns + 1, 3)); wc = np.zeros((nw + 1, 3))
        sce = np.zeros((ns + 1, 2), np.int32); wce = np.zeros((nw + 1, 2), np.int32)
        m = self.L.ref_er_track(self.ctx, _p(planes6, _u8p), _p(ycrcb, _u8p), w, h, _p(strong, _i32p), ns, _p(weak, _i32p), nw,
                                _p(tr, _i32p), _p(sc, _f64p), _p(wc, _f64p), _p(sce, _i32p), _p(wce, _i32p))
        return dict(tracked=tr[:m].copy(), strong_color=sc[:ns], weak_color=wc[:nw], strong_center=sce[:ns], weak_center=wce[:nw])

    def er_grouping(self, rows, overlap_sup=False, inner_sup=True, dedupe=False):
        """ERFilter::er_grouping (src/ER.cpp:612-692) [+ er_ocr's duplicate removal, src/ER.cpp:702-724].  rows [n,11] =
        ch, x, y, w, h, area, cx, cy, color1..3.  Returns dict(after = all_er afterwards (input indices), bounds [n,6] =
        x, y, w, h, cx, cy afterwards, texts = [(slope, [input indices])])."""
        rows = np.ascontiguousarray(rows, dtype=np.float64).reshape(-1, 11)
        n = len(rows)
        n_after = C.c_int32(0)
        after = np.zeros(n + 1, np.int32); bounds = np.zeros((n + 1, 6), np.int32)
        tcap, ecap = 4 * n + 4, 64 * n + 64
        toff = np.zeros(tcap + 1, np.int32); ters = np.zeros(ecap, np.int32); tsl = np.zeros(tcap, np.float64)
        nt = self.L.ref_er_grouping(self.ctx, _p(rows, _f64p), n, int(overlap_sup), int(inner_sup), int(dedupe), C.byref(n_after), _p(after, _i32p),
                                    _p(bounds, _i32p), _p(toff, _i32p), _p(ters, _i32p), ecap, _p(tsl, _f64p), tcap)
        if nt < 0:
            raise ValueError("er_grouping output capacity exceeded")
        texts = [(float(tsl[t]), ters[toff[t]:toff[t + 1]].tolist()) for t in range(nt)]
        return dict(after=after[:n_after.value].tolist(), bounds=bounds[:n].copy(), texts=texts)

    def chain_run(self, crop, thresh=0, slope=0.0):
        """OCR::chain_run verbatim (src/OCR.cpp:67-140): returns table[label] + prob[label]."""
        crop = np.ascontiguousarray(crop, dtype=np.uint8)
        h, w = crop.shape
        return self.L.ref_chain_run(self.ctx, _p(crop, _u8p), w, h, w, thresh, slope)

    def ocr_features(self, crop, slope=0.0):
        """chain_run's pre-processing + extract_feature: (30x30 image, 1800 feature bytes = value*255)."""
        crop = np.ascontiguousarray(crop, dtype=np.uint8)
        h, w = crop.shape
        img = np.zeros((30, 30), np.uint8); feat = np.zeros(1800, np.uint8)
        n = self.L.ref_ocr_features(self.ctx, _p(crop, _u8p), w, h, w, slope, _p(img, _u8p), _p(feat, _u8p))
        if n < 0:
            raise ValueError("empty image after rotation")
        return img, feat

    def rotate_mat(self, img, rad, crop=True):
        img = np.ascontiguousarray(img, dtype=np.uint8)
        h, w = img.shape
        cap = 16 * (h + w + 4) * (h + w + 4)
        out = np.zeros(cap, np.uint8)
        r = self.L.ref_rotate_mat(self.ctx, _p(img, _u8p), w, h, w, rad, 1 if crop else 0, _p(out, _u8p), cap)
        if r < 0:
            raise ValueError("rotate_mat output too large")
        rows, cols = r >> 16, r & 0xffff
        return out[:rows * cols].reshape(rows, cols).copy()

    def detect_frames(self, bgr, mode=1, nthreads=1):
        """bgr [F,H,W,3] -> (wall seconds, counts[F,4]=kept,pool,strong,weak, stage seconds[3])."""
        bgr = np.ascontiguousarray(bgr, dtype=np.uint8)
        f, h, w, _ = bgr.shape
        counts = np.zeros((f, 4), np.int32); st = np.zeros(3, np.float64)
        sec = self.L.ref_detect_frames(self.ctx, _p(bgr, _u8p), f, w, h, mode, nthreads, _p(counts, _i32p), _p(st, _f64p))
        return sec, counts, st


class PortOracle:
    """The C restatement (oracle/er_port.c).  Same call surface as RefOracle."""

    kind = "port"

    def __init__(self, with_svm=False, **params):
        path = os.path.join(HERE, "libert_port.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path + " (run `make -C oracle libert_port.so`)")
        L = self.L = C.CDLL(path)
        p = dict(DEFAULTS); p.update(params); self.params = p
        L.port_channels.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, _u8p]
        L.port_quantize.argtypes = [_u8p, C.c_int, C.c_int, _u8p]
        L.port_tree_extract.restype = C.c_void_p
        L.port_tree_extract.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.port_tree_size.argtypes = [C.c_void_p]
        L.port_tree_dump.argtypes = [C.c_void_p, _i32p]
        L.port_nms.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double]
        L.port_pool_indices.argtypes = [C.c_void_p, _i32p]
        L.port_tree_free.argtypes = [C.c_void_p]
        L.port_resize.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _u8p]
        L.port_aran_minor.argtypes = [C.c_int, C.c_int, C.c_int]
        L.port_aran.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, C.c_int, _u8p]
        L.port_lbp_hist.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, _f64p]
        L.port_cascade_load.restype = C.c_void_p
        L.port_cascade_load.argtypes = [C.c_char_p]
        L.port_cascade_info.argtypes = [C.c_void_p, _i32p, _i32p, _i32p]
        L.port_cascade_predict_batch.argtypes = [C.c_void_p, _f64p, C.c_int, C.c_int, _f64p]
        L.port_cascade_free.argtypes = [C.c_void_p]
        L.port_classify.argtypes = [C.c_void_p, _u8p, C.c_int, C.c_void_p, C.c_void_p, _i32p, _f64p, _f64p]
        L.port_svm_load.restype = C.c_void_p
        L.port_svm_load.argtypes = [C.c_char_p]
        L.port_svm_info.argtypes = [C.c_void_p, _i32p, _i32p, _f64p, _i32p]
        L.port_svm_dense_sv.argtypes = [C.c_void_p, C.c_int, _f64p]
        L.port_svm_predict_probability.restype = C.c_double
        L.port_svm_predict_probability.argtypes = [C.c_void_p, _f64p, C.c_int, _f64p, _f64p, _f64p]
        L.port_svm_free.argtypes = [C.c_void_p]
        L.port_canonical_nodes.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _i32p, C.c_int]
        L.port_sort_children.argtypes = [C.c_void_p, C.c_int]
        L.port_calc_color.argtypes = [_u8p, _u8p, C.c_int, C.c_int, _i32p, C.c_int, _f64p]
        L.port_er_track.argtypes = [_u8p, _u8p, C.c_int, C.c_int, _i32p, C.c_int, _i32p, C.c_int, _i32p, _f64p, _f64p, _i32p, _i32p]
        L.port_ocr_features.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, C.c_double, _u8p, _u8p]
        self.casc = [L.port_cascade_load(os.path.join(ASSETS, "strong.classifier").encode()),
                     L.port_cascade_load(os.path.join(ASSETS, "weak.classifier").encode())]
        self.svm = L.port_svm_load(svm_model_path().encode()) if with_svm else None

    def close(self):
        for c in self.casc:
            self.L.port_cascade_free(c)
        self.casc = []
        if self.svm:
            self.L.port_svm_free(self.svm)
            self.svm = None

    def plane(self, plane, classify=True, scores=False, canonical_order=False):
        """canonical_order=True re-orders every child list the way the GPU path visits siblings
        (descending bbox.y, bbox.x, level, area) before NMS -- the order-independent statement of the path."""
        plane = np.ascontiguousarray(plane, dtype=np.uint8)
        h, w = plane.shape
        p = self.params
        t = self.L.port_tree_extract(_p(plane, _u8p), w, h, w, p["thresh_step"], p["min_area"])
        if canonical_order:
            self.L.port_sort_children(t, 9)
        n = self.L.port_tree_size(t)
        nodes = np.zeros((n, 8), np.int32)
        self.L.port_tree_dump(t, _p(nodes, _i32p))
        m = self.L.port_nms(t, p["min_area"], p["max_area"], p["stability_t"], p["overlap_coef"])
        pool = np.zeros(m, np.int32)
        self.L.port_pool_indices(t, _p(pool, _i32p))
        out = dict(nodes=nodes, pool=pool)
        if classify:
            label = np.zeros(m, np.int32)
            ss = np.zeros(m, np.float64); ws = np.zeros(m, np.float64)
            self.L.port_classify(t, _p(plane, _u8p), w, self.casc[0], self.casc[1], _p(label, _i32p), _p(ss, _f64p), _p(ws, _f64p))
            out.update(label=label, strong_score=ss, weak_score=ws)
        self.L.port_tree_free(t)
        return out

    # -- rows after the detect path (same call surface as RefOracle) ----------------------------
    def calc_color(self, plane, ycrcb, rects):
        plane = np.ascontiguousarray(plane, dtype=np.uint8)
        ycrcb = np.ascontiguousarray(ycrcb, dtype=np.uint8)
        rects = np.ascontiguousarray(rects, dtype=np.int32).reshape(-1, 4)
        h, w = plane.shape
        out = np.zeros((len(rects), 3), np.float64)
        self.L.port_calc_color(_p(plane, _u8p), _p(ycrcb, _u8p), w, h, _p(rects, _i32p), len(rects), _p(out, _f64p))
        return out

    def er_track(self, planes6, ycrcb, strong, weak):
        planes6 = np.ascontiguousarray(planes6, dtype=np.uint8)
        ycrcb = np.ascontiguousarray(ycrcb, dtype=np.uint8)
        strong = np.ascontiguousarray(strong, dtype=np.int32).reshape(-1, 6)
        weak = np.ascontiguousarray(weak, dtype=np.int32).reshape(-1, 6)
        _, h, w = planes6.shape
        ns, nw = len(strong), len(weak)
        tr = np.zeros((ns + nw + 1, 2), np.int32)
        sc = np.zeros((ns + 1, 3)); wc = np.zeros((nw + 1, 3))
        sce = np.zeros((ns + 1, 2), np.int32); wce = np.zeros((nw + 1, 2), np.int32)
        m = self.L.port_er_track(_p(planes6, _u8p), _p(ycrcb, _u8p), w, h, _p(strong, _i32p), ns, _p(weak, _i32p), nw,
                                 _p(tr, _i32p), _p(sc, _f64p), _p(wc, _f64p), _p(sce, _i32p), _p(wce, _i32p))
        return dict(tracked=tr[:m].copy(), strong_color=sc[:ns], weak_color=wc[:nw], strong_center=sce[:ns], weak_center=wce[:nw])

    def ocr_features(self, crop, slope=0.0):
        crop = np.ascontiguousarray(crop, dtype=np.uint8)
        h, w = crop.shape
        img = np.zeros((30, 30), np.uint8); feat = np.zeros(1800, np.uint8)
        n = self.L.port_ocr_features(_p(crop, _u8p), w, h, w, slope, _p(img, _u8p), _p(feat, _u8p))
        if n < 0:
            raise ValueError("empty image after rotation / ARAN")
        return img, feat

    def canonical_nodes(self, plane, min_area=None):
        plane = np.ascontiguousarray(plane, dtype=np.uint8)
        h, w = plane.shape
        p = self.params
        ma = p["min_area"] if min_area is None else min_area
        cap = h * w + 8
        out = np.zeros((cap, 6), np.int32)
        n = self.L.port_canonical_nodes(_p(plane, _u8p), w, h, w, p["thresh_step"], ma, _p(out, _i32p), cap)
        return out[:n].copy()

    def channels(self, bgr):
        bgr = np.ascontiguousarray(bgr, dtype=np.uint8)
        h, w, _ = bgr.shape
        out = np.zeros((6, h, w), np.uint8)
        self.L.port_channels(_p(bgr, _u8p), w, h, w * 3, _p(out, _u8p))
        return out

    def quantize(self, vals, step):
        vals = np.ascontiguousarray(vals, dtype=np.uint8)
        out = np.zeros_like(vals)
        self.L.port_quantize(_p(vals, _u8p), vals.size, step, _p(out, _u8p))
        return out

    def lbp_hist(self, crop):
        crop = np.ascontiguousarray(crop, dtype=np.uint8)
        h, w = crop.shape
        out = np.zeros(1024, np.float64)
        self.L.port_lbp_hist(_p(crop, _u8p), w, h, w, _p(out, _f64p))
        return out

    def aran(self, crop, L=26):
        crop = np.ascontiguousarray(crop, dtype=np.uint8)
        h, w = crop.shape
        out = np.zeros((L, L), np.uint8)
        self.L.port_aran(_p(crop, _u8p), w, h, w, L, _p(out, _u8p))
        return out

    def resize(self, src, dw, dh):
        src = np.ascontiguousarray(src, dtype=np.uint8)
        h, w = src.shape
        out = np.zeros((dh, dw), np.uint8)
        self.L.port_resize(_p(src, _u8p), w, h, w, dw, dh, _p(out, _u8p))
        return out

    def cascade_info(self, which):
        ns = C.c_int32(0)
        sl = np.zeros(64, np.int32); st = np.zeros(64, np.int32)
        n = self.L.port_cascade_info(self.casc[which], C.byref(ns), _p(sl, _i32p), _p(st, _i32p))
        return n, sl[:ns.value].copy(), st[:ns.value].copy()

    def cascade_predict(self, which, fv):
        fv = np.ascontiguousarray(fv, dtype=np.float64)
        n, d = fv.shape
        out = np.zeros(n, np.float64)
        self.L.port_cascade_predict_batch(self.casc[which], _p(fv, _f64p), n, d, _p(out, _f64p))
        return out

    def svm_predict_probability(self, x, nthreads=1, want_kvalue=False):
        x = np.ascontiguousarray(x, dtype=np.float64)
        n, d = x.shape
        k = C.c_int32(0); l = C.c_int32(0)
        self.L.port_svm_info(self.svm, C.byref(k), C.byref(l), None, None)
        k, l = k.value, l.value
        label = np.zeros(n, np.float64); prob = np.zeros((n, k), np.float64)
        kv = np.zeros((n, l), np.float64) if want_kvalue else None
        for i in range(n):
            label[i] = self.L.port_svm_predict_probability(self.svm, _p(x[i], _f64p), d, _p(prob[i], _f64p),
                                                           _p(kv[i], _f64p) if want_kvalue else None, None)
        return (label, prob, kv) if want_kvalue else (label, prob)


def best_oracle(**kw):
    """The reference-backed oracle when its prebuilt .so is present, else the port."""
    try:
        return RefOracle(**kw)
    except (FileNotFoundError, OSError):
        return PortOracle(**kw)
