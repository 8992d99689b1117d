// oracle/ref_capi.cpp -- TEST INFRASTRUCTURE (the checker), never linked into the product.
//
// A plain-C wrapper around the reference's OWN hot-path functions.  oracle/build_ref.sh
// extracts the function bodies by line range from /root/reference/src/ER.cpp and
// src/OCR.cpp into oracle/_ref/ref_hotpath.cpp (git-ignored, never committed), compiles
// them unmodified against oracle/cvshim + the reference's own headers, together with the
// unmodified src/adaboost.cpp and src/svm.cpp, and links this file on top.
//
// What is exposed (all results are flattened to arrays so python/ctypes can compare):
//   ref_tree_*      ERFilter::er_tree_extract            (src/ER.cpp:240-413)
//   ref_nms         ERFilter::non_maximum_supression     (src/ER.cpp:416-505)
//   ref_classify    ERFilter::classify                   (src/ER.cpp:507-528)
//   ref_lbp_hist    ERFilter::make_LBP_hist              (src/ER.cpp:789-845, src/OCR.cpp:394-430)
//   ref_cascade_*   CascadeBoost::predict                (src/adaboost.cpp:507-542)
//   ref_svm_*       svm_predict_probability              (src/svm.cpp:2592-2629)
//   ref_detect_frames  the per-channel loop of ERFilter::text_detect (src/ER.cpp:50-60), timed
#include "ER.h"
#include <omp.h>
#include <chrono>

namespace {

struct RefCtx {
	ERFilter *erf;
	CascadeBoost *stc;
	CascadeBoost *wtc;
	OCR *ocr;
	svm_model *svm;
};

struct RefTree {
	ER *root;
	std::vector<ER *> nodes;          // DFS pre-order, children in the reference's list order
	std::vector<int> parent;          // index into nodes, -1 for root
	std::vector<ER *> pool, strong, weak;
	cv::Mat plane;                    // owning copy of the input plane
};

void flatten(RefTree *t)
{
	// iterative pre-order following child / next exactly as the reference links them
	t->nodes.clear(); t->parent.clear();
	std::vector<std::pair<ER *, int> > st;
	st.push_back(std::make_pair(t->root, -1));
	while (!st.empty()) {
		std::pair<ER *, int> cur = st.back(); st.pop_back();
		int idx = (int)t->nodes.size();
		t->nodes.push_back(cur.first);
		t->parent.push_back(cur.second);
		// push children in reverse so that the first child is visited first
		std::vector<ER *> ch;
		for (ER *c = cur.first->child; c; c = c->next) ch.push_back(c);
		for (int i = (int)ch.size() - 1; i >= 0; i--) st.push_back(std::make_pair(ch[i], idx));
	}
}

int index_of(const RefTree *t, const ER *e)
{
	for (size_t i = 0; i < t->nodes.size(); i++) if (t->nodes[i] == e) return (int)i;
	return -1;
}

// BGR -> the six planes of ERFilter::compute_channels (src/ER.cpp:114-128).  cvtColor is not
// available in the shim; this is the integer restatement of OpenCV's 8-bit BGR2YCrCb
// (pinned against cv2 in tests/test_oracle.py).
void channels_from_bgr(const uchar *bgr, int w, int h, int stride, std::vector<cv::Mat> &ch)
{
	ch.clear();
	for (int k = 0; k < 6; k++) ch.push_back(cv::Mat(h, w, CV_8UC1));
	for (int y = 0; y < h; y++) {
		const uchar *p = bgr + (size_t)y * stride;
		for (int x = 0; x < w; x++) {
			int B = p[3 * x], G = p[3 * x + 1], R = p[3 * x + 2];
			int Y = (R * 4899 + G * 9617 + B * 1868 + 8192) >> 14;
			int Cr = ((R - Y) * 11682 + (128 << 14) + 8192) >> 14;
			int Cb = ((B - Y) * 9241 + (128 << 14) + 8192) >> 14;
			Cr = Cr < 0 ? 0 : (Cr > 255 ? 255 : Cr);
			Cb = Cb < 0 ? 0 : (Cb > 255 ? 255 : Cb);
			ch[0].ptr(y)[x] = (uchar)Y; ch[1].ptr(y)[x] = (uchar)Cr; ch[2].ptr(y)[x] = (uchar)Cb;
			ch[3].ptr(y)[x] = (uchar)(255 - Y); ch[4].ptr(y)[x] = (uchar)(255 - Cr); ch[5].ptr(y)[x] = (uchar)(255 - Cb);
		}
	}
}

} // namespace

extern "C" {

void *ref_create(int thresh_step, int min_area, int max_area, int stability_t, double overlap_coef,
                 const char *strong_path, const char *weak_path, const char *svm_path)
{
	RefCtx *c = new RefCtx();
	c->erf = new ERFilter(thresh_step, min_area, max_area, stability_t, overlap_coef, 0.15);
	c->stc = strong_path ? new CascadeBoost(strong_path) : nullptr;
	c->wtc = weak_path ? new CascadeBoost(weak_path) : nullptr;
	c->ocr = new OCR();
	c->svm = svm_path ? svm_load_model(svm_path) : nullptr;
	c->erf->stc = c->stc; c->erf->wtc = c->wtc; c->erf->ocr = c->ocr;
	return c;
}

int ref_cascade_num_stumps(void *ctx, int which)
{
	RefCtx *c = (RefCtx *)ctx;
	CascadeBoost *b = which == 0 ? c->stc : c->wtc;
	return b ? b->get_num_iter() : -1;
}

// ---- tree ------------------------------------------------------------------------------------
void *ref_tree_extract(void *ctx, const uchar *plane, int w, int h, int stride)
{
	RefCtx *c = (RefCtx *)ctx;
	RefTree *t = new RefTree();
	cv::Mat view(h, w, CV_8UC1, (void *)plane, (size_t)stride);
	t->plane = view.clone();
	t->root = c->erf->er_tree_extract(t->plane);
	flatten(t);
	return t;
}

int ref_tree_size(void *tree) { return (int)((RefTree *)tree)->nodes.size(); }

// out: n x 8 int32 = level, area, x, y, w, h, parent(index in this dump), n_children
void ref_tree_dump(void *tree, int *out)
{
	RefTree *t = (RefTree *)tree;
	for (size_t i = 0; i < t->nodes.size(); i++) {
		ER *e = t->nodes[i];
		int nc = 0;
		for (ER *k = e->child; k; k = k->next) nc++;
		int *o = out + 8 * i;
		o[0] = e->level; o[1] = e->area; o[2] = e->bound.x; o[3] = e->bound.y;
		o[4] = e->bound.width; o[5] = e->bound.height; o[6] = t->parent[i]; o[7] = nc;
	}
}

// runs NMS (mutates done/stability like the reference), returns pool size
int ref_nms(void *ctx, void *tree)
{
	RefCtx *c = (RefCtx *)ctx; RefTree *t = (RefTree *)tree;
	ERs all;
	t->pool.clear();
	c->erf->non_maximum_supression(t->root, all, t->pool, t->plane);
	return (int)t->pool.size();
}

void ref_pool_indices(void *tree, int *out)
{
	RefTree *t = (RefTree *)tree;
	for (size_t i = 0; i < t->pool.size(); i++) out[i] = index_of(t, t->pool[i]);
}

// classify the pool; label per pool entry: 2 strong, 1 weak, 0 rejected.  Also the raw
// predict() return values of both cascades (weak evaluated for every entry for the test's benefit).
void ref_classify(void *ctx, void *tree, int *label, double *strong_score, double *weak_score)
{
	RefCtx *c = (RefCtx *)ctx; RefTree *t = (RefTree *)tree;
	t->strong.clear(); t->weak.clear();
	c->erf->classify(t->pool, t->strong, t->weak, t->plane);
	size_t si = 0, wi = 0;
	for (size_t i = 0; i < t->pool.size(); i++) {
		label[i] = 0;
		if (si < t->strong.size() && t->strong[si] == t->pool[i]) { label[i] = 2; si++; }
		else if (wi < t->weak.size() && t->weak[wi] == t->pool[i]) { label[i] = 1; wi++; }
		if (strong_score || weak_score) {
			vector<double> fv = c->erf->make_LBP_hist(t->plane(t->pool[i]->bound), 2, 24);
			if (strong_score) strong_score[i] = c->stc->predict(fv);
			if (weak_score) weak_score[i] = c->wtc->predict(fv);
		}
	}
}

void ref_tree_free(void *ctx, void *tree)
{
	RefCtx *c = (RefCtx *)ctx; RefTree *t = (RefTree *)tree;
	c->erf->er_delete(t->root);
	delete t;
}

// ---- features / classifiers ------------------------------------------------------------------
void ref_lbp_hist(void *ctx, const uchar *crop, int w, int h, int stride, double *hist1024)
{
	RefCtx *c = (RefCtx *)ctx;
	cv::Mat view(h, w, CV_8UC1, (void *)crop, (size_t)stride);
	vector<double> fv = c->erf->make_LBP_hist(view, 2, 24);
	for (int i = 0; i < 1024; i++) hist1024[i] = fv[i];
}

// ARAN-normalised 26x26 patch (the contiguous buffer calc_LBP reads), for the resize parity test
void ref_aran(void *ctx, const uchar *crop, int w, int h, int stride, int L, uchar *outLL)
{
	RefCtx *c = (RefCtx *)ctx;
	cv::Mat view(h, w, CV_8UC1, (void *)crop, (size_t)stride), dst;
	c->ocr->ARAN(view, dst, L);
	for (int i = 0; i < L; i++) memcpy(outLL + (size_t)i * L, dst.ptr(i), (size_t)L);
}

void ref_resize(const uchar *src, int sw, int sh, int stride, int dw, int dh, uchar *out)
{
	cv::Mat view(sh, sw, CV_8UC1, (void *)src, (size_t)stride), dst;
	cv::resize(view, dst, cv::Size(dw, dh));
	for (int i = 0; i < dh; i++) memcpy(out + (size_t)i * dw, dst.ptr(i), (size_t)dw);
}

void ref_divide(const uchar *src, int n, int step, uchar *out)
{
	cv::Mat m(1, n, CV_8UC1);
	memcpy(m.data, src, (size_t)n);
	m /= step;
	memcpy(out, m.data, (size_t)n);
}

void ref_cascade_predict(void *ctx, int which, const double *fv, int n, int dims, double *score)
{
	RefCtx *c = (RefCtx *)ctx;
	CascadeBoost *b = which == 0 ? c->stc : c->wtc;
	for (int i = 0; i < n; i++) {
		vector<double> v(fv + (size_t)i * dims, fv + (size_t)(i + 1) * dims);
		score[i] = b->predict(v);
	}
}

int ref_svm_nr_class(void *ctx) { RefCtx *c = (RefCtx *)ctx; return c->svm ? svm_get_nr_class(c->svm) : -1; }
int ref_svm_total_sv(void *ctx) { RefCtx *c = (RefCtx *)ctx; return c->svm ? c->svm->l : -1; }
void ref_svm_labels(void *ctx, int *labels) { RefCtx *c = (RefCtx *)ctx; svm_get_labels(c->svm, labels); }

// x: n x dims dense doubles; zeros are omitted exactly as OCR::extract_feature does
// (src/OCR.cpp:203-218: index = position, 0-based, value != 0 only).
void ref_svm_predict_probability(void *ctx, const double *x, int n, int dims, int nthreads, double *label, double *prob)
{
	RefCtx *c = (RefCtx *)ctx;
	const int k = svm_get_nr_class(c->svm);
	if (nthreads < 1) nthreads = 1;
#pragma omp parallel for num_threads(nthreads) schedule(dynamic, 1)
	for (int i = 0; i < n; i++) {
		std::vector<svm_node> nodes;
		const double *xi = x + (size_t)i * dims;
		for (int d = 0; d < dims; d++) if (xi[d] != 0) { svm_node nd; nd.index = d; nd.value = xi[d]; nodes.push_back(nd); }
		svm_node end; end.index = -1; end.value = 0; nodes.push_back(end);
		label[i] = svm_predict_probability(c->svm, nodes.data(), prob + (size_t)i * k);
	}
}

// ---- whole-frame driver (the CPU baseline leg) --------------------------------------------------
// Mirrors the per-channel loop of ERFilter::text_detect (src/ER.cpp:50-60): for each of the six
// planes er_tree_extract -> non_maximum_supression -> classify.
// mode 0 ("R", reference-faithful): frames one at a time, omp parallel for over the 6 planes.
// mode 1 ("T", throughput-fair):   frames spread over nthreads, planes sequential inside a frame.
// mode 2 ("U", throughput-fair at a fixed step): the n_frames * 6 (frame, plane) units spread over nthreads, so that a
//         step of few frames still keeps every host thread busy.
// counts: per frame 4 ints = kept nodes, pool, strong, weak.  Returns wall seconds.
double ref_detect_frames(void *ctx, const uchar *bgr, int n_frames, int w, int h, int mode, int nthreads,
                         int *counts, double *stage_seconds /*3: extract,nms,classify summed over planes*/)
{
	RefCtx *c = (RefCtx *)ctx;
	if (nthreads < 1) nthreads = 1;
	double st_e = 0, st_n = 0, st_c = 0;
	auto t0 = std::chrono::high_resolution_clock::now();
	auto do_plane = [&](cv::Mat &plane, int *cnt, double *se, double *sn, double *sc) {
		auto a = std::chrono::high_resolution_clock::now();
		ER *root = c->erf->er_tree_extract(plane);
		auto b = std::chrono::high_resolution_clock::now();
		ERs all, pool, strong, weak;
		c->erf->non_maximum_supression(root, all, pool, plane);
		auto d = std::chrono::high_resolution_clock::now();
		c->erf->classify(pool, strong, weak, plane);
		auto e = std::chrono::high_resolution_clock::now();
		// count kept nodes
		int kept = 0;
		{
			std::vector<ER *> st; st.push_back(root);
			while (!st.empty()) { ER *x = st.back(); st.pop_back(); kept++; for (ER *k = x->child; k; k = k->next) st.push_back(k); }
		}
		cnt[0] = kept; cnt[1] = (int)pool.size(); cnt[2] = (int)strong.size(); cnt[3] = (int)weak.size();
		*se = std::chrono::duration<double>(b - a).count();
		*sn = std::chrono::duration<double>(d - b).count();
		*sc = std::chrono::duration<double>(e - d).count();
		c->erf->er_delete(root);
	};
	if (mode == 0) {
		for (int f = 0; f < n_frames; f++) {
			std::vector<cv::Mat> ch;
			channels_from_bgr(bgr + (size_t)f * w * h * 3, w, h, w * 3, ch);
			int cnt[6][4]; double se[6], sn[6], sc[6];
#pragma omp parallel for num_threads(nthreads)
			for (int i = 0; i < 6; i++) do_plane(ch[i], cnt[i], &se[i], &sn[i], &sc[i]);
			for (int k = 0; k < 4; k++) { counts[4 * f + k] = 0; for (int i = 0; i < 6; i++) counts[4 * f + k] += cnt[i][k]; }
			for (int i = 0; i < 6; i++) { st_e += se[i]; st_n += sn[i]; st_c += sc[i]; }
		}
	} else if (mode == 2) {
		std::vector<std::vector<cv::Mat> > chs((size_t)n_frames);
#pragma omp parallel for num_threads(nthreads) schedule(dynamic, 1)
		for (int f = 0; f < n_frames; f++) channels_from_bgr(bgr + (size_t)f * w * h * 3, w, h, w * 3, chs[(size_t)f]);
		std::vector<int> cnt((size_t)n_frames * 6 * 4, 0);
#pragma omp parallel for num_threads(nthreads) schedule(dynamic, 1) reduction(+ : st_e, st_n, st_c)
		for (int u = 0; u < n_frames * 6; u++) {
			double se, sn, sc;
			do_plane(chs[(size_t)(u / 6)][(size_t)(u % 6)], &cnt[(size_t)u * 4], &se, &sn, &sc);
			st_e += se; st_n += sn; st_c += sc;
		}
		for (int f = 0; f < n_frames; f++)
			for (int k = 0; k < 4; k++) { counts[4 * f + k] = 0; for (int i = 0; i < 6; i++) counts[4 * f + k] += cnt[((size_t)f * 6 + i) * 4 + k]; }
	} else {
#pragma omp parallel for num_threads(nthreads) schedule(dynamic, 1) reduction(+ : st_e, st_n, st_c)
		for (int f = 0; f < n_frames; f++) {
			std::vector<cv::Mat> ch;
			channels_from_bgr(bgr + (size_t)f * w * h * 3, w, h, w * 3, ch);
			int cnt[6][4]; double se, sn, sc;
			for (int i = 0; i < 6; i++) { do_plane(ch[i], cnt[i], &se, &sn, &sc); st_e += se; st_n += sn; st_c += sc; }
			for (int k = 0; k < 4; k++) { counts[4 * f + k] = 0; for (int i = 0; i < 6; i++) counts[4 * f + k] += cnt[i][k]; }
		}
	}
	auto t1 = std::chrono::high_resolution_clock::now();
	if (stage_seconds) { stage_seconds[0] = st_e; stage_seconds[1] = st_n; stage_seconds[2] = st_c; }
	return std::chrono::duration<double>(t1 - t0).count();
}

void ref_channels(const uchar *bgr, int w, int h, uchar *planes6)
{
	std::vector<cv::Mat> ch;
	channels_from_bgr(bgr, w, h, w * 3, ch);
	for (int k = 0; k < 6; k++) memcpy(planes6 + (size_t)k * w * h, ch[k].data, (size_t)w * h);
}

void ref_destroy(void *ctx)
{
	RefCtx *c = (RefCtx *)ctx;
	delete c->erf; delete c->stc; delete c->wtc; delete c->ocr;
	if (c->svm) svm_free_and_destroy_model(&c->svm);
	delete c;
}

} // extern "C"
