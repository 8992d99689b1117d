// ERFilter.hpp -- header-only C++ host facade over the C ABI (include/ertext.h).
//
// Mirrors the reference's call surface for the hot path so that its callers (image_mode
// src/utils.cpp:49, video_mode src/utils.cpp:115-140, the offline helpers) read the same:
//   ERFilter(thresh_step, min_area, max_area, stability_t, overlap_coef, min_ocr_prob)   inc/ER.h:113
//   AdaBoost *stc, *wtc                                                                  inc/ER.h:117-118
//   text_detect / er_tree_extract / non_maximum_supression / classify / er_delete        inc/ER.h:123-128
//   make_LBP_hist / set_thresh_step / set_min_area                                       inc/ER.h:132-136
//   CascadeBoost(filename), predict(vector<double>) with -DBL_MAX = rejected             inc/adaboost.h:159-166
//   struct ER with level/area/bound/parent/child/next/done/stability                     inc/ER.h:42-80
//   er_track(strong, weak, tracked, ...) incl. calc_color                                 inc/ER.h:129, src/ER.cpp:532-609
//   OCR(svm_file, img_L, feature_L), chain_run(src, thresh, slope)                        inc/OCR.h:31-35, src/OCR.cpp:67-140
//   struct Text                                                                          inc/ER.h:84-97
// Downstream CPU stages of the reference (er_track, er_grouping, er_ocr) consume the ER* trees
// this facade rebuilds from the device results.  With -DERT_WITH_OPENCV the cv::Mat / cv::Rect types
// are used directly; otherwise minimal stand-ins with the same member names are provided.
//
// Error behaviour follows the reference: a non-8UC1 input throws (CV_Assert, src/ER.cpp:242);
// loaders print and report false; anything the device reports is thrown as std::runtime_error.
#pragma once
#include "../../include/ertext.h"

#include <cfloat>
#include <cmath>
#include <chrono>
#include <cstdint>
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>

#ifdef ERT_WITH_OPENCV
#include <opencv2/core.hpp>
namespace ertx { using cv::Mat; using cv::Rect; using cv::Point; }
#else
namespace ertx {
struct Point { int x = 0, y = 0; };
struct Rect {
	int x = 0, y = 0, width = 0, height = 0;
	Rect() {}
	Rect(int x_, int y_, int w_, int h_) : x(x_), y(y_), width(w_), height(h_) {}
	int area() const { return width * height; }
};
// non-owning 8-bit image view with cv::Mat's member names (rows, cols, step, data, channels())
struct Mat {
	int rows = 0, cols = 0, chans = 1;
	size_t step = 0;
	unsigned char *data = nullptr;
	Mat() {}
	Mat(int r, int c, int channels_, void *d, size_t step_ = 0) : rows(r), cols(c), chans(channels_), step(step_ ? step_ : (size_t)c * channels_), data((unsigned char *)d) {}
	int channels() const { return chans; }
	bool empty() const { return !data || !rows || !cols; }
	Mat operator()(const Rect &r) const { Mat m(r.height, r.width, chans, data + (size_t)r.y * step + (size_t)r.x * chans, step); return m; }
};
} // namespace ertx
#endif

namespace ertx {

struct ER {
	ER() {}
	ER(int level_, int pixel_, int x_, int y_) : pixel(pixel_), level(level_), x(x_), y(y_) { bound = Rect(x_, y_, 1, 1); }
	int pixel = 0, level = 0, x = 0, y = 0;
	int area = 1;
	Rect bound;
	Point center;
	double color1 = 0, color2 = 0, color3 = 0, stkw = 0;
	bool done = false;
	double stability = 0;
	ER *parent = nullptr, *child = nullptr, *next = nullptr, *sibling_L = nullptr, *sibling_R = nullptr;
	int ch = 0;
	char letter = 0;
	double prob = 0;
	int node_index = -1;   // position in the plane's DFS node array (facade bookkeeping)
};
typedef std::vector<ER *> ERs;

// inc/ER.h:84-97
struct Text {
	Text() {}
	ERs ers;
	double slope = 0;
	Rect box;
	std::string word;
};

// er_grouping / overlap_suppression / inner_suppression / fitline_avgslope / er_ocr (src/ER.cpp:612-786, 893-964,
// 1362-1389) are NOT part of this facade: they are small-N, order-dependent host logic that the reference keeps (SURVEY 8f
// rank 3).  A program that needs them links the reference's own definitions -- either against the reference's unmodified
// headers with host/dropin/erfilter_dropin.cpp underneath (the tested route), or compiled
// against this header's ER / Text, whose members carry the reference's names.

inline void throw_last(const char *what) { throw std::runtime_error(std::string(what) + ": " + ert_last_error()); }

// One device context shared by the facade objects of a thread.
class Device {
public:
	explicit Device(const ert_params &p, int device = 0) { ctx_ = ert_create(&p, device); if (!ctx_) throw_last("ert_create"); }
	~Device() { ert_destroy(ctx_); }
	ert_ctx *ctx() const { return ctx_; }
private:
	ert_ctx *ctx_;
	Device(const Device &);
	Device &operator=(const Device &);
};

class AdaBoost {
public:
	virtual ~AdaBoost() {}
	virtual double predict(std::vector<double> fv) = 0;
	virtual bool load_classifier(std::string filename) = 0;
};

// CascadeBoost(filename): the file is parsed by the library (same text format as the reference).
class CascadeBoost : public AdaBoost {
public:
	CascadeBoost(Device &dev, int which, const std::string &filename) : dev_(dev), which_(which) { load_classifier(filename); }
	bool load_classifier(std::string filename)
	{
		n_ = ert_load_cascade(dev_.ctx(), which_, filename.c_str());
		if (n_ < 0) { std::cout << "Error: " << filename << " is not opened!!" << std::endl; return false; }
		return true;
	}
	int get_num_iter() const { return n_; }
	double predict(std::vector<double> fv)
	{
		double s = 0;
		if (ert_cascade_predict_batch(dev_.ctx(), which_, fv.data(), 1, (int)fv.size(), &s)) throw_last("cascade predict");
		return s;
	}
private:
	Device &dev_;
	int which_, n_ = -1;
};

// OCR(svm_file_name, img_L, feature_L) + chain_run: the SVM model is parsed by the library (libsvm text format).
class OCR {
public:
	OCR(Device &dev, const char *svm_file_name, int img_L = 30, int feature_L = 15) : dev_(dev)
	{
		if (img_L != 30 || feature_L != 15) throw std::runtime_error("OCR: only img_L=30, feature_L=15 (the trained model's contract, inc/utils.h:12-13)");
		loaded_ = ert_load_svm(dev_.ctx(), svm_file_name) == 0;      // svm_load_model returns NULL on failure; so does this, quietly
	}
	bool loaded() const { return loaded_; }
	// double chain_run(Mat src, int thresh, double slope = 0)  (src/OCR.cpp:67): returns table[label] + probability
	double chain_run(const Mat &src, int thresh, double slope = 0)
	{
		(void)thresh;   // THRESH_OTSU ignores it (src/OCR.cpp:72)
		if (src.empty() || src.channels() != 1) throw std::runtime_error("chain_run: 8UC1 image expected");
		ert_ocr_region reg; reg.frame = 0; reg.plane = 0; reg.x = 0; reg.y = 0; reg.w = src.cols; reg.h = src.rows; reg.slope = slope;
		const ert_ocr_result *r = nullptr;
		if (ert_ocr_chain_run_plane(dev_.ctx(), src.data, src.cols, src.rows, (int)src.step, &reg, 1, &r)) throw_last("chain_run");
		return r->value[0];
	}
	// the batched form er_ocr wants (src/ER.cpp:728-735 loops chain_run over the ERs of a text line):
	// letter / prob are written into the ERs exactly as er_ocr does (src/ER.cpp:733-734)
	void chain_run(const Mat &channel, ERs &ers, double slope)
	{
		if (ers.empty()) return;
		std::vector<ert_ocr_region> regs(ers.size());
		for (size_t i = 0; i < ers.size(); i++) {
			regs[i].frame = 0; regs[i].plane = 0; regs[i].x = ers[i]->bound.x; regs[i].y = ers[i]->bound.y;
			regs[i].w = ers[i]->bound.width; regs[i].h = ers[i]->bound.height; regs[i].slope = slope;
		}
		const ert_ocr_result *r = nullptr;
		if (ert_ocr_chain_run_plane(dev_.ctx(), channel.data, channel.cols, channel.rows, (int)channel.step, regs.data(), (int)regs.size(), &r))
			throw_last("chain_run");
		for (size_t i = 0; i < ers.size(); i++) {
			const double result = r->value[i];
			ers[i]->letter = (char)floor(result);
			ers[i]->prob = result - floor(result);
		}
	}
private:
	Device &dev_;
	bool loaded_ = false;
};

class ERFilter {
public:
	// Defaults: the values every hot-path caller of the reference passes (src/main.cpp:22, inc/utils.h:6-11).  The reference's
	// header default THRESH_STEP = 2 (inc/ER.h:113) is outside the device path's range: thresh_step must be 5..255 (levels are
	// bytes with 255 reserved for walls and 6 bits in the global keys; steps 1..4 give up to 256 levels) -- ert_create fails loudly.
	ERFilter(int thresh_step = 8, int min_area = 120, int max_area = 900000, int stability_t = 2, double overlap_coef = 0.7,
	         double min_ocr_prob = 0.15, int device = 0)
	    : dev_(make_params(thresh_step, min_area, max_area, stability_t, overlap_coef, min_ocr_prob), device), min_ocr_prob_(min_ocr_prob) {}
	~ERFilter() {}

	AdaBoost *stc = nullptr;   // strong text classifier (assign a CascadeBoost built on device())
	AdaBoost *wtc = nullptr;   // weak text classifier
	Device &device() { return dev_; }

	void set_thresh_step(int t) { if (ert_set_thresh_step(dev_.ctx(), t)) throw_last("set_thresh_step"); }
	void set_min_area(int m) { ert_set_min_area(dev_.ctx(), m); }

	// ERFilter::text_detect up to classify (src/ER.cpp:33-60).  src: 8UC3 BGR.  times[0..2] = extract, nms, classify
	// seconds (device time), times[6] = wall seconds, like the reference's return value.
	std::vector<double> text_detect(const Mat &src, ERs &root, std::vector<ERs> &all, std::vector<ERs> &pool, std::vector<ERs> &strong,
	                                std::vector<ERs> &weak)
	{
		if (src.empty() || src.channels() != 3) throw std::runtime_error("text_detect: 8UC3 BGR image expected");
		const auto t0 = std::chrono::high_resolution_clock::now();
		const ert_result *r = nullptr;
		if (ert_detect_classify(dev_.ctx(), src.data, 1, src.cols, src.rows, (int)src.step, ERT_STAGE_CLASSIFY, &r)) throw_last("text_detect");
		check_status(r);
		root.assign(6, nullptr); all.assign(6, ERs()); pool.assign(6, ERs()); strong.assign(6, ERs()); weak.assign(6, ERs());
		for (int p = 0; p < 6; p++) {
			std::vector<ER *> nodes;
			root[p] = build_tree(r, p, nodes);
			for (int k = r->pool_offset[p]; k < r->pool_offset[p + 1]; k++) {
				ER *e = nodes[r->pool_node[k]];
				pool[p].push_back(e);
				if (r->pool_label[k] == ERT_LABEL_STRONG) strong[p].push_back(e);
				else if (r->pool_label[k] == ERT_LABEL_WEAK) weak[p].push_back(e);
			}
		}
		std::vector<double> times(7, 0);
		times[0] = r->stage_ms[0] * 1e-3; times[1] = r->stage_ms[1] * 1e-3; times[2] = r->stage_ms[2] * 1e-3;
		times[6] = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
		return times;
	}

	// text_detect through er_track (src/ER.cpp:33-63): as above, plus `tracked`; times[3] = er_track seconds.
	std::vector<double> text_detect(const Mat &src, ERs &root, std::vector<ERs> &all, std::vector<ERs> &pool, std::vector<ERs> &strong,
	                                std::vector<ERs> &weak, ERs &tracked)
	{
		if (src.empty() || src.channels() != 3) throw std::runtime_error("text_detect: 8UC3 BGR image expected");
		const auto t0 = std::chrono::high_resolution_clock::now();
		const ert_result *r = nullptr;
		if (ert_detect_classify(dev_.ctx(), src.data, 1, src.cols, src.rows, (int)src.step, ERT_STAGE_TRACK, &r)) throw_last("text_detect");
		check_status(r);
		const ert_track_result *t = nullptr;
		if (ert_er_track(dev_.ctx(), &t)) throw_last("er_track");
		root.assign(6, nullptr); all.assign(6, ERs()); pool.assign(6, ERs()); strong.assign(6, ERs()); weak.assign(6, ERs());
		std::vector<std::vector<ER *> > nodes(6);
		for (int p = 0; p < 6; p++) {
			root[p] = build_tree(r, p, nodes[(size_t)p]);
			for (int k = r->pool_offset[p]; k < r->pool_offset[p + 1]; k++) {
				ER *e = nodes[(size_t)p][(size_t)r->pool_node[k]];
				pool[p].push_back(e);
				if (r->pool_label[k] == ERT_LABEL_STRONG) strong[p].push_back(e);
				else if (r->pool_label[k] == ERT_LABEL_WEAK) weak[p].push_back(e);
			}
		}
		std::vector<ER *> cand((size_t)(t->cand_offset[1] - t->cand_offset[0]));
		for (size_t i = 0; i < cand.size(); i++) {
			const ert_tracked &c = t->cand[i];
			ER *e = nodes[(size_t)c.plane][(size_t)c.node];
			e->color1 = c.color1; e->color2 = c.color2; e->color3 = c.color3;
			e->center.x = c.center_x; e->center.y = c.center_y; e->ch = c.plane;
			cand[i] = e;
		}
		tracked.clear();
		for (int k = t->track_offset[0]; k < t->track_offset[1]; k++) tracked.push_back(cand[(size_t)t->tracked[k]]);
		std::vector<double> times(7, 0);
		times[0] = r->stage_ms[0] * 1e-3; times[1] = r->stage_ms[1] * 1e-3; times[2] = r->stage_ms[2] * 1e-3; times[3] = t->track_ms * 1e-3;
		times[6] = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
		return times;
	}

	// er_track(vector<ERs> &strong, vector<ERs> &weak, ERs &all_er, vector<Mat> &channel, Mat Ycrcb)  (src/ER.cpp:532).
	// The channel planes and the YCrCb frame are derived on the device from the BGR frame, so the facade takes `src`
	// instead of (channel, Ycrcb).  Sets color1..3, center, ch on every strong / weak ER and appends to all_er.
	void er_track(std::vector<ERs> &strong, std::vector<ERs> &weak, ERs &all_er, const Mat &src)
	{
		if (src.empty() || src.channels() != 3) throw std::runtime_error("er_track: 8UC3 BGR image expected");
		std::vector<int32_t> rs, rw;
		std::vector<ER *> es, ew;
		for (int pass = 0; pass < 2; pass++) {
			std::vector<ERs> &v = pass ? weak : strong;
			std::vector<int32_t> &rows = pass ? rw : rs;
			std::vector<ER *> &flat = pass ? ew : es;
			for (size_t ch = 0; ch < v.size(); ch++)
				for (ER *e : v[ch]) {
					const int32_t row[6] = {(int32_t)ch, e->bound.x, e->bound.y, e->bound.width, e->bound.height, e->area};
					rows.insert(rows.end(), row, row + 6);
					flat.push_back(e);
				}
		}
		const ert_track_result *t = nullptr;
		if (ert_er_track_regions(dev_.ctx(), src.data, src.cols, src.rows, (int)src.step, rs.data(), (int)es.size(), rw.data(), (int)ew.size(), &t))
			throw_last("er_track");
		std::vector<ER *> cand(es);
		cand.insert(cand.end(), ew.begin(), ew.end());
		for (size_t i = 0; i < cand.size(); i++) {
			const ert_tracked &c = t->cand[i];
			cand[i]->color1 = c.color1; cand[i]->color2 = c.color2; cand[i]->color3 = c.color3;
			cand[i]->center.x = c.center_x; cand[i]->center.y = c.center_y; cand[i]->ch = c.plane;
		}
		for (int k = t->track_offset[0]; k < t->track_offset[1]; k++) all_er.push_back(cand[(size_t)t->tracked[k]]);
	}

	// compute_channels(src, YCrcb, channels)  (src/ER.cpp:114-128): the six planes as one contiguous buffer
	// (plane k at channels6.data() + k * rows * cols); wrap them in Mat headers as needed.
	void compute_channels(const Mat &src, std::vector<unsigned char> &channels6)
	{
		if (src.empty() || src.channels() != 3) throw std::runtime_error("compute_channels: 8UC3 BGR image expected");
		channels6.resize((size_t)src.rows * src.cols * 6);
		if (ert_compute_channels(dev_.ctx(), src.data, src.cols, src.rows, (int)src.step, channels6.data())) throw_last("compute_channels");
	}

	// ER* er_tree_extract(Mat input)  (src/ER.cpp:240): 8UC1 plane -> heap tree owned by the caller (er_delete).
	ER *er_tree_extract(const Mat &input)
	{
		if (input.empty() || input.channels() != 1) throw std::runtime_error("CV_Assert failed: input.type() == CV_8UC1");
		const ert_result *r = nullptr;
		if (ert_planes_detect(dev_.ctx(), input.data, 1, input.cols, input.rows, (int)input.step, 0, ERT_STAGE_EXTRACT, &r)) throw_last("er_tree_extract");
		check_status(r);
		std::vector<ER *> nodes;
		return build_tree(r, 0, nodes);
	}

	// non_maximum_supression(ER *er, ERs &all, ERs &pool, Mat input)  (src/ER.cpp:416): runs on the caller's tree
	// with the caller's child order; marks done like the reference does not matter downstream, pool is filled.
	void non_maximum_supression(ER *er, ERs &all, ERs &pool, const Mat &input)
	{
		(void)all;
		std::vector<ER *> flat; std::vector<ert_node> nodes;
		flatten(er, flat, nodes);
		std::vector<int32_t> idx(flat.size() + 1);
		int np = 0;
		if (ert_nms_nodes(dev_.ctx(), nodes.data(), (int)nodes.size(), input.cols, input.rows, idx.data(), (int)idx.size(), &np)) throw_last("nms");
		er->parent = er;   // src/ER.cpp:424
		for (int i = 0; i < np; i++) pool.push_back(flat[idx[i]]);
	}

	// classify(ERs &pool, ERs &strong, ERs &weak, Mat input)  (src/ER.cpp:507)
	void classify(ERs &pool, ERs &strong, ERs &weak, const Mat &input)
	{
		if (pool.empty()) return;
		std::vector<int32_t> rects(4 * pool.size()), label(pool.size());
		for (size_t i = 0; i < pool.size(); i++) { rects[4 * i] = pool[i]->bound.x; rects[4 * i + 1] = pool[i]->bound.y; rects[4 * i + 2] = pool[i]->bound.width; rects[4 * i + 3] = pool[i]->bound.height; }
		if (ert_classify_regions(dev_.ctx(), input.data, input.cols, input.rows, (int)input.step, rects.data(), (int)pool.size(), label.data(), nullptr, nullptr, nullptr))
			throw_last("classify");
		for (size_t i = 0; i < pool.size(); i++) {
			if (label[i] == ERT_LABEL_STRONG) strong.push_back(pool[i]);
			else if (label[i] == ERT_LABEL_WEAK) weak.push_back(pool[i]);
		}
	}

	// vector<double> make_LBP_hist(Mat input, N = 2, normalize_size = 24)  (src/ER.cpp:789)
	std::vector<double> make_LBP_hist(const Mat &input, const int N = 2, const int normalize_size = 24)
	{
		if (N != 2 || normalize_size != 24) throw std::runtime_error("make_LBP_hist: only N=2, normalize_size=24 (the trained classifiers' contract)");
		std::vector<double> h(1024);
		const int32_t rect[4] = {0, 0, input.cols, input.rows};
		if (ert_lbp_hist(dev_.ctx(), input.data, input.cols, input.rows, (int)input.step, rect, 1, h.data())) throw_last("make_LBP_hist");
		return h;
	}

	// er_delete(ER *er)  (src/ER.cpp:194-233)
	void er_delete(ER *er)
	{
		std::vector<ER *> st;
		if (er) st.push_back(er);
		while (!st.empty()) {
			ER *e = st.back(); st.pop_back();
			for (ER *c = e->child; c; c = c->next) st.push_back(c);
			delete e;
		}
	}

private:
	Device dev_;
	double min_ocr_prob_;

	static ert_params make_params(int ts, int mina, int maxa, int st, double oc, double mp)
	{
		ert_params p; p.thresh_step = ts; p.min_area = mina; p.max_area = maxa; p.stability_t = st; p.overlap_coef = oc; p.min_ocr_prob = mp;
		return p;
	}
	static void check_status(const ert_result *r)
	{
		if (r->status) throw std::runtime_error(std::string("device status: ") + ert_status_string(r->status));
	}
	// rebuild the linked ER tree (parent/child/next in visiting order) from the plane's DFS node array
	static ER *build_tree(const ert_result *r, int plane, std::vector<ER *> &nodes)
	{
		const int a = r->node_offset[plane], b = r->node_offset[plane + 1];
		nodes.assign((size_t)(b - a), nullptr);
		std::vector<ER *> last_child((size_t)(b - a), nullptr);
		for (int i = a; i < b; i++) {
			const ert_node &n = r->nodes[i];
			ER *e = new ER(n.level, n.y * r->width + n.x, n.x, n.y);
			e->area = n.area; e->bound = Rect(n.x, n.y, n.w, n.h); e->node_index = i - a;
			nodes[(size_t)(i - a)] = e;
			if (n.parent >= 0) {
				ER *p = nodes[(size_t)n.parent];
				e->parent = p;
				if (!p->child) p->child = e; else last_child[(size_t)n.parent]->next = e;
				last_child[(size_t)n.parent] = e;
			}
		}
		return nodes.empty() ? nullptr : nodes[0];
	}
	static void flatten(ER *root, std::vector<ER *> &flat, std::vector<ert_node> &nodes)
	{
		std::vector<std::pair<ER *, int> > st;
		st.push_back(std::make_pair(root, -1));
		while (!st.empty()) {
			std::pair<ER *, int> cur = st.back(); st.pop_back();
			const int idx = (int)flat.size();
			flat.push_back(cur.first);
			ert_node n; n.level = cur.first->level; n.area = cur.first->area; n.x = cur.first->bound.x; n.y = cur.first->bound.y;
			n.w = cur.first->bound.width; n.h = cur.first->bound.height; n.parent = cur.second; n.n_children = 0;
			std::vector<ER *> ch;
			for (ER *c = cur.first->child; c; c = c->next) ch.push_back(c);
			n.n_children = (int)ch.size();
			nodes.push_back(n);
			for (int i = (int)ch.size() - 1; i >= 0; i--) st.push_back(std::make_pair(ch[(size_t)i], idx));
		}
	}
};

} // namespace ertx
