// erfilter_dropin.cpp -- the hot path of the reference's OWN classes, re-defined on top of libertext.so.
//
// This translation unit is compiled against the reference's UNMODIFIED headers (inc/ER.h, inc/adaboost.h; add the
// reference's inc/ to the include path) and defines exactly the member functions of the path -- with the reference's
// own signatures, so every caller (ERFilter::text_detect src/ER.cpp:33-111, video_mode src/utils.cpp:115-140,
// image_mode, the offline helpers) compiles and links unchanged:
//
//   ERFilter::compute_channels        src/ER.cpp:114-128      -> ert_compute_channels
//   ERFilter::er_tree_extract         src/ER.cpp:240-374      -> ert_planes_detect(ERT_STAGE_EXTRACT)  (+ process_stack,
//                                                               er_accumulate, er_merge: private helpers, no longer used)
//   ERFilter::non_maximum_supression  src/ER.cpp:416-505      -> ert_nms_nodes (the caller's tree, the caller's child order)
//   ERFilter::classify                src/ER.cpp:507-528      -> ert_classify_regions
//   ERFilter::er_track                src/ER.cpp:532-609      -> ert_er_track_regions_ycc (calc_color included)
//   ERFilter::make_LBP_hist           src/ER.cpp:789-816      -> ert_lbp_hist
//   ERFilter::calc_LBP                src/ER.cpp:819-845      -> ert_calc_lbp
//   CascadeBoost::predict             src/adaboost.cpp:507-542 -> ert_cascade_predict_batch
//
// The reference keeps everything else (text_detect itself, er_delete, er_grouping, er_ocr, the word graph, training).
// The maintainer's side of the binding is one guard around those eight definitions; libsvm is
// replaced at link time by libertext_svm.so (svm_shim.cpp), no source change.
//
// State: the reference's class layouts cannot grow, so device contexts live in a registry keyed by the object's
// address.  The stage functions are called concurrently from the 6 OpenMP threads of text_detect (src/ER.cpp:50-60):
// every call leases a context from the filter's pool (one per concurrent caller, created on demand).
// Errors: CV_Assert-style std::runtime_error for wrong Mat types (the reference throws cv::Exception there,
// src/ER.cpp:242); anything the device reports is thrown as std::runtime_error -- there is no CPU path behind this.
#include "ER.h"
#include "../../../include/ertext.h"

#include <cfloat>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <stdexcept>

namespace ertdrop {

inline void fail(const char *what) { throw std::runtime_error(std::string(what) + ": " + ert_last_error()); }

inline int device_ordinal()
{
	const char *e = getenv("ERTEXT_DEVICE");
	return e ? atoi(e) : 0;
}

// ---- cascades handed over by the reference's CascadeBoost objects ------------------------------------------------
struct CascadeTable {
	std::vector<int> stage_len, stage_thr, dim;
	std::vector<double> thr, cp, cn;
	unsigned long long version = 0;
};
struct Registry {
	std::mutex mu;
	std::map<const AdaBoost *, CascadeTable> cascades;
	unsigned long long next_version = 1;
};
inline Registry &registry() { static Registry r; return r; }

// ---- context pools ------------------------------------------------------------------------------------------
struct Ctx {
	ert_ctx *c = nullptr;
	unsigned long long casc_version[2] = {0, 0};
};
struct Pool {
	std::mutex mu;
	std::vector<Ctx *> idle;
};
struct Pools {
	std::mutex mu;
	std::map<const void *, Pool *> by_owner;
};
inline Pools &pools() { static Pools p; return p; }

inline Pool *pool_of(const void *owner)
{
	Pools &P = pools();
	std::lock_guard<std::mutex> lk(P.mu);
	Pool *&p = P.by_owner[owner];
	if (!p) p = new Pool();
	return p;
}

class Lease {
public:
	Lease(const void *owner, const ert_params &prm) : pool_(pool_of(owner))
	{
		{
			std::lock_guard<std::mutex> lk(pool_->mu);
			if (!pool_->idle.empty()) { ctx_ = pool_->idle.back(); pool_->idle.pop_back(); }
		}
		if (!ctx_) {
			ctx_ = new Ctx();
			ctx_->c = ert_create(&prm, device_ordinal());
			if (!ctx_->c) { delete ctx_; ctx_ = nullptr; fail("ert_create"); }
		} else if (ert_set_params(ctx_->c, &prm)) fail("ert_set_params");
	}
	~Lease() { if (ctx_) { std::lock_guard<std::mutex> lk(pool_->mu); pool_->idle.push_back(ctx_); } }
	ert_ctx *c() const { return ctx_->c; }
	// make sure slot `which` of the leased context holds the table of classifier `a`
	void use_cascade(int which, AdaBoost *a)
	{
		if (!a) throw std::runtime_error("classify: ERFilter::stc / wtc is not set");
		Registry &R = registry();
		CascadeTable t;
		{
			std::lock_guard<std::mutex> lk(R.mu);
			std::map<const AdaBoost *, CascadeTable>::const_iterator it = R.cascades.find(a);
			if (it != R.cascades.end()) t = it->second;
		}
		if (!t.version) {
			// not seen yet: CascadeBoost::predict (below) publishes the object's table on its first call
			a->predict(std::vector<double>(1024, 0.0));
			std::lock_guard<std::mutex> lk(R.mu);
			std::map<const AdaBoost *, CascadeTable>::const_iterator it = R.cascades.find(a);
			if (it == R.cascades.end()) throw std::runtime_error("classify: stc / wtc must be CascadeBoost objects (REAL boosting of decision stumps) on the device path");
			t = it->second;
		}
		if (ctx_->casc_version[which] == t.version) return;
		if (ert_set_cascade(ctx_->c, which, (int)t.stage_len.size(), t.stage_len.data(), t.stage_thr.data(), (int)t.dim.size(), t.dim.data(), t.thr.data(),
		                    t.cp.data(), t.cn.data()) < 0) fail("ert_set_cascade");
		ctx_->casc_version[which] = t.version;
	}
private:
	Pool *pool_;
	Ctx *ctx_ = nullptr;
};

inline void require_8uc1(const Mat &m, const char *who)
{
	if (m.empty() || m.type() != CV_8UC1) throw std::runtime_error(std::string(who) + ": CV_Assert failed: input.type() == CV_8UC1");
}

// rebuild the linked ER tree (parent / child / next, children in visiting order) from a plane's DFS node array
inline ER *build_tree(const ert_result *r, int plane)
{
	const int a = r->node_offset[plane], b = r->node_offset[plane + 1];
	std::vector<ER *> nodes((size_t)(b - a), nullptr), last_child((size_t)(b - a), nullptr);
	for (int i = a; i < b; i++) {
		const ert_node &n = r->nodes[i];
		ER *e = new ER(n.level, n.y * r->width + n.x, n.x, n.y);
		e->area = n.area;
		e->bound = Rect(n.x, n.y, n.w, n.h);
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

inline void flatten(ER *root, std::vector<ER *> &flat, std::vector<ert_node> &nodes)
{
	std::vector<std::pair<ER *, int> > st;
	st.push_back(std::make_pair(root, -1));
	while (!st.empty()) {
		const std::pair<ER *, int> cur = st.back();
		st.pop_back();
		const int idx = (int)flat.size();
		flat.push_back(cur.first);
		ert_node n;
		n.level = cur.first->level; n.area = cur.first->area; n.x = cur.first->bound.x; n.y = cur.first->bound.y;
		n.w = cur.first->bound.width; n.h = cur.first->bound.height; n.parent = cur.second;
		std::vector<ER *> ch;
		for (ER *c = cur.first->child; c; c = c->next) ch.push_back(c);
		n.n_children = (int)ch.size();
		nodes.push_back(n);
		for (int i = (int)ch.size() - 1; i >= 0; i--) st.push_back(std::make_pair(ch[(size_t)i], idx));
	}
}

} // namespace ertdrop

// private parameters of the filter as the C ABI wants them (members are accessible: these ARE member functions)
#define ERTDROP_PARAMS(name) \
	ert_params name; name.thresh_step = THRESH_STEP; name.min_area = MIN_AREA; name.max_area = MAX_AREA; name.stability_t = STABILITY_T; \
	name.overlap_coef = OVERLAP_COEF; name.min_ocr_prob = MIN_OCR_PROB

// ---------------------------------------------------------------------------------------------
// ERFilter::compute_channels(Mat &src, Mat &YCrcb, vector<Mat> &channels)   inc/ER.h:124, src/ER.cpp:114-128
// ---------------------------------------------------------------------------------------------
void ERFilter::compute_channels(Mat &src, Mat &YCrcb, vector<Mat> &channels)
{
	if (src.empty() || src.type() != CV_8UC3) throw std::runtime_error("compute_channels: 8UC3 BGR image expected");
	ERTDROP_PARAMS(prm);
	ertdrop::Lease L(this, prm);
	const int W = src.cols, H = src.rows;
	std::vector<unsigned char> planes((size_t)W * H * 6);
	if (ert_compute_channels(L.c(), src.data, W, H, (int)src.step, planes.data())) ertdrop::fail("compute_channels");
	channels.clear();
	for (int k = 0; k < 6; k++) {
		Mat m(H, W, CV_8UC1);
		for (int y = 0; y < H; y++) memcpy(m.ptr(y), planes.data() + ((size_t)k * H + y) * W, (size_t)W);
		channels.push_back(m);
	}
	YCrcb = Mat(H, W, CV_8UC3);
	for (int y = 0; y < H; y++) {
		unsigned char *o = YCrcb.ptr(y);
		const unsigned char *p0 = planes.data() + (size_t)y * W, *p1 = p0 + (size_t)W * H, *p2 = p1 + (size_t)W * H;
		for (int x = 0; x < W; x++) { o[3 * x] = p0[x]; o[3 * x + 1] = p1[x]; o[3 * x + 2] = p2[x]; }
	}
}

// ---------------------------------------------------------------------------------------------
// ER* ERFilter::er_tree_extract(Mat input)   inc/ER.h:125, src/ER.cpp:240-374.  Heap tree owned by the caller (er_delete).
// ---------------------------------------------------------------------------------------------
ER *ERFilter::er_tree_extract(Mat input)
{
	ertdrop::require_8uc1(input, "er_tree_extract");
	ERTDROP_PARAMS(prm);
	ertdrop::Lease L(this, prm);
	const ert_result *r = nullptr;
	if (ert_planes_detect(L.c(), input.data, 1, input.cols, input.rows, (int)input.step, 0, ERT_STAGE_EXTRACT, &r)) ertdrop::fail("er_tree_extract");
	if (r->status) throw std::runtime_error(std::string("er_tree_extract: device status: ") + ert_status_string(r->status));
	return ertdrop::build_tree(r, 0);
}

// (the reference's private flood helpers er_accumulate / er_merge / process_stack, inc/ER.h:149-151, are not needed any more;
// they may stay in src/ER.cpp or go with the guard -- nothing calls them)

// ---------------------------------------------------------------------------------------------
// void ERFilter::non_maximum_supression(ER *er, ERs &all, ERs &pool, Mat input)   inc/ER.h:126, src/ER.cpp:416-505
// ---------------------------------------------------------------------------------------------
void ERFilter::non_maximum_supression(ER *er, ERs &all, ERs &pool, Mat input)
{
	(void)all;   // GET_ALL_ER is off in the reference (inc/ER.h:24): `all` stays empty
	if (!er) return;
	ERTDROP_PARAMS(prm);
	ertdrop::Lease L(this, prm);
	std::vector<ER *> flat;
	std::vector<ert_node> nodes;
	ertdrop::flatten(er, flat, nodes);
	std::vector<int32_t> idx(flat.size() + 1);
	int np = 0;
	if (ert_nms_nodes(L.c(), nodes.data(), (int)nodes.size(), input.cols, input.rows, idx.data(), (int)idx.size(), &np)) ertdrop::fail("non_maximum_supression");
	er->parent = er;   // src/ER.cpp:424
	for (int i = 0; i < np; i++) pool.push_back(flat[(size_t)idx[i]]);
}

// ---------------------------------------------------------------------------------------------
// void ERFilter::classify(ERs &pool, ERs &strong, ERs &weak, Mat input)   inc/ER.h:127, src/ER.cpp:507-528
// ---------------------------------------------------------------------------------------------
void ERFilter::classify(ERs &pool, ERs &strong, ERs &weak, Mat input)
{
	if (pool.empty()) return;
	ertdrop::require_8uc1(input, "classify");
	ERTDROP_PARAMS(prm);
	ertdrop::Lease L(this, prm);
	L.use_cascade(ERT_CASCADE_STRONG, stc);
	L.use_cascade(ERT_CASCADE_WEAK, wtc);
	std::vector<int32_t> rects(4 * pool.size()), label(pool.size());
	for (size_t i = 0; i < pool.size(); i++) {
		rects[4 * i] = pool[i]->bound.x; rects[4 * i + 1] = pool[i]->bound.y; rects[4 * i + 2] = pool[i]->bound.width; rects[4 * i + 3] = pool[i]->bound.height;
	}
	if (ert_classify_regions(L.c(), input.data, input.cols, input.rows, (int)input.step, rects.data(), (int)pool.size(), label.data(), nullptr, nullptr, nullptr))
		ertdrop::fail("classify");
	for (size_t i = 0; i < pool.size(); i++) {
		if (label[i] == ERT_LABEL_STRONG) strong.push_back(pool[i]);
		else if (label[i] == ERT_LABEL_WEAK) weak.push_back(pool[i]);
	}
}

// ---------------------------------------------------------------------------------------------
// void ERFilter::er_track(vector<ERs> &strong, vector<ERs> &weak, ERs &all_er, vector<Mat> &channel, Mat Ycrcb)
// inc/ER.h:129, src/ER.cpp:532-609 (calc_color src/ER.cpp:1391-1437 runs on the device with it)
// ---------------------------------------------------------------------------------------------
void ERFilter::er_track(vector<ERs> &strong, vector<ERs> &weak, ERs &all_er, vector<Mat> &channel, Mat Ycrcb)
{
	(void)Ycrcb;   // its three planes are channel[0..2] (src/ER.cpp:122-124)
	if (channel.size() < 3) throw std::runtime_error("er_track: the six channels of compute_channels expected");
	for (int k = 0; k < 3; k++) {
		ertdrop::require_8uc1(channel[(size_t)k], "er_track");
		if (channel[(size_t)k].cols != channel[0].cols || channel[(size_t)k].rows != channel[0].rows || channel[(size_t)k].step != channel[0].step)
			throw std::runtime_error("er_track: channel planes of different geometry");
	}
	ERTDROP_PARAMS(prm);
	ertdrop::Lease L(this, prm);
	std::vector<int32_t> rs, rw;
	std::vector<ER *> es, ew;
	for (int pass = 0; pass < 2; pass++) {
		vector<ERs> &v = pass ? weak : strong;
		std::vector<int32_t> &rows = pass ? rw : rs;
		std::vector<ER *> &flat = pass ? ew : es;
		for (size_t ch = 0; ch < v.size(); ch++)
			for (size_t i = 0; i < v[ch].size(); i++) {
				ER *e = v[ch][i];
				const int32_t row[6] = {(int32_t)ch, e->bound.x, e->bound.y, e->bound.width, e->bound.height, e->area};
				rows.insert(rows.end(), row, row + 6);
				flat.push_back(e);
			}
	}
	const ert_track_result *t = nullptr;
	if (ert_er_track_regions_ycc(L.c(), channel[0].data, channel[1].data, channel[2].data, channel[0].cols, channel[0].rows, (int)channel[0].step, rs.data(),
	                             (int)es.size(), rw.data(), (int)ew.size(), &t)) ertdrop::fail("er_track");
	std::vector<ER *> cand(es);
	cand.insert(cand.end(), ew.begin(), ew.end());
	for (size_t i = 0; i < cand.size(); i++) {
		const ert_tracked &c = t->cand[i];
		cand[i]->color1 = c.color1; cand[i]->color2 = c.color2; cand[i]->color3 = c.color3;        // calc_color
		cand[i]->center = Point(c.center_x, c.center_y);                                           // src/ER.cpp:545
		cand[i]->ch = c.plane;                                                                     // src/ER.cpp:546
	}
	for (int k = t->track_offset[0]; k < t->track_offset[1]; k++) all_er.push_back(cand[(size_t)t->tracked[k]]);
}

// ---------------------------------------------------------------------------------------------
// vector<double> ERFilter::make_LBP_hist(Mat input, const int N, const int normalize_size)   inc/ER.h:132, src/ER.cpp:789-816
// Mat ERFilter::calc_LBP(Mat input, const int size)                                          inc/ER.h:134, src/ER.cpp:819-845
// The device kernels implement the trained classifiers' contract (N = 2, 24 x 24; inc/ER.h defaults, the only values any
// caller of the reference passes: src/ER.cpp:517, src/utils.cpp:1451-1469).
// ---------------------------------------------------------------------------------------------
vector<double> ERFilter::make_LBP_hist(Mat input, const int N, const int normalize_size)
{
	ertdrop::require_8uc1(input, "make_LBP_hist");
	if (N != 2 || normalize_size != 24) throw std::runtime_error("make_LBP_hist: the device path implements N = 2, normalize_size = 24 (the trained classifiers' feature)");
	ERTDROP_PARAMS(prm);
	ertdrop::Lease L(this, prm);
	vector<double> h(1024);
	const int32_t rect[4] = {0, 0, input.cols, input.rows};
	if (ert_lbp_hist(L.c(), input.data, input.cols, input.rows, (int)input.step, rect, 1, h.data())) ertdrop::fail("make_LBP_hist");
	return h;
}

Mat ERFilter::calc_LBP(Mat input, const int size)
{
	ertdrop::require_8uc1(input, "calc_LBP");
	if (size != 24) throw std::runtime_error("calc_LBP: the device path implements size = 24");
	ERTDROP_PARAMS(prm);
	ertdrop::Lease L(this, prm);
	Mat LBP(size, size, CV_8UC1);
	std::vector<unsigned char> codes(576);
	const int32_t rect[4] = {0, 0, input.cols, input.rows};
	if (ert_calc_lbp(L.c(), input.data, input.cols, input.rows, (int)input.step, rect, 1, codes.data())) ertdrop::fail("calc_LBP");
	for (int y = 0; y < size; y++) memcpy(LBP.ptr(y), codes.data() + (size_t)y * size, (size_t)size);
	return LBP;
}

// ---------------------------------------------------------------------------------------------
// double CascadeBoost::predict(vector<double> fv)   inc/adaboost.h:163, src/adaboost.cpp:507-542
// The object keeps the reference's parsed state (classifier[], num_of_iter[], thresh[], filled by the reference's own
// load_classifier or training); its table is published once and scored on the device.  -DBL_MAX = rejected.
// ---------------------------------------------------------------------------------------------
double CascadeBoost::predict(vector<double> fv)
{
	using namespace ertdrop;
	Registry &R = registry();
	CascadeTable t;
	{
		std::lock_guard<std::mutex> lk(R.mu);
		std::map<const AdaBoost *, CascadeTable>::iterator it = R.cascades.find(this);
		if (it == R.cascades.end() || it->second.dim.size() != classifier.size()) {
			if (boost_type != REAL) throw std::runtime_error("CascadeBoost::predict: the device path implements REAL boosting of decision stumps");
			CascadeTable n;
			n.stage_len = num_of_iter;
			n.stage_thr = thresh;
			for (size_t j = 0; j < classifier.size(); j++) {
				const vector<double> para = classifier[j]->get_para();     // RealDecisionStump: dim, thresh, cp, cn (src/adaboost.cpp:138-146)
				if (para.size() != 4) throw std::runtime_error("CascadeBoost::predict: RealDecisionStump base classifiers expected");
				n.dim.push_back((int)para[0]); n.thr.push_back(para[1]); n.cp.push_back(para[2]); n.cn.push_back(para[3]);
			}
			n.version = R.next_version++;
			R.cascades[this] = n;
			it = R.cascades.find(this);
		}
		t = it->second;
	}
	if (fv.size() < 1024) throw std::runtime_error("CascadeBoost::predict: 1024-bin feature vector expected");
	ert_params prm; prm.thresh_step = 8; prm.min_area = 120; prm.max_area = 900000; prm.stability_t = 2; prm.overlap_coef = 0.7; prm.min_ocr_prob = 0.15;
	Lease L(this, prm);
	L.use_cascade(ERT_CASCADE_STRONG, this);
	double s = 0;
	if (ert_cascade_predict_batch(L.c(), ERT_CASCADE_STRONG, fv.data(), 1, (int)fv.size(), &s)) fail("CascadeBoost::predict");
	return s;
}
