// oracle/ref_capi_next.cpp -- TEST INFRASTRUCTURE (the checker), never linked into the product.
//
// C wrapper around the reference's OWN code for the rows that FOLLOW the detect path:
//   ref_calc_color / ref_er_track   calc_color, ERFilter::er_track   (src/ER.cpp:1391-1437, 532-609)
//   ref_chain_run / ref_ocr_features OCR::chain_run, extract_feature, rotate_mat, ARAN
//                                                                   (src/OCR.cpp:67-140, 144-250, 254-360, 394-430)
// The function bodies are extracted by line range by oracle/build_ref.sh and compiled unmodified against
// oracle/cvshim (whose threshold / findContours / GaussianBlur / normalize are pinned against cv2 through the
// ref_prim_* entry points below).
#include <stdio.h>
#include <string.h>
#include <math.h>
#include <time.h>
#include <string>
#include <algorithm>
#include <iostream>
#include <iterator>
#include <fstream>
#include <sstream>
#include <vector>
#include <stack>
#include <map>
#include <set>
#include <thread>
#include <chrono>
#include <numeric>
#include <forward_list>
#include <memory>
#include <omp.h>
#include <opencv2/opencv.hpp>   // oracle/cvshim
#define private public          // OCR keeps img_L / feature_L / model private; the checker sets them directly
#include "ER.h"
#undef private

namespace {

struct RefCtxView {             // same leading layout as RefCtx in ref_capi.cpp
	ERFilter *erf;
	CascadeBoost *stc;
	CascadeBoost *wtc;
	OCR *ocr;
	svm_model *svm;
};

OCR *ocr_of(void *ctx)
{
	RefCtxView *c = (RefCtxView *)ctx;
	c->ocr->img_L = 30;         // new OCR("classifier/OCR.model", OCR_IMG_L, OCR_FEATURE_L)  (src/main.cpp:25, inc/utils.h:12-13)
	c->ocr->feature_L = 15;
	c->ocr->model = c->svm;
	return c->ocr;
}

} // namespace

extern "C" {

// ---- shim primitives, exposed so that the tests can pin them against python cv2 ---------------------
int ref_prim_threshold_otsu(const uchar *src, int w, int h, int stride, uchar *out)
{
	cv::Mat view(h, w, CV_8UC1, (void *)src, (size_t)stride), dst;
	double t = cv::threshold(view, dst, 128, 255, cv::THRESH_OTSU);
	for (int i = 0; i < h; i++) memcpy(out + (size_t)i * w, dst.ptr(i), (size_t)w);
	return (int)t;
}

// contours flattened: pts = (x, y) pairs, sizes[k] = points in contour k; returns the number of contours
int ref_prim_find_contours(const uchar *src, int w, int h, int stride, int *pts, int pts_cap, int *sizes, int sizes_cap)
{
	cv::Mat view(h, w, CV_8UC1, (void *)src, (size_t)stride);
	std::vector<std::vector<cv::Point> > cs;
	cv::findContours(view, cs, cv::RETR_LIST, cv::CHAIN_APPROX_NONE);
	int np = 0;
	for (size_t k = 0; k < cs.size(); k++) {
		if ((int)k < sizes_cap) sizes[k] = (int)cs[k].size();
		for (size_t j = 0; j < cs[k].size(); j++, np++)
			if (np < pts_cap) { pts[2 * np] = cs[k][j].x; pts[2 * np + 1] = cs[k][j].y; }
	}
	return (int)cs.size();
}

void ref_prim_gaussian7(const uchar *src, int w, int h, int stride, uchar *out)
{
	cv::Mat view(h, w, CV_8UC1, (void *)src, (size_t)stride), dst;
	cv::GaussianBlur(view, dst, cv::Size(7, 7), 0);
	for (int i = 0; i < h; i++) memcpy(out + (size_t)i * w, dst.ptr(i), (size_t)w);
}

void ref_prim_normalize_minmax(const uchar *src, int w, int h, int stride, uchar *out)
{
	cv::Mat view(h, w, CV_8UC1, (void *)src, (size_t)stride), dst;
	cv::normalize(view, dst, 0, 255, cv::NORM_MINMAX, CV_8U);
	for (int i = 0; i < h; i++) memcpy(out + (size_t)i * w, dst.ptr(i), (size_t)w);
}

// ---- calc_color (src/ER.cpp:1391-1437) ---------------------------------------------------------------
// plane = the ER's channel image (w x h); ycrcb = the 3-channel interleaved YCrCb frame; rects = n x (x,y,w,h).
void ref_calc_color(const uchar *plane, const uchar *ycrcb, int w, int h, const int *rects, int n, double *color3)
{
	cv::Mat ch(h, w, CV_8UC1, (void *)plane, (size_t)w);
	cv::Mat ycc(h, 3 * w, CV_8UC1, (void *)ycrcb, (size_t)3 * w);    // 8UC3 seen as rows of 3*w bytes: calc_color only uses ptr(i)
	for (int i = 0; i < n; i++) {
		ER e(0, 0, 0, 0);
		e.bound = cv::Rect(rects[4 * i], rects[4 * i + 1], rects[4 * i + 2], rects[4 * i + 3]);
		calc_color(&e, ch, ycc);
		color3[3 * i] = e.color1; color3[3 * i + 1] = e.color2; color3[3 * i + 2] = e.color3;
	}
}

// ---- ERFilter::er_track (src/ER.cpp:532-609) ---------------------------------------------------------
// planes6 = the six channel images; strong / weak = rows of (ch, x, y, w, h, area), channel-major (the order
// classify fills strong[ch] / weak[ch]).  tracked_out receives, in all_er order, (kind, index): kind 0 = strong
// row, 1 = weak row.  *_color = rows x 3 doubles, *_center = rows x 2.  Returns the length of all_er.
int ref_er_track(void *ctx, const uchar *planes6, const uchar *ycrcb, int w, int h, const int *strong, int ns, const int *weak, int nw,
                 int *tracked_out, double *strong_color, double *weak_color, int *strong_center, int *weak_center)
{
	RefCtxView *c = (RefCtxView *)ctx;
	std::vector<cv::Mat> channel;
	for (int k = 0; k < 6; k++) channel.push_back(cv::Mat(h, w, CV_8UC1, (void *)(planes6 + (size_t)k * w * h), (size_t)w));
	cv::Mat ycc(h, 3 * w, CV_8UC1, (void *)ycrcb, (size_t)3 * w);
	std::vector<ER> S((size_t)ns), Wk((size_t)nw);
	vector<ERs> vs(6), vw(6);
	for (int i = 0; i < ns; i++) {
		const int *r = strong + 6 * i;
		S[i] = ER(0, 0, 0, 0); S[i].bound = cv::Rect(r[1], r[2], r[3], r[4]); S[i].area = r[5];
		vs[r[0]].push_back(&S[i]);
	}
	for (int i = 0; i < nw; i++) {
		const int *r = weak + 6 * i;
		Wk[i] = ER(0, 0, 0, 0); Wk[i].bound = cv::Rect(r[1], r[2], r[3], r[4]); Wk[i].area = r[5];
		vw[r[0]].push_back(&Wk[i]);
	}
	ERs all_er;
	c->erf->er_track(vs, vw, all_er, channel, ycc);
	for (size_t i = 0; i < all_er.size(); i++) {
		ER *e = all_er[i];
		if (ns && e >= &S[0] && e < &S[0] + ns) { tracked_out[2 * i] = 0; tracked_out[2 * i + 1] = (int)(e - &S[0]); }
		else { tracked_out[2 * i] = 1; tracked_out[2 * i + 1] = (int)(e - &Wk[0]); }
	}
	for (int i = 0; i < ns; i++) {
		strong_color[3 * i] = S[i].color1; strong_color[3 * i + 1] = S[i].color2; strong_color[3 * i + 2] = S[i].color3;
		strong_center[2 * i] = S[i].center.x; strong_center[2 * i + 1] = S[i].center.y;
	}
	for (int i = 0; i < nw; i++) {
		weak_color[3 * i] = Wk[i].color1; weak_color[3 * i + 1] = Wk[i].color2; weak_color[3 * i + 2] = Wk[i].color3;
		weak_center[2 * i] = Wk[i].center.x; weak_center[2 * i + 1] = Wk[i].center.y;
	}
	return (int)all_er.size();
}

// ---- ERFilter::er_grouping (src/ER.cpp:612-692) + the duplicate removal at the head of er_ocr (src/ER.cpp:702-724) ---
// in:  all_er as rows of 11 doubles: ch, x, y, w, h, area, center_x, center_y, color1, color2, color3 (what er_track left)
// out: n_after / after[] = all_er after the call (indices into the input; inner_suppression erases),
//      bounds[] = every input ER's bound and centre afterwards (x, y, w, h, cx, cy; overlap_suppression edits them in place),
//      text_off[] / text_ers[] / text_slope[] = the Text groups (members as input indices, in the group's final order).
// dedupe != 0 additionally runs the reference's per-text duplicate removal (er_ocr's first step) on every group.
int ref_er_grouping(void *ctx, const double *rows, int n, int overlap_sup, int inner_sup, int dedupe, int *n_after, int *after, int *bounds,
                    int *text_off, int *text_ers, int text_ers_cap, double *text_slope, int text_cap)
{
	RefCtxView *c = (RefCtxView *)ctx;
	std::vector<ER> E((size_t)n);
	ERs all_er;
	for (int i = 0; i < n; i++) {
		const double *r = rows + 11 * i;
		E[i] = ER(0, 0, 0, 0);
		E[i].ch = (int)r[0]; E[i].bound = cv::Rect((int)r[1], (int)r[2], (int)r[3], (int)r[4]); E[i].area = (int)r[5];
		E[i].center = cv::Point((int)r[6], (int)r[7]); E[i].color1 = r[8]; E[i].color2 = r[9]; E[i].color3 = r[10];
		all_er.push_back(&E[i]);
	}
	vector<Text> text;
	c->erf->er_grouping(all_er, text, overlap_sup != 0, inner_sup != 0);
	if (dedupe) {
		for (int i = (int)text.size() - 1; i >= 0; i--)          // the loop header of src/ER.cpp:700-701
		{
#include "ref_er_ocr_dedupe.inc"
		}
	}
	*n_after = (int)all_er.size();
	for (size_t i = 0; i < all_er.size(); i++) after[i] = (int)(all_er[i] - &E[0]);
	for (int i = 0; i < n; i++) {
		int *b = bounds + 6 * i;
		b[0] = E[i].bound.x; b[1] = E[i].bound.y; b[2] = E[i].bound.width; b[3] = E[i].bound.height; b[4] = E[i].center.x; b[5] = E[i].center.y;
	}
	int k = 0;
	if ((int)text.size() > text_cap) return -1;
	for (size_t t = 0; t < text.size(); t++) {
		text_off[t] = k;
		text_slope[t] = text[t].slope;
		for (size_t j = 0; j < text[t].ers.size(); j++) { if (k >= text_ers_cap) return -1; text_ers[k++] = (int)(text[t].ers[j] - &E[0]); }
	}
	text_off[text.size()] = k;
	return (int)text.size();
}

// ---- OCR::chain_run (src/OCR.cpp:67-140) -------------------------------------------------------------
// the verbatim call: returns table[label] + prob[label]
double ref_chain_run(void *ctx, const uchar *crop, int w, int h, int stride, int thresh, double slope)
{
	OCR *o = ocr_of(ctx);
	cv::Mat view(h, w, CV_8UC1, (void *)crop, (size_t)stride);
	return o->chain_run(view, thresh, slope);
}

// The stages of chain_run made visible: the same member calls in the same order as src/OCR.cpp:72-84
// (threshold(255-src, OTSU) -> rotate_mat when |slope| > 0.01 -> ARAN(30) -> extract_feature), returning the
// 30x30 normalised image and the dense 1800-d feature vector as bytes (value * 255; src/OCR.cpp:203-218).
// ref_chain_run above is the unmodified path; tests check that both give the same label / probability.
int ref_ocr_features(void *ctx, const uchar *crop, int w, int h, int stride, double slope, uchar *img30, uchar *feat1800)
{
	OCR *o = ocr_of(ctx);
	cv::Mat view(h, w, CV_8UC1, (void *)crop, (size_t)stride), ocr_img;
	cv::threshold(255 - view, ocr_img, 128, 255, cv::THRESH_OTSU);
	if (abs(slope) > 0.01) {
		double rad = atan2(slope, 1);
		o->rotate_mat(ocr_img, ocr_img, rad, true);
	}
	if (ocr_img.rows < 1 || ocr_img.cols < 1) return -1;
	o->ARAN(ocr_img, ocr_img, o->img_L);
	if (img30) for (int i = 0; i < 30; i++) memcpy(img30 + 30 * i, ocr_img.ptr(i), 30);
	std::vector<svm_node> fv(8 * 15 * 15 + 1);
	o->extract_feature(ocr_img, fv.data());
	memset(feat1800, 0, 1800);
	int nnz = 0;
	for (int j = 0; fv[j].index != -1; j++, nnz++) feat1800[fv[j].index] = (uchar)lrint(fv[j].value * 255.0);
	return nnz;
}

// rotate_mat alone (src/OCR.cpp:254-360); out must hold out_cap bytes; returns rows<<16 | cols
int ref_rotate_mat(void *ctx, const uchar *src, int w, int h, int stride, double rad, int crop, uchar *out, int out_cap)
{
	OCR *o = ocr_of(ctx);
	cv::Mat view(h, w, CV_8UC1, (void *)src, (size_t)stride), dst;
	cv::Mat s = view.clone();
	o->rotate_mat(s, dst, rad, crop != 0);
	if (dst.rows * dst.cols > out_cap) return -1;
	for (int i = 0; i < dst.rows; i++) memcpy(out + (size_t)i * dst.cols, dst.ptr(i), (size_t)dst.cols);
	return (dst.rows << 16) | dst.cols;
}

} // extern "C"
