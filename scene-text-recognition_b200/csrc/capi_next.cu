// capi_next.cu -- C ABI of the rows that follow the detect path: ERFilter::er_track and the OCR
// feature path + SVM of OCR::chain_run.  Same rules as capi.cu: C++ host code, plain pointers, no CPU
// implementation of the work -- the host only prepares job records (sizes, libm constants) and moves bytes.
#include "ctx.h"

#include <cmath>
#include <cstring>
#include <algorithm>

using namespace ert;

namespace {

template <typename T>
int hmalloc(T **p, size_t n)
{
	ERT_CUDA_CHECK(cudaHostAlloc((void **)p, sizeof(T) * std::max<size_t>(n, 1), cudaHostAllocMapped));
	return 0;
}
template <typename T>
int dmalloc(T **p, size_t n)
{
	ERT_CUDA_CHECK(cudaMalloc((void **)p, sizeof(T) * std::max<size_t>(n, 1)));
	return 0;
}

void free_track(ert_ctx *c)
{
	cudaFree(c->tk.cand); cudaFree(c->tk.n_cand); cudaFree(c->tk.n_strong); cudaFree(c->tk.tracked); cudaFree(c->tk.n_tracked);
	cudaFreeHost(c->h_cand); cudaFreeHost(c->h_cand_off); cudaFreeHost(c->h_nstrong); cudaFreeHost(c->h_track_off); cudaFreeHost(c->h_tracked);
	c->tk = TrackWork{};
	c->h_cand = nullptr; c->h_cand_off = c->h_nstrong = c->h_track_off = c->h_tracked = nullptr;
	c->track_frames_cap = 0; c->track_cand_cap = 0;
}

// candidate buffers for n_frames frames of up to cand_cap strong + weak regions each
int ensure_track(ert_ctx *c, int n_frames, int cand_cap)
{
	cand_cap = std::max(cand_cap, 32);
	if (n_frames <= c->track_frames_cap && (size_t)n_frames * cand_cap <= (size_t)c->track_frames_cap * c->track_cand_cap) {
		c->tk.cand_cap = cand_cap;
		return 0;
	}
	const int F = std::max(n_frames, c->track_frames_cap), C = std::max(cand_cap, c->track_cand_cap);
	free_track(c);
	const size_t tot = (size_t)F * C;
	if (dmalloc(&c->tk.cand, tot) || dmalloc(&c->tk.tracked, tot) || dmalloc(&c->tk.n_cand, (size_t)F) || dmalloc(&c->tk.n_strong, (size_t)F) ||
	    dmalloc(&c->tk.n_tracked, (size_t)F)) return -1;
	if (hmalloc(&c->h_cand, tot) || hmalloc(&c->h_tracked, tot) || hmalloc(&c->h_cand_off, (size_t)F + 1) || hmalloc(&c->h_track_off, (size_t)F + 1) ||
	    hmalloc(&c->h_nstrong, (size_t)F)) return -1;
	c->track_frames_cap = F; c->track_cand_cap = C;
	c->tk.cand_cap = cand_cap;
	return 0;
}

int finish_track(ert_ctx *c, int n_frames, const ert_track_result **out)
{
	ERT_CUDA_CHECK(cudaStreamSynchronize(c->stream));
	c->track_pending = false;
	ert_track_result &r = c->tres;
	r.n_frames = n_frames;
	r.cand_offset = c->h_cand_off; r.n_strong = c->h_nstrong; r.cand = c->h_cand;
	r.track_offset = c->h_track_off; r.tracked = c->h_tracked;
	float ms = 0.f;
	cudaEventElapsedTime(&ms, c->ev[10], c->ev[11]);
	r.track_ms = (double)ms;
	if (out) *out = &r;
	return 0;
}

// ---- OCR::chain_run job records ---------------------------------------------------------------------
// The host evaluates, with the same C library calls the reference makes, everything that is not per-pixel work:
// rotate_mat's output geometry (src/OCR.cpp:256-287) and ARAN's target size (src/OCR.cpp:396-397, 405, 419).
int make_ocr_job(OcrJob &J, const uint8_t *src, int pitch, int invert, const ert_ocr_region &R, int idx)
{
	const int L = 30;
	J.src = src; J.pitch = pitch; J.invert = invert;
	J.x0 = R.x; J.y0 = R.y; J.w = R.w; J.h = R.h;
	J.rot = 0; J.cs = 1.0; J.sn = 0.0; J.cx = J.cy = 0; J.min_x = J.min_y = J.crop_h = 0;
	int sw = R.w, sh = R.h;
	if (std::abs(R.slope) > 0.01) {
		const double rad = atan2(R.slope, 1);
		const int cols = R.w, rows = R.h;
		const int x0 = (int)((cols - 1) / 2.0), y0 = (int)((rows - 1) / 2.0);
		const int x1 = 0 - x0, y1 = 0 - y0, x2 = (cols - 1) - x0, y2 = 0 - y0;
		const int x3 = (cols - 1) - x0, y3 = (rows - 1) - y0, x4 = 0 - x0, y4 = (rows - 1) - y0;
		const int nx1 = (int)round(x1 * cos(rad) - y1 * sin(rad)), ny1 = (int)round(x1 * sin(rad) + y1 * cos(rad));
		const int nx2 = (int)round(x2 * cos(rad) - y2 * sin(rad)), ny2 = (int)round(x2 * sin(rad) + y2 * cos(rad));
		const int nx3 = (int)round(x3 * cos(rad) - y3 * sin(rad)), ny3 = (int)round(x3 * sin(rad) + y3 * cos(rad));
		const int nx4 = (int)round(x4 * cos(rad) - y4 * sin(rad)), ny4 = (int)round(x4 * sin(rad) + y4 * cos(rad));
		const int max_x = std::max(nx1, std::max(nx2, std::max(nx3, nx4))), max_y = std::max(ny1, std::max(ny2, std::max(ny3, ny4)));
		const int min_x = std::min(nx1, std::min(nx2, std::min(nx3, nx4))), min_y = std::min(ny1, std::min(ny2, std::min(ny3, ny4)));
		const int crop_height = (int)((nx2 - nx1) * tan(rad) * 0.5);
		J.cs = cos(rad); J.sn = sin(rad); J.cx = x0; J.cy = y0; J.min_x = min_x; J.min_y = min_y;
		sw = max_x - min_x + 1;
		if (max_y - min_y + 1 - 2 * crop_height <= 0) { J.rot = 2; J.crop_h = 0; sh = max_y - min_y + 1; }
		else { J.rot = 1; J.crop_h = crop_height; sh = max_y - min_y + 1 - 2 * crop_height; }
	}
	if (sw < 1 || sh < 1) { set_error("region %d: empty image after rotation", idx); return -1; }
	J.sw = sw; J.sh = sh;
	const double R1 = (sw > sh) ? (double)sh / sw : (double)sw / sh;
	const int minor = (int)(L * pow(R1, 0.5));
	if (minor < 1) { set_error("region %d: %dx%d collapses to an empty image in ARAN (cv::resize would throw)", idx, sw, sh); return -1; }
	J.dw = (sw > sh) ? L : minor; J.dh = (sw > sh) ? minor : L;
	J.offx = J.offy = 0;
	if (J.dw > J.dh) J.offy = (L - J.dh) / 2; else J.offx = (L - J.dw) / 2;
	return 0;
}

// the class table of the reference's OCR (src/OCR.cpp:10): 10 digits, 52 letters, 3 symbols
const char OCR_TABLE[] = "0123456789ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz&()";

int run_ocr(ert_ctx *c, const std::vector<OcrJob> &jobs, bool with_svm, const ert_ocr_result **out)
{
	const int n = (int)jobs.size();
	cudaStream_t st = c->stream;
	ert_ocr_result &r = c->ores;
	r = ert_ocr_result{};
	r.n = n; r.nr_class = with_svm ? c->svm.nr_class : 0;
	if (n == 0) { if (out) *out = &r; return 0; }
	const size_t feat_b = (size_t)n * 1800, img_b = (size_t)n * 900;
	if (c->o2.ensure(sizeof(OcrJob) * (size_t)n) || c->o3.ensure(feat_b) || c->o4.ensure(img_b)) return -1;
	ERT_CUDA_CHECK(cudaMemcpyAsync(c->o2.p, jobs.data(), sizeof(OcrJob) * (size_t)n, cudaMemcpyHostToDevice, st));
	ERT_CUDA_CHECK(cudaMemsetAsync(c->o4.p, 0, img_b, st));
	ERT_CUDA_CHECK(cudaEventRecord(c->ev[10], st));
	if (launch_ocr_features((const OcrJob *)c->o2.p, n, (uint8_t *)c->o3.p, (uint8_t *)c->o4.p, st)) return -1;
	c->ocr_feat.resize(feat_b); c->ocr_img.resize(img_b);
	double *d_label = nullptr, *d_prob = nullptr;
	if (with_svm) {
		const SvmHost &m = c->svm;
		if (m.dims != 1800) { set_error("the loaded SVM has %d dimensions, OCR features have 1800", m.dims); return -1; }
		if (c->s1.ensure(svm_ws_bytes(m.dev(), n)) || c->s3.ensure(sizeof(double) * (size_t)n * (m.nr_class + 1))) return -1;
		d_label = (double *)c->s3.p; d_prob = d_label + n;
		uint8_t *tcws = nullptr;
		if (m.use_tc && m.d_svj) { if (c->s4.ensure(svm_tc_ws_bytes(n))) return -1; tcws = (uint8_t *)c->s4.p; }
		if (launch_svm_predict(m.dev(), nullptr, (const uint8_t *)c->o3.p, n, (double *)c->s1.p, d_label, d_prob, st, tcws)) return -1;
	}
	ERT_CUDA_CHECK(cudaEventRecord(c->ev[11], st));
	ERT_CUDA_CHECK(cudaMemcpyAsync(c->ocr_feat.data(), c->o3.p, feat_b, cudaMemcpyDeviceToHost, st));
	ERT_CUDA_CHECK(cudaMemcpyAsync(c->ocr_img.data(), c->o4.p, img_b, cudaMemcpyDeviceToHost, st));
	std::vector<double> lab;
	if (with_svm) {
		lab.resize((size_t)n);
		c->ocr_prob.resize((size_t)n * c->svm.nr_class);
		ERT_CUDA_CHECK(cudaMemcpyAsync(lab.data(), d_label, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, st));
		ERT_CUDA_CHECK(cudaMemcpyAsync(c->ocr_prob.data(), d_prob, sizeof(double) * (size_t)n * c->svm.nr_class, cudaMemcpyDeviceToHost, st));
	}
	ERT_CUDA_CHECK(cudaStreamSynchronize(st));
	if (with_svm && c->svm.use_tc && c->svm.d_svj && svm_gemm_flag_check((uint8_t *)c->s4.p, n)) return -1;
	float ms = 0.f;
	cudaEventElapsedTime(&ms, c->ev[10], c->ev[11]);
	r.ocr_ms = (double)ms;
	r.feat = c->ocr_feat.data(); r.img = c->ocr_img.data();
	if (with_svm) {
		const int k = c->svm.nr_class;
		c->ocr_label.resize((size_t)n); c->ocr_value.resize((size_t)n);
		for (int i = 0; i < n; i++) {
			const int label = (int)lab[i];                        // const int label = (int)svm_predict_probability(...)  (src/OCR.cpp:92)
			c->ocr_label[i] = label;
			const double prob = (label >= 0 && label < k) ? c->ocr_prob[(size_t)i * k + label] : 0.0;   // pv[label]
			const char ch = (label >= 0 && label < (int)sizeof(OCR_TABLE) - 1) ? OCR_TABLE[label] : '?';
			c->ocr_value[i] = ch + prob;                          // return table[label] + prob  (src/OCR.cpp:139)
		}
		r.value = c->ocr_value.data(); r.label = c->ocr_label.data(); r.prob_all = c->ocr_prob.data();
	}
	if (out) *out = &r;
	return 0;
}

int ocr_plane_common(ert_ctx *c, const uint8_t *plane, int W, int H, int stride, const ert_ocr_region *regions, int n, bool with_svm,
                     const ert_ocr_result **out)
{
	if (!c || !plane || W < 1 || H < 1 || stride < W || n < 0 || (n && !regions)) { set_error("bad arguments"); return -1; }
	if (with_svm && !c->svm.loaded) { set_error("svm model is not loaded"); return -1; }
	ERT_CUDA_CHECK(cudaSetDevice(c->device));
	const int pitch = extract_pitch(W);
	if (c->o0.ensure((size_t)pitch * H)) return -1;
	std::vector<OcrJob> jobs((size_t)n);
	for (int i = 0; i < n; i++) {
		const ert_ocr_region &R = regions[i];
		if (R.x < 0 || R.y < 0 || R.w < 1 || R.h < 1 || R.x + R.w > W || R.y + R.h > H) { set_error("region %d outside the plane", i); return -1; }
		if (make_ocr_job(jobs[i], (const uint8_t *)c->o0.p, pitch, 0, R, i)) return -1;
	}
	ERT_CUDA_CHECK(cudaMemcpy2DAsync(c->o0.p, (size_t)pitch, plane, (size_t)stride, (size_t)W, (size_t)H, cudaMemcpyHostToDevice, c->stream));
	return run_ocr(c, jobs, with_svm, out);
}

} // namespace

namespace ert {

// enqueue er_track for the batch whose classify results sit in the context's device buffers (BGR layout)
int enqueue_track(ert_ctx *c, int n_frames)
{
	if (ensure_track(c, n_frames, 6 * c->pool_cap)) return -1;
	cudaStream_t st = c->pending ? c->work_stream() : c->stream;   // fused with a batch: the batch's post stream; stand-alone: the context's stream
	ERT_CUDA_CHECK(cudaEventRecord(c->ev[10], st));
	if (launch_track_gather(c->tk, n_frames, c->d_out_nodes, c->d_out_pool, c->d_out_counts, c->d_label, c->kept_cap, c->pool_cap, st)) return -1;
	if (launch_calc_color(c->tk, n_frames, c->d_ycc, c->ycc_bytes, c->pitch, st)) return -1;
	if (launch_track(c->tk, n_frames, c->h_cand, c->h_cand_off, c->h_nstrong, c->h_track_off, c->h_tracked, st)) return -1;
	ERT_CUDA_CHECK(cudaEventRecord(c->ev[11], st));
	c->launches += 4;
	c->track_pending = true;
	c->pending_frames = n_frames;
	return 0;
}

void free_next(ert_ctx *c)
{
	free_track(c);
	c->o0.release(); c->o1.release(); c->o2.release(); c->o3.release(); c->o4.release();
}

} // namespace ert

extern "C" {

void *ert_host_alloc(size_t bytes)
{
	void *p = nullptr;
	if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) { set_error("cudaHostAlloc(%zu) failed", bytes); return nullptr; }
	return p;
}
void ert_host_free(void *p) { if (p) cudaFreeHost(p); }

int ert_er_track(ert_ctx *c, const ert_track_result **out)
{
	if (!c) { set_error("bad arguments"); return -1; }
	ERT_CUDA_CHECK(cudaSetDevice(c->device));
	if (c->track_pending) return finish_track(c, c->pending_frames, out);
	if (c->frames_cap != -1 || c->pending_planes < 6 || c->pending_upto < ERT_STAGE_CLASSIFY) {
		set_error("ert_er_track: the context's last batch was not a classified BGR batch");
		return -1;
	}
	const int n_frames = c->pending_planes / 6;
	if (enqueue_track(c, n_frames)) return -1;
	return finish_track(c, n_frames, out);
}

// strong / weak rows -> candidate records; frame = either one BGR image (converted on the device) or its three
// Y / Cr / Cb planes as compute_channels delivers them (channel[0..2], src/ER.cpp:122-124)
static int track_regions_common(ert_ctx *c, const uint8_t *bgr, const uint8_t *const ycc[3], int W, int H, int stride, const int32_t *strong, int ns,
                                const int32_t *weak, int nw, const ert_track_result **out)
{
	const int min_stride = bgr ? 3 * W : W;
	if (!c || W < 1 || H < 1 || stride < min_stride || ns < 0 || nw < 0 || (ns && !strong) || (nw && !weak)) { set_error("bad arguments"); return -1; }
	const int n = ns + nw;
	std::vector<ert_tracked> hc((size_t)std::max(n, 1));
	for (int i = 0; i < n; i++) {
		const int32_t *r = i < ns ? strong + 6 * i : weak + 6 * (i - ns);
		if (r[0] < 0 || r[0] > 5 || r[1] < 0 || r[2] < 0 || r[3] < 1 || r[4] < 1 || r[1] + r[3] > W || r[2] + r[4] > H) {
			set_error("region %d: channel or rectangle outside the frame", i);
			return -1;
		}
		if (i != 0 && i != ns && r[0] < r[-6]) { set_error("region %d: rows must be channel-major (strong[0..5], weak[0..5])", i); return -1; }
		ert_tracked &t = hc[i];
		t.plane = r[0]; t.pool_index = -1; t.node = -1; t.label = i < ns ? ERT_LABEL_STRONG : ERT_LABEL_WEAK; t.level = 0; t.area = r[5];
		t.x = r[1]; t.y = r[2]; t.w = r[3]; t.h = r[4];
		t.center_x = r[1] + r[3] / 2; t.center_y = r[2] + r[4] / 2;
		t.color1 = t.color2 = t.color3 = 0.0;
	}
	ERT_CUDA_CHECK(cudaSetDevice(c->device));
	cudaStream_t st = c->stream;
	const int pitch = extract_pitch(W);
	const size_t in_b = (size_t)stride * H, plane_b = (size_t)pitch * H;
	if ((bgr && c->o0.ensure(in_b)) || c->o1.ensure(plane_b * 3)) return -1;
	// the track buffers may be shared with a batch in flight: this entry point is synchronous and owns them while it runs
	if (ensure_track(c, 1, n)) return -1;
	if (bgr) {
		ERT_CUDA_CHECK(cudaMemcpyAsync(c->o0.p, bgr, in_b, cudaMemcpyHostToDevice, st));
		if (launch_channels((const uint8_t *)c->o0.p, in_b, stride, W, H, 1, (uint8_t *)c->o1.p, pitch, st)) return -1;
	} else {
		for (int k = 0; k < 3; k++)
			ERT_CUDA_CHECK(cudaMemcpy2DAsync((uint8_t *)c->o1.p + (size_t)k * plane_b, (size_t)pitch, ycc[k], (size_t)stride, (size_t)W, (size_t)H, cudaMemcpyHostToDevice, st));
	}
	const int32_t cnt[2] = {n, ns};
	if (n) ERT_CUDA_CHECK(cudaMemcpyAsync(c->tk.cand, hc.data(), sizeof(ert_tracked) * (size_t)n, cudaMemcpyHostToDevice, st));
	ERT_CUDA_CHECK(cudaMemcpyAsync(c->tk.n_cand, &cnt[0], sizeof(int32_t), cudaMemcpyHostToDevice, st));
	ERT_CUDA_CHECK(cudaMemcpyAsync(c->tk.n_strong, &cnt[1], sizeof(int32_t), cudaMemcpyHostToDevice, st));
	ERT_CUDA_CHECK(cudaEventRecord(c->ev[10], st));
	if (launch_calc_color(c->tk, 1, (const uint8_t *)c->o1.p, plane_b, pitch, st)) return -1;
	if (launch_track(c->tk, 1, c->h_cand, c->h_cand_off, c->h_nstrong, c->h_track_off, c->h_tracked, st)) return -1;
	ERT_CUDA_CHECK(cudaEventRecord(c->ev[11], st));
	return finish_track(c, 1, out);     // synchronises: the staging vector and cnt[] may go out of scope afterwards
}

int ert_er_track_regions(ert_ctx *c, const uint8_t *bgr, int W, int H, int stride, const int32_t *strong, int ns, const int32_t *weak, int nw,
                         const ert_track_result **out)
{
	if (!bgr) { set_error("bad arguments"); return -1; }
	return track_regions_common(c, bgr, nullptr, W, H, stride, strong, ns, weak, nw, out);
}

int ert_er_track_regions_ycc(ert_ctx *c, const uint8_t *y, const uint8_t *cr, const uint8_t *cb, int W, int H, int stride, const int32_t *strong, int ns,
                             const int32_t *weak, int nw, const ert_track_result **out)
{
	if (!y || !cr || !cb) { set_error("bad arguments"); return -1; }
	const uint8_t *const ycc[3] = {y, cr, cb};
	return track_regions_common(c, nullptr, ycc, W, H, stride, strong, ns, weak, nw, out);
}

int ert_ocr_chain_run_plane(ert_ctx *c, const uint8_t *plane, int W, int H, int stride, const ert_ocr_region *regions, int n,
                            const ert_ocr_result **out)
{
	return ocr_plane_common(c, plane, W, H, stride, regions, n, true, out);
}

int ert_ocr_features_plane(ert_ctx *c, const uint8_t *plane, int W, int H, int stride, const ert_ocr_region *regions, int n,
                           const ert_ocr_result **out)
{
	return ocr_plane_common(c, plane, W, H, stride, regions, n, false, out);
}

int ert_ocr_chain_run_batch(ert_ctx *c, const ert_ocr_region *regions, int n, const ert_ocr_result **out)
{
	if (!c || n < 0 || (n && !regions)) { set_error("bad arguments"); return -1; }
	if (!c->svm.loaded) { set_error("svm model is not loaded"); return -1; }
	if (c->frames_cap != -1 || c->pending_planes < 6) { set_error("ert_ocr_chain_run_batch: the context's last batch was not a BGR batch"); return -1; }
	ERT_CUDA_CHECK(cudaSetDevice(c->device));
	const int n_frames = c->pending_planes / 6;
	std::vector<OcrJob> jobs((size_t)n);
	for (int i = 0; i < n; i++) {
		const ert_ocr_region &R = regions[i];
		if (R.frame < 0 || R.frame >= n_frames || R.plane < 0 || R.plane > 5) { set_error("region %d: frame / channel outside the batch", i); return -1; }
		if (R.x < 0 || R.y < 0 || R.w < 1 || R.h < 1 || R.x + R.w > c->W || R.y + R.h > c->H) { set_error("region %d outside the frame", i); return -1; }
		const uint8_t *src = c->d_ycc + ((size_t)R.frame * 3 + (R.plane % 3)) * c->ycc_bytes;
		if (make_ocr_job(jobs[i], src, c->pitch, R.plane >= 3, R, i)) return -1;
	}
	return run_ocr(c, jobs, true, out);
}

} // extern "C"
