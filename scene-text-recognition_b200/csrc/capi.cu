// capi.cu -- the C ABI of libertext.so (include/ertext.h): context, model-file parsers,
// workspace management and the batched orchestration of the kernels.  C++ host code, no torch,
// no CPU implementation of the path: every entry point fails if the CUDA device is unavailable.
#include "../../include/ertext.h"
#include "common.cuh"
#include "kernels.h"

#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>
#include <fstream>
#include <sstream>
#include <algorithm>
#include <mutex>

namespace ert {

static thread_local char g_err[1024] = "";
void set_error(const char *fmt, ...)
{
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_err, sizeof g_err, fmt, ap);
	va_end(ap);
}

// ---------------------------------------------------------------------------------------------
// result compaction: per-plane fixed-capacity device arrays -> contiguous host-mapped arrays
// ---------------------------------------------------------------------------------------------
__global__ void k_compact_results(int n_planes, int node_cap, int pool_cap, const int32_t *__restrict__ counts,
                                  const OutNode *__restrict__ nodes, const int32_t *__restrict__ pool, const int32_t *__restrict__ label,
                                  const double *__restrict__ ss, const double *__restrict__ ws, const uint8_t *__restrict__ hist,
                                  int32_t *__restrict__ h_node_off, int32_t *__restrict__ h_pool_off, OutNode *__restrict__ h_nodes,
                                  int32_t *__restrict__ h_pool, int32_t *__restrict__ h_label, double *__restrict__ h_ss,
                                  double *__restrict__ h_ws, uint8_t *__restrict__ h_hist, int with_labels,
                                  const int32_t *__restrict__ order_sens, int32_t *__restrict__ h_order_sens)
{
	const int plane = blockIdx.x;
	if (threadIdx.x == 0 && h_order_sens) h_order_sens[plane] = order_sens ? order_sens[plane] : -1;
	int noff = 0, poff = 0;
	for (int p = 0; p < plane; p++) { noff += counts[2 * p]; poff += counts[2 * p + 1]; }
	const int nn = counts[2 * plane], np = counts[2 * plane + 1];
	if (threadIdx.x == 0) {
		h_node_off[plane] = noff; h_pool_off[plane] = poff;
		if (plane == n_planes - 1) { h_node_off[n_planes] = noff + nn; h_pool_off[n_planes] = poff + np; }
	}
	const uint4 *src = reinterpret_cast<const uint4 *>(nodes + (size_t)plane * node_cap);
	uint4 *dst = reinterpret_cast<uint4 *>(h_nodes + noff);
	for (int i = threadIdx.x; i < nn * 2; i += blockDim.x) dst[i] = src[i];
	for (int i = threadIdx.x; i < np; i += blockDim.x) {
		const size_t s = (size_t)plane * pool_cap + i;
		h_pool[poff + i] = pool[s];
		if (with_labels) { h_label[poff + i] = label[s]; h_ss[poff + i] = ss[s]; h_ws[poff + i] = ws[s]; }
		else { h_label[poff + i] = 0; h_ss[poff + i] = 0.0; h_ws[poff + i] = 0.0; }
	}
	if (h_hist && with_labels) {
		const uint32_t *hs = reinterpret_cast<const uint32_t *>(hist + (size_t)plane * pool_cap * 1024);
		uint32_t *hd = reinterpret_cast<uint32_t *>(h_hist + (size_t)poff * 1024);
		for (int i = threadIdx.x; i < np * 256; i += blockDim.x) hd[i] = hs[i];
	}
}

} // namespace ert

#include "ctx.h"

using namespace ert;

namespace {

void free_workspace(ert_ctx *c)
{
	cudaFree(c->d_ycc); cudaFree(c->d_planes);
	cudaFree(c->wk.par); cudaFree(c->wk.attr); cudaFree(c->wk.node_key); cudaFree(c->wk.node_count); cudaFree(c->wk.start_key); cudaFree(c->wk.ring_rec);
	cudaFree(c->wk.reach_root); cudaFree(c->wk.lone_level); cudaFree(c->wk.kept); cudaFree(c->wk.kept_count); cudaFree(c->wk.status);
	cudaFree(c->d_nms_scratch); cudaFree(c->d_out_nodes); cudaFree(c->d_out_pool); cudaFree(c->d_out_counts); cudaFree(c->d_label);
	cudaFree(c->d_ss); cudaFree(c->d_ws); cudaFree(c->d_hist); cudaFree(c->d_order_sens); cudaFreeHost(c->h_order_sens);
	c->d_order_sens = nullptr; c->h_order_sens = nullptr;
	cudaFreeHost(c->h_node_off); cudaFreeHost(c->h_pool_off); cudaFreeHost(c->h_pool); cudaFreeHost(c->h_label);
	cudaFreeHost(c->h_nodes); cudaFreeHost(c->h_ss); cudaFreeHost(c->h_ws); cudaFreeHost(c->h_hist); cudaFreeHost(c->h_status);
	c->d_ycc = nullptr; c->d_planes = nullptr; c->wk = ExtractWork{};
	c->d_nms_scratch = nullptr; c->d_out_nodes = nullptr; c->d_out_pool = nullptr; c->d_out_counts = nullptr; c->d_label = nullptr;
	c->d_ss = c->d_ws = nullptr; c->d_hist = nullptr;
	c->h_node_off = c->h_pool_off = c->h_pool = c->h_label = nullptr; c->h_nodes = nullptr; c->h_ss = c->h_ws = nullptr;
	c->h_hist = nullptr; c->h_status = nullptr;
	c->planes_cap = 0; c->frames_cap = 0; c->W = c->H = 0;
	c->cap_np = c->cap_ycc = c->cap_ring = 0; c->cap_nodes = 0;
	c->table_planes = 0; c->table_W = c->table_H = 0;
}

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

// compact tables of a cascade for u8 histograms + upload (h.stumps / stage_len / stage_thr are filled)
int upload_cascade(CascadeHost &h)
{
	h.cpcn.resize(h.stumps.size()); h.dimthr.resize(h.stumps.size());
	for (size_t j = 0; j < h.stumps.size(); j++) {
		const Stump &s0 = h.stumps[j];
		h.cpcn[j] = make_double2(s0.cp, s0.cn);
		const double ct = ceil(s0.thr);       // integer counts: h < thr  <=>  h < ceil(thr)
		const uint32_t ithr = (uint32_t)(ct < 0.0 ? 0.0 : (ct > 65535.0 ? 65535.0 : ct));
		h.dimthr[j] = (uint32_t)s0.dim | (ithr << 16);
	}
	if (dev_upload(&h.d_stumps, h.stumps) || dev_upload(&h.d_len, h.stage_len) || dev_upload(&h.d_thr, h.stage_thr) ||
	    dev_upload(&h.d_cpcn, h.cpcn) || dev_upload(&h.d_dimthr, h.dimthr)) return -1;
	return 0;
}

int ensure_cascade_scratch(ert_ctx *c, int rows, int planes)
{
	CascadeScratch &sc = c->csc;
	if (rows <= sc.rows_cap && planes <= sc.planes_cap) return 0;
	ERT_CUDA_CHECK(cudaStreamSynchronize(c->stream));
	if (c->post_stream) ERT_CUDA_CHECK(cudaStreamSynchronize(c->post_stream));
	const int R = std::max(rows, sc.rows_cap), P = std::max(planes, sc.planes_cap);
	cudaFree(sc.stage_sum); cudaFree(sc.done); cudaFree(sc.pool_prefix);
	sc = CascadeScratch{};
	if (dmalloc(&sc.stage_sum, (size_t)R * 32) || dmalloc(&sc.done, (size_t)R) || dmalloc(&sc.pool_prefix, (size_t)P + 1)) return -1;
	ERT_CUDA_CHECK(cudaMemset(sc.done, 0, sizeof(uint32_t) * (size_t)R));
	sc.rows_cap = R; sc.planes_cap = P;
	return 0;
}

// Workspace by CAPACITY: buffers grow to the largest (planes x pixels) seen and are reused for anything smaller, so
// alternating plane sizes (the scales of a pyramid, BASELINE config 4) costs no allocation after the first pass.
// Nothing in the workspace needs clearing between batches (sparse slots are initialised by the kernel that creates them).
int ensure_workspace(ert_ctx *c, int n_planes, int W, int H)
{
	if (W < 1 || H < 1 || (long long)W * H > (1ll << KEY_IDX_BITS) || W > 8191 || H > 8191) {
		set_error("unsupported plane size %dx%d (max 8191 per side, %d pixels)", W, H, 1 << KEY_IDX_BITS);
		return -1;
	}
	const int pitch = extract_pitch(W);
	const size_t N = (size_t)W * H;
	// slots of the global node arrays per plane: the tile-local nodes that leave their tiles (seam-touching ones and
	// interior ones the reference keeps) -- measured 13 per 64x32 tile on the benchmark frames, ~300 on uniform noise;
	// default one slot per 4 pixels (at least 8192: any plane of up to 4 tiles fits whatever it holds); with a tiny MIN_AREA
	// nearly every tile-local node is one the reference keeps, so every pixel may need a slot
	const size_t node_cap = c->node_cap_user > 0 ? (size_t)c->node_cap_user : std::max<size_t>(8192, c->prm.min_area >= 32 ? N / 4 : N);
	const size_t need_np = node_cap * n_planes, need_ycc = (size_t)pitch * H * n_planes, need_ring = ring_words_per_plane(W, H) * n_planes;
	if (n_planes <= c->planes_cap && need_np <= c->cap_np && need_ycc <= c->cap_ycc && need_ring <= c->cap_ring) {
		c->W = W; c->H = H; c->pitch = pitch; c->ycc_bytes = (size_t)pitch * H;
		c->wk.node_cap = (int)node_cap;
		return 0;
	}
	const int P = std::max(n_planes, c->planes_cap);
	const size_t cap_np = std::max(need_np, c->cap_np), cap_ycc = std::max(need_ycc, c->cap_ycc), cap_ring = std::max(need_ring, c->cap_ring);
	ERT_CUDA_CHECK(cudaSetDevice(c->device));
	ERT_CUDA_CHECK(cudaStreamSynchronize(c->stream));
	if (c->post_stream) ERT_CUDA_CHECK(cudaStreamSynchronize(c->post_stream));
	free_workspace(c);
	c->W = W; c->H = H; c->pitch = pitch;
	c->ycc_bytes = (size_t)pitch * H;
	if (dmalloc(&c->d_ycc, cap_ycc + 256)) return -1;
	ERT_CUDA_CHECK(cudaMemset(c->d_ycc, 0, cap_ycc + 256));
	if (dmalloc(&c->d_planes, (size_t)P)) return -1;
	if (dmalloc(&c->wk.par, cap_np) || dmalloc(&c->wk.attr, cap_np) || dmalloc(&c->wk.node_key, cap_np) || dmalloc(&c->wk.start_key, (size_t)4 * P)) return -1;
	c->wk.node_cap = (int)node_cap;
	if (dmalloc(&c->wk.ring_rec, cap_ring)) return -1;
	if (dmalloc(&c->wk.node_count, (size_t)P) || dmalloc(&c->wk.reach_root, (size_t)P) || dmalloc(&c->wk.lone_level, (size_t)P)) return -1;
	if (dmalloc(&c->wk.kept, (size_t)P * c->kept_cap) || dmalloc(&c->wk.kept_count, (size_t)P) || dmalloc(&c->wk.status, 1)) return -1;
	ERT_CUDA_CHECK(cudaMemset(c->wk.status, 0, sizeof(uint32_t)));
	c->wk.node_blocks = 32;
	c->wk.tile_cfg = c->tile_cfg;
	c->wk.seam_list = c->seam_list;
	c->wk.post_ctas = c->post_ctas_per_sm * c->sm_count;
	c->wk.node_cap = (int)node_cap;
	c->wk.prof = c->d_prof;
	c->nms_stride = nms_scratch_stride(c->kept_cap);
	if (dmalloc(&c->d_nms_scratch, c->nms_stride * P)) return -1;
	if (dmalloc(&c->d_out_nodes, (size_t)P * c->kept_cap) || dmalloc(&c->d_out_pool, (size_t)P * c->pool_cap)) return -1;
	if (dmalloc(&c->d_out_counts, (size_t)2 * P) || dmalloc(&c->d_label, (size_t)P * c->pool_cap)) return -1;
	if (dmalloc(&c->d_ss, (size_t)P * c->pool_cap) || dmalloc(&c->d_ws, (size_t)P * c->pool_cap)) return -1;
	if (dmalloc(&c->d_hist, (size_t)P * c->pool_cap * 1024)) return -1;
	if (hmalloc(&c->h_node_off, (size_t)P + 1) || hmalloc(&c->h_pool_off, (size_t)P + 1)) return -1;
	if (dmalloc(&c->d_order_sens, (size_t)P) || hmalloc(&c->h_order_sens, (size_t)P)) return -1;
	if (hmalloc(&c->h_nodes, (size_t)P * c->kept_cap) || hmalloc(&c->h_pool, (size_t)P * c->pool_cap)) return -1;
	if (hmalloc(&c->h_label, (size_t)P * c->pool_cap) || hmalloc(&c->h_ss, (size_t)P * c->pool_cap) || hmalloc(&c->h_ws, (size_t)P * c->pool_cap)) return -1;
	// the pooled-region histograms travel to the host only on request: 1 KB per region, 100 MB of pinned memory at the defaults
	if (hmalloc(&c->h_hist, c->return_hist ? (size_t)P * c->pool_cap * 1024 : 1) || hmalloc(&c->h_status, 1)) return -1;
	c->planes_cap = P; c->cap_np = cap_np; c->cap_ycc = cap_ycc; c->cap_ring = cap_ring;
	return 0;
}

int set_plane_table(ert_ctx *c, int n_planes, bool bgr_mode)
{
	std::vector<PlaneSrc> t((size_t)n_planes);
	for (int p = 0; p < n_planes; p++) {
		if (bgr_mode) {
			const int ppf = c->planes_per_frame;
			const int f = p / ppf, ch = p % ppf;
			t[p].z = f * 3 + (ch % 3);
			t[p].src = c->d_ycc + (size_t)t[p].z * c->ycc_bytes;
			t[p].invert = ch >= 3;
		} else {
			t[p].z = p;
			t[p].src = c->d_ycc + (size_t)p * c->ycc_bytes;
			t[p].invert = 0;
		}
	}
	// the tile kernel's view of the plane buffer: (x, y, source plane), one haloed box per tile through the TMA unit
	if (make_tile_tensor_map(&c->wk.tmap, c->d_ycc, c->W, c->H, c->pitch, bgr_mode ? 3 * (n_planes / c->planes_per_frame) : n_planes)) return -1;
	ERT_CUDA_CHECK(cudaMemcpyAsync(c->d_planes, t.data(), sizeof(PlaneSrc) * n_planes, cudaMemcpyHostToDevice, c->stream));
	// the table is tiny; make the pageable staging copy safe before `t` dies
	ERT_CUDA_CHECK(cudaStreamSynchronize(c->stream));
	return 0;
}

ExtractParams make_extract_params(const ert_ctx *c, int n_planes)
{
	ExtractParams P;
	P.W = c->W; P.H = c->H; P.pitch = c->pitch; P.n_planes = n_planes;
	P.hi = 255 / c->prm.thresh_step + 1;
	P.qscale = (float)(1.0 / (double)c->prm.thresh_step);
	P.min_area = c->prm.min_area;
	P.kept_cap = c->kept_cap;
	P.node_cap = c->wk.node_cap;
	return P;
}

NmsParams make_nms_params(const ert_ctx *c, int W, int H)
{
	NmsParams P;
	P.W = W; P.H = H; P.N = (size_t)c->wk.node_cap;
	P.kept_cap = c->kept_cap; P.pool_cap = c->pool_cap;
	P.min_area = c->prm.min_area; P.max_area = c->prm.max_area; P.stability_t = c->prm.stability_t;
	P.overlap_coef = c->prm.overlap_coef;
	P.sequential_walk = c->nms_sequential;
	return P;
}

// FIFO of the SM-filling tile kernels across the contexts of a device.  Several contexts (streams) are used
// round-robin to overlap one batch's narrow kernels (seams, refit, NMS, cascades) with another batch's tile kernel.
// Left alone, the block scheduler favours whatever grid has CTAs ready: with >= 4 batches in flight the newest
// batches' tile kernels (49 k CTAs each) crowd out the last small kernels of the OLDEST batch -- the one the host is
// waiting for -- and the pipeline degenerates to lock-step (measured: 3.4 -> 19 ms per step at 4 contexts).  Chaining
// the tile kernels by an event keeps them in submission order; a tile kernel fills the GPU by itself, so nothing is lost.
struct TileFifo { std::mutex mu; cudaEvent_t last[64] = {}; const ert_ctx *owner[64] = {}; };
TileFifo g_fifo;

// planes already in d_ycc + plane table set: run extract -> nms -> classify -> compaction (all async)
int enqueue_pipeline(ert_ctx *c, int n_planes, int upto)
{
	cudaStream_t st = c->stream;
	const ExtractParams EP = make_extract_params(c, n_planes);
	const int dslot = c->device & 63;
	// the device status word belongs to ONE batch: an overflow of an earlier batch must not poison this one
	ERT_CUDA_CHECK(cudaMemsetAsync(c->wk.status, 0, sizeof(uint32_t), st));
	if (c->tile_fifo) {
		std::lock_guard<std::mutex> lk(g_fifo.mu);
		if (g_fifo.last[dslot] && g_fifo.owner[dslot] != c) ERT_CUDA_CHECK(cudaStreamWaitEvent(st, g_fifo.last[dslot], 0));
	}
	if (launch_extract(EP, c->d_planes, c->wk, st, c->ev[8], c->ev[9], c->work_stream())) return -1;
	const cudaStream_t st_main = st;
	st = c->work_stream();
	if (c->tile_fifo) {
		std::lock_guard<std::mutex> lk(g_fifo.mu);
		g_fifo.last[dslot] = c->ev[9]; g_fifo.owner[dslot] = c;      // ev[9] is recorded right after the tile kernel
	}
	c->launches += 5;
	ERT_CUDA_CHECK(cudaEventRecord(c->ev[2], st));
	const NmsParams NP = make_nms_params(c, c->W, c->H);
	if (launch_nms(NP, n_planes, c->wk.kept, c->wk.kept_count, c->wk.attr, c->wk.reach_root, c->wk.lone_level, nullptr, nullptr,
	               c->d_nms_scratch, c->nms_stride, c->d_out_nodes, c->d_out_pool, c->d_out_counts, c->wk.status, st, c->d_order_sens)) return -1;
	c->launches += 1;
	ERT_CUDA_CHECK(cudaEventRecord(c->ev[3], st));
	const bool do_classify = upto >= ERT_STAGE_CLASSIFY;
	if (do_classify) {
		if (!c->casc[0].loaded || !c->casc[1].loaded) { set_error("classify requested but cascades are not loaded"); return -1; }
		ClassifyParams CP; CP.pitch = c->pitch; CP.pool_cap = c->pool_cap; CP.node_cap = c->kept_cap;
		if (ensure_cascade_scratch(c, n_planes * c->pool_cap, n_planes)) return -1;
		if (launch_pool_prefix(c->d_out_counts, n_planes, c->pool_cap, c->csc.pool_prefix, st)) return -1;
		if (launch_lbp_hist(CP, n_planes, c->d_planes, c->d_out_nodes, c->d_out_pool, c->csc.pool_prefix, c->d_aran_tbl, c->d_hist, st)) return -1;
		if (launch_cascade_u8(c->d_hist, 1024, n_planes * c->pool_cap, c->csc.pool_prefix, n_planes, c->pool_cap, c->casc[0].dev(), c->casc[1].dev(),
		                      c->d_label, c->d_ss, c->d_ws, c->csc, st)) return -1;
		c->launches += 3;
	}
	ERT_CUDA_CHECK(cudaEventRecord(c->ev[4], st));
	k_compact_results<<<n_planes, 256, 0, st>>>(n_planes, c->kept_cap, c->pool_cap, c->d_out_counts, c->d_out_nodes, c->d_out_pool, c->d_label,
	                                           c->d_ss, c->d_ws, c->d_hist, c->h_node_off, c->h_pool_off, c->h_nodes, c->h_pool, c->h_label,
	                                           c->h_ss, c->h_ws, (c->return_hist && do_classify) ? c->h_hist : nullptr, do_classify ? 1 : 0,
	                                           c->nms_sequential ? nullptr : c->d_order_sens, c->h_order_sens);
	ERT_CUDA_CHECK(cudaGetLastError());
	c->launches += 1;
	ERT_CUDA_CHECK(cudaMemcpyAsync(c->h_status, c->wk.status, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
	ERT_CUDA_CHECK(cudaEventRecord(c->ev[5], st));
	c->pending = true; c->pending_planes = n_planes; c->pending_upto = upto;
	c->track_pending = false;
	if (upto >= ERT_STAGE_TRACK) {
		if (c->frames_cap != -1) { set_error("ERT_STAGE_TRACK needs a BGR batch (er_track reads the YCrCb frame)"); return -1; }
		if (enqueue_track(c, n_planes / 6)) return -1;
	}
	if (st != st_main) {
		// join: whatever is enqueued on the context's stream afterwards (and a synchronize on it) sees the finished batch
		ERT_CUDA_CHECK(cudaEventRecord(c->ev_post_done, st));
		ERT_CUDA_CHECK(cudaStreamWaitEvent(st_main, c->ev_post_done, 0));
	}
	return 0;
}

int finish_result(ert_ctx *c, const ert_result **out)
{
	if (!c->pending) { set_error("no batch in flight"); return -1; }
	ERT_CUDA_CHECK(cudaStreamSynchronize(c->stream));
	c->pending = false;
	ert_result &r = c->res;
	r.n_planes = c->pending_planes; r.width = c->W; r.height = c->H;
	r.node_offset = c->h_node_off; r.nodes = reinterpret_cast<const ert_node *>(c->h_nodes);
	r.pool_offset = c->h_pool_off; r.pool_node = c->h_pool; r.pool_label = c->h_label;
	r.pool_strong_score = c->h_ss; r.pool_weak_score = c->h_ws;
	r.pool_hist = (c->return_hist && c->pending_upto >= ERT_STAGE_CLASSIFY) ? c->h_hist : nullptr;
	r.status = *c->h_status;
	r.plane_order_sensitive = c->h_order_sens;
	r.order_sensitive_total = 0;
	for (int p = 0; p < r.n_planes; p++) { if (c->h_order_sens[p] < 0) { r.order_sensitive_total = -1; break; } r.order_sensitive_total += c->h_order_sens[p]; }
	float ms;
	auto el = [&](int a, int b) { ms = 0.f; cudaEventElapsedTime(&ms, c->ev[a], c->ev[b]); return (double)ms; };
	r.stage_ms[3] = el(0, 1); r.stage_ms[0] = el(1, 2); r.stage_ms[1] = el(2, 3); r.stage_ms[2] = el(3, 4); r.stage_ms[4] = el(4, 5);
	r.stage_ms[5] = el(0, 5);
	r.stage_ms[6] = el(8, 9); r.stage_ms[7] = r.stage_ms[0] - r.stage_ms[6];
	if (r.status & ERR_LOOP_GUARD) { set_error("device loop guard tripped (internal error)"); return -2; }
	if (out) *out = &r;
	return 0;
}

int check_cuda_device(int device)
{
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess || n <= 0) { set_error("no CUDA device available (%s): libertext has no CPU path", cudaGetErrorString(e)); return -1; }
	if (device < 0 || device >= n) { set_error("device %d out of range (have %d)", device, n); return -1; }
	return 0;
}

void build_aran_table(ert_ctx *c)
{
	// for minor = (int)(L * pow(R1, 0.5)) when L*sqrt(R1) is an exact integer m: what does libm give?
	const int L = 26;
	for (int m = 0; m < 64; m++) c->aran_tbl_h[m] = (uint8_t)std::min(m, L);
	for (int m = 1; m <= L; m++) {
		long long p = (long long)m * m, q = (long long)L * L, a = p, b = q;
		while (b) { long long t = a % b; a = b; b = t; }
		p /= a; q /= a;
		const double R1 = (double)p / (double)q;
		c->aran_tbl_h[m] = (uint8_t)(int)(L * pow(R1, 0.5));
	}
}

} // namespace

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
extern "C" {

int ert_abi_version(void) { return 2; }   // 2: ert_result grew (plane_order_sensitive), round-2 entry points
const char *ert_last_error(void) { return g_err; }

const char *ert_status_string(uint32_t s)
{
	static thread_local char buf[256];
	buf[0] = 0;
	if (!s) return "ok";
	if (s & ERR_LOOP_GUARD) strcat(buf, "loop-guard ");
	if (s & ERR_KEPT_OVERFLOW) strcat(buf, "kept-overflow ");
	if (s & ERR_POOL_OVERFLOW) strcat(buf, "pool-overflow ");
	if (s & ERR_NMS_OVERFLOW) strcat(buf, "nms-overflow ");
	if (s & ERR_NODE_OVERFLOW) strcat(buf, "node-overflow ");
	return buf;
}

ert_ctx *ert_create(const ert_params *params, int device)
{
	if (check_cuda_device(device)) return nullptr;
	if (cudaSetDevice(device) != cudaSuccess) { set_error("cudaSetDevice(%d) failed", device); return nullptr; }
	ert_ctx *c = new ert_ctx();
	if (params) c->prm = *params;
	else { c->prm.thresh_step = 8; c->prm.min_area = 120; c->prm.max_area = 900000; c->prm.stability_t = 2; c->prm.overlap_coef = 0.7; c->prm.min_ocr_prob = 0.15; }
	if (c->prm.thresh_step < 5 || c->prm.thresh_step > 255) { set_error("thresh_step %d unsupported (5..255)", c->prm.thresh_step); delete c; return nullptr; }
	c->device = device;
	if (cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || c->sm_count < 1) c->sm_count = 148;
	if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { set_error("stream create failed"); delete c; return nullptr; }
	for (int i = 0; i < 12; i++) cudaEventCreate(&c->ev[i]);
	{
		int least = 0, greatest = 0;
		cudaDeviceGetStreamPriorityRange(&least, &greatest);
		if (cudaStreamCreateWithPriority(&c->post_stream, cudaStreamNonBlocking, greatest) != cudaSuccess) c->post_stream = nullptr;
		cudaEventCreateWithFlags(&c->ev_post_done, cudaEventDisableTiming);
		cudaEventCreateWithFlags(&c->ev_planes, cudaEventDisableTiming);
		cudaEventCreateWithFlags(&c->ev_resized, cudaEventDisableTiming);
		cudaEventCreate(&c->ev_decode[0]); cudaEventCreate(&c->ev_decode[1]);
	}
	build_aran_table(c);
	if (cudaMalloc((void **)&c->d_aran_tbl, 64) != cudaSuccess || cudaMemcpy(c->d_aran_tbl, c->aran_tbl_h, 64, cudaMemcpyHostToDevice) != cudaSuccess) {
		set_error("aran table upload failed"); delete c; return nullptr;
	}
	return c;
}

void ert_destroy(ert_ctx *c)
{
	if (!c) return;
	cudaSetDevice(c->device);
	cudaStreamSynchronize(c->stream);
	{
		std::lock_guard<std::mutex> lk(g_fifo.mu);
		if (g_fifo.owner[c->device & 63] == c) { g_fifo.owner[c->device & 63] = nullptr; g_fifo.last[c->device & 63] = nullptr; }
	}
	free_workspace(c);
	cudaFree(c->d_bgr); cudaFree(c->d_aran_tbl);
	for (int k = 0; k < 2; k++) { cudaFree(c->casc[k].d_stumps); cudaFree(c->casc[k].d_len); cudaFree(c->casc[k].d_thr); cudaFree(c->casc[k].d_cpcn); cudaFree(c->casc[k].d_dimthr); }
	cudaFree(c->csc.stage_sum); cudaFree(c->csc.done); cudaFree(c->csc.pool_prefix);
	cudaFree(c->svm.d_sv); cudaFree(c->svm.d_coef); cudaFree(c->svm.d_coefT); cudaFree(c->svm.d_pair_ij); cudaFree(c->svm.d_rho); cudaFree(c->svm.d_probA); cudaFree(c->svm.d_probB);
	cudaFree(c->svm.d_label); cudaFree(c->svm.d_nsv); cudaFree(c->svm.d_start);
	cudaFree(c->svm.d_svj); cudaFree(c->svm.d_sve); cudaFree(c->svm.d_ss);
	c->s0.release(); c->s1.release(); c->s2.release(); c->s3.release(); c->s4.release();
	free_next(c);
	for (int i = 0; i < 12; i++) cudaEventDestroy(c->ev[i]);
	if (c->post_stream) { cudaStreamSynchronize(c->post_stream); cudaStreamDestroy(c->post_stream); }
	if (c->ev_post_done) cudaEventDestroy(c->ev_post_done);
	if (c->ev_planes) cudaEventDestroy(c->ev_planes);
	if (c->ev_resized) cudaEventDestroy(c->ev_resized);
	for (cudaEvent_t e : c->ev_decode) if (e) cudaEventDestroy(e);
	if (c->jpeg) { jpeg_decoder_destroy(c->jpeg); c->jpeg = nullptr; }
	if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
	delete c;
}

int ert_set_thresh_step(ert_ctx *c, int step)
{
	if (step < 5 || step > 255) { set_error("thresh_step %d unsupported (5..255)", step); return -1; }
	c->prm.thresh_step = step; return 0;
}
int ert_set_min_area(ert_ctx *c, int m) { c->prm.min_area = m; return 0; }
int ert_set_params(ert_ctx *c, const ert_params *p)
{
	if (!c || !p) { set_error("bad arguments"); return -1; }
	if (p->thresh_step < 5 || p->thresh_step > 255) { set_error("thresh_step %d unsupported (5..255)", p->thresh_step); return -1; }
	c->prm = *p;     // read at enqueue time: applies to the next batch
	return 0;
}
int ert_set_return_hist(ert_ctx *c, int on)
{
	if (on && !c->return_hist && c->planes_cap) { cudaStreamSynchronize(c->stream); free_workspace(c); }   // re-allocate with the host histogram buffer
	c->return_hist = on;
	return 0;
}
int ert_set_nms_sequential(ert_ctx *c, int on) { c->nms_sequential = on ? 1 : 0; return 0; }
int ert_set_tile_fifo(ert_ctx *c, int on) { c->tile_fifo = on ? 1 : 0; return 0; }
int ert_set_stream_split(ert_ctx *c, int on) { cudaStreamSynchronize(c->stream); if (c->post_stream) cudaStreamSynchronize(c->post_stream); c->split_streams = on ? 1 : 0; return 0; }
int ert_debug_phase_cycles(ert_ctx *c, int enable, unsigned long long *out16)
{
	if (enable && !c->d_prof) { ERT_CUDA_CHECK(cudaMalloc((void **)&c->d_prof, 16 * sizeof(unsigned long long))); ERT_CUDA_CHECK(cudaMemset(c->d_prof, 0, 16 * sizeof(unsigned long long))); }
	if (out16 && c->d_prof) { ERT_CUDA_CHECK(cudaStreamSynchronize(c->stream)); ERT_CUDA_CHECK(cudaMemcpy(out16, c->d_prof, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost)); ERT_CUDA_CHECK(cudaMemset(c->d_prof, 0, 16 * sizeof(unsigned long long))); }
	if (!enable && c->d_prof) { cudaFree(c->d_prof); c->d_prof = nullptr; }
	c->wk.prof = c->d_prof;
	return 0;
}
int ert_set_post_footprint(ert_ctx *c, int ctas_per_sm)
{
	if (!c || ctas_per_sm < 0 || ctas_per_sm > 32) { set_error("bad arguments"); return -1; }
	c->post_ctas_per_sm = ctas_per_sm; c->wk.post_ctas = ctas_per_sm * c->sm_count;
	return 0;
}

int ert_set_seam_list(ert_ctx *c, int on) { c->seam_list = on ? 1 : 0; c->wk.seam_list = c->seam_list; return 0; }
int ert_set_tile_config(ert_ctx *c, int id)
{
	if (id < 0 || id >= tile_config_count()) { set_error("tile config %d out of range", id); return -1; }
	c->tile_cfg = id; c->wk.tile_cfg = id; return 0;
}
int ert_set_capacity(ert_ctx *c, int kept, int pool)
{
	// pool: k_track keeps one bit per candidate of a frame (6 planes x pool) in 48 KB of shared memory
	if (kept < 16 || pool < 16 || kept > (1 << 22) || pool > 32768) { set_error("bad capacity (kept 16..4194304, pool 16..32768)"); return -1; }
	cudaStreamSynchronize(c->stream);
	free_workspace(c);
	c->kept_cap = kept; c->pool_cap = pool;
	return 0;
}

int ert_set_node_capacity(ert_ctx *c, int slots_per_plane)
{
	if (slots_per_plane < 0 || slots_per_plane > (1 << KEY_IDX_BITS)) { set_error("bad node capacity"); return -1; }
	cudaStreamSynchronize(c->stream);
	if (c->post_stream) cudaStreamSynchronize(c->post_stream);
	free_workspace(c);
	c->node_cap_user = slots_per_plane;
	return 0;
}

int ert_set_stream(ert_ctx *c, uint64_t s)
{
	cudaStreamSynchronize(c->stream);
	if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
	if (s == 0) { c->own_stream = true; if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { set_error("stream create failed"); return -1; } }
	else { c->own_stream = false; c->stream = (cudaStream_t)(uintptr_t)s; }
	return 0;
}
uint64_t ert_get_stream(ert_ctx *c) { return (uint64_t)(uintptr_t)c->stream; }
int ert_last_launch_count(ert_ctx *c) { return c->launches; }

// ---- model files --------------------------------------------------------------------------------
int ert_load_cascade(ert_ctx *c, int which, const char *path)
{
	if (which < 0 || which > 1) { set_error("bad cascade id"); return -1; }
	std::ifstream fin(path);
	if (!fin.is_open()) { set_error("Error: %s is not opened!!", path); return -1; }
	CascadeHost &h = c->casc[which];
	h.stage_len.clear(); h.stage_thr.clear(); h.stumps.clear(); h.loaded = false;
	std::string tok;
	bool real = true;
	fin >> tok;
	if (tok == "boost_type") { fin >> tok; real = (tok != "DISCRETE"); fin >> tok; }
	if (tok == "base_type") { fin >> tok; fin >> tok; }
	if (tok == "num_of_iter") {
		while (fin >> tok) {
			char *e = nullptr;
			const double v = strtod(tok.c_str(), &e);
			if (e == tok.c_str()) break;     // first non-numeric token ends the list ("threshold")
			h.stage_len.push_back((int)v);
		}
	}
	if (tok == "threshold") {
		for (size_t j = 0; j < h.stage_len.size(); j++) { fin >> tok; h.stage_thr.push_back((int)strtod(tok.c_str(), nullptr)); }
	}
	if (!real) { set_error("%s: only REAL cascades of decision stumps are supported on the device path", path); return -1; }
	std::string line;
	std::getline(fin, line);
	while (std::getline(fin, line)) {
		const char *s = line.c_str();
		char *e;
		double v[5];
		int k = 0;
		while (k < 5) { v[k] = strtod(s, &e); if (e == s) break; s = e; k++; }
		if (k < 5) continue;
		Stump st; st.dim = (int)v[1]; st.thr = v[2]; st.cp = v[3]; st.cn = v[4]; st.pad = 0;
		h.stumps.push_back(st);
	}
	int total = 0;
	for (int n : h.stage_len) total += n;
	if (h.stage_len.empty() || h.stage_thr.size() != h.stage_len.size() || total != (int)h.stumps.size()) {
		set_error("%s: malformed cascade (stages %zu, thresholds %zu, stumps %zu, expected %d)", path, h.stage_len.size(), h.stage_thr.size(), h.stumps.size(), total);
		return -1;
	}
	for (const Stump &st : h.stumps) if (st.dim < 0 || st.dim >= 1024) { set_error("%s: stump dimension %d outside the 1024-bin feature", path, st.dim); return -1; }
	ERT_CUDA_CHECK(cudaSetDevice(c->device));
	ERT_CUDA_CHECK(cudaStreamSynchronize(c->stream));   // a batch in flight may still read the old tables
	if (c->post_stream) ERT_CUDA_CHECK(cudaStreamSynchronize(c->post_stream));
	if (upload_cascade(h)) return -1;
	h.loaded = true;
	return (int)h.stumps.size();
}

int ert_set_cascade(ert_ctx *c, int which, int n_stages, const int *stage_len, const int *stage_thr, int n_stumps, const int *dim, const double *thr,
                    const double *cp, const double *cn)
{
	if (!c || which < 0 || which > 1 || n_stages < 1 || n_stumps < 1 || !stage_len || !stage_thr || !dim || !thr || !cp || !cn) { set_error("bad arguments"); return -1; }
	long long total = 0;
	for (int i = 0; i < n_stages; i++) { if (stage_len[i] < 0) { set_error("negative stage length"); return -1; } total += stage_len[i]; }
	if (total != n_stumps) { set_error("stage lengths sum to %lld, %d stumps given", total, n_stumps); return -1; }
	for (int j = 0; j < n_stumps; j++) if (dim[j] < 0 || dim[j] >= 1024) { set_error("stump dimension %d outside the 1024-bin feature", dim[j]); return -1; }
	CascadeHost &h = c->casc[which];
	h.loaded = false;
	h.stage_len.assign(stage_len, stage_len + n_stages);
	h.stage_thr.assign(stage_thr, stage_thr + n_stages);
	h.stumps.resize((size_t)n_stumps);
	for (int j = 0; j < n_stumps; j++) { Stump st; st.dim = dim[j]; st.thr = thr[j]; st.cp = cp[j]; st.cn = cn[j]; st.pad = 0; h.stumps[(size_t)j] = st; }
	ERT_CUDA_CHECK(cudaSetDevice(c->device));
	ERT_CUDA_CHECK(cudaStreamSynchronize(c->stream));   // a batch in flight may still read the old tables
	if (c->post_stream) ERT_CUDA_CHECK(cudaStreamSynchronize(c->post_stream));
	if (upload_cascade(h)) return -1;
	h.loaded = true;
	return n_stumps;
}

int ert_load_svm(ert_ctx *c, const char *path)
{
	FILE *f = fopen(path, "rb");
	if (!f) { set_error("can't open model file %s", path); return -1; }
	fseek(f, 0, SEEK_END); const long sz = ftell(f); fseek(f, 0, SEEK_SET);
	std::vector<char> txt((size_t)sz + 1);
	if (fread(txt.data(), 1, (size_t)sz, f) != (size_t)sz) { fclose(f); set_error("short read on %s", path); return -1; }
	fclose(f); txt[sz] = 0;
	SvmHost &m = c->svm;
	{
		// a reload replaces the model: release the previous device tables, keep the caller's tensor-core choice
		const int keep_tc = m.use_tc;
		cudaFree(m.d_sv); cudaFree(m.d_coef); cudaFree(m.d_coefT); cudaFree(m.d_pair_ij); m.d_pair_ij = nullptr; cudaFree(m.d_rho); cudaFree(m.d_probA); cudaFree(m.d_probB);
		cudaFree(m.d_label); cudaFree(m.d_nsv); cudaFree(m.d_start); cudaFree(m.d_svj); cudaFree(m.d_sve); cudaFree(m.d_ss);
		m = SvmHost();
		m.use_tc = keep_tc;
	}
	char *p = txt.data();
	auto next_line = [&](char *&line) { line = p; char *e = strchr(p, '\n'); if (!e) { p += strlen(p); return; } *e = 0; p = e + 1; };
	std::string svm_type, kernel_type;
	for (;;) {
		char *line; next_line(line);
		char *save = nullptr;
		char *key = strtok_r(line, " \t\r", &save);
		if (!key) { if (!*p) break; continue; }
		if (!strcmp(key, "SV")) break;
		auto tok = [&]() { return strtok_r(nullptr, " \t\r", &save); };
		if (!strcmp(key, "svm_type")) { char *t = tok(); svm_type = t ? t : ""; }
		else if (!strcmp(key, "kernel_type")) { char *t = tok(); kernel_type = t ? t : ""; }
		else if (!strcmp(key, "gamma")) { char *t = tok(); m.gamma = t ? strtod(t, nullptr) : 0; }
		else if (!strcmp(key, "nr_class")) { char *t = tok(); m.nr_class = t ? atoi(t) : 0; }
		else if (!strcmp(key, "total_sv")) { char *t = tok(); m.l = t ? atoi(t) : 0; }
		else if (!strcmp(key, "rho") || !strcmp(key, "probA") || !strcmp(key, "probB")) {
			std::vector<double> &a = (key[0] == 'r') ? m.rho : (key[4] == 'A' ? m.probA : m.probB);
			const int n = m.nr_class * (m.nr_class - 1) / 2;
			a.assign((size_t)n, 0.0);
			for (int i = 0; i < n; i++) { char *t = tok(); if (t) a[i] = strtod(t, nullptr); }
		} else if (!strcmp(key, "label") || !strcmp(key, "nr_sv")) {
			std::vector<int> &a = (key[0] == 'l') ? m.label : m.nsv;
			a.assign((size_t)m.nr_class, 0);
			for (int i = 0; i < m.nr_class; i++) { char *t = tok(); if (t) a[i] = atoi(t); }
		}
	}
	if (svm_type != "c_svc" || kernel_type != "rbf" || m.probA.empty() || m.probB.empty() || m.nr_class < 2 || m.l < 1) {
		set_error("%s: only c_svc + rbf models with probability estimates are supported (svm_type=%s kernel_type=%s)", path, svm_type.c_str(), kernel_type.c_str());
		return -1;
	}
	{
		const size_t npair = (size_t)m.nr_class * (m.nr_class - 1) / 2;
		long long nsv_sum = 0;
		for (int v : m.nsv) nsv_sum += v;
		if (m.rho.size() != npair || m.probA.size() != npair || m.probB.size() != npair || m.label.size() != (size_t)m.nr_class ||
		    m.nsv.size() != (size_t)m.nr_class || nsv_sum != m.l || m.nr_class > 4096 || m.l > (1 << 22)) {
			set_error("%s: malformed model header (nr_class %d, total_sv %d, rho %zu, label %zu, nr_sv %zu summing to %lld)", path, m.nr_class, m.l,
			          m.rho.size(), m.label.size(), m.nsv.size(), nsv_sum);
			m = SvmHost();
			return -1;
		}
	}
	const int k1 = m.nr_class - 1;
	m.coef.assign((size_t)k1 * m.l, 0.0);
	std::vector<int> idx; std::vector<double> val; std::vector<int> rowstart((size_t)m.l + 1, 0);
	int maxidx = -1;
	for (int i = 0; i < m.l; i++) {
		char *line; next_line(line);
		char *s = line, *e;
		for (int j = 0; j < k1; j++) { m.coef[(size_t)j * m.l + i] = strtod(s, &e); s = e; }
		rowstart[i] = (int)idx.size();
		for (;;) {
			while (*s == ' ' || *s == '\t' || *s == '\r') s++;
			if (!*s) break;
			const long id = strtol(s, &e, 10);
			if (e == s || *e != ':') break;
			s = e + 1;
			const double v = strtod(s, &e);
			s = e;
			if (id < 0 || id >= (1 << 20)) { set_error("%s: support vector %d has feature index %ld outside 0..%d", path, i, id, (1 << 20) - 1); m = SvmHost(); return -1; }
			idx.push_back((int)id); val.push_back(v);
			if (id > maxidx) maxidx = (int)id;
		}
	}
	rowstart[m.l] = (int)idx.size();
	m.dims = maxidx + 1;
	if (m.dims < 1) { set_error("%s: no support vector entries", path); m = SvmHost(); return -1; }
	if ((size_t)m.l * (size_t)m.dims > ((size_t)1 << 31)) { set_error("%s: %d support vectors x %d dimensions is more than this loader densifies", path, m.l, m.dims); m = SvmHost(); return -1; }
	m.sv.assign((size_t)m.l * m.dims, 0.0);
	for (int i = 0; i < m.l; i++)
		for (int q = rowstart[i]; q < rowstart[i + 1]; q++) if (idx[q] >= 0) m.sv[(size_t)i * m.dims + idx[q]] = val[q];
	m.start.assign((size_t)m.nr_class, 0);
	for (int i = 1; i < m.nr_class; i++) m.start[i] = m.start[i - 1] + m.nsv[i - 1];
	ERT_CUDA_CHECK(cudaSetDevice(c->device));
	// tensor-core tables: v = j/255 + eps with j = round(255 v); e = round(S * eps), S = 127 / max|eps|
	if (m.dims <= svm_tc_kpad() && m.l <= svm_tc_npad()) {
		const int KP = svm_tc_kpad(), NP = svm_tc_npad();
		m.svj.assign((size_t)NP * KP, 0); m.sve.assign((size_t)NP * KP, 0); m.ss.assign((size_t)NP, 0.0);
		double maxeps = 0;
		bool ok = true;
		for (int i = 0; i < m.l && ok; i++)
			for (int d = 0; d < m.dims; d++) {
				const double v = m.sv[(size_t)i * m.dims + d];
				if (v < 0 || v > 1.0 + 1e-9) { ok = false; break; }
				const double j = floor(v * 255.0 + 0.5);
				maxeps = std::max(maxeps, fabs(v - j / 255.0));
			}
		if (ok) {
			const double S = (maxeps > 0) ? 127.0 / maxeps : 1.0;
			for (int i = 0; i < m.l; i++) {
				double acc = 0;
				for (int d = 0; d < m.dims; d++) {
					const double v = m.sv[(size_t)i * m.dims + d];
					const double j = floor(v * 255.0 + 0.5);
					m.svj[(size_t)i * KP + d] = (uint8_t)j;
					m.sve[(size_t)i * KP + d] = (int8_t)lrint((v - j / 255.0) * S);
					acc += v * v;
				}
				m.ss[i] = acc;
			}
			m.inv_s255 = 1.0 / (255.0 * S);
			if (dev_upload(&m.d_svj, m.svj) || dev_upload(&m.d_sve, m.sve) || dev_upload(&m.d_ss, m.ss)) return -1;
		}
	}
	ERT_CUDA_CHECK(cudaSetDevice(c->device));
	m.coefT.assign((size_t)m.l * k1, 0.0);
	for (int j = 0; j < k1; j++) for (int i = 0; i < m.l; i++) m.coefT[(size_t)i * k1 + j] = m.coef[(size_t)j * m.l + i];
	if (dev_upload(&m.d_coefT, m.coefT)) return -1;
	m.pair_ij.clear();
	for (int i = 0; i < m.nr_class; i++) for (int j = i + 1; j < m.nr_class; j++) m.pair_ij.push_back((uint16_t)(i << 8 | j));
	if (dev_upload(&m.d_pair_ij, m.pair_ij)) return -1;
	if (dev_upload(&m.d_sv, m.sv) || dev_upload(&m.d_coef, m.coef) || dev_upload(&m.d_rho, m.rho) || dev_upload(&m.d_probA, m.probA) ||
	    dev_upload(&m.d_probB, m.probB) || dev_upload(&m.d_label, m.label) || dev_upload(&m.d_nsv, m.nsv) || dev_upload(&m.d_start, m.start)) return -1;
	m.loaded = true;
	return m.l;
}

int ert_cascade_stage_info(ert_ctx *c, int which, int *stage_len, int *stage_thr, int cap)
{
	if (!c || which < 0 || which > 1 || !c->casc[which].loaded) { set_error("cascade %d is not loaded", which); return -1; }
	const CascadeHost &h = c->casc[which];
	const int n = (int)h.stage_len.size();
	for (int i = 0; i < n && i < cap; i++) { if (stage_len) stage_len[i] = h.stage_len[i]; if (stage_thr) stage_thr[i] = h.stage_thr[i]; }
	return n;
}

int ert_svm_nr_class(ert_ctx *c) { return c->svm.loaded ? c->svm.nr_class : -1; }
int ert_svm_total_sv(ert_ctx *c) { return c->svm.loaded ? c->svm.l : -1; }
int ert_svm_labels(ert_ctx *c, int *label)
{
	if (!c || !c->svm.loaded || !label) { set_error("svm model is not loaded"); return -1; }
	for (int i = 0; i < c->svm.nr_class; i++) label[i] = c->svm.label[i];
	return c->svm.nr_class;
}
double ert_svm_gamma(ert_ctx *c) { return c->svm.loaded ? c->svm.gamma : 0.0; }
int ert_set_svm_tensor_cores(ert_ctx *c, int on) { if (!c || on < 0 || on > 2) { set_error("bad arguments"); return -1; } c->svm.use_tc = on; return 0; }
int ert_set_svm_legacy_prob(ert_ctx *c, int on) { c->svm.legacy_prob = on ? 1 : 0; return 0; }
int ert_svm_dims(ert_ctx *c) { return c->svm.loaded ? c->svm.dims : -1; }

// contexts that build pyramid levels from this context's planes must have read them before the next batch overwrites them
static int wait_for_readers(ert_ctx *c)
{
	for (cudaEvent_t e : c->readers) ERT_CUDA_CHECK(cudaStreamWaitEvent(c->stream, e, 0));
	c->readers.clear();
	return 0;
}

// ---- the batched hot path ----------------------------------------------------------------------
static int detect_common(ert_ctx *c, const uint8_t *bgr, bool on_device, int n_frames, int W, int H, int stride, int upto)
{
	if (!c || !bgr || n_frames < 1) { set_error("bad arguments"); return -1; }
	if (stride < 3 * W) { set_error("stride %d < 3*width", stride); return -1; }
	ERT_CUDA_CHECK(cudaSetDevice(c->device));
	const int n_planes = c->planes_per_frame * n_frames;
	if (upto >= ERT_STAGE_TRACK && c->planes_per_frame != 6) { set_error("ERT_STAGE_TRACK needs all six channels (ert_set_planes_per_frame)"); return -1; }
	if (wait_for_readers(c)) return -1;       // pyramid levels still reading this context's planes / plane table
	if (ensure_workspace(c, n_planes, W, H)) return -1;
	// the plane table (plane -> source pointer, invert flag) depends on the layout, the plane pitch and the plane count
	if (c->frames_cap != -1 || c->table_W != W || c->table_H != H || n_planes > c->table_planes || c->table_ppf != c->planes_per_frame) {
		if (set_plane_table(c, n_planes, true)) return -1;
		c->frames_cap = -1; c->table_W = W; c->table_H = H; c->table_planes = n_planes; c->table_ppf = c->planes_per_frame;
	}
	c->launches = 0;
	cudaStream_t st = c->stream;
	ERT_CUDA_CHECK(cudaEventRecord(c->ev[0], st));
	const uint8_t *d_in = bgr;
	const size_t frame_bytes = (size_t)stride * H;
	if (!on_device) {
		if (c->bgr_cap < frame_bytes * n_frames) {
			cudaFree(c->d_bgr); c->d_bgr = nullptr; c->bgr_cap = 0;
			ERT_CUDA_CHECK(cudaMalloc((void **)&c->d_bgr, frame_bytes * n_frames));
			c->bgr_cap = frame_bytes * n_frames;
		}
		ERT_CUDA_CHECK(cudaMemcpyAsync(c->d_bgr, bgr, frame_bytes * n_frames, cudaMemcpyHostToDevice, st));
		d_in = c->d_bgr;
	}
	ERT_CUDA_CHECK(cudaEventRecord(c->ev[1], st));
	if (launch_channels(d_in, frame_bytes, stride, W, H, n_frames, c->d_ycc, c->pitch, st)) return -1;
	c->launches += 1;
	ERT_CUDA_CHECK(cudaEventRecord(c->ev_planes, st));
	return enqueue_pipeline(c, n_planes, upto);
}

int ert_detect_classify(ert_ctx *c, const uint8_t *bgr, int n_frames, int W, int H, int stride, int upto, const ert_result **out)
{
	if (detect_common(c, bgr, false, n_frames, W, H, stride, upto)) return -1;
	return finish_result(c, out);
}

int ert_enqueue_host(ert_ctx *c, const uint8_t *bgr, int n_frames, int W, int H, int stride, int upto)
{
	return detect_common(c, bgr, false, n_frames, W, H, stride, upto);
}

int ert_detect_classify_device(ert_ctx *c, const void *d_bgr, int n_frames, int W, int H, int stride, int upto)
{
	return detect_common(c, (const uint8_t *)d_bgr, true, n_frames, W, H, stride, upto);
}

int ert_fetch_result(ert_ctx *c, const ert_result **out) { return finish_result(c, out); }

int ert_batch_done(ert_ctx *c)
{
	if (!c || !c->pending) return -1;
	const cudaError_t e = cudaEventQuery(c->track_pending ? c->ev[11] : c->ev[5]);
	if (e == cudaSuccess) return 1;
	if (e == cudaErrorNotReady) return 0;
	set_error("cudaEventQuery: %s", cudaGetErrorString(e));
	return -1;
}

int ert_compute_channels(ert_ctx *c, const uint8_t *bgr, int W, int H, int stride, uint8_t *planes6)
{
	if (!c || !bgr || !planes6 || W < 1 || H < 1 || stride < 3 * W) { set_error("bad arguments"); return -1; }
	ERT_CUDA_CHECK(cudaSetDevice(c->device));
	cudaStream_t st = c->stream;
	const int pitch = extract_pitch(W);
	const size_t in_b = (size_t)stride * H, ycc_b = (size_t)pitch * H * 3, out_b = (size_t)W * H * 6;
	if (c->s0.ensure(in_b) || c->s1.ensure(ycc_b) || c->s2.ensure(out_b)) return -1;
	ERT_CUDA_CHECK(cudaMemcpyAsync(c->s0.p, bgr, in_b, cudaMemcpyHostToDevice, st));
	if (launch_channels((const uint8_t *)c->s0.p, in_b, stride, W, H, 1, (uint8_t *)c->s1.p, pitch, st)) return -1;
	if (launch_unpack_planes((const uint8_t *)c->s1.p, pitch, W, H, (uint8_t *)c->s2.p, st)) return -1;
	ERT_CUDA_CHECK(cudaMemcpyAsync(planes6, c->s2.p, out_b, cudaMemcpyDeviceToHost, st));
	ERT_CUDA_CHECK(cudaStreamSynchronize(st));
	return 0;
}

int ert_enqueue_planes(ert_ctx *c, const uint8_t *planes, int n_planes, int W, int H, int stride, size_t plane_stride, int upto)
{
	if (!c || !planes || n_planes < 1 || stride < W) { set_error("bad arguments"); return -1; }
	if (upto >= ERT_STAGE_TRACK) { set_error("ERT_STAGE_TRACK needs a BGR batch (er_track reads the YCrCb frame)"); return -1; }
	ERT_CUDA_CHECK(cudaSetDevice(c->device));
	if (wait_for_readers(c)) return -1;
	if (ensure_workspace(c, n_planes, W, H)) return -1;
	if (set_plane_table(c, n_planes, false)) return -1;
	c->frames_cap = 0; c->table_planes = 0;   // plane table no longer in BGR layout
	c->launches = 0;
	cudaStream_t st = c->stream;
	ERT_CUDA_CHECK(cudaEventRecord(c->ev[0], st));
	for (int p = 0; p < n_planes; p++)
		ERT_CUDA_CHECK(cudaMemcpy2DAsync(c->d_ycc + (size_t)p * c->ycc_bytes, (size_t)c->pitch, planes + (size_t)p * plane_stride, (size_t)stride,
		                                 (size_t)W, (size_t)H, cudaMemcpyHostToDevice, st));
	ERT_CUDA_CHECK(cudaEventRecord(c->ev[1], st));
	ERT_CUDA_CHECK(cudaEventRecord(c->ev_planes, st));
	return enqueue_pipeline(c, n_planes, upto);
}

// A pyramid level on the device: every plane of the batch `src` has in flight (for a BGR batch the channels of
// compute_channels, inverted ones inverted first), resized by cv::resize(INTER_LINEAR) semantics to (width / div,
// height / div) straight into `dst`'s plane buffer, then the same per-plane path on `dst`.  Nothing but the frame itself
// ever crosses the bus.
int ert_enqueue_pyramid_level(ert_ctx *dst, ert_ctx *src, int div, int upto)
{
	if (!dst || !src || dst == src || div < 2) { set_error("bad arguments"); return -1; }
	if (dst->device != src->device) { set_error("ert_enqueue_pyramid_level: both contexts must live on one device"); return -1; }
	if (!src->pending || src->pending_planes < 1) { set_error("ert_enqueue_pyramid_level: the source context has no batch in flight"); return -1; }
	if (upto >= ERT_STAGE_TRACK) { set_error("ert_enqueue_pyramid_level: er_track runs on the native-scale batch only"); return -1; }
	const int dw = src->W / div, dh = src->H / div;
	if (dw < 1 || dh < 1) { set_error("ert_enqueue_pyramid_level: %dx%d / %d is empty", src->W, src->H, div); return -1; }
	ERT_CUDA_CHECK(cudaSetDevice(dst->device));
	const int n_planes = src->pending_planes;
	if (wait_for_readers(dst)) return -1;
	if (ensure_workspace(dst, n_planes, dw, dh)) return -1;
	if (set_plane_table(dst, n_planes, false)) return -1;
	dst->frames_cap = 0; dst->table_planes = 0;        // the level holds its planes one by one (no BGR layout)
	dst->launches = 0;
	cudaStream_t st = dst->stream;
	ERT_CUDA_CHECK(cudaEventRecord(dst->ev[0], st));
	ERT_CUDA_CHECK(cudaStreamWaitEvent(st, src->ev_planes, 0));
	ERT_CUDA_CHECK(cudaEventRecord(dst->ev[1], st));
	if (launch_resize_planes(src->d_planes, n_planes, src->W, src->H, src->pitch, dst->d_ycc, dw, dh, dst->pitch, dst->ycc_bytes, st)) return -1;
	dst->launches += 1;
	ERT_CUDA_CHECK(cudaEventRecord(dst->ev_resized, st));
	ERT_CUDA_CHECK(cudaEventRecord(dst->ev_planes, st));
	src->readers.push_back(dst->ev_resized);
	return enqueue_pipeline(dst, n_planes, upto);
}

int ert_set_planes_per_frame(ert_ctx *c, int n)
{
	if (n != 3 && n != 6) { set_error("planes per frame: 6 (Y, Cr, Cb and their inverses) or 3 (Y, Cr, Cb)"); return -1; }
	c->planes_per_frame = n;
	return 0;
}

int ert_planes_detect(ert_ctx *c, const uint8_t *planes, int n_planes, int W, int H, int stride, size_t plane_stride, int upto, const ert_result **out)
{
	if (ert_enqueue_planes(c, planes, n_planes, W, H, stride, plane_stride, upto)) return -1;
	return finish_result(c, out);
}

// ---- stage entry points on caller data --------------------------------------------------------
int ert_nms_nodes(ert_ctx *c, const ert_node *nodes, int n, int W, int H, int32_t *pool_out, int pool_cap, int *n_pool)
{
	if (!c || !nodes || n < 1) { set_error("bad arguments"); return -1; }
	if (n > c->kept_cap) { set_error("ert_nms_nodes: %d nodes exceed the kept capacity %d (ert_set_capacity)", n, c->kept_cap); return -1; }
	ERT_CUDA_CHECK(cudaSetDevice(c->device));
	cudaStream_t st = c->stream;
	const size_t stride = nms_scratch_stride(c->kept_cap);
	if (c->s0.ensure(sizeof(OutNode) * (size_t)n) || c->s1.ensure(stride) || c->s2.ensure(sizeof(OutNode) * (size_t)c->kept_cap) ||
	    c->s3.ensure(sizeof(int32_t) * (size_t)(c->pool_cap + 8)) || c->s4.ensure(64)) return -1;
	int32_t offs[2] = {0, n};
	int32_t *d_offs = (int32_t *)c->s4.p;             // [0..1] offsets, [2..3] counts, [4] status
	int32_t *d_counts = d_offs + 2;
	uint32_t *d_status = (uint32_t *)(d_offs + 4);
	ERT_CUDA_CHECK(cudaMemsetAsync(c->s4.p, 0, 64, st));
	ERT_CUDA_CHECK(cudaMemcpyAsync(d_offs, offs, sizeof offs, cudaMemcpyHostToDevice, st));
	ERT_CUDA_CHECK(cudaMemcpyAsync(c->s0.p, nodes, sizeof(OutNode) * (size_t)n, cudaMemcpyHostToDevice, st));
	NmsParams NP = make_nms_params(c, W, H);
	if (launch_nms(NP, 1, nullptr, nullptr, nullptr, nullptr, nullptr, (const OutNode *)c->s0.p, d_offs, (uint8_t *)c->s1.p, stride,
	               (OutNode *)c->s2.p, (int32_t *)c->s3.p, d_counts, d_status, st)) return -1;
	int32_t counts[3] = {0, 0, 0};
	ERT_CUDA_CHECK(cudaMemcpyAsync(counts, d_counts, sizeof(int32_t) * 3, cudaMemcpyDeviceToHost, st));
	ERT_CUDA_CHECK(cudaStreamSynchronize(st));
	const int np = counts[1];
	if (np > pool_cap) { set_error("pool_out too small: need %d", np); return -1; }
	// pool indices come back in the kernel's DFS numbering, which equals the caller's numbering
	// because the caller's order was used verbatim
	ERT_CUDA_CHECK(cudaMemcpy(pool_out, c->s3.p, sizeof(int32_t) * (size_t)np, cudaMemcpyDeviceToHost));
	if (n_pool) *n_pool = np;
	return 0;
}

static int classify_common(ert_ctx *c, const uint8_t *plane, int W, int H, int stride, const int32_t *rects, int n, int32_t *label,
                           double *ss, double *ws, uint8_t *hist_u8, double *hist_f64, bool need_cascade, uint8_t *codes = nullptr)
{
	if (!c || !plane || !rects || n < 0 || stride < W) { set_error("bad arguments"); return -1; }
	if (n == 0) return 0;
	if (need_cascade && (!c->casc[0].loaded || !c->casc[1].loaded)) { set_error("cascades are not loaded"); return -1; }
	for (int i = 0; i < n; i++) {
		const int32_t *r = rects + 4 * i;
		if (r[0] < 0 || r[1] < 0 || r[2] < 1 || r[3] < 1 || r[0] + r[2] > W || r[1] + r[3] > H) { set_error("rect %d outside the plane", i); return -1; }
	}
	ERT_CUDA_CHECK(cudaSetDevice(c->device));
	cudaStream_t st = c->stream;
	const int pitch = extract_pitch(W);
	const size_t nodes_b = sizeof(OutNode) * (size_t)n, pool_b = sizeof(int32_t) * (size_t)n;
	// s0: plane, s1: nodes + pool + counts + PlaneSrc, s2: hist, s3: label + scores
	if (c->s0.ensure((size_t)pitch * H) || c->s1.ensure(nodes_b + pool_b + 64 + sizeof(PlaneSrc)) || c->s2.ensure((size_t)n * 1024) ||
	    c->s3.ensure((size_t)n * (sizeof(int32_t) + 2 * sizeof(double)) + 64) || (codes && c->s4.ensure((size_t)n * 576))) return -1;
	ERT_CUDA_CHECK(cudaMemcpy2DAsync(c->s0.p, (size_t)pitch, plane, (size_t)stride, (size_t)W, (size_t)H, cudaMemcpyHostToDevice, st));
	std::vector<OutNode> hn((size_t)n);
	std::vector<int32_t> hp((size_t)n);
	for (int i = 0; i < n; i++) {
		hn[i].level = 0; hn[i].area = 0; hn[i].x = rects[4 * i]; hn[i].y = rects[4 * i + 1]; hn[i].w = rects[4 * i + 2]; hn[i].h = rects[4 * i + 3];
		hn[i].parent = -1; hn[i].nchild = 0; hp[i] = i;
	}
	uint8_t *b1 = (uint8_t *)c->s1.p;
	OutNode *d_nodes = (OutNode *)b1;
	int32_t *d_pool = (int32_t *)(b1 + nodes_b);
	int32_t *d_counts = (int32_t *)(b1 + nodes_b + ((pool_b + 15) / 16) * 16);
	PlaneSrc *d_ps = (PlaneSrc *)((uint8_t *)d_counts + 32);
	int32_t counts[2] = {n, n};
	PlaneSrc ps; ps.src = (const uint8_t *)c->s0.p; ps.invert = 0;
	ERT_CUDA_CHECK(cudaMemcpyAsync(d_nodes, hn.data(), nodes_b, cudaMemcpyHostToDevice, st));
	ERT_CUDA_CHECK(cudaMemcpyAsync(d_pool, hp.data(), pool_b, cudaMemcpyHostToDevice, st));
	ERT_CUDA_CHECK(cudaMemcpyAsync(d_counts, counts, sizeof counts, cudaMemcpyHostToDevice, st));
	ERT_CUDA_CHECK(cudaMemcpyAsync(d_ps, &ps, sizeof ps, cudaMemcpyHostToDevice, st));
	ClassifyParams CP; CP.pitch = pitch; CP.pool_cap = n; CP.node_cap = n;
	if (ensure_cascade_scratch(c, n, 1)) return -1;
	if (launch_pool_prefix(d_counts, 1, n, c->csc.pool_prefix, st)) return -1;
	if (launch_lbp_hist(CP, 1, d_ps, d_nodes, d_pool, c->csc.pool_prefix, c->d_aran_tbl, (uint8_t *)c->s2.p, st, codes ? (uint8_t *)c->s4.p : nullptr)) return -1;
	int32_t *d_label = (int32_t *)c->s3.p;
	double *d_ss = (double *)((uint8_t *)c->s3.p + (((size_t)n * 4 + 63) / 64) * 64);
	double *d_ws = d_ss + n;
	if (need_cascade) {
		if (launch_cascade_u8((const uint8_t *)c->s2.p, 1024, n, nullptr, 0, n, c->casc[0].dev(), c->casc[1].dev(), d_label, d_ss, d_ws, c->csc, st)) return -1;
	}
	ERT_CUDA_CHECK(cudaStreamSynchronize(st));   // staging vectors hn/hp may go out of scope now
	if (label) ERT_CUDA_CHECK(cudaMemcpy(label, d_label, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost));
	if (ss) ERT_CUDA_CHECK(cudaMemcpy(ss, d_ss, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost));
	if (ws) ERT_CUDA_CHECK(cudaMemcpy(ws, d_ws, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost));
	if (hist_u8) ERT_CUDA_CHECK(cudaMemcpy(hist_u8, c->s2.p, (size_t)n * 1024, cudaMemcpyDeviceToHost));
	if (codes) ERT_CUDA_CHECK(cudaMemcpy(codes, c->s4.p, (size_t)n * 576, cudaMemcpyDeviceToHost));
	if (hist_f64) {
		std::vector<uint8_t> tmp((size_t)n * 1024);
		ERT_CUDA_CHECK(cudaMemcpy(tmp.data(), c->s2.p, (size_t)n * 1024, cudaMemcpyDeviceToHost));
		for (size_t i = 0; i < tmp.size(); i++) hist_f64[i] = (double)tmp[i];   // vector<double> of counts, as make_LBP_hist returns
	}
	return 0;
}

int ert_classify_regions(ert_ctx *c, const uint8_t *plane, int W, int H, int stride, const int32_t *rects, int n, int32_t *label,
                         double *ss, double *ws, uint8_t *hist1024)
{
	return classify_common(c, plane, W, H, stride, rects, n, label, ss, ws, hist1024, nullptr, true);
}

int ert_lbp_hist(ert_ctx *c, const uint8_t *plane, int W, int H, int stride, const int32_t *rects, int n, double *hist)
{
	return classify_common(c, plane, W, H, stride, rects, n, nullptr, nullptr, nullptr, nullptr, hist, false);
}

int ert_calc_lbp(ert_ctx *c, const uint8_t *plane, int W, int H, int stride, const int32_t *rects, int n, uint8_t *codes)
{
	if (!codes) { set_error("bad arguments"); return -1; }
	return classify_common(c, plane, W, H, stride, rects, n, nullptr, nullptr, nullptr, nullptr, nullptr, false, codes);
}

int ert_cascade_predict_batch(ert_ctx *c, int which, const double *fv, int n, int dims, double *score)
{
	if (!c || !fv || !score || which < 0 || which > 1 || n < 0) { set_error("bad arguments"); return -1; }
	if (!c->casc[which].loaded) { set_error("cascade %d is not loaded", which); return -1; }
	if (dims < 1024) { set_error("feature vectors must have at least 1024 dims (stump dims index 0..1023)"); return -1; }
	if (n == 0) return 0;
	ERT_CUDA_CHECK(cudaSetDevice(c->device));
	cudaStream_t st = c->stream;
	if (c->s0.ensure(sizeof(double) * (size_t)n * dims) || c->s3.ensure((size_t)n * (sizeof(int32_t) + 2 * sizeof(double)) + 64)) return -1;
	ERT_CUDA_CHECK(cudaMemcpyAsync(c->s0.p, fv, sizeof(double) * (size_t)n * dims, cudaMemcpyHostToDevice, st));
	int32_t *d_label = (int32_t *)c->s3.p;
	double *d_ss = (double *)((uint8_t *)c->s3.p + (((size_t)n * 4 + 63) / 64) * 64);
	double *d_ws = d_ss + n;
	// evaluate the requested cascade in the "strong" slot; the other slot gets the same table (cheap, keeps one kernel)
	const CascadeDev cd = c->casc[which].dev();
	if (launch_cascade_f64((const double *)c->s0.p, (size_t)dims, n, cd, cd, d_label, d_ss, d_ws, st)) return -1;
	ERT_CUDA_CHECK(cudaMemcpyAsync(score, d_ss, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, st));
	ERT_CUDA_CHECK(cudaStreamSynchronize(st));
	return 0;
}

int ert_cascade_classify_u8(ert_ctx *c, const uint8_t *hist, int n, int32_t *label, double *ss, double *ws)
{
	if (!c || !hist || n < 0) { set_error("bad arguments"); return -1; }
	if (!c->casc[0].loaded || !c->casc[1].loaded) { set_error("cascades are not loaded"); return -1; }
	if (n == 0) return 0;
	ERT_CUDA_CHECK(cudaSetDevice(c->device));
	cudaStream_t st = c->stream;
	if (c->s2.ensure((size_t)n * 1024) || c->s3.ensure((size_t)n * (sizeof(int32_t) + 2 * sizeof(double)) + 64)) return -1;
	ERT_CUDA_CHECK(cudaMemcpyAsync(c->s2.p, hist, (size_t)n * 1024, cudaMemcpyHostToDevice, st));
	int32_t *d_label = (int32_t *)c->s3.p;
	double *d_ss = (double *)((uint8_t *)c->s3.p + (((size_t)n * 4 + 63) / 64) * 64);
	double *d_ws = d_ss + n;
	if (n < 32768 && ensure_cascade_scratch(c, n, 1)) return -1;
	if (launch_cascade_u8((const uint8_t *)c->s2.p, 1024, n, nullptr, 0, n, c->casc[0].dev(), c->casc[1].dev(), d_label, d_ss, d_ws, c->csc, st)) return -1;
	if (label) ERT_CUDA_CHECK(cudaMemcpyAsync(label, d_label, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost, st));
	if (ss) ERT_CUDA_CHECK(cudaMemcpyAsync(ss, d_ss, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, st));
	if (ws) ERT_CUDA_CHECK(cudaMemcpyAsync(ws, d_ws, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, st));
	ERT_CUDA_CHECK(cudaStreamSynchronize(st));
	return 0;
}

static int svm_common(ert_ctx *c, const double *xf, const uint8_t *xu, int n, double *label, double *prob)
{
	if (!c || (!xf && !xu) || n < 0) { set_error("bad arguments"); return -1; }
	if (!c->svm.loaded) { set_error("svm model is not loaded"); return -1; }
	if (n == 0) return 0;
	ERT_CUDA_CHECK(cudaSetDevice(c->device));
	cudaStream_t st = c->stream;
	const SvmHost &m = c->svm;
	const size_t xb = xf ? sizeof(double) * (size_t)n * m.dims : (size_t)n * m.dims;
	if (c->s0.ensure(xb) || c->s1.ensure(svm_ws_bytes(m.dev(), n)) || c->s3.ensure(sizeof(double) * (size_t)n * (m.nr_class + 1))) return -1;
	ERT_CUDA_CHECK(cudaMemcpyAsync(c->s0.p, xf ? (const void *)xf : (const void *)xu, xb, cudaMemcpyHostToDevice, st));
	double *d_label = (double *)c->s3.p, *d_prob = d_label + n;
	uint8_t *tcws = nullptr;
	if (xu && m.use_tc && m.d_svj) { if (c->s4.ensure(svm_tc_ws_bytes(n))) return -1; tcws = (uint8_t *)c->s4.p; }
	if (launch_svm_predict(m.dev(), xf ? (const double *)c->s0.p : nullptr, xu ? (const uint8_t *)c->s0.p : nullptr, n, (double *)c->s1.p, d_label, d_prob, st, tcws)) return -1;
	if (label) ERT_CUDA_CHECK(cudaMemcpyAsync(label, d_label, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, st));
	if (prob) ERT_CUDA_CHECK(cudaMemcpyAsync(prob, d_prob, sizeof(double) * (size_t)n * m.nr_class, cudaMemcpyDeviceToHost, st));
	ERT_CUDA_CHECK(cudaStreamSynchronize(st));
	return svm_gemm_flag_check(tcws, n);
}

int ert_svm_predict_probability_batch(ert_ctx *c, const double *x, int n, double *label, double *prob) { return svm_common(c, x, nullptr, n, label, prob); }
int ert_svm_predict_probability_batch_u8(ert_ctx *c, const uint8_t *x, int n, double *label, double *prob) { return svm_common(c, nullptr, x, n, label, prob); }

int ert_bench_cascade_u8(ert_ctx *c, const uint8_t *hist, int n, int iters, double *ms_per_iter)
{
	if (!c || !hist || n < 1 || iters < 1) { set_error("bad arguments"); return -1; }
	if (!c->casc[0].loaded || !c->casc[1].loaded) { set_error("cascades are not loaded"); return -1; }
	ERT_CUDA_CHECK(cudaSetDevice(c->device));
	cudaStream_t st = c->stream;
	if (c->s2.ensure((size_t)n * 1024) || c->s3.ensure((size_t)n * (sizeof(int32_t) + 2 * sizeof(double)) + 64)) return -1;
	ERT_CUDA_CHECK(cudaMemcpyAsync(c->s2.p, hist, (size_t)n * 1024, cudaMemcpyHostToDevice, st));
	int32_t *d_label = (int32_t *)c->s3.p;
	double *d_ss = (double *)((uint8_t *)c->s3.p + (((size_t)n * 4 + 63) / 64) * 64);
	double *d_ws = d_ss + n;
	if (n < 32768 && ensure_cascade_scratch(c, n, 1)) return -1;
	for (int w = 0; w < 3; w++)
		if (launch_cascade_u8((const uint8_t *)c->s2.p, 1024, n, nullptr, 0, n, c->casc[0].dev(), c->casc[1].dev(), d_label, d_ss, d_ws, c->csc, st)) return -1;
	ERT_CUDA_CHECK(cudaEventRecord(c->ev[6], st));
	for (int i = 0; i < iters; i++)
		if (launch_cascade_u8((const uint8_t *)c->s2.p, 1024, n, nullptr, 0, n, c->casc[0].dev(), c->casc[1].dev(), d_label, d_ss, d_ws, c->csc, st)) return -1;
	ERT_CUDA_CHECK(cudaEventRecord(c->ev[7], st));
	ERT_CUDA_CHECK(cudaStreamSynchronize(st));
	float ms = 0;
	ERT_CUDA_CHECK(cudaEventElapsedTime(&ms, c->ev[6], c->ev[7]));
	*ms_per_iter = (double)ms / iters;
	return 0;
}

int ert_bench_svm_u8(ert_ctx *c, const uint8_t *x, int n, int iters, double *ms_per_iter)
{
	if (!c || !x || n < 1 || iters < 1) { set_error("bad arguments"); return -1; }
	if (!c->svm.loaded) { set_error("svm model is not loaded"); return -1; }
	ERT_CUDA_CHECK(cudaSetDevice(c->device));
	cudaStream_t st = c->stream;
	const SvmHost &m = c->svm;
	if (c->s0.ensure((size_t)n * m.dims) || c->s1.ensure(svm_ws_bytes(m.dev(), n)) || c->s3.ensure(sizeof(double) * (size_t)n * (m.nr_class + 1))) return -1;
	ERT_CUDA_CHECK(cudaMemcpyAsync(c->s0.p, x, (size_t)n * m.dims, cudaMemcpyHostToDevice, st));
	double *d_label = (double *)c->s3.p, *d_prob = d_label + n;
	uint8_t *tcws = nullptr;
	if (m.use_tc && m.d_svj) { if (c->s4.ensure(svm_tc_ws_bytes(n))) return -1; tcws = (uint8_t *)c->s4.p; }
	if (launch_svm_predict(m.dev(), nullptr, (const uint8_t *)c->s0.p, n, (double *)c->s1.p, d_label, d_prob, st, tcws)) return -1;
	ERT_CUDA_CHECK(cudaEventRecord(c->ev[6], st));
	for (int i = 0; i < iters; i++)
		if (launch_svm_predict(m.dev(), nullptr, (const uint8_t *)c->s0.p, n, (double *)c->s1.p, d_label, d_prob, st, tcws)) return -1;
	ERT_CUDA_CHECK(cudaEventRecord(c->ev[7], st));
	ERT_CUDA_CHECK(cudaStreamSynchronize(st));
	float ms = 0;
	ERT_CUDA_CHECK(cudaEventElapsedTime(&ms, c->ev[6], c->ev[7]));
	*ms_per_iter = (double)ms / iters;
	return svm_gemm_flag_check(tcws, n);
}

} // extern "C"
