// ctx.h -- the context behind the C ABI (internal to libertext.so): host-side model tables, workspaces.
#pragma once
#include "../../include/ertext.h"
#include "common.cuh"
#include "kernels.h"
#include <vector>

namespace ert {

// ---------------------------------------------------------------------------------------------
// host-side model tables
// ---------------------------------------------------------------------------------------------
struct CascadeHost {
	std::vector<int> stage_len, stage_thr;
	std::vector<Stump> stumps;
	bool loaded = false;
	Stump *d_stumps = nullptr;
	int *d_len = nullptr, *d_thr = nullptr;
	std::vector<double2> cpcn; std::vector<uint32_t> dimthr;      // compact tables for u8 histograms
	double2 *d_cpcn = nullptr; uint32_t *d_dimthr = nullptr;
	CascadeDev dev() const
	{
		CascadeDev c; c.stumps = d_stumps; c.stage_len = d_len; c.stage_thr = d_thr; c.n_stages = (int)stage_len.size();
		c.cpcn = d_cpcn; c.dimthr = d_dimthr;
		return c;
	}
};

struct SvmHost {
	bool loaded = false;
	int nr_class = 0, l = 0, dims = 0;
	double gamma = 0;
	std::vector<double> rho, probA, probB, coef, sv;
	std::vector<int> label, nsv, start;
	double *d_sv = nullptr, *d_coef = nullptr, *d_rho = nullptr, *d_probA = nullptr, *d_probB = nullptr;
	std::vector<double> coefT; double *d_coefT = nullptr;
	std::vector<uint16_t> pair_ij; uint16_t *d_pair_ij = nullptr;   // pair p of the rho order -> i << 8 | j
	int legacy_prob = 0;
	int *d_label = nullptr, *d_nsv = nullptr, *d_start = nullptr;
	std::vector<uint8_t> svj; std::vector<int8_t> sve; std::vector<double> ss;
	uint8_t *d_svj = nullptr; int8_t *d_sve = nullptr; double *d_ss = nullptr;
	double inv_s255 = 0;
	int use_tc = 1;          // 0 FP64 distance kernel, 1 TMA-pipelined tcgen05 GEMM, 2 round-1 single-stage tcgen05 GEMM
	SvmDev dev() const
	{
		SvmDev m; m.nr_class = nr_class; m.l = l; m.ldk = (l + 127) & ~127; m.dims = dims; m.gamma = gamma; m.sv = d_sv; m.coef = d_coef; m.coefT = d_coefT; m.pair_ij = d_pair_ij; m.legacy_prob = legacy_prob; m.tc_variant = use_tc;
		m.rho = d_rho; m.probA = d_probA; m.probB = d_probB; m.label = d_label; m.nsv = d_nsv; m.start = d_start;
		m.svj = use_tc ? d_svj : nullptr; m.sve = d_sve; m.ss = d_ss; m.inv_s255 = inv_s255;
		return m;
	}
};

// after the stream has been synchronised: did an mbarrier wait in the tensor-core GEMM give up? (bounded waits, svm_gemm.cu)
static inline int svm_gemm_flag_check(uint8_t *tcws, int n)
{
	if (!tcws) return 0;
	uint32_t f = 0;
	ERT_CUDA_CHECK(cudaMemcpy(&f, svm_tc_flag(tcws, n), sizeof f, cudaMemcpyDeviceToHost));
	if (f) { set_error("svm: the tensor-core GEMM pipeline timed out (flag %u); results are invalid", f); return -1; }
	return 0;
}

void jpeg_decoder_destroy(void *decoder);   // ingest.cu

template <typename T>
static int dev_upload(T **dptr, const std::vector<T> &v)
{
	if (*dptr) { cudaFree(*dptr); *dptr = nullptr; }
	if (v.empty()) return 0;
	ERT_CUDA_CHECK(cudaMalloc((void **)dptr, sizeof(T) * v.size()));
	ERT_CUDA_CHECK(cudaMemcpy(*dptr, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice));
	return 0;
}

struct Scratch {
	void *p = nullptr;
	size_t cap = 0;
	int ensure(size_t bytes)
	{
		if (bytes <= cap) return 0;
		if (p) cudaFree(p);
		p = nullptr; cap = 0;
		ERT_CUDA_CHECK(cudaMalloc(&p, bytes));
		cap = bytes;
		return 0;
	}
	void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

} // namespace ert

struct ert_ctx {
	ert_params prm;
	int device = 0;
	cudaStream_t stream = nullptr;
	bool own_stream = true;
	cudaStream_t post_stream = nullptr;   // highest priority: everything after the tile kernel (seams .. result compaction, er_track)
	cudaEvent_t ev_post_done = nullptr;   // joins the post stream back into `stream`
	int split_streams = 1;
	int planes_per_frame = 6;             // BGR entry points: 6 = Y, Cr, Cb and their inverses (compute_channels); 3 = Y, Cr, Cb only
	cudaEvent_t ev_planes = nullptr;      // the batch's source planes are complete in d_ycc (pyramid levels of other contexts wait on it)
	cudaEvent_t ev_decode[2] = {nullptr, nullptr};   // around nvjpegDecodeBatched (ingest.cu)
	void *jpeg = nullptr;                 // JpegDecoder (ingest.cu), created by the first ert_enqueue_jpeg
	int jpeg_backend = -1, jpeg_cpu_threads = 4, jpeg_frames = 0, jpeg_W = 0, jpeg_H = 0;
	cudaEvent_t ev_resized = nullptr;     // this context finished READING another context's planes (ert_enqueue_pyramid_level)
	std::vector<cudaEvent_t> readers;     // events of contexts that still read this context's planes: the next batch waits for them
	cudaStream_t work_stream() const { return (split_streams && post_stream) ? post_stream : stream; }
	cudaEvent_t ev[12];
	int nms_sequential = 0;  // 1: run the reference's walk on one thread per plane (audit / A-B) instead of the level-parallel form
	int tile_fifo = 0;       // 1: chain the tile kernels of all contexts on the device in submission order (round-1 default; with the
	                         // post-tile kernels capped, unchained tile kernels fill each other's tails: 4 619 -> 4 760 frames/s)
	int tile_cfg = 0;
	int sm_count = 148;
	int post_ctas_per_sm = 1; // footprint of the post-tile kernels: they run at high priority under the next tile kernel and must leave its CTAs room
	int seam_list = 1;       // seam kernel: compacted edge lists (1) or one thread per seam position (0, round 1)
	int return_hist = 0;
	int kept_cap = 16384, pool_cap = 2048;
	int launches = 0;

	ert::CascadeHost casc[2];
	ert::SvmHost svm;
	uint8_t aran_tbl_h[64];
	uint8_t *d_aran_tbl = nullptr;
	unsigned long long *d_prof = nullptr;

	// workspace geometry
	int W = 0, H = 0, pitch = 0, planes_cap = 0, frames_cap = 0;
	size_t cap_np = 0, cap_ycc = 0, cap_ring = 0, cap_nodes = 0;   // capacities: planes x node slots, plane bytes, seam-record words
	int node_cap_user = 0;   // ert_set_node_capacity: slots per plane (0 = one per 4 pixels)
	int table_planes = 0, table_W = 0, table_H = 0, table_ppf = 0; // what the BGR-layout plane table on the device was built for
	uint8_t *d_bgr = nullptr; size_t bgr_cap = 0;
	uint8_t *d_ycc = nullptr;          // frames_cap*3 planes (BGR mode) or planes_cap planes (plane mode)
	size_t ycc_bytes = 0;
	ert::PlaneSrc *d_planes = nullptr;
	ert::ExtractWork wk{};
	uint8_t *d_nms_scratch = nullptr; size_t nms_stride = 0;
	ert::OutNode *d_out_nodes = nullptr;
	int32_t *d_out_pool = nullptr, *d_out_counts = nullptr, *d_label = nullptr;
	double *d_ss = nullptr, *d_ws = nullptr;
	uint8_t *d_hist = nullptr;
	// host-mapped result buffers
	int32_t *h_node_off = nullptr, *h_pool_off = nullptr, *h_pool = nullptr, *h_label = nullptr;
	ert::OutNode *h_nodes = nullptr;
	double *h_ss = nullptr, *h_ws = nullptr;
	uint8_t *h_hist = nullptr;
	uint32_t *h_status = nullptr;
	int32_t *d_order_sens = nullptr, *h_order_sens = nullptr;   // per plane: nodes where the NMS outcome depends on the sibling order
	ert_result res{};
	int pending_planes = 0, pending_upto = 0;
	bool pending = false;

	ert::Scratch s0, s1, s2, s3, s4;
	ert::CascadeScratch csc{};          // stage sums / arrival counters / pool prefix of the cascade kernel

	// ---- rows after the detect path (capi_next.cu): er_track and OCR::chain_run ----
	ert::TrackWork tk{};                    // device candidate / colour / tracked buffers
	int track_frames_cap = 0, track_cand_cap = 0;
	ert_tracked *h_cand = nullptr;     // host-mapped
	int32_t *h_cand_off = nullptr, *h_nstrong = nullptr, *h_track_off = nullptr, *h_tracked = nullptr;
	ert_track_result tres{};
	bool track_pending = false;
	int pending_frames = 0;
	uint8_t aran30_tbl_h[64];
	uint8_t *d_aran30_tbl = nullptr;
	ert::Scratch o0, o1, o2, o3, o4;
	std::vector<double> ocr_value, ocr_prob;
	std::vector<int32_t> ocr_label;
	std::vector<uint8_t> ocr_feat, ocr_img;
	ert_ocr_result ores{};
};

namespace ert {
// capi_next.cu
int enqueue_track(ert_ctx *c, int n_frames);
void free_next(ert_ctx *c);
} // namespace ert
