// kernels.h -- host-side launchers of the sm_100a kernels (internal to libertext.so)
#pragma once
#include "common.cuh"
#include "../../include/ertext.h"

namespace ert {

// CUtensorMap (128 bytes, 64-byte aligned) without pulling <cuda.h> into every translation unit
struct alignas(64) TileTensorMap { uint64_t opaque[16]; };

struct ExtractWork {
	uint32_t *par;          // [n_planes][node_cap] keyed forest over node slots
	NodeAttr *attr;         // [n_planes][node_cap] node attributes
	uint32_t *node_key;     // [n_planes][node_cap] level << 26 | pixel index of the node's level root (its largest own-level pixel)
	uint32_t *node_count;   // [n_planes] slots handed out
	uint32_t *start_key;    // [n_planes][4] keys (level, slot) of the nodes holding the flood's start candidates: pixels 0, 1, W
	uint32_t *reach_root;   // [n_planes]
	int32_t *lone_level;    // [n_planes]
	KeptRec *kept;          // [n_planes][kept_cap]
	uint32_t *kept_count;   // [n_planes]
	uint32_t *status;       // [1]
	int node_blocks;        // grid.x of the node-list kernels
	int post_ctas;          // CTAs the post-tile kernels (seam / fold / refit / emit) may put on the device at once; 0 = no cap
	int tile_cfg;           // index into the tile configuration table
	unsigned long long *prof;   // optional [16] per-phase cycle sums of k_tile_build (debug)
	uint32_t *ring_rec;     // [n_planes][tiles][2*(TW+TH)] root keys on the tile sides (seam records)
	int seam_list;          // 1: k_seam_link_list (compacted edge list per CTA, converged drain); 0: k_seam_link_rec (one thread per seam position)
	int node_cap;           // slots per plane
	TileTensorMap tmap;     // (x, y, source plane) view of the plane buffer for k_tile_build2's haloed TMA box
};

struct NmsParams {
	int W, H;
	size_t N;               // stride of the per-plane node attribute array (slots per plane)
	int kept_cap, pool_cap;
	int min_area, max_area, stability_t;
	double overlap_coef;
	int sequential_walk;    // 1: the one-thread walk (audit / A-B); 0: the level-parallel formulation of the same result
};

struct ClassifyParams {
	int pitch;
	int pool_cap, node_cap;
};

struct Stump { double thr, cp, cn; int dim; int pad; };

struct CascadeDev {
	const Stump *stumps;
	const int *stage_len;
	const int *stage_thr;
	int n_stages;
	// compact tables for u8 histograms (built at load time): {cp, cn} and dim | ceil(thr) << 16 per stump
	const double2 *cpcn;
	const uint32_t *dimthr;
};

// scratch of the warp-per-(region, stage) cascade kernel: stage sums, per-region arrival counters (self-resetting: zero
// between launches), first region of every plane
struct CascadeScratch {
	double *stage_sum;      // [rows][32]
	uint32_t *done;         // [rows]
	int32_t *pool_prefix;   // [planes + 1]
	int rows_cap, planes_cap;
};

struct SvmDev {
	int nr_class, l, dims;
	int ldk;                 // row stride of the kernel-value matrix K[vectors][ldk]: l rounded up to 128 (whole tiles, 16-byte stores)
	double gamma;
	const double *sv;        // [l][dims] dense support vectors
	const double *coef;      // [nr_class-1][l]
	const double *coefT;     // [l][nr_class-1] the same table SV-major (a class block's coefficients are contiguous)
	const uint16_t *pair_ij; // [nr_class*(nr_class-1)/2] classes of pair p (rho order): i << 8 | j
	int legacy_prob;         // 1: k_svm_prob (one warp per vector, A/B); 0: k_svm_decide + k_svm_couple
	int tc_variant;          // u8 features: 1 = k_svm_kvalue_tma (TMA ring, double-buffered TMEM), 2 = round-1 single-stage tcgen05 kernel (A/B)
	const double *rho, *probA, *probB;   // [nr_class*(nr_class-1)/2]
	const int *label, *nsv, *start;      // [nr_class]
	// tensor-core tables (u8 features): j = round(255 v) and e = round(S (v - j/255)), padded [2048][1920]; |sv|^2
	const uint8_t *svj;
	const int8_t *sve;
	const double *ss;
	double inv_s255;        // 1 / (255 * S)
};

// er_track (er_track.cu): per-frame candidate lists (strong then weak), colours, tracked order
struct TrackWork {
	ert_tracked *cand;      // [frames][cand_cap]
	int32_t *n_cand;        // [frames]
	int32_t *n_strong;      // [frames]
	int32_t *tracked;       // [frames][cand_cap] indices into the frame's candidates, all_er order
	int32_t *n_tracked;     // [frames]
	int cand_cap;
};

int launch_track_gather(TrackWork &tk, int n_frames, const OutNode *nodes, const int32_t *pool, const int32_t *counts, const int32_t *label,
                        int node_cap, int pool_cap, cudaStream_t st);
int launch_calc_color(TrackWork &tk, int n_frames, const uint8_t *d_ycc, size_t plane_bytes, int pitch, cudaStream_t st);
int launch_track(TrackWork &tk, int n_frames, ert_tracked *h_cand, int32_t *h_cand_off, int32_t *h_nstrong, int32_t *h_track_off,
                 int32_t *h_tracked, cudaStream_t st);

// OCR::chain_run pre-processing + extract_feature (ocr_feat.cu).  Everything that depends on libm (atan2 / cos / sin /
// tan / pow) is evaluated by the host with the C library the reference uses and shipped in the job record.
struct OcrJob {
	const uint8_t *src;     // channel plane (device), row pitch below
	int pitch, invert;      // invert: channel value = 255 - src
	int x0, y0, w, h;       // er->bound inside the plane
	int rot;                // 0: no rotation (|slope| <= 0.01); 1: OCR::rotate_mat(crop = true); 2: its crop = false fallback
	double cs, sn;          // cos(rad), sin(rad), rad = atan2(slope, 1)
	int cx, cy;             // rotate_mat's x0, y0
	int min_x, min_y, crop_h;
	int sw, sh;             // size of the image that enters ARAN (the crop, or the rotated image)
	int dw, dh, offx, offy; // ARAN: resize target and paste offset inside the 30x30 image
};

int launch_ocr_features(const OcrJob *d_jobs, int n, uint8_t *d_feat1800, uint8_t *d_img30, cudaStream_t st);

int extract_pitch(int W);
int make_tile_tensor_map(TileTensorMap *out, const uint8_t *d_planes0, int W, int H, int pitch, int n_src_planes);
int launch_tile_v2(const ExtractParams &P, const PlaneSrc *d_planes, ExtractWork &wk, cudaStream_t st, int opt);
int tile_config_count();
size_t ring_words_per_plane(int W, int H);
int launch_channels(const uint8_t *d_bgr, size_t frame_stride, int row_stride, int W, int H, int n_frames, uint8_t *d_ycc, int pitch, cudaStream_t st);
int launch_resize_planes(const PlaneSrc *d_src_planes, int n_planes, int sw, int sh, int spitch, uint8_t *d_dst, int dw, int dh, int dpitch, size_t dplane,
                         cudaStream_t st);
int launch_unpack_planes(const uint8_t *d_ycc, int pitch, int W, int H, uint8_t *d_out6, cudaStream_t st);
int launch_extract(const ExtractParams &P, const PlaneSrc *d_planes, ExtractWork &wk, cudaStream_t st,
                   cudaEvent_t ev_tile_begin = nullptr, cudaEvent_t ev_tile_end = nullptr, cudaStream_t st_post = nullptr);

size_t nms_scratch_stride(int kept_cap);
int launch_nms(const NmsParams &P, int n_planes, const KeptRec *kept, const uint32_t *kept_count, const NodeAttr *attr,
               const uint32_t *reach_root, const int32_t *lone_level, const OutNode *in_nodes, const int32_t *in_offsets,
               uint8_t *scratch, size_t scratch_stride, OutNode *out_nodes, int32_t *out_pool, int32_t *out_counts,
               uint32_t *status, cudaStream_t st, int32_t *order_sens = nullptr);

// first region of every plane: pool_prefix[p] = sum over q < p of min(pool count of plane q, pool_cap)
int launch_pool_prefix(const int32_t *counts, int n_planes, int pool_cap, int32_t *pool_prefix, cudaStream_t st);
int launch_lbp_hist(const ClassifyParams &P, int n_planes, const PlaneSrc *planes, const OutNode *nodes, const int32_t *pool,
                    const int32_t *pool_prefix, const uint8_t *aran_tbl, uint8_t *hist_out, cudaStream_t st, uint8_t *codes_out = nullptr);
int launch_cascade_u8(const uint8_t *hist, size_t row_stride, int n_rows, const int32_t *pool_prefix, int n_planes, int pool_cap, const CascadeDev &strong,
                      const CascadeDev &weak, int32_t *label, double *sscore, double *wscore, const CascadeScratch &sc, cudaStream_t st);
int launch_cascade_f64(const double *fv, size_t row_stride, int n_rows, const CascadeDev &strong, const CascadeDev &weak,
                       int32_t *label, double *sscore, double *wscore, cudaStream_t st);

// ws: svm_ws_bytes(m, n) bytes (kernel values + the two decision-term arrays of one pass of at most 32768 vectors)
size_t svm_ws_bytes(const SvmDev &m, int n);
int launch_svm_predict(const SvmDev &m, const double *x_f64, const uint8_t *x_u8, int n, double *ws, double *label, double *prob,
                       cudaStream_t st, uint8_t *tc_ws = nullptr);
size_t svm_tc_ws_bytes(int n);
uint32_t *svm_tc_flag(uint8_t *tc_ws, int n);   // device word after the workspace: nonzero = an mbarrier wait in the GEMM gave up
int launch_svm_kvalue_tma(const SvmDev &m, const uint8_t *xp, const uint32_t *xx, int n_rows, int row0, int n, double *kv, uint32_t *flag, cudaStream_t st);
int svm_tc_kpad();
int svm_tc_npad();

} // namespace ert
