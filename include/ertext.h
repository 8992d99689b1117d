/* ertext.h -- C ABI of libertext.so: the B200-native Extremal-Region detect + classify path.
 *
 * This is the drop-in boundary for the hot path of HsiehYiChia/Scene-text-recognition.  Every
 * entry point names the reference interface (file:line in that repository) it replaces.  Plain
 * pointers and sizes only; all buffers are caller owned unless stated; result views returned by
 * the library stay valid until the next call on the same context.  Functions return 0 on
 * success and a negative value on error (message via ert_last_error()); there is no CPU
 * fallback -- without a CUDA device ert_create() fails.
 *
 * Thread model: one ert_ctx per host thread / CUDA stream (the reference calls its stage
 * functions from 6 OpenMP threads on one ERFilter, src/ER.cpp:50-60; here the 6 planes of a frame
 * are one batched launch instead).
 */
#ifndef ERTEXT_H
#define ERTEXT_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ERT_API __attribute__((visibility("default")))

typedef struct ert_ctx ert_ctx;

/* Constructor arguments of ERFilter (inc/ER.h:113, src/ER.cpp:14-19; defaults = inc/utils.h:6-11
 * as passed by src/main.cpp:22). */
typedef struct {
	int thresh_step;      /* THRESH_STEP   8      (supported range 5..255) */
	int min_area;         /* MIN_AREA      120    */
	int max_area;         /* MAX_AREA      900000 */
	int stability_t;      /* STABILITY_T   2      */
	double overlap_coef;  /* OVERLAP_COEF  0.7    */
	double min_ocr_prob;  /* MIN_OCR_PROB  0.15   (carried for the facade, unused on this path) */
} ert_params;

/* One tree node = the fields of `struct ER` (inc/ER.h:42-80) that the path produces:
 * level, area (reference semantics: pixels + nodes of the subtree), bound, tree links.
 * Nodes of a plane are stored in DFS pre-order, children in visiting order; `parent` indexes
 * the same plane-local array (-1 for the root). */
typedef struct {
	int32_t level, area, x, y, w, h, parent, n_children;
} ert_node;

enum { ERT_LABEL_NONE = 0, ERT_LABEL_WEAK = 1, ERT_LABEL_STRONG = 2 };
enum { ERT_CASCADE_STRONG = 0, ERT_CASCADE_WEAK = 1 };
enum { ERT_STAGE_EXTRACT = 1, ERT_STAGE_NMS = 2, ERT_STAGE_CLASSIFY = 3, ERT_STAGE_TRACK = 4 };

/* Result of a batch: n_planes = 6 * n_frames for BGR input (plane order of
 * ERFilter::compute_channels, src/ER.cpp:122-127: Y, Cr, Cb, 255-Y, 255-Cr, 255-Cb), or the
 * number of planes given to ert_planes_detect.  Per plane p:
 *   nodes      [node_offset[p] .. node_offset[p+1])      kept nodes  (root[] / the ER tree)
 *   pool_*     [pool_offset[p] .. pool_offset[p+1])      NMS survivors in visiting order (pool[])
 *   pool_node  index of the pooled node inside the plane's node range
 *   pool_label ERT_LABEL_STRONG / WEAK / NONE             (strong[] / weak[] of classify)
 *   pool_strong_score / pool_weak_score  CascadeBoost::predict return values (-DBL_MAX = rejected) */
typedef struct {
	int32_t n_planes;
	int32_t width, height;
	const int32_t *node_offset;     /* n_planes + 1 */
	const ert_node *nodes;
	const int32_t *pool_offset;     /* n_planes + 1 */
	const int32_t *pool_node;
	const int32_t *pool_label;
	const double *pool_strong_score;
	const double *pool_weak_score;
	const uint8_t *pool_hist;       /* 1024 bins per pooled region, or NULL unless requested */
	uint32_t status;                /* 0 = ok; bit flags on capacity overflow (see ert_status_string) */
	/* device time per stage in milliseconds (CUDA events on the context's stream), mirroring
	 * the reference's times[] of ERFilter::text_detect (src/ER.cpp:99-110):
	 * [0] extract (incl. channels) [1] nms [2] classify [3] h2d [4] d2h [5] total
	 * [6] the tile-build kernel alone (the dominant kernel; roofline numerator) [7] rest of extract */
	double stage_ms[8];
	/* The one documented deviation made observable: per plane, the number of tree nodes into which the
	 * overlap chains of TWO OR MORE children could continue (src/ER.cpp:455-462).  Only there does non_maximum_supression's
	 * result depend on the order of siblings -- the reference's flood order, here a canonical order; everywhere else the
	 * pool is provably the reference's.  -1 per plane / in total when the sequential audit walk ran (it does not count). */
	const int32_t *plane_order_sensitive;   /* n_planes */
	int32_t order_sensitive_total;
} ert_result;

/* One strong or weak region as ERFilter::er_track leaves it (src/ER.cpp:536-558): the fields of `struct ER`
 * (inc/ER.h:42-80) that er_track / calc_color set, plus where the region sits in ert_result. */
typedef struct {
	int32_t plane;                 /* er->ch: channel 0..5 inside its frame */
	int32_t pool_index;            /* position inside the plane's pool range of ert_result (-1: caller-supplied region) */
	int32_t node;                  /* index inside the plane's node range of ert_result (-1: caller-supplied region) */
	int32_t label;                 /* ERT_LABEL_STRONG or ERT_LABEL_WEAK */
	int32_t level, area;           /* er->level, er->area */
	int32_t x, y, w, h;            /* er->bound */
	int32_t center_x, center_y;    /* er->center = bound.tl + bound.size / 2 */
	double color1, color2, color3; /* calc_color (src/ER.cpp:1391-1437); NaN when the OTSU mask is empty (0/0 there too) */
} ert_tracked;

/* Result of er_track for a batch.  Per frame f:
 *   cand    [cand_offset[f] .. cand_offset[f+1])   every strong region (strong[0..5], pool order) followed by every
 *                                                  weak region (weak[0..5]); the first n_strong[f] are the strong ones
 *   tracked [track_offset[f] .. track_offset[f+1]) `ERs &tracked` (all_er) in the reference's order, as indices
 *                                                  into the frame's candidate range */
typedef struct {
	int32_t n_frames;
	const int32_t *cand_offset;
	const int32_t *n_strong;
	const ert_tracked *cand;
	const int32_t *track_offset;
	const int32_t *tracked;
	double track_ms;               /* device time of the er_track kernels (CUDA events) */
} ert_track_result;

/* One region handed to OCR::chain_run (src/ER.cpp:732: channel[er->ch](er->bound), er->level*THRESH_STEP, text.slope).
 * The `thresh` argument of chain_run is not carried: THRESH_OTSU ignores it (src/OCR.cpp:72). */
typedef struct {
	int32_t frame;          /* frame inside the context's last batch (ert_ocr_chain_run_batch); ignored by the plane form */
	int32_t plane;          /* er->ch, channel 0..5 of that frame; ignored by the plane form */
	int32_t x, y, w, h;     /* er->bound */
	double slope;           /* Text::slope; |slope| <= 0.01 means no rotation (src/OCR.cpp:73) */
} ert_ocr_region;

/* Result of a chain_run batch; arrays of n entries, valid until the next OCR call on the context. */
typedef struct {
	int32_t n, nr_class;
	const double *value;    /* chain_run's return value: table[label] + prob_estimates[label] (src/OCR.cpp:139);
	                           the caller splits it as the reference does: letter = (char)floor(v), prob = v - floor(v) */
	const int32_t *label;   /* svm_predict_probability's label = index into the reference's table[] (src/OCR.cpp:10) */
	const double *prob_all; /* n x nr_class probability estimates */
	const uint8_t *feat;    /* n x 1800 feature bytes (svm_node value = byte / 255; zero = absent node; src/OCR.cpp:203-218) */
	const uint8_t *img;     /* n x 30 x 30: the image extract_feature received (after OTSU, rotate_mat, ARAN) */
	double ocr_ms;          /* device time, features + SVM (CUDA events) */
} ert_ocr_result;

ERT_API int ert_abi_version(void);   /* 2 since ert_result carries plane_order_sensitive (round 2) */
ERT_API const char *ert_last_error(void);
ERT_API const char *ert_status_string(uint32_t status);

/* new ERFilter(...)  (src/main.cpp:22).  device = CUDA ordinal. */
ERT_API ert_ctx *ert_create(const ert_params *params, int device);
ERT_API void ert_destroy(ert_ctx *ctx);
/* ERFilter::set_thresh_step / set_min_area  (src/ER.cpp:21-30) */
ERT_API int ert_set_thresh_step(ert_ctx *ctx, int step);
ERT_API int ert_set_min_area(ert_ctx *ctx, int min_area);
/* all constructor parameters at once (applies from the next batch on) */
ERT_API int ert_set_params(ert_ctx *ctx, const ert_params *params);
/* option: also return the 1024-bin histograms of pooled regions (costs a D2H copy) */
ERT_API int ert_set_return_hist(ert_ctx *ctx, int on);
/* audit / A-B: 1 = non_maximum_supression's walk on ONE thread per plane, literally as the reference orders it;
 * 0 (default) = the level-parallel statement of the same result (er_nms.cu).  Pools are identical either way. */
ERT_API int ert_set_nms_sequential(ert_ctx *ctx, int on);
/* scheduling: 1 = the tile-build kernels of all contexts on a device run in submission order (an event chain: the oldest
 * batch in flight is never starved; round-1 default); 0 (default) = unchained: with the post-tile kernels capped
 * (ert_set_post_footprint) the tile kernels of neighbouring batches fill each other's tails, +3 % throughput */
ERT_API int ert_set_tile_fifo(ert_ctx *ctx, int on);
/* scheduling: 1 (default) = the stages after the tile-build kernel (seams, fold, refit, NMS, classify, result compaction,
 * er_track) run on a second, highest-priority stream of the context and are joined back into the context's stream:
 * their narrow kernels then take CTA slots between the tile CTAs of the other contexts' batches; 0 = one stream */
ERT_API int ert_set_stream_split(ert_ctx *ctx, int on);
/* tuning / A-B: variant of the tile-build kernel (k_tile_build2: 64x32 tile + halo through one tensor-map TMA box, 256
 * threads, 4 pixels per lane): 0 (default) = arrival-counter fold + horizontal edge skip; 1 = neither; 2 = fold only;
 * 3 = edge skip only.  All produce identical results. */
ERT_API int ert_set_tile_config(ert_ctx *ctx, int id);
/* A-B: 1 (default) = seams through k_seam_link_list (edges compacted per CTA, warp-converged drain); 0 = k_seam_link_rec */
ERT_API int ert_set_seam_list(ert_ctx *ctx, int on);
/* scheduling: CTAs per SM the post-tile kernels (seam / fold / refit / emit) may occupy (default 1, applied to batches of more than 12 planes; 0 = no cap).  They run at
 * high priority under the NEXT batch's tile kernel; uncapped they fill the SMs first and the two run one after the other. */
ERT_API int ert_set_post_footprint(ert_ctx *ctx, int ctas_per_sm);
/* debug: per-phase cycle sums (clock64, thread 0 of every CTA) of the tile-build kernel since the last call */
ERT_API int ert_debug_phase_cycles(ert_ctx *ctx, int enable, unsigned long long *out16);
/* capacity hints (defaults: 16384 kept nodes and 2048 pooled regions per plane) */
ERT_API int ert_set_capacity(ert_ctx *ctx, int kept_per_plane, int pool_per_plane);
/* slots of the global node arrays per plane (tile-local nodes that leave their tiles); 0 = default: one per 4 pixels, at
 * least 8192 (one per pixel while MIN_AREA < 32).  A batch that needs more reports the status flag "node-overflow". */
ERT_API int ert_set_node_capacity(ert_ctx *ctx, int slots_per_plane);

/* new CascadeBoost(path) -> CascadeBoost::load_classifier  (src/adaboost.cpp:498-501, 873-951).
 * Parses the reference's text format byte-compatibly.  which = ERT_CASCADE_STRONG | _WEAK. */
ERT_API int ert_load_cascade(ert_ctx *ctx, int which, const char *path);
/* A cascade handed over in memory instead of a file: what CascadeBoost holds after load_classifier / training
 * (num_of_iter[], thresh[], and per stump RealDecisionStump::get_para() = dim, thresh, cp, cn; src/adaboost.cpp:138-146). */
ERT_API int ert_set_cascade(ert_ctx *ctx, int which, int n_stages, const int *stage_len, const int *stage_thr, int n_stumps, const int *dim,
                            const double *thr, const double *cp, const double *cn);
/* the header of a loaded cascade (num_of_iter / threshold lines, src/adaboost.cpp:896-921): returns the number of stages and
 * fills up to cap entries of stage_len / stage_thr (either may be NULL); CascadeBoost::get_num_iter = sum of stage_len */
ERT_API int ert_cascade_stage_info(ert_ctx *ctx, int which, int *stage_len, int *stage_thr, int cap);
/* svm_load_model  (src/svm.cpp:2876; inc/svm.h:77).  c_svc + rbf probability models. */
ERT_API int ert_load_svm(ert_ctx *ctx, const char *path);
ERT_API int ert_svm_nr_class(ert_ctx *ctx);   /* svm_get_nr_class (inc/svm.h:81) */
ERT_API int ert_svm_total_sv(ert_ctx *ctx);   /* svm_get_nr_sv (inc/svm.h:84) */
ERT_API int ert_svm_labels(ert_ctx *ctx, int *label);   /* svm_get_labels (inc/svm.h:82); returns nr_class */
ERT_API double ert_svm_gamma(ert_ctx *ctx);   /* model->param.gamma */
/* u8 features: 1 (default) = RBF distances as two integer tcgen05 GEMMs (kind::i8; TMA ring, TMEM double-buffered);
 * 2 = the same GEMMs through the round-1 single-stage kernel (A-B); 0 = FP64 CUDA-core kernel */
ERT_API int ert_set_svm_tensor_cores(ert_ctx *ctx, int on);
/* A-B: 1 = the round-1 probability kernel (one warp per vector walks the coefficient table); 0 (default) = k_svm_decide + k_svm_couple */
ERT_API int ert_set_svm_legacy_prob(ert_ctx *ctx, int on);
ERT_API int ert_svm_dims(ert_ctx *ctx);

/* ---- the batched hot path ---------------------------------------------------------------
 * ERFilter::text_detect up to and including classify (src/ER.cpp:33-60): compute_channels ->
 * per plane er_tree_extract -> non_maximum_supression -> classify, for n_frames BGR frames
 * (8-bit, 3 channels interleaved, row stride in bytes) in HOST memory.  upto = ERT_STAGE_*. */
ERT_API int ert_detect_classify(ert_ctx *ctx, const uint8_t *bgr, int n_frames, int width, int height, int stride_bytes,
                                int upto, const ert_result **out);
/* Asynchronous form of the above: enqueue the H2D copy (host memory should be pinned for a truly
 * asynchronous copy) and all kernels on the context's stream, return immediately; the result is
 * collected by ert_fetch_result.  Two contexts used alternately overlap copy and compute. */
ERT_API int ert_enqueue_host(ert_ctx *ctx, const uint8_t *bgr, int n_frames, int width, int height, int stride_bytes, int upto);
/* Same, frames already resident on the device (device pointer); enqueue only. The result is
 * fetched (one stream sync + small D2H) by ert_fetch_result. */
ERT_API int ert_detect_classify_device(ert_ctx *ctx, const void *d_bgr, int n_frames, int width, int height, int stride_bytes, int upto);
ERT_API int ert_fetch_result(ert_ctx *ctx, const ert_result **out);
/* non-blocking: 1 = the batch enqueued last on this context has finished (ert_fetch_result will not wait),
 * 0 = still running, -1 = nothing in flight */
ERT_API int ert_batch_done(ert_ctx *ctx);

/* ERFilter::compute_channels(src, YCrCb, channels)  (src/ER.cpp:114-128): BGR -> the six 8-bit planes
 * Y, Cr, Cb, 255-Y, 255-Cr, 255-Cb (OpenCV's 8-bit BGR2YCrCb arithmetic), written to planes6 as six
 * contiguous width*height images in host memory.  The batched entry points above fuse this step; this
 * call exists for callers that need the planes themselves (er_track / er_ocr read them). */
ERT_API int ert_compute_channels(ert_ctx *ctx, const uint8_t *bgr, int width, int height, int stride_bytes, uint8_t *planes6);

/* ERFilter::er_tree_extract(Mat) / non_maximum_supression / classify on caller-supplied
 * single-channel planes (src/ER.cpp:240, 416, 507; inc/ER.h:125-127): n_planes images of the same
 * size, plane k at planes + k*plane_stride_bytes, rows stride_bytes apart, host memory. */
ERT_API int ert_planes_detect(ert_ctx *ctx, const uint8_t *planes, int n_planes, int width, int height, int stride_bytes,
                              size_t plane_stride_bytes, int upto, const ert_result **out);
/* Asynchronous form (collect with ert_fetch_result; the host planes must stay valid until then unless pinned copies were
 * enqueued).  One context per scale runs the levels of a pyramid concurrently (BASELINE config 4). */
ERT_API int ert_enqueue_planes(ert_ctx *ctx, const uint8_t *planes, int n_planes, int width, int height, int stride_bytes,
                               size_t plane_stride_bytes, int upto);

/* Image pyramid on the device (BASELINE configs 2 and 4; the reference itself runs native scale only, src/ER.cpp:122-127,
 * and SURVEY 8d defines a level as the same per-plane path on the cv::resize(INTER_LINEAR)'d plane, the primitive of
 * src/OCR.cpp:401).  `src` = a context with a batch in flight (any of the enqueue calls above); `dst` = another context of
 * the same device: receives src's source planes resized to (width / div, height / div) -- bit-exact cv::resize semantics
 * incl. the exact-2x area path -- and runs the path on them.  Collect each level with ert_fetch_result(dst).  src's next
 * batch waits on the device until every level has read its planes. */
ERT_API int ert_enqueue_pyramid_level(ert_ctx *dst, ert_ctx *src, int div, int upto);
/* BGR entry points: 6 (default) = the six channels of compute_channels; 3 = Y, Cr, Cb only (BASELINE config 4) */
ERT_API int ert_set_planes_per_frame(ert_ctx *ctx, int n);

/* ERFilter::non_maximum_supression(ER *root, ..., pool, input) on a caller tree (src/ER.cpp:416):
 * nodes in DFS pre-order with the caller's child order (as ert_result delivers them, or as
 * flattened from a reference ER* tree).  pool_out receives node indices in push order. */
ERT_API int ert_nms_nodes(ert_ctx *ctx, const ert_node *nodes, int n_nodes, int width, int height, int32_t *pool_out, int pool_cap,
                          int *n_pool);

/* ERFilter::classify(pool, strong, weak, input) on caller rectangles (src/ER.cpp:507-528):
 * rects = n x (x, y, w, h) inside the given plane.  Any of the outputs may be NULL. */
ERT_API int ert_classify_regions(ert_ctx *ctx, const uint8_t *plane, int width, int height, int stride_bytes, const int32_t *rects, int n,
                                 int32_t *label, double *strong_score, double *weak_score, uint8_t *hist1024);
/* ERFilter::make_LBP_hist(input(rect), 2, 24)  (src/ER.cpp:789-816): hist = n x 1024 doubles */
ERT_API int ert_lbp_hist(ert_ctx *ctx, const uint8_t *plane, int width, int height, int stride_bytes, const int32_t *rects, int n,
                         double *hist);

/* ERFilter::calc_LBP(input(rect), 24)  (src/ER.cpp:819-845): the 24 x 24 mean-LBP code image of every rect, codes = n x 576 bytes */
ERT_API int ert_calc_lbp(ert_ctx *ctx, const uint8_t *plane, int width, int height, int stride_bytes, const int32_t *rects, int n,
                         uint8_t *codes);

/* CascadeBoost::predict(vector<double> fv)  (src/adaboost.cpp:507-542) for n feature vectors of
 * `dims` doubles each; score = predict's return value (-DBL_MAX when a stage rejects). */
ERT_API int ert_cascade_predict_batch(ert_ctx *ctx, int which, const double *fv, int n, int dims, double *score);
/* Both cascades on 1024-bin u8 histograms (the pipeline's own feature format), device-side sweep entry */
ERT_API int ert_cascade_classify_u8(ert_ctx *ctx, const uint8_t *hist, int n, int32_t *label, double *strong_score, double *weak_score);

/* svm_predict_probability(model, x, prob_estimates)  (src/svm.cpp:2592; inc/svm.h:88) for a batch:
 * x = n dense rows of ert_svm_dims() doubles (zero = absent svm_node), label[n] = returned label,
 * prob = n x nr_class.  The _u8 form takes features as k in 0..255 meaning k/255.0
 * (OCR::extract_feature, src/OCR.cpp:203-218). */
ERT_API int ert_svm_predict_probability_batch(ert_ctx *ctx, const double *x, int n, double *label, double *prob);
ERT_API int ert_svm_predict_probability_batch_u8(ert_ctx *ctx, const uint8_t *x, int n, double *label, double *prob);

/* ---- after the detect path ------------------------------------------------------------
 * ERFilter::er_track(strong, weak, tracked, channel, Ycrcb)  (src/ER.cpp:532-609; called at src/ER.cpp:63 and
 * src/utils.cpp:140): calc_color + center + ch for every strong / weak region, then the strong-seeded greedy
 * growth of `tracked`.  Runs on the regions of the batch this context processed last (BGR entry points; the
 * planes and labels are still on the device).  A batch enqueued with upto = ERT_STAGE_TRACK runs it in the
 * same stream submission and ert_er_track only collects the result. */
ERT_API int ert_er_track(ert_ctx *ctx, const ert_track_result **out);
/* The same on caller-supplied regions of ONE host BGR frame: strong / weak = rows of 6 ints
 * (ch, x, y, w, h, area), channel-major as classify fills strong[ch] / weak[ch]. */
ERT_API int ert_er_track_regions(ert_ctx *ctx, const uint8_t *bgr, int width, int height, int stride_bytes, const int32_t *strong,
                                 int n_strong, const int32_t *weak, int n_weak, const ert_track_result **out);
/* The same with the frame given as the three planes compute_channels produced (channel[0..2] = Y, Cr, Cb, rows
 * stride_bytes apart): exactly what ERFilter::er_track's (channel, Ycrcb) arguments hold (src/ER.cpp:532, 545-546). */
ERT_API int ert_er_track_regions_ycc(ert_ctx *ctx, const uint8_t *y, const uint8_t *cr, const uint8_t *cb, int width, int height, int stride_bytes,
                                     const int32_t *strong, int n_strong, const int32_t *weak, int n_weak, const ert_track_result **out);

/* OCR::chain_run(src, thresh, slope) for a batch of regions (src/OCR.cpp:67-140; called from ERFilter::er_ocr,
 * src/ER.cpp:728-735): threshold(255 - src, OTSU) -> rotate_mat when |slope| > 0.01 -> ARAN(30) -> extract_feature
 * (findContours chain codes, GaussianBlur, normalize, resize) -> svm_predict_probability.  Needs ert_load_svm.
 * _batch: regions of the BGR batch this context processed last (planes still on the device; only the region records
 * travel).  _plane: regions of one caller-supplied single-channel image in host memory. */
ERT_API int ert_ocr_chain_run_batch(ert_ctx *ctx, const ert_ocr_region *regions, int n, const ert_ocr_result **out);
ERT_API int ert_ocr_chain_run_plane(ert_ctx *ctx, const uint8_t *plane, int width, int height, int stride_bytes,
                                    const ert_ocr_region *regions, int n, const ert_ocr_result **out);
/* OCR::extract_feature path only (no SVM model needed): out->value/label/prob_all are NULL. */
ERT_API int ert_ocr_features_plane(ert_ctx *ctx, const uint8_t *plane, int width, int height, int stride_bytes,
                                   const ert_ocr_region *regions, int n, const ert_ocr_result **out);

/* ---- multi-GPU: the final region gather -----------------------------------------------------------
 * One process per GPU; rank r of G owns frames {f : f mod G = r}; no collective on the compute path.  After a batch the
 * labelled regions (strong[] / weak[] of ERFilter::classify, src/ER.cpp:507-528) of all ranks are gathered on every rank:
 * packed on the device, per-rank counts by ncclAllGather, then ONE grouped exchange of exactly the records each rank
 * holds, on the gather's own side stream, pipelined behind the data path (collected one or two batches later).
 * NCCL is loaded at run time (libnccl.so.2); world = 1 needs no NCCL at all. */
typedef struct ert_dist ert_dist;
typedef struct {
	int32_t frame, plane;          /* global frame id (frame_ids[local frame]) and channel 0..5 */
	int32_t level, area, x, y, w, h;
	int32_t label;                 /* ERT_LABEL_STRONG / ERT_LABEL_WEAK */
	int32_t pool_index;            /* position inside the plane's pool (classify's push order) */
} ert_region_record;
typedef struct {
	int32_t world;
	int32_t n_records;
	const int32_t *rank_offset;            /* world + 1: records of rank r are [rank_offset[r], rank_offset[r+1]) */
	const ert_region_record *records;      /* pinned host memory, valid until the next collect on this handle */
	long long sequence;                    /* which enqueue this is (0, 1, 2, ...) */
} ert_gather_result;
/* rank 0: a fresh NCCL unique id (128 bytes) to hand to every rank out of band (torch.distributed, MPI, a file) */
ERT_API int ert_dist_unique_id(void *id128);
/* collective over all ranks (ncclCommInitRank); id128 may be NULL for world = 1 */
ERT_API ert_dist *ert_dist_create(int device, int rank, int world, const void *id128, int max_records_per_rank);
ERT_API void ert_dist_destroy(ert_dist *d);
/* call right after ert_enqueue_host / ert_detect_classify_device on ctx (batch still in flight), on every rank, in the same
 * order: packs the batch's labelled regions on the device and starts the gather; returns at once.  frame_ids[n_frames] = the
 * global ids of the batch's frames (NULL: 0..n_frames-1).  At most 11 gathers may be outstanding; the exact-size exchange of a
 * batch is issued six enqueue calls later (or by the collect call that needs it): keep at most six batches in flight on the
 * data path, otherwise that call waits for the batch's counts and the host stops running ahead of the device (measured:
 * lag 4 under 5 contexts cost 14 % of the host-input throughput at 2 GPUs). */
ERT_API int ert_gather_regions_enqueue(ert_dist *d, ert_ctx *ctx, const int32_t *frame_ids, int n_frames);
/* the oldest outstanding gather; blocks only if it has not finished (collective when it has to drain) */
ERT_API int ert_gather_regions_collect(ert_dist *d, const ert_gather_result **out);
ERT_API int ert_gather_regions_outstanding(ert_dist *d);

/* ---- compressed input ------------------------------------------------------ */
/* replaces `cap >> frame` + compute_channels of video_mode (src/utils.cpp:107-113) for JPEG frames: n_frames baseline-JPEG
 * bitstreams (host memory), each W x H, are decoded by nvJPEG straight into the context's device BGR buffer on the batch's
 * stream; then as ert_detect_classify_device.  Collect with ert_fetch_result.  nvJPEG is opened with dlopen at first use. */
ERT_API int ert_enqueue_jpeg(ert_ctx *ctx, const uint8_t *const *data, const size_t *sizes, int n_frames, int W, int H, int upto);
/* the decoded pixels of the last ert_enqueue_jpeg ([n_frames][H][W][3] BGR): what parity is stated on (nvJPEG's IDCT and
 * chroma upsampling differ from libjpeg's in the last bit) */
ERT_API int ert_jpeg_fetch_frames(ert_ctx *ctx, uint8_t *bgr_out);
/* backend -1 = best available (hardware engine, else GPU Huffman, else default) or an nvjpegBackend_t value; cpu_threads >= 1 */
ERT_API int ert_set_jpeg_backend(ert_ctx *ctx, int backend, int cpu_threads);
ERT_API const char *ert_jpeg_backend_name(ert_ctx *ctx);
ERT_API double ert_jpeg_decode_ms(ert_ctx *ctx);

/* ---- plumbing -------------------------------------------------------------------------------- */
/* use an external CUDA stream (cudaStream_t as integer, e.g. torch.cuda.current_stream().cuda_stream); 0 = own stream */
ERT_API int ert_set_stream(ert_ctx *ctx, uint64_t cuda_stream);
ERT_API uint64_t ert_get_stream(ert_ctx *ctx);
/* page-locked host memory for frame staging: ert_enqueue_host copies asynchronously only from pinned memory */
ERT_API void *ert_host_alloc(size_t bytes);
ERT_API void ert_host_free(void *p);
/* number of kernel launches issued by the last batch call (bench's gpu_launches) */
ERT_API int ert_last_launch_count(ert_ctx *ctx);
/* device-side benchmark helpers for the classifier sweeps (inputs generated/resident on device):
 * time `iters` back-to-back launches with CUDA events; returns milliseconds per iteration. */
ERT_API int ert_bench_cascade_u8(ert_ctx *ctx, const uint8_t *hist_host, int n, int iters, double *ms_per_iter);
ERT_API int ert_bench_svm_u8(ert_ctx *ctx, const uint8_t *x_host, int n, int iters, double *ms_per_iter);

#ifdef __cplusplus
}
#endif
#endif /* ERTEXT_H */
