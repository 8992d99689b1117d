// er_classify.cu -- stage-2 features and the two Real-AdaBoost cascades
// (replaces ERFilter::classify / make_LBP_hist / calc_LBP, src/ER.cpp:507-528, 789-845;
//  OCR::ARAN, src/OCR.cpp:394-430; CascadeBoost::predict, src/adaboost.cpp:507-542).
//
//   k_lbp_hist : one warp per candidate region.  crop -> ARAN aspect-preserving bilinear resize
//                into a zeroed 26x26 patch (OpenCV's 11-bit fixed-point INTER_LINEAR, incl. the
//                exact-2x INTER_AREA switch) -> 24x24 mean-LBP codes with the reference's literal
//                stride-24-on-stride-26 neighbour offsets -> 2x2 blocks x 256 bins, counts as u8.
//   k_cascade  : one thread per region walks the stump tables in file order with non-fused FP64
//                adds so every stage sum rounds exactly like the reference; all lanes of a warp
//                read the same stump (broadcast), gather their own histogram bin.
// Not a GEMM (gather + compare + select + ordered sum): no tensor cores by design.
#include "common.cuh"
#include "kernels.h"
#include <algorithm>

namespace ert {

// (int)(L * pow(min/max, 0.5)) without calling pow on the device: the exact floor, except when
// L*sqrt(a/b) is an integer in exact arithmetic -- there glibc's rounding decides and the host
// supplies its answer in exact_tbl[m] (see capi.cu: build_aran_table).
__device__ __forceinline__ int aran_minor(int w, int h, int L, const uint8_t *exact_tbl)
{
	const long long a = min(w, h), b = max(w, h);
	const long long num = (long long)L * L * a;
	int m = 0;
	while ((long long)(m + 1) * (m + 1) * b <= num) m++;
	if ((long long)m * m * b == num) return exact_tbl[m];
	return m;
}

__device__ __forceinline__ int src_px(const PlaneSrc &ps, int pitch, int x, int y)
{
	const int v = __ldg(ps.src + (size_t)y * pitch + x);
	return ps.invert ? 255 - v : v;
}

constexpr int HIST_WARPS = 8;

__global__ void __launch_bounds__(HIST_WARPS * 32) k_lbp_hist(ClassifyParams P, const PlaneSrc *__restrict__ planes,
                                                               const OutNode *__restrict__ nodes, const int32_t *__restrict__ pool,
                                                               const int32_t *__restrict__ counts, const uint8_t *__restrict__ aran_tbl,
                                                               uint8_t *__restrict__ hist_out, uint8_t *__restrict__ codes_out)
{
	__shared__ uint8_t s_patch[HIST_WARPS][26 * 26 + 28];
	__shared__ uint32_t s_hist[HIST_WARPS][1024];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int plane = blockIdx.y;
	const int r = blockIdx.x * HIST_WARPS + warp;
	const int npool = min(counts[2 * plane + 1], P.pool_cap);
	if (r >= npool) return;
	const OutNode nd = nodes[(size_t)plane * P.node_cap + pool[(size_t)plane * P.pool_cap + r]];
	const PlaneSrc ps = planes[plane];
	uint8_t *patch = s_patch[warp];
	uint32_t *hist = s_hist[warp];
	for (int i = lane; i < 26 * 26 + 28; i += 32) patch[i] = 0;
	for (int i = lane; i < 1024; i += 32) hist[i] = 0;
	__syncwarp();

	const int L = 26;
	const int sw = nd.w, sh = nd.h;
	const int minor = aran_minor(sw, sh, L, aran_tbl);
	const int dw = (sw > sh) ? L : minor, dh = (sw > sh) ? minor : L;
	// paste offset: (int)round((L - minor) / 2) with integer division first (src/OCR.cpp:405,418)
	const int off = (L - minor) / 2;
	const int offx = (dw > dh) ? 0 : off, offy = (dw > dh) ? off : 0;
	if (dw > 0 && dh > 0) {
		const int X0 = nd.x, Y0 = nd.y;
		if (sw == 2 * dw && sh == 2 * dh) {
			for (int t = lane; t < dw * dh; t += 32) {
				const int oy = t / dw, ox = t % dw;
				const int s = src_px(ps, P.pitch, X0 + 2 * ox, Y0 + 2 * oy) + src_px(ps, P.pitch, X0 + 2 * ox + 1, Y0 + 2 * oy) +
				              src_px(ps, P.pitch, X0 + 2 * ox, Y0 + 2 * oy + 1) + src_px(ps, P.pitch, X0 + 2 * ox + 1, Y0 + 2 * oy + 1);
				patch[(oy + offy) * L + ox + offx] = (uint8_t)((s + 2) >> 2);
			}
		} else {
			const double scx = 1.0 / ((double)dw / (double)sw), scy = 1.0 / ((double)dh / (double)sh);
			for (int t = lane; t < dw * dh; t += 32) {
				const int oy = t / dw, ox = t % dw;
				float fy = (float)(((double)oy + 0.5) * scy - 0.5);
				int iy = (int)floorf(fy);
				fy -= (float)iy;
				const int wy0 = __float2int_rn((1.f - fy) * 2048.f), wy1 = __float2int_rn(fy * 2048.f);
				const int r0 = min(max(iy, 0), sh - 1), r1 = min(max(iy + 1, 0), sh - 1);
				float fx = (float)(((double)ox + 0.5) * scx - 0.5);
				int ix = (int)floorf(fx);
				fx -= (float)ix;
				if (ix < 0) { ix = 0; fx = 0.f; }
				if (ix >= sw - 1) { ix = sw - 1; fx = 0.f; }
				const int wx0 = __float2int_rn((1.f - fx) * 2048.f), wx1 = __float2int_rn(fx * 2048.f);
				const int ix1 = min(ix + 1, sw - 1);
				const int h0 = src_px(ps, P.pitch, X0 + ix, Y0 + r0) * wx0 + src_px(ps, P.pitch, X0 + ix1, Y0 + r0) * wx1;
				const int h1 = src_px(ps, P.pitch, X0 + ix, Y0 + r1) * wx0 + src_px(ps, P.pitch, X0 + ix1, Y0 + r1) * wx1;
				int v = (((wy0 * (h0 >> 4)) >> 16) + ((wy1 * (h1 >> 4)) >> 16) + 2) >> 2;
				v = min(max(v, 0), 255);
				patch[(oy + offy) * L + ox + offx] = (uint8_t)v;
			}
		}
	}
	__syncwarp();
	// mean-LBP with the reference's offsets relative to base = (i+1)*26 + 1 + j  (src/ER.cpp:831-841)
	for (int t = lane; t < 24 * 24; t += 32) {
		const int i = t / 24, j = t % 24;
		const int base = (i + 1) * 26 + 1 + j;
		const int v0 = patch[base - 25], v1 = patch[base - 24], v2 = patch[base - 23], v3 = patch[base + 1];
		const int v4 = patch[base + 25], v5 = patch[base + 24], v6 = patch[base + 23], v7 = patch[base - 1];
		const int s = v0 + v1 + v2 + v3 + v4 + v5 + v6 + v7;   // bit set iff v > s/8.0  <=>  8v > s
		const int code = (8 * v0 > s) | ((8 * v1 > s) << 1) | ((8 * v2 > s) << 2) | ((8 * v3 > s) << 3) |
		                 ((8 * v4 > s) << 4) | ((8 * v5 > s) << 5) | ((8 * v6 > s) << 6) | ((8 * v7 > s) << 7);
		atomicAdd(&hist[(i / 12) * 512 + (j / 12) * 256 + code], 1u);
		if (codes_out) codes_out[((size_t)plane * P.pool_cap + r) * 576 + t] = (uint8_t)code;   // calc_LBP's 24x24 image (src/ER.cpp:819-845)
	}
	__syncwarp();
	uint32_t *out = reinterpret_cast<uint32_t *>(hist_out + ((size_t)plane * P.pool_cap + r) * 1024);
	for (int it = 0; it < 8; it++) {
		const int b = (it * 32 + lane) * 4;
		out[it * 32 + lane] = hist[b] | (hist[b + 1] << 8) | (hist[b + 2] << 16) | (hist[b + 3] << 24);
	}
}

// ---------------------------------------------------------------------------------------------
// cascades
// ---------------------------------------------------------------------------------------------
#define ERT_NEG_DBL_MAX (-1.7976931348623157e308)

template <typename T>
__device__ __forceinline__ double cascade_eval(const CascadeDev &c, const T *__restrict__ fv)
{
	double score = 0.0;
	int off = 0;
	for (int s = 0; s < c.n_stages; s++) {
		score = 0.0;
		const int len = c.stage_len[s];
		for (int j = off; j < off + len; j++) {
			const Stump st = c.stumps[j];
			const double f = (double)fv[st.dim];
			score = __dadd_rn(score, (f < st.thr) ? st.cp : st.cn);
		}
		if (score < (double)c.stage_thr[s]) return ERT_NEG_DBL_MAX;
		off += len;
	}
	return score;
}

// pipeline flavour: regions are (plane, k) with k < counts[2*plane+1]; API flavour: counts == nullptr, n rows.
template <typename T>
__global__ void k_cascade(const T *__restrict__ fv, size_t row_stride, int n_rows, const int32_t *__restrict__ counts, int pool_cap,
                          CascadeDev strong, CascadeDev weak, int32_t *__restrict__ label, double *__restrict__ sscore,
                          double *__restrict__ wscore)
{
	const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= n_rows) return;
	if (counts) {
		const int plane = (int)(idx / pool_cap), k = (int)(idx % pool_cap);
		if (k >= min(counts[2 * plane + 1], pool_cap)) return;
	}
	const T *row = fv + (size_t)idx * row_stride;
	const double s = cascade_eval<T>(strong, row);
	const double w = cascade_eval<T>(weak, row);
	label[idx] = (s > ERT_NEG_DBL_MAX) ? 2 : ((w > ERT_NEG_DBL_MAX) ? 1 : 0);
	if (sscore) sscore[idx] = s;
	if (wscore) wscore[idx] = w;
}

// ---------------------------------------------------------------------------------------------
// Warp-cooperative stage evaluation for u8 histograms.  Stump tables live in shared memory in a compact
// form (dim, integer threshold, cp, cn: for integer counts h < thr <=> h < ceil(thr)); 32 lanes gather and select
// 32 stumps at once, then the stage sum is accumulated in FILE ORDER by broadcasting the 32 selected values one
// after the other (every lane performs the same non-fused double adds) -- bit-identical stage sums, ~20x less
// latency than one thread walking the table.
// ---------------------------------------------------------------------------------------------
struct StumpC { double cp, cn; uint16_t dim, ithr; uint32_t pad; };

// ---------------------------------------------------------------------------------------------
// k_cascade_stage : one CTA per region, one WARP per cascade STAGE.  A stage's sum starts from zero
// (src/adaboost.cpp:528-529), so the stages of both cascades are independent computations: each warp evaluates one
// stage as described above (32 stumps gathered at once, added in FILE ORDER with non-fused double adds),
// and one thread then applies the stage thresholds in order.  Same bits as walking the stages one after the other,
// but the latency is the longest stage (1210 stumps) instead of the sum of the stages a region survives (up to 4014).
// Persistent CTAs: the stump tables are loaded into shared memory once per CTA.
// ---------------------------------------------------------------------------------------------
constexpr int CS_WARPS = 12;

__global__ void __launch_bounds__(CS_WARPS * 32) k_cascade_stage(const uint8_t *__restrict__ hist, int n_rows, const int32_t *__restrict__ counts,
                                                                 int n_planes, int pool_cap, CascadeDev strong, CascadeDev weak, int n_strong,
                                                                 int n_weak, int32_t *__restrict__ label, double *__restrict__ sscore,
                                                                 double *__restrict__ wscore)
{
	extern __shared__ __align__(16) uint8_t csm[];
	StumpC *st = reinterpret_cast<StumpC *>(csm);                 // strong stumps, then weak stumps
	double *s_score = reinterpret_cast<double *>(st + n_strong + n_weak);   // [32] stage sums of the current region
	int *meta = reinterpret_cast<int *>(s_score + 32);            // [0..15] strong len, [16..31] strong thr, [32..47] weak len, [48..63] weak thr
	int *stage_off = meta + 64;                                   // [32] first stump of stage s (strong stages, then weak stages)
	uint8_t *hs = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(stage_off + 32) + 15) & ~(uintptr_t)15);   // 1024 histogram bytes (uint4 copies)
	int *pref = reinterpret_cast<int *>(hs + 1024);               // [n_planes + 1] first region of each plane (pipeline flavour)
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	for (int i = tid; i < n_strong + n_weak; i += blockDim.x) {
		const Stump s0 = (i < n_strong) ? strong.stumps[i] : weak.stumps[i - n_strong];
		StumpC c;
		c.cp = s0.cp; c.cn = s0.cn; c.dim = (uint16_t)s0.dim;
		const double ct = ceil(s0.thr);
		c.ithr = (uint16_t)(ct < 0.0 ? 0.0 : (ct > 65535.0 ? 65535.0 : ct));
		c.pad = 0;
		st[i] = c;
	}
	if (tid < 16) {
		meta[tid] = (tid < strong.n_stages) ? strong.stage_len[tid] : 0;
		meta[16 + tid] = (tid < strong.n_stages) ? strong.stage_thr[tid] : 0;
		meta[32 + tid] = (tid < weak.n_stages) ? weak.stage_len[tid] : 0;
		meta[48 + tid] = (tid < weak.n_stages) ? weak.stage_thr[tid] : 0;
	}
	__syncthreads();
	const int ns = strong.n_stages, nw = weak.n_stages, nst = ns + nw;
	if (tid == 0) {
		int o = 0;
		for (int s2 = 0; s2 < ns; s2++) { stage_off[s2] = o; o += meta[s2]; }
		o = n_strong;
		for (int s2 = 0; s2 < nw; s2++) { stage_off[ns + s2] = o; o += meta[32 + s2]; }
		if (counts) {
			int acc = 0;
			for (int p = 0; p < n_planes; p++) { pref[p] = acc; acc += min(counts[2 * p + 1], pool_cap); }
			pref[n_planes] = acc;
		}
	}
	__syncthreads();
	const int total = counts ? pref[n_planes] : n_rows;
	for (int r = blockIdx.x; r < total; r += gridDim.x) {
		size_t idx = (size_t)r;
		if (counts) {
			int lo = 0, hi = n_planes - 1;                             // largest plane with pref[plane] <= r
			while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (pref[mid] <= r) lo = mid; else hi = mid - 1; }
			idx = (size_t)lo * pool_cap + (size_t)(r - pref[lo]);
		}
		__syncthreads();                                               // the previous region's readers are done with hs / s_score
		if (tid < 64) reinterpret_cast<uint4 *>(hs)[tid] = reinterpret_cast<const uint4 *>(hist + idx * 1024)[tid];
		__syncthreads();
		for (int s2 = warp; s2 < nst; s2 += CS_WARPS) {
			const StumpC *tbl = st + stage_off[s2];
			const int len = (s2 < ns) ? meta[s2] : meta[32 + s2 - ns];
			double score = 0.0;
			for (int c0 = 0; c0 < len; c0 += 32) {
				const int j = c0 + lane;
				double v = 0.0;
				if (j < len) { const StumpC t = tbl[j]; v = ((int)hs[t.dim] < (int)t.ithr) ? t.cp : t.cn; }
				const int m = min(32, len - c0);
				for (int i = 0; i < m; i++) score = __dadd_rn(score, __shfl_sync(0xFFFFFFFFu, v, i));
			}
			if (lane == 0) s_score[s2] = score;
		}
		__syncthreads();
		if (tid == 0) {
			double sres = 0.0, wres = 0.0;                              // CascadeBoost::predict: the last stage's sum, -DBL_MAX on the first failing stage
			for (int s2 = 0; s2 < ns; s2++) { sres = s_score[s2]; if (sres < (double)meta[16 + s2]) { sres = ERT_NEG_DBL_MAX; break; } }
			for (int s2 = 0; s2 < nw; s2++) { wres = s_score[ns + s2]; if (wres < (double)meta[48 + s2]) { wres = ERT_NEG_DBL_MAX; break; } }
			label[idx] = (sres > ERT_NEG_DBL_MAX) ? 2 : ((wres > ERT_NEG_DBL_MAX) ? 1 : 0);
			if (sscore) sscore[idx] = sres;
			if (wscore) wscore[idx] = wres;
		}
	}
}

int launch_lbp_hist(const ClassifyParams &P, int n_planes, const PlaneSrc *planes, const OutNode *nodes, const int32_t *pool,
                    const int32_t *counts, const uint8_t *aran_tbl, uint8_t *hist_out, cudaStream_t st, uint8_t *codes_out)
{
	dim3 grid((P.pool_cap + HIST_WARPS - 1) / HIST_WARPS, n_planes);
	k_lbp_hist<<<grid, HIST_WARPS * 32, 0, st>>>(P, planes, nodes, pool, counts, aran_tbl, hist_out, codes_out);
	ERT_CUDA_CHECK(cudaGetLastError());
	return 0;
}

int launch_cascade_u8(const uint8_t *hist, size_t row_stride, int n_rows, const int32_t *counts, int pool_cap, const CascadeDev &strong,
                      const CascadeDev &weak, int n_strong, int n_weak, int32_t *label, double *sscore, double *wscore, cudaStream_t st)
{
	if (n_rows <= 0) return 0;
	const int n_planes = counts ? n_rows / pool_cap : 0;
	const size_t smem = (size_t)(n_strong + n_weak) * sizeof(StumpC) + 32 * sizeof(double) + 96 * sizeof(int) + 16 + 1024 + (size_t)(n_planes + 1) * sizeof(int);
	// few regions (the pipeline: tens per plane): one CTA per region, one warp per cascade stage -- the latency is the
	// longest stage; very many regions (candidate sweeps): one thread per region keeps 32 independent sums per warp in flight
	const bool few = (counts != nullptr) || n_rows < 32768;
	if (few && row_stride == 1024 && strong.n_stages <= 16 && weak.n_stages <= 16 && smem <= 200 * 1024) {
		ERT_CUDA_CHECK(cudaFuncSetAttribute(k_cascade_stage, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
		const int grid = counts ? 148 * 2 : (int)std::min<long long>(n_rows, 148 * 2);
		k_cascade_stage<<<grid, CS_WARPS * 32, smem, st>>>(hist, n_rows, counts, n_planes, pool_cap, strong, weak, n_strong, n_weak, label, sscore, wscore);
	} else {
		k_cascade<uint8_t><<<(n_rows + 127) / 128, 128, 0, st>>>(hist, row_stride, n_rows, counts, pool_cap, strong, weak, label, sscore, wscore);
	}
	ERT_CUDA_CHECK(cudaGetLastError());
	return 0;
}

int launch_cascade_f64(const double *fv, size_t row_stride, int n_rows, const CascadeDev &strong, const CascadeDev &weak,
                       int32_t *label, double *sscore, double *wscore, cudaStream_t st)
{
	if (n_rows <= 0) return 0;
	k_cascade<double><<<(n_rows + 127) / 128, 128, 0, st>>>(fv, row_stride, n_rows, nullptr, 1, strong, weak, label, sscore, wscore);
	ERT_CUDA_CHECK(cudaGetLastError());
	return 0;
}

} // namespace ert
