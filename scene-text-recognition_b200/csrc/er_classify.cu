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

constexpr int HIST_WARPS = 4;
constexpr int HIST_GRID = 148 * 4;

// largest plane with pool_prefix[plane] <= r
__device__ __forceinline__ int plane_of_region(const int32_t *__restrict__ pool_prefix, int n_planes, int r)
{
	int lo = 0, hi = n_planes - 1;
	while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (__ldg(pool_prefix + mid) <= r) lo = mid; else hi = mid - 1; }
	return lo;
}

__global__ void k_pool_prefix(const int32_t *__restrict__ counts, int n_planes, int pool_cap, int32_t *__restrict__ pool_prefix)
{
	const int lane = threadIdx.x;
	int carry = 0;
	for (int p0 = 0; p0 < n_planes; p0 += 32) {
		const int p = p0 + lane;
		const int c = (p < n_planes) ? min(counts[2 * p + 1], pool_cap) : 0;
		int inc = c;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (lane >= o) inc += t; }
		if (p < n_planes) pool_prefix[p] = carry + inc - c;
		carry += __shfl_sync(0xFFFFFFFFu, inc, 31);
	}
	if (lane == 0) pool_prefix[n_planes] = carry;
}

// One WARP per pooled region, warps stride over all regions of the batch (found through pool_prefix): the grid does not
// depend on the per-plane capacity, and a CTA (4 warps, 7 KB of shared memory, histogram bins packed four to a word --
// a bin never exceeds the 144 pixels of its block) fits beside the resident tile-kernel CTAs of other batches.
__global__ void __launch_bounds__(HIST_WARPS * 32) k_lbp_hist(ClassifyParams P, int n_planes, const PlaneSrc *__restrict__ planes,
                                                               const OutNode *__restrict__ nodes, const int32_t *__restrict__ pool,
                                                               const int32_t *__restrict__ pool_prefix, const uint8_t *__restrict__ aran_tbl,
                                                               uint8_t *__restrict__ hist_out, uint8_t *__restrict__ codes_out)
{
	__shared__ uint8_t s_patch[HIST_WARPS][26 * 26 + 28];
	__shared__ uint32_t s_hist[HIST_WARPS][256];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int total = pool_prefix[n_planes];
	uint8_t *patch = s_patch[warp];
	uint32_t *hist = s_hist[warp];
	for (int item = blockIdx.x * HIST_WARPS + warp; item < total; item += gridDim.x * HIST_WARPS) {
	const int plane = plane_of_region(pool_prefix, n_planes, item);
	const int r = item - pool_prefix[plane];
	const OutNode nd = nodes[(size_t)plane * P.node_cap + pool[(size_t)plane * P.pool_cap + r]];
	const PlaneSrc ps = planes[plane];
	for (int i = lane; i < 26 * 26 + 28; i += 32) patch[i] = 0;
	for (int i = lane; i < 256; i += 32) hist[i] = 0;
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
		const int bin = (i / 12) * 512 + (j / 12) * 256 + code;
		atomicAdd(&hist[bin >> 2], 1u << (8 * (bin & 3)));
		if (codes_out) codes_out[((size_t)plane * P.pool_cap + r) * 576 + t] = (uint8_t)code;   // calc_LBP's 24x24 image (src/ER.cpp:819-845)
	}
	__syncwarp();
	uint32_t *out = reinterpret_cast<uint32_t *>(hist_out + ((size_t)plane * P.pool_cap + r) * 1024);
	for (int it = 0; it < 8; it++) out[it * 32 + lane] = hist[it * 32 + lane];     // bin 4w + b is byte b of word w: the u8[1024] layout
	__syncwarp();
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
// k_cascade_stage : one WARP per (region, cascade STAGE).  A stage's sum starts from zero (src/adaboost.cpp:528-529), so
// the 4 + 6 stages of the two cascades are independent computations: a warp evaluates one stage -- 32 lanes gather and
// select 32 stumps at once (compact tables {cp, cn}, dim | ceil(thr) << 16 read coalesced through L2; for integer counts
// h < thr <=> h < ceil(thr)), then the stage sum is accumulated in FILE ORDER by broadcasting the 32 selected values one
// after the other with non-fused double adds: bit-identical stage sums.  The warp that delivers a region's last stage
// applies the stage thresholds in order (CascadeBoost::predict: the last stage's sum, -DBL_MAX at the first failing
// stage).  Latency = the longest stage (1210 stumps), not the sum of the stages a region survives (up to 4014); no
// shared-memory tables (the round-1 form staged 96 KB per CTA, which cannot sit beside the tile kernel's CTAs).
// ---------------------------------------------------------------------------------------------
constexpr int CS_WARPS = 4;
constexpr int CS_GRID = 148 * 4;

__global__ void __launch_bounds__(CS_WARPS * 32) k_cascade_stage(const uint8_t *__restrict__ hist, int n_rows, const int32_t *__restrict__ pool_prefix,
                                                                 int n_planes, int pool_cap, CascadeDev strong, CascadeDev weak,
                                                                 int32_t *__restrict__ label, double *__restrict__ sscore, double *__restrict__ wscore,
                                                                 double *__restrict__ stage_sum, uint32_t *__restrict__ done)
{
	__shared__ double s_v[CS_WARPS][32];            // the 32 selected values of a chunk, re-read by broadcast
	const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
	const int gw = blockIdx.x * CS_WARPS + wib, nw_total = gridDim.x * CS_WARPS;
	const int ns = strong.n_stages, nw = weak.n_stages, nst = ns + nw;
	const int total = pool_prefix ? pool_prefix[n_planes] : n_rows;
	const long long items = (long long)total * nst;
	double *sv = s_v[wib];
	for (long long it = gw; it < items; it += nw_total) {
		__syncwarp();                                 // lane 0 may still be in the previous item's epilogue
		const int r = (int)(it / nst), s2 = (int)(it % nst);
		size_t idx = (size_t)r;
		if (pool_prefix) {
			const int plane = plane_of_region(pool_prefix, n_planes, r);
			idx = (size_t)plane * pool_cap + (size_t)(r - pool_prefix[plane]);
		}
		const bool in_strong = s2 < ns;
		const int sl = in_strong ? s2 : s2 - ns;
		const int *slen = in_strong ? strong.stage_len : weak.stage_len;
		const double2 *cpcn = in_strong ? strong.cpcn : weak.cpcn;
		const uint32_t *dimthr = in_strong ? strong.dimthr : weak.dimthr;
		int off = 0;
		for (int q = 0; q < sl; q++) off += slen[q];
		const int len = slen[sl];
		const uint8_t *hs = hist + idx * 1024;
		double score = 0.0;
		// All lanes execute the same instruction stream (indices clamped, values selected): the warp stays converged, the
		// next chunk's table entries are in flight while this chunk's 32 adds retire, and the 32 selected values travel
		// through shared memory (one store, broadcast reads) so the add chain is the only dependent sequence.
		int jc = max(min(lane, len - 1), 0);
		double2 cc = make_double2(0.0, 0.0);
		uint32_t dt = 0;
		if (len > 0) { cc = __ldg(cpcn + off + jc); dt = __ldg(dimthr + off + jc); }      // warp-uniform condition
		for (int c0 = 0; c0 < len; c0 += 32) {
			const int j = c0 + lane;
			const uint32_t h = (uint32_t)__ldg(hs + (dt & 0xFFFFu));
			const double v = (j < len) ? ((h < (dt >> 16)) ? cc.x : cc.y) : 0.0;
			jc = min(j + 32, len - 1);
			cc = __ldg(cpcn + off + jc);
			dt = __ldg(dimthr + off + jc);
			sv[lane] = v;
			__syncwarp();
			const int m = min(32, len - c0);
			if (m == 32) {
#pragma unroll
				for (int i = 0; i < 32; i += 2) {
					const double2 p = *reinterpret_cast<const double2 *>(sv + i);
					score = __dadd_rn(score, p.x);
					score = __dadd_rn(score, p.y);
				}
			} else {
				for (int i = 0; i < m; i++) score = __dadd_rn(score, sv[i]);
			}
			__syncwarp();
		}
		uint32_t arrived = 0;
		if (lane == 0) {
			stage_sum[(size_t)r * 32 + s2] = score;
			__threadfence();
			arrived = atomicAdd(&done[r], 1u);
		}
		arrived = __shfl_sync(0xFFFFFFFFu, arrived, 0);
		if (arrived == (uint32_t)(nst - 1) && lane == 0) {
			__threadfence();
			const volatile double *ss = stage_sum + (size_t)r * 32;
			double sres = 0.0, wres = 0.0;
			for (int q = 0; q < ns; q++) { sres = ss[q]; if (sres < (double)strong.stage_thr[q]) { sres = ERT_NEG_DBL_MAX; break; } }
			for (int q = 0; q < nw; q++) { wres = ss[ns + q]; if (wres < (double)weak.stage_thr[q]) { wres = ERT_NEG_DBL_MAX; break; } }
			label[idx] = (sres > ERT_NEG_DBL_MAX) ? 2 : ((wres > ERT_NEG_DBL_MAX) ? 1 : 0);
			if (sscore) sscore[idx] = sres;
			if (wscore) wscore[idx] = wres;
			done[r] = 0;                                   // ready for the next launch
		}
	}
}

int launch_pool_prefix(const int32_t *counts, int n_planes, int pool_cap, int32_t *pool_prefix, cudaStream_t st)
{
	k_pool_prefix<<<1, 32, 0, st>>>(counts, n_planes, pool_cap, pool_prefix);
	ERT_CUDA_CHECK(cudaGetLastError());
	return 0;
}

int launch_lbp_hist(const ClassifyParams &P, int n_planes, const PlaneSrc *planes, const OutNode *nodes, const int32_t *pool,
                    const int32_t *pool_prefix, const uint8_t *aran_tbl, uint8_t *hist_out, cudaStream_t st, uint8_t *codes_out)
{
	k_lbp_hist<<<HIST_GRID, HIST_WARPS * 32, 0, st>>>(P, n_planes, planes, nodes, pool, pool_prefix, aran_tbl, hist_out, codes_out);
	ERT_CUDA_CHECK(cudaGetLastError());
	return 0;
}

int launch_cascade_u8(const uint8_t *hist, size_t row_stride, int n_rows, const int32_t *pool_prefix, int n_planes, int pool_cap, const CascadeDev &strong,
                      const CascadeDev &weak, int32_t *label, double *sscore, double *wscore, const CascadeScratch &sc, cudaStream_t st)
{
	if (n_rows <= 0) return 0;
	// few regions (the pipeline: tens per plane): one warp per (region, stage) -- the latency is the longest stage; very many
	// regions (candidate sweeps): one thread per region keeps 32 independent sums per warp in flight
	const bool few = (pool_prefix != nullptr) || n_rows < 32768;
	if (few && row_stride == 1024 && strong.n_stages + weak.n_stages <= 32 && strong.cpcn && weak.cpcn && sc.stage_sum && n_rows <= sc.rows_cap) {
		const long long items = (long long)n_rows * (strong.n_stages + weak.n_stages);
		const int grid = pool_prefix ? CS_GRID : (int)std::min<long long>((items + CS_WARPS - 1) / CS_WARPS, CS_GRID);
		k_cascade_stage<<<grid, CS_WARPS * 32, 0, st>>>(hist, n_rows, pool_prefix, n_planes, pool_cap, strong, weak, label, sscore, wscore, sc.stage_sum, sc.done);
	} else {
		if (pool_prefix) { set_error("launch_cascade_u8: the per-plane form needs the cascade scratch"); return -1; }
		k_cascade<uint8_t><<<(n_rows + 127) / 128, 128, 0, st>>>(hist, row_stride, n_rows, nullptr, pool_cap, strong, weak, label, sscore, wscore);
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
