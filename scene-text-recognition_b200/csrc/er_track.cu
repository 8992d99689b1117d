// er_track.cu -- the step right after classify: ERFilter::er_track + calc_color
// (reference src/ER.cpp:532-609, 1391-1437), on the device, for every frame of a batch at once.
//
//   k_track_gather  strong[0..5] then weak[0..5] of a frame, each in pool order -> one candidate list per frame
//   k_calc_color    per candidate: per-warp histograms of 255 - channel over the bound, OpenCV's OTSU threshold (the FP64
//                   recurrence of getThreshVal_Otsu_8u evaluated in the same order, no FMA contraction), then the
//                   mean YCrCb over the mask -- read at (row, col) counted from the IMAGE origin, not the bound's,
//                   because that is what the reference does (color_img.ptr(i), src/ER.cpp:1404)
//   k_track         one warp per frame runs the strong-seeded growth of all_er: the outer loop over the growing
//                   list is sequential by definition, the inner scan over the weak regions is a ballot + ordered
//                   append, which reproduces the (m, n) loop order exactly
//   k_track_emit    per-frame lists -> contiguous host-mapped arrays
// Integer work except the OTSU recurrence and the colour means / comparisons, which are IEEE double operations
// in the reference's order => results are bit-identical to the CPU code.
#include "kernels.h"

namespace ert {

namespace {

constexpr int CC_THREADS = 256;
constexpr int CC_WARPS = CC_THREADS / 32;

__global__ void __launch_bounds__(32) k_track_gather(TrackWork tk, const OutNode *__restrict__ nodes, const int32_t *__restrict__ pool,
                                                     const int32_t *__restrict__ counts, const int32_t *__restrict__ label, int node_cap,
                                                     int pool_cap)
{
	const int f = blockIdx.x, lane = threadIdx.x;
	ert_tracked *cand = tk.cand + (size_t)f * tk.cand_cap;
	int n = 0, ns = 0;
	for (int pass = 0; pass < 2; pass++) {
		const int want = pass == 0 ? ERT_LABEL_STRONG : ERT_LABEL_WEAK;
		for (int ch = 0; ch < 6; ch++) {
			const int p = f * 6 + ch;
			const int np = counts[2 * p + 1];
			for (int base = 0; base < np; base += 32) {
				const int i = base + lane;
				bool ok = false;
				int node = 0;
				if (i < np) { ok = label[(size_t)p * pool_cap + i] == want; node = pool[(size_t)p * pool_cap + i]; }
				const unsigned m = __ballot_sync(0xffffffffu, ok);
				if (ok) {
					const int pos = n + __popc(m & ((1u << lane) - 1u));
					if (pos < tk.cand_cap) {
						const OutNode nd = nodes[(size_t)p * node_cap + node];
						ert_tracked t;
						t.plane = ch; t.pool_index = i; t.node = node; t.label = want; t.level = nd.level; t.area = nd.area;
						t.x = nd.x; t.y = nd.y; t.w = nd.w; t.h = nd.h;
						t.center_x = nd.x + nd.w / 2; t.center_y = nd.y + nd.h / 2;     // src/ER.cpp:545
						t.color1 = t.color2 = t.color3 = 0.0;
						cand[pos] = t;
					}
				}
				n += __popc(m);
			}
		}
		if (pass == 0) ns = n;
	}
	if (lane == 0) { tk.n_cand[f] = min(n, tk.cand_cap); tk.n_strong[f] = min(ns, tk.cand_cap); }
}

// OpenCV getThreshVal_Otsu_8u (modules/imgproc/src/thresh.cpp) as the reference reaches it through
// threshold(255 - img, img, 128, 255, THRESH_OTSU) (src/ER.cpp:1395): same operations, same order, round-to-nearest
// doubles without contraction.
__device__ int otsu_threshold(const int *h, int total)
{
	double mu = 0.0;
	const double scale = __ddiv_rn(1.0, (double)total);
	for (int i = 0; i < 256; i++) mu = __dadd_rn(mu, __dmul_rn((double)i, (double)h[i]));
	mu = __dmul_rn(mu, scale);
	double mu1 = 0.0, q1 = 0.0, max_sigma = 0.0;
	int max_val = 0;
	const double feps = 1.1920928955078125e-07;
	for (int i = 0; i < 256; i++) {
		const double p_i = __dmul_rn((double)h[i], scale);
		mu1 = __dmul_rn(mu1, q1);
		q1 = __dadd_rn(q1, p_i);
		const double q2 = __dsub_rn(1.0, q1);
		if (fmin(q1, q2) < feps || fmax(q1, q2) > 1.0 - feps) continue;
		mu1 = __ddiv_rn(__dadd_rn(mu1, __dmul_rn((double)i, p_i)), q1);
		const double mu2 = __ddiv_rn(__dsub_rn(mu, __dmul_rn(q1, mu1)), q2);
		const double d = __dsub_rn(mu1, mu2);
		const double sigma = __dmul_rn(__dmul_rn(__dmul_rn(q1, q2), d), d);
		if (sigma > max_sigma) { max_sigma = sigma; max_val = i; }
	}
	return max_val;
}

__global__ void __launch_bounds__(CC_THREADS) k_calc_color(TrackWork tk, const uint8_t *__restrict__ d_ycc, size_t plane_bytes, int pitch)
{
	__shared__ int hist[CC_WARPS][256];
	__shared__ int s_thr;
	__shared__ unsigned long long red[CC_WARPS][4];
	const int f = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int n = tk.n_cand[f];
	ert_tracked *cand = tk.cand + (size_t)f * tk.cand_cap;
	const uint8_t *Y = d_ycc + (size_t)f * 3 * plane_bytes, *Cr = Y + plane_bytes, *Cb = Cr + plane_bytes;
	for (int ci = blockIdx.x; ci < n; ci += gridDim.x) {
		const int ch = cand[ci].plane, x0 = cand[ci].x, y0 = cand[ci].y, w = cand[ci].w, h = cand[ci].h;
		const uint8_t *src = Y + (size_t)(ch % 3) * plane_bytes;
		const bool inv = ch >= 3;                 // channel value = 255 - plane; the mask image is 255 - channel
		for (int i = tid; i < CC_WARPS * 256; i += CC_THREADS) (&hist[0][0])[i] = 0;
		__syncthreads();
		for (int r = warp; r < h; r += CC_WARPS) {
			const uint8_t *row = src + (size_t)(y0 + r) * pitch + x0;
			for (int xb = 0; xb < w; xb += 32) {
				const int x = xb + lane;
				const bool in = x < w;
				int u = 0;
				if (in) { const int v = row[x]; u = inv ? v : 255 - v; }
				// per-warp histogram; same-bin lanes serialise in the shared-memory atomic unit (measured faster than aggregating
				// equal values with __match_any_sync first: 0.265 -> 0.224 ms per 8-frame batch)
				if (in) atomicAdd(&hist[warp][u], 1);
			}
		}
		__syncthreads();
		for (int i = tid; i < 256; i += CC_THREADS) {
			int s = 0;
#pragma unroll
			for (int k = 0; k < CC_WARPS; k++) s += hist[k][i];
			hist[0][i] = s;
		}
		__syncthreads();
		if (tid == 0) s_thr = otsu_threshold(hist[0], w * h);
		__syncthreads();
		const int thr = s_thr;
		unsigned long long cnt = 0, s1 = 0, s2 = 0, s3 = 0;
		for (int r = warp; r < h; r += CC_WARPS) {
			const uint8_t *row = src + (size_t)(y0 + r) * pitch + x0;
			// colour rows / columns are counted from the image origin (src/ER.cpp:1403-1404), not from the bound
			const uint8_t *ry = Y + (size_t)r * pitch, *rcr = Cr + (size_t)r * pitch, *rcb = Cb + (size_t)r * pitch;
			for (int x = lane; x < w; x += 32) {
				const int v = row[x];
				const int u = inv ? v : 255 - v;
				if (u > thr) { cnt++; s1 += ry[x]; s2 += rcr[x]; s3 += rcb[x]; }
			}
		}
#pragma unroll
		for (int o = 16; o; o >>= 1) {
			cnt += __shfl_down_sync(0xffffffffu, cnt, o); s1 += __shfl_down_sync(0xffffffffu, s1, o);
			s2 += __shfl_down_sync(0xffffffffu, s2, o); s3 += __shfl_down_sync(0xffffffffu, s3, o);
		}
		if (lane == 0) { red[warp][0] = cnt; red[warp][1] = s1; red[warp][2] = s2; red[warp][3] = s3; }
		__syncthreads();
		if (tid == 0) {
			unsigned long long c = 0, a = 0, b = 0, d = 0;
			for (int k = 0; k < CC_WARPS; k++) { c += red[k][0]; a += red[k][1]; b += red[k][2]; d += red[k][3]; }
			cand[ci].color1 = __ddiv_rn((double)a, (double)c);      // 0/0 = NaN, exactly as the reference
			cand[ci].color2 = __ddiv_rn((double)b, (double)c);
			cand[ci].color3 = __ddiv_rn((double)d, (double)c);
		}
		__syncthreads();
	}
}

// the acceptance test of src/ER.cpp:579-590 (USE_STROKE_WIDTH is off, inc/ER.h:25)
__device__ __forceinline__ bool track_match(const ert_tracked &s, const ert_tracked &w)
{
	return abs(s.center_x - w.center_x) + abs(s.center_y - w.center_y) < (max(s.w, s.h) << 1) &&
	       abs(s.h - w.h) < min(s.h, w.h) &&
	       abs(s.w - w.w) < ((s.w + w.w) >> 1) &&
	       fabs(__dsub_rn(s.color1, w.color1)) < 25.0 &&
	       fabs(__dsub_rn(s.color2, w.color2)) < 25.0 &&
	       fabs(__dsub_rn(s.color3, w.color3)) < 25.0 &&
	       abs(s.area - w.area) < min(s.area, w.area) * 3;
}

__global__ void __launch_bounds__(32) k_track(TrackWork tk)
{
	extern __shared__ unsigned s_flags[];          // one bit per weak candidate: already in all_er
	const int f = blockIdx.x, lane = threadIdx.x;
	const int n = tk.n_cand[f], ns = tk.n_strong[f];
	const ert_tracked *cand = tk.cand + (size_t)f * tk.cand_cap;
	int32_t *tracked = tk.tracked + (size_t)f * tk.cand_cap;
	const int words = (n - ns + 31) >> 5;
	for (int i = lane; i < words; i += 32) s_flags[i] = 0;
	for (int i = lane; i < ns; i += 32) tracked[i] = i;          // all_er starts as strong[0] ++ ... ++ strong[5]
	__syncwarp();
	int qlen = ns;
	for (int qi = 0; qi < qlen; qi++) {
		const ert_tracked s = cand[tracked[qi]];
		for (int base = ns; base < n; base += 32) {
			const int j = base + lane;
			const unsigned done = s_flags[(base - ns) >> 5];
			bool ok = false;
			if (j < n && !((done >> lane) & 1u)) ok = track_match(s, cand[j]);
			const unsigned m = __ballot_sync(0xffffffffu, ok);
			if (ok) tracked[qlen + __popc(m & ((1u << lane) - 1u))] = j;
			__syncwarp();
			if (lane == 0 && m) s_flags[(base - ns) >> 5] = done | m;
			qlen += __popc(m);
			__syncwarp();
		}
	}
	if (lane == 0) tk.n_tracked[f] = qlen;
}

// per-frame device lists -> contiguous host-mapped arrays
__global__ void __launch_bounds__(128) k_track_emit(TrackWork tk, int n_frames, ert_tracked *__restrict__ h_cand,
                                                    int32_t *__restrict__ h_cand_off, int32_t *__restrict__ h_nstrong,
                                                    int32_t *__restrict__ h_track_off, int32_t *__restrict__ h_tracked)
{
	const int f = blockIdx.x, tid = threadIdx.x;
	int coff = 0, toff = 0;
	for (int g = 0; g < f; g++) { coff += tk.n_cand[g]; toff += tk.n_tracked[g]; }
	const int n = tk.n_cand[f], nt = tk.n_tracked[f];
	if (tid == 0) {
		h_cand_off[f] = coff; h_track_off[f] = toff; h_nstrong[f] = tk.n_strong[f];
		if (f == n_frames - 1) { h_cand_off[n_frames] = coff + n; h_track_off[n_frames] = toff + nt; }
	}
	// ert_tracked = 12 ints + 3 doubles = 72 bytes = 9 x 8-byte words
	const unsigned long long *src = reinterpret_cast<const unsigned long long *>(tk.cand + (size_t)f * tk.cand_cap);
	unsigned long long *dst = reinterpret_cast<unsigned long long *>(h_cand + coff);
	for (int i = tid; i < n * 9; i += 128) dst[i] = src[i];
	const int32_t *ts = tk.tracked + (size_t)f * tk.cand_cap;
	for (int i = tid; i < nt; i += 128) h_tracked[toff + i] = ts[i];
}

} // namespace

int launch_track_gather(TrackWork &tk, int n_frames, const OutNode *nodes, const int32_t *pool, const int32_t *counts, const int32_t *label,
                        int node_cap, int pool_cap, cudaStream_t st)
{
	k_track_gather<<<n_frames, 32, 0, st>>>(tk, nodes, pool, counts, label, node_cap, pool_cap);
	ERT_CUDA_CHECK(cudaGetLastError());
	return 0;
}

int launch_calc_color(TrackWork &tk, int n_frames, const uint8_t *d_ycc, size_t plane_bytes, int pitch, cudaStream_t st)
{
	k_calc_color<<<dim3(96, n_frames), CC_THREADS, 0, st>>>(tk, d_ycc, plane_bytes, pitch);
	ERT_CUDA_CHECK(cudaGetLastError());
	return 0;
}

int launch_track(TrackWork &tk, int n_frames, ert_tracked *h_cand, int32_t *h_cand_off, int32_t *h_nstrong, int32_t *h_track_off,
                 int32_t *h_tracked, cudaStream_t st)
{
	const size_t sh = sizeof(unsigned) * (size_t)((tk.cand_cap + 31) / 32 + 1);
	k_track<<<n_frames, 32, sh, st>>>(tk);
	ERT_CUDA_CHECK(cudaGetLastError());
	k_track_emit<<<n_frames, 128, 0, st>>>(tk, n_frames, h_cand, h_cand_off, h_nstrong, h_track_off, h_tracked);
	ERT_CUDA_CHECK(cudaGetLastError());
	return 0;
}

} // namespace ert
