// ocr_feat.cu -- the OCR feature path that feeds the SVM: OCR::chain_run's pre-processing and
// OCR::extract_feature (reference src/OCR.cpp:67-84, 144-218, 254-360, 394-430, 602-622), one CTA per region.
//
//   1. threshold(255 - src, THRESH_OTSU): histogram of the bound, OpenCV's OTSU recurrence in FP64 (no contraction)
//   2. rotate_mat (only when |slope| > 0.01) and ARAN(30): nothing is materialised -- every tap of the fixed-point
//      bilinear resize evaluates the rotated / thresholded pixel it needs straight from the plane
//   3. findContours(RETR_LIST, CHAIN_APPROX_NONE): Suzuki-Abe border following on the 30x30 image in shared memory
//      (sequential by definition, ~10^3 steps, one thread); each step (pixel, direction) sets the pixel in one of the
//      eight chain-code bitmaps -- contour ORDER does not matter because the bitmaps are only ever set to 255
//   4. GaussianBlur 7x7 (OpenCV's 8-bit fixed-point kernel {8,28,56,72,56,28,8}/256, BORDER_REFLECT_101),
//      normalize(NORM_MINMAX 0..255: float scale/shift, one FMA, round-half-even), resize 30 -> 15 (2x2 box)
//   5. the 8 x 15 x 15 bytes are the SVM feature vector (value = byte / 255, src/OCR.cpp:203-218)
// Integer / byte work is bit-exact; the FP64 steps follow the reference's operation order with round-to-nearest
// intrinsics so that they round identically.
#include "kernels.h"

namespace ert {

namespace {

constexpr int OCR_THREADS = 128;
constexpr int OCR_WARPS = OCR_THREADS / 32;
constexpr int IL = 30;     // OCR_IMG_L      (inc/utils.h:12)
constexpr int FL = 15;     // OCR_FEATURE_L  (inc/utils.h:13)

__device__ int otsu_threshold_ocr(const int *h, int total)
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

// pixel of the OTSU image (threshold(255 - src, ..., THRESH_OTSU), src/OCR.cpp:72) at bound coordinates
__device__ __forceinline__ int bin_px(const OcrJob &J, int thr, int x, int y)
{
	const int v = __ldg(J.src + (size_t)(J.y0 + y) * J.pitch + J.x0 + x);
	const int u = J.invert ? v : 255 - v;
	return u > thr ? 255 : 0;
}

// pixel (rj, ri) of OCR::rotate_mat's output (src/OCR.cpp:254-360), evaluated on demand
__device__ int rot_px(const OcrJob &J, int thr, int rj, int ri)
{
	if (ri >= J.sh - 1 || rj >= J.sw - 1) return 0;            // loops stop one short of max_y / max_x
	if (J.rot == 1 && ri == 0) return 0;                         // crop mode: i > min_y + crop_height
	const int i = ri + J.min_y + J.crop_h, j = rj + J.min_x;
	const double ii = (double)(i - J.crop_h), jj = (double)j;
	const double new_j = __dadd_rn(__dsub_rn(__dmul_rn(J.cs, jj), __dmul_rn(J.sn, ii)), (double)J.cx);
	const double new_i = __dadd_rn(__dadd_rn(__dmul_rn(J.sn, jj), __dmul_rn(J.cs, ii)), (double)J.cy);
	if (!(new_i > 0 && new_j > 0 && new_i < (double)(J.h - 1) && new_j < (double)(J.w - 1))) return 0;
	const int yi = (int)new_i, xj = (int)new_j;
	const double fi = floor(new_i), fj = floor(new_j);
	if (new_i == fi && new_j == fj) return bin_px(J, thr, xj, yi);
	const double alpha = __dsub_rn(new_i, fi), beta = __dsub_rn(new_j, fj);
	const double A = (double)bin_px(J, thr, xj, yi), B = (double)bin_px(J, thr, xj + 1, yi);
	const double C = (double)bin_px(J, thr, xj, yi + 1), D = (double)bin_px(J, thr, xj + 1, yi + 1);
	const double na = __dsub_rn(1.0, alpha), nb = __dsub_rn(1.0, beta);
	double v = __dmul_rn(__dmul_rn(na, nb), A);
	v = __dadd_rn(v, __dmul_rn(__dmul_rn(na, beta), B));
	v = __dadd_rn(v, __dmul_rn(__dmul_rn(alpha, nb), C));
	v = __dadd_rn(v, __dmul_rn(__dmul_rn(alpha, beta), D));
	return (int)(unsigned char)(int)round(v);
}

__device__ __forceinline__ int aran_src(const OcrJob &J, int thr, int x, int y)
{
	return J.rot ? rot_px(J, thr, x, y) : bin_px(J, thr, x, y);
}

__global__ void __launch_bounds__(OCR_THREADS) k_ocr_features(const OcrJob *__restrict__ jobs, uint8_t *__restrict__ feat_out,
                                                               uint8_t *__restrict__ img_out)
{
	__shared__ int s_hist[OCR_WARPS][256];
	__shared__ int s_thr;
	__shared__ signed char s_img[32 * 32];             // 30x30 image with a zero frame (findContours works on a padded copy)
	__shared__ uint8_t s_f[8][IL * IL];                // chain-code bitmaps, later the normalised maps
	__shared__ unsigned short s_h[8][IL * IL];         // horizontal blur pass, 8.8 fixed point
	__shared__ float s_scale[8], s_shift[8];
	const OcrJob J = jobs[blockIdx.x];
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

	// ---- 1. OTSU threshold of 255 - channel over the bound ----
	for (int i = tid; i < OCR_WARPS * 256; i += OCR_THREADS) (&s_hist[0][0])[i] = 0;
	for (int i = tid; i < 32 * 32; i += OCR_THREADS) s_img[i] = 0;
	for (int i = tid; i < 8 * IL * IL; i += OCR_THREADS) (&s_f[0][0])[i] = 0;
	__syncthreads();
	for (int r = warp; r < J.h; r += OCR_WARPS) {
		const uint8_t *row = J.src + (size_t)(J.y0 + r) * J.pitch + J.x0;
		for (int xb = 0; xb < J.w; xb += 32) {
			const int x = xb + lane;
			const bool in = x < J.w;
			int u = 0;
			if (in) { const int v = row[x]; u = J.invert ? v : 255 - v; }
			if (in) atomicAdd(&s_hist[warp][u], 1);
		}
	}
	__syncthreads();
	for (int i = tid; i < 256; i += OCR_THREADS) {
		int s = 0;
#pragma unroll
		for (int k = 0; k < OCR_WARPS; k++) s += s_hist[k][i];
		s_hist[0][i] = s;
	}
	__syncthreads();
	if (tid == 0) s_thr = otsu_threshold_ocr(s_hist[0], J.w * J.h);
	__syncthreads();
	const int thr = s_thr;

	// ---- 2. (rotate_mat) + ARAN(30): OpenCV INTER_LINEAR 8-bit fixed point, exact-2x -> INTER_AREA box ----
	{
		const int sw = J.sw, sh = J.sh, dw = J.dw, dh = J.dh;
		if (sw == 2 * dw && sh == 2 * dh) {
			for (int t = tid; t < dw * dh; t += OCR_THREADS) {
				const int oy = t / dw, ox = t % dw;
				const int s = aran_src(J, thr, 2 * ox, 2 * oy) + aran_src(J, thr, 2 * ox + 1, 2 * oy) + aran_src(J, thr, 2 * ox, 2 * oy + 1) +
				              aran_src(J, thr, 2 * ox + 1, 2 * oy + 1);
				const int v = (s + 2) >> 2;
				s_img[(oy + J.offy + 1) * 32 + ox + J.offx + 1] = (signed char)(v != 0);
				if (img_out) img_out[(size_t)blockIdx.x * IL * IL + (oy + J.offy) * IL + ox + J.offx] = (uint8_t)v;
			}
		} else {
			const double scx = 1.0 / ((double)dw / (double)sw), scy = 1.0 / ((double)dh / (double)sh);
			for (int t = tid; t < dw * dh; t += OCR_THREADS) {
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
				const int h0 = aran_src(J, thr, ix, r0) * wx0 + aran_src(J, thr, ix1, r0) * wx1;
				const int h1 = aran_src(J, thr, ix, r1) * wx0 + aran_src(J, thr, ix1, r1) * wx1;
				int v = (((wy0 * (h0 >> 4)) >> 16) + ((wy1 * (h1 >> 4)) >> 16) + 2) >> 2;
				v = min(max(v, 0), 255);
				s_img[(oy + J.offy + 1) * 32 + ox + J.offx + 1] = (signed char)(v != 0);
				if (img_out) img_out[(size_t)blockIdx.x * IL * IL + (oy + J.offy) * IL + ox + J.offx] = (uint8_t)v;
			}
		}
	}
	__syncthreads();

	// ---- 3. findContours(RETR_LIST, CHAIN_APPROX_NONE) -> chain-code bitmaps (src/OCR.cpp:153-169) ----
	// pixel states: 0 background, 1 untouched foreground, 2 visited, 2|-128 visited with the border on its east side.
	// directions: 0=E 1=NE 2=N 3=NW 4=W 5=SW 6=S 7=SE; chain_code_direction(next, cur) of the reference = (4 - s) & 7.
	if (tid == 0) {
		const int dx[8] = {1, 1, 0, -1, -1, -1, 0, 1}, dy[8] = {0, -1, -1, -1, 0, 1, 1, 1};
		int delta[16];
#pragma unroll
		for (int k = 0; k < 16; k++) delta[k] = dy[k & 7] * 32 + dx[k & 7];
		for (int y = 1; y <= IL; y++) {
			int prev = 0;
			for (int x = 1; x <= IL + 1; x++) {
				int p = s_img[y * 32 + x];
				if (p == prev) continue;
				bool is_hole = false, start = false;
				if (prev == 0 && p == 1) start = true;
				else if (p == 0 && prev >= 1) { start = true; is_hole = true; }
				if (start) {
					const int i0 = y * 32 + x - (is_hole ? 1 : 0);
					int s_end = is_hole ? 0 : 4, s = s_end, i1;
					do { s = (s - 1) & 7; i1 = i0 + delta[s]; } while (s_img[i1] == 0 && s != s_end);
					if (s == s_end) {
						s_img[i0] = (signed char)(2 | -128);        // isolated pixel: a 1-point contour, skipped by the reference
					} else {
						int i3 = i0, i4 = i0;
						for (int guard = 0; guard < 8 * 32 * 32; guard++) {
							s_end = s;
							while (s < 15) { i4 = i3 + delta[++s]; if (s_img[i4] != 0) break; }
							s &= 7;
							if ((unsigned)(s - 1) < (unsigned)s_end) s_img[i3] = (signed char)(2 | -128);
							else if (s_img[i3] == 1) s_img[i3] = 2;
							const int py = (i3 >> 5) - 1, px = (i3 & 31) - 1;
							s_f[(4 - s) & 7][py * IL + px] = 255;
							if (i4 == i0 && i3 == i1) break;
							i3 = i4;
							s = (s + 4) & 7;
						}
					}
					p = s_img[y * 32 + x];
				}
				prev = p;
			}
		}
	}
	__syncthreads();

	// ---- 4. GaussianBlur 7x7 (fixed point), normalize MINMAX, resize 30 -> 15 ----
	for (int t = tid; t < 8 * IL * IL; t += OCR_THREADS) {
		const int c = t / (IL * IL), r = (t % (IL * IL)) / IL, x = t % IL;
		const uint8_t *row = &s_f[c][r * IL];
		const int K[7] = {8, 28, 56, 72, 56, 28, 8};
		int s = 0;
#pragma unroll
		for (int k = 0; k < 7; k++) { int xi = x + k - 3; xi = xi < 0 ? -xi : (xi >= IL ? 2 * IL - 2 - xi : xi); s += K[k] * row[xi]; }
		s_h[c][r * IL + x] = (unsigned short)s;
	}
	__syncthreads();
	for (int t = tid; t < 8 * IL * IL; t += OCR_THREADS) {
		const int c = t / (IL * IL), r = (t % (IL * IL)) / IL, x = t % IL;
		const int K[7] = {8, 28, 56, 72, 56, 28, 8};
		int s = 0;
#pragma unroll
		for (int k = 0; k < 7; k++) { int yi = r + k - 3; yi = yi < 0 ? -yi : (yi >= IL ? 2 * IL - 2 - yi : yi); s += K[k] * s_h[c][yi * IL + x]; }
		s_f[c][r * IL + x] = (uint8_t)((s + 32768) >> 16);
	}
	__syncthreads();
	for (int c = warp; c < 8; c += OCR_WARPS) {
		int mn = 255, mx = 0;
		for (int i = lane; i < IL * IL; i += 32) { const int v = s_f[c][i]; mn = min(mn, v); mx = max(mx, v); }
#pragma unroll
		for (int o = 16; o; o >>= 1) { mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o)); mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o)); }
		if (lane == 0) {
			// cv::normalize: scale = (255 - 0) * (1 / (smax - smin)) (or 0), shift = 0 - smin * scale, both cast to float
			const double d = (double)(mx - mn);
			const double scale = __dmul_rn(255.0, d > 2.220446049250313e-16 ? __ddiv_rn(1.0, d) : 0.0);
			const double shift = __dsub_rn(0.0, __dmul_rn((double)mn, scale));
			s_scale[c] = (float)scale; s_shift[c] = (float)shift;
		}
	}
	__syncthreads();
	for (int t = tid; t < 8 * IL * IL; t += OCR_THREADS) {
		const int c = t / (IL * IL), i = t % (IL * IL);
		int v = __float2int_rn(__fmaf_rn((float)s_f[c][i], s_scale[c], s_shift[c]));
		s_f[c][i] = (uint8_t)min(max(v, 0), 255);
	}
	__syncthreads();
	uint8_t *fo = feat_out + (size_t)blockIdx.x * (8 * FL * FL);
	for (int t = tid; t < 8 * FL * FL; t += OCR_THREADS) {
		const int c = t / (FL * FL), p = t % (FL * FL), y = p / FL, x = p % FL;
		const uint8_t *a = &s_f[c][(2 * y) * IL + 2 * x];
		fo[t] = (uint8_t)((a[0] + a[1] + a[IL] + a[IL + 1] + 2) >> 2);
	}
}

} // namespace

int launch_ocr_features(const OcrJob *d_jobs, int n, uint8_t *d_feat1800, uint8_t *d_img30, cudaStream_t st)
{
	if (n <= 0) return 0;
	k_ocr_features<<<n, OCR_THREADS, 0, st>>>(d_jobs, d_feat1800, d_img30);
	ERT_CUDA_CHECK(cudaGetLastError());
	return 0;
}

} // namespace ert
