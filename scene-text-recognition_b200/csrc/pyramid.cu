// pyramid.cu -- image pyramid levels on the device (BASELINE configs 2 and 4).
//
// The reference runs its six planes at native scale only (src/ER.cpp:122-127); SURVEY 8d defines a pyramid level as the
// same per-plane path applied to the plane resized by cv::resize(INTER_LINEAR) -- the primitive the path already leans on
// (src/OCR.cpp:401).  k_resize_planes is that primitive, bit-exact against OpenCV 4.x for 8-bit single-channel images
// (pinned against cv2 in tests/test_gpu_pipeline.py):
//   * exact 2x reduction in both axes -> OpenCV switches to INTER_AREA: (a + b + c + d + 2) >> 2
//   * otherwise 11-bit fixed-point bilinear: per axis f = (d + 0.5) * scale - 0.5, s = floor(f), horizontal taps clamped
//     with the weight snapped (s < 0 -> s = 0, f = 0; s >= sw - 1 -> s = sw - 1, f = 0), vertical taps clamp the two ROW
//     indices and keep the weights; out = ((b0 * (H0 >> 4)) >> 16) + ((b1 * (H1 >> 4)) >> 16) + 2) >> 2
// One thread per 4 output pixels of a row (one 32-bit store); source rows come through the read-only path.
#include "common.cuh"
#include "kernels.h"

namespace ert {

// source plane p = what the source context's plane table says (pointer + invert flag): an inverted channel is inverted
// BEFORE it is resized, as cv::resize(255 - channel) would see it (resizing and inverting do not commute bit for bit)
#define ERT_SRC(ptr) (inv ? 255 - (int)__ldg(ptr) : (int)__ldg(ptr))
__global__ void __launch_bounds__(128) k_resize_planes(const PlaneSrc *__restrict__ planes, int sw, int sh, int spitch, uint8_t *__restrict__ dst,
                                                       int dw, int dh, int dpitch, size_t dplane)
{
	const int plane = blockIdx.z, oy = blockIdx.y;
	const int ox0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
	if (ox0 >= dw) return;
	const PlaneSrc ps = planes[plane];
	const uint8_t *S = ps.src;
	const bool inv = ps.invert != 0;
	uint8_t out[4] = {0, 0, 0, 0};
	if (sw == 2 * dw && sh == 2 * dh) {
		const uint8_t *r0 = S + (size_t)(2 * oy) * spitch, *r1 = r0 + spitch;
#pragma unroll
		for (int k = 0; k < 4; k++) {
			const int ox = ox0 + k;
			if (ox < dw) out[k] = (uint8_t)((ERT_SRC(r0 + 2 * ox) + ERT_SRC(r0 + 2 * ox + 1) + ERT_SRC(r1 + 2 * ox) + ERT_SRC(r1 + 2 * ox + 1) + 2) >> 2);
		}
	} else {
		const double scx = 1.0 / ((double)dw / (double)sw), scy = 1.0 / ((double)dh / (double)sh);
		float fy = (float)(((double)oy + 0.5) * scy - 0.5);
		int iy = (int)floorf(fy);
		fy -= (float)iy;
		const int wy0 = __float2int_rn((1.f - fy) * 2048.f), wy1 = __float2int_rn(fy * 2048.f);
		const uint8_t *r0 = S + (size_t)min(max(iy, 0), sh - 1) * spitch, *r1 = S + (size_t)min(max(iy + 1, 0), sh - 1) * spitch;
#pragma unroll
		for (int k = 0; k < 4; k++) {
			const int ox = ox0 + k;
			if (ox >= dw) break;
			float fx = (float)(((double)ox + 0.5) * scx - 0.5);
			int ix = (int)floorf(fx);
			fx -= (float)ix;
			if (ix < 0) { ix = 0; fx = 0.f; }
			if (ix >= sw - 1) { ix = sw - 1; fx = 0.f; }
			const int wx0 = __float2int_rn((1.f - fx) * 2048.f), wx1 = __float2int_rn(fx * 2048.f);
			const int ix1 = min(ix + 1, sw - 1);
			const int h0 = ERT_SRC(r0 + ix) * wx0 + ERT_SRC(r0 + ix1) * wx1;
			const int h1 = ERT_SRC(r1 + ix) * wx0 + ERT_SRC(r1 + ix1) * wx1;
			int v = (((wy0 * (h0 >> 4)) >> 16) + ((wy1 * (h1 >> 4)) >> 16) + 2) >> 2;
			out[k] = (uint8_t)min(max(v, 0), 255);
		}
	}
	// dpitch is a multiple of 128 and ox0 of 4: the 4-byte store is aligned (pad bytes are don't-care)
	*reinterpret_cast<uchar4 *>(dst + (size_t)plane * dplane + (size_t)oy * dpitch + ox0) = make_uchar4(out[0], out[1], out[2], out[3]);
}

#undef ERT_SRC

int launch_resize_planes(const PlaneSrc *d_src_planes, int n_planes, int sw, int sh, int spitch, uint8_t *d_dst, int dw, int dh, int dpitch, size_t dplane,
                         cudaStream_t st)
{
	if (n_planes < 1 || dw < 1 || dh < 1) return 0;
	dim3 grid(((dw + 3) / 4 + 127) / 128, dh, n_planes);
	k_resize_planes<<<grid, 128, 0, st>>>(d_src_planes, sw, sh, spitch, d_dst, dw, dh, dpitch, dplane);
	ERT_CUDA_CHECK(cudaGetLastError());
	return 0;
}

} // namespace ert
