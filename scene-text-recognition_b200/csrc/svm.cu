// svm.cu -- batched libsvm C-SVC / RBF probability prediction
// (replaces svm_predict_probability, src/svm.cpp:2592-2629, and what it calls: svm_predict_values
//  :2501-2575, Kernel::k_function RBF :325-365, sigmoid_predict :1818-1826, multiclass_probability
//  :1829-1890) for a BATCH of dense feature vectors (the reference scores one ER at a time).
//
//   k_svm_kvalue : K[n][s] = exp(-gamma * sum_d (x[n][d] - sv[s][d])^2) as a register-tiled FP64
//                  "distance GEMM" (64x64 output tile per CTA, 4x4 per thread, operands staged in
//                  shared memory).  FP64 keeps the parsed support-vector values exact; the integer
//                  tensor-core formulation (u8 x u8 -> s32) is the planned fast path.
//   k_svm_prob   : one warp per vector: 2080 pairwise decision values with the reference's
//                  sequential, non-fused accumulation order, Platt sigmoid, clamp, and the
//                  Wu-Lin-Weng coupling iteration with the reference's exact operation order
//                  (so any difference to the CPU comes only from the last bits of K).
#include "common.cuh"
#include "kernels.h"

namespace ert {

constexpr int KT = 64;   // output tile
constexpr int KD = 16;   // dims per stage

template <typename XT>
__global__ void __launch_bounds__(256) k_svm_kvalue(SvmDev m, const XT *__restrict__ x, int n, double *__restrict__ kv)
{
	__shared__ double xs[KD][KT + 1];
	__shared__ double ss[KD][KT + 1];
	const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
	const int n0 = blockIdx.y * KT, s0 = blockIdx.x * KT;
	double acc[4][4];
#pragma unroll
	for (int i = 0; i < 4; i++)
#pragma unroll
		for (int j = 0; j < 4; j++) acc[i][j] = 0.0;
	for (int d0 = 0; d0 < m.dims; d0 += KD) {
#pragma unroll
		for (int k = 0; k < 4; k++) {
			const int idx = tid + k * 256;
			const int row = idx / KD, dd = idx % KD;
			const int d = d0 + dd;
			double xv = 0.0, sv = 0.0;
			if (d < m.dims) {
				if (n0 + row < n) {
					if (sizeof(XT) == 1) xv = (double)x[(size_t)(n0 + row) * m.dims + d] / 255.0;   // value = u8 / 255.0 (src/OCR.cpp:212)
					else xv = (double)x[(size_t)(n0 + row) * m.dims + d];
				}
				if (s0 + row < m.l) sv = m.sv[(size_t)(s0 + row) * m.dims + d];
			}
			xs[dd][row] = xv;
			ss[dd][row] = sv;
		}
		__syncthreads();
#pragma unroll
		for (int dd = 0; dd < KD; dd++) {
			double a[4], b[4];
#pragma unroll
			for (int i = 0; i < 4; i++) { a[i] = xs[dd][ty * 4 + i]; b[i] = ss[dd][tx * 4 + i]; }
#pragma unroll
			for (int i = 0; i < 4; i++)
#pragma unroll
				for (int j = 0; j < 4; j++) { const double df = a[i] - b[j]; acc[i][j] = fma(df, df, acc[i][j]); }
		}
		__syncthreads();
	}
#pragma unroll
	for (int i = 0; i < 4; i++)
#pragma unroll
		for (int j = 0; j < 4; j++) {
			const int r = n0 + ty * 4 + i, s = s0 + tx * 4 + j;
			if (r < n && s < m.l) kv[(size_t)r * m.l + s] = exp(-m.gamma * acc[i][j]);
		}
}

constexpr int PROB_WARPS = 2;
constexpr int MAXK = 96;   // classes supported by the 3-slots-per-lane layout

__device__ __forceinline__ double sigmoid_predict_dev(double dec, double A, double B)
{
	const double f = __dadd_rn(__dmul_rn(dec, A), B);
	if (f >= 0) { const double e = exp(-f); return e / (1.0 + e); }
	return 1.0 / (1.0 + exp(f));
}

__global__ void __launch_bounds__(PROB_WARPS * 32) k_svm_prob(SvmDev m, const double *__restrict__ kv, int n, double *__restrict__ label_out,
                                                              double *__restrict__ prob_out)
{
	extern __shared__ __align__(16) double dsm[];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int k = m.nr_class;
	double *r = dsm + (size_t)warp * ((size_t)k * k + 3 * MAXK);
	double *p = r + (size_t)k * k;
	double *Qp = p + MAXK;
	double *Qtt = Qp + MAXK;
	const int v = blockIdx.x * PROB_WARPS + warp;
	if (v >= n) return;
	const double *kvv = kv + (size_t)v * m.l;

	// pairwise decision values -> pairwise probabilities
	for (int i = 0; i < k; i++) {
		if (lane == 0) r[i * k + i] = 0.0;
		for (int j = i + 1 + lane; j < k; j += 32) {
			const int pidx = i * k - i * (i + 1) / 2 + (j - i - 1);
			const double *c1 = m.coef + (size_t)(j - 1) * m.l, *c2 = m.coef + (size_t)i * m.l;
			const int si = m.start[i], sj = m.start[j], ci = m.nsv[i], cj = m.nsv[j];
			double sum = 0.0;
			for (int q = 0; q < ci; q++) sum = __dadd_rn(sum, __dmul_rn(c1[si + q], kvv[si + q]));
			for (int q = 0; q < cj; q++) sum = __dadd_rn(sum, __dmul_rn(c2[sj + q], kvv[sj + q]));
			sum = __dadd_rn(sum, -m.rho[pidx]);
			double pr = sigmoid_predict_dev(sum, m.probA[pidx], m.probB[pidx]);
			const double lo = 1e-7;
			pr = fmin(fmax(pr, lo), 1.0 - lo);
			r[i * k + j] = pr;
			r[j * k + i] = 1.0 - pr;
		}
	}
	__syncwarp();

	// multiclass_probability: Q[t][t] = sum_{j != t} r[j][t]^2 ; Q[t][j] = -r[j][t] * r[t][j]
	for (int t = lane; t < k; t += 32) {
		double q = 0.0;
		for (int j = 0; j < k; j++) if (j != t) q = __dadd_rn(q, __dmul_rn(r[j * k + t], r[j * k + t]));
		Qtt[t] = q;
		p[t] = 1.0 / k;
	}
	__syncwarp();
	const int max_iter = max(100, k);
	const double eps = 0.005 / k;
	for (int iter = 0; iter < max_iter; iter++) {
		for (int t = lane; t < k; t += 32) {
			double s = 0.0;
			for (int j = 0; j < k; j++) {
				const double q = (j == t) ? Qtt[t] : -__dmul_rn(r[j * k + t], r[t * k + j]);
				s = __dadd_rn(s, __dmul_rn(q, p[j]));
			}
			Qp[t] = s;
		}
		__syncwarp();
		double pQp = 0.0;
		for (int t = 0; t < k; t++) pQp = __dadd_rn(pQp, __dmul_rn(p[t], Qp[t]));   // every lane, same order
		double max_err = 0.0;
		for (int t = 0; t < k; t++) max_err = fmax(max_err, fabs(__dadd_rn(Qp[t], -pQp)));
		if (max_err < eps) break;
		for (int t = 0; t < k; t++) {
			const double qtt = Qtt[t], qpt = Qp[t];
			const double diff = __ddiv_rn(__dadd_rn(-qpt, pQp), qtt);
			const double one_d = __dadd_rn(1.0, diff);
			pQp = __ddiv_rn(__ddiv_rn(__dadd_rn(pQp, __dmul_rn(diff, __dadd_rn(__dmul_rn(diff, qtt), __dmul_rn(2.0, qpt)))), one_d), one_d);
			__syncwarp();
			for (int j = lane; j < k; j += 32) {
				const double q = (j == t) ? qtt : -__dmul_rn(r[j * k + t], r[t * k + j]);
				const double pj = (j == t) ? __dadd_rn(p[j], diff) : p[j];
				Qp[j] = __ddiv_rn(__dadd_rn(Qp[j], __dmul_rn(diff, q)), one_d);
				p[j] = __ddiv_rn(pj, one_d);
			}
			__syncwarp();
		}
	}
	__syncwarp();
	int best = 0;
	for (int t = 1; t < k; t++) if (p[t] > p[best]) best = t;
	for (int t = lane; t < k; t += 32) prob_out[(size_t)v * k + t] = p[t];
	if (lane == 0) label_out[v] = (double)m.label[best];
}

int launch_svm_predict(const SvmDev &m, const double *x_f64, const uint8_t *x_u8, int n, double *kvalue_ws, double *label, double *prob,
                       cudaStream_t st)
{
	if (n <= 0) return 0;
	if (m.nr_class > MAXK) { set_error("svm: nr_class %d > %d unsupported", m.nr_class, MAXK); return -1; }
	dim3 grid((m.l + KT - 1) / KT, (n + KT - 1) / KT);
	if (x_u8) k_svm_kvalue<uint8_t><<<grid, 256, 0, st>>>(m, x_u8, n, kvalue_ws);
	else k_svm_kvalue<double><<<grid, 256, 0, st>>>(m, x_f64, n, kvalue_ws);
	ERT_CUDA_CHECK(cudaGetLastError());
	const size_t smem = (size_t)PROB_WARPS * ((size_t)m.nr_class * m.nr_class + 3 * MAXK) * sizeof(double);
		ERT_CUDA_CHECK(cudaFuncSetAttribute(k_svm_prob, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
	k_svm_prob<<<(n + PROB_WARPS - 1) / PROB_WARPS, PROB_WARPS * 32, smem, st>>>(m, kvalue_ws, n, label, prob);
	ERT_CUDA_CHECK(cudaGetLastError());
	return 0;
}

} // namespace ert
