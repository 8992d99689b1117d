// svm.cu -- batched libsvm C-SVC / RBF probability prediction
// (replaces svm_predict_probability, src/svm.cpp:2592-2629, and what it calls: svm_predict_values
//  :2501-2575, Kernel::k_function RBF :325-365, sigmoid_predict :1818-1826, multiclass_probability
//  :1829-1890) for a BATCH of dense feature vectors (the reference scores one ER at a time).
//
//   k_svm_prep_x     : u8 features padded to 1920 bytes per row + their squared norms
//   k_svm_kvalue_tma : (svm_gemm.cu) K[n][s] = exp(-gamma |x_n - sv_s|^2) for u8 features as two integer GEMMs on the tensor
//                      cores (tcgen05 kind::i8, TMA ring, TMEM double buffer) with an FP64 epilogue; k_svm_kvalue_tc below is the
//                      round-1 single-stage form of the same GEMMs (A-B, bit-identical results)
//   k_svm_kvalue     : the same matrix as a register-tiled FP64 "distance GEMM" on CUDA cores: arbitrary double inputs, and
//                      the audit path (parsed support-vector values enter as they are: ~1e-13 against the reference)
//   k_svm_decide     : per class block [vectors x nsv_c] x [nsv_c x 64] -> S[vector][class][other] (cp.async ring)
//   k_svm_couple     : one warp per vector: decision values from S, Platt sigmoid, clamp, Wu-Lin-Weng coupling on the full
//                      symmetric matrix in shared memory (last-bit differences from libsvm only)
//   k_svm_prob       : round-1 form of the last two (one warp per vector for everything), kept for A-B (ert_set_svm_legacy_prob)
#include "common.cuh"
#include "kernels.h"
#include "svm_math.cuh"

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
			if (r < n && s < m.l) kv[(size_t)r * m.ldk + s] = exp(-m.gamma * acc[i][j]);
		}
}

// ---------------------------------------------------------------------------------------------
// Tensor-core formulation of the RBF distance (u8 features only).
//   x.sv = sum_d (k_d/255) * (j_d/255 + eps_d)  with k_d the u8 feature, j_d = round(255 v_d), eps_d = v_d - j_d/255
//        = DJ / 255^2 + DE / (255 * S),   DJ = sum k_d j_d  (u8 x u8 -> s32, exact),
//                                         DE = sum k_d e_d  (u8 x s8 -> s32), e_d = round(S * eps_d), S = 127 / max|eps|
//   d^2  = |x|^2 + |sv|^2 - 2 x.sv ;  |x|^2 = sum k^2 / 255^2 (exact integer sum), |sv|^2 in FP64 from the parsed values.
// The OCR.model values are 6-significant-digit decimals of j/255 (|eps| <= 4.9e-7), so the second GEMM restores
// them to ~1e-9; both are dense [N x 1800] x [1800 x 1910] GEMMs -> tcgen05.mma kind::i8, accumulators in TMEM.
//
// One CTA per 128 (vectors) x 256 (support vectors) output tile; K is consumed in 128-byte chunks: all threads
// stage A (128 rows), B_J and B_E (256 rows each) into shared memory in the canonical no-swizzle K-major
// core-matrix layout (8 rows x 16 B), one thread issues 4 + 4 MMAs (K = 32 each), tcgen05.commit signals an
// mbarrier, everybody waits, next chunk.  Epilogue: tcgen05.ld 32 lanes x 16 columns per warp -> exp -> FP64 K.
// (Single-stage on purpose in round 1: the whole GEMM is ~2 % of the SVM time once it runs on tensor cores.)
// ---------------------------------------------------------------------------------------------
constexpr int TC_M = 128, TC_N = 256, TC_KC = 128;        // tile and K-chunk (bytes)
constexpr int TC_KPAD = 1920;                             // 1800 padded to a multiple of TC_KC
constexpr int TC_NPAD = 2048;                             // 1910 support vectors padded to 8 tiles of 256

__device__ __forceinline__ uint32_t tc_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t tc_make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
	// UMMA shared-memory matrix descriptor, SWIZZLE_NONE, K-major: ((8,n),2):((1,SBO),LBO) in 16-byte units
	uint64_t d = 0;
	d |= (uint64_t)((smem_addr >> 4) & 0x3FFFu);
	d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
	d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
	d |= (uint64_t)1 << 46;   // descriptor version for sm_100
	return d;
}

__global__ void __launch_bounds__(128) k_svm_prep_x(const uint8_t *__restrict__ x, int n, int dims, uint8_t *__restrict__ xp, uint32_t *__restrict__ xx)
{
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int row = blockIdx.x * 4 + warp;
	if (row >= n) return;
	const uint8_t *src = x + (size_t)row * dims;
	uint8_t *dst = xp + (size_t)row * TC_KPAD;
	uint32_t acc = 0;
	for (int d = lane; d < TC_KPAD; d += 32) {
		const uint32_t v = (d < dims) ? src[d] : 0u;
		dst[d] = (uint8_t)v;
		acc += v * v;
	}
	acc = __reduce_add_sync(0xFFFFFFFFu, acc);
	if (lane == 0) xx[row] = acc;
}

__global__ void __launch_bounds__(128, 1) k_svm_kvalue_tc(const uint8_t *__restrict__ xp, const uint32_t *__restrict__ xx, int n,
                                                         const uint8_t *__restrict__ svj, const int8_t *__restrict__ sve,
                                                         const double *__restrict__ ss, int l, int ldk, double gamma, double inv_s255,
                                                         double *__restrict__ kv)
{
	extern __shared__ __align__(1024) uint8_t tsm[];
	uint8_t *sA = tsm;                         // 128 rows x 128 B  = 16 KB
	uint8_t *sBJ = tsm + 16384;                // 256 rows x 128 B  = 32 KB
	uint8_t *sBE = tsm + 16384 + 32768;        // 32 KB
	__shared__ __align__(8) uint64_t bar;
	__shared__ uint32_t s_tmem;
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const int n0 = blockIdx.y * TC_M, s0 = blockIdx.x * TC_N;

	if (warp == 0) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(&s_tmem)), "r"(512u) : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
	}
	if (tid == 0) {
		asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tc_smem_u32(&bar)), "r"(1u) : "memory");
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	const uint32_t tmem = s_tmem;

	// instruction descriptors: D = S32, A = u8 (K-major), B = u8 / s8 (K-major), M = 128, N = 256
	const uint32_t idesc_base = (2u << 4) | (0u << 7) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
	const uint32_t idesc_j = idesc_base | (0u << 10);
	const uint32_t idesc_e = idesc_base | (1u << 10);
	const uint32_t aA = tc_smem_u32(sA), aBJ = tc_smem_u32(sBJ), aBE = tc_smem_u32(sBE);

	uint32_t phase = 0;
	for (int kc = 0; kc < TC_KPAD / TC_KC; ++kc) {
		// stage the chunk: granule (r, g) of 16 bytes -> core matrix (g, r/8), row r%8
		for (int q = tid; q < TC_M * 8; q += 128) {
			const int r = q >> 3, g = q & 7;
			uint4 v = make_uint4(0, 0, 0, 0);
			if (n0 + r < n) v = *reinterpret_cast<const uint4 *>(xp + (size_t)(n0 + r) * TC_KPAD + kc * TC_KC + g * 16);
			*reinterpret_cast<uint4 *>(sA + ((g * (TC_M / 8) + (r >> 3)) * 8 + (r & 7)) * 16) = v;
		}
		for (int q = tid; q < TC_N * 8; q += 128) {
			const int r = q >> 3, g = q & 7;
			const size_t off = (size_t)(s0 + r) * TC_KPAD + kc * TC_KC + g * 16;
			const int so = ((g * (TC_N / 8) + (r >> 3)) * 8 + (r & 7)) * 16;
			*reinterpret_cast<uint4 *>(sBJ + so) = *reinterpret_cast<const uint4 *>(svj + off);
			*reinterpret_cast<uint4 *>(sBE + so) = *reinterpret_cast<const uint4 *>(reinterpret_cast<const uint8_t *>(sve) + off);
		}
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> visible to the tensor core
		asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
		__syncthreads();
		asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
		if (tid == 0) {
#pragma unroll
			for (int k4 = 0; k4 < TC_KC / 32; ++k4) {
				const uint64_t da = tc_make_desc(aA + (uint32_t)(2 * k4) * (TC_M / 8) * 128, (TC_M / 8) * 128, 128);
				const uint64_t dj = tc_make_desc(aBJ + (uint32_t)(2 * k4) * (TC_N / 8) * 128, (TC_N / 8) * 128, 128);
				const uint64_t de = tc_make_desc(aBE + (uint32_t)(2 * k4) * (TC_N / 8) * 128, (TC_N / 8) * 128, 128);
				const uint32_t acc = (kc > 0 || k4 > 0) ? 1u : 0u;
				asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
				             "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
				             ::"r"(tmem), "l"(da), "l"(dj), "r"(idesc_j), "r"(acc) : "memory");
				asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
				             "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
				             ::"r"(tmem + 256u), "l"(da), "l"(de), "r"(idesc_e), "r"(acc) : "memory");
			}
			asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tc_smem_u32(&bar)) : "memory");
		}
		// wait until the MMAs of this chunk have consumed shared memory
		{
			uint32_t ok;
			int spins = 0;
			do {
				asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
				             : "=r"(ok) : "r"(tc_smem_u32(&bar)), "r"(phase) : "memory");
			} while (!ok && ++spins < (1 << 22));   // bounded: a lost commit must not hang the GPU
		}
		phase ^= 1u;
	}
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

	// epilogue: warp w owns TMEM lanes 32w..32w+31 = tile rows; 16 columns per tcgen05.ld
	const int row = n0 + warp * 32 + lane;
	const double xxr = (row < n) ? (double)xx[row] / 65025.0 : 0.0;
	const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
	for (int c0 = 0; c0 < TC_N; c0 += 16) {
		uint32_t rj[16], re[16];
		asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
		             : "=r"(rj[0]), "=r"(rj[1]), "=r"(rj[2]), "=r"(rj[3]), "=r"(rj[4]), "=r"(rj[5]), "=r"(rj[6]), "=r"(rj[7]),
		               "=r"(rj[8]), "=r"(rj[9]), "=r"(rj[10]), "=r"(rj[11]), "=r"(rj[12]), "=r"(rj[13]), "=r"(rj[14]), "=r"(rj[15])
		             : "r"(lane_base + (uint32_t)c0) : "memory");
		asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
		             : "=r"(re[0]), "=r"(re[1]), "=r"(re[2]), "=r"(re[3]), "=r"(re[4]), "=r"(re[5]), "=r"(re[6]), "=r"(re[7]),
		               "=r"(re[8]), "=r"(re[9]), "=r"(re[10]), "=r"(re[11]), "=r"(re[12]), "=r"(re[13]), "=r"(re[14]), "=r"(re[15])
		             : "r"(lane_base + 256u + (uint32_t)c0) : "memory");
		asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
		if (row < n) {
#pragma unroll
			for (int j = 0; j < 16; ++j) {
				const int s = s0 + c0 + j;
				if (s < l) {
					const double dot = fma((double)(int32_t)rj[j], 1.0 / 65025.0, (double)(int32_t)re[j] * inv_s255);   // no division: its slow-path call would serialise the 16 chains
					const double d2 = xxr + ss[s] - 2.0 * dot;
					kv[(size_t)row * ldk + s] = exp_nonpos(-gamma * fmax(d2, 0.0));
				}
			}
		}
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

constexpr int PROB_WARPS = 4;
constexpr int MAXK = 96;   // classes supported by the 3-slots-per-lane layout

__device__ __forceinline__ double sigmoid_predict_dev(double dec, double A, double B)
{
	const double f = __dadd_rn(__dmul_rn(dec, A), B);
	if (f >= 0) { const double e = exp(-f); return e / (1.0 + e); }
	return 1.0 / (1.0 + exp(f));
}

// 1 / x to within an ulp or two: hardware seed (>= 20 bits) + two Newton steps.  The Gauss-Seidel sweep below is one long
// dependency chain per vector, and the IEEE division sequence was most of its length.
__device__ __forceinline__ double rcp_newton(double x)
{
	double r;
	asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
	double e = fma(-x, r, 1.0);
	r = fma(r, e, r);
	e = fma(-x, r, 1.0);
	return fma(r, e, r);
}

#define QIDX(a, b) ((a) * k - (a) * ((a) - 1) / 2 + (b) - (a))
// One warp: pairwise probabilities r[i][j] (i < j) in Q's upper triangle -> multiclass_probability -> prob_out / label_out.
__device__ __forceinline__ void svm_couple_warp(const SvmDev &m, double *Q, double *ps, int k, int lane, int v, double *__restrict__ label_out,
                                                double *__restrict__ prob_out)
{
	// multiclass_probability (src/svm.cpp:1829-1890): Q[t][t] = sum_{j != t} r[j][t]^2, Q[t][j] = -r[j][t] r[t][j].
	// p and Qp live in registers (3 slots per lane: t = lane, lane+32, lane+64); the Gauss-Seidel sweep broadcasts
	// Qp[t] by shuffle, so the inner loop has no shared-memory writes and no barriers; 1/Q[t][t] is computed once per
	// vector and 1/(1+diff) once per step (Newton reciprocal) instead of dividing every element (results differ from
	// libsvm's in the last bits only).
	double qtt[3], iqt[3], pr_[3], qp[3];
#pragma unroll
	for (int sl = 0; sl < 3; sl++) {
		const int t = lane + 32 * sl;
		double q = 0.0;
		if (t < k) for (int j = 0; j < k; j++) if (j != t) { const double rjt = (j < t) ? Q[QIDX(j, t)] : 1.0 - Q[QIDX(t, j)]; q = fma(rjt, rjt, q); }
		qtt[sl] = q; iqt[sl] = (t < k) ? 1.0 / q : 0.0; pr_[sl] = (t < k) ? 1.0 / k : 0.0; qp[sl] = 0.0;
	}
	__syncwarp();
	for (int i = 0; i < k; i++)
		for (int j = i + 1 + lane; j < k; j += 32) {
			const double sij = Q[QIDX(i, j)];
			Q[QIDX(i, j)] = -((1.0 - sij) * sij);
		}
#pragma unroll
	for (int sl = 0; sl < 3; sl++) { const int t = lane + 32 * sl; if (t < k) { Q[QIDX(t, t)] = qtt[sl]; ps[t] = pr_[sl]; } }
	__syncwarp();

	const int max_iter = max(100, k);
	const double eps = 0.005 / k;
	for (int iter = 0; iter < max_iter; iter++) {
		double part = 0.0, err = 0.0;
#pragma unroll
		for (int sl = 0; sl < 3; sl++) {
			const int t = lane + 32 * sl;
			double sacc = 0.0;
			if (t < k) {
				for (int j = 0; j < t; j++) sacc = fma(Q[QIDX(j, t)], ps[j], sacc);
				const double *row = Q + QIDX(t, t) - t;   // row[j] = Q[t][j] for j >= t
				for (int j = t; j < k; j++) sacc = fma(row[j], ps[j], sacc);
			}
			qp[sl] = sacc;
			part = fma(pr_[sl], sacc, part);
		}
		double pQp = part;
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) pQp += __shfl_xor_sync(0xFFFFFFFFu, pQp, o);
#pragma unroll
		for (int sl = 0; sl < 3; sl++) if (lane + 32 * sl < k) err = fmax(err, fabs(qp[sl] - pQp));
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) err = fmax(err, __shfl_xor_sync(0xFFFFFFFFu, err, o));
		if (err < eps) break;
		for (int t = 0; t < k; t++) {
			const int sl_t = t >> 5, owner = t & 31;
			const double qpt = __shfl_sync(0xFFFFFFFFu, sl_t == 0 ? qp[0] : (sl_t == 1 ? qp[1] : qp[2]), owner);
			const double qt = __shfl_sync(0xFFFFFFFFu, sl_t == 0 ? qtt[0] : (sl_t == 1 ? qtt[1] : qtt[2]), owner);
			const double iq = __shfl_sync(0xFFFFFFFFu, sl_t == 0 ? iqt[0] : (sl_t == 1 ? iqt[1] : iqt[2]), owner);
			const double diff = (pQp - qpt) * iq;
			const double inv = rcp_newton(1.0 + diff);
			pQp = (pQp + diff * fma(diff, qt, 2.0 * qpt)) * inv * inv;
#pragma unroll
			for (int sl = 0; sl < 3; sl++) {
				const int j = lane + 32 * sl;
				if (j < k) {
					const double qtj = (j < t) ? Q[QIDX(j, t)] : Q[QIDX(t, j)];
					qp[sl] = fma(diff, qtj, qp[sl]) * inv;
					pr_[sl] = ((j == t) ? pr_[sl] + diff : pr_[sl]) * inv;
				}
			}
		}
		__syncwarp();
#pragma unroll
		for (int sl = 0; sl < 3; sl++) { const int t = lane + 32 * sl; if (t < k) ps[t] = pr_[sl]; }
		__syncwarp();
	}
	__syncwarp();
#pragma unroll
	for (int sl = 0; sl < 3; sl++) { const int t = lane + 32 * sl; if (t < k) { ps[t] = pr_[sl]; prob_out[(size_t)v * k + t] = pr_[sl]; } }
	__syncwarp();
	if (lane == 0) {
		int best = 0;
		for (int t = 1; t < k; t++) if (ps[t] > ps[best]) best = t;
		label_out[v] = (double)m.label[best];
	}
}
#undef QIDX

__global__ void __launch_bounds__(PROB_WARPS * 32) k_svm_prob(SvmDev m, const double *__restrict__ kv, int n, int v_out0,
                                                              double *__restrict__ label_out, double *__restrict__ prob_out)
{
	extern __shared__ __align__(16) double dsm[];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int k = m.nr_class;
	const int tri = k * (k + 1) / 2;                          // upper triangle incl. diagonal: idx(a<=b) = a*k - a*(a-1)/2 + b - a
	double *Q = dsm + (size_t)warp * ((size_t)tri + MAXK);    // first r[i][j] (i<j; r[j][i] = 1 - r[i][j]), then Q in place
	double *ps = Q + tri;                                     // p, shared copy for the matrix-vector product
#define QIDX(a, b) ((a) * k - (a) * ((a) - 1) / 2 + (b) - (a))
	const int v = blockIdx.x * PROB_WARPS + warp;
	if (v >= n) return;
	const double *kvv = kv + (size_t)v * m.ldk;

	// pairwise decision values (svm_predict_values, src/svm.cpp:2527-2551, same summation order per pair)
	// -> sigmoid_predict -> clamp [1e-7, 1-1e-7]
	for (int i = 0; i < k; i++) {
		for (int j = i + 1 + lane; j < k; j += 32) {
			const int pidx = i * k - i * (i + 1) / 2 + (j - i - 1);
			const double *c1 = m.coef + (size_t)(j - 1) * m.l, *c2 = m.coef + (size_t)i * m.l;
			const int si = m.start[i], sj = m.start[j], ci = m.nsv[i], cj = m.nsv[j];
			double sum = 0.0;
			for (int q = 0; q < ci; q++) sum = fma(c1[si + q], kvv[si + q], sum);
			for (int q = 0; q < cj; q++) sum = fma(c2[sj + q], kvv[sj + q], sum);
			sum -= m.rho[pidx];
			double pr = sigmoid_predict_dev(sum, m.probA[pidx], m.probB[pidx]);
			const double lo = 1e-7;
			pr = fmin(fmax(pr, lo), 1.0 - lo);
			Q[QIDX(i, j)] = pr;
		}
	}
	__syncwarp();

	svm_couple_warp(m, Q, ps, k, lane, v_out0 + v, label_out, prob_out);
#undef QIDX
}

// ---------------------------------------------------------------------------------------------
// k_svm_decide + k_svm_couple : decision values and probabilities, two kernels.
// The pairwise decision value of (i, j), i < j, is S_i[j'] + S_j[i'] - rho with S_c[o'] = sum over the support vectors s of
// class c of coef[o'][s] * K[s]  (svm_predict_values, src/svm.cpp:2527-2551; o' = the row of sv_coef that faces the other
// class).  Per class block that is a small matrix product [vectors x nsv_c] x [nsv_c x (k-1)]:
//   k_svm_decide : CTA = 64 vectors x a group of class blocks, 4 x 4 outputs per thread from shared memory (coefficients
//                  from the SV-major copy, contiguous per block).  Output S[vector][class c][o'] -- the term of every pair
//                  (c, other) that sums over c's support vectors -- as whole 32-byte sectors; no atomics, no zero fill.
//   k_svm_couple : one warp per vector: dec(i, j) = S[i][j'] + S[j][i'] - rho, Platt sigmoid, clamp, Wu-Lin-Weng coupling.
// (k_svm_prob, one warp per vector for everything, reads 2 x 2080 x ~30 coefficients per VECTOR through L2, uncoalesced;
//  a fused 8-vectors-per-CTA variant was latency-bound at one CTA per SM: 430 us per CTA whatever the grid, profiles/README.)
// ---------------------------------------------------------------------------------------------
constexpr int DEC_V = 64, DEC_Q = 32, DEC_O = 64, DEC_CG = 5;
constexpr int DEC_KLD = DEC_Q + 2;                                            // row stride of the K stage (doubles): 16-byte aligned pairs, rows 4 apart on different banks
constexpr int DEC_KS = DEC_V * DEC_KLD, DEC_CS = DEC_Q * DEC_O;               // doubles per stage
constexpr size_t DEC_SMEM = 2 * (size_t)(DEC_KS + DEC_CS) * sizeof(double);   // two stages

__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gsrc)
{
	asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc)
{
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}

// grid (class groups, vector tiles, output tiles): the CTAs of one vector tile are neighbours in launch order, so a vector's
// kernel values (contiguous over the group's classes) and its R / C rows are touched together (L2).  A CTA walks its
// DEC_CG classes in chunks of DEC_Q support vectors through a two-stage cp.async ring: the next chunk is in flight
// while this one is multiplied; one barrier per chunk.  Stage layout: K[vector][q] (lanes copy along q: contiguous on both
// sides) and coef[q][output]; a thread multiplies two q at a time from 16-byte loads (4 x 4 outputs, 8 loads per 32 FMAs).
__global__ void __launch_bounds__(256, 3) k_svm_decide(SvmDev m, const double *__restrict__ kv, int n, double *__restrict__ S, int so)
{
	extern __shared__ __align__(16) double dsm[];
	__shared__ int s_nsv[DEC_CG], s_start[DEC_CG];
	const int tid = threadIdx.x, tv = tid >> 4, to = tid & 15;
	const int k = m.nr_class, k1 = k - 1;
	const int c_begin = blockIdx.x * DEC_CG, c_end = min(k, c_begin + DEC_CG), o0 = blockIdx.z * DEC_O;
	const int v0 = blockIdx.y * DEC_V;
	if (tid < c_end - c_begin) { s_nsv[tid] = m.nsv[c_begin + tid]; s_start[tid] = m.start[c_begin + tid]; }
	__syncthreads();
	const bool wide = (k1 & 1) == 0 && o0 + DEC_O <= k1;     // coefficient rows can be copied 16 bytes at a time
	// this thread's K rows: vv = (tid >> 5) + 8 i, clamped to the last vector (rows past n are computed and dropped)
	const double *krow = kv + (size_t)min(v0 + (tid >> 5), n - 1) * m.ldk + (tid & 31);
	const size_t kstep = (size_t)8 * m.ldk;
	const int vlast = n - 1 - v0 - (tid >> 5);               // rows i with 8 i > vlast are past the end

	auto issue = [&](int c, int q0, int buf) {
		double *Ks = dsm + (size_t)buf * (DEC_KS + DEC_CS), *Cs = Ks + DEC_KS;
		const int nq = min(DEC_Q, s_nsv[c - c_begin] - q0), sq = s_start[c - c_begin] + q0;
		if ((tid & 31) < nq) {
			double *dst = Ks + (tid >> 5) * DEC_KLD + (tid & 31);
#pragma unroll
			for (int i = 0; i < DEC_V / 8; i++) cp_async8(dst + i * 8 * DEC_KLD, (8 * i <= vlast ? krow + i * kstep : krow) + sq);
		}
		if (wide) {
			const double *src = m.coefT + (size_t)sq * k1 + o0;
#pragma unroll
			for (int i = 0; i < DEC_Q * DEC_O / 512; i++) {
				const int idx = tid + i * 256, q = idx >> 5, o2 = (idx & 31) * 2;
				if (q < nq) cp_async16(&Cs[q * DEC_O + o2], src + (size_t)q * k1 + o2);
			}
		} else {
#pragma unroll 2
			for (int i = 0; i < DEC_Q * DEC_O / 256; i++) {
				const int idx = tid + i * 256, q = idx >> 6, o = idx & 63;
				if (q < nq) cp_async8(&Cs[q * DEC_O + o], m.coefT + (size_t)(sq + q) * k1 + min(o0 + o, k1 - 1));
			}
		}
		asm volatile("cp.async.commit_group;" ::: "memory");
	};

	double acc[4][4];
#pragma unroll
	for (int i = 0; i < 4; i++)
#pragma unroll
		for (int j = 0; j < 4; j++) acc[i][j] = 0.0;
	int c = c_begin, q0 = 0, buf = 0;
	if (c < c_end) issue(c, 0, 0);
	while (c < c_end) {
		const int ncls = s_nsv[c - c_begin];
		const int nq = min(DEC_Q, ncls - q0);
		const bool last_chunk = q0 + DEC_Q >= ncls;
		const int c_next = last_chunk ? c + 1 : c, q_next = last_chunk ? 0 : q0 + DEC_Q;
		asm volatile("cp.async.wait_group 0;" ::: "memory");
		__syncthreads();                                     // this chunk has landed; everybody is done with the other stage
		if (c_next < c_end) issue(c_next, q_next, buf ^ 1);
		const double *Ks = dsm + (size_t)buf * (DEC_KS + DEC_CS) + tv * 4 * DEC_KLD, *Cs = dsm + (size_t)buf * (DEC_KS + DEC_CS) + DEC_KS + to * 4;
		int q = 0;
#pragma unroll 2
		for (; q + 1 < nq; q += 2) {
			double2 a[4];
#pragma unroll
			for (int i = 0; i < 4; i++) a[i] = *reinterpret_cast<const double2 *>(Ks + i * DEC_KLD + q);
			const double2 b01 = *reinterpret_cast<const double2 *>(Cs + q * DEC_O), b23 = *reinterpret_cast<const double2 *>(Cs + q * DEC_O + 2);
			const double2 d01 = *reinterpret_cast<const double2 *>(Cs + (q + 1) * DEC_O), d23 = *reinterpret_cast<const double2 *>(Cs + (q + 1) * DEC_O + 2);
			const double b[4] = {b01.x, b01.y, b23.x, b23.y}, d[4] = {d01.x, d01.y, d23.x, d23.y};
#pragma unroll
			for (int i = 0; i < 4; i++)
#pragma unroll
				for (int j = 0; j < 4; j++) acc[i][j] = fma(a[i].y, d[j], fma(a[i].x, b[j], acc[i][j]));
		}
		if (q < nq) {
			const double2 b01 = *reinterpret_cast<const double2 *>(Cs + q * DEC_O), b23 = *reinterpret_cast<const double2 *>(Cs + q * DEC_O + 2);
			const double b[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
			for (int i = 0; i < 4; i++) {
				const double ai = Ks[i * DEC_KLD + q];
#pragma unroll
				for (int j = 0; j < 4; j++) acc[i][j] = fma(ai, b[j], acc[i][j]);
			}
		}
		if (last_chunk) {
			// class c is complete: its row of S (the term of every pair (c, other) that sums over c's support vectors), 32 bytes per thread and vector
#pragma unroll
			for (int i = 0; i < 4; i++) {
				const int v = v0 + tv * 4 + i;
				if (v < n && o0 + to * 4 < so) {
					double2 *dst = reinterpret_cast<double2 *>(S + ((size_t)v * k + c) * so + o0 + to * 4);
					dst[0] = make_double2(acc[i][0], acc[i][1]);
					dst[1] = make_double2(acc[i][2], acc[i][3]);
				}
#pragma unroll
				for (int j = 0; j < 4; j++) acc[i][j] = 0.0;
			}
		}
		c = c_next; q0 = q_next; buf ^= 1;
	}
}

// One warp per vector, two warps per CTA.  Shared memory per warp: the FULL symmetric matrix Q (row stride ks odd: a lane
// walking its own row and 32 lanes reading one row are both conflict-free), then p and 1/Q[t][t].
//   1. the vector's 65 x 64 decision terms S -> Q (coalesced rows); r[i][j] = clamp(sigmoid((Q[i][j] + Q[j][i] - rho) * A + B)),
//      r[j][i] = 1 - r[i][j], flat over the pairs
//   2. Q[t][t] = sum_{j != t} r[j][t]^2 (column t, libsvm's order);  Q[i][j] = Q[j][i] = -r[i][j] r[j][i] in place
//   3. Wu-Lin-Weng iteration (multiclass_probability, src/svm.cpp:1829-1890): Qp from scratch (each lane its rows),
//      convergence test, then the Gauss-Seidel sweep.  The sweep is one dependency chain of k steps; per step: one shuffle
//      (Qp[t] from its owner), one broadcast load (1/Q[t][t]), a Newton reciprocal, and per lane three fused updates from
//      row t.  diff * Q[t][t] is replaced by the numerator it came from (pQp - Qp[t]); with the reciprocals this differs
//      from libsvm in the last bits only.
// (The packed triangle of the previous version held twelve vectors per SM instead of six but spent 2/3 of its instructions
//  on index selects: 41 k warp instructions per vector against ~15 k here; profiles/README.)
constexpr int CPL_WARPS = 2;

__global__ void __launch_bounds__(CPL_WARPS * 32) k_svm_couple(SvmDev m, const double *__restrict__ S, int so, int n, int v_out0,
                                                               double *__restrict__ label_out, double *__restrict__ prob_out)
{
	extern __shared__ __align__(16) double dsm[];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int k = m.nr_class, np = k * (k - 1) / 2, kp = (k + 7) & ~7, ks = k | 1;
	double *Q = dsm + (size_t)warp * ((size_t)ks * k + 2 * kp);
	double *ps = Q + (size_t)ks * k, *iqs = ps + kp;
	const int v = blockIdx.x * CPL_WARPS + warp;
	if (v >= n) return;
	// S[v][c][o'] -> Q[c][o]: the two terms of pair (i, j) land at Q[i][j] and Q[j][i] (coalesced rows, one store each)
	const double *Sv = S + (size_t)v * k * so;
	// (eight rows = up to 24 independent loads per lane in flight: the sweep is pure latency otherwise)
	for (int c0 = 0; c0 < k; c0 += 8) {
		double sv_[8][3];
#pragma unroll
		for (int u = 0; u < 8; u++)
#pragma unroll
			for (int sl = 0; sl < 3; sl++) {
				const int c = c0 + u, o1 = lane + 32 * sl;
				sv_[u][sl] = (c < k && o1 < k - 1) ? Sv[c * so + o1] : 0.0;
			}
#pragma unroll
		for (int u = 0; u < 8; u++)
#pragma unroll
			for (int sl = 0; sl < 3; sl++) {
				const int c = c0 + u, o1 = lane + 32 * sl;
				if (c < k && o1 < k - 1) Q[c * ks + o1 + (o1 >= c ? 1 : 0)] = sv_[u][sl];
			}
	}
	__syncwarp();
#pragma unroll 4
	for (int p = lane; p < np; p += 32) {
		const int ij0 = m.pair_ij[p], i0 = ij0 >> 8, j0 = ij0 & 255;
		const double dec = (Q[i0 * ks + j0] + Q[j0 * ks + i0]) - m.rho[p];
		const double f = __dadd_rn(__dmul_rn(dec, m.probA[p]), m.probB[p]);   // sigmoid_predict, src/svm.cpp:1818-1826
		const double e = exp_nonpos(-fabs(f));
		const double r1 = rcp_newton(1.0 + e);
		const double pr = fmin(fmax((f >= 0) ? e * r1 : r1, 1e-7), 1.0 - 1e-7);   // src/svm.cpp:2606-2611
		Q[i0 * ks + j0] = pr;
		Q[j0 * ks + i0] = 1.0 - pr;
	}
	int tt[3];
	double qtt[3], pr_[3], qp[3];
#pragma unroll
	for (int sl = 0; sl < 3; sl++) {
		const int t = lane + 32 * sl;
		tt[sl] = min(t, k - 1);                                  // lanes past k shadow the last class and are masked out
		if (t < k) Q[t * ks + t] = 0.0;
		qtt[sl] = 0.0; qp[sl] = 0.0;
		pr_[sl] = (t < k) ? 1.0 / k : 0.0;
	}
	__syncwarp();
	for (int j = 0; j < k; j++) {
		const double *row = Q + j * ks;
		const double r0 = row[tt[0]], r1 = row[tt[1]], r2 = row[tt[2]];   // the diagonal holds 0: j == t adds nothing
		qtt[0] = fma(r0, r0, qtt[0]); qtt[1] = fma(r1, r1, qtt[1]); qtt[2] = fma(r2, r2, qtt[2]);
	}
	__syncwarp();
#pragma unroll 4
	for (int p = lane; p < np; p += 32) {
		const int ij = m.pair_ij[p], i = ij >> 8, j = ij & 255;
		const double sij = Q[i * ks + j];
		const double qv = -((1.0 - sij) * sij);
		Q[i * ks + j] = qv;
		Q[j * ks + i] = qv;
	}
#pragma unroll
	for (int sl = 0; sl < 3; sl++) { const int t = lane + 32 * sl; if (t < k) { Q[t * ks + t] = qtt[sl]; ps[t] = pr_[sl]; iqs[t] = 1.0 / qtt[sl]; } }
	__syncwarp();

	const int max_iter = max(100, k);
	const double eps = 0.005 / k;
	const double *row0 = Q + tt[0] * ks, *row1 = Q + tt[1] * ks, *row2 = Q + tt[2] * ks;
	for (int iter = 0; iter < max_iter; iter++) {
		double s0 = 0.0, s1 = 0.0, s2 = 0.0;
#pragma unroll 5
		for (int j = 0; j < k; j++) {
			const double pj = ps[j];
			s0 = fma(row0[j], pj, s0); s1 = fma(row1[j], pj, s1); s2 = fma(row2[j], pj, s2);
		}
		qp[0] = s0; qp[1] = s1; qp[2] = s2;
		double pQp = fma(pr_[2], s2, fma(pr_[1], s1, pr_[0] * s0));
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) pQp += __shfl_xor_sync(0xFFFFFFFFu, pQp, o);
		double err = 0.0;
#pragma unroll
		for (int sl = 0; sl < 3; sl++) if (lane + 32 * sl < k) err = fmax(err, fabs(qp[sl] - pQp));
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) err = fmax(err, __shfl_xor_sync(0xFFFFFFFFu, err, o));
		if (err < eps) break;
#pragma unroll
		for (int SL = 0; SL < 3; SL++) {
			const int t_end = min(k, 32 * SL + 32);
			for (int t = 32 * SL; t < t_end; t++) {
				const double qpt = __shfl_sync(0xFFFFFFFFu, qp[SL], t - 32 * SL);
				const double *rowt = Q + t * ks;
				const double q0 = rowt[tt[0]], q1 = rowt[tt[1]], q2 = rowt[tt[2]];
				const double num = pQp - qpt;
				const double diff = num * iqs[t];
				const double inv = rcp_newton(1.0 + diff);
				pQp = (pQp + diff * (num + 2.0 * qpt)) * inv * inv;
				if (lane == t - 32 * SL) pr_[SL] += diff;
				qp[0] = fma(diff, q0, qp[0]) * inv; qp[1] = fma(diff, q1, qp[1]) * inv; qp[2] = fma(diff, q2, qp[2]) * inv;
				pr_[0] *= inv; pr_[1] *= inv; pr_[2] *= inv;
			}
		}
#pragma unroll
		for (int sl = 0; sl < 3; sl++) { const int t = lane + 32 * sl; if (t < k) ps[t] = pr_[sl]; }
		__syncwarp();
	}
	// argmax, first maximum wins (svm_predict_probability, src/svm.cpp:2620-2626)
	double bv = -1.0; int bi = 0;
#pragma unroll
	for (int sl = 0; sl < 3; sl++) {
		const int t = lane + 32 * sl;
		if (t < k) { prob_out[(size_t)(v_out0 + v) * k + t] = pr_[sl]; if (pr_[sl] > bv) { bv = pr_[sl]; bi = t; } }
	}
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) {
		const double ov = __shfl_xor_sync(0xFFFFFFFFu, bv, o);
		const int oi = __shfl_xor_sync(0xFFFFFFFFu, bi, o);
		if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
	}
	if (lane == 0) label_out[v_out0 + v] = (double)m.label[bi];
}

// vectors per pass: bounds the workspace (K, R, C: ~48 KB per vector for OCR.model) whatever the batch
constexpr int SVM_PASS = 32768;

size_t svm_ws_bytes(const SvmDev &m, int n)
{
	const size_t c = (size_t)min(n, SVM_PASS);
	return sizeof(double) * c * ((size_t)m.ldk + (size_t)m.nr_class * (size_t)((m.nr_class + 2) & ~3)) + 1024;
}

int launch_svm_predict(const SvmDev &m, const double *x_f64, const uint8_t *x_u8, int n, double *ws, double *label, double *prob,
                       cudaStream_t st, uint8_t *tc_ws)
{
	if (n <= 0) return 0;
	if (m.nr_class > MAXK) { set_error("svm: nr_class %d > %d unsupported", m.nr_class, MAXK); return -1; }
	const size_t tri = (size_t)m.nr_class * (m.nr_class + 1) / 2;
	const bool tc = x_u8 && m.svj && tc_ws && m.dims <= TC_KPAD && m.l <= TC_NPAD;
	uint8_t *xp = tc_ws;
	uint32_t *xx = tc ? reinterpret_cast<uint32_t *>(tc_ws + (((size_t)n * TC_KPAD + 255) / 256) * 256) : nullptr;
	if (tc) {
		ERT_CUDA_CHECK(cudaMemsetAsync(svm_tc_flag(tc_ws, n), 0, sizeof(uint32_t), st));
		k_svm_prep_x<<<(n + 3) / 4, 128, 0, st>>>(x_u8, n, m.dims, xp, xx);
		ERT_CUDA_CHECK(cudaGetLastError());
	}
	const size_t smem_tc = 16384 + 2 * 32768;
	const size_t smem_q = (size_t)PROB_WARPS * (tri + MAXK) * sizeof(double);
	const size_t smem_c = (size_t)CPL_WARPS * ((size_t)(m.nr_class | 1) * m.nr_class + 2 * (size_t)((m.nr_class + 7) & ~7)) * sizeof(double);
	ERT_CUDA_CHECK(cudaFuncSetAttribute(k_svm_kvalue_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tc));
	ERT_CUDA_CHECK(cudaFuncSetAttribute(k_svm_prob, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
	ERT_CUDA_CHECK(cudaFuncSetAttribute(k_svm_couple, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
	ERT_CUDA_CHECK(cudaFuncSetAttribute(k_svm_decide, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DEC_SMEM));
	const int cap = min(n, SVM_PASS);
	const int so = (m.nr_class + 2) & ~3;                  // row length of S: nr_class - 1 rounded up to 4 doubles
	double *kv = ws, *S = ws + (size_t)cap * m.ldk;
	for (int v0 = 0; v0 < n; v0 += SVM_PASS) {
		const int nc = min(SVM_PASS, n - v0);
		if (tc && m.tc_variant != 2) {
			if (launch_svm_kvalue_tma(m, xp, xx, n, v0, nc, kv, svm_tc_flag(tc_ws, n), st)) return -1;
		} else if (tc) {
			dim3 g2(TC_NPAD / TC_N, (nc + TC_M - 1) / TC_M);
			k_svm_kvalue_tc<<<g2, 128, smem_tc, st>>>(xp + (size_t)v0 * TC_KPAD, xx + v0, nc, m.svj, m.sve, m.ss, m.l, m.ldk, m.gamma, m.inv_s255, kv);
		} else {
			dim3 grid((m.l + KT - 1) / KT, (nc + KT - 1) / KT);
			if (x_u8) k_svm_kvalue<uint8_t><<<grid, 256, 0, st>>>(m, x_u8 + (size_t)v0 * m.dims, nc, kv);
			else k_svm_kvalue<double><<<grid, 256, 0, st>>>(m, x_f64 + (size_t)v0 * m.dims, nc, kv);
		}
		ERT_CUDA_CHECK(cudaGetLastError());
		if (m.coefT && m.pair_ij && !m.legacy_prob) {
			dim3 g3((m.nr_class + DEC_CG - 1) / DEC_CG, (nc + DEC_V - 1) / DEC_V, (m.nr_class - 1 + DEC_O - 1) / DEC_O);
			k_svm_decide<<<g3, 256, DEC_SMEM, st>>>(m, kv, nc, S, so);
			ERT_CUDA_CHECK(cudaGetLastError());
			k_svm_couple<<<(nc + CPL_WARPS - 1) / CPL_WARPS, CPL_WARPS * 32, smem_c, st>>>(m, S, so, nc, v0, label, prob);
		} else {
			k_svm_prob<<<(nc + PROB_WARPS - 1) / PROB_WARPS, PROB_WARPS * 32, smem_q, st>>>(m, kv, nc, v0, label, prob);
		}
		ERT_CUDA_CHECK(cudaGetLastError());
	}
	return 0;
}

size_t svm_tc_ws_bytes(int n) { return (((size_t)n * TC_KPAD + 255) / 256) * 256 + (size_t)n * 4 + 256; }
uint32_t *svm_tc_flag(uint8_t *tc_ws, int n) { return reinterpret_cast<uint32_t *>(tc_ws + (((size_t)n * TC_KPAD + 255) / 256) * 256) + n; }
int svm_tc_kpad() { return TC_KPAD; }
int svm_tc_npad() { return TC_NPAD; }

} // namespace ert
