// svm_gemm.cu -- the RBF kernel values K[v][s] = exp(-gamma |x_v - sv_s|^2) of a batch of u8 feature vectors on the 5th-gen
// tensor cores: the distance is two integer GEMMs (svm.cu, "Tensor-core formulation": DJ = X.J^T exact u8 x u8, DE = X.E^T
// u8 x s8 for the sub-1/255 residue of the model's decimals), [n x 1920] x [1920 x 1920] each, accumulated in TMEM.
// (replaces Kernel::k_function RBF, src/svm.cpp:325-365, as called per support vector by svm_predict_values :2516-2518)
//
// k_svm_kvalue_tma: persistent, warp-specialised, one CTA per SM (320 threads):
//   warp 0      TMA producer: per 128-byte K chunk three tensor-map boxes (A 128 rows, B_J and B_E 128 rows each, SWIZZLE_128B)
//               into a 4-stage ring (48 KB per stage), mbarrier expect_tx / complete_tx
//   warp 1      MMA issuer: per stage 4 + 4 tcgen05.mma kind::i8 (M 128, N 128, K 32), tcgen05.commit frees the stage;
//               the two accumulators of a tile are 256 TMEM columns, double-buffered (512 columns), so the next tile's
//               main loop runs under this tile's epilogue
//   warps 2..9  epilogue (two warps per TMEM lane quarter, 64 columns each): tcgen05.ld 32 lanes x 16 columns, d^2 = |x|^2 + |sv|^2 - 2 (DJ / 255^2 + DE / (255 S)), exp in
//               FP64, 128-byte row segments to K
// Tiles: 128 vectors x 128 support vectors, support-vector tile fastest, so the CTAs in flight share a few A tiles and
// the 7.4 MB of B through L2.  Rows past n are zero-filled by TMA and never stored.
// Every mbarrier wait is bounded; a wait that gives up sets *flag (the caller reports it) instead of hanging the GPU.
#include "common.cuh"
#include "kernels.h"
#include "svm_math.cuh"
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cstring>

namespace ert {

namespace g2 {
constexpr int BM = 128, BN = 128, BK = 128;                 // tile; BK in bytes = one SWIZZLE_128B atom
constexpr int STAGES = 4, EPI_WARPS = 8, NT = 64 + EPI_WARPS * 32;
constexpr int A_BYTES = BM * BK, B_BYTES = BN * BK, STAGE_BYTES = A_BYTES + 2 * B_BYTES;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024;     // + slack for the 1024-byte alignment SWIZZLE_128B wants
constexpr int KPAD = 1920, NPAD = 2048;                     // same padding as svm.cu (TC_KPAD, TC_NPAD)
constexpr uint32_t SPIN = 1u << 24;
}

__device__ __forceinline__ uint32_t g2_smem(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ bool g2_wait(uint32_t bar, uint32_t parity)
{
	uint32_t ok = 0;
	for (uint32_t spins = 0; !ok && spins < g2::SPIN; ++spins)
		asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
		             : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
	return ok != 0;
}

// UMMA shared-memory matrix descriptor: K-major, SWIZZLE_128B (layout type 2), 8-row groups 1024 bytes apart, version 1
__device__ __forceinline__ uint64_t g2_desc(uint32_t smem_addr)
{
	return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(1024u >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

__device__ __forceinline__ void g2_tma_2d(uint32_t dst, const CUtensorMap *tm, int c0, int c1, uint32_t bar)
{
	asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
	             ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(bar) : "memory");
}

__global__ void __launch_bounds__(g2::NT, 1) k_svm_kvalue_tma(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmJ,
                                                             const __grid_constant__ CUtensorMap tmE, const uint32_t *__restrict__ xx, int row0, int n,
                                                             const double *__restrict__ ss, int l, int ldk, double gamma, double inv_s255,
                                                             double *__restrict__ kv, uint32_t *__restrict__ flag)
{
	using namespace g2;
	extern __shared__ uint8_t g2_raw[];
	__shared__ __align__(8) uint64_t bar_full[STAGES], bar_empty[STAGES], bar_tfull[2], bar_tempty[2];
	__shared__ uint32_t s_tmem;
	const uint32_t smem0 = (g2_smem(g2_raw) + 1023u) & ~1023u;
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const int n_tiles_n = (l + BN - 1) / BN, n_tiles_m = (n + BM - 1) / BM, n_tiles = n_tiles_m * n_tiles_n;
	constexpr int KCH = KPAD / BK;

	if (tid == 0) {
		for (int i = 0; i < STAGES; i++) {
			asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(g2_smem(&bar_full[i])), "r"(1u) : "memory");
			asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(g2_smem(&bar_empty[i])), "r"(1u) : "memory");
		}
		for (int i = 0; i < 2; i++) {
			asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(g2_smem(&bar_tfull[i])), "r"(1u) : "memory");
			asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(g2_smem(&bar_tempty[i])), "r"((uint32_t)EPI_WARPS) : "memory");   // one arrival per epilogue warp
		}
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	if (warp == 1) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(g2_smem(&s_tmem)), "r"(512u) : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	const uint32_t tmem = s_tmem;

	if (warp == 0) {
		// ---------------- TMA producer ----------------
		if (lane == 0) {
			uint32_t stage = 0, phase = 0;
			bool ok = true;
			for (int tile = blockIdx.x; tile < n_tiles && ok; tile += gridDim.x) {
				const int tm_ = tile / n_tiles_n, tn = tile - tm_ * n_tiles_n;
				for (int kc = 0; kc < KCH && ok; kc++) {
					ok = g2_wait(g2_smem(&bar_empty[stage]), phase ^ 1u);
					const uint32_t sA = smem0 + stage * STAGE_BYTES, sJ = sA + A_BYTES, sE = sJ + B_BYTES, fb = g2_smem(&bar_full[stage]);
					asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fb), "r"((uint32_t)STAGE_BYTES) : "memory");
					g2_tma_2d(sA, &tmA, kc * BK, row0 + tm_ * BM, fb);
					g2_tma_2d(sJ, &tmJ, kc * BK, tn * BN, fb);
					g2_tma_2d(sE, &tmE, kc * BK, tn * BN, fb);
					if (++stage == STAGES) { stage = 0; phase ^= 1u; }
				}
			}
			if (!ok) atomicOr(flag, 1u);
		}
	} else if (warp == 1) {
		// ---------------- MMA issuer ----------------
		if (lane == 0) {
			// instruction descriptor: D = S32, A = u8, B = u8 / s8, both K-major, N = 128, M = 128
			const uint32_t idesc_base = (2u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
			const uint32_t idesc_j = idesc_base, idesc_e = idesc_base | (1u << 10);
			uint32_t stage = 0, phase = 0;
			bool ok = true;
			int t = 0;
			for (int tile = blockIdx.x; tile < n_tiles && ok; tile += gridDim.x, ++t) {
				const uint32_t buf = (uint32_t)t & 1u, use = (uint32_t)t >> 1;
				ok = g2_wait(g2_smem(&bar_tempty[buf]), (use & 1u) ^ 1u);           // the epilogue has drained this accumulator pair
				asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
				const uint32_t dJ = tmem + buf * 256u, dE = dJ + 128u;
				for (int kc = 0; kc < KCH && ok; kc++) {
					ok = g2_wait(g2_smem(&bar_full[stage]), phase);
					asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
					const uint32_t sA = smem0 + stage * STAGE_BYTES, sJ = sA + A_BYTES, sE = sJ + B_BYTES;
#pragma unroll
					for (int k4 = 0; k4 < BK / 32; ++k4) {
						const uint64_t da = g2_desc(sA + k4 * 32), dj = g2_desc(sJ + k4 * 32), de = g2_desc(sE + k4 * 32);
						const uint32_t acc = (kc > 0 || k4 > 0) ? 1u : 0u;
						asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
						             "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
						             ::"r"(dJ), "l"(da), "l"(dj), "r"(idesc_j), "r"(acc) : "memory");
						asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
						             "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
						             ::"r"(dE), "l"(da), "l"(de), "r"(idesc_e), "r"(acc) : "memory");
					}
					// frees the stage once the MMAs above have read it (commit implies fence::before_thread_sync)
					asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(g2_smem(&bar_empty[stage])) : "memory");
					if (++stage == STAGES) { stage = 0; phase ^= 1u; }
				}
				asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(g2_smem(&bar_tfull[buf])) : "memory");
			}
			if (!ok) atomicOr(flag, 2u);
		}
	} else {
		// ---------------- epilogue: warp w may touch TMEM lanes 32 (w % 4) .. +31; two warps per lane quarter split the columns ----------------
		const int q = warp & 3, half = (warp - 2) >> 2;
		constexpr int CW = BN / (EPI_WARPS / 4);                 // columns per warp
		bool ok = true;
		int t = 0;
		for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++t) {
			const int tm_ = tile / n_tiles_n, tn = tile - tm_ * n_tiles_n;
			const uint32_t buf = (uint32_t)t & 1u, use = (uint32_t)t >> 1;
			if (ok) ok = g2_wait(g2_smem(&bar_tfull[buf]), use & 1u);
			asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
			const int row = tm_ * BM + q * 32 + lane;
			const double xxr = (row < n) ? (double)xx[row0 + row] / 65025.0 : 0.0;
			const uint32_t tJ = tmem + buf * 256u + ((uint32_t)(q * 32) << 16), tE = tJ + 128u;
			double *out = kv + (size_t)row * ldk + tn * BN;      // ldk is a multiple of BN: whole tiles, 16-byte aligned rows
			const double *sst = ss + tn * BN;                   // |sv|^2, padded to the tile grid
#pragma unroll 1
			for (int c0 = half * CW; c0 < half * CW + CW; c0 += 16) {
				uint32_t rj[16], re[16];
				asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
				             : "=r"(rj[0]), "=r"(rj[1]), "=r"(rj[2]), "=r"(rj[3]), "=r"(rj[4]), "=r"(rj[5]), "=r"(rj[6]), "=r"(rj[7]),
				               "=r"(rj[8]), "=r"(rj[9]), "=r"(rj[10]), "=r"(rj[11]), "=r"(rj[12]), "=r"(rj[13]), "=r"(rj[14]), "=r"(rj[15])
				             : "r"(tJ + (uint32_t)c0) : "memory");
				asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
				             : "=r"(re[0]), "=r"(re[1]), "=r"(re[2]), "=r"(re[3]), "=r"(re[4]), "=r"(re[5]), "=r"(re[6]), "=r"(re[7]),
				               "=r"(re[8]), "=r"(re[9]), "=r"(re[10]), "=r"(re[11]), "=r"(re[12]), "=r"(re[13]), "=r"(re[14]), "=r"(re[15])
				             : "r"(tE + (uint32_t)c0) : "memory");
				asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
				// 16 independent chains, no branches: the compiler interleaves them (columns past l are computed and never read)
				double kvv[16];
#pragma unroll
				for (int j = 0; j < 16; ++j) {
					const double dot = fma((double)(int32_t)rj[j], 1.0 / 65025.0, (double)(int32_t)re[j] * inv_s255);   // no division: its slow-path call would serialise the 16 chains
					const double d2 = xxr + sst[c0 + j] - 2.0 * dot;
					kvv[j] = exp_nonpos(-gamma * fmax(d2, 0.0));
				}
				if (row < n) {
#pragma unroll
					for (int j = 0; j < 16; j += 2) *reinterpret_cast<double2 *>(out + c0 + j) = make_double2(kvv[j], kvv[j + 1]);
				}
			}
			asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
			__syncwarp();
			if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(g2_smem(&bar_tempty[buf])) : "memory");
		}
		if (!ok && lane == 0) atomicOr(flag, 4u);
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

static int g2_make_map(CUtensorMap *out, const void *base, uint64_t rows, uint32_t box_rows)
{
	static PFN_cuTensorMapEncodeTiled_v12000 encode = nullptr;
	if (!encode) {
		void *fn = nullptr;
		cudaDriverEntryPointQueryResult qres;
		if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
			set_error("cuTensorMapEncodeTiled is not available from the CUDA driver");
			return -1;
		}
		encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
	}
	const cuuint64_t gdim[2] = {(cuuint64_t)g2::KPAD, (cuuint64_t)rows};
	const cuuint64_t gstride[1] = {(cuuint64_t)g2::KPAD};
	const cuuint32_t box[2] = {(cuuint32_t)g2::BK, box_rows};
	const cuuint32_t estr[2] = {1u, 1u};
	const CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void *>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
	                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d) for the svm operand (%llu rows)", (int)r, (unsigned long long)rows); return -1; }
	return 0;
}

// xp: [n_rows][1920] padded u8 features (k_svm_prep_x), xx: their squared norms; scores rows row0 .. row0 + n - 1 into kv[n][l]
int launch_svm_kvalue_tma(const SvmDev &m, const uint8_t *xp, const uint32_t *xx, int n_rows, int row0, int n, double *kv, uint32_t *flag, cudaStream_t st)
{
	static int sm_count = 0;
	if (!sm_count) {
		int dev = 0;
		ERT_CUDA_CHECK(cudaGetDevice(&dev));
		ERT_CUDA_CHECK(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
	}
	CUtensorMap tmA, tmJ, tmE;
	if (g2_make_map(&tmA, xp, (uint64_t)n_rows, g2::BM) || g2_make_map(&tmJ, m.svj, g2::NPAD, g2::BN) || g2_make_map(&tmE, m.sve, g2::NPAD, g2::BN)) return -1;
	ERT_CUDA_CHECK(cudaFuncSetAttribute(k_svm_kvalue_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, g2::SMEM_BYTES));
	const int n_tiles = ((n + g2::BM - 1) / g2::BM) * ((m.l + g2::BN - 1) / g2::BN);
	k_svm_kvalue_tma<<<min(n_tiles, sm_count), g2::NT, g2::SMEM_BYTES, st>>>(tmA, tmJ, tmE, xx, row0, n, m.ss, m.l, m.ldk, m.gamma, m.inv_s255, kv, flag);
	ERT_CUDA_CHECK(cudaGetLastError());
	return 0;
}

} // namespace ert
