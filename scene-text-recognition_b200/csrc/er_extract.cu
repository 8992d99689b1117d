// er_extract.cu -- component-tree build (replaces ERFilter::er_tree_extract, src/ER.cpp:240-413,
// and compute_channels, src/ER.cpp:114-128) as hand-written sm_100a kernels.
//
// Pipeline per batch of planes (a plane = one u8 channel image, 6 per BGR frame):
//   k_channels      BGR -> Y, Cr, Cb planes (inverted planes are derived on the fly)
//   k_tile_build    per 64x32 tile (512 threads): TMA bulk-copy the tile into shared memory, quantise to levels,
//                   build the tile-local component forest with a keyed lock-free union-find in
//                   shared memory, count own-level pixels / bbox per tile-local node, emit nodes
//   k_seam_link     stitch tiles: the same keyed union-find on the (few) edges that cross tile seams
//   k_fold          fold nodes that were merged across seams into their final node; count children
//   k_refit         bottom-up accumulation of (pixels, nodes, bbox) with arrival counters, no grid sync
//   k_emit_kept     the flood's start-pixel rule (src/ER.cpp:267-341) picks the reference's tree; compact the nodes
//                   the reference keeps (area > MIN_AREA, or the root)
//
// Exact reference semantics reproduced: level = rint_half_even(v/step); levels >= hi
// are walls; one node per (level L, 4-connected component of {level<=L} holding a level-L pixel);
// area = pixels + number of nodes in the subtree; bbox exact; kept = area > MIN_AREA or root.
#include "common.cuh"
#include "kernels.h"
#include <algorithm>

namespace ert {

// ---------------------------------------------------------------------------------------------
// k_channels : OpenCV 8-bit BGR2YCrCb in integer arithmetic (14 fractional bits)
// ---------------------------------------------------------------------------------------------
__global__ void k_channels(const uint8_t *__restrict__ bgr, size_t frame_stride, int row_stride, int W, int H,
                           uint8_t *__restrict__ ycc, int pitch)
{
	const int f = blockIdx.z;
	const int y = blockIdx.y;
	const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
	if (x4 >= W) return;
	const uint8_t *row = bgr + (size_t)f * frame_stride + (size_t)y * row_stride + (size_t)x4 * 3;
	const size_t plane_bytes = (size_t)pitch * H;
	uint8_t *o0 = ycc + ((size_t)f * 3 + 0) * plane_bytes + (size_t)y * pitch + x4;
	uint8_t *o1 = o0 + plane_bytes, *o2 = o1 + plane_bytes;
	const int n = min(4, W - x4);
	uint8_t yy[4] = {0, 0, 0, 0}, cr[4] = {0, 0, 0, 0}, cb[4] = {0, 0, 0, 0};
#pragma unroll
	for (int i = 0; i < 4; i++) {
		if (i < n) {
			const int b = __ldg(row + 3 * i), g = __ldg(row + 3 * i + 1), r = __ldg(row + 3 * i + 2);
			const int lum = (4899 * r + 9617 * g + 1868 * b + 8192) >> 14;
			int vr = ((r - lum) * 11682 + (128 << 14) + 8192) >> 14;
			int vb = ((b - lum) * 9241 + (128 << 14) + 8192) >> 14;
			vr = min(max(vr, 0), 255);
			vb = min(max(vb, 0), 255);
			yy[i] = (uint8_t)lum; cr[i] = (uint8_t)vr; cb[i] = (uint8_t)vb;
		}
	}
	// pitch is a multiple of 64 and x4 a multiple of 4: 4-byte stores are aligned (pad bytes are don't-care)
	*reinterpret_cast<uchar4 *>(o0) = make_uchar4(yy[0], yy[1], yy[2], yy[3]);
	*reinterpret_cast<uchar4 *>(o1) = make_uchar4(cr[0], cr[1], cr[2], cr[3]);
	*reinterpret_cast<uchar4 *>(o2) = make_uchar4(cb[0], cb[1], cb[2], cb[3]);
}

// ---------------------------------------------------------------------------------------------
// shared-memory keyed union-find.  local key = level << 16 | local pixel index
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t find_s(volatile uint32_t *par, uint32_t k)
{
	for (;;) {
		const uint32_t p = par[k & 0xFFFFu];
		if (p == KEY_NONE || (p >> 16) != (k >> 16)) return k;
		k = p;
	}
}

// Insert the edge a--b.  par[x] only ever decreases (atomicMin) and whatever it displaces is
// re-linked, so the final forest does not depend on the interleaving of concurrent links.
__device__ __forceinline__ void link_s(uint32_t *par, uint32_t a, uint32_t b, uint32_t *status)
{
	for (int guard = 0; guard < (1 << 20); ++guard) {
		a = find_s(par, a);
		b = find_s(par, b);
		if (a == b) return;
		if (a > b) { const uint32_t t = a; a = b; b = t; }
		const uint32_t old = atomicMin(&par[a & 0xFFFFu], b);
		if (old == b || old == KEY_NONE) return;
		if (old < b) a = old;           // a already had a closer ancestor: b must sit above it
		else { a = b; b = old; }        // b slipped in between a and its old ancestor
	}
	atomicOr(status, ERR_LOOP_GUARD);
}

// One step of a walk towards the level root with path halving.  Returns true when k is a level root.
// Halving rewrites par[k] from its same-level parent to its same-level grandparent with a plain
// store: all three are pixels of the same node, so connectivity per level is unchanged, and a
// concurrent atomicMin that the store might overwrite has already queued the re-link of what it
// displaced ("why compression is safe").
__device__ __forceinline__ bool climb_s(volatile uint32_t *par, uint32_t &k)
{
	const uint32_t p = par[k & 0xFFFFu];
	if (p == KEY_NONE || (p >> 16) != (k >> 16)) return true;
	const uint32_t g = par[p & 0xFFFFu];
	if (g != KEY_NONE && (g >> 16) == (k >> 16)) { par[k & 0xFFFFu] = g; k = g; return false; }
	k = p;          // p's own pointer leaves the level: p IS the level root -- no extra round to find that out
	return true;
}

// ---------------------------------------------------------------------------------------------
// TMA (bulk async copy) + mbarrier helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
	uint32_t ok;
	do {
		asm volatile(
			"{\n\t.reg .pred p;\n\t"
			"mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
			"selp.u32 %0, 1, 0, p;\n\t}"
			: "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
	} while (!ok);
}
__device__ __forceinline__ void tma_row_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
	             ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ---------------------------------------------------------------------------------------------
// k_tile_build
// ---------------------------------------------------------------------------------------------
template <int TW, int TH, int NT, int RF, bool CHUNK>
__global__ void __launch_bounds__(NT) k_tile_build(ExtractParams P, const PlaneSrc *__restrict__ planes,
                                                   uint32_t *__restrict__ par_g, NodeAttr *__restrict__ attr_g,
                                                   uint32_t *__restrict__ node_list, uint32_t *__restrict__ node_count,
                                                   uint32_t *status, int tiles_x, int local_union, unsigned long long *prof, uint32_t *__restrict__ ring_rec)
{
	long long t_prev = prof ? clock64() : 0;
#define ERT_PHASE(i) do { if (prof && threadIdx.x == 0) { const long long t_now = clock64(); atomicAdd(&prof[i], (unsigned long long)(t_now - t_prev)); t_prev = t_now; } } while (0)
	constexpr int TPX = TW * TH;
	constexpr int SEGS = TPX / 32;
	constexpr int NWARP = NT / 32;
	static_assert(TW % 32 == 0 && TPX <= 32768 && TH <= 32, "tile shape");
	extern __shared__ __align__(128) uint8_t smem[];
	uint8_t *lvl = smem;                                     // TMA destination, converted to levels in place
	uint32_t *par = reinterpret_cast<uint32_t *>(smem + TPX);
	uint32_t *cnt = par + TPX;
	uint32_t *xmn = cnt + TPX;
	uint32_t *xmx = xmn + TPX;
	uint32_t *ymask = xmx + TPX;                             // one bit per tile row (TH <= 32)
	uint16_t *rootlist = reinterpret_cast<uint16_t *>(ymask + TPX);
	__shared__ __align__(8) uint64_t bar;
	__shared__ uint32_t s_nroots, s_base, s_cursor, s_nlinks, s_nemit, s_minlvl, s_maxlvl;
	__shared__ uint16_t s_ringA[2 * (TW + TH)];              // per tile-side position: the node that stands for it across the seam (0xFFFF: none)

	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int plane = blockIdx.y;
	const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
	const int X0 = tx * TW, Y0 = ty * TH;
	const int rows = min(TH, P.H - Y0), cols = min(TW, P.W - X0);
	const PlaneSrc ps = planes[plane];
	const uint8_t *src = ps.src + (size_t)Y0 * P.pitch + X0;
	const size_t N = (size_t)P.W * P.H;
	uint32_t *parP = par_g + (size_t)plane * N;
	NodeAttr *attrP = attr_g + (size_t)plane * N;

	// ---- stage the tile through TMA (one bulk copy per row, all completing on one mbarrier) ----
	if (tid == 0) { mbar_init(&bar, 1); s_nroots = 0; s_cursor = 0; s_nlinks = 0; s_nemit = 0; s_minlvl = 255; s_maxlvl = 0; }
	for (int i = tid; i < 2 * (TW + TH); i += NT) s_ringA[i] = 0xFFFFu;
	__syncthreads();
	if (warp == 0) {
		if (lane == 0) mbar_expect_tx(&bar, (uint32_t)(rows * TW));
		__syncwarp();
		for (int r = lane; r < rows; r += 32) tma_row_g2s(lvl + r * TW, src + (size_t)r * P.pitch, TW, &bar);
	}
	if (warp == 0) mbar_wait(&bar, 0);   // one warp polls; the others park at the CTA barrier (no issue slots burnt on polling)
	__syncthreads();

	ERT_PHASE(0);
	// ---- phase A: quantise, horizontal same-level runs become chains without atomics ----
	uint32_t wmin = 255u, wmax = 0u;
	for (int seg = warp; seg < SEGS; seg += NWARP) {
		const int p = seg * 32 + lane;
		const int y = p / TW, x = p % TW;
		int L = 255;
		if (y < rows && x < cols) {
			int v = lvl[p];
			if (ps.invert) v = 255 - v;
			L = quantize_level(v, P.qscale);
			if (L >= P.hi) L = 255;
		}
		const int Lr = __shfl_down_sync(0xFFFFFFFFu, L, 1);
		const bool same = local_union && (lane < 31) && (Lr == L) && (L != 255);
		const uint32_t bmask = __ballot_sync(0xFFFFFFFFu, !same);
		const int end = lane + __ffs(bmask >> lane) - 1;
		lvl[p] = (uint8_t)L;   // every lane rewrites only the byte it read itself
		par[p] = (L == 255 || end == lane) ? KEY_NONE : (((uint32_t)L << 16) | (uint32_t)(seg * 32 + end));
		wmin = min(wmin, (uint32_t)L);
		wmax = max(wmax, (L == 255) ? 0u : (uint32_t)L);
	}
	{
		const uint32_t lmin = __reduce_min_sync(0xFFFFFFFFu, wmin), lmax = __reduce_max_sync(0xFFFFFFFFu, wmax);
		if (lane == 0) { if (lmin < 255u) atomicMin(&s_minlvl, lmin); atomicMax(&s_maxlvl, lmax); }
	}
	__syncthreads();

	ERT_PHASE(1);
	// ---- phase B: the remaining in-tile edges.  B1 compacts them into a work list in shared memory
	// (aliasing cnt/xmn, which are not live yet); B2 drains the list with warp-converged state machines:
	// every lane owns one edge at a time, all lanes advance one hop per iteration (no divergent inner
	// loops), idle lanes refill from the list -- so a long chain stalls one lane, not the CTA. ----
	uint32_t *links = cnt;   // up to 2*TPX packed (p << 16 | q)
	if (local_union) {
		for (int seg = warp; seg < SEGS; seg += NWARP) {
			const int p = seg * 32 + lane;
			const int y = p / TW, x = p % TW;
			const uint32_t L = lvl[p];
			const uint32_t Lb = (y + 1 < TH) ? (uint32_t)lvl[p + TW] : 255u;        // pixel below
			// horizontal neighbours come from the neighbouring lanes; only the segment's end lanes touch shared memory
			uint32_t Lrt = __shfl_down_sync(0xFFFFFFFFu, L, 1), Llf = __shfl_up_sync(0xFFFFFFFFu, L, 1), Lbl = __shfl_up_sync(0xFFFFFFFFu, Lb, 1);
			if (lane == 31) Lrt = (x + 1 < TW) ? (uint32_t)lvl[p + 1] : 255u;
			if (lane == 0) { Llf = (x > 0) ? (uint32_t)lvl[p - 1] : 255u; Lbl = (x > 0 && y + 1 < TH) ? (uint32_t)lvl[p + TW - 1] : 255u; }
			bool eh = false, ev = false;
			if (L != 255) {
				eh = (Lrt != 255) && (Lrt != L || lane == 31);
				if (Lb != 255) ev = !((x > 0) && (Llf == L) && (Lbl == Lb));
			}
			const uint32_t mh = __ballot_sync(0xFFFFFFFFu, eh), mv = __ballot_sync(0xFFFFFFFFu, ev);
			uint32_t base = 0;
			if (lane == 0 && (mh | mv)) base = atomicAdd(&s_nlinks, (uint32_t)(__popc(mh) + __popc(mv)));
			base = __shfl_sync(0xFFFFFFFFu, base, 0);
			const uint32_t lt = (1u << lane) - 1u;
			// endpoints are stored as the LAST pixel of each pixel's same-level run (what phase A made it point at):
			// that pixel is the run's level root until the run is merged, so most edges start with both ends at a root
			if (eh | ev) {
				const uint32_t pp = par[p];
				const uint32_t ep = (pp != KEY_NONE && (pp >> 16) == L) ? (pp & 0xFFFFu) : (uint32_t)p;
				if (eh) {
					const uint32_t q = (uint32_t)(p + 1), pq = par[q];
					const uint32_t eq = (pq != KEY_NONE && (pq >> 16) == (uint32_t)lvl[q]) ? (pq & 0xFFFFu) : q;
					links[base + __popc(mh & lt)] = (ep << 16) | eq;
				}
				if (ev) {
					const uint32_t q = (uint32_t)(p + TW), pq = par[q];
					const uint32_t eq = (pq != KEY_NONE && (pq >> 16) == (uint32_t)lvl[q]) ? (pq & 0xFFFFu) : q;
					links[base + __popc(mh) + __popc(mv & lt)] = (ep << 16) | eq;
				}
			}
		}
		__syncthreads();
		ERT_PHASE(2);
		{
			const uint32_t nl = s_nlinks;
			uint32_t a = 0, b = 0;
			bool act = false, more = true;
			// CHUNK: every warp owns a contiguous share of the (roughly raster-ordered) list: no shared cursor
			const uint32_t per_warp = (nl + NWARP - 1) / NWARP;
			uint32_t wnext = (uint32_t)warp * per_warp;
			const uint32_t wend = min(nl, wnext + per_warp);
			for (int guard = 0; guard < (1 << 22); ++guard) {
				const uint32_t idle = __ballot_sync(0xFFFFFFFFu, !act);
				if (more && (idle == 0xFFFFFFFFu || __popc(idle) >= RF)) {
					uint32_t base = 0, lim = nl;
					if (CHUNK) { base = wnext; wnext += (uint32_t)__popc(idle); lim = wend; }
					else {
						if (lane == 0) base = atomicAdd(&s_cursor, (uint32_t)__popc(idle));
						base = __shfl_sync(0xFFFFFFFFu, base, 0);
					}
					if (!act) {
						const uint32_t i = base + (uint32_t)__popc(idle & ((1u << lane) - 1u));
						if (i < lim) {
							// the list is drained from its END (bottom-right of the tile first): the level root of a
							// node is its last pixel in raster order, so it is met first and stays put while the rows
							// above attach to it directly -- chains stay ~1 hop deep instead of one hop per row
							const uint32_t e = links[nl - 1u - i];
							const uint32_t p = e >> 16, q = e & 0xFFFFu;
							a = ((uint32_t)lvl[p] << 16) | p;
							b = ((uint32_t)lvl[q] << 16) | q;
							act = true;
						}
					}
					if (base + (uint32_t)__popc(idle) >= lim) more = false;
				}
				if (!__any_sync(0xFFFFFFFFu, act)) { if (!more) break; else continue; }
				if (act) {
					const bool ra = climb_s(par, a);
					const bool rb = climb_s(par, b);
					if (ra && rb) {
						if (a == b) act = false;
						else {
							if (a > b) { const uint32_t t = a; a = b; b = t; }
							const uint32_t old = atomicMin(&par[a & 0xFFFFu], b);
							if (old == b || old == KEY_NONE) act = false;
							else if (old < b) a = old;
							else { a = b; b = old; }
						}
					}
				}
			}
		}
		__syncthreads();

		ERT_PHASE(3);
	}
	ERT_PHASE(4);
	for (int p = tid; p < TPX; p += NT) { cnt[p] = 0; xmn[p] = 0xFFFFFFFFu; xmx[p] = 0; ymask[p] = 0; }
	__syncthreads();

	// ---- phase D: own-level pixel count and bbox per tile-local node, one update per same-level run; the run's
	// last pixel looks its level root up with a warp-converged walk (no separate flatten pass over all pixels).
	// acc word: bits 0..14 pixels, bits 15..29 nodes, bit 31 = node touches a seam (BORDER) ----
	constexpr uint32_t ACC_NODE = 1u << 15, ACC_MASK = 0x7FFFu, ACC_BORDER = 0x80000000u;
	{
		volatile uint32_t *vpar = par;
		static_assert(SEGS / NWARP <= 8, "a warp keeps the root masks of its segments in registers");
		uint32_t rootmask[8] = {0, 0, 0, 0, 0, 0, 0, 0}, nroot_w = 0;
		int jseg = 0;
#pragma unroll
		for (int seg = warp; seg < SEGS; seg += NWARP) {
			const int p = seg * 32 + lane;
			const int y = p / TW, x = p % TW;
			const uint32_t L = lvl[p];
			const uint32_t Lr = __shfl_down_sync(0xFFFFFFFFu, L, 1);
			const bool same = local_union && (lane < 31) && (Lr == L) && (L != 255);
			const uint32_t bmask = __ballot_sync(0xFFFFFFFFu, !same);
			const bool runend = (L != 255) && !same;
			uint32_t k = (L << 16) | (uint32_t)p;
			bool act = runend;
			bool isroot = false;
			while (__any_sync(0xFFFFFFFFu, act)) {
				if (act) {
					const uint32_t q = vpar[k & 0xFFFFu];
					if (q == KEY_NONE || (q >> 16) != L) { act = false; isroot = ((k & 0xFFFFu) == (uint32_t)p); }
					else {
						const uint32_t g = vpar[q & 0xFFFFu];           // two hops per round; q is the root if its pointer leaves the level
						if (g == KEY_NONE || (g >> 16) != L) { k = q; act = false; }
						else k = g;
					}
				}
			}
			if (runend) {
				const uint32_t prev = bmask & ((1u << lane) - 1u);
				const int start = prev ? (32 - __clz(prev)) : 0;
				const int len = lane - start + 1;
				const uint32_t r = k & 0xFFFFu;
				atomicAdd(&cnt[r], (uint32_t)len + (isroot ? ACC_NODE : 0u));
				atomicMin(&xmn[r], (uint32_t)(x - (lane - start)));
				atomicMax(&xmx[r], (uint32_t)x);
				atomicOr(&ymask[r], 1u << y);
			}
			const uint32_t rmask = __ballot_sync(0xFFFFFFFFu, isroot);
			if (jseg < 8) { rootmask[jseg] = rmask; }
			nroot_w += (uint32_t)__popc(rmask);
			++jseg;
		}
		// one reservation in the root list per warp (order inside the list is irrelevant)
		uint32_t wb = 0;
		if (lane == 0 && nroot_w) wb = atomicAdd(&s_nroots, nroot_w);
		wb = __shfl_sync(0xFFFFFFFFu, wb, 0);
		jseg = 0;
		for (int seg = warp; seg < SEGS; seg += NWARP, ++jseg) {
			const uint32_t rmask = rootmask[jseg];
			if ((rmask >> lane) & 1u) rootlist[wb + __popc(rmask & ((1u << lane) - 1u))] = (uint16_t)(seg * 32 + lane);
			wb += (uint32_t)__popc(rmask);
		}
	}
	__syncthreads();
	ERT_PHASE(9);

	// ---- phase C: every level root points at its parent's LEVEL ROOT (roots only: a few % of the pixels) ----
	if (local_union) {
		volatile uint32_t *vpar = par;
		const uint32_t nr = s_nroots;
		for (uint32_t i0 = warp * 32; i0 < nr; i0 += NT) {
			const uint32_t i = i0 + lane;
			uint32_t p = 0, k = KEY_NONE;
			if (i < nr) { p = rootlist[i]; k = vpar[p]; }
			const uint32_t k0 = k;
			bool act = (k != KEY_NONE);
			while (__any_sync(0xFFFFFFFFu, act)) {
				if (act) {
					const uint32_t q = vpar[k & 0xFFFFu];
					if (q == KEY_NONE || (q >> 16) != (k >> 16)) act = false;
					else k = q;
				}
			}
			if (k != k0) par[p] = k;   // readers that still see k0 walk the same chain to the same root
		}
	}
	__syncthreads();

	ERT_PHASE(5);
	// ---- phase D2: which tile-local nodes can still change?  A pixel p on a side of the tile that faces another tile
	// meets its outside neighbour q at level M = max(level p, level q): what that edge can change is the tile-local
	// component holding p at threshold M -- the HIGHEST ancestor-or-self A(p) of p's node with level <= M -- and
	// everything above it.  Nodes below A(p) on p's root path are final unless another side pixel says otherwise; a wall
	// outside makes no edge at all.  So BORDER = the A(p) of all side pixels (plus the nodes of the flood's start
	// candidates, pixels 0 / 1 / W of the plane) and all their ancestors; everything else is INTERIOR: its subtree is
	// final here and never has to leave the SM.  A(p) also stands for p in the seam record (phase E): it is connected to
	// p inside the tile at a level <= M, so linking it to the other side makes the same union.  (CPU model on the bench
	// frames: 21.7 -> 13.6 global nodes per tile.) ----
	if (local_union) {
		const int ring = 2 * (TW + TH);
		for (int i = tid; i < ring + 3; i += NT) {
			int x, y;
			bool on = true;
			if (i < TW) { x = i; y = 0; on = (Y0 > 0); }
			else if (i < 2 * TW) { x = i - TW; y = rows - 1; on = (Y0 + rows < P.H); }
			else if (i < 2 * TW + TH) { x = 0; y = i - 2 * TW; on = (X0 > 0); }
			else if (i < ring) { x = cols - 1; y = i - 2 * TW - TH; on = (X0 + cols < P.W); }
			else {   // start candidates of the flood: global pixels 0, 1, W
				const int gi = i - ring;
				const int gx = (gi == 1) ? 1 : 0, gy = (gi == 2) ? 1 : 0;
				x = gx - X0; y = gy - Y0;
			}
			if (!on || x < 0 || y < 0 || x >= cols || y >= rows) continue;
			const int p = y * TW + x;
			const uint32_t L = lvl[p];
			if (L == 255) continue;
			uint32_t M = L;                                 // start candidates and the record-less debug mode: the pixel's own node
			if (ring_rec && i < ring) {
				int gx = X0 + x, gy = Y0 + y;
				if (i < TW) gy -= 1; else if (i < 2 * TW) gy += 1; else if (i < 2 * TW + TH) gx -= 1; else gx += 1;
				int v = __ldg(ps.src + (size_t)gy * P.pitch + gx);
				if (ps.invert) v = 255 - v;
				const int Lq = quantize_level(v, P.qscale);
				if (Lq >= P.hi) continue;                    // a wall outside: no edge across the seam here
				M = max(L, (uint32_t)Lq);
			}
			uint32_t kk = (L << 16) | (uint32_t)p;
			for (int guard = 0; guard < 65536; ++guard) {   // to the level root
				const uint32_t q = par[kk & 0xFFFFu];
				if (q == KEY_NONE || (q >> 16) != L) break;
				kk = q;
			}
			uint32_t r = kk & 0xFFFFu;
			for (int guard = 0; guard < 64; ++guard) {      // level roots point at their parent's level root (phase C): climb while level <= M
				const uint32_t up = par[r];
				if (up == KEY_NONE || (up >> 16) > M) break;
				r = up & 0xFFFFu;
			}
			if (i < ring) s_ringA[i] = (uint16_t)r;
			for (int guard = 0; guard < 64; ++guard) {
				const uint32_t old = atomicOr(&cnt[r], ACC_BORDER);
				if (old & ACC_BORDER) break;
				const uint32_t up = par[r];
				if (up == KEY_NONE) break;
				r = up & 0xFFFFu;
			}
		}
	} else {
		for (int p = tid; p < TPX; p += NT) if (lvl[p] != 255) cnt[p] |= ACC_BORDER;
	}
	__syncthreads();

	ERT_PHASE(6);
	// ---- phase D3: fold interior subtrees bottom-up, one level per round (a tile holds few distinct levels) ----
	const uint32_t nroots = s_nroots;
	if (local_union) {
		const int lo = (int)s_minlvl, hi_l = (int)s_maxlvl;
		for (int Lc = lo; Lc < hi_l; ++Lc) {
			for (uint32_t i = tid; i < nroots; i += NT) {
				const uint32_t p = rootlist[i];
				if (lvl[p] != Lc) continue;
				const uint32_t acc = cnt[p];
				if (acc & ACC_BORDER) continue;
				const uint32_t up = par[p];
				if (up == KEY_NONE) continue;
				const uint32_t q = up & 0xFFFFu;
				atomicAdd(&cnt[q], acc);           // pixels and node count travel together
				atomicMin(&xmn[q], xmn[p]); atomicMax(&xmx[q], xmx[p]);
				atomicOr(&ymask[q], ymask[p]);
			}
			__syncthreads();
		}
	}

	ERT_PHASE(7);
	// ---- phase E: emit.  BORDER nodes go to the global forest with what they have gathered (own pixels +
	// interior descendants); interior nodes are emitted only if the reference would keep them
	// (area > MIN_AREA), already complete (pend = NODE_COMPLETE); seam pixels publish their root. ----
	uint32_t my_emit = 0;
	for (uint32_t i = tid; i < nroots; i += NT) {
		const uint32_t acc = cnt[rootlist[i]];
		const bool emit = (acc & ACC_BORDER) || (int)((acc & ACC_MASK) + ((acc >> 15) & ACC_MASK)) > P.min_area;
		my_emit += emit ? 1u : 0u;
	}
	my_emit = __reduce_add_sync(0xFFFFFFFFu, my_emit);
	if (lane == 0 && my_emit) atomicAdd(&s_nemit, my_emit);
	__syncthreads();
	if (tid == 0) { s_base = s_nemit ? atomicAdd(&node_count[plane], s_nemit) : 0u; s_cursor = 0; }
	__syncthreads();
	for (uint32_t i0 = warp * 32; i0 < nroots; i0 += NT) {
		const uint32_t i = i0 + lane;
		bool emit = false;
		uint32_t p = 0, acc = 0;
		if (i < nroots) {
			p = rootlist[i];
			acc = cnt[p];
			emit = (acc & ACC_BORDER) || (int)((acc & ACC_MASK) + ((acc >> 15) & ACC_MASK)) > P.min_area;
		}
		const uint32_t emask = __ballot_sync(0xFFFFFFFFu, emit);
		uint32_t wbase = 0;
		if (lane == 0 && emask) wbase = atomicAdd(&s_cursor, (uint32_t)__popc(emask));
		wbase = __shfl_sync(0xFFFFFFFFu, wbase, 0);
		if (!emit) continue;
		const int y = (int)p / TW, x = (int)p % TW;
		const uint32_t L = lvl[p];
		const uint32_t gidx = (uint32_t)(Y0 + y) * (uint32_t)P.W + (uint32_t)(X0 + x);
		const uint32_t pk = par[p];
		uint32_t gpar = KEY_NONE;
		if (pk != KEY_NONE) {
			const uint32_t q = pk & 0xFFFFu;
			gpar = make_key(pk >> 16, (uint32_t)(Y0 + (int)(q / TW)) * (uint32_t)P.W + (uint32_t)(X0 + (int)(q % TW)));
		}
		parP[gidx] = gpar;
		uint4 *a = reinterpret_cast<uint4 *>(&attrP[gidx]);
		a[0] = make_uint4(acc & ACC_MASK, (acc >> 15) & ACC_MASK, (acc & ACC_BORDER) ? 0u : NODE_COMPLETE, 0u);
		const uint32_t ym = ymask[p];
		a[1] = make_uint4((uint32_t)X0 + xmn[p], (uint32_t)Y0 + (uint32_t)(__ffs(ym) - 1), (uint32_t)X0 + xmx[p], (uint32_t)Y0 + (uint32_t)(31 - __clz(ym)));
		const uint32_t pos = s_base + wbase + (uint32_t)__popc(emask & ((1u << lane) - 1u));
		node_list[(size_t)plane * N + pos] = make_key(L, gidx);
	}
	{
		const int ring = 2 * (TW + TH);
		// seam records: for every position on the four tile sides the GLOBAL key of the level root of the node that
		// stands for the pixel there (KEY_NONE where no edge crosses), laid out contiguously per tile so that
		// k_seam_link_rec reads both sides of a seam coalesced and starts every union at a root
		uint32_t *rec = ring_rec ? ring_rec + ((size_t)plane * gridDim.x + blockIdx.x) * ring : nullptr;
		for (int i = tid; i < ring + 3; i += NT) {
			int x, y;
			if (i < TW) { x = i; y = 0; }
			else if (i < 2 * TW) { x = i - TW; y = rows - 1; }
			else if (i < 2 * TW + TH) { x = 0; y = i - 2 * TW; }
			else if (i < ring) { x = cols - 1; y = i - 2 * TW - TH; }
			else { const int gi = i - ring; x = ((gi == 1) ? 1 : 0) - X0; y = ((gi == 2) ? 1 : 0) - Y0; }
			if (rec && i < ring) {
				// the node that stands for this side position (phase D2); KEY_NONE for walls, pixels outside the plane,
				// sides on the plane's border and positions whose outside neighbour is a wall
				const uint32_t a = s_ringA[i];
				rec[i] = (a == 0xFFFFu) ? KEY_NONE
				                        : make_key((uint32_t)lvl[a], (uint32_t)(Y0 + (int)(a / TW)) * (uint32_t)P.W + (uint32_t)(X0 + (int)(a % TW)));
				continue;
			}
			uint32_t rootkey = KEY_NONE;
			bool is_root = false;
			if (x >= 0 && y >= 0 && x < cols && y < rows) {
				const int p = y * TW + x;
				const uint32_t L = lvl[p];
				if (L != 255) {
					uint32_t kk = (L << 16) | (uint32_t)p;
					for (int guard = 0; guard < 65536; ++guard) {
						const uint32_t q2 = par[kk & 0xFFFFu];
						if (q2 == KEY_NONE || (q2 >> 16) != L) break;
						kk = q2;
					}
					const uint32_t q = kk & 0xFFFFu;
					is_root = (q == (uint32_t)p);
					rootkey = make_key(L, (uint32_t)(Y0 + (int)(q / TW)) * (uint32_t)P.W + (uint32_t)(X0 + (int)(q % TW)));
				}
			}
			// non-root pixels publish their root in par[] when something will look them up by pixel:
			// the flood's start candidates always, seam pixels only in the record-less (debug) mode
			if (rootkey != KEY_NONE && !is_root)
				parP[(uint32_t)(Y0 + y) * (uint32_t)P.W + (uint32_t)(X0 + x)] = rootkey;
		}
	}
	ERT_PHASE(8);
#undef ERT_PHASE
}


// ---------------------------------------------------------------------------------------------
// global-memory keyed union-find (seams)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t find_g(const uint32_t *par, uint32_t k)
{
	for (;;) {
		const uint32_t p = ld_relaxed(par + key_idx(k));
		if (p == KEY_NONE || key_level(p) != key_level(k)) return k;
		k = p;
	}
}

// find with path halving for the seam kernel: chains of tile roots of one big same-level region (background) grow with
// the number of tiles it spans.  par[k] = same-level grandparent is a relaxed store; the argument of climb_s carries
// over: whatever a concurrent atomicMin puts below a same-level parent is itself a same-level pixel of the node and is
// re-linked above by the thread that displaced it, so per-level connectivity never changes.
__device__ __forceinline__ uint32_t find_halve_g(uint32_t *par, uint32_t k)
{
	for (;;) {
		const uint32_t p = ld_relaxed(par + key_idx(k));
		if (p == KEY_NONE || key_level(p) != key_level(k)) return k;
		const uint32_t g = ld_relaxed(par + key_idx(p));
		if (g == KEY_NONE || key_level(g) != key_level(k)) return p;
		asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(par + key_idx(k)), "r"(g) : "memory");
		k = g;
	}
}

__device__ __forceinline__ void link_g(uint32_t *par, uint32_t a, uint32_t b, uint32_t *status)
{
	for (int guard = 0; guard < (1 << 22); ++guard) {
		a = find_halve_g(par, a);
		b = find_halve_g(par, b);
		if (a == b) return;
		if (a > b) { const uint32_t t = a; a = b; b = t; }
		const uint32_t old = atomicMin(&par[key_idx(a)], b);
		if (old == b || old == KEY_NONE) return;
		if (old < b) a = old;
		else { a = b; b = old; }
	}
	atomicOr(status, ERR_LOOP_GUARD);
}

__device__ __forceinline__ int level_at(const PlaneSrc &ps, const ExtractParams &P, int x, int y)
{
	int v = __ldg(ps.src + (size_t)y * P.pitch + x);
	if (ps.invert) v = 255 - v;
	const int L = quantize_level(v, P.qscale);
	return L >= P.hi ? 255 : L;
}

__global__ void k_seam_link(ExtractParams P, const PlaneSrc *__restrict__ planes, uint32_t *__restrict__ par_g,
                            uint32_t *status, int TW, int TH)
{
	const int plane = blockIdx.y;
	const int nvs = (P.W - 1) / TW, nhs = (P.H - 1) / TH;
	const long long nv = (long long)nvs * P.H, nh = (long long)nhs * P.W;
	const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (e >= nv + nh) return;
	const PlaneSrc ps = planes[plane];
	uint32_t *parP = par_g + (size_t)plane * P.W * P.H;
	int xa, ya, xb, yb;
	bool skip = false;
	int la, lb;
	if (e < nv) {
		const int k = (int)(e / P.H) + 1, y = (int)(e % P.H);
		xa = k * TW - 1; xb = xa + 1; ya = yb = y;
		la = level_at(ps, P, xa, ya); lb = level_at(ps, P, xb, yb);
		if (la == 255 || lb == 255) return;
		// the pair one row up makes the same union -- but only if each of its pixels is already united with
		// the pixel below it INSIDE a tile (not across a horizontal seam, whose own skip rule would lean on us)
		if (y % TH != 0) skip = (level_at(ps, P, xa, y - 1) == la) && (level_at(ps, P, xb, y - 1) == lb);
	} else {
		const long long e2 = e - nv;
		const int k = (int)(e2 / P.W) + 1, x = (int)(e2 % P.W);
		ya = k * TH - 1; yb = ya + 1; xa = xb = x;
		la = level_at(ps, P, xa, ya); lb = level_at(ps, P, xb, yb);
		if (la == 255 || lb == 255) return;
		if (x % TW != 0) skip = (level_at(ps, P, x - 1, ya) == la) && (level_at(ps, P, x - 1, yb) == lb);
	}
	if (skip) return;
	link_g(parP, make_key((uint32_t)la, (uint32_t)(ya * P.W + xa)), make_key((uint32_t)lb, (uint32_t)(yb * P.W + xb)), status);
}

// seams from the tile kernel's records: edge = (record of the pixel on one side, record on the other side); a pair
// identical to the pair one step earlier along the seam (inside the same tile pair) makes the same union and is skipped
__global__ void k_seam_link_rec(ExtractParams P, const uint32_t *__restrict__ ring_rec, uint32_t *__restrict__ par_g, uint32_t *status,
                                int TW, int TH, int tiles_x, int tiles_per_plane)
{
	const int plane = blockIdx.y;
	const int RINGW = 2 * (TW + TH);
	const int nvs = (P.W - 1) / TW, nhs = (P.H - 1) / TH;
	const long long nv = (long long)nvs * P.H, nh = (long long)nhs * P.W;
	const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (e >= nv + nh) return;
	const uint32_t *recP = ring_rec + (size_t)plane * tiles_per_plane * RINGW;
	uint32_t *parP = par_g + (size_t)plane * P.W * P.H;
	size_t ia, ib;
	bool has_prev;
	if (e < nv) {
		// consecutive threads walk down one seam line: records of a tile side are contiguous
		const int k = (int)(e / P.H) + 1, y = (int)(e % P.H);
		const int ty = y / TH, yy = y % TH;
		ia = ((size_t)ty * tiles_x + (k - 1)) * RINGW + 2 * TW + TH + yy;   // right side of the left tile
		ib = ((size_t)ty * tiles_x + k) * RINGW + 2 * TW + yy;             // left side of the right tile
		has_prev = yy != 0;
	} else {
		const long long e2 = e - nv;
		const int k = (int)(e2 / P.W) + 1, x = (int)(e2 % P.W);
		const int tx = x / TW, xx = x % TW;
		ia = ((size_t)(k - 1) * tiles_x + tx) * RINGW + TW + xx;            // bottom side of the upper tile
		ib = ((size_t)k * tiles_x + tx) * RINGW + xx;                       // top side of the lower tile
		has_prev = xx != 0;
	}
	const uint32_t ra = recP[ia], rb = recP[ib];
	if (ra == KEY_NONE || rb == KEY_NONE) return;
	if (has_prev && recP[ia - 1] == ra && recP[ib - 1] == rb) return;
	link_g(parP, ra, rb, status);
}

// ---------------------------------------------------------------------------------------------
// k_fold : aliases (nodes merged into a same-level node of another tile) hand their own-level
// pixels to the final node; final nodes resolve their final parent and register as its child.
// ---------------------------------------------------------------------------------------------
__global__ void k_fold(ExtractParams P, uint32_t *__restrict__ par_g, NodeAttr *__restrict__ attr_g,
                       const uint32_t *__restrict__ node_list, const uint32_t *__restrict__ node_count)
{
	const int plane = blockIdx.y;
	const size_t N = (size_t)P.W * P.H;
	uint32_t *parP = par_g + (size_t)plane * N;
	NodeAttr *attrP = attr_g + (size_t)plane * N;
	const uint32_t n = node_count[plane];
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		const uint32_t g = node_list[(size_t)plane * N + i];
		NodeAttr *ag = &attrP[key_idx(g)];
		if (ag->pend == NODE_COMPLETE) {
			// interior node, finished inside its tile; only its parent may have been merged across a seam
			const uint32_t pk = parP[key_idx(g)];
			if (pk != KEY_NONE) {
				const uint32_t fp = find_g(parP, pk);
				if (fp != pk) parP[key_idx(g)] = fp;
			}
			continue;
		}
		const uint32_t f = find_g(parP, g);
		if (f != g) {
			NodeAttr *af = &attrP[key_idx(f)];
			atomicAdd(&af->cnt, ag->cnt);
			if (ag->nn > 1) atomicAdd(&af->nn, ag->nn - 1u);   // its interior descendants are nodes of the final node's subtree
			atomicMin(&af->x0, ag->x0); atomicMin(&af->y0, ag->y0);
			atomicMax(&af->x1, ag->x1); atomicMax(&af->y1, ag->y1);
			ag->nn = 0;   // alias marker
		} else {
			const uint32_t pk = ld_relaxed(parP + key_idx(g));
			if (pk != KEY_NONE) {
				const uint32_t fp = find_g(parP, pk);
				if (fp != pk) parP[key_idx(g)] = fp;
				atomicAdd(&attrP[key_idx(fp)].pend, 1u);
			}
		}
	}
}

// ---------------------------------------------------------------------------------------------
// k_refit : leaves start; a thread carries a node's finished totals into its parent and continues
// upward only if it was the last child to arrive (no grid-wide synchronisation).
// ---------------------------------------------------------------------------------------------
__global__ void k_refit(ExtractParams P, const uint32_t *__restrict__ par_g, NodeAttr *__restrict__ attr_g,
                        const uint32_t *__restrict__ node_list, const uint32_t *__restrict__ node_count)
{
	const int plane = blockIdx.y;
	const size_t N = (size_t)P.W * P.H;
	const uint32_t *parP = par_g + (size_t)plane * N;
	NodeAttr *attrP = attr_g + (size_t)plane * N;
	const uint32_t n = node_count[plane];
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		uint32_t cur = node_list[(size_t)plane * N + i];
		{
			const NodeAttr *a = &attrP[key_idx(cur)];
			if (a->nn == 0 || a->pend != 0) continue;   // alias, or not a leaf (pend is constant in this kernel)
		}
		for (int guard = 0; guard < 64; ++guard) {
			const uint32_t p = parP[key_idx(cur)];
			if (p == KEY_NONE) break;
			const uint32_t *a = reinterpret_cast<const uint32_t *>(&attrP[key_idx(cur)]);
			const uint32_t c_cnt = ld_relaxed(a + 0), c_nn = ld_relaxed(a + 1);
			const uint32_t c_x0 = ld_relaxed(a + 4), c_y0 = ld_relaxed(a + 5), c_x1 = ld_relaxed(a + 6), c_y1 = ld_relaxed(a + 7);
			NodeAttr *ap = &attrP[key_idx(p)];
			atomicAdd(&ap->cnt, c_cnt);
			atomicAdd(&ap->nn, c_nn);
			atomicMin(&ap->x0, c_x0); atomicMin(&ap->y0, c_y0);
			atomicMax(&ap->x1, c_x1); atomicMax(&ap->y1, c_y1);
			__threadfence();
			const uint32_t t = atomicAdd(&ap->arr, 1u);
			if (t + 1u != ap->pend) break;
			__threadfence();
			cur = p;
		}
	}
}

// ---------------------------------------------------------------------------------------------
// k_reach_root : the flood starts at pixel 0; if that is a wall it escapes to pixel 1, else to
// pixel W (neighbour order right, bottom); only that tree is the reference's result.
// ---------------------------------------------------------------------------------------------
__device__ uint32_t reach_root_of(const ExtractParams &P, const PlaneSrc &ps, const uint32_t *parP, int32_t *lone)
{
	int s = -1, ls = 255;
	const int l0 = level_at(ps, P, 0, 0);
	if (l0 != 255) { s = 0; ls = l0; }
	else if (P.W > 1 && (ls = level_at(ps, P, 1, 0)) != 255) s = 1;
	else if (P.H > 1 && (ls = level_at(ps, P, 0, 1)) != 255) s = P.W;
	if (s < 0) {
		int v = __ldg(ps.src);
		if (ps.invert) v = 255 - v;
		*lone = quantize_level(v, P.qscale);
		return KEY_NONE;
	}
	uint32_t k = find_g(parP, make_key((uint32_t)ls, (uint32_t)s));
	for (int guard = 0; guard < 64; ++guard) {
		const uint32_t p = parP[key_idx(k)];
		if (p == KEY_NONE) break;
		k = p;
	}
	*lone = -1;
	return k;
}

// ---------------------------------------------------------------------------------------------
// k_emit_kept : what er_merge leaves alive (src/ER.cpp:167-180): area > MIN_AREA, plus the root
// ---------------------------------------------------------------------------------------------
__global__ void k_emit_kept(ExtractParams P, const uint32_t *__restrict__ par_g, NodeAttr *__restrict__ attr_g,
                            const uint32_t *__restrict__ node_list, const uint32_t *__restrict__ node_count,
                            const PlaneSrc *__restrict__ planes, uint32_t *__restrict__ reach_root, int32_t *__restrict__ lone_level,
                            KeptRec *__restrict__ kept, uint32_t *__restrict__ kept_count, uint32_t *status)
{
	const int plane = blockIdx.y;
	const size_t N = (size_t)P.W * P.H;
	const uint32_t *parP = par_g + (size_t)plane * N;
	NodeAttr *attrP = attr_g + (size_t)plane * N;
	const uint32_t n = node_count[plane];
	// the flood's start-pixel rule, evaluated once per block (a ~40-load dependent chain, cheaper than its own launch)
	__shared__ uint32_t s_rr;
	if (threadIdx.x == 0) {
		int32_t lone = -1;
		s_rr = reach_root_of(P, planes[plane], parP, &lone);
		if (blockIdx.x == 0) { reach_root[plane] = s_rr; lone_level[plane] = lone; }
	}
	__syncthreads();
	const uint32_t rr = s_rr;
	if (rr == KEY_NONE) return;
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		const uint32_t g = node_list[(size_t)plane * N + i];
		NodeAttr *a = &attrP[key_idx(g)];
		if (a->nn == 0) continue;
		const int area = (int)(a->cnt + a->nn);
		if (!(area > P.min_area || g == rr)) continue;
		uint32_t t = g;
		for (int guard = 0; guard < 64; ++guard) {
			const uint32_t p = parP[key_idx(t)];
			if (p == KEY_NONE) break;
			t = p;
		}
		if (t != rr) continue;
		const uint32_t pos = atomicAdd(&kept_count[plane], 1u);
		if (pos >= (uint32_t)P.kept_cap) { atomicOr(status, ERR_KEPT_OVERFLOW); continue; }
		a->arr = pos;
		KeptRec r;
		r.gidx = key_idx(g);
		const uint32_t pk = parP[key_idx(g)];
		r.parent = (pk == KEY_NONE) ? KEY_NONE : key_idx(pk);
		r.level = (int32_t)key_level(g);
		r.area = area;
		r.x0 = (uint16_t)a->x0; r.y0 = (uint16_t)a->y0; r.x1 = (uint16_t)a->x1; r.y1 = (uint16_t)a->y1;
		kept[(size_t)plane * P.kept_cap + pos] = r;
	}
}

// ---------------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------------
int extract_pitch(int W) { return (W + 127) / 128 * 128; }

// tile configurations (selectable at run time for tuning; id 0 is the default)
struct TileCfg { int tw, th, nt; };
// 0 = k_tile_build2 (er_tile.cu); 1 = the round-1 kernel (also the record-less debug mode local_union = 0); 2 = round-1 with a shared work queue
// 3..5 = k_tile_build2 variants for A/B (see er_tile.cu: OPT bits)
static const TileCfg g_tile_cfgs[] = {{64, 32, 256}, {64, 32, 512}, {64, 32, 512}, {64, 32, 256}, {64, 32, 256}, {64, 32, 256}};
int tile_config_count() { return (int)(sizeof(g_tile_cfgs) / sizeof(g_tile_cfgs[0])); }
size_t ring_words_per_plane(int W, int H)
{
	size_t m = 0;
	for (int i = 0; i < tile_config_count(); i++) {
		const TileCfg c = g_tile_cfgs[i];
		m = std::max(m, (size_t)((W + c.tw - 1) / c.tw) * ((H + c.th - 1) / c.th) * 2 * (c.tw + c.th));
	}
	return m;
}

template <int TW, int TH, int NT, int RF, bool CHUNK>
static int launch_tile(const ExtractParams &P, const PlaneSrc *d_planes, ExtractWork &wk, int local_union, cudaStream_t st)
{
	const size_t smem = (size_t)TW * TH * (1 + 5 * 4 + 2);
	ERT_CUDA_CHECK(cudaFuncSetAttribute(k_tile_build<TW, TH, NT, RF, CHUNK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	const int tiles_x = (P.W + TW - 1) / TW, tiles_y = (P.H + TH - 1) / TH;
	dim3 grid(tiles_x * tiles_y, P.n_planes);
	k_tile_build<TW, TH, NT, RF, CHUNK><<<grid, NT, smem, st>>>(P, d_planes, wk.par, wk.attr, wk.node_list, wk.node_count, wk.status, tiles_x, local_union, wk.prof,
	                                                 local_union ? wk.ring_rec : nullptr);
	ERT_CUDA_CHECK(cudaGetLastError());
	return 0;
}

__global__ void k_unpack_planes(const uint8_t *__restrict__ ycc, int pitch, int W, int H, uint8_t *__restrict__ out6)
{
	const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
	if (x >= W) return;
	const size_t plane_bytes = (size_t)pitch * H, n = (size_t)W * H;
#pragma unroll
	for (int k = 0; k < 3; k++) {
		const uint8_t v = ycc[k * plane_bytes + (size_t)y * pitch + x];
		out6[k * n + (size_t)y * W + x] = v;
		out6[(k + 3) * n + (size_t)y * W + x] = (uint8_t)(255 - v);
	}
}

int launch_unpack_planes(const uint8_t *d_ycc, int pitch, int W, int H, uint8_t *d_out6, cudaStream_t st)
{
	dim3 grid((W + 255) / 256, H);
	k_unpack_planes<<<grid, 256, 0, st>>>(d_ycc, pitch, W, H, d_out6);
	ERT_CUDA_CHECK(cudaGetLastError());
	return 0;
}

int launch_channels(const uint8_t *d_bgr, size_t frame_stride, int row_stride, int W, int H, int n_frames, uint8_t *d_ycc, int pitch, cudaStream_t st)
{
	dim3 block(128), grid(((W + 3) / 4 + 127) / 128, H, n_frames);
	k_channels<<<grid, block, 0, st>>>(d_bgr, frame_stride, row_stride, W, H, d_ycc, pitch);
	ERT_CUDA_CHECK(cudaGetLastError());
	return 0;
}

int launch_extract(const ExtractParams &P, const PlaneSrc *d_planes, ExtractWork &wk, int local_union, cudaStream_t st,
                   cudaEvent_t ev_tile_begin, cudaEvent_t ev_tile_end, cudaStream_t st_post)
{
	const TileCfg tc = g_tile_cfgs[(wk.tile_cfg >= 0 && wk.tile_cfg < tile_config_count()) ? wk.tile_cfg : 0];
	const int TILE_W = tc.tw, TILE_H = tc.th;
	ERT_CUDA_CHECK(cudaMemsetAsync(wk.node_count, 0, sizeof(uint32_t) * P.n_planes, st));
	ERT_CUDA_CHECK(cudaMemsetAsync(wk.kept_count, 0, sizeof(uint32_t) * P.n_planes, st));
	if (ev_tile_begin) ERT_CUDA_CHECK(cudaEventRecord(ev_tile_begin, st));
	int rc = -1;
	switch (local_union ? wk.tile_cfg : 1) {
	case 1: rc = launch_tile<64, 32, 512, 8, true>(P, d_planes, wk, local_union, st); break;
	case 2: rc = launch_tile<64, 32, 512, 12, false>(P, d_planes, wk, local_union, st); break;   // shared work queue instead of per-warp shares
	case 3: rc = launch_tile_v2(P, d_planes, wk, st, 0); break;
	case 4: rc = launch_tile_v2(P, d_planes, wk, st, 1); break;
	case 5: rc = launch_tile_v2(P, d_planes, wk, st, 2); break;
	default: rc = launch_tile_v2(P, d_planes, wk, st, 3); break;
	}
	if (rc) return rc;
	if (ev_tile_end) ERT_CUDA_CHECK(cudaEventRecord(ev_tile_end, st));
	if (st_post && st_post != st) {
		// everything after the SM-filling tile kernel runs on the context's high-priority stream: its narrow, latency-bound
		// kernels get CTA slots as soon as tile CTAs of OTHER batches retire, instead of queueing behind whole tile grids
		if (!ev_tile_end) { set_error("launch_extract: the split-stream mode needs the tile-end event"); return -1; }
		ERT_CUDA_CHECK(cudaStreamWaitEvent(st_post, ev_tile_end, 0));
		st = st_post;
	}
	{
		// with local_union == 0 every pixel is its own tile-local node and ALL edges are seams (debug A/B mode)
		const int tw = local_union ? TILE_W : 1, th = local_union ? TILE_H : 1;
		const long long edges = (long long)((P.W - 1) / tw) * P.H + (long long)((P.H - 1) / th) * P.W;
		if (edges > 0) {
			dim3 grid((unsigned)((edges + 255) / 256), P.n_planes);
			if (local_union && wk.ring_rec) {
				const int tiles_x = (P.W + tw - 1) / tw, tiles_y = (P.H + th - 1) / th;
				k_seam_link_rec<<<grid, 256, 0, st>>>(P, wk.ring_rec, wk.par, wk.status, tw, th, tiles_x, tiles_x * tiles_y);
			} else k_seam_link<<<grid, 256, 0, st>>>(P, d_planes, wk.par, wk.status, tw, th);
			ERT_CUDA_CHECK(cudaGetLastError());
		}
	}
	{
		dim3 grid(wk.node_blocks, P.n_planes);
		k_fold<<<grid, 256, 0, st>>>(P, wk.par, wk.attr, wk.node_list, wk.node_count);
		ERT_CUDA_CHECK(cudaGetLastError());
		k_refit<<<grid, 256, 0, st>>>(P, wk.par, wk.attr, wk.node_list, wk.node_count);
		ERT_CUDA_CHECK(cudaGetLastError());
		k_emit_kept<<<grid, 256, 0, st>>>(P, wk.par, wk.attr, wk.node_list, wk.node_count, d_planes, wk.reach_root, wk.lone_level, wk.kept,
		                                  wk.kept_count, wk.status);
		ERT_CUDA_CHECK(cudaGetLastError());
	}
	return 0;
}

} // namespace ert
