// er_extract.cu -- component-tree build (replaces ERFilter::er_tree_extract, src/ER.cpp:240-413,
// and compute_channels, src/ER.cpp:114-128) as hand-written sm_100a kernels.
//
// Pipeline per batch of planes (a plane = one u8 channel image, 6 per BGR frame):
//   k_channels        BGR -> Y, Cr, Cb planes (inverted planes are derived on the fly)
//   k_tile_build2     (er_tile.cu) per 64x32 tile: one tensor-map TMA box with halo, quantise to levels, tile-local
//                     component forest by a keyed lock-free union-find in shared memory, own-level pixel counts / bbox,
//                     interior subtrees folded on chip; nodes that touch a seam (or that the reference keeps) get a SLOT
//                     in the plane's dense node arrays; seam records name the slot that stands for every side pixel
//   k_seam_link_list  stitch tiles: the same keyed union-find over slots on the edges that cross tile seams
//   k_fold            fold nodes that were merged across seams into their final node; count children
//   k_refit           bottom-up accumulation of (pixels, nodes, bbox) with arrival counters, no grid sync
//   k_emit_kept       the flood's start-pixel rule (src/ER.cpp:267-341) picks the reference's tree; compact the nodes
//                     the reference keeps (area > MIN_AREA, or the root)
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
	uint32_t *parP = par_g + (size_t)plane * P.node_cap;
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

// The same seams, two phases per CTA: (1) every thread looks at SEAM_CHUNK / NT seam positions (coalesced record reads,
// identical neighbouring pairs dropped) and the positions that carry an edge -- ~15 % on natural frames -- are compacted
// into a shared list; (2) the list is drained by warp-converged state machines (one edge per lane, one hop or one
// atomicMin per iteration, idle lanes refill).  k_seam_link_rec keeps a whole warp resident for the L2-latency-bound
// chain of its few linking lanes; here every resident warp slot works on 32 chains, so the kernel holds ~6x fewer
// CTA slots while it overlaps the tile kernels of the other batches in flight.
constexpr int SEAM_NT = 256, SEAM_CHUNK = 2048;

__device__ __forceinline__ void seam_position(long long e, const ExtractParams &P, int TW, int TH, int tiles_x, int RINGW, long long nv, size_t &ia, size_t &ib,
                                              bool &has_prev)
{
	if (e < nv) {
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
}

__global__ void __launch_bounds__(SEAM_NT) k_seam_link_list(ExtractParams P, const uint32_t *__restrict__ ring_rec, uint32_t *__restrict__ par_g,
                                                           uint32_t *status, int TW, int TH, int tiles_x, int tiles_per_plane)
{
	__shared__ uint16_t s_pos[SEAM_CHUNK];
	__shared__ uint32_t s_n, s_cursor;
	const int plane = blockIdx.y, tid = threadIdx.x, lane = tid & 31;
	const int RINGW = 2 * (TW + TH);
	const int nvs = (P.W - 1) / TW, nhs = (P.H - 1) / TH;
	const long long nv = (long long)nvs * P.H, nh = (long long)nhs * P.W;
	const uint32_t *recP = ring_rec + (size_t)plane * tiles_per_plane * RINGW;
	uint32_t *parP = par_g + (size_t)plane * P.node_cap;
	bool guard_hit = false;
	// a CTA walks several chunks when the launcher caps the grid (the kernel runs under another batch's tile kernel and must
	// not take the SMs away from it: launch_extract)
	for (long long e0 = (long long)blockIdx.x * SEAM_CHUNK; e0 < nv + nh; e0 += (long long)gridDim.x * SEAM_CHUNK) {
	__syncthreads();
	if (tid == 0) { s_n = 0; s_cursor = 0; }
	__syncthreads();
	for (int k = tid; k < SEAM_CHUNK; k += SEAM_NT) {
		const long long e = e0 + k;
		bool edge = false;
		if (e < nv + nh) {
			size_t ia, ib;
			bool has_prev;
			seam_position(e, P, TW, TH, tiles_x, RINGW, nv, ia, ib, has_prev);
			const uint32_t ra = recP[ia], rb = recP[ib];
			edge = ra != KEY_NONE && rb != KEY_NONE && !(has_prev && recP[ia - 1] == ra && recP[ib - 1] == rb);
		}
		const uint32_t m = __ballot_sync(0xFFFFFFFFu, edge);
		uint32_t base = 0;
		if (lane == 0 && m) base = atomicAdd(&s_n, (uint32_t)__popc(m));
		base = __shfl_sync(0xFFFFFFFFu, base, 0);
		if (edge) s_pos[base + __popc(m & ((1u << lane) - 1u))] = (uint16_t)k;
	}
	__syncthreads();
	const uint32_t n = s_n;
	uint32_t a = 0, b = 0;      // a == b: the lane is idle
	bool more = true;
	for (int guard = 0; guard < (1 << 22); ++guard) {
		const uint32_t idle = __ballot_sync(0xFFFFFFFFu, a == b);
		if (more && idle) {
			uint32_t base = 0;
			if (lane == 0) base = atomicAdd(&s_cursor, (uint32_t)__popc(idle));
			base = __shfl_sync(0xFFFFFFFFu, base, 0);
			if (a == b) {
				const uint32_t i = base + (uint32_t)__popc(idle & ((1u << lane) - 1u));
				if (i < n) {
					size_t ia, ib;
					bool has_prev;
					seam_position(e0 + s_pos[i], P, TW, TH, tiles_x, RINGW, nv, ia, ib, has_prev);
					a = recP[ia]; b = recP[ib];
				}
			}
			if (base + (uint32_t)__popc(idle) >= n) more = false;
		}
		if (__all_sync(0xFFFFFFFFu, a == b)) { if (!more) break; else continue; }
		if (a != b) {
			// one step of link_g: both ends one hop towards their level roots (path halving), or the link itself
			const uint32_t pa = ld_relaxed(parP + key_idx(a)), pb = ld_relaxed(parP + key_idx(b));
			bool ra = pa == KEY_NONE || key_level(pa) != key_level(a), rb = pb == KEY_NONE || key_level(pb) != key_level(b);
			if (!ra) {
				const uint32_t g = ld_relaxed(parP + key_idx(pa));
				if (g == KEY_NONE || key_level(g) != key_level(a)) { a = pa; ra = true; }
				else { asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(parP + key_idx(a)), "r"(g) : "memory"); a = g; }
			}
			if (!rb) {
				const uint32_t g = ld_relaxed(parP + key_idx(pb));
				if (g == KEY_NONE || key_level(g) != key_level(b)) { b = pb; rb = true; }
				else { asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(parP + key_idx(b)), "r"(g) : "memory"); b = g; }
			}
			if (ra && rb && a != b) {
				const uint32_t lo = min(a, b), hi = max(a, b);
				const uint32_t old = atomicMin(&parP[key_idx(lo)], hi);
				if (old == hi || old == KEY_NONE) a = b = 0;
				else { a = min(hi, old); b = max(hi, old); }
			} else if (ra && rb) a = b = 0;
		}
	}
	guard_hit |= (a != b);
	}
	if (guard_hit) atomicOr(status, ERR_LOOP_GUARD);
}

// ---------------------------------------------------------------------------------------------
// k_fold : aliases (nodes merged into a same-level node of another tile) hand their own-level
// pixels to the final node; final nodes resolve their final parent and register as its child.
// Node i of a plane is slot i of the dense arrays par / attr / node_key; its key is (level, i).
// ---------------------------------------------------------------------------------------------
__global__ void k_fold(ExtractParams P, uint32_t *__restrict__ par_g, NodeAttr *__restrict__ attr_g, uint32_t *__restrict__ node_key,
                       const uint32_t *__restrict__ node_count)
{
	const int plane = blockIdx.y;
	uint32_t *parP = par_g + (size_t)plane * P.node_cap;
	NodeAttr *attrP = attr_g + (size_t)plane * P.node_cap;
	uint32_t *keyP = node_key + (size_t)plane * P.node_cap;
	const uint32_t n = node_count[plane];
	if (n > (uint32_t)P.node_cap) return;   // node-overflow (flagged by the tile kernel): some slots were never written, the plane is void
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		const uint32_t nk = keyP[i];
		const uint32_t g = make_key(key_level(nk), i);
		NodeAttr *ag = &attrP[i];
		if (ag->pend == NODE_COMPLETE) {
			// interior node, finished inside its tile; only its parent may have been merged across a seam
			const uint32_t pk = parP[i];
			if (pk != KEY_NONE) {
				const uint32_t fp = find_g(parP, pk);
				if (fp != pk) parP[i] = fp;
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
			atomicMax(&keyP[key_idx(f)], nk);                  // same level: the final node's root pixel is the largest own-level pixel of all its parts
			ag->nn = 0;   // alias marker
		} else {
			const uint32_t pk = ld_relaxed(parP + i);
			if (pk != KEY_NONE) {
				const uint32_t fp = find_g(parP, pk);
				if (fp != pk) parP[i] = fp;
				atomicAdd(&attrP[key_idx(fp)].pend, 1u);
			}
		}
	}
}

// ---------------------------------------------------------------------------------------------
// k_refit : leaves start; a thread carries a node's finished totals into its parent and continues
// upward only if it was the last child to arrive (no grid-wide synchronisation).
// ---------------------------------------------------------------------------------------------
__global__ void k_refit(ExtractParams P, const uint32_t *__restrict__ par_g, NodeAttr *__restrict__ attr_g, const uint32_t *__restrict__ node_count)
{
	const int plane = blockIdx.y;
	const uint32_t *parP = par_g + (size_t)plane * P.node_cap;
	NodeAttr *attrP = attr_g + (size_t)plane * P.node_cap;
	const uint32_t n = node_count[plane];
	if (n > (uint32_t)P.node_cap) return;
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		uint32_t cur = i;
		{
			const NodeAttr *a = &attrP[cur];
			if (a->nn == 0 || a->pend != 0) continue;   // alias, or not a leaf (pend is constant in this kernel)
		}
		for (int guard = 0; guard < 64; ++guard) {
			const uint32_t pk = parP[cur];
			if (pk == KEY_NONE) break;
			const uint32_t p = key_idx(pk);
			const uint32_t *a = reinterpret_cast<const uint32_t *>(&attrP[cur]);
			const uint32_t c_cnt = ld_relaxed(a + 0), c_nn = ld_relaxed(a + 1);
			const uint32_t c_x0 = ld_relaxed(a + 4), c_y0 = ld_relaxed(a + 5), c_x1 = ld_relaxed(a + 6), c_y1 = ld_relaxed(a + 7);
			NodeAttr *ap = &attrP[p];
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
// the flood starts at pixel 0; if that is a wall it escapes to pixel 1, else to pixel W (neighbour order right,
// bottom, src/ER.cpp:298-341); only that tree is the reference's result.  The tile kernel left the keys of the three
// candidates' nodes in start_key; if all three are walls the result is a lone node (level(pixel 0), area 2).
// ---------------------------------------------------------------------------------------------
__device__ uint32_t reach_root_of(const ExtractParams &P, const PlaneSrc &ps, const uint32_t *parP, const uint32_t *start_key, int32_t *lone)
{
	uint32_t s = start_key[0];
	if (s == KEY_NONE && P.W > 1) s = start_key[1];
	if (s == KEY_NONE && P.H > 1) s = start_key[2];
	if (s == KEY_NONE) {
		int v = __ldg(ps.src);
		if (ps.invert) v = 255 - v;
		*lone = quantize_level(v, P.qscale);
		return KEY_NONE;
	}
	uint32_t k = find_g(parP, s);
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
__global__ void k_emit_kept(ExtractParams P, const uint32_t *__restrict__ par_g, NodeAttr *__restrict__ attr_g, const uint32_t *__restrict__ node_key,
                            const uint32_t *__restrict__ node_count, const uint32_t *__restrict__ start_key, const PlaneSrc *__restrict__ planes,
                            uint32_t *__restrict__ reach_root, int32_t *__restrict__ lone_level, KeptRec *__restrict__ kept,
                            uint32_t *__restrict__ kept_count, uint32_t *status)
{
	const int plane = blockIdx.y;
	const uint32_t *parP = par_g + (size_t)plane * P.node_cap;
	NodeAttr *attrP = attr_g + (size_t)plane * P.node_cap;
	const uint32_t *keyP = node_key + (size_t)plane * P.node_cap;
	const uint32_t n = node_count[plane];
	if (n > (uint32_t)P.node_cap) {           // node-overflow: no tree for this plane (the batch carries the status flag)
		if (blockIdx.x == 0 && threadIdx.x == 0) { reach_root[plane] = KEY_NONE; lone_level[plane] = 0; }
		return;
	}
	// the flood's start-pixel rule, evaluated once per block (a short dependent chain, cheaper than its own launch)
	__shared__ uint32_t s_rr;
	if (threadIdx.x == 0) {
		int32_t lone = -1;
		s_rr = reach_root_of(P, planes[plane], parP, start_key + (size_t)plane * 4, &lone);
		if (blockIdx.x == 0) { reach_root[plane] = s_rr; lone_level[plane] = lone; }
	}
	__syncthreads();
	const uint32_t rr = s_rr;
	if (rr == KEY_NONE) return;
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		NodeAttr *a = &attrP[i];
		if (a->nn == 0) continue;
		const uint32_t nk = keyP[i];
		const uint32_t g = make_key(key_level(nk), i);
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
		r.gidx = key_idx(nk);
		const uint32_t pk = parP[i];
		r.parent = (pk == KEY_NONE) ? KEY_NONE : key_idx(pk);
		r.level = (int32_t)key_level(nk);
		r.area = area;
		r.x0 = (uint16_t)a->x0; r.y0 = (uint16_t)a->y0; r.x1 = (uint16_t)a->x1; r.y1 = (uint16_t)a->y1;
		kept[(size_t)plane * P.kept_cap + pos] = r;
	}
}

// ---------------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------------
int extract_pitch(int W) { return (W + 127) / 128 * 128; }

// tile-kernel variants selectable at run time for A/B (ert_set_tile_config): 0 = default (both options of er_tile.cu),
// 1 = neither, 2 = arrival-counter fold only, 3 = horizontal edge skip only, 4 = as 0 with 48 registers
int tile_config_count() { return 5; }
size_t ring_words_per_plane(int W, int H) { return (size_t)((W + 63) / 64) * ((H + 31) / 32) * 2 * (64 + 32); }

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

int launch_extract(const ExtractParams &P, const PlaneSrc *d_planes, ExtractWork &wk, cudaStream_t st, cudaEvent_t ev_tile_begin, cudaEvent_t ev_tile_end,
                   cudaStream_t st_post)
{
	constexpr int TILE_W = 64, TILE_H = 32;
	ERT_CUDA_CHECK(cudaMemsetAsync(wk.node_count, 0, sizeof(uint32_t) * P.n_planes, st));
	ERT_CUDA_CHECK(cudaMemsetAsync(wk.kept_count, 0, sizeof(uint32_t) * P.n_planes, st));
	if (ev_tile_begin) ERT_CUDA_CHECK(cudaEventRecord(ev_tile_begin, st));
	static const int opt_of_cfg[5] = {3, 0, 1, 2, 4};     // 4: as 0, register allocation for 5 CTAs per SM (48 registers)
	if (launch_tile_v2(P, d_planes, wk, st, opt_of_cfg[(wk.tile_cfg >= 0 && wk.tile_cfg < 5) ? wk.tile_cfg : 0])) return -1;
	if (ev_tile_end) ERT_CUDA_CHECK(cudaEventRecord(ev_tile_end, st));
	if (st_post && st_post != st) {
		// everything after the SM-filling tile kernel runs on the context's high-priority stream: its narrow, latency-bound
		// kernels get CTA slots as soon as tile CTAs of OTHER batches retire, instead of queueing behind whole tile grids
		if (!ev_tile_end) { set_error("launch_extract: the split-stream mode needs the tile-end event"); return -1; }
		ERT_CUDA_CHECK(cudaStreamWaitEvent(st_post, ev_tile_end, 0));
		st = st_post;
	}
	{
		const long long edges = (long long)((P.W - 1) / TILE_W) * P.H + (long long)((P.H - 1) / TILE_H) * P.W;
		if (edges > 0) {
			const int tiles_x = (P.W + TILE_W - 1) / TILE_W, tiles_y = (P.H + TILE_H - 1) / TILE_H;
			if (wk.seam_list) {
				unsigned gx = (unsigned)((edges + SEAM_CHUNK - 1) / SEAM_CHUNK);
				if (wk.post_ctas > 0 && P.n_planes > 12) gx = std::min(gx, (unsigned)std::max(1, wk.post_ctas / P.n_planes));   // small batches are latency-oriented: no cap
				dim3 g2(gx, P.n_planes);
				k_seam_link_list<<<g2, SEAM_NT, 0, st>>>(P, wk.ring_rec, wk.par, wk.status, TILE_W, TILE_H, tiles_x, tiles_x * tiles_y);
			} else {
				dim3 grid((unsigned)((edges + 255) / 256), P.n_planes);
				k_seam_link_rec<<<grid, 256, 0, st>>>(P, wk.ring_rec, wk.par, wk.status, TILE_W, TILE_H, tiles_x, tiles_x * tiles_y);
			}
			ERT_CUDA_CHECK(cudaGetLastError());
		}
	}
	{
		int nb = wk.node_blocks;
		if (wk.post_ctas > 0 && P.n_planes > 12) nb = std::min(nb, std::max(1, wk.post_ctas / P.n_planes));
		dim3 grid(nb, P.n_planes);
		k_fold<<<grid, 256, 0, st>>>(P, wk.par, wk.attr, wk.node_key, wk.node_count);
		ERT_CUDA_CHECK(cudaGetLastError());
		k_refit<<<grid, 256, 0, st>>>(P, wk.par, wk.attr, wk.node_count);
		ERT_CUDA_CHECK(cudaGetLastError());
		k_emit_kept<<<grid, 256, 0, st>>>(P, wk.par, wk.attr, wk.node_key, wk.node_count, wk.start_key, d_planes, wk.reach_root, wk.lone_level, wk.kept,
		                                  wk.kept_count, wk.status);
		ERT_CUDA_CHECK(cudaGetLastError());
	}
	return 0;
}

} // namespace ert
