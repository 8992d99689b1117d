// er_nms.cu -- tree non-maximum suppression (replaces ERFilter::non_maximum_supression,
// src/ER.cpp:416-505).  One CTA per plane.
//
// The reference walks the tree so that a node is processed after all of its children, children in
// list order, and the result depends on that order through the `done` flags.  List order in the
// reference is the (sequential) flood's completion order; the parallel build has no such order, so
// siblings are visited in a CANONICAL order instead: descending (bbox.y, bbox.x), ties by
// descending level, area, x1, y1, root pixel.  On the 211 ICDAR frames of the reference repo this
// changes 1 of 14 668 pooled regions (tests/test_sibling_order.py); against the oracle run with
// the same canonical order the result is bit-exact.  When the caller supplies an explicit child
// order (ert_nms_nodes, used by the drop-in facade on trees it got from us or from the reference)
// that order is used verbatim.
//
// Steps: (1) load the kept list, resolve parent positions; (2) bitonic sort by (parent, sibling
// key) so every child list is contiguous and ordered; (3) the reference's post-order walk with its
// `done` chains (arithmetic in IEEE double exactly as the reference: overlap = area(bbox∩) /
// area(parent bbox) > coef, stability = a_i / (a_{i+T} - a_i)); (4) all threads scatter nodes into
// DFS pre-order (the oracle's dump order) and the pool.
//
// Step (3) exists twice.  The reference's walk is sequential, but its RESULT has a parallel statement:
// every node ends up in exactly one chain (the one that marked it done); a chain is a contiguous piece
// of a root path that starts at the node D visited first; the chain that owns a child c is the only one
// that can continue into its parent p, it does so when D is visited (long before p), and if several
// children's chains qualify the one visited first wins.  Hence
//     owner(p) = owner(c*)  with c* = the FIRST child in visiting order whose owner passes the overlap
//                test against p,            owner(p) = p if there is none,
// a bottom-up recurrence over the tree levels (a parent's level is strictly higher than its children's).
// Chains are then evaluated independently (one thread per chain start), and the pool order -- the order
// in which the walk meets the chain starts -- is the post-order index of the start, which follows from
// subtree sizes and a prefix sum.  The one-thread walk is kept for trees whose levels do not increase
// towards the root (caller-supplied nodes), for overlap_coef >= 1, and as an audit path
// (ert_set_nms_sequential); tests run both and require identical pools.
#include "common.cuh"
#include "kernels.h"

namespace ert {

struct NmsView {
	unsigned long long *keyA;   // [n2] sort key, high part
	uint32_t *keyB;             // [n2] sort key, low part (tie break)
	uint32_t *ord;              // [n2] payload: original position
	int32_t *area;              // [n] sorted-position attributes
	uint16_t *bx;               // [4n] x0,y0,x1,y1
	int32_t *parent;            // [n]
	int32_t *first;             // [n] first child (sorted position) or -1
	int32_t *pre;               // [n] DFS pre-order index
	int32_t *newpos;            // [n] original position -> sorted position
	uint8_t *level;             // [n]
	uint8_t *done;              // [n]
};

__host__ __device__ inline size_t nms_bytes_per_node() { return 8 + 4 + 4 + 4 + 8 + 4 + 4 + 4 + 4 + 1 + 1; }

__host__ __device__ inline size_t nms_view_bytes(int n2) { return (size_t)n2 * nms_bytes_per_node() + 64; }

__device__ inline NmsView make_view(uint8_t *base, int n2)
{
	NmsView v;
	v.keyA = reinterpret_cast<unsigned long long *>(base); base += (size_t)n2 * 8;
	v.keyB = reinterpret_cast<uint32_t *>(base); base += (size_t)n2 * 4;
	v.ord = reinterpret_cast<uint32_t *>(base); base += (size_t)n2 * 4;
	v.area = reinterpret_cast<int32_t *>(base); base += (size_t)n2 * 4;
	v.bx = reinterpret_cast<uint16_t *>(base); base += (size_t)n2 * 8;
	v.parent = reinterpret_cast<int32_t *>(base); base += (size_t)n2 * 4;
	v.first = reinterpret_cast<int32_t *>(base); base += (size_t)n2 * 4;
	v.pre = reinterpret_cast<int32_t *>(base); base += (size_t)n2 * 4;
	v.newpos = reinterpret_cast<int32_t *>(base); base += (size_t)n2 * 4;
	v.level = base; base += n2;
	v.done = base;
	return v;
}

__device__ __forceinline__ int bb_area(const uint16_t *b) { return ((int)b[2] - (int)b[0] + 1) * ((int)b[3] - (int)b[1] + 1); }
__device__ __forceinline__ int bb_inter(const uint16_t *a, const uint16_t *b)
{
	const int x0 = max((int)a[0], (int)b[0]), y0 = max((int)a[1], (int)b[1]);
	const int x1 = min((int)a[2], (int)b[2]), y1 = min((int)a[3], (int)b[3]);
	if (x1 < x0 || y1 < y0) return 0;
	return (x1 - x0 + 1) * (y1 - y0 + 1);
}

// (double)inter / (double)parea > coef, exactly as the reference evaluates it, but without the FP64 divide unless the
// quotient is within 1e-9 (relative) of the threshold: there the rounding of the division decides and it is performed.
__device__ __forceinline__ bool overlap_exceeds(int inter, int parea, double coef)
{
	const double a = (double)inter, t = coef * (double)parea;
	if (fabs(a - t) > 1e-9 * t) return a > t;
	return a / (double)parea > coef;
}

// exclusive prefix sum of a[0..n) in place; returns the total.  Every thread of the CTA must call it.
template <int NT>
__device__ int block_exclusive_scan(int32_t *a, int n, int32_t *s_warp /* [NT/32 + 1] shared */)
{
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int C = (n + NT - 1) / NT;
	const int b = min(n, tid * C), e = min(n, b + C);
	int sum = 0;
	for (int i = b; i < e; i++) sum += a[i];
	int inc = sum;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (lane >= o) inc += t; }
	if (lane == 31) s_warp[warp] = inc;
	__syncthreads();
	if (warp == 0) {
		const int w = (lane < NT / 32) ? s_warp[lane] : 0;
		int winc = w;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xFFFFFFFFu, winc, o); if (lane >= o) winc += t; }
		if (lane < NT / 32) s_warp[lane] = winc - w;
		if (lane == NT / 32 - 1) s_warp[NT / 32] = winc;
	}
	__syncthreads();
	int run = s_warp[warp] + inc - sum;
	for (int i = b; i < e; i++) { const int t = a[i]; a[i] = run; run += t; }
	const int total = s_warp[NT / 32];
	__syncthreads();
	return total;
}

// in:  either the extract stage's kept list (kept != nullptr; parents given as root-pixel indices,
//      resolved through attr[].arr) or caller nodes (in_nodes: DFS pre-order with explicit order).
template <int NT>
__global__ void __launch_bounds__(NT) k_nms(NmsParams P, const KeptRec *__restrict__ kept, const uint32_t *__restrict__ kept_count,
                                            const NodeAttr *__restrict__ attr_g, const uint32_t *__restrict__ reach_root,
                                            const int32_t *__restrict__ lone_level,
                                            const OutNode *__restrict__ in_nodes, const int32_t *__restrict__ in_offsets,
                                            uint8_t *__restrict__ scratch, size_t scratch_stride, int smem_cap,
                                            OutNode *__restrict__ out_nodes, int32_t *__restrict__ out_pool,
                                            int32_t *__restrict__ out_counts /* per plane: n_nodes, n_pool */, uint32_t *status,
                                            int32_t *__restrict__ order_sens /* per plane, may be null */)
{
	extern __shared__ __align__(16) uint8_t nms_smem[];
	if (order_sens && threadIdx.x == 0) order_sens[blockIdx.x] = 0;
	__shared__ int s_npool, s_lmin, s_lmax, s_bad;
	__shared__ int32_t s_scan[NT / 32 + 1];
	const int tid = threadIdx.x;
	const int plane = blockIdx.x;
	const bool from_kept = (kept != nullptr);
	int n;
	const KeptRec *kp = nullptr;
	const OutNode *inp = nullptr;
	if (from_kept) {
		n = (int)min(kept_count[plane], (uint32_t)P.kept_cap);
		kp = kept + (size_t)plane * P.kept_cap;
	} else {
		n = in_offsets[plane + 1] - in_offsets[plane];
		inp = in_nodes + in_offsets[plane];
	}
	OutNode *outn = out_nodes + (size_t)plane * P.kept_cap;
	int32_t *outp = out_pool + (size_t)plane * P.pool_cap;

	if (from_kept && reach_root[plane] == KEY_NONE) {
		// the flood never left a wall pixel: a lone component of level(pixel 0), area 2 (ctor 1 + one pixel)
		if (tid == 0) {
			OutNode o; o.level = lone_level[plane]; o.area = 2; o.x = 0; o.y = 0; o.w = 1; o.h = 1; o.parent = -1; o.nchild = 0;
			outn[0] = o;
			out_counts[2 * plane] = 1; out_counts[2 * plane + 1] = 0;
		}
		return;
	}
	if (n == 0) {
		if (tid == 0) { out_counts[2 * plane] = 0; out_counts[2 * plane + 1] = 0; }
		return;
	}
	int n2 = 4;                       // at least 4: the per-node byte arrays are also accessed as 32-bit words
	while (n2 < n) n2 <<= 1;
	uint8_t *base = (n2 <= smem_cap) ? nms_smem : (scratch + (size_t)plane * scratch_stride);
	NmsView v = make_view(base, n2);

	// ---- (1) keys ----
	for (int i = tid; i < n2; i += NT) {
		if (i < n) {
			int ppos, lvl, ar; uint32_t tie; int x0, y0, x1, y1;
			if (from_kept) {
				const KeptRec r = kp[i];
				ppos = (r.parent == KEY_NONE) ? -1 : (int)attr_g[(size_t)plane * P.N + r.parent].arr;
				if (ppos >= n) ppos = -1;   // parent fell off an overflowing kept list (status already flags the batch)
				lvl = r.level; ar = r.area; tie = r.gidx; x0 = r.x0; y0 = r.y0; x1 = r.x1; y1 = r.y1;
			} else {
				const OutNode r = inp[i];
				ppos = r.parent; lvl = r.level; ar = r.area; tie = (uint32_t)i; x0 = r.x; y0 = r.y; x1 = r.x + r.w - 1; y1 = r.y + r.h - 1;
			}
			unsigned long long a;
			uint32_t b;
			if (from_kept) {
				// parent | ~y0 | ~x0 | ~level  ;  ~area | ~x1(13) ... tie by root pixel
				a = ((unsigned long long)(uint32_t)(ppos + 1) << 40) | ((unsigned long long)(8191 - y0) << 27) |
				    ((unsigned long long)(8191 - x0) << 14) | ((unsigned long long)(63 - lvl) << 8);
				b = 0xFFFFFFFFu - (uint32_t)ar;
				// (x1,y1,root pixel) decide only if level, area, x0, y0 all tie: fold them into the low byte order via ord compare
				v.keyA[i] = a; v.keyB[i] = b;
			} else {
				v.keyA[i] = ((unsigned long long)(uint32_t)(ppos + 1) << 40) | (unsigned long long)tie;   // explicit order
				v.keyB[i] = 0;
			}
			v.ord[i] = (uint32_t)i;
			v.newpos[i] = ppos;   // stash the unsorted parent position
		} else {
			v.keyA[i] = ~0ull; v.keyB[i] = ~0u; v.ord[i] = (uint32_t)i;
		}
	}
	__syncthreads();

	// ---- (2) bitonic sort by (keyA, keyB, tie) ----
	for (int k = 2; k <= n2; k <<= 1) {
		for (int j = k >> 1; j > 0; j >>= 1) {
			for (int i = tid; i < n2; i += NT) {
				const int l = i ^ j;
				if (l > i) {
					const unsigned long long ai = v.keyA[i], al = v.keyA[l];
					const uint32_t bi = v.keyB[i], bl = v.keyB[l];
					const uint32_t oi = v.ord[i], ol = v.ord[l];
					bool gt;
					if (ai != al) gt = ai > al;
					else if (bi != bl) gt = bi > bl;
					else {
						// full tie on (parent,y0,x0,level,area): x1 desc, y1 desc, root pixel desc
						if (from_kept && oi < (uint32_t)n && ol < (uint32_t)n) {
							const KeptRec ri = kp[oi], rl = kp[ol];
							if (ri.x1 != rl.x1) gt = ri.x1 < rl.x1;
							else if (ri.y1 != rl.y1) gt = ri.y1 < rl.y1;
							else gt = ri.gidx < rl.gidx;
						} else gt = oi > ol;
					}
					const bool up = ((i & k) == 0);
					if (gt == up) {
						v.keyA[i] = al; v.keyA[l] = ai;
						v.keyB[i] = bl; v.keyB[l] = bi;
						v.ord[i] = ol; v.ord[l] = oi;
					}
				}
			}
			__syncthreads();
		}
	}

	// ---- (3) build the sorted tree arrays ----
	// newpos currently holds unsorted parent positions; move them aside through `pre`
	for (int i = tid; i < n; i += NT) v.pre[i] = v.newpos[i];
	__syncthreads();
	for (int j = tid; j < n; j += NT) v.newpos[v.ord[j]] = j;
	__syncthreads();
	for (int j = tid; j < n; j += NT) {
		const int o = (int)v.ord[j];
		int lvl, ar, x0, y0, x1, y1;
		if (from_kept) { const KeptRec r = kp[o]; lvl = r.level; ar = r.area; x0 = r.x0; y0 = r.y0; x1 = r.x1; y1 = r.y1; }
		else { const OutNode r = inp[o]; lvl = r.level; ar = r.area; x0 = r.x; y0 = r.y; x1 = r.x + r.w - 1; y1 = r.y + r.h - 1; }
		v.level[j] = (uint8_t)lvl; v.area[j] = ar;
		v.bx[4 * j] = (uint16_t)x0; v.bx[4 * j + 1] = (uint16_t)y0; v.bx[4 * j + 2] = (uint16_t)x1; v.bx[4 * j + 3] = (uint16_t)y1;
		const int pp = v.pre[o];
		v.parent[j] = (pp < 0) ? -1 : v.newpos[pp];
		v.first[j] = -1;
		v.done[j] = 0;
	}
	__syncthreads();
	if (tid == 0) { s_lmin = 255; s_lmax = 0; s_bad = 0; }
	__syncthreads();
	for (int j = tid; j < n; j += NT) {
		const int p = v.parent[j];
		if (p >= 0 && (j == 0 || v.parent[j - 1] != p)) v.first[p] = j;
		const int lj = v.level[j];
		atomicMin(&s_lmin, lj); atomicMax(&s_lmax, lj);
		if ((p >= 0 && v.level[p] <= lj) || (p < 0 && j != 0)) s_bad = 1;   // not a level-ordered single tree: use the walk
	}
	__syncthreads();
	const bool parallel_walk = !P.sequential_walk && !s_bad && P.overlap_coef < 1.0;

	if (parallel_walk) {
		// ---- (4p) level-parallel statement of the walk's result (see the header) ----
		int32_t *owner = reinterpret_cast<int32_t *>(v.keyA);     // chain start that owns the node   (sort keys are dead now)
		int32_t *size = owner + n2;                               // subtree size
		int32_t *fpass = reinterpret_cast<int32_t *>(v.keyB);     // first child whose chain continues into this node; later: pool marks
		int32_t *S = reinterpret_cast<int32_t *>(v.ord);          // prefix sums
		int32_t *depth = v.newpos;
		const int T = P.stability_t;
		const int lmin = s_lmin, lmax = s_lmax;
		for (int j = tid; j < n; j += NT) { size[j] = 1; fpass[j] = 0x7FFFFFFF; }
		// v.done (a byte per node, the sequential walk's flag) counts here how many children's chains may continue into a
		// node: bit 0 = one, bit 1 = two or more.  With two or more the walk's outcome depends on which sibling is visited
		// first -- the ONLY place where the canonical sibling order can differ from the reference's flood order.
		uint32_t *pass_bits = reinterpret_cast<uint32_t *>(v.done);
		for (int j = tid; j < (n + 3) / 4; j += NT) pass_bits[j] = 0;
		__syncthreads();
		for (int L = lmin; L <= lmax; ++L) {                      // bottom-up: children (lower levels) are final
			for (int j = tid; j < n; j += NT) {
				if (v.level[j] != L) continue;
				const int fp = fpass[j];
				const int o = (fp != 0x7FFFFFFF) ? owner[fp] : j;
				owner[j] = o;
				const int p = v.parent[j];
				if (p >= 0) {
					atomicAdd(&size[p], size[j]);
					// the test the walk makes when the chain started at o, having taken j, looks at p (src/ER.cpp:457)
					if (overlap_exceeds(bb_inter(&v.bx[4 * o], &v.bx[4 * p]), bb_area(&v.bx[4 * p]), P.overlap_coef)) {
						atomicMin(&fpass[p], j);
						const uint32_t sh = 8u * ((uint32_t)p & 3u);
						if (atomicOr(&pass_bits[p >> 2], 1u << sh) & (1u << sh)) atomicOr(&pass_bits[p >> 2], 2u << sh);
					}
				}
			}
			__syncthreads();
		}
		if (order_sens) {
			int contested = 0;
			for (int j = tid; j < n; j += NT) contested += (v.done[j] >> 1) & 1;
			contested = __reduce_add_sync(0xFFFFFFFFu, contested);
			if ((tid & 31) == 0 && contested) atomicAdd(&order_sens[plane], contested);
		}
		// DFS pre-order index and depth, top-down: pre(child) = pre(parent) + 1 + sizes of the siblings before it
		for (int j = tid; j < n; j += NT) S[j] = size[j];
		__syncthreads();
		block_exclusive_scan<NT>(S, n, s_scan);
		for (int L = lmax; L >= lmin; --L) {
			for (int j = tid; j < n; j += NT) {
				if (v.level[j] != L) continue;
				const int p = v.parent[j];
				if (p < 0) { v.pre[j] = 0; depth[j] = 0; }
				else { v.pre[j] = v.pre[p] + 1 + (S[j] - S[v.first[p]]); depth[j] = depth[p] + 1; }
			}
			__syncthreads();
		}
		// chains: one thread per chain start; an accepted chain marks its start's POST-order slot with the chosen node
		for (int j = tid; j < n; j += NT) fpass[j] = 0;
		__syncthreads();
		for (int j = tid; j < n; j += NT) {
			if (owner[j] != j) continue;
			int len = 0, hi = j;
			for (int p = j; p >= 0 && owner[p] == j; p = v.parent[p]) { if (len == T) hi = p; len++; }   // hi = T-th element of the chain
			if (len > 72) { atomicOr(status, ERR_NMS_OVERFLOW); len = 72; }
			if (len < 1 + T) continue;
			// stability_i = a_i / (a_{i+T} - a_i), arg-max with ties to the smaller area then the earlier element (src/ER.cpp:470-484)
			int best = j, best_area = 0, lo = j;
			long long bn = 0, bd = 1;
			double best_s = 0.0;
			const bool small = (long long)P.W * P.H < (1ll << 24);
			for (int i = 0; i < len - T; i++) {
				const int ai = bb_area(&v.bx[4 * lo]), aj = bb_area(&v.bx[4 * hi]);
				if (small) {
					const long long n_i = ai, d_i = (long long)aj - ai;
					int cmp;
					if (i == 0) cmp = 1;
					else if (d_i == 0 || bd == 0) cmp = (d_i == 0 && bd == 0) ? 0 : (d_i == 0 ? 1 : -1);
					else { const long long l = n_i * bd, r = bn * d_i; cmp = (l > r) - (l < r); }
					if (cmp > 0 || (cmp == 0 && i > 0 && ai < best_area)) { best = lo; best_area = ai; bn = n_i; bd = d_i; }
				} else {
					const double sdiv = (double)ai / (double)(aj - ai);
					if (i == 0 || sdiv > best_s || (sdiv == best_s && ai < best_area)) { best = lo; best_area = ai; best_s = sdiv; }
				}
				lo = v.parent[lo]; hi = v.parent[hi];
			}
			const int w = (int)v.bx[4 * best + 2] - (int)v.bx[4 * best] + 1, h = (int)v.bx[4 * best + 3] - (int)v.bx[4 * best + 1] + 1;
			const double ar = (double)w / (double)h;
			if (ar < 2.0 && ar > 0.10 && v.area[best] < P.max_area && v.area[best] > P.min_area && h < P.H * 0.8 && w < P.W * 0.8)
				fpass[v.pre[j] - depth[j] + size[j] - 1] = best + 1;      // post-order index of the chain start
		}
		__syncthreads();
		for (int j = tid; j < n; j += NT) S[j] = fpass[j] != 0;
		__syncthreads();
		const int total = block_exclusive_scan<NT>(S, n, s_scan);
		for (int j = tid; j < n; j += NT) {
			const int b = fpass[j];
			if (b != 0 && S[j] < P.pool_cap) outp[S[j]] = v.pre[b - 1];
		}
		if (tid == 0) {
			if (total > P.pool_cap) atomicOr(status, ERR_POOL_OVERFLOW);
			s_npool = min(total, P.pool_cap);
		}
		__syncthreads();
	} else {
	// one 16-byte record per node for the sequential walk (aliases the sort keys, which are dead now):
	// {parent, first child, x0 | y0 << 16, x1 | y1 << 16} -> one 128-bit load per visited node
	uint4 *rec = reinterpret_cast<uint4 *>(v.keyA);
	for (int j = tid; j < n; j += NT)
		rec[j] = make_uint4((uint32_t)v.parent[j], (uint32_t)v.first[j], (uint32_t)v.bx[4 * j] | ((uint32_t)v.bx[4 * j + 1] << 16),
		                    (uint32_t)v.bx[4 * j + 2] | ((uint32_t)v.bx[4 * j + 3] << 16));
	__syncthreads();

	// ---- (4) the reference's sequential walk (src/ER.cpp:426-502) ----
	__shared__ int stack[72], chain[72], chain_area[72];
	if (tid == 0) {
		int sp = 0, pre = 0, npool = 0;
		const int T = P.stability_t;
		int cur = 0;   // the root sorts first (parent -1)
		for (long long guard = 0; guard < 4ll * n + 8; ++guard) {
			while (cur >= 0) {
				if (sp >= 72) { atomicOr(status, ERR_NMS_OVERFLOW); cur = -1; break; }
				stack[sp++] = cur; v.pre[cur] = pre++; cur = (int)rec[cur].y;
			}
			if (sp == 0) break;
			cur = stack[--sp];
			const uint4 rc = rec[cur];
			if (!v.done[cur]) {
				int len = 0, p = cur;
				const int cx0 = rc.z & 0xFFFF, cy0 = rc.z >> 16, cx1 = rc.w & 0xFFFF, cy1 = rc.w >> 16;
				uint4 rp = rc;
				for (;;) {
					if (v.done[p]) break;
					const int px0 = rp.z & 0xFFFF, py0 = rp.z >> 16, px1 = rp.w & 0xFFFF, py1 = rp.w >> 16;
					const int ix0 = max(cx0, px0), iy0 = max(cy0, py0), ix1 = min(cx1, px1), iy1 = min(cy1, py1);
					const int inter = (ix1 < ix0 || iy1 < iy0) ? 0 : (ix1 - ix0 + 1) * (iy1 - iy0 + 1);
					const int parea = (px1 - px0 + 1) * (py1 - py0 + 1);
					if (!overlap_exceeds(inter, parea, P.overlap_coef)) break;
					v.done[p] = 1;
					if (len < 72) { chain[len] = p; chain_area[len] = parea; }
					len++;
					const int pp = (int)rp.x;
					if (pp >= 0) { p = pp; rp = rec[p]; }   // root->parent = root (src/ER.cpp:424): same p, now done
				}
				if (len > 72) { atomicOr(status, ERR_NMS_OVERFLOW); len = 72; }
				if (len >= 1 + T) {
					// stability_i = a_i / (a_{i+T} - a_i) as IEEE doubles in the reference (x/0 = +inf).  For areas below 2^24
					// two such quotients are equal as doubles iff they are equal as rationals, so the arg-max is decided by
					// exact 64-bit cross-multiplication (no FP64 divides on this one-thread critical path).
					int best = 0;
					long long bn = 0, bd = 1;   // best = bn / bd (bd == 0: +inf)
					double best_s = 0.0;
					const bool small = (long long)P.W * P.H < (1ll << 24);
					for (int i = 0; i < len - T; i++) {
						const int ai = chain_area[i], aj = chain_area[i + T];
						int cmp;   // sign of (s_i - s_best)
						if (small) {
							const long long n_i = ai, d_i = (long long)aj - ai;
							if (i == 0) cmp = 1;
							else if (d_i == 0 || bd == 0) cmp = (d_i == 0 && bd == 0) ? 0 : (d_i == 0 ? 1 : -1);
							else { const long long l = n_i * bd, r = bn * d_i; cmp = (l > r) - (l < r); }
							if (cmp > 0 || (cmp == 0 && i > 0 && ai < chain_area[best])) { best = i; bn = n_i; bd = d_i; }
						} else {
							const double sdiv = (double)ai / (double)(aj - ai);
							if (i == 0) { best = 0; best_s = sdiv; }
							else if (sdiv > best_s) { best = i; best_s = sdiv; }
							else if (sdiv == best_s && ai < chain_area[best]) { best = i; best_s = sdiv; }
						}
					}
					const int b = chain[best];
					const uint4 rb = rec[b];
					const int w = (int)(rb.w & 0xFFFF) - (int)(rb.z & 0xFFFF) + 1, h = (int)(rb.w >> 16) - (int)(rb.z >> 16) + 1;
					const double ar = (double)w / (double)h;
					if (ar < 2.0 && ar > 0.10 && v.area[b] < P.max_area && v.area[b] > P.min_area &&
					    h < P.H * 0.8 && w < P.W * 0.8) {
						if (npool < P.pool_cap) v.newpos[npool] = b;   // newpos is free after the build step
						else atomicOr(status, ERR_POOL_OVERFLOW);
						npool++;
					}
				}
			}
			// next sibling: the following sorted position if it has the same parent
			const int nx = cur + 1;
			cur = (nx < n && (int)rc.x >= 0 && (int)rec[nx].x == (int)rc.x) ? nx : -1;
		}
		s_npool = min(npool, P.pool_cap);
	}
	__syncthreads();
	}

	// ---- (5) scatter into DFS pre-order ----
	for (int j = tid; j < n; j += NT) {
		OutNode o;
		o.level = v.level[j]; o.area = v.area[j];
		o.x = v.bx[4 * j]; o.y = v.bx[4 * j + 1];
		o.w = v.bx[4 * j + 2] - v.bx[4 * j] + 1; o.h = v.bx[4 * j + 3] - v.bx[4 * j + 1] + 1;
		o.parent = (v.parent[j] < 0) ? -1 : v.pre[v.parent[j]];
		int nc = 0;
		const int f = v.first[j];
		if (f >= 0) { nc = 1; while (f + nc < n && v.parent[f + nc] == j) nc++; }
		o.nchild = nc;
		outn[v.pre[j]] = o;
	}
	const int npool = s_npool;
	if (!parallel_walk) for (int k = tid; k < npool; k += NT) outp[k] = v.pre[v.newpos[k]];
	if (tid == 0) { out_counts[2 * plane] = n; out_counts[2 * plane + 1] = npool; }
}

constexpr int NMS_NT = 512;
constexpr int NMS_SMEM_NODES = 2048;   // larger planes fall back to the global scratch view

size_t nms_scratch_stride(int kept_cap)
{
	int n2 = 1;
	while (n2 < kept_cap) n2 <<= 1;
	return (nms_view_bytes(n2) + 255) / 256 * 256;
}

int launch_nms(const NmsParams &P, int n_planes, const KeptRec *kept, const uint32_t *kept_count, const NodeAttr *attr,
               const uint32_t *reach_root, const int32_t *lone_level, const OutNode *in_nodes, const int32_t *in_offsets,
               uint8_t *scratch, size_t scratch_stride, OutNode *out_nodes, int32_t *out_pool, int32_t *out_counts,
               uint32_t *status, cudaStream_t st, int32_t *order_sens)
{
	const size_t smem = nms_view_bytes(NMS_SMEM_NODES);
		ERT_CUDA_CHECK(cudaFuncSetAttribute(k_nms<NMS_NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	k_nms<NMS_NT><<<n_planes, NMS_NT, smem, st>>>(P, kept, kept_count, attr, reach_root, lone_level, in_nodes, in_offsets,
	                                              scratch, scratch_stride, NMS_SMEM_NODES, out_nodes, out_pool, out_counts, status, order_sens);
	ERT_CUDA_CHECK(cudaGetLastError());
	return 0;
}

} // namespace ert
