// er_tile.cu -- k_tile_build2: the tile-local component-tree build of er_tree_extract (src/ER.cpp:240-413),
// second generation.  Global outputs: nodes that leave the tile get a dense SLOT in the plane's par / attr / node_key arrays,
// seam records name the (level, slot) that stands for every side pixel.  Rebuilt around four ideas measured against the
// round-1 kernel (k_tile_build: issue-slot bound, ~25 warp instructions per pixel, 130 SASS instructions per iteration
// of its union loop):
//   * the 64x32 tile arrives WITH its one-pixel halo as ONE tensor-map TMA box (cp.async.bulk.tensor.3d, SASS UTMALDG;
//     96x34 bytes, out-of-plane bytes zero-filled by the TMA unit) instead of 32 row copies + ~190 scalar halo loads;
//   * the per-pixel sweeps (quantise, horizontal runs, edge list) are vectorised: one lane owns 4 consecutive pixels
//     (one 32-bit word of levels, one 128-bit vector of forest words), a warp owns two full tile rows, runs span the
//     whole row, run ends and edges are compacted into lists with ballots / one shuffle scan per 128 pixels;
//   * per-run work (own-level pixel count, bbox) runs over the compacted RUN LIST (one run per lane, all lanes busy);
//   * the union loop works on pre-keyed 64-bit edge records through 32-bit shared-memory addresses (explicit
//     ld/st/atom.shared): no address re-materialisation, no byte loads, one warp vote per iteration.
// Semantics are unchanged: keyed lock-free union-find (key = level << 16 | pixel), DESIGN.md section 5.1.
#include "common.cuh"
#include "kernels.h"
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cstring>

namespace ert {

namespace t2 {
constexpr int TW = 64, TH = 32, TPX = TW * TH, NT = 256, NWARP = NT / 32;
constexpr int BOXW = 96, BOXH = 34, XOFF = 16, YOFF = 1;      // haloed box: tile pixel (x, y) sits at box[(y + 1) * 96 + 16 + x]
constexpr int BOX_BYTES = BOXW * BOXH;                         // 3264
constexpr int LV0 = YOFF * BOXW + XOFF;                        // 112
constexpr int OFF_BOX = 0;
constexpr int OFF_PAR = 3328;                                  // u32[TPX]   keyed forest
constexpr int OFF_ATTR = OFF_PAR + TPX * 4;                    // u32[4][TPX] cnt / xmn / xmx / ymask, aliased by the edge records (uint2[<= 4000])
constexpr int OFF_RUNS = OFF_ATTR + 4 * TPX * 4;               // u16[TPX]   run records, 128 per row pair
constexpr int OFF_ROOTS = OFF_RUNS + TPX * 2;                  // u16[TPX]   level roots
constexpr int SMEM_BYTES = OFF_ROOTS + TPX * 2 + 128;          // + slack for the 128-byte alignment of the TMA destination
constexpr int RING = 2 * (TW + TH);
constexpr uint32_t FULL = 0xFFFFFFFFu;
constexpr uint32_t REC_WALL = 0x800u, REC_IDX = 0x7FFu;
}

// ---- explicit shared-memory accesses through 32-bit shared addresses ------------------------------------------
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ uint32_t lds_u16(uint32_t a) { uint32_t v; asm volatile("{\n\t.reg .u16 t;\n\tld.shared.u16 t, [%1];\n\tcvt.u32.u16 %0, t;\n\t}" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ uint32_t lds_u8(uint32_t a) { uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ uint2 lds_v2(uint32_t a) { uint2 v; asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ uint4 lds_v4(uint32_t a) { uint4 v; asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_u16(uint32_t a, uint32_t v) { asm volatile("{\n\t.reg .u16 t;\n\tcvt.u16.u32 t, %1;\n\tst.shared.u16 [%0], t;\n\t}" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_v2(uint32_t a, uint32_t x, uint32_t y) { asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(x), "r"(y) : "memory"); }
__device__ __forceinline__ void sts_v4(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) { asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory"); }
__device__ __forceinline__ uint32_t atoms_min(uint32_t a, uint32_t v) { uint32_t o; asm volatile("atom.shared.min.u32 %0, [%1], %2;" : "=r"(o) : "r"(a), "r"(v) : "memory"); return o; }
__device__ __forceinline__ void sts_u32_if(bool p, uint32_t a, uint32_t v)
{
	asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %0, 0;\n\t@q st.shared.u32 [%1], %2;\n\t}" ::"r"((uint32_t)p), "r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t atoms_min_if(bool p, uint32_t a, uint32_t v, uint32_t dflt)
{
	uint32_t o = dflt;
	asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %1, 0;\n\t@q atom.shared.min.u32 %0, [%2], %3;\n\t}" : "+r"(o) : "r"((uint32_t)p), "r"(a), "r"(v) : "memory");
	return o;
}
__device__ __forceinline__ void reds_add(uint32_t a, uint32_t v) { asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void reds_min(uint32_t a, uint32_t v) { asm volatile("red.shared.min.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void reds_max(uint32_t a, uint32_t v) { asm volatile("red.shared.max.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void reds_or(uint32_t a, uint32_t v) { asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

__device__ __forceinline__ uint32_t t2_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// levels differ (also true when q is KEY_NONE: its level field is 0xFFFF)
__device__ __forceinline__ bool lvl_differs(uint32_t q, uint32_t k) { return (q ^ k) > 0xFFFFu; }

// OPT bit 0: fold interior subtrees with arrival counters (two CTA barriers) instead of one CTA barrier per level
// OPT bit 1: skip a horizontal edge whose pixel pair repeats the pair right above it (joined through two same-level vertical pairs)
// MINB: resident CTAs per SM the register allocation aims at (shared memory allows 4; 5 caps the kernel at 48 registers,
// which leaves room for two small CTAs of the other stages beside four tile CTAs)
template <int OPT, int MINB>
__global__ void __launch_bounds__(t2::NT, MINB)
k_tile_build2(const __grid_constant__ CUtensorMap tmap, ExtractParams P, const PlaneSrc *__restrict__ planes, uint32_t *__restrict__ par_g,
              NodeAttr *__restrict__ attr_g, uint32_t *__restrict__ node_key, uint32_t *__restrict__ node_count, uint32_t *__restrict__ start_key,
              uint32_t *status, int tiles_x, unsigned long long *prof, uint32_t *__restrict__ ring_rec)
{
	using namespace t2;
	long long t_prev = prof ? clock64() : 0;
#define ERT_PHASE(i) do { if (prof && threadIdx.x == 0) { const long long t_now = clock64(); atomicAdd(&prof[i], (unsigned long long)(t_now - t_prev)); t_prev = t_now; } } while (0)
	extern __shared__ uint8_t smem_raw[];
	__shared__ __align__(8) uint64_t bar;
	__shared__ uint32_t s_nroots, s_base, s_cursor, s_nlinks, s_nemit, s_minlvl, s_maxlvl;
	__shared__ uint16_t s_ringA[RING];   // per tile-side position: the node that stands for it across the seam (0xFFFF: none)

	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int plane = blockIdx.y;
	const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
	const int X0 = tx * TW, Y0 = ty * TH;
	const int rows = min(TH, P.H - Y0), cols = min(TW, P.W - X0);
	const PlaneSrc ps = planes[plane];

	const uint32_t sraw = t2_smem_u32(smem_raw);
	const uint32_t sb = (sraw + 127u) & ~127u;
	uint8_t *sm = smem_raw + (sb - sraw);
	const uint32_t box_s = sb + OFF_BOX, par_s = sb + OFF_PAR, attr_s = sb + OFF_ATTR, runs_s = sb + OFF_RUNS;
	uint8_t *box = sm + OFF_BOX;
	uint32_t *par = reinterpret_cast<uint32_t *>(sm + OFF_PAR);
	uint32_t *cnt = reinterpret_cast<uint32_t *>(sm + OFF_ATTR);
	uint32_t *xmn = cnt + TPX, *xmx = xmn + TPX, *ymask = xmx + TPX;
	uint16_t *rootlist = reinterpret_cast<uint16_t *>(sm + OFF_ROOTS);
#define ERT_LV(p) ((uint32_t)box[((p) >> 6) * BOXW + ((p) & 63) + LV0])

	// ---- phase 0: ONE tensor-map TMA copy brings the tile and its halo ring (out-of-plane bytes arrive as zeros) ----
	if (tid == 0) {
		asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(t2_smem_u32(&bar)), "r"(1) : "memory");
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
		s_nroots = 0; s_cursor = 0; s_nlinks = 0; s_nemit = 0; s_minlvl = 255; s_maxlvl = 0;
		asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(t2_smem_u32(&bar)), "r"((uint32_t)BOX_BYTES) : "memory");
		asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
		             ::"r"(box_s), "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(X0 - XOFF), "r"(Y0 - YOFF), "r"(ps.z), "r"(t2_smem_u32(&bar)) : "memory");
	}
	for (int i = tid; i < RING; i += NT) s_ringA[i] = 0xFFFFu;
	if (warp == 0) {
		// one warp polls, the others park at the CTA barrier (no issue slots burnt on polling)
		uint32_t ok;
		do {
			asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
			             : "=r"(ok) : "r"(t2_smem_u32(&bar)), "r"(0) : "memory");
		} while (!ok);
	}
	__syncthreads();
	ERT_PHASE(0);

	// ---- phase Q: quantise the box in place (columns 12..83 = the tile and one halo pixel per side, word granular);
	// walls (level >= hi, src/ER.cpp:343-358) and everything outside the plane become 255 ----
	for (int i = tid; i < BOXH * 18; i += NT) {
		const int r = i / 18, wx = 3 + (i - r * 18);
		const uint32_t a = box_s + (uint32_t)(r * BOXW + wx * 4);
		uint32_t w = lds_u32(a);
		if (ps.invert) w = ~w;
		const int gy = Y0 - YOFF + r, gx0 = X0 - XOFF + wx * 4;
		const bool rowok = (unsigned)gy < (unsigned)P.H;
		uint32_t out = 0;
#pragma unroll
		for (int j = 0; j < 4; j++) {
			int L = quantize_level((int)((w >> (8 * j)) & 255u), P.qscale);
			if (L >= P.hi || !rowok || (unsigned)(gx0 + j) >= (unsigned)P.W) L = 255;
			out |= (uint32_t)L << (8 * j);
		}
		sts_u32(a, out);
	}
	__syncthreads();
	ERT_PHASE(1);

	// ---- phase A: horizontal same-level runs.  A lane owns 4 consecutive pixels, a warp two tile rows; every pixel of a
	// run points at the run's LAST pixel (no atomics), every run end (walls too, flagged) becomes a run record ----
	uint32_t nruns[2];
	const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
	for (int pass = 0; pass < 2; ++pass) {
		const int rp = warp + NWARP * pass;
		const int y = 2 * rp + (lane >> 4), x0 = (lane & 15) * 4;
		const uint32_t p0 = (uint32_t)(y * TW + x0);
		const uint32_t Lw = lds_u32(box_s + (uint32_t)((y + YOFF) * BOXW + XOFF + x0));
		const uint32_t L0 = Lw & 255u, L1 = (Lw >> 8) & 255u, L2 = (Lw >> 16) & 255u, L3 = Lw >> 24;
		const uint32_t Ln = __shfl_down_sync(FULL, Lw, 1) & 255u;
		const bool e0 = (L0 == L1) && (L0 != 255u);
		const bool e1 = (L1 == L2) && (L1 != 255u);
		const bool e2 = (L2 == L3) && (L2 != 255u);
		const bool e3 = (L3 == Ln) && (L3 != 255u) && ((lane & 15) != 15);
		const uint32_t cm = __ballot_sync(FULL, e0 && e1 && e2 && e3);     // lanes whose 4 pixels all continue to the right
		const int f = !e0 ? 0 : (!e1 ? 1 : (!e2 ? 2 : 3));                 // first run end inside the lane
		const uint32_t nc = (~cm >> 1) >> lane;                            // bit i: lane + 1 + i does not fully continue
		const int ln = lane + __ffs(nc);                                   // (lanes 15 / 31 never continue: e3 is false there)
		const int fn = __shfl_sync(FULL, f, ln & 31);
		const uint32_t cont_end = (uint32_t)(y * TW + (ln & 15) * 4 + fn);
		const uint32_t end3 = e3 ? cont_end : p0 + 3u;
		const uint32_t end2 = e2 ? end3 : p0 + 2u;
		const uint32_t end1 = e1 ? end2 : p0 + 1u;
		const uint32_t end0 = e0 ? end1 : p0;
		sts_v4(par_s + p0 * 4u, e0 ? ((L0 << 16) | end0) : KEY_NONE, e1 ? ((L1 << 16) | end1) : KEY_NONE, e2 ? ((L2 << 16) | end2) : KEY_NONE,
		       e3 ? ((L3 << 16) | end3) : KEY_NONE);
		const uint32_t b0 = __ballot_sync(FULL, !e0), b1 = __ballot_sync(FULL, !e1), b2 = __ballot_sync(FULL, !e2), b3 = __ballot_sync(FULL, !e3);
		uint32_t ra = runs_s + (uint32_t)(rp * 128 + __popc(b0 & lt) + __popc(b1 & lt) + __popc(b2 & lt) + __popc(b3 & lt)) * 2u;
		if (!e0) { sts_u16(ra, p0 | (L0 == 255u ? REC_WALL : 0u)); ra += 2u; }
		if (!e1) { sts_u16(ra, (p0 + 1u) | (L1 == 255u ? REC_WALL : 0u)); ra += 2u; }
		if (!e2) { sts_u16(ra, (p0 + 2u) | (L2 == 255u ? REC_WALL : 0u)); ra += 2u; }
		if (!e3) { sts_u16(ra, (p0 + 3u) | (L3 == 255u ? REC_WALL : 0u)); }
		nruns[pass] = (uint32_t)(__popc(b0) + __popc(b1) + __popc(b2) + __popc(b3));
		const uint32_t mn = min(min(L0, L1), min(L2, L3));
		const uint32_t mx = max(max(L0 == 255u ? 0u : L0, L1 == 255u ? 0u : L1), max(L2 == 255u ? 0u : L2, L3 == 255u ? 0u : L3));
		const uint32_t wmn = __reduce_min_sync(FULL, mn), wmx = __reduce_max_sync(FULL, mx);
		if (lane == 0) { if (wmn < 255u) atomicMin(&s_minlvl, wmn); atomicMax(&s_maxlvl, wmx); }
	}
	__syncthreads();
	ERT_PHASE(2);

	// ---- phase B1: the remaining in-tile edges as pre-keyed 64-bit records (key of p's run end, key of q's run end).
	// Horizontal: neighbouring runs of different level.  Vertical: pixel over pixel, skipped when the pair to its left
	// has the same two levels (it joins the same two runs). ----
	const uint32_t links_s = attr_s;
#pragma unroll
	for (int pass = 0; pass < 2; ++pass) {
		const int rp = warp + NWARP * pass;
		const int y = 2 * rp + (lane >> 4), x0 = (lane & 15) * 4;
		const uint32_t p0 = (uint32_t)(y * TW + x0);
		const bool has_below = y < TH - 1;
		const uint32_t Lw = lds_u32(box_s + (uint32_t)((y + YOFF) * BOXW + XOFF + x0));
		const uint32_t Lbw = has_below ? lds_u32(box_s + (uint32_t)((y + YOFF + 1) * BOXW + XOFF + x0)) : 0xFFFFFFFFu;
		// row above (inside the tile only): 0xFE never equals a level, so nothing is skipped in the tile's first row
		const uint32_t Law = ((OPT & 2) && y > 0) ? lds_u32(box_s + (uint32_t)((y + YOFF - 1) * BOXW + XOFF + x0)) : 0xFEFEFEFEu;
		const uint4 pw = lds_v4(par_s + p0 * 4u);
		uint4 qw = make_uint4(KEY_NONE, KEY_NONE, KEY_NONE, KEY_NONE);
		if (has_below) qw = lds_v4(par_s + (p0 + TW) * 4u);
		const uint32_t L0 = Lw & 255u, L1 = (Lw >> 8) & 255u, L2 = (Lw >> 16) & 255u, L3 = Lw >> 24;
		const uint32_t B0 = Lbw & 255u, B1 = (Lbw >> 8) & 255u, B2 = (Lbw >> 16) & 255u, B3 = Lbw >> 24;
		// keys of the run ends (a pixel that is not a run end points at its run end: same level by construction)
		const uint32_t ka0 = (L0 << 16) | (pw.x != KEY_NONE ? (pw.x & 0xFFFFu) : p0);
		const uint32_t ka1 = (L1 << 16) | (pw.y != KEY_NONE ? (pw.y & 0xFFFFu) : p0 + 1u);
		const uint32_t ka2 = (L2 << 16) | (pw.z != KEY_NONE ? (pw.z & 0xFFFFu) : p0 + 2u);
		const uint32_t ka3 = (L3 << 16) | (pw.w != KEY_NONE ? (pw.w & 0xFFFFu) : p0 + 3u);
		const uint32_t kb0 = (B0 << 16) | (qw.x != KEY_NONE ? (qw.x & 0xFFFFu) : p0 + TW);
		const uint32_t kb1 = (B1 << 16) | (qw.y != KEY_NONE ? (qw.y & 0xFFFFu) : p0 + TW + 1u);
		const uint32_t kb2 = (B2 << 16) | (qw.z != KEY_NONE ? (qw.z & 0xFFFFu) : p0 + TW + 2u);
		const uint32_t kb3 = (B3 << 16) | (qw.w != KEY_NONE ? (qw.w & 0xFFFFu) : p0 + TW + 3u);
		const uint32_t ka4 = __shfl_down_sync(FULL, ka0, 1);                 // first pixel of the lane to the right
		const uint32_t Lm = __shfl_up_sync(FULL, L3, 1), Bm = __shfl_up_sync(FULL, B3, 1);   // last pixel of the lane to the left
		const bool first = (lane & 15) == 0, last = (lane & 15) == 15;
		const uint32_t L4 = ka4 >> 16;
		bool h0 = (L0 != L1) && (L0 != 255u) && (L1 != 255u);
		bool h1 = (L1 != L2) && (L1 != 255u) && (L2 != 255u);
		bool h2 = (L2 != L3) && (L2 != 255u) && (L3 != 255u);
		bool h3 = !last && (L3 != L4) && (L3 != 255u) && (L4 != 255u);
		if (OPT & 2) {
			// the pair right above has the same two levels: (x, y) ~ (x, y-1) and (x+1, y) ~ (x+1, y-1) are same-level vertical
			// pairs (united by their own edges, or by what their skip rule leans on: pairs further LEFT), and the pair above is
			// joined by its own horizontal edge (or, recursively, by the row above it) -- no cycle, row 0 keeps all its edges
			const uint32_t A0 = Law & 255u, A1 = (Law >> 8) & 255u, A2 = (Law >> 16) & 255u, A3 = Law >> 24;
			const uint32_t A4 = __shfl_down_sync(FULL, Law, 1) & 255u;
			h0 = h0 && !(A0 == L0 && A1 == L1);
			h1 = h1 && !(A1 == L1 && A2 == L2);
			h2 = h2 && !(A2 == L2 && A3 == L3);
			h3 = h3 && !(A3 == L3 && A4 == L4);
		}
		const bool v0 = (L0 != 255u) && (B0 != 255u) && (first || Lm != L0 || Bm != B0);
		const bool v1 = (L1 != 255u) && (B1 != 255u) && (L0 != L1 || B0 != B1);
		const bool v2 = (L2 != 255u) && (B2 != 255u) && (L1 != L2 || B1 != B2);
		const bool v3 = (L3 != 255u) && (B3 != 255u) && (L2 != L3 || B2 != B3);
		const uint32_t n = (uint32_t)h0 + (uint32_t)h1 + (uint32_t)h2 + (uint32_t)h3 + (uint32_t)v0 + (uint32_t)v1 + (uint32_t)v2 + (uint32_t)v3;
		uint32_t s = n;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(FULL, s, d); if (lane >= d) s += t; }
		uint32_t base = 0;
		if (lane == 31 && s) base = atomicAdd(&s_nlinks, s);
		base = __shfl_sync(FULL, base, 31);
		uint32_t la = links_s + (base + s - n) * 8u;
		if (h0) { sts_v2(la, ka0, ka1); la += 8u; }
		if (v0) { sts_v2(la, ka0, kb0); la += 8u; }
		if (h1) { sts_v2(la, ka1, ka2); la += 8u; }
		if (v1) { sts_v2(la, ka1, kb1); la += 8u; }
		if (h2) { sts_v2(la, ka2, ka3); la += 8u; }
		if (v2) { sts_v2(la, ka2, kb2); la += 8u; }
		if (h3) { sts_v2(la, ka3, ka4); la += 8u; }
		if (v3) { sts_v2(la, ka3, kb3); }
	}
	__syncthreads();
	ERT_PHASE(3);

	// ---- phase B2: drain the edge list.  Every lane owns one edge at a time, all lanes advance one step per iteration
	// (warp-converged state machine), idle lanes refill from the warp's contiguous share of the list.  The list is drained
	// from its END (bottom of the tile first): the level root of a node is its last pixel in raster order, so it is met
	// first and stays put while the rows above attach to it -- chains stay ~1 hop deep. ----
	{
		const uint32_t nl = s_nlinks;
		const uint32_t per_warp = (nl + NWARP - 1) / NWARP;
		uint32_t wnext = (uint32_t)warp * per_warp;
		const uint32_t wend = min(nl, wnext + per_warp);
		// state per lane: the edge (a, b); a == b means the lane is idle.  The step is branch-free (selects and predicated
		// stores / atomics): idle lanes and lanes that are already at a root re-read harmlessly.
		uint32_t a = 0, b = 0;
		int guard = 0;
		for (; guard < (1 << 22); ++guard) {
			const uint32_t idle = __ballot_sync(FULL, a == b);
			if (wnext < wend) {
				if (a == b) {
					const uint32_t i = wnext + (uint32_t)__popc(idle & lt);
					if (i < wend) {
						const uint2 e = lds_v2(links_s + (nl - 1u - i) * 8u);
						a = e.x; b = e.y;
					}
				}
				wnext += (uint32_t)__popc(idle);
			} else if (idle == FULL) break;
			const bool act = a != b;
			const uint32_t aa = par_s + ((a & 0xFFFFu) << 2), ab = par_s + ((b & 0xFFFFu) << 2);
			const uint32_t pa = lds_u32(aa), pb = lds_u32(ab);
			const bool ra = lvl_differs(pa, a), rb = lvl_differs(pb, b);
			// second hop: the grandparent when the parent is in the same level (path halving), else a harmless re-read
			const uint32_t ga = lds_u32(par_s + (((ra ? a : pa) & 0xFFFFu) << 2));
			const uint32_t gb = lds_u32(par_s + (((rb ? b : pb) & 0xFFFFu) << 2));
			const bool ha = !ra && !lvl_differs(ga, a), hb = !rb && !lvl_differs(gb, b);   // halve: grandparent still in the level
			sts_u32_if(act && ha, aa, ga);
			sts_u32_if(act && hb, ab, gb);
			// a root is reached when the own pointer leaves the level (ra) or the parent's does (!ha: pa IS the level root)
			const uint32_t a1 = ra ? a : (ha ? ga : pa), b1 = rb ? b : (hb ? gb : pb);
			const bool both = !ha && !hb;
			const uint32_t lo = min(a1, b1), hi = max(a1, b1);
			const bool link = act && both && (lo != hi);
			const uint32_t old = atoms_min_if(link, par_s + ((lo & 0xFFFFu) << 2), hi, KEY_NONE);
			// link: done if nothing was displaced (old == NONE) or the link existed (old == hi); else continue with (hi, old):
			// old < hi: lo already had a closer ancestor; old > hi: hi slipped in below it and old must go above hi
			const bool more = link && (old != hi) && (old != KEY_NONE);
			a = both ? (more ? min(hi, old) : 0u) : a1;
			b = both ? (more ? max(hi, old) : 0u) : b1;
		}
		if (guard >= (1 << 22)) atomicOr(status, ERR_LOOP_GUARD);
	}
	__syncthreads();
	ERT_PHASE(4);

	// ---- phase Z: the attribute words (they aliased the edge list) ----
	for (int i = tid; i < TPX / 4; i += NT) {
		sts_v4(attr_s + (uint32_t)i * 16u, 0u, 0u, 0u, 0u);
		sts_v4(attr_s + (uint32_t)(TPX * 4 + i * 16), 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
		sts_v4(attr_s + (uint32_t)(2 * TPX * 4 + i * 16), 0u, 0u, 0u, 0u);
		sts_v4(attr_s + (uint32_t)(3 * TPX * 4 + i * 16), 0u, 0u, 0u, 0u);
	}
	__syncthreads();

	// ---- phase D: own-level pixel count and bbox per tile-local node, one update per RUN (run list: one run per lane);
	// the run's last pixel finds its level root with a warp-converged walk; a run end that is its own root is a node.
	// acc word: bits 0..14 pixels, bits 15..29 nodes, bit 31 = node touches a seam (BORDER) ----
	constexpr uint32_t ACC_NODE = 1u << 15, ACC_MASK = 0x7FFFu, ACC_BORDER = 0x80000000u;
#pragma unroll
	for (int pass = 0; pass < 2; ++pass) {
		const int rp = warp + NWARP * pass;
		const uint32_t seg = runs_s + (uint32_t)(rp * 128) * 2u;
		const uint32_t n = nruns[pass];
		for (uint32_t j0 = 0; j0 < n; j0 += 32) {
			const uint32_t j = j0 + (uint32_t)lane;
			const bool valid = j < n;
			const uint32_t rec = valid ? lds_u16(seg + j * 2u) : REC_WALL;
			const uint32_t prev = (valid && j > 0) ? lds_u16(seg + j * 2u - 2u) : 0xFFFFu;
			const uint32_t e = rec & REC_IDX;
			const uint32_t y = e >> 6, xe = e & 63u;
			const uint32_t xs = (prev != 0xFFFFu && ((prev & REC_IDX) >> 6) == y) ? (prev & 63u) + 1u : 0u;
			const bool live = valid && !(rec & REC_WALL);
			const uint32_t L = lds_u8(box_s + y * BOXW + xe + LV0);
			uint32_t k = (L << 16) | e;
			bool act = live;
			while (__any_sync(FULL, act)) {
				if (act) {
					const uint32_t q = lds_u32(par_s + ((k & 0xFFFFu) << 2));
					if (lvl_differs(q, k)) act = false;
					else {
						const uint32_t g = lds_u32(par_s + ((q & 0xFFFFu) << 2));   // two hops per round; q is the root if its pointer leaves the level
						if (lvl_differs(g, k)) { k = q; act = false; }
						else k = g;
					}
				}
			}
			const uint32_t r = k & 0xFFFFu;
			const bool isroot = live && (r == e);
			if (live) {
				const uint32_t ar = attr_s + (r << 2);
				reds_add(ar, xe - xs + 1u + (isroot ? ACC_NODE : 0u));
				reds_min(ar + TPX * 4, xs);
				reds_max(ar + 2 * TPX * 4, xe);
				reds_or(ar + 3 * TPX * 4, 1u << y);
			}
			const uint32_t rmask = __ballot_sync(FULL, isroot);
			if (rmask) {
				uint32_t wb = 0;
				if (lane == 0) wb = atomicAdd(&s_nroots, (uint32_t)__popc(rmask));
				wb = __shfl_sync(FULL, wb, 0);
				if (isroot) rootlist[wb + __popc(rmask & lt)] = (uint16_t)e;
			}
		}
	}
	__syncthreads();
	ERT_PHASE(9);

	// ---- phase C: every level root points at its parent's LEVEL ROOT (roots only: a few % of the pixels) ----
	if (OPT & 1) for (int i = tid; i < TPX / 8; i += NT) sts_v4(runs_s + (uint32_t)i * 16u, 0u, 0u, 0u, 0u);   // arrival counters of phase D3 (the run records are dead)
	{
		const uint32_t nr = s_nroots;
		for (uint32_t i0 = warp * 32; i0 < nr; i0 += NT) {
			const uint32_t i = i0 + lane;
			uint32_t p = 0, k = KEY_NONE;
			if (i < nr) { p = rootlist[i]; k = lds_u32(par_s + (p << 2)); }
			const uint32_t k0 = k;
			bool act = (k != KEY_NONE);
			while (__any_sync(FULL, act)) {
				if (act) {
					const uint32_t q = lds_u32(par_s + ((k & 0xFFFFu) << 2));
					if (lvl_differs(q, k)) act = false;
					else k = q;
				}
			}
			if (k != k0) sts_u32(par_s + (p << 2), k);   // readers that still see k0 walk the same chain to the same root
		}
	}
	__syncthreads();
	ERT_PHASE(5);

	// ---- phase D2: which tile-local nodes can still change?  A pixel p on a side of the tile that faces another tile
	// meets its outside neighbour q (the halo ring of the box) at level M = max(level p, level q): what that edge can
	// change is the tile-local component holding p at threshold M -- the HIGHEST ancestor-or-self A(p) of p's node with
	// level <= M -- and everything above it.  A wall or the plane's border outside makes no edge at all.  BORDER = the
	// A(p) of all side pixels (plus the nodes of the flood's start candidates, pixels 0 / 1 / W of the plane) and all
	// their ancestors; everything else is INTERIOR: its subtree is final here and never has to leave the SM.  A(p) also
	// stands for p in the seam record (phase E). ----
	for (int i = tid; i < RING + 3; i += NT) {
		int x, y, dx = 0, dy = 0;
		if (i < TW) { x = i; y = 0; dy = -1; }
		else if (i < 2 * TW) { x = i - TW; y = rows - 1; dy = 1; }
		else if (i < 2 * TW + TH) { x = 0; y = i - 2 * TW; dx = -1; }
		else if (i < RING) { x = cols - 1; y = i - 2 * TW - TH; dx = 1; }
		else {   // start candidates of the flood: global pixels 0, 1, W
			const int gi = i - RING;
			x = ((gi == 1) ? 1 : 0) - X0; y = ((gi == 2) ? 1 : 0) - Y0;
		}
		if (x < 0 || y < 0 || x >= cols || y >= rows) continue;
		const int p = y * TW + x;
		const uint32_t L = ERT_LV(p);
		if (L == 255u) continue;
		uint32_t M = L;                                 // start candidates: the pixel's own node
		if (i < RING) {
			const uint32_t Lq = box[(y + YOFF + dy) * BOXW + XOFF + x + dx];
			if (Lq == 255u) continue;                   // a wall or nothing outside: no edge across the seam here
			M = max(L, Lq);
		}
		uint32_t kk = (L << 16) | (uint32_t)p;
		for (int guard = 0; guard < 65536; ++guard) {   // to the level root
			const uint32_t q = par[kk & 0xFFFFu];
			if (lvl_differs(q, kk)) break;
			kk = q;
		}
		uint32_t r = kk & 0xFFFFu;
		for (int guard = 0; guard < 64; ++guard) {      // level roots point at their parent's level root (phase C): climb while level <= M
			const uint32_t up = par[r];
			if (up == KEY_NONE || (up >> 16) > M) break;
			r = up & 0xFFFFu;
		}
		if (i < RING) s_ringA[i] = (uint16_t)r;
		for (int guard = 0; guard < 64; ++guard) {
			const uint32_t old = atomicOr(&cnt[r], ACC_BORDER);
			if (old & ACC_BORDER) break;
			const uint32_t up = par[r];
			if (up == KEY_NONE) break;
			r = up & 0xFFFFu;
		}
	}
	__syncthreads();
	ERT_PHASE(6);

	// ---- phase D3: fold interior subtrees bottom-up (pixels, node count, bbox travel to the parent) ----
	const uint32_t nroots = s_nroots;
	if (OPT & 1) {
		// arrival counters (u16 pairs in the run-list words, zeroed during phase C): an interior node with no interior child
		// starts; a thread carries a finished node into its parent and continues upward only if it was the last child to
		// arrive -- three CTA barriers instead of one per level.  BORDER is upward closed, so interior nodes have interior children only.
		uint32_t *pend = reinterpret_cast<uint32_t *>(sm + OFF_RUNS);
		for (uint32_t i = tid; i < nroots; i += NT) {
			const uint32_t p = rootlist[i];
			if (cnt[p] & ACC_BORDER) continue;
			const uint32_t up = par[p];
			if (up == KEY_NONE) continue;
			const uint32_t q = up & 0xFFFFu;
			atomicAdd(&pend[q >> 1], 1u << ((q & 1u) * 16u));
		}
		__syncthreads();
		// who starts is decided BEFORE any counter moves (a node whose counter reaches zero later is carried by its last child)
		uint32_t starts = 0;
		static_assert(TPX / NT <= 32, "one start bit per root handled by a thread");
		for (uint32_t i = tid, k = 0; i < nroots; i += NT, ++k) {
			const uint32_t cur = rootlist[i];
			if (!(cnt[cur] & ACC_BORDER) && !((pend[cur >> 1] >> ((cur & 1u) * 16u)) & 0xFFFFu)) starts |= 1u << k;
		}
		__syncthreads();
		for (uint32_t i = tid, k = 0; i < nroots; i += NT, ++k) {
			if (!((starts >> k) & 1u)) continue;
			uint32_t cur = rootlist[i];
			for (int guard = 0; guard < 64; ++guard) {
				const uint32_t up = par[cur];
				if (up == KEY_NONE) break;
				const uint32_t q = up & 0xFFFFu;
				const uint32_t acc = *reinterpret_cast<volatile uint32_t *>(&cnt[cur]);
				atomicAdd(&cnt[q], acc);
				atomicMin(&xmn[q], *reinterpret_cast<volatile uint32_t *>(&xmn[cur]));
				atomicMax(&xmx[q], *reinterpret_cast<volatile uint32_t *>(&xmx[cur]));
				atomicOr(&ymask[q], *reinterpret_cast<volatile uint32_t *>(&ymask[cur]));
				__threadfence_block();
				const uint32_t sh = (q & 1u) * 16u;
				const uint32_t old = atomicSub(&pend[q >> 1], 1u << sh);
				if (((old >> sh) & 0xFFFFu) != 1u) break;                          // not the last child
				if (*reinterpret_cast<volatile uint32_t *>(&cnt[q]) & ACC_BORDER) break;   // BORDER nodes keep what they gathered and go global
				__threadfence_block();
				cur = q;
			}
		}
		__syncthreads();
	} else {
		const int lo = (int)s_minlvl, hi_l = (int)s_maxlvl;
		for (int Lc = lo; Lc < hi_l; ++Lc) {
			for (uint32_t i = tid; i < nroots; i += NT) {
				const uint32_t p = rootlist[i];
				if (ERT_LV(p) != (uint32_t)Lc) continue;
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

	// ---- phase E: emit.  BORDER nodes go to the global forest with what they have gathered (own pixels + interior
	// descendants); interior nodes are emitted only if the reference would keep them (area > MIN_AREA), already
	// complete (pend = NODE_COMPLETE).  Every emitted node gets a SLOT in the plane's dense node arrays; the parent of
	// an emitted node is always emitted too (BORDER is upward closed, areas grow towards the root), so parent pointers
	// and seam records name slots.  Three steps: reserve the tile's slot range, write the records (and remember every
	// node's slot in its cnt word), then resolve parent slots / seam records / start candidates. ----
	constexpr uint32_t SLOT_SET = 0x80000000u;
	uint32_t my_emit = 0;
	for (uint32_t i = tid; i < nroots; i += NT) {
		const uint32_t acc = cnt[rootlist[i]];
		const bool emit = (acc & ACC_BORDER) || (int)((acc & ACC_MASK) + ((acc >> 15) & ACC_MASK)) > P.min_area;
		my_emit += emit ? 1u : 0u;
	}
	my_emit = __reduce_add_sync(FULL, my_emit);
	if (lane == 0 && my_emit) atomicAdd(&s_nemit, my_emit);
	__syncthreads();
	if (tid == 0) {
		s_base = s_nemit ? atomicAdd(&node_count[plane], s_nemit) : 0u;
		s_cursor = 0;
		if (s_base + s_nemit > (uint32_t)P.node_cap) atomicOr(status, ERR_NODE_OVERFLOW);
	}
	__syncthreads();
	const bool fits = s_base + s_nemit <= (uint32_t)P.node_cap;     // an overflowing tile publishes nothing (the batch is flagged)
	uint32_t *parP = par_g + (size_t)plane * P.node_cap;
	NodeAttr *attrP = attr_g + (size_t)plane * P.node_cap;
	uint32_t *keyP = node_key + (size_t)plane * P.node_cap;
	for (uint32_t i0 = warp * 32; i0 < nroots; i0 += NT) {
		const uint32_t i = i0 + lane;
		bool emit = false;
		uint32_t p = 0, acc = 0;
		if (i < nroots) {
			p = rootlist[i];
			acc = cnt[p];
			emit = (acc & ACC_BORDER) || (int)((acc & ACC_MASK) + ((acc >> 15) & ACC_MASK)) > P.min_area;
		}
		const uint32_t emask = __ballot_sync(FULL, emit);
		uint32_t wbase = 0;
		if (lane == 0 && emask) wbase = atomicAdd(&s_cursor, (uint32_t)__popc(emask));
		wbase = __shfl_sync(FULL, wbase, 0);
		if (i < nroots) cnt[p] = 0;                       // no slot (only this thread reads / writes this root's cnt word here)
		if (!emit || !fits) continue;
		const uint32_t slot = s_base + wbase + (uint32_t)__popc(emask & lt);
		const int y = (int)p / TW, x = (int)p % TW;
		const uint32_t gidx = (uint32_t)(Y0 + y) * (uint32_t)P.W + (uint32_t)(X0 + x);
		uint4 *a = reinterpret_cast<uint4 *>(&attrP[slot]);
		a[0] = make_uint4(acc & ACC_MASK, (acc >> 15) & ACC_MASK, (acc & ACC_BORDER) ? 0u : NODE_COMPLETE, 0u);
		const uint32_t ym = ymask[p];
		a[1] = make_uint4((uint32_t)X0 + xmn[p], (uint32_t)Y0 + (uint32_t)(__ffs(ym) - 1), (uint32_t)X0 + xmx[p], (uint32_t)Y0 + (uint32_t)(31 - __clz(ym)));
		keyP[slot] = make_key(ERT_LV(p), gidx);
		cnt[p] = SLOT_SET | slot;
	}
	__syncthreads();
	for (uint32_t i = tid; i < nroots; i += NT) {
		const uint32_t p = rootlist[i];
		const uint32_t sp = cnt[p];
		if (!(sp & SLOT_SET)) continue;
		const uint32_t pk = par[p];
		uint32_t gpar = KEY_NONE;
		if (pk != KEY_NONE) gpar = make_key(pk >> 16, cnt[pk & 0xFFFFu] & ~SLOT_SET);
		parP[sp & ~SLOT_SET] = gpar;
	}
	{
		// seam records: for every position on the four tile sides the GLOBAL key (level, slot) of the node that stands
		// for the pixel there (KEY_NONE where no edge crosses), laid out contiguously per tile so that the seam kernel
		// reads both sides of a seam coalesced and starts every union at a root
		uint32_t *rec = ring_rec + ((size_t)plane * gridDim.x + blockIdx.x) * RING;
		for (int i = tid; i < RING + 3; i += NT) {
			if (i < RING) {
				const uint32_t a = s_ringA[i];
				rec[i] = (a == 0xFFFFu || !fits) ? KEY_NONE : make_key(ERT_LV(a), cnt[a] & ~SLOT_SET);
				continue;
			}
			// the flood's start candidates (global pixels 0, 1, W; all inside tile 0): the key of the node that holds each of
			// them (phase D2 made those nodes BORDER, so they have a slot); KEY_NONE for a wall or a pixel that does not exist
			if (blockIdx.x != 0) continue;
			const int gi = i - RING;
			const int x = (gi == 1) ? 1 : 0, y = (gi == 2) ? 1 : 0;
			uint32_t key = KEY_NONE;
			if (x < cols && y < rows && fits) {
				const int p = y * TW + x;
				const uint32_t L = ERT_LV(p);
				if (L != 255u) {
					uint32_t kk = (L << 16) | (uint32_t)p;
					for (int guard = 0; guard < 65536; ++guard) {
						const uint32_t q2 = par[kk & 0xFFFFu];
						if (lvl_differs(q2, kk)) break;
						kk = q2;
					}
					key = make_key(L, cnt[kk & 0xFFFFu] & ~SLOT_SET);
				}
			}
			start_key[(size_t)plane * 4 + gi] = key;
		}
	}
	ERT_PHASE(8);
#undef ERT_PHASE
#undef ERT_LV
}

// ---------------------------------------------------------------------------------------------
// host side: the tensor map over the source planes (x = W bytes, y = H rows, z = planes) and the launcher
// ---------------------------------------------------------------------------------------------
int make_tile_tensor_map(TileTensorMap *out, const uint8_t *d_planes0, int W, int H, int pitch, int n_src_planes)
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
	static_assert(sizeof(TileTensorMap) == sizeof(CUtensorMap), "tensor map size");
	const cuuint64_t gdim[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)n_src_planes};
	const cuuint64_t gstride[2] = {(cuuint64_t)pitch, (cuuint64_t)pitch * (cuuint64_t)H};
	const cuuint32_t box[3] = {(cuuint32_t)t2::BOXW, (cuuint32_t)t2::BOXH, 1u};
	const cuuint32_t estr[3] = {1u, 1u, 1u};
	const CUresult r = encode(reinterpret_cast<CUtensorMap *>(out), CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t *>(d_planes0), gdim, gstride, box, estr,
	                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d) for %dx%d pitch %d planes %d", (int)r, W, H, pitch, n_src_planes); return -1; }
	return 0;
}

template <int OPT, int MINB>
static int launch_tile_v2_opt(const ExtractParams &P, const PlaneSrc *d_planes, ExtractWork &wk, cudaStream_t st)
{
	ERT_CUDA_CHECK(cudaFuncSetAttribute(k_tile_build2<OPT, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, t2::SMEM_BYTES));
	const int tiles_x = (P.W + t2::TW - 1) / t2::TW, tiles_y = (P.H + t2::TH - 1) / t2::TH;
	dim3 grid(tiles_x * tiles_y, P.n_planes);
	CUtensorMap tm;
	memcpy(&tm, &wk.tmap, sizeof tm);
	k_tile_build2<OPT, MINB><<<grid, t2::NT, t2::SMEM_BYTES, st>>>(tm, P, d_planes, wk.par, wk.attr, wk.node_key, wk.node_count, wk.start_key, wk.status, tiles_x, wk.prof, wk.ring_rec);
	ERT_CUDA_CHECK(cudaGetLastError());
	return 0;
}

int launch_tile_v2(const ExtractParams &P, const PlaneSrc *d_planes, ExtractWork &wk, cudaStream_t st, int opt)
{
	switch (opt) {
	case 0: return launch_tile_v2_opt<0, 4>(P, d_planes, wk, st);
	case 1: return launch_tile_v2_opt<1, 4>(P, d_planes, wk, st);
	case 2: return launch_tile_v2_opt<2, 4>(P, d_planes, wk, st);
	case 4: return launch_tile_v2_opt<3, 5>(P, d_planes, wk, st);
	default: return launch_tile_v2_opt<3, 4>(P, d_planes, wk, st);
	}
}

} // namespace ert
