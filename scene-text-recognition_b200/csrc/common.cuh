// common.cuh -- shared device/host definitions for the ER detect+classify path (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace ert {

// ---------------------------------------------------------------------------------------------
// Keys.  Inside a tile a pixel is named by key = level << 16 | tile-local pixel index; in the GLOBAL forest a
// tile-local node that leaves its tile is named by key = level << 26 | slot, slot = its position in the plane's node
// array (handed out by the tile kernel).  Ordering keys as unsigned integers orders nodes by (level, index); the
// component tree is held as a forest par[] in which par[i] is always the key of an ANCESTOR of i (strictly larger
// key), or KEY_NONE.  See DESIGN.md "keyed lock-free union-find".
// ---------------------------------------------------------------------------------------------
constexpr int      KEY_IDX_BITS = 26;
constexpr uint32_t KEY_IDX_MASK = (1u << KEY_IDX_BITS) - 1u;
constexpr uint32_t KEY_NONE     = 0xFFFFFFFFu;
constexpr int      MAX_LEVELS   = 63;           // level field is 6 bits
constexpr uint32_t NODE_COMPLETE = 0xFFFFFFFFu;  // NodeAttr::pend marker: interior node, totals final after the tile pass

__host__ __device__ __forceinline__ uint32_t make_key(uint32_t level, uint32_t idx) { return (level << KEY_IDX_BITS) | idx; }
__host__ __device__ __forceinline__ uint32_t key_level(uint32_t k) { return k >> KEY_IDX_BITS; }
__host__ __device__ __forceinline__ uint32_t key_idx(uint32_t k) { return k & KEY_IDX_MASK; }

// Per-node attributes, one 32-byte sector per node slot (dense: slots are handed out consecutively per plane).
struct __align__(32) NodeAttr {
	uint32_t cnt;    // pixels: own-level pixels after the tile pass, whole subtree after refit
	uint32_t nn;     // number of tree nodes in the subtree (1 = self); 0 marks an alias (merged across a seam)
	uint32_t pend;   // number of child nodes (constant during refit)
	uint32_t arr;    // children arrived during refit; afterwards: position in the kept list
	uint32_t x0, y0, x1, y1;   // inclusive bbox
};

// A kept node as handed from the extract stage to the NMS stage (unordered list per plane)
struct KeptRec {
	uint32_t gidx;     // pixel index of the node's level root
	uint32_t parent;   // pixel index of the parent's level root, KEY_NONE for the root
	int32_t level, area;
	uint16_t x0, y0, x1, y1;
};

// Output node record: same columns as the oracle's dump (level, area, x, y, w, h, parent, n_children)
struct OutNode { int32_t level, area, x, y, w, h, parent, nchild; };

struct PlaneSrc {
	const uint8_t *src;   // u8 plane, row pitch = pitch bytes
	int invert;           // 1: value = 255 - src
	int z;                // index of the source plane inside the context's plane buffer (third coordinate of the tile tensor map)
};

struct ExtractParams {
	int W, H, pitch;        // plane geometry (pitch in bytes, multiple of 16)
	int n_planes;
	int hi;                 // 255/step + 1 : levels >= hi are walls
	float qscale;           // (float)(1.0/step)
	int min_area;
	int kept_cap;           // capacity of the kept list per plane
	int node_cap;           // capacity of the global node arrays per plane (slots)
};

// device-side status flags (bit-or'ed)
enum : uint32_t {
	ERR_LOOP_GUARD   = 1u,   // a bounded loop hit its guard (indicates a bug)
	ERR_KEPT_OVERFLOW = 2u,
	ERR_POOL_OVERFLOW = 4u,
	ERR_NMS_OVERFLOW  = 8u,
	ERR_NODE_OVERFLOW = 16u, // more tile-local nodes left their tiles than the plane's node array holds (ert_set_node_capacity)
};

__device__ __forceinline__ int quantize_level(int v, float qscale)
{
	// cv::Mat /= step  ==  saturate_cast<uchar>(cvRound(v * (float)(1/step)))   (src/ER.cpp:250)
	int q = __float2int_rn((float)v * qscale);
	return q > 255 ? 255 : q;
}

// relaxed, L1-bypassing load for words that other CTAs update with atomics in the same kernel
__device__ __forceinline__ uint32_t ld_relaxed(const uint32_t *p)
{
	uint32_t v;
	asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}

#define ERT_CUDA_CHECK(expr)                                                                     \
	do {                                                                                         \
		cudaError_t _e = (expr);                                                                 \
		if (_e != cudaSuccess) {                                                                 \
			ert::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
			return -1;                                                                           \
		}                                                                                        \
	} while (0)

void set_error(const char *fmt, ...);

} // namespace ert
