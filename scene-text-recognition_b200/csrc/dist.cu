// dist.cu -- the one collective of the path: the final gather of region records.
//
// Frames are independent units (rank r of G owns frames {f : f mod G = r}); nothing is exchanged on the compute path.
// After a batch, every rank contributes the labelled regions of its frames.  The gather lives in the library, on the
// device, off the data path:
//   enqueue(batch s) : k_pack_regions compacts the batch's strong / weak regions into a send buffer ON THE DEVICE (the
//                      batch's own high-priority stream, right behind its result compaction), then on the gather's side
//                      stream: ncclAllGather of the per-rank counts -> pinned host;
//                      for batch s-4 (its counts reached the host long ago, nothing waits): ONE grouped exchange of EXACTLY
//                      the records each rank holds (ncclSend / ncclRecv per peer, ncclGroupStart / End), records -> pinned host
//   collect()        : the oldest outstanding gather: pointers into pinned host memory.
// No padding travels, nothing on the host touches the records, no stream of the data path is ever synchronised.
// NCCL is bound at run time (dlopen of the libnccl.so.2 the process already has, else the system one): libertext.so
// has no link-time NCCL dependency and single-GPU users never load it.
#include "ctx.h"
#include <dlfcn.h>
#include <nccl.h>
#include <algorithm>
#include <cstring>
#include <vector>

namespace ert {

// pool entries with a label (strong / weak) of every plane -> records {frame, plane, level, area, x, y, w, h, label, pool index}
__global__ void k_pack_regions(int n_planes, int node_cap, int pool_cap, const int32_t *__restrict__ counts, const OutNode *__restrict__ nodes,
                               const int32_t *__restrict__ pool, const int32_t *__restrict__ label, const int32_t *__restrict__ frame_ids,
                               ert_region_record *__restrict__ out, int out_cap, int32_t *__restrict__ out_count)
{
	const int plane = blockIdx.x;
	const int np = min(counts[2 * plane + 1], pool_cap);
	__shared__ int s_base, s_n;
	if (threadIdx.x == 0) s_n = 0;
	__syncthreads();
	// order inside a plane follows the pool order (classify's push order); planes are concatenated in arrival order and
	// sorted by (frame, plane) on the host side of the collect call only if the caller asks for it
	for (int i0 = 0; i0 < np; i0 += blockDim.x) {
		const int i = i0 + threadIdx.x;
		const size_t s = (size_t)plane * pool_cap + i;
		const bool keep = i < np && label[s] > 0;
		const unsigned m = __ballot_sync(0xFFFFFFFFu, keep);
		int wbase = 0;
		if ((threadIdx.x & 31) == 0 && m) wbase = atomicAdd(&s_n, __popc(m));
		wbase = __shfl_sync(0xFFFFFFFFu, wbase, 0);
		__syncthreads();
		if (threadIdx.x == 0) s_base = s_n ? atomicAdd(out_count, s_n) : 0;
		__syncthreads();
		if (keep) {
			const int pos = s_base + wbase + __popc(m & ((1u << (threadIdx.x & 31)) - 1u));
			if (pos < out_cap) {
				const OutNode nd = nodes[(size_t)plane * node_cap + pool[s]];
				ert_region_record r;
				r.frame = frame_ids ? frame_ids[plane / 6] : plane / 6; r.plane = plane % 6; r.level = nd.level; r.area = nd.area;
				r.x = nd.x; r.y = nd.y; r.w = nd.w; r.h = nd.h; r.label = label[s]; r.pool_index = i;
				out[pos] = r;
			}
		}
		__syncthreads();
		if (threadIdx.x == 0) s_n = 0;
		__syncthreads();
	}
}

struct NcclApi {
	void *h = nullptr;
	ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
	ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
	ncclResult_t (*CommInitRankConfig)(ncclComm_t *, int, ncclUniqueId, int, ncclConfig_t *) = nullptr;   // optional (NCCL >= 2.14)
	ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
	ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*GroupStart)() = nullptr;
	ncclResult_t (*GroupEnd)() = nullptr;
	const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi *nccl_api()
{
	static NcclApi api;
	static bool tried = false;
	if (tried) return api.h ? &api : nullptr;
	tried = true;
	// the copy the process already loaded (torch ships its own libnccl.so.2), else whatever the loader finds
	void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
	if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
	if (!h) { set_error("libnccl.so.2 cannot be loaded: %s", dlerror()); return nullptr; }
#define ERT_NCCL_SYM(field, name) *(void **)(&api.field) = dlsym(h, name); if (!api.field) { set_error("libnccl.so.2 lacks %s", name); return nullptr; }
	ERT_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
	ERT_NCCL_SYM(CommInitRank, "ncclCommInitRank")
	*(void **)(&api.CommInitRankConfig) = dlsym(h, "ncclCommInitRankConfig");
	ERT_NCCL_SYM(CommDestroy, "ncclCommDestroy")
	ERT_NCCL_SYM(AllGather, "ncclAllGather")
	ERT_NCCL_SYM(Send, "ncclSend")
	ERT_NCCL_SYM(Recv, "ncclRecv")
	ERT_NCCL_SYM(GroupStart, "ncclGroupStart")
	ERT_NCCL_SYM(GroupEnd, "ncclGroupEnd")
	ERT_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef ERT_NCCL_SYM
	api.h = h;
	return &api;
}

#define ERT_NCCL_CHECK(expr)                                                                                       \
	do {                                                                                                           \
		ncclResult_t _r = (expr);                                                                                  \
		if (_r != ncclSuccess) { set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, api->GetErrorString(_r)); return -1; } \
	} while (0)

} // namespace ert

using namespace ert;

namespace {
constexpr int DEPTH = 12;      // gathers in flight
constexpr int LAG = 6;         // the exact-size exchange of batch s - LAG is issued when batch s is enqueued (same order on every rank)
enum SlotState { FREE = 0, PACKED = 1, EXCHANGED = 2 };
struct Slot {
	int state = FREE;
	ert_region_record *d_send = nullptr, *d_recv = nullptr, *h_recv = nullptr;
	int32_t *d_count = nullptr, *d_counts_all = nullptr, *h_counts = nullptr;   // own count; all ranks' counts (device / pinned host)
	int32_t *d_frame_ids = nullptr;
	cudaEvent_t ev_packed = nullptr, ev_counts = nullptr, ev_data = nullptr;
	std::vector<int32_t> offsets;   // world + 1, records
	long long seq = -1;
};
} // namespace

struct ert_dist {
	int device = 0, rank = 0, world = 1;
	int cap = 0;                      // records per rank and batch
	ncclComm_t comm = nullptr;
	cudaStream_t stream = nullptr;    // the gather's own side stream
	Slot slot[DEPTH];
	long long n_enqueued = 0, n_collected = 0;
	ert_gather_result res{};
};

static int exchange_slot(ert_dist *d, Slot &s)
{
	NcclApi *api = d->world > 1 ? nccl_api() : nullptr;
	if (d->world > 1 && !api) return -1;
	ERT_CUDA_CHECK(cudaEventSynchronize(s.ev_counts));        // recorded a whole batch ago: does not wait in steady state
	s.offsets.assign((size_t)d->world + 1, 0);
	for (int r = 0; r < d->world; r++) {
		const int n = s.h_counts[r];
		if (n < 0 || n > d->cap) { set_error("rank %d reports %d region records, capacity is %d per batch (ert_dist_create)", r, n, d->cap); return -2; }
		s.offsets[(size_t)r + 1] = s.offsets[(size_t)r] + n;
	}
	const int mine = s.h_counts[d->rank];
	if (mine) ERT_CUDA_CHECK(cudaMemcpyAsync(s.d_recv + s.offsets[(size_t)d->rank], s.d_send, sizeof(ert_region_record) * (size_t)mine, cudaMemcpyDeviceToDevice, d->stream));
	if (d->world > 1) {
		const size_t words = sizeof(ert_region_record) / sizeof(int32_t);
		ERT_NCCL_CHECK(api->GroupStart());
		for (int p = 0; p < d->world; p++) {
			if (p == d->rank) continue;
			if (mine) ERT_NCCL_CHECK(api->Send(s.d_send, (size_t)mine * words, ncclInt32, p, d->comm, d->stream));
			const int theirs = s.h_counts[p];
			if (theirs) ERT_NCCL_CHECK(api->Recv(s.d_recv + s.offsets[(size_t)p], (size_t)theirs * words, ncclInt32, p, d->comm, d->stream));
		}
		ERT_NCCL_CHECK(api->GroupEnd());
	}
	const int total = s.offsets[(size_t)d->world];
	if (total) ERT_CUDA_CHECK(cudaMemcpyAsync(s.h_recv, s.d_recv, sizeof(ert_region_record) * (size_t)total, cudaMemcpyDeviceToHost, d->stream));
	ERT_CUDA_CHECK(cudaEventRecord(s.ev_data, d->stream));
	s.state = EXCHANGED;
	return 0;
}

extern "C" {

int ert_dist_unique_id(void *id128)
{
	NcclApi *api = nccl_api();
	if (!api || !id128) return -1;
	static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
	ncclUniqueId id;
	ERT_NCCL_CHECK(api->GetUniqueId(&id));
	memcpy(id128, &id, sizeof id);
	return 0;
}

ert_dist *ert_dist_create(int device, int rank, int world, const void *id128, int max_records_per_rank)
{
	if (world < 1 || rank < 0 || rank >= world || max_records_per_rank < 1 || (world > 1 && !id128)) { set_error("bad arguments"); return nullptr; }
	if (cudaSetDevice(device) != cudaSuccess) { set_error("cudaSetDevice(%d) failed", device); return nullptr; }
	ert_dist *d = new ert_dist();
	d->device = device; d->rank = rank; d->world = world; d->cap = max_records_per_rank;
	// highest priority: the NCCL kernels are tiny, but at default priority their CTAs queue behind every tile-build CTA already
	// pending on the device (several kernels of 49 k CTAs when the contexts' tile kernels are not chained) -- the counts then
	// arrive milliseconds late and the enqueue call that needs them blocks the host (measured at 2 GPUs: -9 % / -17 %)
	int prio_least = 0, prio_greatest = 0;
	cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest);
	bool ok = cudaStreamCreateWithPriority(&d->stream, cudaStreamNonBlocking, prio_greatest) == cudaSuccess;
	for (int i = 0; i < DEPTH && ok; i++) {
		Slot &s = d->slot[i];
		ok = ok && cudaMalloc((void **)&s.d_send, sizeof(ert_region_record) * (size_t)d->cap) == cudaSuccess;
		ok = ok && cudaMalloc((void **)&s.d_recv, sizeof(ert_region_record) * (size_t)d->cap * world) == cudaSuccess;
		ok = ok && cudaHostAlloc((void **)&s.h_recv, sizeof(ert_region_record) * (size_t)d->cap * world, cudaHostAllocDefault) == cudaSuccess;
		ok = ok && cudaMalloc((void **)&s.d_count, sizeof(int32_t)) == cudaSuccess;
		ok = ok && cudaMalloc((void **)&s.d_counts_all, sizeof(int32_t) * (size_t)world) == cudaSuccess;
		ok = ok && cudaHostAlloc((void **)&s.h_counts, sizeof(int32_t) * (size_t)world, cudaHostAllocDefault) == cudaSuccess;
		ok = ok && cudaMalloc((void **)&s.d_frame_ids, sizeof(int32_t) * 4096) == cudaSuccess;
		ok = ok && cudaEventCreateWithFlags(&s.ev_packed, cudaEventDisableTiming) == cudaSuccess;
		ok = ok && cudaEventCreateWithFlags(&s.ev_counts, cudaEventDisableTiming) == cudaSuccess;
		ok = ok && cudaEventCreateWithFlags(&s.ev_data, cudaEventDisableTiming) == cudaSuccess;
	}
	if (!ok) { set_error("ert_dist_create: allocation failed (%s)", cudaGetErrorString(cudaGetLastError())); ert_dist_destroy(d); return nullptr; }
	if (world > 1) {
		NcclApi *api = nccl_api();
		if (!api) { ert_dist_destroy(d); return nullptr; }
		ncclUniqueId id;
		memcpy(&id, id128, sizeof id);
		// the payload is a few KB per step: one CTA per NCCL kernel is plenty, and every CTA more is one that spins on an SM the
		// tile kernel wants while the peer is a step behind
		ncclResult_t r = ncclInternalError;
		if (api->CommInitRankConfig) {
			ncclConfig_t cfg = NCCL_CONFIG_INITIALIZER;
			cfg.minCTAs = 1; cfg.maxCTAs = 1;
			r = api->CommInitRankConfig(&d->comm, world, id, rank, &cfg);
			if (r != ncclSuccess) d->comm = nullptr;
		}
		if (r != ncclSuccess) r = api->CommInitRank(&d->comm, world, id, rank);
		if (r != ncclSuccess) { set_error("ncclCommInitRank: %s", api->GetErrorString(r)); d->comm = nullptr; ert_dist_destroy(d); return nullptr; }
	}
	return d;
}

void ert_dist_destroy(ert_dist *d)
{
	if (!d) return;
	cudaSetDevice(d->device);
	if (d->stream) cudaStreamSynchronize(d->stream);
	if (d->comm) { NcclApi *api = nccl_api(); if (api) api->CommDestroy(d->comm); }
	for (int i = 0; i < DEPTH; i++) {
		Slot &s = d->slot[i];
		cudaFree(s.d_send); cudaFree(s.d_recv); cudaFreeHost(s.h_recv); cudaFree(s.d_count); cudaFree(s.d_counts_all); cudaFreeHost(s.h_counts);
		cudaFree(s.d_frame_ids);
		if (s.ev_packed) cudaEventDestroy(s.ev_packed);
		if (s.ev_counts) cudaEventDestroy(s.ev_counts);
		if (s.ev_data) cudaEventDestroy(s.ev_data);
	}
	if (d->stream) cudaStreamDestroy(d->stream);
	delete d;
}

int ert_gather_regions_enqueue(ert_dist *d, ert_ctx *c, const int32_t *frame_ids, int n_frames)
{
	if (!d || !c || n_frames < 1 || n_frames > 4096) { set_error("bad arguments"); return -1; }
	if (!c->pending || c->pending_upto < ERT_STAGE_CLASSIFY || c->pending_planes != 6 * n_frames) {
		set_error("ert_gather_regions_enqueue: the context has no classified BGR batch of %d frames in flight", n_frames);
		return -1;
	}
	if (d->n_enqueued - d->n_collected >= DEPTH - 1) { set_error("ert_gather_regions_enqueue: %d gathers outstanding, collect first", DEPTH - 1); return -1; }
	ERT_CUDA_CHECK(cudaSetDevice(d->device));
	NcclApi *api = d->world > 1 ? nccl_api() : nullptr;
	if (d->world > 1 && !api) return -1;
	Slot &s = d->slot[d->n_enqueued % DEPTH];
	s.seq = d->n_enqueued;
	// (1) pack on the batch's own stream, right behind its result compaction (the device result buffers are still this batch's)
	cudaStream_t ws = c->work_stream();
	ERT_CUDA_CHECK(cudaMemsetAsync(s.d_count, 0, sizeof(int32_t), ws));
	if (frame_ids) ERT_CUDA_CHECK(cudaMemcpyAsync(s.d_frame_ids, frame_ids, sizeof(int32_t) * (size_t)n_frames, cudaMemcpyHostToDevice, ws));
	k_pack_regions<<<c->pending_planes, 256, 0, ws>>>(c->pending_planes, c->kept_cap, c->pool_cap, c->d_out_counts, c->d_out_nodes, c->d_out_pool, c->d_label,
	                                                 frame_ids ? s.d_frame_ids : nullptr, s.d_send, d->cap, s.d_count);
	ERT_CUDA_CHECK(cudaGetLastError());
	ERT_CUDA_CHECK(cudaEventRecord(s.ev_packed, ws));
	if (ws != c->stream) {       // keep the context's own stream ordered behind the pack (the next batch reuses the buffers it reads)
		ERT_CUDA_CHECK(cudaEventRecord(c->ev_post_done, ws));
		ERT_CUDA_CHECK(cudaStreamWaitEvent(c->stream, c->ev_post_done, 0));
	}
	c->launches += 1;
	// (2) counts of every rank, on the gather's side stream
	ERT_CUDA_CHECK(cudaStreamWaitEvent(d->stream, s.ev_packed, 0));
	if (d->world > 1) ERT_NCCL_CHECK(api->AllGather(s.d_count, s.d_counts_all, 1, ncclInt32, d->comm, d->stream));
	else ERT_CUDA_CHECK(cudaMemcpyAsync(s.d_counts_all, s.d_count, sizeof(int32_t), cudaMemcpyDeviceToDevice, d->stream));
	ERT_CUDA_CHECK(cudaMemcpyAsync(s.h_counts, s.d_counts_all, sizeof(int32_t) * (size_t)d->world, cudaMemcpyDeviceToHost, d->stream));
	ERT_CUDA_CHECK(cudaEventRecord(s.ev_counts, d->stream));
	s.state = PACKED;
	d->n_enqueued++;
	// (3) the exact-size exchange of the batch enqueued LAG calls ago: with a few batches in flight (the data path runs
	// several contexts round-robin) that batch has long finished, so its counts are on the host and nothing waits.  The lag
	// is a constant, not a poll: every rank issues its NCCL calls in the same order.
	if (d->n_enqueued > LAG) {
		Slot &p = d->slot[(d->n_enqueued - 1 - LAG) % DEPTH];
		if (p.state == PACKED) { const int rc = exchange_slot(d, p); if (rc) return rc; }
	}
	return 0;
}

int ert_gather_regions_collect(ert_dist *d, const ert_gather_result **out)
{
	if (!d || !out) { set_error("bad arguments"); return -1; }
	if (d->n_collected >= d->n_enqueued) { set_error("ert_gather_regions_collect: nothing outstanding"); return -1; }
	ERT_CUDA_CHECK(cudaSetDevice(d->device));
	Slot &s = d->slot[d->n_collected % DEPTH];
	if (s.state == PACKED) { const int rc = exchange_slot(d, s); if (rc) return rc; }     // draining: the exchange has not been issued yet
	ERT_CUDA_CHECK(cudaEventSynchronize(s.ev_data));
	d->res.world = d->world; d->res.rank_offset = s.offsets.data(); d->res.records = s.h_recv; d->res.n_records = s.offsets[(size_t)d->world];
	d->res.sequence = s.seq;
	s.state = FREE;
	d->n_collected++;
	*out = &d->res;
	return 0;
}

int ert_gather_regions_outstanding(ert_dist *d) { return d ? (int)(d->n_enqueued - d->n_collected) : -1; }

} // extern "C"
