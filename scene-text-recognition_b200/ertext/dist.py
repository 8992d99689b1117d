"""Data-parallel sharding + the final region gather.

Frames are independent units: rank r of G owns frames {f : f mod G == r}.  Nothing is exchanged on
the compute path; after a batch every rank contributes its region records (frame, plane, level,
area, x, y, w, h, label) to ONE all_gather (NCCL over NVLink on the GPU box, gloo in the CPU tests).
"""
import numpy as np
import torch
import torch.distributed as dist

REC_COLS = 9   # frame, plane, level, area, x, y, w, h, label


def shard_frames(n_frames, rank, world):
    return list(range(rank, n_frames, world))


def pack_records(batch_result, frame_ids):
    """BatchResult (6 planes per local frame) -> int32 [n, REC_COLS] of pooled regions with a label (vectorised)."""
    noff, nodes, poff, pool, label = batch_result.flat
    if len(pool) == 0:
        return np.zeros((0, REC_COLS), np.int32)
    plane = np.repeat(np.arange(len(poff) - 1), np.diff(poff))
    keep = label > 0
    plane, pool, label = plane[keep], pool[keep], label[keep]
    rows = nodes[noff[plane] + pool]
    fid = np.asarray(frame_ids, np.int32)[plane // 6]
    out = np.empty((len(plane), REC_COLS), np.int32)
    out[:, 0] = fid; out[:, 1] = plane % 6; out[:, 2:8] = rows[:, :6]; out[:, 8] = label
    return out


def gather_records(records, device, max_rows=None):
    """all_gather of variable-length record lists through torch.distributed (host-side reference of the library's gather,
    used by the gloo tests): the counts are gathered first and the payload is sized from the largest one, so nothing is
    ever dropped.  max_rows, if given, is a hard limit: exceeding it raises instead of truncating."""
    world = dist.get_world_size()
    n = len(records)
    if max_rows is not None and n > max_rows:
        raise ValueError("%d region records on this rank exceed max_rows = %d" % (n, max_rows))
    cnt = torch.tensor([n], dtype=torch.int32, device=device)
    cnts = torch.empty((world,), dtype=torch.int32, device=device)
    dist.all_gather_into_tensor(cnts, cnt)
    max_rows = int(cnts.max().item())
    buf = torch.zeros((max_rows + 1, REC_COLS), dtype=torch.int32)
    buf[0, 0] = n
    if n:
        buf[1:n + 1] = torch.from_numpy(records[:n])
    buf = buf.to(device, non_blocking=True)
    out = torch.empty((world, max_rows + 1, REC_COLS), dtype=torch.int32, device=device)
    dist.all_gather_into_tensor(out.view(-1, REC_COLS), buf)
    out = out.cpu().numpy()
    parts = [out[r, 1:1 + int(out[r, 0, 0])] for r in range(world)]
    allrec = np.concatenate(parts) if parts else np.zeros((0, REC_COLS), np.int32)
    order = np.lexsort(allrec.T[::-1]) if len(allrec) else []
    return allrec[order] if len(allrec) else allrec


class RegionGatherer:
    """Pipelined form of gather_records: every submit() starts one asynchronous all_gather on a rotating set of
    buffers and returns immediately; results become available `depth - 1` submits later (or at drain()).  The data
    path never waits for the collective -- it only has to be finished before its buffers are reused."""

    def __init__(self, device, max_rows=4096, depth=3):
        self.dev, self.max_rows, self.depth = device, max_rows, depth
        self.world = dist.get_world_size()
        pin = device.type == "cuda"
        self.host = [torch.zeros((max_rows + 1, REC_COLS), dtype=torch.int32, pin_memory=pin) for _ in range(depth)]
        self.src = [torch.zeros((max_rows + 1, REC_COLS), dtype=torch.int32, device=device) for _ in range(depth)]
        self.dst = [torch.empty((self.world * (max_rows + 1), REC_COLS), dtype=torch.int32, device=device) for _ in range(depth)]
        self.out_host = [torch.empty((self.world * (max_rows + 1), REC_COLS), dtype=torch.int32, pin_memory=pin) for _ in range(depth)]
        self.work = [None] * depth
        self.n = 0
        self.done = []          # gathered arrays (rank 0 keeps them; other ranks only count)

    def _finish(self, slot):
        w = self.work[slot]
        if w is None:
            return
        w.wait()
        if self.dev.type == "cuda":
            torch.cuda.current_stream(self.dev).synchronize()
        self.out_host[slot].copy_(self.dst[slot])
        out = self.out_host[slot].numpy().reshape(self.world, self.max_rows + 1, REC_COLS)
        parts = [out[r, 1:1 + int(out[r, 0, 0])] for r in range(self.world)]
        self.done.append(np.concatenate(parts) if parts else np.zeros((0, REC_COLS), np.int32))
        self.work[slot] = None

    def submit(self, records):
        slot = self.n % self.depth
        self._finish(slot)                       # the collective issued `depth` submits ago
        if len(records) > self.max_rows:
            raise ValueError("%d region records on this rank exceed max_rows = %d (nothing is dropped silently)" % (len(records), self.max_rows))
        n = len(records)
        h = self.host[slot]
        h[0, 0] = n
        if n:
            h[1:n + 1] = torch.from_numpy(records[:n])
        self.src[slot].copy_(h, non_blocking=True)
        self.work[slot] = dist.all_gather_into_tensor(self.dst[slot], self.src[slot], async_op=True)
        self.n += 1

    def drain(self):
        for k in range(self.depth):
            self._finish((self.n + k) % self.depth)
        res, self.done = self.done, []
        return res


# ---------------------------------------------------------------------------------------------
# the library's own gather (csrc/dist.cu): packed on the device, NCCL inside libertext.so, pipelined
# ---------------------------------------------------------------------------------------------
import ctypes as _C


class _Rec(_C.Structure):
    _fields_ = [(n, _C.c_int32) for n in ("frame", "plane", "level", "area", "x", "y", "w", "h", "label", "pool_index")]


class _GatherResult(_C.Structure):
    _fields_ = [("world", _C.c_int32), ("n_records", _C.c_int32), ("rank_offset", _C.POINTER(_C.c_int32)), ("records", _C.POINTER(_Rec)),
                ("sequence", _C.c_longlong)]


REC_DTYPE = np.dtype([(n, np.int32) for n in ("frame", "plane", "level", "area", "x", "y", "w", "h", "label", "pool_index")])


class LibraryGather:
    """ert_dist_* of include/ertext.h.  The NCCL unique id travels through torch.distributed (any backend); everything
    else -- device-side packing, count all-gather, exact-size grouped send / recv, pinned-host landing -- happens inside
    libertext.so on its own side stream."""

    def __init__(self, device_index, rank, world, max_records_per_rank=16384):
        import ertext
        self.L = ertext.load_library()
        L = self.L
        L.ert_dist_create.restype = _C.c_void_p
        L.ert_dist_create.argtypes = [_C.c_int, _C.c_int, _C.c_int, _C.c_void_p, _C.c_int]
        L.ert_dist_destroy.argtypes = [_C.c_void_p]
        L.ert_dist_unique_id.argtypes = [_C.c_void_p]
        L.ert_gather_regions_enqueue.argtypes = [_C.c_void_p, _C.c_void_p, _C.c_void_p, _C.c_int]
        L.ert_gather_regions_collect.argtypes = [_C.c_void_p, _C.POINTER(_C.POINTER(_GatherResult))]
        L.ert_gather_regions_outstanding.argtypes = [_C.c_void_p]
        idbuf = (_C.c_ubyte * 128)()
        if world > 1:
            if rank == 0 and L.ert_dist_unique_id(idbuf):
                raise ertext.ErtError(L.ert_last_error().decode())
            t = torch.tensor(list(idbuf), dtype=torch.uint8)
            if dist.get_backend() == "nccl":
                t = t.cuda(device_index)
            dist.broadcast(t, 0)
            idbuf = (_C.c_ubyte * 128)(*t.cpu().tolist())
        self.h = L.ert_dist_create(device_index, rank, world, idbuf if world > 1 else None, max_records_per_rank)
        if not self.h:
            raise ertext.ErtError(L.ert_last_error().decode())
        self.world = world

    def enqueue(self, ctx, frame_ids):
        import ertext
        ids = np.ascontiguousarray(frame_ids, dtype=np.int32)
        if self.L.ert_gather_regions_enqueue(self.h, ctx.ctx, ids.ctypes.data, len(ids)):
            raise ertext.ErtError(self.L.ert_last_error().decode())

    def outstanding(self):
        return self.L.ert_gather_regions_outstanding(self.h)

    def collect(self):
        """-> (records as a structured array copied out of the pinned buffer, rank offsets [world + 1], sequence number)"""
        import ertext
        rp = _C.POINTER(_GatherResult)()
        if self.L.ert_gather_regions_collect(self.h, _C.byref(rp)):
            raise ertext.ErtError(self.L.ert_last_error().decode())
        r = rp.contents
        n = r.n_records
        off = np.ctypeslib.as_array(r.rank_offset, shape=(r.world + 1,)).copy()
        rec = np.frombuffer((_C.c_char * (n * REC_DTYPE.itemsize)).from_address(_C.addressof(r.records.contents)), dtype=REC_DTYPE).copy() if n else np.zeros(0, REC_DTYPE)
        return rec, off, int(r.sequence)

    def close(self):
        if self.h:
            self.L.ert_dist_destroy(self.h)
            self.h = None
