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


def gather_records(records, device, max_rows=4096):
    """all_gather of variable-length record lists: one fixed-size buffer per rank (count in row 0)."""
    world = dist.get_world_size()
    buf = torch.zeros((max_rows + 1, REC_COLS), dtype=torch.int32)
    n = min(len(records), max_rows)
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
        n = min(len(records), self.max_rows)
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
