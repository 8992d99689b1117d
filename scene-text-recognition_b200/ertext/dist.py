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
