"""Multi-GPU check of the library's region gather (run under torchrun, one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py
Every rank runs the path on its own frames, the labelled regions are gathered by ert_gather_regions_* (device pack +
NCCL inside libertext.so); the gathered records of every step must equal, on every rank, the union of what the ranks
pack on the host from their own results (exchanged through torch.distributed as the reference plumbing)."""
import os
import sys
import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scene-text-recognition_b200"))
import ertext  # noqa: E402
from ertext import dist as edist, synth  # noqa: E402


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    g = edist.LibraryGather(lr, rank, world, max_records_per_rank=4096)
    ctxs = [ertext.ErText(device=lr) for _ in range(2)]
    fpg, W, H = 2, 640, 480
    ok = True
    expected = []
    steps = 7
    for s in range(steps):
        ids = [rank + world * i + 100 * s for i in range(fpg)]
        frames = np.stack([synth.s_text_frame(1000 + i, W, H, n_glyphs=40 if (s + rank) % 3 else 0) for i in ids])
        c = ctxs[s % 2]
        c.enqueue_host_array(frames)
        g.enqueue(c, ids)
        r = c.fetch()
        mine = edist.pack_records(r, ids)
        allr = [None] * world
        dist.all_gather_object(allr, mine)
        expected.append(np.concatenate(allr))
        if g.outstanding() >= 9:
            rec, off, seq = g.collect()
            ok &= check(rec, off, expected[seq], world, rank, seq)
    while g.outstanding():
        rec, off, seq = g.collect()
        ok &= check(rec, off, expected[seq], world, rank, seq)
    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("dist_check world %d: %s (%d steps, %d records in the last gather)" % (world, "OK" if int(t.item()) else "FAILED", steps, len(expected[-1])))
    g.close()
    dist.destroy_process_group()
    sys.exit(0 if int(t.item()) else 1)


def check(rec, off, exp, world, rank, seq):
    got = sorted((int(r["frame"]), int(r["plane"]), int(r["level"]), int(r["area"]), int(r["x"]), int(r["y"]), int(r["w"]), int(r["h"]), int(r["label"])) for r in rec)
    want = sorted(tuple(int(v) for v in row) for row in exp)
    good = got == want and len(off) == world + 1 and off[-1] == len(rec)
    if not good:
        print("rank %d step %d: gathered %d records, expected %d" % (rank, seq, len(got), len(want)))
    return good


if __name__ == "__main__":
    main()
