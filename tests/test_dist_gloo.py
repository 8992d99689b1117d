"""CPU, world_size 2 over gloo: the data-parallel plumbing (frame sharding + the final region gather)."""
import os
import sys
import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from conftest import ROOT, PKG


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    for p in (ROOT, PKG):
        if p not in sys.path:
            sys.path.insert(0, p)
    from ertext import dist as edist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    frames = edist.shard_frames(7, rank, world)
    rng = np.random.RandomState(100)           # same stream on every rank: records keyed by frame id
    allrec = rng.randint(0, 1000, (40, edist.REC_COLS)).astype(np.int32)
    allrec[:, 0] = rng.randint(0, 7, 40)
    mine = allrec[np.isin(allrec[:, 0], frames)]
    got = edist.gather_records(mine, torch.device("cpu"), max_rows=64)
    exp = allrec[np.lexsort(allrec.T[::-1])]
    q.put((rank, frames, bool((got == exp).all() and got.shape == exp.shape)))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_and_gather_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(60)
    assert res[0][1] == [0, 2, 4, 6] and res[1][1] == [1, 3, 5]
    assert res[0][2] and res[1][2]


def _worker_async(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    for p in (ROOT, PKG):
        if p not in sys.path:
            sys.path.insert(0, p)
    from ertext import dist as edist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = edist.RegionGatherer(torch.device("cpu"), max_rows=32, depth=3)
    rng = np.random.RandomState(7)
    exp = []
    for step in range(7):                              # more submits than buffers: exercises reuse
        allrec = rng.randint(0, 1000, (20, edist.REC_COLS)).astype(np.int32)
        allrec[:, 0] = rng.randint(0, 6, 20)
        mine = allrec[allrec[:, 0] % world == rank]
        g.submit(mine)
        exp.append(np.concatenate([allrec[allrec[:, 0] % world == r] for r in range(world)]))
    got = g.drain()
    ok = len(got) == len(exp) and all(a.shape == b.shape and (a == b).all() for a, b in zip(got, exp))
    q.put((rank, ok))
    dist.barrier()
    dist.destroy_process_group()


def test_pipelined_gather_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker_async, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(60)
    assert res == [(0, True), (1, True)]
