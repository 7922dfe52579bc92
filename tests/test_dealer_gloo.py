"""CPU, world_size 2 over gloo: batches are dealt round-robin, processed independently per rank (no data-path collective)
and re-ordered by start_idx on rank 0."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    from bwaflow_b200 import dealer
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_batches, per = 7, 10
    mine = dealer.deal(n_batches, world, rank)
    # "process" a batch: a deterministic function of its reads (stands in for the GPU stages)
    results = [(b * per, per, (np.arange(b * per, (b + 1) * per) ** 2).tolist()) for b in mine]
    gathered = [None] * world
    dist.all_gather_object(gathered, results)              # result collection only; nothing is exchanged on the data path
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)               # the bench's max-over-ranks timing reduction
    if rank == 0:
        ro = dealer.Reorder()
        out = []
        for r in range(world - 1, -1, -1):                 # arrive out of order
            for start, n, payload in gathered[r]:
                for _, p in ro.push(start, n, payload):
                    out.extend(p)
        q.put((out, ro.drained(), t.item()))
    dist.destroy_process_group()


def test_round_robin_deal_and_reorder_world2():
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out, drained, tmax = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert drained and tmax == 2.0
    assert out == [i * i for i in range(70)]


def test_deal_covers_every_batch_once():
    sys.path.insert(0, ROOT)
    from bwaflow_b200 import dealer
    for world in (1, 2, 4, 8):
        seen = sorted(b for r in range(world) for b in dealer.deal(19, world, r))
        assert seen == list(range(19))
