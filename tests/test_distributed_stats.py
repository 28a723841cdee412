"""CPU, world_size 2 over gloo: the only cross-rank step (reward statistics) and the env-sharding
arithmetic used by bench.py / the multi-GPU path."""
from __future__ import annotations

import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sorrel_b200.stats import all_reduce_partial, all_reduce_partial_async, summarize

    total = 1000
    n = total // world
    rng = np.random.RandomState(0)
    allv = rng.randint(-40, 90, total).astype(np.float64)
    mine = allv[rank * n:(rank + 1) * n]  # shard owned by this rank: envs [rank*n, (rank+1)*n)
    part = torch.tensor([mine.sum(), (mine * mine).sum(), mine.min(), mine.max(), len(mine)], dtype=torch.float64)
    s = summarize(all_reduce_partial(part))
    ok = (s["count"] == total and s["min"] == allv.min() and s["max"] == allv.max()
          and abs(s["mean"] - allv.mean()) < 1e-12 and abs(s["std"] - allv.std()) < 1e-9)
    # the overlapped form (collective left in flight, joined later) gives the same numbers; several may be pending
    pend = [all_reduce_partial_async(part * k) for k in (1.0, 2.0)]
    s2 = summarize(pend[1].result())
    ok = ok and summarize(pend[0].result()) == s and s2["count"] == 2 * total and s2["max"] == 2 * allv.max()
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_reward_stats_all_reduce_gloo_world2():
    world, port = 2, _free_port()
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}
