"""CPU: replica sharding + per-step gather on a world_size-2 gloo group."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rlvacdiffsim_b200.sharding import allreduce_gradients, gather_step, pack_records, shard_indices


def test_round_robin_partition():
    parts = [shard_indices(11, r, 4) for r in range(4)]
    assert sorted(np.concatenate(parts).tolist()) == list(range(11))
    assert parts[1].tolist() == [1, 5, 9]


def test_gather_single_process():
    Q = [np.arange(12, dtype=np.float32), np.arange(5, dtype=np.float32)]
    q_out, acts = gather_step([0, 1], Q, [3, 4], n_replicas=2)
    assert np.array_equal(q_out[1], Q[1]) and acts.tolist() == [3, 4]
    assert pack_records([], [], [], 12, 3)[:, 0].tolist() == [-1.0, -1.0, -1.0]


def _worker(rank, world, port, n_rep, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard_indices(n_rep, rank, world)
    Q = [np.full(12 if r % 2 == 0 else 7, float(r), dtype=np.float32) + np.arange(12 if r % 2 == 0 else 7) for r in mine]
    chosen = [int(r) % 5 for r in mine]
    q_out, acts = gather_step(mine.tolist(), Q, chosen, n_replicas=n_rep)
    ok = all(len(q_out[r]) == (12 if r % 2 == 0 else 7) and q_out[r][0] == float(r) for r in range(n_rep))
    ok = ok and acts.tolist() == [r % 5 for r in range(n_rep)]
    p = torch.nn.Parameter(torch.zeros(5))
    p.grad = torch.full((5,), float(rank + 1))
    allreduce_gradients([p])
    ok = ok and torch.allclose(p.grad, torch.full((5,), 1.5))
    ret[rank] = ok
    dist.destroy_process_group()


def test_gather_two_ranks_gloo():
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 9, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret[0] and ret[1]
