"""Replica sharding across the GPUs of one box.

Replicas are independent (the reference loops them serially, ``rlsim/drl/deploy.py:75-94``), so the data
path needs NO collective: rank r owns replicas ``r, r+W, r+2W, …`` and scores them on its own GPU.  The
only exchange is the per-step gather of ``(rl_q[≤A], selected action)`` per replica for the rank-0
trajectory writer — one ``all_gather`` of a fixed-width fp32 record over NCCL/NVLink (gloo in CPU tests).
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist


def shard_indices(n_items: int, rank: int, world_size: int) -> np.ndarray:
    """Round-robin ownership: item i belongs to rank i % world_size."""
    return np.arange(rank, n_items, world_size, dtype=np.int64)


def record_width(max_actions: int) -> int:
    return 3 + max_actions   # replica id, n_actions, chosen action, Q[max_actions]


def pack_records(replica_ids: Sequence[int], Q_per_replica: Sequence[np.ndarray], chosen: Sequence[int],
                 max_actions: int, rows: int) -> torch.Tensor:
    """Fixed-shape [rows, 3+max_actions] fp32 block; unused rows carry replica id -1."""
    rec = np.full((rows, record_width(max_actions)), np.nan, dtype=np.float32)
    rec[:, 0] = -1.0
    for k, (rid, q, a) in enumerate(zip(replica_ids, Q_per_replica, chosen)):
        q = np.asarray(q, dtype=np.float32)
        if len(q) > max_actions:
            raise ValueError(f"replica {rid} has {len(q)} candidates > max_actions={max_actions}")
        rec[k, 0], rec[k, 1], rec[k, 2] = float(rid), float(len(q)), float(a)
        rec[k, 3:3 + len(q)] = q
    return torch.from_numpy(rec)


def gather_step(replica_ids: Sequence[int], Q_per_replica: Sequence[np.ndarray], chosen: Sequence[int],
                n_replicas: int, max_actions: int = 12, device=None, group=None
                ) -> Tuple[List[np.ndarray], np.ndarray]:
    """All-gather one step's results; every rank returns (Q per replica, chosen action per replica) in
    global replica order.  Works without an initialised process group (world size 1)."""
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rows = (n_replicas + world - 1) // world
    local = pack_records(replica_ids, Q_per_replica, chosen, max_actions, rows)
    if world == 1:
        blocks = [local]
    else:
        if device is not None:
            local = local.to(device)
        blocks = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(blocks, local, group=group)
        blocks = [b.cpu() for b in blocks]
    Q_out: List[np.ndarray] = [np.zeros(0, dtype=np.float32) for _ in range(n_replicas)]
    act_out = np.full(n_replicas, -1, dtype=np.int64)
    for b in blocks:
        arr = b.numpy()
        for row in arr:
            rid = int(row[0])
            if rid < 0:
                continue
            n = int(row[1])
            Q_out[rid] = row[3:3 + n].copy()
            act_out[rid] = int(row[2])
    return Q_out, act_out


def allreduce_gradients(parameters, group=None) -> None:
    """Sum-all-reduce of DQN gradients for rl-train; one flat 2.46 MB bucket."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    grads = [p.grad for p in parameters if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat /= dist.get_world_size(group)
    off = 0
    for g in grads:
        g.copy_(flat[off:off + g.numel()].view_as(g))
        off += g.numel()


def gather_step_device(engine, first: torch.Tensor, rl_q: torch.Tensor, chosen: torch.Tensor, E_R: torch.Tensor, rank: int,
                       world: int, max_actions: int = 12, group=None) -> torch.Tensor:
    """The per-step exchange of SURVEY.md §8e, device side: every rank packs one fixed-width fp32 row per replica
    (``rlvd_pack_step_records``: replica id = rank + world * local index, n_actions, chosen, E_R, Q[max_actions]) and ONE
    ``all_gather_into_tensor`` over NCCL returns all rows to every rank — no numpy round trip, no per-row Python.
    → f32[world * R, 4 + max_actions] on the device (ranks must hold equally many replicas)."""
    local = engine.pack_step_records(first, rl_q, chosen, E_R, rank, world, max_actions)
    if world == 1 or not (dist.is_available() and dist.is_initialized()):
        return local
    out = torch.empty((world * local.shape[0], local.shape[1]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local, group=group)
    return out


def unpack_records(records) -> Tuple[List[np.ndarray], np.ndarray, np.ndarray]:
    """Host view of gathered rows: (Q per replica, chosen, E_R) in global replica order."""
    arr = records.cpu().numpy() if torch.is_tensor(records) else np.asarray(records)
    ids = arr[:, 0].astype(np.int64)
    order = np.argsort(ids, kind="stable")
    arr = arr[order]
    Q = [row[4:4 + int(row[1])].copy() for row in arr]
    return Q, arr[:, 2].astype(np.int64), arr[:, 3].copy()
