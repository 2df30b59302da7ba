"""Multi-GPU plumbing: utterances shard across ranks with no data-path collective.

The reference has no multi-GPU inference (svc_inference_batch.py:39-43 is a serial loop of
subprocesses).  Here one process per GPU holds a full copy of the model; rank 0 packs the
checkpoint once and its packed blob is broadcast over NCCL (NVLink/NVSwitch) — the only
collective besides the final timing reduction.  Work items are dealt longest-first,
round-robin, so ranks finish together.
"""
from __future__ import annotations

import os
from typing import List, Sequence

import torch
import torch.distributed as dist


def env_world():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def init(backend: str | None = None):
    rank, local, world = env_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, local, world


def assign(lengths: Sequence[int], world: int) -> List[List[int]]:
    """Deal item indices to ranks: sort by length (longest first), then round-robin in a
    boustrophedon so per-rank totals stay balanced.  Returns world lists of indices."""
    order = sorted(range(len(lengths)), key=lambda i: (-lengths[i], i))
    out: List[List[int]] = [[] for _ in range(world)]
    for n, i in enumerate(order):
        r = n % world
        if (n // world) % 2:
            r = world - 1 - r
        out[r].append(i)
    return out


def broadcast_blob(blob: torch.Tensor | None, table, device, src: int = 0):
    """Broadcast rank `src`'s packed weight blob + table to every rank (one collective)."""
    rank, _, world = env_world()
    if world == 1:
        return blob, table
    meta = [None]
    if rank == src:
        meta = [(int(blob.numel()), table)]
    dist.broadcast_object_list(meta, src=src)
    numel, table = meta[0]
    if rank != src:
        blob = torch.empty(numel, dtype=torch.float32, device=device)
    dist.broadcast(blob, src=src)
    return blob, table


def max_over_ranks(value: float, device) -> float:
    rank, _, world = env_world()
    if world == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device) -> float:
    rank, _, world = env_world()
    if world == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def barrier():
    if dist.is_initialized():
        dist.barrier()
