"""Multi-GPU plumbing: one process per GPU (torch.distributed; NCCL over NVLink/NVSwitch on the
B200 box, gloo in CPU tests).  The path shards by *clip*: clips are independent units
(SURVEY.md section 8e), every rank runs K1 -> VS -> VD, A, alignment and K4 on its own clips, and the
only collective is one all-gather of the small per-frame tensors (labels and/or probabilities).
The reference has no distributed code at all (single process, literal "cuda:0": src/run.py:253).
"""
from __future__ import annotations

import os
from typing import List, Sequence

import torch
import torch.distributed as dist


def init_from_env(backend: str | None = None) -> tuple[int, int, int]:
    """Returns (rank, world, local_rank); initialises the default process group when WORLD_SIZE > 1."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local_rank


def shard_clips(costs: Sequence[float], world: int) -> List[List[int]]:
    """Longest-processing-time-first assignment of clips to ranks; cost ~ N frames + alpha * windows.
    Deterministic: ties go to the lowest rank; clip order inside a rank is ascending."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    load = [0.0] * world
    shards: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        shards[r].append(i)
        load[r] += costs[i]
    return [sorted(s) for s in shards]


def allgather_rows(local: torch.Tensor, counts: Sequence[int]) -> torch.Tensor:
    """All-gather row blocks of different heights: local [counts[rank], C] -> [sum(counts), C] on every
    rank (rank-major).  Blocks are padded to the tallest shard so a single all_gather_into_tensor moves
    everything (<= 88 B per frame: sub-millisecond on NVLink for 1.5 M frames)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    rank = dist.get_rank()
    assert local.shape[0] == counts[rank]
    tall = max(counts)
    cols = local.shape[1:]
    send = local.new_zeros((tall,) + tuple(cols))
    send[: counts[rank]] = local
    recv = local.new_empty((world * tall,) + tuple(cols))
    dist.all_gather_into_tensor(recv, send)
    return torch.cat([recv[r * tall: r * tall + counts[r]] for r in range(world)], dim=0)


def scatter_back(gathered: torch.Tensor, shards: Sequence[Sequence[int]], sizes: Sequence[int]) -> List[torch.Tensor]:
    """Undo the rank-major order of allgather_rows: returns one tensor per clip in original clip order."""
    out: List[torch.Tensor] = [None] * len(sizes)  # type: ignore[list-item]
    pos = 0
    for shard in shards:
        for i in shard:
            out[i] = gathered[pos: pos + sizes[i]]
            pos += sizes[i]
    return out
