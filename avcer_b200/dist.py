"""Multi-GPU plumbing: one process per GPU (torch.distributed; NCCL over NVLink/NVSwitch on the
B200 box, gloo in CPU tests).  The path shards by *clip*: clips are independent units
(SURVEY.md section 8e), every rank runs K1 -> VS -> VD, A, alignment and K4 on its own clips, and the
only collective is one all-gather of the small per-frame tensors (labels and/or probabilities).
The reference has no distributed code at all (single process, literal "cuda:0": src/run.py:253).
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence

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


# Relative cost of one 4 s audio window against one face crop (108 us vs 9.3 us per unit at round-1 speeds): the
# longest-processing-time-first assignment balances frames + ALPHA * windows per rank.
COST_ALPHA = 11.6


def clip_costs(n_frames: Sequence[int], n_windows: Sequence[int], alpha: float = COST_ALPHA) -> List[float]:
    return [float(f) + alpha * float(w) for f, w in zip(n_frames, n_windows)]


class ShardedRunner:
    """Clip-sharded execution of the whole path on `world` ranks (BASELINE config 4).

    Every rank knows the METADATA of all clips (frame counts, which clips start with a missing crop) and owns the DATA of
    its shard only.  A step = the rank's clips through K1 -> VS -> VD, A and alignment, with the per-frame rows (VS
    probabilities [n,7], VD logits [n,7], audio mean logits [n,8]; 88 B per frame) written STRAIGHT into this rank's block of
    the send buffer; ONE all_gather_into_tensor (NCCL over NVLink on the box, gloo in the CPU tests); then every rank
    runs the tail (permute, softmax, K4) over each gathered block, K4 storing its labels directly into that block's slot
    of the [4, all frames] result (label pitch) -- no transposes, no zero-filled staging, no concatenation.
    The result is rank-major; `clip_slices` maps it back to clips."""

    ROW = 7 + 7 + 8          # floats per frame in the exchange buffer

    def __init__(self, engine, n_frames: Sequence[int], n_windows: Sequence[int], f64_flags: Sequence[bool],
                 rank: int, world: int, alpha: float = COST_ALPHA):
        self.engine, self.rank, self.world = engine, rank, world
        self.n_frames = [int(f) for f in n_frames]
        self.f64_flags = [bool(f) for f in f64_flags]
        self.shards = shard_clips(clip_costs(n_frames, n_windows, alpha), world)
        self.counts = [sum(self.n_frames[i] for i in sh) for sh in self.shards]
        self.tall = max(self.counts) if self.counts else 0
        self.total = sum(self.counts)
        self.offsets = [0]
        for c in self.counts:
            self.offsets.append(self.offsets[-1] + c)
        dev = engine.device if engine is not None else torch.device("cpu")
        self.ncls = engine.a.num_classes if engine is not None and engine.a is not None else 8
        self.send = torch.zeros(max(self.tall, 1) * self.ROW, dtype=torch.float32, device=dev)
        self.recv = torch.empty((world, max(self.tall, 1) * self.ROW), dtype=torch.float32, device=dev)
        self.labels = torch.empty((4, max(self.total, 1)), dtype=torch.int64, device=dev)
        self.clip_slices: Dict[int, slice] = {}
        for r, sh in enumerate(self.shards):
            pos = self.offsets[r]
            for i in sh:
                self.clip_slices[i] = slice(pos, pos + self.n_frames[i])
                pos += self.n_frames[i]

    @property
    def my_clips(self) -> List[int]:
        return self.shards[self.rank]

    def block_views(self, buf: torch.Tensor, n: int):
        """(stat [n,7], dyn [n,7], audio [n,ncls]) views of one rank's block of `tall * ROW` floats."""
        t = self.tall
        return (buf[: n * 7].view(n, 7), buf[7 * t: 7 * t + n * 7].view(n, 7),
                buf[14 * t: 14 * t + n * self.ncls].view(n, self.ncls))

    def exchange(self) -> torch.Tensor:
        """The one collective of the path."""
        if self.world == 1:
            self.recv[0].copy_(self.send)
            return self.recv
        dist.all_gather_into_tensor(self.recv.view(-1), self.send)
        return self.recv

    def fuse_gathered(self, weights_1, weights_2, ce_weights_type: bool, ce_mask: bool) -> torch.Tensor:
        """Tail over every gathered block; returns labels [4, total frames] (rank-major)."""
        eng = self.engine
        for r, sh in enumerate(self.shards):
            n = self.counts[r]
            if n == 0:
                continue
            stat, dyn, a = self.block_views(self.recv[r], n)
            eng.fuse_clips(stat, dyn, a, [self.f64_flags[i] for i in sh], [self.n_frames[i] for i in sh], weights_1, weights_2,
                           ce_weights_type, ce_mask, labels=self.labels[:, self.offsets[r]: self.offsets[r] + n])
        return self.labels[:, : self.total]

    def step(self, crops_u8: torch.Tensor, exists_list, fps_list, wav_cat: torch.Tensor, wav_lens, weights_1, weights_2,
             ce_weights_type: bool, ce_mask: bool, **kw) -> torch.Tensor:
        """One pass: this rank's clips (inputs in `my_clips` order) -> labels of ALL clips on every rank."""
        n = self.counts[self.rank]
        if n:
            self.engine.run_clips(crops_u8, exists_list, fps_list, wav_cat, wav_lens, None, None, False, False,
                                  rows_out=self.block_views(self.send, n), fuse=False, **kw)
        self.exchange()
        return self.fuse_gathered(weights_1, weights_2, ce_weights_type, ce_mask)
