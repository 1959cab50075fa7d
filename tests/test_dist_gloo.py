"""CPU: the N>1 path (clip sharding + all-gather of per-frame rows) with world_size 2 over gloo."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from avcer_b200 import dist as adist


def test_shard_clips_lpt_is_balanced_and_complete():
    rng = np.random.default_rng(0)
    costs = rng.integers(250, 15000, size=37).tolist()
    for world in (1, 2, 4, 8):
        shards = adist.shard_clips(costs, world)
        assert sorted(i for s in shards for i in s) == list(range(37))
        loads = [sum(costs[i] for i in s) for s in shards]
        assert max(loads) - min(loads) <= max(costs)


def _worker(rank, world, port, sizes, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    adist.init_from_env("gloo")
    shards = adist.shard_clips(sizes, world)
    # every rank "processes" its clips: row value encodes (clip, frame)
    local = torch.cat([torch.stack([torch.full((7,), float(c)), torch.arange(7.0)]).sum(0).repeat(sizes[c], 1) +
                       torch.arange(sizes[c]).float()[:, None] * 1000 for c in shards[rank]]) if shards[rank] else torch.zeros(0, 7)
    counts = [sum(sizes[c] for c in s) for s in shards]
    full = adist.allgather_rows(local, counts)
    per_clip = adist.scatter_back(full, shards, sizes)
    ok = all(per_clip[c].shape[0] == sizes[c] and float(per_clip[c][0, 0]) == float(c) and
             float(per_clip[c][-1, 0]) == c + (sizes[c] - 1) * 1000 for c in range(len(sizes)))
    q.put((rank, ok, full.shape[0]))
    dist.barrier()
    dist.destroy_process_group()


def test_allgather_rows_world2_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    sizes = [5, 12, 3, 9, 7]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, sizes, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res) and all(n == sum(sizes) for _, _, n in res)


def _runner_worker(rank, world, port, n_frames, n_windows, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    adist.init_from_env("gloo")
    run = adist.ShardedRunner(None, n_frames, n_windows, [False] * len(n_frames), rank, world)
    n = run.counts[rank]
    stat, dyn, a = run.block_views(run.send, n)
    # the rank's "results": value = 1000 * clip + frame (+ 0.25 / 0.5 for the dyn / audio streams), written in place
    pos = 0
    for c in run.my_clips:
        f = torch.arange(n_frames[c], dtype=torch.float32)[:, None] + 1000.0 * c
        stat[pos: pos + n_frames[c]] = f
        dyn[pos: pos + n_frames[c]] = f + 0.25
        a[pos: pos + n_frames[c]] = f + 0.5
        pos += n_frames[c]
    recv = run.exchange()
    ok = True
    for r, sh in enumerate(run.shards):
        st, dy, au = run.block_views(recv[r], run.counts[r])
        pos = 0
        for c in sh:
            want = torch.arange(n_frames[c], dtype=torch.float32) + 1000.0 * c
            ok &= bool(torch.equal(st[pos: pos + n_frames[c], 3], want) and torch.equal(dy[pos: pos + n_frames[c], 0], want + 0.25)
                       and torch.equal(au[pos: pos + n_frames[c], 7], want + 0.5))
            ok &= run.clip_slices[c] == slice(run.offsets[r] + pos, run.offsets[r] + pos + n_frames[c])
            pos += n_frames[c]
    q.put((rank, ok, run.total, run.counts))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_runner_exchange_world2_gloo():
    """dist.ShardedRunner's host logic on CPU: LPT shards from frame / window counts, per-rank blocks of the packed send
    buffer (three row arrays at fixed offsets), ONE all_gather_into_tensor, block views and clip slices of the result."""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    n_frames = [250, 1500, 750, 3000, 500, 1500, 100]
    n_windows = [21, 121, 61, 241, 41, 121, 9]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_runner_worker, args=(r, 2, port, n_frames, n_windows, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _, _ in res) and all(t == sum(n_frames) for _, _, t, _ in res)
    counts = res[0][3]
    assert abs(counts[0] - counts[1]) <= max(n_frames) and sum(counts) == sum(n_frames)
