"""GPU: the clip-sharded path (dist.ShardedRunner: per-frame rows written straight into the all-gather send buffer, one
collective, K4 per gathered block with a label pitch) against the single-GPU Engine.run_clips on the same clips.

world 1 runs on any B200 box; world 2 needs two GPUs (NCCL) and is skipped otherwise -- run it with
`gpurun --gpus 2 -- python -m pytest tests/test_gpu_dist.py -m gpu`."""
import os
import socket

import numpy as np
import pytest
import torch

from avcer_b200 import synthetic as syn

pytestmark = pytest.mark.gpu

CLIP_FRAMES = [60, 25, 110, 40, 75]          # mixed lengths; clip 1 starts with missing crops (float64 tables), clip 2 has a gap
FPS = 25


def _clips():
    exists, crops, wavs = [], [], []
    for i, n in enumerate(CLIP_FRAMES):
        ex = np.ones(n, bool)
        if i == 1:
            ex[:3] = False
        if i == 2:
            ex[[50, 51]] = False
        exists.append(ex)
        crops.append(syn.make_crops(900 + i, int(ex.sum())))
        wavs.append(syn.make_wav(950 + i, int(n / FPS * 16000) - 160))
    return exists, crops, wavs


def _engine(dev):
    from avcer_b200.pipeline import Engine

    return Engine(syn.make_vs_state_dict(0, "spread"), syn.make_vd_state_dict(1), syn.make_audio_state_dict(2, 8, "spread", 2),
                  precision="bf16", device=dev, vs_batch=64, a_batch=8)


def _run_sharded(eng, rank, world, exists, crops, wavs, w1, w2):
    from avcer_b200 import dist as adist
    from avcer_b200.pipeline import Engine, plan_audio

    n_windows = [len(plan_audio(len(w), FPS).starts) for w in wavs]
    run = adist.ShardedRunner(eng, CLIP_FRAMES, n_windows, [Engine.needs_f64(e, FPS) for e in exists], rank, world)
    mine = run.my_clips
    c = torch.from_numpy(np.concatenate([crops[i] for i in mine])) if mine else torch.zeros((0, 224, 224, 3), dtype=torch.uint8)
    w = torch.from_numpy(np.concatenate([wavs[i] for i in mine])) if mine else torch.zeros(0)
    labels = run.step(c, [exists[i] for i in mine], [FPS] * len(mine), w, [len(wavs[i]) for i in mine], w1, w2, False, True)
    torch.cuda.synchronize()
    return run, labels.cpu().numpy()


def _single_gpu_labels(eng, exists, crops, wavs, w1, w2):
    out = eng.run_clips(torch.from_numpy(np.concatenate(crops)), exists, [FPS] * len(exists), torch.from_numpy(np.concatenate(wavs)),
                        [len(w) for w in wavs], w1, w2, False, True)
    return out["labels"].cpu().numpy()


def test_sharded_runner_world1_equals_run_clips(cuda_lib):
    from avcer_b200 import get_weights_matrices as gwm

    exists, crops, wavs = _clips()
    eng = _engine("cuda:0")
    for w1 in (gwm.class_weights(gwm.weights_3), None):
        ref = _single_gpu_labels(eng, exists, crops, wavs, w1, [1, 1, 1])
        run, got = _run_sharded(eng, 0, 1, exists, crops, wavs, w1, [1, 1, 1])
        assert got.shape == ref.shape == (4, sum(CLIP_FRAMES))
        base = np.r_[0, np.cumsum(CLIP_FRAMES)]
        for i in range(len(CLIP_FRAMES)):
            assert np.array_equal(got[:, run.clip_slices[i]], ref[:, base[i]: base[i + 1]]), i


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from avcer_b200 import dist as adist, get_weights_matrices as gwm

    adist.init_from_env("nccl")
    exists, crops, wavs = _clips()
    eng = _engine(f"cuda:{rank}")
    w1, w2 = gwm.class_weights(gwm.weights_3), [1, 1, 1]
    run, got = _run_sharded(eng, rank, world, exists, crops, wavs, w1, w2)
    ok = True
    if rank == 0:
        ref = _single_gpu_labels(eng, exists, crops, wavs, w1, w2)
        base = np.r_[0, np.cumsum(CLIP_FRAMES)]
        ok = all(np.array_equal(got[:, run.clip_slices[i]], ref[:, base[i]: base[i + 1]]) for i in range(len(CLIP_FRAMES)))
    q.put((rank, ok, [len(s) for s in run.shards], got.sum()))
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


def test_sharded_runner_world2_nccl_equals_single_gpu(cuda_lib):
    """Two ranks, two GPUs, NCCL: the gathered labels on every rank equal the 1-GPU labels of the same clips."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert all(ok for _, ok, _, _ in res)
    assert res[0][3] == res[1][3] and all(n > 0 for n in res[0][2])          # same labels on both ranks; both ranks had work
