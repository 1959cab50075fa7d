#!/usr/bin/env python
"""GPU probe: face detector (RetinaFace-ResNet50) on synthetic 720p / 1080p frames -- whole detect() per
batch (CUDA events, median, L2 flushed between), per-kernel breakdown of one eager batch, end-to-end detect_batch from host
frames, and the oracle (the reference's network restated in torch fp32) on the host cores for the same frame size."""
import collections
import os
import statistics
import sys
import time
from types import SimpleNamespace

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from avcer_b200 import nets, ops, synthetic as syn      # noqa: E402
from avcer_b200.data.face_detection import RetinaFacePredictor, cfg_re50      # noqa: E402

DEV = "cuda:0"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
sd = syn.make_retinaface_state_dict(5, "spread")
FLOP = {}


def timed(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts)


for prec in ("bf16", "fp16"):
    net = nets.RetinaFaceNet(sd, prec, DEV)
    for (h, w, n) in ((720, 1280, 8), (1080, 1920, 4), (1080, 1920, 8)):
        frames = torch.from_numpy(syn.make_frames(7, n, h, w)).to(DEV)
        t = timed(lambda: net.detect(frames))
        ops.PROFILE = prof = []
        net.detect(frames)
        torch.cuda.synchronize()
        ops.PROFILE = None
        agg = collections.OrderedDict()
        for tag, work, a, b, nbytes in prof:
            d = agg.setdefault(tag, [0.0, 0.0, 0])
            d[0] += work; d[1] += a.elapsed_time(b); d[2] += 1
        flop = sum(v[0] for v in agg.values())
        print(f"{prec} {h}x{w} batch {n}: detect {t:.2f} ms = {n / t * 1e3:.0f} frames/s, {flop / n / 1e9:.1f} GFLOP/frame, {flop / t / 1e9:.0f} TFLOP/s", flush=True)
        if prec == "bf16" and n == 8 and h == 1080:
            for tag, (work, ms, cnt) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                print(f"    {tag:45s} {ms:8.3f} ms  n={cnt:3d}  {work / max(ms, 1e-9) / 1e9:8.0f} TFLOP/s")

pred = RetinaFacePredictor(threshold=0.8, device=DEV, model=SimpleNamespace(weights=sd, config=SimpleNamespace(**cfg_re50)), precision="bf16")
host = syn.make_frames(8, 8, 1080, 1920)
pred.detect_batch(host)
t0 = time.perf_counter()
for _ in range(3):
    out = pred.detect_batch(host)
dt = (time.perf_counter() - t0) / 3
t0 = time.perf_counter()
for _ in range(3):
    dev_frames = pred._upload(host)
    torch.cuda.synchronize()
t_up = (time.perf_counter() - t0) / 3
t0 = time.perf_counter()
for _ in range(3):
    pred.detect_batch(dev_frames)
t_dev = (time.perf_counter() - t0) / 3
print(f"detect_batch from host frames (H2D + detect + select + NMS), 8 x 1080p: {dt * 1e3:.1f} ms = {8 / dt:.0f} frames/s "
      f"(upload {t_up * 1e3:.1f} ms, from device frames {t_dev * 1e3:.1f} ms); detections per frame {[len(o) for o in out]} "
      "(synthetic weights: hundreds of boxes per frame reach the host-side NMS; a trained detector yields a handful)")
strict = RetinaFacePredictor(threshold=0.999, device=DEV, model=SimpleNamespace(weights=sd, config=SimpleNamespace(**cfg_re50)), precision="bf16")
strict.detect_batch(host)
t0 = time.perf_counter()
for _ in range(3):
    out = strict.detect_batch(host)
dt = (time.perf_counter() - t0) / 3
print(f"same with threshold 0.999 ({[len(o) for o in out]} detections per frame): {dt * 1e3:.1f} ms = {8 / dt:.0f} frames/s")

from oracle import face as ofa      # noqa: E402  (CPU baseline leg: the reference's network restated in torch fp32)
torch.set_num_threads(os.cpu_count() or 1)
x = ofa.prepare(host[0])
ofa.forward(x, sd)
t0 = time.perf_counter()
ofa.predict(sd, host[0])
print(f"oracle (torch fp32, {torch.get_num_threads()} host threads), one 1080p frame: {(time.perf_counter() - t0) * 1e3:.0f} ms")
