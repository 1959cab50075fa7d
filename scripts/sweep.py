#!/usr/bin/env python
"""BASELINE config 5: throughput sweep on one GPU -- clip length 10 s .. 10 min x VS batch 32 .. 1024 crops (audio batch 64),
whole path (K1 -> VS -> VD, A, alignment, K4) through Engine.run_clips on device-resident synthetic clips.
Prints one JSON object (frames/s per cell, device-timed with CUDA events, 2 warm-up + 3 timed passes per cell).

    python scripts/sweep.py > gpurun_out/r02_sweep.json
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from avcer_b200 import get_weights_matrices as gwm, synthetic as syn      # noqa: E402
from avcer_b200.pipeline import Engine                                     # noqa: E402

dev = "cuda:0"
sds = (syn.make_vs_state_dict(0, "default"), syn.make_vd_state_dict(1), syn.make_audio_state_dict(2, 8, "spread", 12))
w1, w2 = gwm.class_weights(gwm.weights_3), [1, 1, 1]
fps = 25
total_seconds = 600                               # every cell processes 10 minutes of material
res = {"unit": "frames/s", "audio_batch": 64, "seconds_per_cell": total_seconds, "cells": []}
g = torch.Generator(device=dev).manual_seed(5)
crops = torch.randint(0, 256, (total_seconds * fps, 224, 224, 3), dtype=torch.uint8, device=dev, generator=g)
wav = torch.randn(total_seconds * 16000, device=dev, generator=g) * 0.1
for vs_batch in (32, 256, 1024):
    eng = Engine(*sds, precision="bf16", device=dev, vs_batch=vs_batch, a_batch=64)
    for clip_s in (10, 60, 600):
        n_clips = total_seconds // clip_s
        exists = [np.ones(clip_s * fps, bool) for _ in range(n_clips)]
        lens = [clip_s * 16000] * n_clips

        def run():
            return eng.run_clips(crops, exists, [float(fps)] * n_clips, wav, lens, w1, w2, False, True)

        for _ in range(2):
            run()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            run()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 3
        res["cells"].append({"clip_seconds": clip_s, "clips": n_clips, "vs_batch": vs_batch, "ms": ms,
                             "frames_per_s": total_seconds * fps / ms * 1e3, "audio_seconds_per_s": total_seconds / ms * 1e3})
    del eng
    torch.cuda.empty_cache()
print(json.dumps(res, indent=1))
