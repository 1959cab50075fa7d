#!/usr/bin/env python
"""Profiling target (run plain first, then under ncu): one face-detector forward on 4 x 1080p frames (bf16) and one audio
conv0 + LayerNorm + GELU on 16 windows, each after one warm-up call."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from avcer_b200 import nets, ops, synthetic as syn      # noqa: E402

DEV = "cuda:0"
net = nets.RetinaFaceNet(syn.make_retinaface_state_dict(5, "spread"), "bf16", DEV)
frames = torch.from_numpy(syn.make_frames(7, 4, 1080, 1920)).to(DEV)
a = nets.ANet(syn.make_audio_state_dict(2, 8, "spread", 12), "bf16", DEV)
x = torch.randn((16, 64000), device=DEV)
h = torch.empty((16, 12799, 512), device=DEV, dtype=torch.bfloat16)
for _ in range(2):
    dets = net.detect(frames)
    ops.w2v_conv0_tc(x, a.w["conv0_tc"], *a.w["conv_ln"][0], h)
    torch.cuda.synchronize()
print("ok", tuple(dets.shape), float(h.float().abs().mean()))
