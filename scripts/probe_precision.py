#!/usr/bin/env python
"""GPU probe: measured per-class probability error of the VS / VD / A forwards against the unmodified reference's outputs
(tests/golden) in every precision of the library: fp32 (SIMT), bf16 (libavcer_b200.so) and fp16 (libavcer_b200_fp16.so)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from avcer_b200 import nets, ops, pipeline, synthetic as syn      # noqa: E402

DEV = "cuda:0"
G = {n: np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", n + ".npz")) for n in ("video", "audio")}
crops = syn.make_crops(11, 6)
print("max |p - p_reference| over inputs and classes (north star: 2e-3 in 16-bit modes, 1e-5 in fp32)")
for init in ("default", "mid", "spread"):
    row = []
    for prec in ("fp32", "bf16", "fp16"):
        net = nets.VSNet(syn.make_vs_state_dict(0, init), prec, DEV)
        x = net.alloc_input(6)
        ops.preprocess(torch.from_numpy(crops).to(DEV), 6, x, net.input_layout)
        p = net.forward(x)[0].cpu().numpy()
        row.append(f"{prec} {np.abs(p - G['video'][f'vs_{init}_probs']).max():.2e}")
    print(f"VS  {init:8s}: " + "   ".join(row))
gen = torch.Generator().manual_seed(5)
xw = torch.relu(torch.randn(12, 10, 512, generator=gen))
row = []
for prec in ("fp32", "bf16", "fp16"):
    net = nets.VDNet(syn.make_vd_state_dict(1), prec, DEV)
    wins = torch.arange(120, dtype=torch.int32).view(12, 10).t().contiguous().to(DEV)
    out = net.forward(xw.reshape(120, 512).to(DEV), wins).cpu()
    row.append(f"{prec} {(torch.softmax(out, 1) - torch.softmax(torch.from_numpy(G['video']['vd_logits']), 1)).abs().max().item():.2e}")
print("VD  spread  : " + "   ".join(row))
wav = syn.make_wav(31, 52800 + 123)
ap = pipeline.plan_audio(len(wav), 25, 0.5)
x = ops.audio_normalize_windows(torch.from_numpy(wav).to(DEV), torch.from_numpy(ap.starts).to(DEV), 64000, "mean")
for init, key in (("default", "a8_default_window_logits"), ("mid", "a8_mid_window_logits"), ("spread", "a8_a_window_logits")):
    row = []
    ref = torch.softmax(torch.from_numpy(G["audio"][key][:, :7]), 1)
    for prec in ("fp32", "bf16", "fp16"):
        net = nets.ANet(syn.make_audio_state_dict(2, 8, init, 12), prec, DEV)
        out = net.forward(x).cpu()
        row.append(f"{prec} {(torch.softmax(out[:, :7], 1) - ref).abs().max().item():.2e}")
    print(f"A   {init:8s}: " + "   ".join(row))
