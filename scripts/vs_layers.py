#!/usr/bin/env python
"""GPU probe: per-launch timing of one eager VS forward (batch 256 by default) -- kernel, ms, TFLOP/s and algorithmic GB/s of
every contraction, in launch order (CUDA events around each launch; L2 is warm from the previous layer as in the real step)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from avcer_b200 import nets, ops, synthetic as syn      # noqa: E402

DEV = "cuda:0"
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
vs = nets.VSNet(syn.make_vs_state_dict(0, "default"), "bf16", DEV)
gen = torch.Generator(device=DEV).manual_seed(0)
crops = torch.randint(0, 256, (n, 224, 224, 3), dtype=torch.uint8, device=DEV, generator=gen)
x = vs.alloc_input(n)
ops.preprocess(crops, n, x, vs.input_layout)
for _ in range(2):
    vs.forward(x)
torch.cuda.synchronize()
runs = []
for rep in range(3):
    ops.PROFILE = prof = []
    vs.forward(x)
    torch.cuda.synchronize()
    ops.PROFILE = None
    runs.append([(t, w, a.elapsed_time(b), nb) for t, w, a, b, nb in prof])
best = [min((r[i] for r in runs), key=lambda e: e[2]) for i in range(len(runs[0]))]
tot = 0.0
for i, (tag, work, ms, nbytes) in enumerate(best):
    tot += ms
    print(f"{i:3d} {tag:48s} {ms * 1e3:8.1f} us  {work / ms / 1e9:7.0f} TFLOP/s  {nbytes / ms / 1e6:7.0f} GB/s  ({nbytes / 1e6:7.1f} MB)")
print(f"sum of timed launches {tot:.3f} ms for batch {n}")
