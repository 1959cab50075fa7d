#!/usr/bin/env python
"""GPU probe: K1 (packed 224x224 crops -> zero-bordered bf16 NHWC4) at 1024 crops, 8 launches per CUDA-graph replay, L2
flushed between replays (the bench's method); AVCER_K1_VARIANT selects an experimental form of the kernel."""
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from avcer_b200 import ops      # noqa: E402

DEV = "cuda:0"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
flush2 = torch.zeros(64 << 20, dtype=torch.float32, device=DEV)
n = 1024
crops = torch.randint(0, 256, (n, 224, 224, 3), dtype=torch.uint8, device=DEV, generator=torch.Generator(device=DEV).manual_seed(0))
x = torch.zeros((n, 232, 240, 4), device=DEV, dtype=torch.bfloat16)
ops.preprocess(crops, n, x, 1)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for _ in range(8):
        ops.preprocess(crops, n, x, 1)
ts = []
for i in range(8):
    flush.zero_()
    flush2.sum()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    g.replay()
    b.record()
    torch.cuda.synchronize()
    if i >= 2:
        ts.append(a.elapsed_time(b) / 8)
t = statistics.median(ts)
print(f"variant {os.environ.get('AVCER_K1_VARIANT', '0')}: {t * 1e3:.1f} us, algorithmic {n * 451584 / t / 1e6:.0f} GB/s = {n * 451584 / t / 1e6 / 6452.8:.3f} of the copy peak; "
      f"sum check {float(x.float().sum()):.1f}")
