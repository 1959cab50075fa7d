#!/usr/bin/env python
"""GPU probe: LayerNorm over [12736, 1024] 16-bit rows (the encoder's shape at 64 windows): cold (L2 flushed) and warm
(input just written, as in the forward) timings; AVCER_LN_VARIANT picks the kernel form."""
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from avcer_b200 import ops      # noqa: E402

DEV = "cuda:0"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
m, k = 64 * 199, 1024
x = torch.randn(m, k, device=DEV).to(torch.bfloat16)
y = torch.empty_like(x)
g1, b1 = torch.ones(k, device=DEV), torch.zeros(k, device=DEV)


def timed(fn, pre, reps=20):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    ts = []
    for _ in range(reps):
        pre()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.replay()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts) * 1e3


cold = timed(lambda: ops.layernorm(x, g1, b1, 1e-5, out=y), lambda: flush.zero_())
warm = timed(lambda: ops.layernorm(x, g1, b1, 1e-5, out=y), lambda: x.copy_(x))
print(f"variant {os.environ.get('AVCER_LN_VARIANT', 'default')}: cold {cold:.1f} us, warm {warm:.1f} us")
