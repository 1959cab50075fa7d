#!/usr/bin/env python
"""GPU probe: VS forward (K1 + ResNet-50, CUDA-graph replay, L2 flushed between replays) with consecutive contractions walking
their tiles in the same direction vs alternating directions (VSNet.alternate); outputs must be bit-identical."""
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from avcer_b200 import nets, ops, synthetic as syn      # noqa: E402

DEV = "cuda:0"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)


def timed(fn, reps=20):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.replay()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts)


vs = nets.VSNet(syn.make_vs_state_dict(0, "default"), "bf16", DEV)
gen = torch.Generator(device=DEV).manual_seed(0)
for n in (256, 750, 1536):
    crops = torch.randint(0, 256, (n, 224, 224, 3), dtype=torch.uint8, device=DEV, generator=gen)
    x = vs.alloc_input(n)

    def run():
        ops.preprocess(crops, n, x, vs.input_layout)
        return vs.forward(x)

    res = {}
    outs = {}
    for alt in (False, True):
        vs.alternate = alt
        outs[alt] = [t.clone() for t in run()]
        res[alt] = timed(run)
    same = all(bool((a == b).all()) for a, b in zip(outs[False], outs[True]))
    print(f"VS batch {n}: same direction {res[False]:.3f} ms | alternating {res[True]:.3f} ms ({res[True] / res[False]:.3f}x) = "
          f"{n * 7.667e9 / (res[True] / 1e3) / 1e12:.0f} TFLOP/s{'' if same else '  OUTPUT DIFFERS'}", flush=True)
