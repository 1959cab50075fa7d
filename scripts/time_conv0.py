#!/usr/bin/env python
"""GPU probe: wav2vec2 conv0 + LayerNorm + GELU, SIMT fp32-arithmetic kernel vs the tcgen05 kernel (64 windows of 4 s), and
the whole audio forward with either (CUDA-graph replays, CUDA events, median of 20, L2 flushed between)."""
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


gen = torch.Generator(device=DEV).manual_seed(0)
for prec in ("bf16", "fp16"):
    a = nets.ANet(syn.make_audio_state_dict(2, 8, "spread", 12), prec, DEV)
    w = a.w
    for nwin in (64, 16):
        xa = torch.randn((nwin, 64000), device=DEV, generator=gen)
        h = torch.empty((nwin, 12799, 512), device=DEV, dtype=a.dtype)
        t_simt = timed(lambda: ops.w2v_conv0_ln_gelu(xa, w["conv0_w"], w["conv0_b"], *w["conv_ln"][0], h))
        t_tc = timed(lambda: ops.w2v_conv0_tc(xa, w["conv0_tc"], *w["conv_ln"][0], h))
        gb = (h.numel() * 2 + xa.numel() * 4) / 1e9
        print(f"{prec} conv0 {nwin} windows: SIMT {t_simt * 1e3:.1f} us ({gb / t_simt * 1e3:.0f} GB/s) | tcgen05 {t_tc * 1e3:.1f} us ({gb / t_tc * 1e3:.0f} GB/s)")
        a.conv0_tc = False
        f0 = timed(lambda: a.forward(xa))
        a.conv0_tc = True
        f1 = timed(lambda: a.forward(xa))
        print(f"{prec} A forward {nwin} windows: SIMT conv0 {f0:.3f} ms | tcgen05 conv0 {f1:.3f} ms")
