#!/usr/bin/env python
"""GPU probe: VS forward (batch 256 / 750, K1 standalone vs fused into the stem) and audio forward (64 windows, LayerNorm
folded vs explicit passes), each as a CUDA-graph replay timed with CUDA events (median of 20, L2 flushed between)."""
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
for n in (256, 750):
    crops = torch.randint(0, 256, (n, 224, 224, 3), dtype=torch.uint8, device=DEV, generator=gen)
    x = vs.alloc_input(n)

    def k1_then_forward():
        ops.preprocess(crops, n, x, vs.input_layout)
        return vs.forward(x)

    t0 = timed(k1_then_forward)
    t1 = timed(lambda: vs.forward_u8(crops))
    t2 = timed(lambda: ops.stem_pool(x, vs.w["stem_packed"], vs.w["stem"].bias))
    t3 = timed(lambda: ops.stem_pool_u8(crops, vs.w["stem_packed"], vs.w["stem"].bias))
    t4 = timed(lambda: ops.preprocess(crops, n, x, vs.input_layout))
    print(f"VS batch {n}: K1 + forward {t0:.3f} ms | fused-K1 forward {t1:.3f} ms | stem_pool {t2 * 1e3:.1f} us, stem_pool_u8 {t3 * 1e3:.1f} us, K1 {t4 * 1e3:.1f} us"
          f" | {n * 7.667e9 / (t1 / 1e3) / 1e12:.0f} TFLOP/s")

a = nets.ANet(syn.make_audio_state_dict(2, 8, "spread", 12), "bf16", DEV)
xa = torch.randn((64, 64000), device=DEV, generator=gen)
print(f"A forward, 64 windows: {timed(lambda: a.forward(xa)):.3f} ms")
m, k = 64 * 199, 1024
bf = torch.bfloat16
h = torch.randn(m, k, device=DEV).to(bf)
for name, n_out, kk, act, res in (("qkv", 3072, 1024, ops.ACT_NONE, False), ("ffn1", 4096, 1024, ops.ACT_GELU, False),
                                  ("o-proj", 1024, 1024, ops.ACT_NONE, True), ("ffn2", 1024, 4096, ops.ACT_NONE, True)):
    x = torch.randn(m, kk, device=DEV).to(bf)
    w = (torch.randn(n_out, kk, device=DEV) / kk ** 0.5).to(bf)
    b = torch.zeros(n_out, device=DEV)
    r = h if res else None
    print(f"GEMM {name:7s} M={m} N={n_out} K={kk}: {timed(lambda: ops.linear(x, w, b, act=act, residual=r)) * 1e3:.1f} us")
xl = torch.randn(m, k, device=DEV).to(bf)
g1, b1 = torch.ones(k, device=DEV), torch.zeros(k, device=DEV)
print(f"layernorm kernel [M,1024]: {timed(lambda: ops.layernorm(xl, g1, b1, 1e-5)) * 1e3:.1f} us")
