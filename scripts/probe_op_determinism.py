"""GPU probe: run single ops of the audio encoder repeatedly on fixed inputs; every run must be bit-identical."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avcer_b200 import ops

dev = "cuda:0"
B, T = int(os.environ.get("B", "64")), 199
M = B * T
g = torch.Generator(device=dev).manual_seed(3)
rn = lambda *s: torch.randn(s, device=dev, generator=g)
bf = torch.bfloat16
h = rn(M, 1024).to(bf)
qkv = rn(M, 3072).to(bf)
wqkv, bqkv = (rn(3072, 1024) * 0.03).to(bf), rn(3072) * 0.1
wo, bo = (rn(1024, 1024) * 0.03).to(bf), rn(1024) * 0.1
w1, b1 = (rn(4096, 1024) * 0.03).to(bf), rn(4096) * 0.1
w2, b2 = (rn(1024, 4096) * 0.02).to(bf), rn(1024) * 0.1
f = rn(M, 4096).to(bf)
gam, bet = rn(1024), rn(1024)
N = int(os.environ.get("RUNS", "40"))
cases = {
    "layernorm": lambda: ops.layernorm(h, gam, bet, 1e-5),
    "qkv gemm": lambda: ops.linear(h, wqkv, bqkv),
    "attention 16x64": lambda: ops.attention(qkv, B, T, 16, 64, 0.125),
    "attention 32x32": lambda: ops.attention(qkv, B, T, 32, 32, 0.17),
    "o gemm + residual": lambda: ops.linear(h, wo, bo, residual=h),
    "ff1 gemm gelu": lambda: ops.linear(h, w1, b1, act=ops.ACT_GELU),
    "ff2 gemm + residual": lambda: ops.linear(f, w2, b2, residual=h),
    "tl ff relu": lambda: ops.linear(h, wo, bo, act=ops.ACT_RELU),
}
for name, fn in cases.items():
    ref = fn().clone()
    bad = 0
    worst = 0.0
    where = None
    for r in range(N):
        out = fn()
        d = (out.float() - ref.float()).abs()
        nb = int((d > 0).sum())
        if nb:
            bad += 1
            worst = max(worst, float(d.max()))
            if where is None:
                where = (r, nb, (d > 0).nonzero()[:3].tolist())
    print(f"{name:22s}: {bad}/{N} runs differ  worst {worst:.4g}  first {where}")
