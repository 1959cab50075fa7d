"""Audio-side non-GEMM kernels for `ncu --set full`: attention (64 windows x 16 heads), conv0+LN+GELU, LayerNorm."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avcer_b200 import ops, _lib
_lib.require_device()
dev = "cuda"; bf = torch.bfloat16
torch.manual_seed(0)
n, t = 64, 199
qkv = torch.randn(n * t, 3072, device=dev).to(bf)
for _ in range(2):
    ops.attention(qkv, n, t, 16, 64, 0.125)
x = torch.randn(n, 64000, device=dev)
w = torch.randn(512, 10, device=dev) * 0.3; b = torch.randn(512, device=dev) * 0.1
g = torch.ones(512, device=dev); be = torch.zeros(512, device=dev)
y = torch.empty(n, 12799, 512, device=dev, dtype=bf)
for _ in range(2):
    ops.w2v_conv0_ln_gelu(x, w, b, g, be, y)
h = torch.randn(n * 6399, 512, device=dev).to(bf)
for _ in range(2):
    ops.layernorm(h, g, be, 1e-5, act=ops.ACT_GELU, out=h)
torch.cuda.synchronize()
print("done")
