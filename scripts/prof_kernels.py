"""A handful of representative hot-path launches for `ncu --set full` (one launch of each after warm-up)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avcer_b200 import ops, _lib

_lib.require_device()
dev = "cuda"
torch.manual_seed(0)
bf = torch.bfloat16
B = 256


def conv(n, h, w, cin, cout, k, stride=1, res=False, reps=2):
    x = torch.randn(n, h, w, cin, device=dev).to(bf)
    wt = (torch.randn(cout, k * k * cin, device=dev) / (cin * k * k) ** 0.5).to(bf)
    b = torch.randn(cout, device=dev)
    ho, wo = (h - 1) // stride + 1, (w - 1) // stride + 1
    r = torch.randn(n, ho, wo, cout, device=dev).to(bf) if res else None
    for _ in range(reps):
        ops.conv2d_nhwc(x, wt, b, kh=k, kw=k, stride=stride, pad_h=(k - 1) // 2, pad_w=(k - 1) // 2, residual=r, act=ops.ACT_RELU)
    torch.cuda.synchronize()


# VS layers at batch 256 (kernel ids in the order launched; two launches each, profile the second)
conv(B, 14, 14, 256, 256, 3)                 # l3.c2  3x3 (MMA bound)
conv(B, 55, 55, 64, 256, 1, res=True)        # l1.c3  1x1 + residual (HBM bound)
conv(B, 55, 55, 64, 64, 3)                   # l1.c2  3x3, N = 64
# audio: ff1 GEMM with GELU, 32 windows
x = torch.randn(32 * 199, 1024, device=dev).to(bf); w = (torch.randn(4096, 1024, device=dev) / 32).to(bf); b = torch.zeros(4096, device=dev)
for _ in range(2):
    ops.linear(x, w, b, act=ops.ACT_GELU)
# K1 and K4 at full size
crops = torch.randint(0, 256, (1024, 224, 224, 3), dtype=torch.uint8, device=dev)
dst = torch.zeros((1024, 232, 232, 4), device=dev, dtype=bf)
for _ in range(2):
    ops.preprocess(crops, 1024, dst, 1)
n = 1_500_000
ps = [torch.softmax(torch.randn(n, 7, device=dev), 1).contiguous() for _ in range(3)]
from avcer_b200 import get_weights_matrices as gwm
for _ in range(2):
    ops.fuse_compound(ps[0], ps[1], ps[2], gwm.class_weights(gwm.weights_3), [1, 1, 1], False, True)
torch.cuda.synchronize()
print("prof_kernels done")
