"""Representative epilogue-bound VS launches for `ncu --set full --import-source on` (second launch of each)."""
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


conv(B, 14, 14, 256, 1024, 1, res=True)      # l3.c3  (62 us; 231 MB)
conv(B, 28, 28, 512, 1024, 1, stride=2)      # l3.0.ds (49 us)
conv(B, 14, 14, 1024, 256, 1)                # l3.1.c1 (35 us)
conv(B, 55, 55, 64, 256, 1, res=True)        # l1.c3  (187 us; 891 MB)
conv(B, 55, 55, 256, 64, 1)                  # l1.1.c1 (84 us; 495 MB)
print("prof_kernels2 done")
