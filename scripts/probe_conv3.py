"""GPU probe: halo-in-smem 3x3 conv kernel (conv3x3.cuh) against torch + timing (AVCER_CONV3=0: generic kernel)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from avcer_b200 import ops, _lib

_lib.require_device()
lib = ctypes.CDLL(_lib.LIB_PATH)
dev = "cuda"; bf = torch.bfloat16
torch.manual_seed(0)
torch.backends.cudnn.allow_tf32 = False


def check(n, h, w, c, cout):
    x = torch.randn(n, h, w, c, device=dev).to(bf)
    w4 = (torch.randn(cout, c, 3, 3, device=dev) / (9 * c) ** 0.5).to(bf)
    b = torch.randn(cout, device=dev)
    wt = w4.permute(0, 2, 3, 1).reshape(cout, 9 * c).contiguous()
    y = ops.conv2d_nhwc(x, wt, b, kh=3, kw=3, pad_h=1, pad_w=1, act=ops.ACT_RELU)
    torch.cuda.synchronize()
    r = F.relu(F.conv2d(x.float().permute(0, 3, 1, 2), w4.float(), b, padding=1)).permute(0, 2, 3, 1)
    err = (y.float() - r).abs()
    bad = (err > 0.05)
    print(f"[{'OK ' if not bad.any() else 'BAD'}] {n}x{h}x{w} {c}->{cout}: max|err|={err.max().item():.4g} bad={int(bad.sum())}/{err.numel()}", flush=True)
    if bad.any():
        idx = bad.nonzero()
        print("    first bad:", idx[:6].tolist(), " bad per (h): ", torch.bincount(idx[:, 1], minlength=h).tolist()[:60], flush=True)
    return x, wt, b


check(3, 55, 55, 64, 64)
check(2, 28, 28, 128, 128)
check(300, 55, 55, 64, 64)
check(200, 28, 28, 128, 128)
check(7, 30, 40, 128, 128)

for (n, h, w, c, cout) in ((256, 55, 55, 64, 64), (256, 28, 28, 128, 128)):
    x = torch.randn(n, h, w, c, device=dev).to(bf)
    wt = (torch.randn(cout, 9 * c, device=dev) / (9 * c) ** 0.5).to(bf)
    b = torch.randn(cout, device=dev)
    out = torch.empty(n, h, w, cout, device=dev, dtype=bf)
    fn = lambda: ops.conv2d_nhwc(x, wt, b, kh=3, kw=3, pad_h=1, pad_w=1, act=ops.ACT_RELU, out=out)
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(10):
            fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    print(f"{n}x{h}x{w} {c}->{cout}: {us:.1f} us  {2 * n * h * w * cout * 9 * c / us / 1e6:.0f} TFLOP/s  (AVCER_CONV3={os.environ.get('AVCER_CONV3')})")
print("probe done")
