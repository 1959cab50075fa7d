"""GPU probe: attention kernel (tcgen05 for dh=64, AVCER_ATT5=0: mma.sync) against torch + timing at 64 windows."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avcer_b200 import ops
dev = "cuda:0"; bf = torch.bfloat16
torch.manual_seed(0)


def check(n, t, heads, dh):
    qkv = (torch.randn(n * t, 3 * heads * dh, device=dev) * 1.5).to(bf)
    scale = dh ** -0.5
    out = ops.attention(qkv, n, t, heads, dh, scale)
    torch.cuda.synchronize()
    q, k, v = (z.float().view(n, t, heads, dh).transpose(1, 2) for z in qkv.split(heads * dh, dim=1))
    ref = (torch.softmax(q @ k.transpose(-1, -2) * scale, -1) @ v).transpose(1, 2).reshape(n * t, heads * dh)
    d = (out.float() - ref).abs()
    print(f"[{'OK ' if d.max().item() < 2e-2 else 'BAD'}] n={n} t={t} heads={heads} dh={dh}: max|d|={d.max().item():.4g} mean|d|={d.mean().item():.3g} nan={int(torch.isnan(out).sum())}", flush=True)
    if d.max().item() >= 2e-2:
        bad = (d > 2e-2).nonzero()
        print("   bad rows(mod t):", torch.unique(bad[:, 0] % t).tolist()[:40], "cols(mod 64):", torch.unique(bad[:, 1] % 64).tolist()[:40], flush=True)


for (n, t) in ((1, 199), (3, 199), (2, 208), (3, 113), (5, 50), (4, 1), (64, 199)):
    check(n, t, 16, 64)
n, t = 64, 199
qkv = torch.randn(n * t, 3072, device=dev).to(bf)
out = torch.empty(n * t, 1024, device=dev, dtype=bf)
fn = lambda: ops.attention(qkv, n, t, 16, 64, 0.125, out=out)
fn(); torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for _ in range(10):
        fn()
g.replay(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
print(f"attention 64 windows x 16 heads: {e0.elapsed_time(e1) * 100:.1f} us  (AVCER_ATT5={os.environ.get('AVCER_ATT5')})")
