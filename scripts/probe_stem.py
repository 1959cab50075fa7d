"""GPU probe: fused stem+pool kernel vs the two-kernel path (bits and time)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avcer_b200 import nets, ops, synthetic as syn
dev = "cuda:0"
N = 256
net = nets.VSNet(syn.make_vs_state_dict(0, "default"), "bf16", dev)
crops = torch.randint(0, 256, (N, 224, 224, 3), dtype=torch.uint8, device=dev)
x = net.alloc_input(N)
ops.preprocess(crops, N, x, net.input_layout)
ref = ops.maxpool3x3s2(net.stem(x))
got = ops.stem_pool(x, net.w["stem_packed"], net.w["stem"].bias)
torch.cuda.synchronize()
d = (got.float() - ref.float()).abs()
print("bit-identical:", torch.equal(got.view(torch.int16), ref.view(torch.int16)), "max|d|", d.max().item(), "bad", int((d > 0).sum()))
if (d > 0).any():
    idx = (d > 0).nonzero()
    print("first bad", idx[:8].tolist(), "bad rows", torch.unique(idx[:, 1]).tolist()[:60])
for name, fn in (("two kernels", lambda: ops.maxpool3x3s2(net.stem(x))), ("fused", lambda: ops.stem_pool(x, net.w["stem_packed"], net.w["stem"].bias))):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(5):
            fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 200
    print(f"{name}: {us:.1f} us per 256 crops  ({2 * N * 112 * 112 * 64 * 147 / us / 1e6:.0f} TFLOP/s on the 147 real taps)")
