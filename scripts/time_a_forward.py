"""GPU probe: graph-replayed audio-network forward (B windows, bf16), L2 flushed between replays."""
import os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avcer_b200 import nets, synthetic as syn

dev = "cuda:0"
B = int(os.environ.get("B", "64"))
net = nets.ANet(syn.make_audio_state_dict(2, 8, "spread", 12), "bf16", dev)
x = torch.randn(B, 64000, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
net.forward(x); torch.cuda.synchronize()
gr = torch.cuda.CUDAGraph()
with torch.cuda.graph(gr):
    y = net.forward(x)
ts = []
for i in range(8):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); gr.replay(); b.record()
    torch.cuda.synchronize()
    if i >= 3:
        ts.append(a.elapsed_time(b))
t = statistics.median(ts)
print(f"A forward {B} windows: {t:.3f} ms  {t / B * 1e3:.1f} us/window  {B * 91.3e9 / (t / 1e3) / 1e12:.0f} TFLOP/s  finite={bool(torch.isfinite(y).all())}")
