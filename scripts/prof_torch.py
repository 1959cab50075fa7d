"""GPU probe: per-kernel device time of one eager VS (batch 256) and A (64 windows) forward via torch.profiler (CUPTI)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from avcer_b200 import nets, ops, synthetic as syn

dev = "cuda:0"
which = sys.argv[1] if len(sys.argv) > 1 else "a"
if which == "a":
    net = nets.ANet(syn.make_audio_state_dict(2, 8, "spread", 12), "bf16", dev)
    x = torch.randn(64, 64000, device=dev)
    fn = lambda: net.forward(x)
else:
    net = nets.VSNet(syn.make_vs_state_dict(0, "default"), "bf16", dev)
    crops = torch.randint(0, 256, (256, 224, 224, 3), dtype=torch.uint8, device=dev)
    x = net.alloc_input(256)
    ops.preprocess(crops, 256, x, net.input_layout)
    fn = lambda: net.forward(x)
for _ in range(3):
    fn()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    fn()
    torch.cuda.synchronize()
rows = [(e.key, e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total, e.count) for e in prof.key_averages()]
rows = [r for r in rows if r[1] > 0]
tot = sum(r[1] for r in rows)
print(f"{which}: total device time {tot / 1e3:.3f} ms")
for k, t, c in sorted(rows, key=lambda r: -r[1])[:14]:
    print(f"  {t:9.1f} us {100 * t / tot:5.1f}%  n={c:4d}  {k[:90]}")
