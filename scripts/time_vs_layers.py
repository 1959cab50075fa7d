"""GPU probe: per-layer CUDA-event times of one eager VS ResNet-50 forward (batch N, bf16) plus the
graph-replayed whole-forward time with L2 flushed between replays.  Usage: N=256 python scripts/time_vs_layers.py [quiet]"""
import os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avcer_b200 import nets, ops, synthetic as syn

dev = "cuda:0"
N = int(os.environ.get("N", "256"))
quiet = "quiet" in sys.argv
net = nets.VSNet(syn.make_vs_state_dict(0, "default"), "bf16", dev)
net.fused_shortcut = os.environ.get("FS", "1") != "0"
net.sampled_tail = int(os.environ.get("ST", "2"))
g = torch.Generator(device=dev).manual_seed(5)
crops = torch.randint(0, 256, (N, 224, 224, 3), dtype=torch.uint8, device=dev, generator=g)
x = net.alloc_input(N)
ops.preprocess(crops, N, x, net.input_layout)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

names = ["stem"]
for li, nb in enumerate((3, 4, 6, 3), 1):
    for b in range(nb):
        names += ([f"l{li}.{b}.ds"] if b == 0 and not net.fused_shortcut else []) + [f"l{li}.{b}.c1", f"l{li}.{b}.c2", f"l{li}.{b}.c3"]
names.append("fc1")

for it in range(3):
    ops.PROFILE = []
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record(); net.forward(x); t1.record()
    torch.cuda.synchronize()
    prof, ops.PROFILE = ops.PROFILE, None
tot = 0.0
agg = {}
for (tag, work, a, b), nm in zip([p for p in prof if p[0].startswith("contract")], names):
    us = a.elapsed_time(b) * 1e3
    tot += us
    kind = nm.split(".")[-1] if "." in nm else nm
    key = (nm.split(".")[0], kind)
    agg[key] = agg.get(key, 0.0) + us
    if not quiet:
        print(f"  {nm:10s} {us:8.1f} us  {work / us / 1e6:7.1f} TF/s")
print("per (stage, kind) us:", "  ".join(f"{k[0]}.{k[1]}={v:.0f}" for k, v in agg.items()))
print(f"eager forward {t0.elapsed_time(t1):.3f} ms, contractions {tot / 1e3:.3f} ms")

net.forward(x); torch.cuda.synchronize()
gr = torch.cuda.CUDAGraph()
with torch.cuda.graph(gr):
    net.forward(x)
ts = []
for i in range(8):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); gr.replay(); b.record()
    torch.cuda.synchronize()
    if i >= 3:
        ts.append(a.elapsed_time(b))
t = statistics.median(ts)
print(f"graph forward batch {N}: {t:.3f} ms  {N / t:.1f} kframes/s  {N * 7.667e9 / (t / 1e3) / 1e12:.0f} TFLOP/s "
      f"env RSLOTS={os.environ.get('AVCER_RSLOTS')} PDL={os.environ.get('AVCER_PDL')}")
