"""GPU probe: per-contraction CUDA-event times of one eager audio forward (B windows, bf16) against the whole forward:
how much of the in-pipeline time is outside the tcgen05 contractions (conv0, LayerNorms, attention, pools)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avcer_b200 import nets, ops, synthetic as syn

dev = "cuda:0"
B = int(os.environ.get("B", "64"))
net = nets.ANet(syn.make_audio_state_dict(2, 8, "spread", 12), "bf16", dev)
x = torch.randn(B, 64000, device=dev)
names = [f"conv{i}" for i in range(1, 7)] + ["proj", "posconv"]
for l in range(12):
    names += [f"L{l}.qkv", f"L{l}.o", f"L{l}.ff1", f"L{l}.ff2"]
for t in ("tl1", "tl2"):
    names += [f"{t}.qkv", f"{t}.o", f"{t}.ff1", f"{t}.ff2"]
names += ["td0", "td4"]
for it in range(3):
    ops.PROFILE = []
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record(); net.forward(x); t1.record()
    torch.cuda.synchronize()
    prof, ops.PROFILE = ops.PROFILE, None
rows = [(nm, a.elapsed_time(b) * 1e3, work) for (tag, work, a, b), nm in zip([p for p in prof if p[0].startswith("contract")], names)]
# gaps between consecutive contractions = the non-GEMM kernels launched in between (+ launch gaps)
evs = [p for p in prof if p[0].startswith("contract")]
gaps = {}
for (p0, n0), (p1, n1) in zip(zip(evs, names), zip(evs[1:], names[1:])):
    gaps[f"{n0}->{n1}"] = p0[3].elapsed_time(p1[2]) * 1e3
tot_g = sum(r[1] for r in rows)
print(f"eager forward {t0.elapsed_time(t1):.3f} ms; contractions {tot_g / 1e3:.3f} ms ({len(rows)} launches); "
      f"before first {t0.elapsed_time(evs[0][2]) * 1e3:.0f} us (normalise + conv0); after last {evs[-1][3].elapsed_time(t1) * 1e3:.0f} us")
agg = {}
for nm, us, work in rows:
    k = nm.split(".")[-1] if "." in nm and nm[0] == "L" else nm
    agg.setdefault(k, [0.0, 0.0]); agg[k][0] += us; agg[k][1] += work
print("contractions:", "  ".join(f"{k}={v[0]:.0f}us({v[1] / v[0] / 1e6:.0f}TF)" for k, v in agg.items()))
gagg = {}
for k, v in gaps.items():
    a, b = k.split("->")
    key = (a.split(".")[-1] if a[0] == "L" else a) + "->" + (b.split(".")[-1] if b[0] == "L" else b)
    gagg[key] = gagg.get(key, 0.0) + v
print("between:", "  ".join(f"{k}={v:.0f}us" for k, v in sorted(gagg.items(), key=lambda kv: -kv[1])[:14]))
