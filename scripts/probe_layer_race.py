"""GPU probe: one wav2vec2 encoder layer, kernel after kernel exactly as ANet._encoder_layer launches them (PDL chain),
repeated; reports the first intermediate that differs from the first run and the shape of the difference."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avcer_b200 import nets, ops, synthetic as syn

dev = "cuda:0"
B, T = int(os.environ.get("B", "64")), 199
net = nets.ANet(syn.make_audio_state_dict(2, 8, "spread", 12), "bf16", dev)
g = torch.Generator(device=dev).manual_seed(3)
h0 = torch.randn((B * T, 1024), device=dev, generator=g).to(torch.bfloat16)
LAYERS = int(os.environ.get("LAYERS", "12"))

def run():
    outs = []
    h = h0
    for L in net.w["layers"][:LAYERS]:
        hin = h
        a = ops.layernorm(h, *L["ln1"], 1e-5)
        qkv = ops.linear(a, L["wqkv"], L["bqkv"])
        att = ops.attention(qkv, B, T, 16, 64, 0.125)
        h1 = ops.linear(att, L["wo"], L["bo"], residual=h)
        f = ops.layernorm(h1, *L["ln2"], 1e-5)
        f1 = ops.linear(f, L["w1"], L["b1"], act=ops.ACT_GELU)
        h = ops.linear(f1, L["w2"], L["b2"], residual=h1)
        outs += [("ln1", a, None), ("qkv", qkv, None), ("att", att, None), ("o+res", h1, hin), ("ln2", f, None), ("ff1", f1, None), ("ff2+res", h, h1)]
    return outs

ref = run()
torch.cuda.synchronize()
for r in range(int(os.environ.get("RUNS", "12"))):
    outs = run()
    torch.cuda.synchronize()
    msg = "identical"
    for i, ((k, a, res), (_, b, _r)) in enumerate(zip(ref, outs)):
        d = (a.float() - b.float()).abs()
        if bool((d > 0).any()):
            nz = (d > 0).nonzero()
            rows = nz[:, 0].unique().tolist()
            cols = nz[:, 1].unique().tolist()
            msg = (f"layer {i // 7} op {k}: {nz.shape[0]} values differ; rows {rows[:12]}{'...' if len(rows) > 12 else ''} (n={len(rows)}) "
                   f"cols {cols[0]}..{cols[-1]} (n={len(cols)}) max {float(d.max()):.4g}")
            if res is not None:
                good, bad, rr = a.float(), b.float(), res.float()
                for (ri, ci) in nz[:: max(1, nz.shape[0] // 10)][:10].tolist():
                    msg += (f"\n      [{ri},{ci}] good {float(good[ri, ci]):8.4f} bad {float(bad[ri, ci]):8.4f} residual {float(rr[ri, ci]):8.4f} "
                            f"good-res {float(good[ri, ci] - rr[ri, ci]):8.4f}  bad-(good-res) {float(bad[ri, ci] - good[ri, ci] + rr[ri, ci]):8.4f}"
                            f"  res[r-128] {float(rr[ri - 128, ci]):8.4f} res[r-256] {float(rr[ri - 256, ci]):8.4f} res[r-512] {float(rr[max(ri - 512, 0), ci]):8.4f}")
            break
    print(f"run {r}: {msg}")
