"""GPU probe: run the audio forward several times on the same input and report the first tap that differs between
runs (a race shows up as run-to-run differences; every kernel is meant to be deterministic)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avcer_b200 import nets, synthetic as syn

dev = "cuda:0"
B = int(os.environ.get("B", "64"))
net = nets.ANet(syn.make_audio_state_dict(2, 8, "spread", 12), "bf16", dev)
g = torch.Generator(device=dev).manual_seed(3)
x = torch.randn((B, 64000), device=dev, generator=g)
junk = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
runs = []
for r in range(int(os.environ.get("RUNS", "4"))):
    taps = {}
    junk.fill_((r * 37 + 1) % 251)            # different stale bytes in freshly allocated buffers? (allocator reuse)
    out = net.forward(x, taps)
    torch.cuda.synchronize()
    taps["logits"] = out
    runs.append({k: v.float().cpu().clone() for k, v in taps.items()})
for r in range(1, len(runs)):
    for k in runs[0]:
        d = (runs[r][k] - runs[0][k]).abs()
        nbad = int((d > 0).sum())
        if nbad:
            idx = (d > 0).nonzero()[:5].tolist()
            print(f"run {r}: tap {k:8s} differs in {nbad} of {d.numel()} values, max {float(d.max()):.4g}, first at {idx}")
            break
    else:
        print(f"run {r}: identical")
