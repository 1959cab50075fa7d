"""GPU probe: time of the VS branch and of the audio branch alone when their persistent kernels are limited to a share
of the SMs (avcer_set_sm_limit), to see which branch is bandwidth- and which is SM-bound."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from avcer_b200 import synthetic as syn
from avcer_b200.pipeline import Engine

dev = "cuda:0"
c = int(os.environ.get("CLIPS", "4"))
n_frames, n_samples = 1500, 960000
sds = (syn.make_vs_state_dict(0, "default"), syn.make_vd_state_dict(1), syn.make_audio_state_dict(2, 8, "spread", 12))
g = torch.Generator(device=dev).manual_seed(1000)
crops = torch.randint(0, 256, (c * n_frames, 224, 224, 3), dtype=torch.uint8, device=dev, generator=g)
wav = (torch.randn(c * n_samples, device=dev, generator=g) * 0.1).contiguous()
ref = None
for lim in [int(v) for v in os.environ.get("LIMITS", "0,0,112,74,48").split(",")]:
    eng = Engine(*sds, precision="bf16", device=dev, vs_batch=1024, a_batch=64, overlap=(lim, lim) if lim else None)
    def vs():
        return eng.vs_forward_u8(crops)
    def au():
        return eng.audio_rows(wav, [n_samples] * c, [25.0] * c, [n_frames] * c)
    out = {}
    for name, fn in (("vs", vs), ("audio", au)):
        for _ in range(2):
            r = fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            r = fn()
        b.record(); torch.cuda.synchronize()
        out[name] = (a.elapsed_time(b) / 3, r)
    logits = out["audio"][1][1].cpu()
    probs = out["vs"][1][0].cpu()
    if ref is None:
        ref = (logits, probs)
    print(f"limit {lim:3d}: vs {out['vs'][0]:7.2f} ms  audio {out['audio'][0]:7.2f} ms   logits==first {bool(torch.equal(torch.nan_to_num(logits, nan=-7.0), torch.nan_to_num(ref[0], nan=-7.0)))} "
          f"maxdiff {float(torch.nan_to_num(logits - ref[0]).abs().max()):.3g}  probs==first {bool((probs == ref[1]).all())}")
    del eng
    torch.cuda.empty_cache()
