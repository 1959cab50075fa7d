"""GPU probe: interleaved A/B of the branch-overlap settings of Engine.run_clips (same process, same clocks).
OVERLAPS="none;0,0;74,74;88,60" CLIPS=8 python scripts/ab_overlap.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from avcer_b200 import get_weights_matrices as gwm, synthetic as syn
from avcer_b200.pipeline import Engine

dev = "cuda:0"
c = int(os.environ.get("CLIPS", "8"))
n_frames, n_samples = 1500, 960000
cfgs = [None if t == "none" else tuple(int(v) for v in t.split(",")) for t in os.environ.get("OVERLAPS", "none;0,0;74,74;88,60").split(";")]
sds = (syn.make_vs_state_dict(0, "default"), syn.make_vd_state_dict(1), syn.make_audio_state_dict(2, 8, "spread", 12))
engs = [Engine(*sds, precision="bf16", device=dev, vs_batch=int(os.environ.get("VSB", "1024")), a_batch=int(os.environ.get("AB", "64")), overlap=o)
        for o in cfgs]
w1, w2 = gwm.class_weights(gwm.weights_3), [1, 1, 1]
g = torch.Generator(device=dev).manual_seed(1000)
crops = torch.randint(0, 256, (c * n_frames, 224, 224, 3), dtype=torch.uint8, device=dev, generator=g)
wav = (torch.randn(c * n_samples, device=dev, generator=g) * 0.1).contiguous()
exists = [np.ones(n_frames, dtype=bool) for _ in range(c)]
args = (crops, exists, [25.0] * c, wav, [n_samples] * c, w1, w2, False, True)
ref = None
for e in engs:
    for _ in range(2):
        out = e.run_clips(*args)
    torch.cuda.synchronize()
    lab = out["labels"].cpu()
    if ref is None:
        ref = (lab, out["window_logits"].cpu(), out["stat"].cpu())
    else:
        print("overlap", e.overlap, "labels equal:", bool((lab == ref[0]).all()), "logits equal:", bool(torch.equal(torch.nan_to_num(out["window_logits"].cpu(), nan=-7.0), torch.nan_to_num(ref[1], nan=-7.0))),
              "stat equal:", bool((out["stat"].cpu() == ref[2]).all()),
              "logits maxdiff:", float(torch.nan_to_num(out["window_logits"].cpu() - ref[1]).abs().max()),
              "nan pattern equal:", bool((torch.isnan(out["window_logits"].cpu()) == torch.isnan(ref[1])).all()))
res = {k: [] for k in range(len(cfgs))}
reps = int(os.environ.get("REPS", "3"))
for rnd in range(4):
    for k, e in enumerate(engs):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            e.run_clips(*args)
        b.record(); torch.cuda.synchronize()
        res[k].append(a.elapsed_time(b) / reps)
for k, o in enumerate(cfgs):
    ms = sorted(res[k])[len(res[k]) // 2]
    print(f"overlap={str(o):10s}: {ms:.2f} ms/step  {c * n_frames / ms:.1f} kframes/s   all: {[round(x, 1) for x in res[k]]}")
