"""GPU probe: interleaved A/B of Engine batch sizes inside one process (same clocks / thermal state)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from avcer_b200 import get_weights_matrices as gwm, synthetic as syn
from avcer_b200.pipeline import Engine

dev = "cuda:0"
c = int(os.environ.get("CLIPS", "4"))
n_frames, n_samples = 1500, 960000
cfgs = [tuple(int(v) for v in t.split(",")) for t in os.environ.get("CFGS", "256,64;512,64;512,128;1024,64").split(";")]
sds = (syn.make_vs_state_dict(0, "default"), syn.make_vd_state_dict(1), syn.make_audio_state_dict(2, 8, "spread", 12))
engs = [Engine(*sds, precision="bf16", device=dev, vs_batch=v, a_batch=a) for v, a in cfgs]
w1, w2 = gwm.class_weights(gwm.weights_3), [1, 1, 1]
g = torch.Generator(device=dev).manual_seed(1000)
crops = torch.randint(0, 256, (c * n_frames, 224, 224, 3), dtype=torch.uint8, device=dev, generator=g)
wav = (torch.randn(c * n_samples, device=dev, generator=g) * 0.1).contiguous()
exists = [np.ones(n_frames, dtype=bool) for _ in range(c)]
args = (crops, exists, [25.0] * c, wav, [n_samples] * c, w1, w2, False, True)
for e in engs:
    for _ in range(2):
        e.run_clips(*args)
torch.cuda.synchronize()
res = {k: [] for k in cfgs}
for rnd in range(4):
    for k, e in zip(cfgs, engs):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(2):
            e.run_clips(*args)
        b.record(); torch.cuda.synchronize()
        res[k].append(a.elapsed_time(b) / 2)
for k in cfgs:
    ms = sorted(res[k])[len(res[k]) // 2]
    print(f"vs_batch={k[0]:4d} a_batch={k[1]:3d}: {ms:.2f} ms/step  {c * n_frames / ms:.1f} kframes/s   all: {[round(x, 1) for x in res[k]]}")
