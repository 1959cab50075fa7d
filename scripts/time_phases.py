"""GPU probe: CUDA-event time of each phase of Engine.run_clips (VS loop, VD + row gathers, A loop, fusion)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from avcer_b200 import get_weights_matrices as gwm, synthetic as syn
from avcer_b200.pipeline import Engine

dev = "cuda:0"
c = int(os.environ.get("CLIPS", "2"))
n_frames, n_samples = 1500, 960000
eng = Engine(syn.make_vs_state_dict(0, "default"), syn.make_vd_state_dict(1), syn.make_audio_state_dict(2, 8, "spread", 12),
             precision="bf16", device=dev, vs_batch=int(os.environ.get("VSB", "256")), a_batch=int(os.environ.get("AB", "64")))
w1, w2 = gwm.class_weights(gwm.weights_3), [1, 1, 1]
g = torch.Generator(device=dev).manual_seed(1000)
crops = torch.randint(0, 256, (c * n_frames, 224, 224, 3), dtype=torch.uint8, device=dev, generator=g)
wav = (torch.randn(c * n_samples, device=dev, generator=g) * 0.1).contiguous()
exists = [np.ones(n_frames, dtype=bool) for _ in range(c)]
fps = [25.0] * c


def step(ev=None):
    def mark(k):
        if ev is not None:
            e = torch.cuda.Event(enable_timing=True); e.record(); ev.append((k, e))
    mark("start")
    probs, feats = eng.vs_forward_u8(crops); mark("vs")
    stat, dyn, plans = eng.video_rows(probs, feats, exists, fps); mark("vd+rows")
    a_rows, logits = eng.audio_rows(wav, [n_samples] * c, fps, [n_frames] * c); mark("audio")
    labels = eng.fuse(stat, dyn, a_rows, w1, w2, False, True); mark("fuse")
    return labels


for _ in range(3):
    step()
torch.cuda.synchronize()
ev = []
step(ev)
torch.cuda.synchronize()
tot = ev[0][1].elapsed_time(ev[-1][1])
print(f"clips={c}: step {tot:.2f} ms -> {c * n_frames / tot:.1f} kframes/s")
for (k0, e0), (k1, e1) in zip(ev[:-1], ev[1:]):
    print(f"  {k1:8s} {e0.elapsed_time(e1):8.3f} ms")
