"""GPU probe: per-kernel device time of one full bench-like step (Engine.run_clips, graph replay) via torch.profiler."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
from avcer_b200 import get_weights_matrices as gwm, synthetic as syn
from avcer_b200.pipeline import Engine

dev = "cuda:0"
c = int(os.environ.get("CLIPS", "2"))
n_frames, n_samples = 1500, 960000
eng = Engine(syn.make_vs_state_dict(0, "default"), syn.make_vd_state_dict(1), syn.make_audio_state_dict(2, 8, "spread", 12),
             precision="bf16", device=dev, vs_batch=256, a_batch=64)
w1, w2 = gwm.class_weights(gwm.weights_3), [1, 1, 1]
g = torch.Generator(device=dev).manual_seed(1000)
crops = torch.randint(0, 256, (c * n_frames, 224, 224, 3), dtype=torch.uint8, device=dev, generator=g)
wav = (torch.randn(c * n_samples, device=dev, generator=g) * 0.1).contiguous()
exists = [np.ones(n_frames, dtype=bool) for _ in range(c)]
step = lambda: eng.run_clips(crops, exists, [25.0] * c, wav, [n_samples] * c, w1, w2, False, True)
for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); step(); e1.record(); torch.cuda.synchronize()
print(f"step (events): {e0.elapsed_time(e1):.2f} ms")
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
rows = [(e.key, getattr(e, "device_time_total", 0) or getattr(e, "cuda_time_total", 0), e.count) for e in prof.key_averages()]
rows = [r for r in rows if r[1] > 0]
tot = sum(r[1] for r in rows)
print(f"sum of kernel device time {tot / 1e3:.2f} ms")
for k, t, n in sorted(rows, key=lambda r: -r[1])[:24]:
    print(f"  {t / 1e3:8.3f} ms {100 * t / tot:5.1f}%  n={n:5d}  {k[:100]}")
