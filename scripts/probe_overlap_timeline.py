"""GPU probe: do the VS branch and the audio branch really run side by side?  Per-stream begin / end times against a
common origin, per VS batch and per audio batch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from avcer_b200 import ops, synthetic as syn
from avcer_b200.pipeline import Engine

dev = "cuda:0"
c = int(os.environ.get("CLIPS", "4"))
lv, la = [int(v) for v in os.environ.get("SPLIT", "74,74").split(",")]
n_frames, n_samples = 1500, 960000
sds = (syn.make_vs_state_dict(0, "default"), syn.make_vd_state_dict(1), syn.make_audio_state_dict(2, 8, "spread", 12))
eng = Engine(*sds, precision="bf16", device=dev, vs_batch=1024, a_batch=64, overlap=(lv, la))
g = torch.Generator(device=dev).manual_seed(1000)
crops = torch.randint(0, 256, (c * n_frames, 224, 224, 3), dtype=torch.uint8, device=dev, generator=g)
wav = (torch.randn(c * n_samples, device=dev, generator=g) * 0.1).contiguous()
for _ in range(2):
    eng.vs_forward_u8(crops); eng.audio_rows(wav, [n_samples] * c, [25.0] * c, [n_frames] * c)
torch.cuda.synchronize()
cur = torch.cuda.current_stream()
side = eng._a_stream
ev = lambda: torch.cuda.Event(enable_timing=True)
t0 = ev(); t0.record(cur)
side.wait_stream(cur)
marks = []
# audio batches on the side stream
xin = torch.empty((64, 64000), device=dev)
with torch.cuda.stream(side):
    for i in range(int(os.environ.get("NA", "8"))):
        a = ev(); a.record(side)
        eng._a_fwd(xin)
        b = ev(); b.record(side)
        marks.append(("A", i, a, b))
x = eng._vs_input(1024)
for i in range(int(os.environ.get("NV", "6"))):
    a = ev(); a.record(cur)
    eng._vs_fwd(x)
    b = ev(); b.record(cur)
    marks.append(("VS", i, a, b))
torch.cuda.synchronize()
for k, i, a, b in marks:
    print(f"{k:3s} {i:2d}: {t0.elapsed_time(a):8.2f} -> {t0.elapsed_time(b):8.2f} ms  ({a.elapsed_time(b):6.2f})")
