#!/usr/bin/env python
"""GPU probe: batched baseline-JPEG decode (avcer_jpeg_decode) of 224x224 crops: host header parse time, device time per
kernel (CUDA events), against cv2.imdecode on the host cores."""
import os
import sys
import time

import cv2
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from avcer_b200 import jpeg, synthetic as syn      # noqa: E402

n = int(os.environ.get("N", "1500"))
base = syn.make_crops(3, 50)
files = [cv2.imencode(".jpg", base[i % 50])[1].tobytes() for i in range(n)]
print(f"{n} crops 224x224, {sum(map(len, files)) / n / 1e3:.1f} KB per file")
t0 = time.perf_counter()
for f in files:
    jpeg.parse(f)
t_parse = time.perf_counter() - t0
t0 = time.perf_counter()
for f in files[:200]:
    cv2.imdecode(np.frombuffer(f, np.uint8), cv2.IMREAD_COLOR)
t_cv = (time.perf_counter() - t0) / 200
jpeg.decode_batch(files, "cuda:0")
torch.cuda.synchronize()
ts = []
for _ in range(5):
    jpeg.PROFILE = []
    t0 = time.perf_counter()
    jpeg.decode_batch(files, "cuda:0")
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3
    ts.append((sum(a.elapsed_time(b) for a, b in jpeg.PROFILE), wall))
jpeg.PROFILE = None
dev_ms, wall_ms = min(t[0] for t in ts), min(t[1] for t in ts)
print(f"host: header parse {t_parse / n * 1e6:.0f} us per file; cv2.imdecode {t_cv * 1e6:.0f} us per file (one core, {os.cpu_count()} cores on the box)")
print(f"decode_batch: {wall_ms:.1f} ms wall ({n / wall_ms * 1e3:.0f} crops/s); the four kernels: {dev_ms:.2f} ms ({n / dev_ms * 1e3:.0f} crops/s), "
      f"output {n * 150528 / dev_ms / 1e6:.1f} GB/s")
