#!/usr/bin/env python
"""GPU probe: where the wall time of the file-based drop-in call (get_prob_video.preprocess_video_and_predict on 1500 JPEG
crops) goes: host phases timed with perf_counter, GPU drained at the marked points."""
import os
import sys
import tempfile
import time

import cv2
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from avcer_b200 import config, get_prob_video as gpv, jpeg, synthetic as syn      # noqa: E402

with tempfile.TemporaryDirectory(prefix="avcer_prof_") as td:
    os.makedirs(os.path.join(td, "clip", "00"))
    base = syn.make_crops(5, 50)
    for i in range(1500):
        cv2.imwrite(os.path.join(td, "clip", "00", f"{i:06d}.jpg"), base[i % 50])
    config.set_state_dicts(vs=syn.make_vs_state_dict(0, "default"), vd=syn.make_vd_state_dict(1))
    for _ in range(2):
        gpv.preprocess_video_and_predict(path_images=os.path.join(td, "clip"), save_path=td, fps=25, total_frames=1500)
    eng = config.video_engine()
    for rep in range(2):
        torch.cuda.synchronize()
        t = [time.perf_counter()]
        paths, exists = gpv._present_frames(os.path.join(td, "clip"), 1500)
        t.append(time.perf_counter())
        files = [open(p, "rb").read() for p in paths]
        t.append(time.perf_counter())
        flat, off, hs, ws, st = jpeg.decode_batch(files, eng.device, defer_status=True)
        t.append(time.perf_counter())
        torch.cuda.synchronize()
        t.append(time.perf_counter())
        probs, feats = eng.vs_forward_ragged(flat, off, hs, ws)
        t.append(time.perf_counter())
        torch.cuda.synchronize()
        t.append(time.perf_counter())
        stat, dyn, plans = eng.video_rows(probs, feats, [exists], [25])
        t.append(time.perf_counter())
        a, b = dyn.cpu().numpy(), stat.cpu().numpy()
        t.append(time.perf_counter())
        names = ["listdir + exists", "read 1500 files", "decode_batch host side", "  .. GPU decode drain", "vs_forward_ragged enqueue", "  .. GPU VS drain",
                 "video_rows (VD) enqueue", "D2H of both tables (drains VD)"]
        print(" | ".join(f"{n} {1e3 * (t[i + 1] - t[i]):.1f} ms" for i, n in enumerate(names)), f"| total {1e3 * (t[-1] - t[0]):.1f} ms", flush=True)
    t0 = time.perf_counter()
    gpv.preprocess_video_and_predict(path_images=os.path.join(td, "clip"), save_path=td, fps=25, total_frames=1500)
    print(f"whole call {1e3 * (time.perf_counter() - t0):.1f} ms")
