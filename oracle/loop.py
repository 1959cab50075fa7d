"""ORACLE (test infrastructure, never shipped on the product path).

The reference's OWN execution shape on a CPU: one frame / one window at a time, batch 1, JPEG decode and PIL
preprocessing inside the loop -- what `bench.py` times as the "reference CPU path" (`cpu_baseline.batch1_loop`)
beside the batched port of oracle/video.py / oracle/audio.py (the "fair" CPU figure).

  * per-frame video loop  -- src/get_prob_video.py:77-187 (listdir membership test, cv2.imread, BGR->RGB, PIL NEAREST
                             resize + PILToTensor + channel flip + mean subtraction (src/data/utils.py:19-39), VS forward
                             at batch 1, the 10-slot feature window and the VD forward at batch 1 on every `step`-th frame,
                             carry-forward / gap rules)
  * per-window audio loop -- src/get_prob_audio_8_cl.py:68-101 (pad, HF zero-mean / unit-variance, model at batch 1,
                             replication of the window's logits to the frame ids it covers)

Same arithmetic as the batched restatements (tests/test_oracle_golden.py::test_batch1_loops_match_batched_oracle pins the
two against each other; the batched ones are pinned against the unmodified reference by oracle/make_golden.py).
"""
from __future__ import annotations

import os
from typing import Dict, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from . import audio as oa
from . import video as ov


def _pil_preprocess(bgr: np.ndarray) -> torch.Tensor:
    """What the reference does to one decoded crop (get_prob_video.py:96-99 + data/utils.py:19-39), with the same
    libraries: cv2 colour conversion, PIL NEAREST resize, CHW uint8 tensor, flip back to BGR, subtract the means."""
    import cv2
    from PIL import Image

    rgb = cv2.cvtColor(bgr, cv2.COLOR_BGR2RGB)
    img = Image.fromarray(rgb).resize((224, 224), Image.Resampling.NEAREST)
    chw = torch.from_numpy(np.asarray(img).copy()).permute(2, 0, 1).to(torch.float32)
    chw = torch.flip(chw, dims=(0,))
    for c, m in enumerate(ov.MEAN_BGR):
        chw[c] -= m
    return chw[None]


def video_loop(path_images: str, fps: float, total_frames: int, sd_vs: Dict[str, torch.Tensor],
               sd_vd: Dict[str, torch.Tensor]) -> Tuple[np.ndarray, np.ndarray]:
    """Frame-at-a-time restatement of preprocess_video_and_predict: returns (dyn [N,7], stat [N,7]) like the DataFrames'
    values (float64 when a zero row was appended)."""
    import cv2

    step = ov.vd_step(fps)
    folder = os.path.join(path_images, "00")
    present = os.listdir(folder)                       # a list: the reference pays the O(N) membership test per frame
    window = []                                        # up to 10 feature rows [1,512]
    last_vd = None
    stat_rows, dyn_rows = [], []
    zero = np.zeros(7)
    for i in range(total_frames):
        name = str(i).zfill(6) + ".jpg"
        if name in present:
            x = _pil_preprocess(cv2.imread(os.path.join(folder, name)))
            logits, feat = ov.resnet50_forward(sd_vs, x)
            stat_rows.append(F.softmax(logits, dim=1).numpy()[0])
            if i % step == 0:
                f = F.relu(feat).numpy()
                window = [f] * 10 if not window else window[1:] + [f]
                last_vd = ov.lstm_forward(sd_vd, torch.from_numpy(np.vstack(window))[None]).numpy()[0]
            dyn_rows.append(last_vd if last_vd is not None else zero)
        else:
            window = []
            if last_vd is not None:
                stat_rows.append(stat_rows[-1])
                dyn_rows.append(dyn_rows[-1])
            else:
                stat_rows.append(zero)
                dyn_rows.append(zero)
    return np.array(dyn_rows), np.array(stat_rows)


def audio_loop(wav: np.ndarray, fps: float, sd: Dict[str, torch.Tensor], step: float = 0.5, window: int = 4, sr: int = 16000,
               padding: str = "mean"):
    """Window-at-a-time restatement of load_audio_features: (rows [R, ncls], frame ids [R])."""
    rows, ids = [], []
    for (s, e, lo, hi) in oa.window_schedule(len(wav), fps, step, window, sr):
        x = oa.zero_mean_unit_var(oa.pad_window(wav[s:e], window * sr, padding))
        logit = oa.audio_model_forward(sd, torch.from_numpy(x[None])).numpy()[0]
        for f in range(lo, hi):
            rows.append(logit)
            ids.append(f)
    return np.asarray(rows, dtype=np.float32), np.asarray(ids, dtype=np.int64)
