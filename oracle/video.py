"""ORACLE (test infrastructure, never shipped on the product path).

CPU restatement (numpy + torch.nn.functional in fp32) of the visual branch of ElenaRyumina/AVCER:
  * face-crop preprocessing        -- src/data/utils.py:19-39 (pth_processing) incl. Pillow's
                                      NEAREST index rule (Pillow Geometry.c affine scale loop)
  * VS ResNet-50 forward           -- src/architectures/video.py:7-166
  * VD LSTM forward                -- src/architectures/video.py:169-185
  * per-frame sampling / windows / gap handling -- src/get_prob_video.py:67-187

Pinned against the reference's own classes and functions by oracle/make_golden.py ->
tests/golden/video_*.npz.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

MEAN_BGR = (91.4953, 103.8827, 131.0912)   # utils.py:27-29
BN_EPS = 1e-3                               # video.py: every BatchNorm2d(eps=0.001)
VS_BLOCKS = (3, 4, 6, 3)


# ------------------------------------------------------------------------------------------ preprocessing
def nearest_index_table(n_in: int, n_out: int = 224) -> np.ndarray:
    """Pillow NEAREST resize source indices: a = in/out (double); o = a*0.5; idx[x] = int(o); o += a."""
    a = n_in / float(n_out)
    o = a * 0.5
    idx = np.empty(n_out, dtype=np.int64)
    for x in range(n_out):
        idx[x] = int(o)
        o += a
    return np.minimum(idx, n_in - 1)


def pth_processing(img_bgr: np.ndarray) -> np.ndarray:
    """uint8 HxWx3 in cv2.imread (BGR) order -> float32 [1,3,224,224] (utils.py:19-39).

    The reference converts BGR->RGB (get_prob_video.py:97), resizes with PIL NEAREST, makes a CHW
    tensor and flips the channel axis back to BGR before subtracting the per-channel means, so the
    result is simply the BGR bytes, nearest-resized, minus MEAN_BGR.
    """
    h, w, _ = img_bgr.shape
    ys, xs = nearest_index_table(h), nearest_index_table(w)
    small = img_bgr[ys][:, xs].astype(np.float32)            # [224,224,3]
    chw = np.ascontiguousarray(small.transpose(2, 0, 1))
    for c in range(3):
        chw[c] -= np.float32(MEAN_BGR[c])
    return chw[None]


# ------------------------------------------------------------------------------------------ VS: ResNet-50
def _bn(x: torch.Tensor, sd: Dict[str, torch.Tensor], p: str) -> torch.Tensor:
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                        training=False, eps=BN_EPS)


def _bottleneck(x: torch.Tensor, sd, p: str, stride: int, has_ds: bool) -> torch.Tensor:
    identity = x
    y = F.relu(_bn(F.conv2d(x, sd[p + ".conv1.weight"], stride=stride), sd, p + ".batch_norm1"))   # stride on conv1
    y = F.relu(_bn(F.conv2d(y, sd[p + ".conv2.weight"], padding=1), sd, p + ".batch_norm2"))       # 3x3 "same"
    y = _bn(F.conv2d(y, sd[p + ".conv3.weight"]), sd, p + ".batch_norm3")
    if has_ds:
        identity = _bn(F.conv2d(x, sd[p + ".i_downsample.0.weight"], stride=stride), sd, p + ".i_downsample.1")
    return F.relu(y + identity)


def resnet50_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, taps: Optional[dict] = None
                     ) -> Tuple[torch.Tensor, torch.Tensor]:
    """x: [B,3,224,224] fp32 -> (logits [B,7], fc1 pre-activation [B,512]) (video.py:115-133)."""
    with torch.no_grad():
        # Conv2dSame, 7x7 stride 2: TF "same" padding = 5 -> 2 before, 3 after (video.py:65-81)
        y = F.conv2d(F.pad(x, [2, 3, 2, 3]), sd["conv_layer_s2_same.weight"], stride=2)
        y = F.relu(_bn(y, sd, "batch_norm1"))
        if taps is not None:
            taps["stem"] = y
        y = F.max_pool2d(y, kernel_size=3, stride=2)                       # no padding: 112 -> 55
        if taps is not None:
            taps["pool"] = y
        for li, blocks in enumerate(VS_BLOCKS, start=1):
            for b in range(blocks):
                y = _bottleneck(y, sd, f"layer{li}.{b}", stride=(2 if (li > 1 and b == 0) else 1), has_ds=(b == 0))
            if taps is not None:
                taps[f"layer{li}"] = y
        y = y.mean(dim=(2, 3))
        feat = F.linear(y, sd["fc1.weight"], sd["fc1.bias"])
        logits = F.linear(F.relu(feat), sd["fc2.weight"], sd["fc2.bias"])
    return logits, feat


# ------------------------------------------------------------------------------------------ VD: LSTM
def _lstm_layer(x: torch.Tensor, w_ih, w_hh, b_ih, b_hh) -> torch.Tensor:
    """batch_first single-layer LSTM from zero state, PyTorch gate order i,f,g,o."""
    bsz, steps, _ = x.shape
    hid = w_hh.shape[1]
    h = x.new_zeros(bsz, hid)
    c = x.new_zeros(bsz, hid)
    outs = []
    for t in range(steps):
        g = x[:, t] @ w_ih.t() + b_ih + h @ w_hh.t() + b_hh
        i, f, gg, o = g.split(hid, dim=1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
        h = torch.sigmoid(o) * torch.tanh(c)
        outs.append(h)
    return torch.stack(outs, dim=1)


def lstm_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor) -> torch.Tensor:
    """x: [M,10,512] -> logits [M,7] (video.py:181-185)."""
    with torch.no_grad():
        y = _lstm_layer(x, sd["lstm1.weight_ih_l0"], sd["lstm1.weight_hh_l0"], sd["lstm1.bias_ih_l0"], sd["lstm1.bias_hh_l0"])
        y = _lstm_layer(y, sd["lstm2.weight_ih_l0"], sd["lstm2.weight_hh_l0"], sd["lstm2.bias_ih_l0"], sd["lstm2.bias_hh_l0"])
        return F.linear(y[:, -1], sd["fc.weight"], sd["fc.bias"])


# ------------------------------------------------------------------------------------------ frame bookkeeping
def vd_step(fps: float) -> int:
    return round((5 * fps) / 25)          # get_prob_video.py:77 (Python banker's rounding)


def plan_video(exists: Sequence[bool], step: int):
    """Index plan equivalent to the frame loop of get_prob_video.py:91-178.

    Returns
      samples  : frame indices that feed the LSTM (existing frames with i % step == 0), in order
      windows  : int array [len(samples), 10] of positions into `samples` (the 10-slot window,
                 first feature repeated on the left, reset after every missing frame)
      stat_src : per frame, the frame index whose VS row it shows (-1 = zeros)
      dyn_src  : per frame, the position into `samples` whose VD row it shows (-1 = zeros)
    """
    n = len(exists)
    samples: List[int] = []
    windows: List[List[int]] = []
    stat_src = np.full(n, -1, dtype=np.int64)
    dyn_src = np.full(n, -1, dtype=np.int64)
    cur: List[int] = []          # sliding window (positions into samples)
    last = -1                    # position of the latest LSTM output
    for i in range(n):
        if exists[i]:
            stat_src[i] = i
            if i % step == 0:
                samples.append(i)
                pos = len(samples) - 1
                cur = [pos] * 10 if not cur else cur[1:] + [pos]
                windows.append(list(cur))
                last = pos
            dyn_src[i] = last
        else:
            cur = []
            if last >= 0 and i > 0:
                stat_src[i] = stat_src[i - 1]
                dyn_src[i] = dyn_src[i - 1]
            # else: zeros for both rows (get_prob_video.py:175-178)
    return np.asarray(samples, dtype=np.int64), np.asarray(windows, dtype=np.int64).reshape(-1, 10), stat_src, dyn_src


def predict_video(frames: Sequence[Optional[np.ndarray]], fps: float, sd_vs, sd_vd, batch: int = 32):
    """frames[i] is the uint8 BGR crop of frame i or None when the face crop is missing.
    Returns (dyn [N,7] VD logits, stat [N,7] VS probabilities) as the reference DataFrames' values
    (columns in VIDEO_ORDER); dtype float64 when a zero row occurs, else float32 (np.array promotion,
    get_prob_video.py:89,182-187)."""
    exists = [f is not None for f in frames]
    step = vd_step(fps)
    samples, windows, stat_src, dyn_src = plan_video(exists, step)
    idx = [i for i, e in enumerate(exists) if e]
    probs = {}
    feats = {}
    for s in range(0, len(idx), batch):
        chunk = idx[s:s + batch]
        x = torch.from_numpy(np.concatenate([pth_processing(frames[i]) for i in chunk], axis=0))
        logits, feat = resnet50_forward(sd_vs, x)
        p = F.softmax(logits, dim=1).numpy()
        fr = F.relu(feat).numpy()
        for j, i in enumerate(chunk):
            probs[i] = p[j]
            feats[i] = fr[j]
    if len(samples):
        fm = np.stack([feats[i] for i in samples])
        vd = lstm_forward(sd_vd, torch.from_numpy(fm[windows])).numpy()
    else:
        vd = np.zeros((0, 7), dtype=np.float32)
    has_zero = bool((stat_src < 0).any() or (dyn_src < 0).any())
    dt = np.float64 if has_zero else np.float32
    n = len(frames)
    stat = np.zeros((n, 7), dtype=dt)
    dyn = np.zeros((n, 7), dtype=dt)
    for i in range(n):
        if stat_src[i] >= 0:
            stat[i] = probs[int(stat_src[i])]
        if dyn_src[i] >= 0:
            dyn[i] = vd[int(dyn_src[i])]
    return dyn, stat
