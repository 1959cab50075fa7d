"""Drop-in for the hot-path helpers of the reference's src/data/utils.py (same names, argument
meaning and error behaviour); the arithmetic runs in libavcer_b200 kernels on the GPU.

  pth_processing          utils.py:19-39     -> K1 (avcer_preprocess_u8, layout 0)
  pad_wav / pad_wav_zeros utils.py:63-89     -> kept as tensor helpers (the batched path pads inside K5a)
  softmax                 utils.py:125-127   -> avcer_softmax7
  get_compound_expression utils.py:222-241   -> avcer_compound_scores
  get_image_location      utils.py:244-247   (string helper)
  save_txt                utils.py:212-219   (file helper)
  convert_mp4_to_mp3      utils.py:42-60     decode is upstream of the path:
                                             reads a 16-bit PCM .wav next to the video, resamples to `sampling_rate`
"""
from __future__ import annotations

import ctypes
import os
import wave

import numpy as np
import torch

from .. import _lib, config, ops


def pth_processing(fp) -> torch.Tensor:
    """fp: PIL RGB image (as built at get_prob_video.py:97-99).  Returns float32 [1,3,224,224]."""
    rgb = np.asarray(fp.convert("RGB") if hasattr(fp, "convert") else fp, dtype=np.uint8)
    bgr = np.ascontiguousarray(rgb[:, :, ::-1])
    dev = config.device()
    h, w, _ = bgr.shape
    src = torch.from_numpy(bgr.reshape(-1)).to(dev)
    out = torch.empty((1, 3, 224, 224), device=dev, dtype=torch.float32)
    ops.preprocess(src, 1, out, 0, offsets=torch.zeros(1, dtype=torch.int64, device=dev),
                   heights=torch.tensor([h], dtype=torch.int32, device=dev), widths=torch.tensor([w], dtype=torch.int32, device=dev))
    return out


def convert_mp4_to_mp3(path, sampling_rate=16000):
    path_save = path[:-3] + "wav"
    if not os.path.exists(path_save):
        raise FileNotFoundError(f"{path_save}: audio decode (ffmpeg) is outside the accelerated path; provide the .wav")
    with wave.open(path_save, "rb") as f:
        sr, nch, sw, n = f.getframerate(), f.getnchannels(), f.getsampwidth(), f.getnframes()
        assert sw == 2, "16-bit PCM expected"
        pcm = np.frombuffer(f.readframes(n), dtype="<i2").reshape(-1, nch).T.astype(np.float32) / 32768.0
    wav = torch.from_numpy(pcm)
    if wav.size(0) > 1:
        wav = wav.mean(dim=0, keepdim=True)
    if sr != sampling_rate:
        import torchaudio

        wav = torchaudio.transforms.Resample(orig_freq=sr, new_freq=sampling_rate)(wav)
        sr = sampling_rate
    assert sr == sampling_rate
    return wav.squeeze(0)


def pad_wav(wav, max_length):
    current_length = len(wav)                        # ZeroDivisionError below for an empty chunk, as in the reference
    if current_length < max_length:
        repetitions = (max_length + current_length - 1) // current_length
        return torch.cat([wav] * repetitions, dim=0)[:max_length]
    return wav[:max_length]


def pad_wav_zeros(wav, max_length, mode="constant"):
    missing = max(0, max_length - wav.shape[0])
    if mode == "mean":
        return torch.nn.functional.pad(wav, (0, missing), mode="constant", value=torch.mean(wav))
    return torch.nn.functional.pad(wav, (0, missing), mode=mode)


def softmax(matrix):
    m = np.ascontiguousarray(matrix)
    if m.ndim != 2 or m.shape[1] != 7 or m.dtype not in (np.float32, np.float64):
        raise ValueError("avcer_b200.data.utils.softmax handles [n,7] float32/float64 matrices")
    return ops.softmax7(torch.from_numpy(m).to(config.device())).cpu().numpy()


def get_compound_expression(pred, com_emo, dict_weights, ce_weights_type, ce_mask):
    pred = np.ascontiguousarray(np.asarray(pred))
    if pred.dtype not in (np.float32, np.float64):
        pred = pred.astype(np.float64)
    pairs, w = [], []
    for _, v in com_emo.items():
        i1, i2 = v[0], v[1]
        if ce_weights_type:
            s = dict_weights[i1] + dict_weights[i2]
            w += [dict_weights[i1] / s, dict_weights[i2] / s]
        else:
            w += [1.0, 1.0]
        pairs += [i1, i2]
    k = len(pairs) // 2
    n, ncols = pred.shape
    dev = config.device()
    out = torch.empty((n, k), device=dev, dtype=torch.float64)
    pa = (ctypes.c_int32 * len(pairs))(*pairs)
    wa = (ctypes.c_double * len(w))(*w)
    x = torch.from_numpy(pred).to(dev)
    _lib.check(_lib.load().avcer_compound_scores(x.data_ptr(), n, ncols, int(pred.dtype == np.float64), pa, wa, k,
                                                 int(bool(ce_mask)), out.data_ptr(), torch.cuda.current_stream().cuda_stream))
    return out.cpu().numpy()


def save_txt(column_names, file_names, labels, save_name):
    with open(save_name, "w") as file:
        file.write(",".join(column_names) + "\n")
        for file_name, label in zip(file_names, labels):
            file.write(f"{file_name},{label}\n")


def get_image_location(curr_video, frame):
    frame = int(frame.split(".")[0]) + 1
    return f"{curr_video}/{str(frame).zfill(5)}.jpg"
