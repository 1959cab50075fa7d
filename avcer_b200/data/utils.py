"""Drop-in for the hot-path helpers of the reference's src/data/utils.py (same names, argument
meaning and error behaviour); the arithmetic runs in libavcer_b200 kernels on the GPU.

  pth_processing          utils.py:19-39     -> K1 (avcer_preprocess_u8, layout 0)
  pad_wav / pad_wav_zeros utils.py:63-89     -> kept as tensor helpers (the batched path pads inside K5a)
  softmax                 utils.py:125-127   -> avcer_softmax7
  get_compound_expression utils.py:222-241   -> avcer_compound_scores
  get_metrics_for_fusion  utils.py:115-122   (precision / F1 / UAR over classes 1..6, sklearn's formulas on a confusion matrix)
  get_weights_prob_model  utils.py:138-163   -> avcer_weight_search_confusion (all Dirichlet candidates in one launch)
  get_weights_v_model     utils.py:166-185   -> same kernel, candidates = weight grid x weight grid
  get_weights_av_model    utils.py:188-209   -> same kernel, candidates = grid^3
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
    """data/utils.py:42-60.  The ffmpeg call (container demux + codec decode) stays upstream: the 16-bit PCM .wav it
    writes next to the video must exist.  What follows it in the reference -- torchaudio.load's 1/32768 scaling, the
    channel mean and torchaudio.transforms.Resample -- runs in one kernel (avcer_pcm16_resample)."""
    path_save = path[:-3] + "wav"
    if not os.path.exists(path_save):
        raise FileNotFoundError(f"{path_save}: audio decode (ffmpeg) is outside the accelerated path; provide the .wav")
    with wave.open(path_save, "rb") as f:
        sr, nch, sw, n = f.getframerate(), f.getnchannels(), f.getsampwidth(), f.getnframes()
        if sw != 2:
            raise ValueError(f"{path_save}: 16-bit PCM expected, got {8 * sw}-bit samples")
        pcm = np.frombuffer(f.readframes(n), dtype="<i2").reshape(-1, nch)
    dev = config.device()
    return ops.pcm16_to_mono(torch.from_numpy(np.array(pcm)).to(dev), sr, sampling_rate).cpu()


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


# ------------------------------------------------------------------------------------------------ weight search
def _prf_from_confusion(cm: np.ndarray):
    """sklearn.metrics.precision_recall_fscore_support on confusion counts cm[..., true, pred] (zero_division -> 0):
    precision = tp / predicted, recall = tp / true, F1 = 2 tp / (true + predicted), all float64."""
    cm = cm.astype(np.float64)
    tp = np.diagonal(cm, axis1=-2, axis2=-1)
    pred_sum = cm.sum(axis=-2)
    true_sum = cm.sum(axis=-1)
    with np.errstate(divide="ignore", invalid="ignore"):
        precision = np.where(pred_sum == 0, 0.0, tp / pred_sum)
        recall = np.where(true_sum == 0, 0.0, tp / true_sum)
        denom = true_sum + pred_sum
        f1 = np.where(denom == 0, 0.0, (2.0 * tp) / denom)
    return precision, f1, recall, true_sum + pred_sum


def _fusion_metrics_from_confusion(cm: np.ndarray) -> np.ndarray:
    """[..., 3] = (precision, f1, uar) averaged over classes 1..6 exactly like get_metrics_for_fusion (utils.py:115-122):
    sequential accumulation over cl = 1..6, then / 6.  A class absent from both truth and prediction is a KeyError there."""
    precision, f1, recall, seen = _prf_from_confusion(cm)
    if np.any(seen[..., 1:7] == 0):
        missing = int(np.argwhere(seen[..., 1:7] == 0)[0][-1]) + 1
        raise KeyError(str(missing))
    out = np.zeros(cm.shape[:-2] + (3,))
    for cl in range(1, 7):
        out[..., 0] += precision[..., cl]
        out[..., 1] += f1[..., cl]
        out[..., 2] += recall[..., cl]
    return out / 6


def get_metrics_for_fusion(true, pred):
    true = np.asarray(true).astype(np.int64)
    pred = np.asarray(pred).astype(np.int64)
    k = int(max(true.max(initial=0), pred.max(initial=0), 6)) + 1
    cm = np.zeros((k, k), dtype=np.int64)
    np.add.at(cm, (true, pred), 1)
    precision, f1, uar = _fusion_metrics_from_confusion(cm)
    return precision, f1, uar


def _search_metrics(ground_truth, predictions, weights: np.ndarray) -> np.ndarray:
    """UAR-over-classes-1..6 (the reference's selection metric) of every candidate weight set [W, M, 7]."""
    gt = np.asarray(ground_truth).astype(np.int64)
    if gt.size and (gt.min() < 0 or gt.max() > 6):
        raise ValueError("ground truth labels must be in 0..6")
    preds = np.stack([np.asarray(p, dtype=np.float64) for p in predictions])
    dev = config.device()
    cm = ops.weight_search_confusion(torch.from_numpy(np.ascontiguousarray(preds)).to(dev),
                                     torch.from_numpy(gt.astype(np.int32)).to(dev),
                                     torch.from_numpy(np.ascontiguousarray(weights, dtype=np.float64)).to(dev)).cpu().numpy()
    return _fusion_metrics_from_confusion(cm)[:, 2]


def _first_strict_best(metric: np.ndarray):
    """`if metric > best` with best starting at 0: index of the first candidate reaching the running maximum, or None."""
    best, idx = 0, None
    run = np.maximum.accumulate(metric)
    cand = np.nonzero((metric == run) & (metric > 0))[0]
    for i in cand:                       # candidates are few: strictly increasing prefix maxima
        if metric[i] > best:
            best, idx = metric[i], int(i)
    return best, idx


def get_weights_prob_model(ground_truth, predictions, num_weights, num_classes):
    num_predictions = len(predictions)
    weights = np.zeros(shape=(num_weights, num_predictions, num_classes))
    for i in range(num_weights):                                   # same RNG stream as the reference loop
        weights[i] = np.random.dirichlet(alpha=np.ones((num_predictions,)), size=num_classes).T
    metric = _search_metrics(ground_truth, predictions, weights)
    best, idx = _first_strict_best(metric)
    best_weights = None if idx is None else weights[idx]
    print("final best metric:%f" % (best))
    print("weights:", best_weights)
    return best_weights


def _grid_search(weights, ground_truth, predictions, n_models):
    grid = np.asarray(list(weights), dtype=np.float64)
    for p in predictions:
        if isinstance(p, np.ndarray) and p.dtype == np.float32:
            raise TypeError("float32 prediction arrays would be fused in float32 by the reference; pass lists / float64 (as get_pred_av.py does)")
    mesh = np.stack(np.meshgrid(*([grid] * n_models), indexing="ij"), axis=-1).reshape(-1, n_models)   # w_s slowest, like the loops
    cand = np.repeat(mesh[:, :, None], 7, axis=2)
    metric = _search_metrics(ground_truth, predictions, cand)
    best, idx = _first_strict_best(metric)
    best_weights = [0] * n_models if idx is None else [weights[j] for j in np.unravel_index(idx, (len(grid),) * n_models)]
    print("final best metric:%f" % (best))
    print("weights:", best_weights)
    return best_weights


def get_weights_v_model(weights, ground_truth, predictions):
    return _grid_search(weights, ground_truth, predictions[:2], 2)


def get_weights_av_model(weights, ground_truth, predictions):
    return _grid_search(weights, ground_truth, predictions[:3], 3)
