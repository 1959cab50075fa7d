"""Drop-in for the reference's src/run.py: `get_c_expr_db_pred` (the fusion stage, :25-189) and
`run_inference` (:192-308).  Alignment bookkeeping stays on the host (integer / string work);
column permutation, softmax, per-frame audio means, weighted fusion, compound rule and argmax run
on the GPU (avcer_gather_rows, avcer_softmax7(_f64), avcer_window_to_frame_mean, avcer_fuse_compound).
"""
from __future__ import annotations

import os
import time
from typing import Optional

import numpy as np
import pandas as pd
import torch

from . import config, ops
from .data.utils import get_image_location, save_txt  # noqa: F401  (re-exported like the reference)
from .pipeline import AUDIO_ORDER
from .tables import AudioTable

NAME_EMO = AUDIO_ORDER                                   # run.py:56-65
COM_EMO = {"Fearfully Surprised": [3, 6], "Happily Surprised": [4, 6], "Sadly Surprised": [5, 6],
           "Disgustedly Surprised": [2, 6], "Angrily Surprised": [1, 6], "Sadly Fearful": [3, 5], "Sadly Angry": [1, 5]}
COLUMN_NAMES = ["image_location", "Fearfully_Surprised", "Happily_Surprised", "Sadly_Surprised", "Disgustedly_Surprised",
                "Angrily_Surprised", "Sadly_Fearful", "Sadly_Angry"]


def audio_frame_rows(audio_df, dev, dropna: bool = False):
    """groupby("frames").mean() of the long-format audio table on the GPU.  Returns (sorted unique
    frame ids [U], per-frame means [U, ncls] device tensor in the table's dtype, value columns).
    audio_df: a DataFrame like the reference's (one row per (window, frame), `frames` strings), or the drivers'
    AudioTable façade, which skips the string table altogether (avcer_b200/tables.py)."""
    if isinstance(audio_df, AudioTable):
        if not dropna and not audio_df.materialized:
            # façade fast path: window logits + frame ranges -> per-frame means, no strings, no host round trip
            logits = audio_df.window_logits
            logits = (logits if isinstance(logits, torch.Tensor) else torch.from_numpy(np.asarray(logits))).to(dev).contiguous()
            uniq = audio_df.frame_ids()
            n_all = int(audio_df.f_hi.max()) if len(audio_df.f_hi) else 0
            lo = torch.from_numpy(np.minimum(audio_df.f_lo, n_all).astype(np.int32)).to(dev)
            hi = torch.from_numpy(np.minimum(audio_df.f_hi, n_all).astype(np.int32)).to(dev)
            means = ops.window_to_frame_mean(logits, lo, hi, n_all)
            if len(uniq) != n_all:
                means = ops.gather_rows(means, torch.from_numpy(uniq.astype(np.int32)).to(dev), len(uniq))
            return uniq, means, list(audio_df.value_columns)
        audio_df = audio_df.materialize()            # someone edited the table, or NaN rows are to be dropped: generic path
    if dropna:
        audio_df = audio_df.dropna()
    cols = [c for c in audio_df.columns if c != "frames"]
    ids = audio_df["frames"].str.slice(0, -4).astype(np.int64).to_numpy()
    # pandas accumulates a group in the column dtype: float32 for the tables the audio drivers return, float64 for
    # tables read back from CSV (get_pred_av.py:246-249)
    f64 = all(audio_df[c].dtype == np.float64 for c in cols)
    vals = np.ascontiguousarray(audio_df[cols].to_numpy(dtype=np.float64 if f64 else np.float32))
    order = np.argsort(ids, kind="stable")               # pandas accumulates each group in row order
    ids_s = ids[order]
    uniq, first = np.unique(ids_s, return_index=True)
    rank = np.searchsorted(uniq, ids_s).astype(np.int32)  # dense group index of every row
    rows = torch.from_numpy(vals[order]).to(dev)
    lo = torch.from_numpy(rank).to(dev)
    hi = torch.from_numpy(rank + 1).to(dev)
    means = ops.window_to_frame_mean(rows, lo, hi, len(uniq))
    return uniq, means, cols


def _device_rows(df: pd.DataFrame, cols, dev) -> torch.Tensor:
    a = df[cols].to_numpy()
    if a.dtype not in (np.float32, np.float64):
        a = a.astype(np.float64)
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def aligned_probabilities(stat_df, dyn_df, audio_df, name_video, dropna_audio=False, softmax_after_fill=False):
    """run.py:76-103 -> (p_vs, p_vd, p_a device tensors [n,7] of one common dtype, image_location)."""
    dev = config.device()
    cols = NAME_EMO[:-1]
    image_location = [f"{name_video}/{str(f + 1).zfill(5)}.jpg" for f in dyn_df.index]
    keep = set(dyn_df.index)
    stat_sel = stat_df[[i in keep for i in stat_df.index]]
    p_vs = _device_rows(stat_sel, cols, dev)
    p_vd = ops.softmax7(_device_rows(dyn_df, cols, dev))
    uniq, means, acols = audio_frame_rows(audio_df, dev, dropna=dropna_audio)
    sel = np.nonzero(np.isin(uniq, np.fromiter(keep, dtype=np.int64)))[0]
    if len(image_location) > len(sel):
        if len(sel) == 0:
            raise IndexError("index -1 is out of bounds for axis 0 with size 0")
        sel = np.r_[sel, np.full(len(image_location) - len(sel), sel[-1])]
    if list(acols[:7]) != cols:
        raise KeyError(f"audio columns must start with {cols}")
    a_rows = ops.gather_rows(means, torch.from_numpy(sel.astype(np.int32)).to(dev), len(sel))
    p_a = ops.softmax7(a_rows)
    # bring the three streams to one dtype (numpy promotes mixed float32/float64 to float64)
    if torch.float64 in (p_vs.dtype, p_vd.dtype, p_a.dtype):
        p_vs, p_vd, p_a = p_vs.double(), p_vd.double(), p_a.double()
    return p_vs.contiguous(), p_vd.contiguous(), p_a.contiguous(), image_location


def get_c_expr_db_pred(stat_df: pd.DataFrame, dyn_df: pd.DataFrame, audio_df: pd.DataFrame, name_video: str,
                       weights_1, weights_2, ce_weights_type: bool, ce_mask: bool, flag_save_prob: bool):
    """Same contract as run.py:25-189: returns (av_ce, vs_ce, vd_ce, a_ce, image_location)."""
    p_vs, p_vd, p_a, image_location = aligned_probabilities(stat_df, dyn_df, audio_df, name_video)
    if len(p_vs) != len(p_vd) or len(p_a) != len(p_vd):
        raise ValueError(f"operands could not be broadcast together with shapes ({len(p_vs)},7) ({len(p_vd)},7) ({len(p_a)},7)")
    labels = ops.fuse_compound(p_vs, p_vd, p_a, weights_1, weights_2, ce_weights_type, ce_mask).cpu().numpy()
    av_ce, vs_ce, vd_ce, a_ce = labels[0], labels[1], labels[2], labels[3]
    if flag_save_prob:
        save_path = "src/pred_results/DF_C_EXPR_DB/"
        os.makedirs(save_path, exist_ok=True)
        save_txt(COLUMN_NAMES, image_location, av_ce,
                 os.path.join(save_path, f"C_EXPR_DB_av_{ce_weights_type}_{ce_mask}_{name_video}.txt"))
    return av_ce, vs_ce, vd_ce, a_ce, image_location


def run_inference(path_video: str = "", path_save_results: str = "", flag_save_prob: bool = False,
                  weights_prob_model: Optional[list] = None, weights_model: Optional[list] = [1, 1, 1],
                  flag_heatmaps: bool = False, model_heatmaps: str = "static", ce_weights_type: bool = True,
                  ce_mask: bool = False, flag_save_plot_pred: bool = True) -> None:
    """run.py:192-308: face crops (detector + tracker) -> VS / VD -> audio -> compound expressions.  The reference detects
    on every call; here crops of track "00" that already exist under <path_save_results>/<clip>/00/ are reused (the
    BASELINE configs start from pre-cropped faces), otherwise `VideoPredictor.process` writes them first.  fps and the frame
    count are the detector's (`int()` of the container's values, get_face_images.py:23-24)."""
    import cv2

    from .get_prob_audio_8_cl import preprocess_audio_and_predict
    from .get_prob_video import preprocess_video_and_predict

    start_time = time.time()
    clip = os.path.basename(path_video)[:-4]
    crops = os.path.join(path_save_results, clip)
    if not os.path.isdir(os.path.join(crops, "00")):
        from .data.get_face_images import VideoPredictor

        print(f"Face images detection in video: {os.path.basename(path_video)}")
        detect = VideoPredictor()
        detect.process(path_video, path_save_results)
        fps, total_frames = detect.fps, detect.total_frames
        if not os.path.isdir(os.path.join(crops, "00")):
            raise FileNotFoundError(f"{crops}/00: no face was detected in {path_video}")
    else:
        cap = cv2.VideoCapture(path_video)
        fps = int(cap.get(cv2.CAP_PROP_FPS))
        total_frames = int(cap.get(cv2.CAP_PROP_FRAME_COUNT))
        cap.release()
    if not fps or total_frames <= 0:
        raise RuntimeError(f"cannot read fps / frame count of {path_video}")
    print("Emotion prediction using visual models")
    df_dyn, df_stat = preprocess_video_and_predict(path_images=crops, save_path=path_save_results, fps=fps,
                                                   total_frames=total_frames, flag_save_prob=flag_save_prob,
                                                   flag_heatmaps=flag_heatmaps, model_heatmaps=model_heatmaps)
    print("Emotion prediction using audio model")
    # the reference passes the Windows literal "src\weights" (run.py:245); on the Linux B200 host that is one odd directory
    # name, so the separator is normalised
    df_audio = preprocess_audio_and_predict(path_video=path_video, path_weights=os.path.join("src", "weights"), fps=fps, step=0.5,
                                            padding="mean", save_path=path_save_results, flag_save_prob=flag_save_prob,
                                            window=4, sr=16000, device=config.device())
    print("Compound expression prediction")
    av, vs, vd, a, _ = get_c_expr_db_pred(stat_df=df_stat, dyn_df=df_dyn, audio_df=df_audio, name_video=clip,
                                          weights_1=weights_prob_model, weights_2=weights_model,
                                          ce_weights_type=ce_weights_type, ce_mask=ce_mask, flag_save_prob=flag_save_prob)
    end_time = time.time()
    if flag_save_plot_pred:
        np.savez(os.path.join(path_save_results, "predicted_CEs.npz"), VS=vs, VD=vd, A=a, AV=av)   # plotting is presentation-only
    print(f"Real-time factor for compound expression prediction: {((end_time - start_time) / (total_frames / fps)):.2f}")
