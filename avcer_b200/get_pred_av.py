"""Drop-in for the fusion arithmetic of the reference's src/get_pred_av.py: the CSV-driven
`get_c_expr_db_pred` (:198-334) and the weighted-fusion argmax of `get_metrics` (:34-40).
Dataset/annotation plumbing, metric reports and plots of that script are outside the path.
"""
from __future__ import annotations

import os

import numpy as np
import pandas as pd
import torch

from . import _lib, config, ops
from .data.utils import get_image_location, save_txt  # noqa: F401
from .run import COLUMN_NAMES, NAME_EMO, audio_frame_rows


def fused_argmax(predictions, weights_1, weights_2):
    """get_metrics :34-40: argmax over the 7 basic emotions of sum_m P_m * W1[m] * W2[m] (float64), on the GPU
    (avcer_fused_argmax).  predictions: n_models arrays [n, 7]; weights_1: [n_models][7]; weights_2: [n_models]."""
    dev = config.device()
    preds = torch.from_numpy(np.ascontiguousarray(np.stack([np.asarray(p, dtype=np.float64) for p in predictions]))).to(dev)
    m, n, k = preds.shape
    if k != 7:
        raise ValueError("fused_argmax handles 7-class probability rows")
    w1 = torch.from_numpy(np.array(np.broadcast_to(np.asarray(weights_1, dtype=np.float64).reshape(m, -1), (m, 7)))).to(dev)
    w2 = torch.from_numpy(np.asarray(weights_2, dtype=np.float64).reshape(m)).to(dev)
    labels = torch.empty(n, device=dev, dtype=torch.int32)
    _lib.check(_lib.load().avcer_fused_argmax(preds.data_ptr(), m, n, w1.data_ptr(), w2.data_ptr(), labels.data_ptr(),
                                              torch.cuda.current_stream().cuda_stream))
    return labels.cpu().numpy()


def get_c_expr_db_pred(prediction_file_format, root, path_preds, name_videos, weights_1, weights_2, modality,
                       weight_type, ce_weights_type, ce_mask):
    dev = config.device()
    cols = NAME_EMO[:-1]
    fmt = pd.read_csv(prediction_file_format)
    fmt["curr_video"] = [i.split("/")[0] for i in fmt.image_location]
    vs_rows, vd_rows, a_rows, image_locations = [], [], [], []
    for curr_video in name_videos:
        stat = pd.read_csv(os.path.join(root, path_preds[0], "static__" + curr_video) + ".csv")
        dyn = pd.read_csv(os.path.join(root, path_preds[0], "dynamic__" + curr_video) + ".csv")
        audio = pd.read_csv(os.path.join(root, path_preds[1], path_preds[2], curr_video) + ".csv")
        image_location = fmt[fmt.curr_video == curr_video].image_location.tolist()
        wanted = np.asarray(sorted({int(l.split("/")[1].split(".")[0]) - 1 for l in image_location}), dtype=np.int64)
        s_sel = stat[np.isin(stat.index.to_numpy(), wanted)][cols].to_numpy(dtype=np.float64)
        d_sel = dyn[np.isin(dyn.index.to_numpy(), wanted)][cols].to_numpy(dtype=np.float64)
        uniq, means, acols = audio_frame_rows(audio, dev, dropna=True)
        sel = np.nonzero(np.isin(uniq, wanted))[0]
        if len(image_location) > len(sel):
            sel = np.r_[sel, np.full(len(image_location) - len(sel), sel[-1])]
        a = ops.gather_rows(means, torch.from_numpy(sel.astype(np.int32)).to(dev), len(sel))
        vs_rows.append(torch.from_numpy(s_sel).to(dev))
        vd_rows.append(ops.softmax7(torch.from_numpy(np.ascontiguousarray(d_sel)).to(dev)))
        a_rows.append(ops.softmax7(a).double())
        image_locations.extend(image_location)
    p_vs, p_vd, p_a = torch.cat(vs_rows).contiguous(), torch.cat(vd_rows).contiguous(), torch.cat(a_rows).contiguous()
    w1 = np.asarray(weights_1, dtype=np.float64).tolist()
    labels = ops.fuse_compound(p_vs, p_vd, p_a, w1, list(weights_2), ce_weights_type, ce_mask).cpu().numpy()
    av_pred = labels[0]
    save_path = "src/pred_results/DF_C_EXPR_DB/"
    os.makedirs(save_path, exist_ok=True)
    save_txt(COLUMN_NAMES, image_locations, av_pred,
             os.path.join(save_path, f"C_EXPR_DB_{modality}_sd_{weight_type}_{ce_weights_type}_{ce_mask}.txt"))
    return av_pred, image_locations
