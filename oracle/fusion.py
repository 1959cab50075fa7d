"""ORACLE (test infrastructure, never shipped on the product path).

CPU restatement in numpy/pandas of the fusion stage of ElenaRyumina/AVCER:
  * softmax                     -- src/data/utils.py:125-127
  * compound-expression rule    -- src/data/utils.py:222-241
  * frame alignment + fusion    -- src/run.py:76-165 (get_c_expr_db_pred)
  * frame id mapping            -- src/data/utils.py:244-247

Pinned against the reference's own functions by oracle/make_golden.py (run in the build container,
where /root/reference is importable) -> tests/golden/fusion_*.npz.
"""
from __future__ import annotations

import numpy as np
import pandas as pd

# audio-model class order (run.py:56-65); the video models use VIDEO_ORDER (get_prob_video.py:56-64)
AUDIO_ORDER = ["Neutral", "Anger", "Disgust", "Fear", "Happiness", "Sadness", "Surprise", "Other"]
VIDEO_ORDER = ["Neutral", "Happiness", "Sadness", "Surprise", "Fear", "Disgust", "Anger"]
# column permutation video order -> audio order (run.py:85-88 selects columns by name)
VIDEO_TO_AUDIO = [VIDEO_ORDER.index(n) for n in AUDIO_ORDER[:7]]  # [0, 6, 5, 4, 1, 2, 3]

# compound classes as index pairs in audio order (run.py:66-74); class id = position
COMPOUND_PAIRS = [(3, 6), (4, 6), (5, 6), (2, 6), (1, 6), (3, 5), (1, 5)]
COMPOUND_NAMES = ["Fearfully Surprised", "Happily Surprised", "Sadly Surprised", "Disgustedly Surprised",
                  "Angrily Surprised", "Sadly Fearful", "Sadly Angry"]
RULE2_WEIGHTS = {1: 5, 2: 6, 3: 5, 4: 6, 5: 4, 6: 2}  # run.py:116-123


def softmax(m: np.ndarray) -> np.ndarray:
    """Row softmax in the dtype of `m` (utils.py:125-127)."""
    e = np.exp(m - np.max(m, axis=1, keepdims=True))
    return e / np.sum(e, axis=1, keepdims=True)


def compound_scores(pred: np.ndarray, ce_weights_type: bool, ce_mask: bool) -> np.ndarray:
    """[n,7] basic-emotion scores -> [n,7] compound scores (utils.py:222-241).

    Rule 1 (ce_mask): scores <= 1/7 are zeroed first (strict >, applied to the already weighted
    scores, in the dtype of `pred`).  Rule 2 (ce_weights_type): the two members of a pair are
    weighted by d[i]/(d[i1]+d[i2]).  The result array is float64 like the reference's np.zeros.
    """
    pred = np.asarray(pred)
    out = np.zeros((len(pred), len(COMPOUND_PAIRS)))
    for k, (i1, i2) in enumerate(COMPOUND_PAIRS):
        if ce_weights_type:
            tot = RULE2_WEIGHTS[i1] + RULE2_WEIGHTS[i2]
            w1, w2 = RULE2_WEIGHTS[i1] / tot, RULE2_WEIGHTS[i2] / tot
        else:
            w1, w2 = 1, 1
        if ce_mask:
            pred = np.where(pred > 1 / 7, pred, 0)
        out[:, k] = pred[:, i1] * w1 + pred[:, i2] * w2
    return out


def fuse_labels(p_vs: np.ndarray, p_vd: np.ndarray, p_a: np.ndarray, weights_1, weights_2,
                ce_weights_type: bool, ce_mask: bool):
    """Weighted fusion + compound rule + argmax for the AV stream and the three single-modality
    streams (run.py:105-165).  Inputs are probabilities in audio class order.  Returns four int64
    label vectors (AV, VS, VD, A)."""
    preds = [p_vs, p_vd, p_a]
    if weights_1:
        single = [preds[m] * weights_1[m] * weights_2[m] for m in range(3)]
        fused = preds[0] * weights_1[0] * weights_2[0]
        for m in (1, 2):
            fused += preds[m] * weights_1[m] * weights_2[m]
    else:
        single = preds
        fused = np.sum(preds, axis=0) / 3
    streams = [fused] + single
    return tuple(np.argmax(compound_scores(s, ce_weights_type, ce_mask)[:, :7], axis=1) for s in streams)


def image_location(video: str, frame_name: str) -> str:
    """'000012.jpg' -> '<video>/00013.jpg' (utils.py:244-247)."""
    return f"{video}/{str(int(frame_name.split('.')[0]) + 1).zfill(5)}.jpg"


def align_streams(stat_df: pd.DataFrame, dyn_df: pd.DataFrame, audio_df: pd.DataFrame, name_video: str):
    """run.py:76-103: returns (p_vs, p_vd, p_a, image_location) with all three [n,7] in audio order."""
    cols = AUDIO_ORDER[:7]
    loc_s = [f"{name_video}/{str(i + 1).zfill(5)}.jpg" for i in stat_df.index]
    loc_d = [f"{name_video}/{str(i + 1).zfill(5)}.jpg" for i in dyn_df.index]
    keep = set(loc_d)
    p_vs = stat_df[[l in keep for l in loc_s]][cols].values
    p_vd = softmax(dyn_df[cols].values)
    a = audio_df.groupby(["frames"]).mean().reset_index()
    a_loc = [image_location(name_video, f) for f in a["frames"]]
    a = a[[l in keep for l in a_loc]][cols].values
    p_a = softmax(a)
    if len(loc_d) > len(p_a):
        p_a = np.vstack((p_a, [p_a[-1]] * (len(loc_d) - len(p_a))))
    return p_vs, p_vd, p_a, loc_d


def get_c_expr_db_pred(stat_df, dyn_df, audio_df, name_video, weights_1, weights_2, ce_weights_type, ce_mask):
    """Restatement of run.get_c_expr_db_pred (run.py:25-189) without the txt side effect."""
    p_vs, p_vd, p_a, loc = align_streams(stat_df.copy(), dyn_df.copy(), audio_df.copy(), name_video)
    av, vs, vd, a = fuse_labels(p_vs, p_vd, p_a, weights_1, weights_2, ce_weights_type, ce_mask)
    return av, vs, vd, a, loc


def pred_av_labels(fmt_df: pd.DataFrame, tables: dict, name_videos, weights_1, weights_2, ce_weights_type, ce_mask):
    """Restatement of get_pred_av.get_c_expr_db_pred (get_pred_av.py:198-334) on DataFrames instead of CSV paths:
    `tables[video] = (static_df, dynamic_df, audio_long_df)` as pd.read_csv returns them (float64 columns), `fmt_df` the
    challenge's frame list (column image_location).  Returns (compound labels [n] int64, image_locations)."""
    cols = AUDIO_ORDER[:7]
    fmt_video = [i.split("/")[0] for i in fmt_df.image_location]
    p_vs, p_vd, p_a, locs = [], [], [], []
    for video in name_videos:
        stat, dyn, audio = tables[video]
        loc_s = [f"{video}/{str(f + 1).zfill(5)}.jpg" for f in stat.index]
        loc_d = [f"{video}/{str(f + 1).zfill(5)}.jpg" for f in dyn.index]
        a = audio.dropna().groupby(["frames"]).mean().reset_index()
        a_loc = [image_location(video, f) for f in a["frames"]]
        wanted = [l for l, v in zip(fmt_df.image_location, fmt_video) if v == video]
        keep = set(wanted)
        vs = stat[[l in keep for l in loc_s]][cols].values
        vd = softmax(dyn[[l in keep for l in loc_d]][cols].values)
        av = a[[l in keep for l in a_loc]][cols].values
        if len(wanted) > len(av):
            av = np.vstack((av, [av[-1]] * (len(wanted) - len(av))))
        p_vs.append(vs); p_vd.append(vd); p_a.append(softmax(av)); locs.extend(wanted)
    preds = [np.concatenate(p_vs), np.concatenate(p_vd), np.concatenate(p_a)]
    w1 = np.asarray(weights_1, dtype=np.float64)
    final = preds[0] * w1[0] * weights_2[0]
    for i in range(1, 3):
        final = final + preds[i] * w1[i] * weights_2[i]
    prob = compound_scores(final, ce_weights_type, ce_mask)
    return np.argmax(prob[:, :7], axis=1), locs


# ------------------------------------------------------------------------------------------------ weight search
def metrics_for_fusion(true, pred):
    """utils.py:115-122: precision, f1 and recall (UAR) averaged over classes 1..6 of sklearn's report."""
    from sklearn.metrics import classification_report

    rep = classification_report(true, pred, output_dict=True, zero_division=0)
    acc = np.zeros(3)
    for cl in range(1, 7):
        for j, key in enumerate(["precision", "f1-score", "recall"]):
            acc[j] += rep[str(cl)][key]
    return tuple(acc / 6)


def search_prob_weights(ground_truth, predictions, weights):
    """utils.py:145-158 for given candidate weights [W, M, C]: (best metric, index of the first strict best)."""
    best, best_idx = 0, None
    for w in range(len(weights)):
        final = np.asarray(predictions[0]) * weights[w, 0]
        for m in range(1, len(predictions)):
            final += np.asarray(predictions[m]) * weights[w, m]
        metric = metrics_for_fusion(ground_truth, np.argmax(final, axis=-1))[2]
        if metric > best:
            best, best_idx = metric, w
    return best, best_idx


def search_av_weights(grid, ground_truth, predictions):
    """utils.py:188-209: triple loop over scalar model weights."""
    p1, p2, p3 = (np.array(p) for p in predictions[:3])
    best_w, best = [0, 0, 0], 0
    for ws in grid:
        for wd in grid:
            for wa in grid:
                acc = metrics_for_fusion(ground_truth, np.argmax(ws * p1 + wd * p2 + wa * p3, axis=1))[2]
                if acc > best:
                    best, best_w = acc, [ws, wd, wa]
    return best, best_w
