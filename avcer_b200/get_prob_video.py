"""Drop-in for the reference's src/get_prob_video.py: same entry point, same DataFrames / CSVs;
the per-frame loop (:91-180) is replaced by one index plan + batched K1 -> VS -> VD launches.
"""
from __future__ import annotations

import os

import cv2
import numpy as np
import pandas as pd
import torch

from . import config
from .pipeline import vd_step

DICT_EMO_VIDEO = {0: "Neutral", 1: "Happiness", 2: "Sadness", 3: "Surprise", 4: "Fear", 5: "Disgust", 6: "Anger"}


def _present_frames(path_images: str, total_frames: int):
    """Frame i is present iff `00/{i:06d}.jpg` exists (get_prob_video.py:79,93-94: only face track "00" is read)."""
    folder = os.path.join(path_images, "00")
    names = set(os.listdir(folder))
    exists = np.zeros(total_frames, dtype=bool)
    paths = []
    for i in range(total_frames):
        name = str(i).zfill(6) + ".jpg"
        if name in names:
            exists[i] = True
            paths.append(os.path.join(folder, name))
    return paths, exists


CHUNK = 512      # crops per decode + VS chunk of the GPU-decoded path: the host reads / parses chunk k+1 while chunk k runs
READ_THREADS = 8  # workers of avcer_read_files: open() + read() from Python cost ~20 us per 30 KB crop of interpreter overhead
_read_buf = None


def _read_files(paths):
    """The files of `paths` read into ONE pinned host buffer by the library's multi-threaded reader (avcer_read_files).
    Returns (buffer, offsets [n], sizes [n]); the buffer is reused by the next call (after its last upload has finished)."""
    import ctypes

    from . import _lib, jpeg

    global _read_buf
    n = len(paths)
    lib = _lib.load()
    arr = (ctypes.c_char_p * n)(*[os.fsencode(p) for p in paths])
    offsets = np.empty(n, dtype=np.int64)
    sizes = np.empty(n, dtype=np.int64)
    needed = ctypes.c_int64(0)
    if _read_buf is not None:
        jpeg.wait_uploaded(_read_buf)
    for attempt in range(2):
        cap = 0 if _read_buf is None else _read_buf.numel()
        rc = lib.avcer_read_files(arr, n, None if _read_buf is None else _read_buf.data_ptr(), cap, offsets.ctypes.data,
                                  sizes.ctypes.data, ctypes.byref(needed), READ_THREADS)
        if rc == 0:
            break
        if needed.value > cap and attempt == 0:            # first call / larger clip: grow the buffer and read again
            _read_buf = torch.empty(max(needed.value * 5 // 4, 1 << 20), dtype=torch.uint8, pin_memory=torch.cuda.is_available())
            continue
        msg = lib.avcer_last_error().decode()
        raise FileNotFoundError(msg) if "cannot" in msg else _lib.AvcerError(msg)
    return _read_buf, offsets, sizes


def _unsupported(e):
    from . import jpeg

    return jpeg.UnsupportedJpeg(f"{e} -- the GPU decoder covers what cv2.imwrite writes by default (baseline, 4:2:0 / 4:4:4); "
                                "avcer_b200.config.set_jpeg_decoder('cv2') decodes on the host instead")


def _vs_from_files_gpu(eng, paths):
    """The crops decoded ON THE GPU (avcer_jpeg_decode, bit-identical to the cv2.imread of get_prob_video.py:95) and fed to
    K1 + VS chunk by chunk: the host only reads file bytes and parses JPEG headers, and does so for the next chunk while
    the GPU decodes and classifies the current one (no synchronisation until every chunk is enqueued)."""
    from . import jpeg

    n = len(paths)
    probs, feats = eng._vs_outputs(n)
    pending = []
    for s in range(0, n, CHUNK):
        buf, foff, fsize = _read_files(paths[s:s + CHUNK])
        try:
            flat, offsets, hs, ws, st = jpeg.decode_packed(buf, foff, fsize, eng.device, defer_status=True)
        except jpeg.UnsupportedJpeg as e:
            raise _unsupported(e) from e
        st.base = s
        pending.append(st)
        eng.vs_forward_ragged(flat, offsets, hs, ws, out=(probs[s:s + len(foff)], feats[s:s + len(foff)]))
    for st in pending:
        try:
            st.check()
        except jpeg.UnsupportedJpeg as e:
            raise _unsupported(e) from e
    return probs, feats


def _load_crops_cv2(paths, device):
    """Host decode with cv2.imread exactly as the reference does (:95); the crops are staged 16-byte aligned."""
    chunks, offsets, hs, ws = [], [], [], []
    off = 0
    for p in paths:
        img = cv2.imread(p)                                          # BGR uint8, what pth_processing ends up consuming
        h, w, _ = img.shape
        chunks.append(np.ascontiguousarray(img).reshape(-1))
        offsets.append(off)
        hs.append(h)
        ws.append(w)
        off += (h * w * 3 + 15) // 16 * 16                           # keep every crop 16-byte aligned
    flat = np.zeros(max(off, 16), dtype=np.uint8)
    for c, o in zip(chunks, offsets):
        flat[o:o + c.size] = c
    return (torch.from_numpy(flat).to(device), np.asarray(offsets, dtype=np.int64), np.asarray(hs, dtype=np.int32),
            np.asarray(ws, dtype=np.int32))


def preprocess_video_and_predict(path_images="", save_path="", fps=30, total_frames=[], flag_save_prob=False,
                                 flag_heatmaps=False, model_heatmaps=None):
    """Returns (df_dynamic, df_static): one row per frame index, columns DICT_EMO_VIDEO; VS rows are
    softmax probabilities, VD rows raw logits (reference :182-187)."""
    if flag_heatmaps:
        raise NotImplementedError("Grad-CAM heatmaps need a backward pass and are outside the accelerated path")
    eng = config.video_engine()
    paths, exists = _present_frames(path_images, total_frames)
    n_present = len(paths)
    if n_present:
        if config.jpeg_decoder() == "gpu":
            probs, feats = _vs_from_files_gpu(eng, paths)
        else:
            flat, offsets, hs, ws = _load_crops_cv2(paths, eng.device)
            probs, feats = eng.vs_forward_ragged(flat, offsets, hs, ws)
    else:
        probs = torch.zeros((1, 7), device=eng.device)
        feats = torch.zeros((1, 512), device=eng.device, dtype=torch.float32)
    stat, dyn, plans = eng.video_rows(probs, feats, [exists], [fps])
    plan = plans[0]
    # np.array() over a list mixing float32 rows and float64 zero rows promotes to float64 (:89,182-187)
    has_zero_row = bool((plan.stat_src < 0).any() or (plan.dyn_src < 0).any())
    dt = np.float64 if has_zero_row else np.float32
    cols = list(DICT_EMO_VIDEO.values())
    df_dynamic = pd.DataFrame(dyn.cpu().numpy().astype(dt), columns=cols)
    df_static = pd.DataFrame(stat.cpu().numpy().astype(dt), columns=cols)
    if flag_save_prob:
        os.makedirs(save_path, exist_ok=True)
        df_dynamic.to_csv(os.path.join(save_path, "dynamic__{}.csv".format(os.path.basename(path_images))), index=False)
        df_static.to_csv(os.path.join(save_path, "static__{}.csv".format(os.path.basename(path_images))), index=False)
    return df_dynamic, df_static
