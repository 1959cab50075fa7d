"""Generates tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) in the build
container, and pins the oracle restatements against it on the way (asserts below).

    python -m oracle.make_golden            # from the repo root; needs /root/reference

Everything is seeded; inputs are regenerated from seeds by the tests (avcer_b200.synthetic), so the
fixtures only hold reference OUTPUTS (plus small inputs where a seed is not enough).
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np
import pandas as pd
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from avcer_b200 import get_weights_matrices as gwm  # noqa: E402
from avcer_b200 import synthetic as syn  # noqa: E402
from oracle import audio as oa  # noqa: E402
from oracle import face as ofa  # noqa: E402
from oracle import fusion as of  # noqa: E402
from oracle import harness  # noqa: E402
from oracle import video as ov  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
FUSION_CONFIGS = [("w3", True, False), ("w3", False, True), ("w3", True, True), ("w3", False, False),
                  ("w2", False, True), ("w2", True, False), ("none", False, True), ("none", True, False),
                  ("none", False, False)]


def fusion_weights(tag):
    if tag == "w3":
        return gwm.class_weights(gwm.weights_3), [1, 1, 1]
    if tag == "w2":
        return gwm.class_weights(gwm.weights_2), gwm.model_weights(gwm.weights_2)
    return None, [1, 1, 1]


def synthetic_fusion_inputs(seed=7, n=300, fps=25):
    """DataFrames shaped like the drivers' outputs, with edge cases baked in."""
    rng = np.random.default_rng(seed)
    stat = rng.dirichlet(np.ones(7) * 0.7, size=n).astype(np.float32)
    dyn = (rng.standard_normal((n, 7)) * 1.5).astype(np.float32)
    stat[5] = 1.0 / 7.0                      # exactly on the Rule-1 threshold (float32(1/7) < 1/7 -> masked)
    stat[6] = 0.0                            # zero row
    dyn[7] = 0.0                             # zero logits -> uniform 1/7
    stat[8] = stat[9]                        # duplicates -> argmax ties
    L = int(n / fps * 16000)
    sched = oa.window_schedule(L, fps, 0.5)
    rows, frames = [], []
    for (s, e, lo, hi) in sched:
        logit = (rng.standard_normal(8) * 1.2).astype(np.float32)
        if e - s == 0:
            logit[:] = np.nan               # the trailing empty window of the "mean" padding
        for f in range(lo, hi):
            rows.append(logit)
            frames.append(str(f).zfill(6) + ".jpg")
    stat_df = pd.DataFrame(stat, columns=of.VIDEO_ORDER)
    dyn_df = pd.DataFrame(dyn, columns=of.VIDEO_ORDER)
    audio_df = pd.DataFrame(np.asarray(rows), columns=of.AUDIO_ORDER)
    audio_df["frames"] = frames
    return stat_df, dyn_df, audio_df


def make_fusion():
    import run as ref_run
    from data.utils import get_compound_expression as ref_gce
    from data.utils import softmax as ref_softmax

    stat_df, dyn_df, audio_df = synthetic_fusion_inputs()
    out = {"stat": stat_df.values, "dyn": dyn_df.values, "audio_rows": audio_df[of.AUDIO_ORDER].values,
           "audio_frames": np.asarray([int(f[:-4]) for f in audio_df["frames"]], dtype=np.int64)}
    p_vs, p_vd, p_a, loc = of.align_streams(stat_df.copy(), dyn_df.copy(), audio_df.copy(), "clip")
    out.update(p_vs=p_vs, p_vd=p_vd, p_a=p_a)
    assert np.array_equal(ref_softmax(dyn_df[of.AUDIO_ORDER[:7]].values), p_vd)
    for tag, cwt, cm in FUSION_CONFIGS:
        w1, w2 = fusion_weights(tag)
        ref = ref_run.get_c_expr_db_pred(stat_df.copy(), dyn_df.copy(), audio_df.copy(), "clip", w1, w2, cwt, cm, False)
        mine = of.get_c_expr_db_pred(stat_df, dyn_df, audio_df, "clip", w1, w2, cwt, cm)
        for a, b in zip(ref[:4], mine[:4]):
            assert np.array_equal(a, b), (tag, cwt, cm)
        assert ref[4] == mine[4]
        out[f"labels_{tag}_{int(cwt)}_{int(cm)}"] = np.stack(ref[:4])
        # compound score arrays of the reference function itself, on the weighted AV stream
        if w1:
            fused = p_vs * w1[0] * w2[0] + p_vd * w1[1] * w2[1] + p_a * w1[2] * w2[2]
            sc = ref_gce(fused, ref_run_com_emo(), {1: 5, 2: 6, 3: 5, 4: 6, 5: 4, 6: 2}, cwt, cm)
            assert np.array_equal(sc, of.compound_scores(fused, cwt, cm), equal_nan=True)
            out[f"scores_{tag}_{int(cwt)}_{int(cm)}"] = sc
    # float64 DataFrames (zero rows appended by the video driver)
    ref64 = ref_run.get_c_expr_db_pred(stat_df.astype(np.float64), dyn_df.astype(np.float64), audio_df.copy(), "clip",
                                       None, [1, 1, 1], False, True, False)
    out["labels_none_f64_0_1"] = np.stack(ref64[:4])
    np.savez_compressed(os.path.join(OUT, "fusion.npz"), **out)
    print("fusion.npz", len(loc), "frames")


def ref_run_com_emo():
    return {"Fearfully Surprised": [3, 6], "Happily Surprised": [4, 6], "Sadly Surprised": [5, 6],
            "Disgustedly Surprised": [2, 6], "Angrily Surprised": [1, 6], "Sadly Fearful": [3, 5], "Sadly Angry": [1, 5]}


def make_preprocess():
    from PIL import Image

    from data.utils import pth_processing as ref_pp
    import cv2

    sizes = sorted(set([1, 2, 3, 7, 64, 100, 111, 112, 150, 200, 223, 224, 225, 256, 300, 333, 448, 449, 480, 512, 640,
                        720, 777, 1000, 1080, 1279, 1280, 1920, 2047] + list(np.random.default_rng(0).integers(8, 2000, 30))))
    tables = np.zeros((len(sizes), 224), dtype=np.int32)
    for i, s in enumerate(sizes):
        ramp = np.arange(s, dtype=np.int32)
        lo = Image.fromarray((ramp % 256).astype(np.uint8)[None, :].repeat(2, 0))
        hi = Image.fromarray((ramp // 256).astype(np.uint8)[None, :].repeat(2, 0))
        lo = np.asarray(lo.resize((224, 224), Image.Resampling.NEAREST))[0].astype(np.int32)
        hi = np.asarray(hi.resize((224, 224), Image.Resampling.NEAREST))[0].astype(np.int32)
        tables[i] = hi * 256 + lo
        assert np.array_equal(tables[i], ov.nearest_index_table(int(s))), s
    shapes = [(224, 224), (100, 150), (480, 640), (77, 333), (225, 223), (1080, 1920)]
    digests = []
    rng = np.random.default_rng(1)
    for (h, w) in shapes:
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        ref = ref_pp(Image.fromarray(cv2.cvtColor(img, cv2.COLOR_BGR2RGB))).numpy()
        assert np.array_equal(ref, ov.pth_processing(img)), (h, w)
        digests.append(hashlib.sha256(ref.tobytes()).hexdigest())
    np.savez_compressed(os.path.join(OUT, "preprocess.npz"), sizes=np.asarray(sizes, dtype=np.int32), tables=tables,
                        shapes=np.asarray(shapes, dtype=np.int32), digests=np.asarray(digests))
    print("preprocess.npz", len(sizes), "sizes")


def make_video(workdir):
    import cv2

    out = {}
    crops = syn.make_crops(11, 6)
    x = torch.from_numpy(np.concatenate([ov.pth_processing(c) for c in crops]))
    for init in ("spread", "default", "mid"):
        sd = syn.make_vs_state_dict(0, init)
        m = harness.reference_resnet(sd)
        with torch.no_grad():
            feat = m.extract_features(x)
            logits = m(x)
        o_logits, o_feat = ov.resnet50_forward(sd, x)
        assert torch.equal(o_logits, logits) and torch.equal(o_feat, feat)
        out[f"vs_{init}_probs"] = torch.softmax(logits, 1).numpy()
        out[f"vs_{init}_feat"] = feat.numpy()
    sd_vd = syn.make_vd_state_dict(1)
    lm = harness.reference_lstm(sd_vd)
    g = torch.Generator().manual_seed(5)
    xw = torch.relu(torch.randn(12, 10, 512, generator=g))
    with torch.no_grad():
        vd = lm(xw)
    assert (ov.lstm_forward(sd_vd, xw) - vd).abs().max() < 5e-6
    out["vd_logits"] = vd.numpy()
    # the stock per-frame driver on JPEG crops with gaps (two scenarios)
    sd_vs = syn.make_vs_state_dict(0, "spread")
    harness.save_video_weights(workdir, sd_vs, sd_vd)
    import get_prob_video as ref_gpv   # loads the weights at import

    for tag, n, fps, missing, size in (("a", 24, 25, {7, 8, 15}, 160), ("b", 20, 30, {0, 1, 11}, 224)):
        frames = syn.make_crops(21 + len(tag), n, size)
        clip = os.path.join(workdir, f"clip_{tag}")
        os.makedirs(os.path.join(clip, "00"), exist_ok=True)
        for i in range(n):
            if i not in missing:
                cv2.imwrite(os.path.join(clip, "00", f"{i:06d}.jpg"), frames[i])
        df_dyn, df_stat = ref_gpv.preprocess_video_and_predict(path_images=clip, save_path=workdir, fps=fps, total_frames=n)
        decoded = [cv2.imread(os.path.join(clip, "00", f"{i:06d}.jpg")) if i not in missing else None for i in range(n)]
        o_dyn, o_stat = ov.predict_video(decoded, fps, sd_vs, sd_vd)
        assert o_dyn.dtype == df_dyn.values.dtype and o_stat.dtype == df_stat.values.dtype, (o_dyn.dtype, df_dyn.values.dtype)
        assert np.abs(o_stat - df_stat.values).max() < 1e-5 and np.abs(o_dyn - df_dyn.values).max() < 1e-4
        out[f"drv_{tag}_dyn"] = df_dyn.values
        out[f"drv_{tag}_stat"] = df_stat.values
        out[f"drv_{tag}_meta"] = np.asarray([n, fps, size] + sorted(missing), dtype=np.int64)
    np.savez_compressed(os.path.join(OUT, "video.npz"), **out)
    print("video.npz")


def make_audio():
    from transformers import Wav2Vec2FeatureExtractor

    out = {}
    for ncls, mod in ((8, "get_prob_audio_8_cl"), (7, "get_prob_audio_7_cl")):
        sd = syn.make_audio_state_dict(2, ncls, "spread", 12)
        model = harness.reference_audio_model(sd, ncls, 12)
        ref_mod = __import__(mod)
        for tag, L, fps, padding, step in (("a", 52800 + 123, 25, "mean", 0.5), ("b", 48000, 30, "mean", 1), ("c", 40000 - 160, 25, "repeat", 1)):
            wav = syn.make_wav(31, L)
            er = ref_mod.EmotionRecognition.__new__(ref_mod.EmotionRecognition)
            er.step, er.window, er.sr, er.device, er.padding, er.flag_save_prob = step, 4, 16000, "cpu", padding, False
            er.model_params, er.save_path = {"model_name": "m"}, ""
            er.processor = Wav2Vec2FeatureExtractor(feature_size=1, sampling_rate=16000, padding_value=0.0, do_normalize=True,
                                                    return_attention_mask=True)
            er.audio_model = model
            ref_mod.convert_mp4_to_mp3 = lambda path, sr, _w=wav: torch.from_numpy(_w)
            df = er.load_audio_features("clip.mp4", fps)
            rows, ids, logits = oa.predict_audio(wav, fps, sd, step=step, padding=padding)
            ref_rows = df[of.AUDIO_ORDER[:ncls]].values
            assert [int(f[:-4]) for f in df["frames"]] == ids.tolist()
            assert np.nanmax(np.abs(ref_rows - rows)) < 2e-5, np.nanmax(np.abs(ref_rows - rows))
            assert np.array_equal(np.isnan(ref_rows), np.isnan(rows))
            # per-window logits of the reference = first row of each window's run
            sched = oa.window_schedule(L, fps, step)
            firsts = np.cumsum([0] + [max(hi - lo, 0) for (_, _, lo, hi) in sched])[:-1]
            out[f"a{ncls}_{tag}_window_logits"] = ref_rows[firsts]
            out[f"a{ncls}_{tag}_meta"] = np.asarray([L, fps, {"mean": 0, "constant": 1, "repeat": 2}[padding], int(step * 1000)], dtype=np.int64)
            # pandas groupby mean of the reference table (what run.py:90 computes)
            gm = df.groupby(["frames"]).mean().reset_index()
            out[f"a{ncls}_{tag}_frame_ids"] = np.asarray([int(f[:-4]) for f in gm["frames"]], dtype=np.int64)
            out[f"a{ncls}_{tag}_frame_means"] = gm[of.AUDIO_ORDER[:ncls]].values
    # PyTorch-default-like random init (the init the north star names) for the bf16 2e-3 check
    sd = syn.make_audio_state_dict(2, 8, "default", 12)
    model = harness.reference_audio_model(sd, 8, 12)
    wav = syn.make_wav(31, 52800 + 123)
    sched = oa.window_schedule(len(wav), 25, 0.5)
    xs = np.stack([oa.zero_mean_unit_var(oa.pad_window(wav[s:e], 64000, "mean")) for (s, e, _, _) in sched])
    with torch.no_grad():
        ref = model(torch.from_numpy(xs)).numpy()
    assert np.abs(oa.audio_model_forward(sd, torch.from_numpy(xs)).numpy() - ref).max() < 2e-5
    out["a8_default_window_logits"] = ref
    # "mid" init (spread with the head scaled to a logit range of ~1): the widest init on which bf16 meets 2e-3
    sd = syn.make_audio_state_dict(2, 8, "mid", 12)
    model = harness.reference_audio_model(sd, 8, 12)
    with torch.no_grad():
        ref = model(torch.from_numpy(xs)).numpy()
    assert np.abs(oa.audio_model_forward(sd, torch.from_numpy(xs)).numpy() - ref).max() < 2e-5
    out["a8_mid_window_logits"] = ref
    # ExprModelV1 (the GRU variant of audio_8_cl.py:18-72), first three windows
    sd = syn.make_audio_state_dict(2, 8, "mid", 12, variant="v1")
    model = harness.reference_audio_model(sd, 8, 12)
    with torch.no_grad():
        ref = model(torch.from_numpy(xs[:3])).numpy()
    assert np.abs(oa.audio_model_forward(sd, torch.from_numpy(xs[:3])).numpy() - ref).max() < 2e-5
    out["a8_v1_window_logits"] = ref
    np.savez_compressed(os.path.join(OUT, "audio.npz"), **out)
    print("audio.npz")


def weight_search_inputs(seed=0, n=600):
    rng = np.random.default_rng(seed)
    gt = rng.integers(0, 7, n)
    preds = [rng.dirichlet(np.ones(7) * 0.5, size=n) for _ in range(3)]
    for k in range(3):
        preds[k][np.arange(n), gt] += 0.4 * (k + 1) / 3        # informative but imperfect streams
    return gt, preds


def make_weight_search():
    import contextlib
    import io

    from data.utils import get_weights_av_model as ref_av
    from data.utils import get_weights_prob_model as ref_prob
    from data.utils import get_weights_v_model as ref_v

    gt, preds = weight_search_inputs()
    preds_l = [p.tolist() for p in preds]
    np.random.seed(42)
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        bw = ref_prob(gt.tolist(), preds_l, 60, 7)
    np.random.seed(42)
    W = np.zeros((60, 3, 7))
    for i in range(60):
        W[i] = np.random.dirichlet(alpha=np.ones((3,)), size=7).T
    best, idx = of.search_prob_weights(gt.tolist(), preds_l, W)
    assert np.array_equal(W[idx], bw)
    grid = [0.1, 0.4, 0.7, 1.0]
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        aw = ref_av(grid, gt.tolist(), preds_l)
        vw = ref_v(grid, gt.tolist(), preds_l[:2])
    assert of.search_av_weights(grid, gt.tolist(), preds_l)[1] == aw
    np.savez_compressed(os.path.join(OUT, "weight_search.npz"), prob_best=bw, prob_idx=np.int64(idx), av_best=np.asarray(aw),
                        v_best=np.asarray(vw), grid=np.asarray(grid))
    print("weight_search.npz")


def resample_inputs():
    """Seeded int16 PCM in the two layouts the reference's ffmpeg step can leave behind (it asks for 44.1 kHz stereo)."""
    rng = np.random.default_rng(77)
    t1 = np.arange(16317) / 44100.0
    stereo = np.stack([8000 * np.sin(2 * np.pi * 440 * t1) + 3000 * rng.standard_normal(t1.size),
                       6000 * np.sin(2 * np.pi * 1250 * t1 + 0.3) + 3000 * rng.standard_normal(t1.size)], axis=1)
    mono48 = 9000 * np.sin(2 * np.pi * 900 * np.arange(9001) / 48000.0) + 2000 * rng.standard_normal(9001)
    return {"stereo_44100": (np.clip(stereo, -32768, 32767).astype(np.int16), 44100),
            "mono_48000": (np.clip(mono48, -32768, 32767).astype(np.int16)[:, None], 48000)}


def make_resample(wd):
    """The reference's own convert_mp4_to_mp3 (src/data/utils.py:42-60) on .wav files that already exist (so its
    ffmpeg call is skipped); only torchaudio.load -- unusable here without torchcodec -- is replaced by a reader with
    the same normalisation (int16 / 32768, channels first)."""
    import wave

    import torchaudio
    from data.utils import convert_mp4_to_mp3 as ref_convert

    def load(path):
        with wave.open(path, "rb") as f:
            sr, nch, n = f.getframerate(), f.getnchannels(), f.getnframes()
            pcm = np.frombuffer(f.readframes(n), dtype="<i2").reshape(-1, nch)
        return torch.from_numpy(pcm.astype(np.float32) / 32768.0).t().contiguous(), sr

    out = {}
    stock_load, torchaudio.load = torchaudio.load, load
    try:
        for name, (pcm, sr) in resample_inputs().items():
            path = os.path.join(wd, name + ".wav")
            with wave.open(path, "wb") as f:
                f.setnchannels(pcm.shape[1]); f.setsampwidth(2); f.setframerate(sr)
                f.writeframes(pcm.astype("<i2").tobytes())
            ref = ref_convert(os.path.join(wd, name + ".mp4"), 16000).numpy()
            mine = oa.pcm16_to_mono_16k(pcm, sr, 16000)
            assert ref.shape == mine.shape and np.abs(ref - mine).max() < 2e-6, (name, ref.shape, mine.shape, np.abs(ref - mine).max())
            out[name + "_pcm"] = pcm
            out[name + "_sr"] = np.int64(sr)
            out[name + "_out"] = ref.astype(np.float32)
    finally:
        torchaudio.load = stock_load
    np.savez_compressed(os.path.join(OUT, "resample.npz"), **out)
    print("resample.npz")


def pred_av_tables():
    """Two clips shaped like the CSVs the drivers write (video columns in VIDEO_ORDER, long-format audio table with NaN
    rows and a `frames` column); the second clip's audio stops early so the repeat-last-row rule applies."""
    tables = {}
    for k, (name, n, seed) in enumerate((("clipA", 120, 21), ("clipB", 90, 22))):
        stat_df, dyn_df, audio_df = synthetic_fusion_inputs(seed, n)
        if k == 1:
            ids = audio_df["frames"].str.slice(0, -4).astype(int)
            audio_df = audio_df[ids < 70].reset_index(drop=True)
        tables[name] = (stat_df, dyn_df, audio_df)
    locs = [f"clipA/{str(f + 1).zfill(5)}.jpg" for f in range(120) if f not in (3, 4, 57, 119)]
    locs += [f"clipB/{str(f + 1).zfill(5)}.jpg" for f in range(90)]
    fmt = pd.DataFrame({"image_location": locs})
    for c in ("Fearfully_Surprised", "Happily_Surprised", "Sadly_Surprised", "Disgustedly_Surprised", "Angrily_Surprised",
              "Sadly_Fearful", "Sadly_Angry"):
        fmt[c] = 0
    return tables, fmt


def write_pred_av_files(root, tables, fmt):
    """The directory layout get_pred_av.get_c_expr_db_pred reads (get_pred_av.py:232-249)."""
    os.makedirs(os.path.join(root, "video"), exist_ok=True)
    os.makedirs(os.path.join(root, "audio_mean_0.5", "model"), exist_ok=True)
    for name, (stat_df, dyn_df, audio_df) in tables.items():
        stat_df.to_csv(os.path.join(root, "video", f"static__{name}.csv"), index=False)
        dyn_df.to_csv(os.path.join(root, "video", f"dynamic__{name}.csv"), index=False)
        audio_df.to_csv(os.path.join(root, "audio_mean_0.5", "model", f"{name}.csv"), index=False)
    fmt_path = os.path.join(root, "format.txt")
    fmt.to_csv(fmt_path, index=False)
    return fmt_path, ["video", "audio_mean_0.5", "model"]


PRED_AV_CONFIGS = [("av8", [1, 1, 1], False, True), ("av8", [1, 1, 1], True, False), ("av8", "double", True, True), ("av7", [1, 1, 1], False, False)]


def pred_av_weights(tag, w2):
    from avcer_b200 import get_weights_matrices as gwm

    table = gwm.weights_3 if tag == "av8" else gwm.weights_2
    w1 = gwm.class_weights(table)
    if tag == "av7":                        # two-stream table (video, audio): the dynamic stream gets zero weight
        w1 = [w1[0], [0.0] * 7, w1[1]]
    return np.asarray(w1, dtype=np.float64), (gwm.model_weights(gwm.weights_3) if w2 == "double" else w2)


def make_pred_av(wd):
    """The unmodified get_pred_av.get_c_expr_db_pred (get_pred_av.py:198-334) on CSV files written here."""
    import get_pred_av as ref_pa

    tables, fmt = pred_av_tables()
    root = os.path.join(wd, "preds")
    fmt_path, path_preds = write_pred_av_files(root, tables, fmt)
    read = {n: (pd.read_csv(os.path.join(root, "video", f"static__{n}.csv")), pd.read_csv(os.path.join(root, "video", f"dynamic__{n}.csv")),
                pd.read_csv(os.path.join(root, "audio_mean_0.5", "model", f"{n}.csv"))) for n in tables}
    out = {}
    for i, (tag, w2, cwt, cm) in enumerate(PRED_AV_CONFIGS):
        w1, w2v = pred_av_weights(tag, w2)
        ref_pa.get_c_expr_db_pred(fmt_path, root, path_preds, list(tables), w1, w2v, tag, f"cfg{i}", cwt, cm)
        txt = pd.read_csv(os.path.join("src", "pred_results", "DF_C_EXPR_DB", f"C_EXPR_DB_{tag}_sd_cfg{i}_{cwt}_{cm}.txt"))
        labels = txt.iloc[:, 1].to_numpy(dtype=np.int64)
        mine, locs = of.pred_av_labels(pd.read_csv(fmt_path), read, list(tables), w1, w2v, cwt, cm)
        assert list(txt.iloc[:, 0]) == locs and np.array_equal(labels, mine), (tag, w2, cwt, cm)
        out[f"labels_{i}"] = labels
    out["n_locations"] = np.int64(len(locs))
    np.savez_compressed(os.path.join(OUT, "pred_av.npz"), **out)
    print("pred_av.npz")


def face_tracker_sequences(seed=3):
    """Box sequences that exercise the tracker on their own: drifting boxes, births, deaths, an empty frame (which clears
    every tracklet), a degenerate zero-area box and two faces swapping order."""
    rng = np.random.default_rng(seed)
    seqs = []
    for s in range(3):
        boxes = rng.uniform(20, 200, (4, 2))
        boxes = np.concatenate([boxes, boxes + rng.uniform(30, 80, (4, 2))], axis=1)
        frames = []
        for t in range(12):
            boxes = boxes + rng.normal(0, 4.0 + 6.0 * s, boxes.shape)
            cur = boxes.copy()
            if t == 4:
                cur = cur[:2]                                   # two faces disappear ...
            if t == 5:
                cur = cur[::-1]                                 # ... come back (new ids) in reverse order
            if t == 7 and s == 1:
                cur = np.empty((0, 4))                          # empty frame: every tracklet is dropped
            if t == 9:
                cur = np.concatenate([cur, [[50, 50, 50, 90]]])     # zero-area box: never tracked (id None)
            frames.append(np.concatenate([cur, rng.uniform(0.8, 1.0, (len(cur), 1)), rng.uniform(0, 200, (len(cur), 10))], axis=1).astype(np.float32))
        seqs.append(frames)
    return seqs


def write_face_video(path, frames, fps=25):
    import cv2

    h, w = frames.shape[1:3]
    vw = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"MJPG"), fps, (w, h))
    assert vw.isOpened()
    for f in frames:
        vw.write(f)
    vw.release()


def make_face(wd):
    """The unmodified RetinaFace / RetinaFacePredictor / SimpleFaceTracker / VideoPredictor.process of the reference
    (data/face_detection/ibug/face_detection, data/get_face_images.py) on synthetic weights, frames and a MJPG video."""
    from types import SimpleNamespace

    import cv2
    from data.face_detection.ibug.face_detection import RetinaFacePredictor
    from data.face_detection.ibug.face_detection.retina_face.config import cfg_re50
    from data.face_detection.ibug.face_detection.utils import SimpleFaceTracker
    from data.get_face_images import VideoPredictor

    sd = syn.make_retinaface_state_dict(5, "spread")
    wpath = os.path.join(wd, "retinaface_synthetic.pth")
    torch.save(dict(sd), wpath)
    pred = RetinaFacePredictor(threshold=0.8, device="cpu", model=SimpleNamespace(weights=wpath, config=SimpleNamespace(**cfg_re50)))
    out = {}
    # (1) raw network outputs on one small frame (odd feature-map sizes: 100 x 136 -> 13 x 17, 7 x 9, 4 x 5)
    fr = syn.make_frames(40, 1, 100, 136)[0]
    x = ofa.prepare(fr)
    with torch.no_grad():
        loc, conf, lm = pred.net(x)
    mine = ofa.forward(x, sd)
    for a, b, name in zip((loc, conf, lm), mine, ("loc", "conf", "landms")):
        assert torch.equal(a, b), name                      # same torch ops in the same order: bit-identical
        out["raw_" + name] = a[0].numpy()
    # (2) the predictor call on a short drifting sequence + the tracker on its detections
    frames = syn.make_frames(41, 6, 150, 200)
    tracker, my_tracker = SimpleFaceTracker(iou_threshold=0.4, minimum_face_size=0.0), ofa.SimpleFaceTracker(0.4, 0.0)
    ids_all = []
    for i, f in enumerate(frames):
        dets = pred(f, rgb=False)
        assert np.array_equal(dets, ofa.predict(sd, f)), i
        ids = tracker(dets)
        assert ids == my_tracker(dets)
        out[f"dets_{i}"] = dets
        ids_all.append(np.asarray(ids, dtype=np.int64))
        assert len(dets) > 0, "synthetic detector produced no face: retune synthetic.RF_CLASS_*"
    out["ids"] = np.concatenate(ids_all)
    out["ids_count"] = np.asarray([len(a) for a in ids_all], dtype=np.int64)
    assert np.array_equal(pred(frames[0][..., ::-1].copy(), rgb=True), out["dets_0"])
    # (3) the tracker alone
    for s, seq in enumerate(face_tracker_sequences()):
        tracker.reset()
        my_tracker.reset()
        got = []
        for boxes in seq:
            ids = tracker(boxes)
            assert ids == my_tracker(boxes)
            got.append(np.asarray([-1 if v is None else v for v in ids], dtype=np.int64))
        out[f"track_{s}"] = np.concatenate(got) if got else np.empty(0, dtype=np.int64)
    # (4) VideoPredictor.process, unmodified, on an MJPG file (constructed without __init__: that one hard-codes cuda:0
    # and the weight file inside the package)
    vpath = os.path.join(wd, "faces_clip.avi")
    write_face_video(vpath, syn.make_frames(42, 8, 150, 200))
    vp = VideoPredictor.__new__(VideoPredictor)
    vp.video_stream, vp.device, vp.count_frame = None, "cpu", None
    vp.model, vp.face_tracker = pred, SimpleFaceTracker(iou_threshold=0.4, minimum_face_size=0.0)
    save = os.path.join(wd, "faces_out")
    vp.process(vpath, save)
    vp.video_stream.release()
    vp.video_stream = None
    cap = cv2.VideoCapture(vpath)
    decoded = []
    while True:
        ok, f = cap.read()
        if not ok:
            break
        decoded.append(f)
    cap.release()
    out["video_frames_sha256"] = np.frombuffer(hashlib.sha256(np.stack(decoded).tobytes()).digest(), dtype=np.uint8)
    names, digests = [], []
    for root, _, files in sorted(os.walk(os.path.join(save, "faces_clip"))):
        for fn in sorted(files):
            rel = os.path.relpath(os.path.join(root, fn), save)
            names.append(rel.replace(os.sep, "/"))
            digests.append(np.frombuffer(hashlib.sha256(open(os.path.join(root, fn), "rb").read()).digest(), dtype=np.uint8))
    assert names, "VideoPredictor.process wrote no crop"
    out["process_files"] = np.asarray(names)
    out["process_sha256"] = np.stack(digests)
    # my restatement of the loop on the same decoded frames
    my_tracker.reset()
    mine_files = []
    for i, f in enumerate(decoded):
        dets = ofa.predict(sd, f)
        for (t, sx, sy, ex, ey) in ofa.crop_boxes(dets, my_tracker(dets), f.shape[1], f.shape[0]):
            mine_files.append(f"faces_clip/{t:02d}/{i:06d}.jpg")
    assert sorted(mine_files) == sorted(names)
    np.savez_compressed(os.path.join(OUT, "face.npz"), **out)
    print("face.npz", len(names), "crops;", [len(out[f"dets_{i}"]) for i in range(6)], "detections per frame")


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 1)
    if len(sys.argv) > 1 and sys.argv[1] == "face":          # only the face-detector fixture (the others are unchanged)
        with harness.reference_env() as wd:
            make_face(wd)
        return
    with harness.reference_env() as wd:
        # run.py imports get_prob_video, which loads the (CWD-relative) weight files at import time
        harness.save_video_weights(wd, syn.make_vs_state_dict(0, "spread"), syn.make_vd_state_dict(1))
        make_fusion()
        make_weight_search()
        make_preprocess()
        make_video(wd)
        make_audio()
        make_resample(wd)
        make_pred_av(wd)
        make_face(wd)


if __name__ == "__main__":
    main()
