"""CPU: the oracle restatements against the golden vectors produced by the unmodified reference
(oracle/make_golden.py).  These pin the checker that the GPU parity tests rely on."""
import hashlib

import numpy as np
import pandas as pd
import pytest
import torch

from avcer_b200 import get_weights_matrices as gwm
from avcer_b200 import synthetic as syn
from oracle import audio as oa
from oracle import fusion as of
from oracle import video as ov
from oracle.make_golden import FUSION_CONFIGS, fusion_weights, synthetic_fusion_inputs


def test_fusion_labels_match_reference(golden):
    g = golden["fusion"]
    stat_df, dyn_df, audio_df = synthetic_fusion_inputs()
    assert np.array_equal(stat_df.values, g["stat"]) and np.array_equal(dyn_df.values, g["dyn"])
    for tag, cwt, cm in FUSION_CONFIGS:
        w1, w2 = fusion_weights(tag)
        got = of.get_c_expr_db_pred(stat_df, dyn_df, audio_df, "clip", w1, w2, cwt, cm)
        assert np.array_equal(np.stack(got[:4]), g[f"labels_{tag}_{int(cwt)}_{int(cm)}"]), (tag, cwt, cm)


def test_fusion_float64_frames(golden):
    g = golden["fusion"]
    stat_df, dyn_df, audio_df = synthetic_fusion_inputs()
    got = of.get_c_expr_db_pred(stat_df.astype(np.float64), dyn_df.astype(np.float64), audio_df, "clip", None, [1, 1, 1], False, True)
    assert np.array_equal(np.stack(got[:4]), g["labels_none_f64_0_1"])


def test_compound_scores_match_reference(golden):
    g = golden["fusion"]
    for tag, cwt, cm in FUSION_CONFIGS:
        w1, w2 = fusion_weights(tag)
        if not w1:
            continue
        fused = g["p_vs"] * w1[0] * w2[0] + g["p_vd"] * w1[1] * w2[1] + g["p_a"] * w1[2] * w2[2]
        assert np.array_equal(of.compound_scores(fused, cwt, cm), g[f"scores_{tag}_{int(cwt)}_{int(cm)}"], equal_nan=True)


def test_weight_tables_match_run_py():
    # run.py:316-344 literal == weights_3[:7].T
    w = gwm.class_weights(gwm.weights_3)
    assert w[0][0] == 0.89900098 and w[1][3] == 0.93791526 and w[2][5] == 0.48672896
    assert gwm.model_weights(gwm.weights_3) == [0.16000000000000003, 0.36000000000000004, 0.01]


def test_nearest_tables_match_pillow(golden):
    g = golden["preprocess"]
    for s, t in zip(g["sizes"], g["tables"]):
        assert np.array_equal(ov.nearest_index_table(int(s)), t), s


def test_preprocess_digests(golden):
    g = golden["preprocess"]
    rng = np.random.default_rng(1)
    for (h, w), d in zip(g["shapes"], g["digests"]):
        img = rng.integers(0, 256, (int(h), int(w), 3), dtype=np.uint8)
        assert hashlib.sha256(ov.pth_processing(img).tobytes()).hexdigest() == str(d)


@pytest.mark.parametrize("init", ["spread", "default", "mid"])
def test_vs_oracle_matches_reference(golden, init):
    g = golden["video"]
    crops = syn.make_crops(11, 6)[:3]
    x = torch.from_numpy(np.concatenate([ov.pth_processing(c) for c in crops]))
    logits, feat = ov.resnet50_forward(syn.make_vs_state_dict(0, init), x)
    assert np.abs(torch.softmax(logits, 1).numpy() - g[f"vs_{init}_probs"][:3]).max() < 1e-6
    assert np.abs(feat.numpy() - g[f"vs_{init}_feat"][:3]).max() < 1e-4


def test_vd_oracle_matches_reference(golden):
    gen = torch.Generator().manual_seed(5)
    xw = torch.relu(torch.randn(12, 10, 512, generator=gen))
    out = ov.lstm_forward(syn.make_vd_state_dict(1), xw).numpy()
    assert np.abs(out - golden["video"]["vd_logits"]).max() < 1e-5


def test_audio_oracle_matches_reference(golden):
    g = golden["audio"]
    L, fps, pad, step = (int(v) for v in g["a8_a_meta"])
    wav = syn.make_wav(31, L)
    rows, ids, logits = oa.predict_audio(wav, fps, syn.make_audio_state_dict(2, 8, "spread", 12), step=step / 1000, padding="mean")
    assert np.nanmax(np.abs(logits - g["a8_a_window_logits"])) < 5e-5
    df = pd.DataFrame(rows)
    df["frames"] = [str(i).zfill(6) + ".jpg" for i in ids]
    gm = df.groupby("frames").mean().reset_index()
    assert [int(f[:-4]) for f in gm["frames"]] == g["a8_a_frame_ids"].tolist()
    assert np.nanmax(np.abs(gm[list(range(8))].values - g["a8_a_frame_means"])) < 5e-5


def test_audio_v1_oracle_matches_reference(golden):
    """ExprModelV1 (GRU variant): the oracle's gru_forward restatement against the reference class's logits."""
    wav = syn.make_wav(31, 52800 + 123)
    sched = oa.window_schedule(len(wav), 25, 0.5)[:2]
    xs = np.stack([oa.zero_mean_unit_var(oa.pad_window(wav[s:e], 64000, "mean")) for (s, e, _, _) in sched])
    out = oa.audio_model_forward(syn.make_audio_state_dict(2, 8, "mid", 12, variant="v1"), torch.from_numpy(xs)).numpy()
    assert np.abs(out - golden["audio"]["a8_v1_window_logits"][:2]).max() < 5e-5


def test_audio_schedule_nan_window(golden):
    # L multiple of step_a: trailing empty window -> NaN logits for exactly one extra frame id
    g = golden["audio"]
    L, fps, pad, step = (int(v) for v in g["a8_b_meta"])
    assert L % int(step / 1000 * 16000) == 0
    wl = g["a8_b_window_logits"]
    assert np.isnan(wl[-1]).all() and not np.isnan(wl[:-1]).any()
    sched = oa.window_schedule(L, fps, step / 1000)
    assert sched[-1][0] == sched[-1][1] == L
    with pytest.raises(ZeroDivisionError):
        oa.pad_window(np.zeros(0, np.float32), 64000, "repeat")


def test_resample_matches_reference(golden):
    """convert_mp4_to_mp3's post-ffmpeg arithmetic (data/utils.py:49-60) on the int16 fixtures the reference itself
    converted (oracle/make_golden.py:make_resample), and the product's own filter-bank restatement against the oracle's."""
    from avcer_b200 import ops
    g = golden["resample"]
    for name in ("stereo_44100", "mono_48000"):
        pcm, sr, ref = g[name + "_pcm"], int(g[name + "_sr"]), g[name + "_out"]
        got = oa.pcm16_to_mono_16k(pcm, sr, 16000)
        assert got.shape == ref.shape and np.abs(got - ref).max() < 2e-6, name
        bank, orig, nnew, width = ops.sinc_resample_bank(sr, 16000)
        okern, owidth, oorig, onew = oa.sinc_resample_kernel(sr, 16000)
        assert (orig, nnew, width) == (oorig, onew, owidth) and np.array_equal(bank, okern), name


def test_pred_av_labels_match_reference(golden, tmp_path):
    """get_pred_av.get_c_expr_db_pred (get_pred_av.py:198-334): the restatement on the CSV tables against the labels the
    unmodified reference wrote for the same files (two clips, dropped frames, NaN audio rows, repeat-last-row tail)."""
    from oracle.make_golden import PRED_AV_CONFIGS, pred_av_tables, pred_av_weights, write_pred_av_files

    g = golden["pred_av"]
    tables, fmt = pred_av_tables()
    root = str(tmp_path / "preds")
    fmt_path, path_preds = write_pred_av_files(root, tables, fmt)
    read = {n: (pd.read_csv(f"{root}/video/static__{n}.csv"), pd.read_csv(f"{root}/video/dynamic__{n}.csv"),
                pd.read_csv(f"{root}/audio_mean_0.5/model/{n}.csv")) for n in tables}
    for i, (tag, w2, cwt, cm) in enumerate(PRED_AV_CONFIGS):
        w1, w2v = pred_av_weights(tag, w2)
        labels, locs = of.pred_av_labels(pd.read_csv(fmt_path), read, list(tables), w1, w2v, cwt, cm)
        assert len(locs) == int(g["n_locations"]) and np.array_equal(labels, g[f"labels_{i}"]), (tag, w2, cwt, cm)


def test_batch1_loops_match_batched_oracle(tmp_path):
    """oracle/loop.py (the reference's own frame-at-a-time / window-at-a-time execution shape, timed by bench.py as the
    "reference CPU path") against the batched restatements that make_golden pins to the unmodified reference."""
    import cv2

    from oracle import loop as ol

    n, fps = 13, 25
    missing = {0, 7}
    frames = syn.make_crops(5, n, 96)
    clip = tmp_path / "clip"
    (clip / "00").mkdir(parents=True)
    for i in range(n):
        if i not in missing:
            cv2.imwrite(str(clip / "00" / f"{i:06d}.jpg"), frames[i])
    sd_vs, sd_vd = syn.make_vs_state_dict(0, "spread"), syn.make_vd_state_dict(1)
    dyn, stat = ol.video_loop(str(clip), fps, n, sd_vs, sd_vd)
    decoded = [cv2.imread(str(clip / "00" / f"{i:06d}.jpg")) if i not in missing else None for i in range(n)]
    o_dyn, o_stat = ov.predict_video(decoded, fps, sd_vs, sd_vd)
    assert dyn.dtype == o_dyn.dtype == np.float64 and stat.shape == o_stat.shape == (n, 7)
    assert np.abs(stat - o_stat).max() < 1e-5 and np.abs(dyn - o_dyn).max() < 1e-4
    sd_a = syn.make_audio_state_dict(2, 8, "spread", 1)
    wav = syn.make_wav(9, 16000)
    rows, ids = ol.audio_loop(wav, fps, sd_a, step=0.5)
    o_rows, o_ids, _ = oa.predict_audio(wav, fps, sd_a, step=0.5)
    assert np.array_equal(ids, o_ids) and np.array_equal(np.isnan(rows), np.isnan(o_rows))
    assert np.nanmax(np.abs(rows - o_rows)) < 5e-5


JPEG_CASES = [(224, 224, 95), (97, 133, 95), (16, 16, 95), (8, 8, 50), (1, 1, 95), (17, 31, 75), (200, 301, 95), (33, 16, 50),
              (2, 2, 95), (15, 15, 95)]


def _jpeg_image(h, w):
    return np.ascontiguousarray(syn.make_crops(h * 1000 + w, 1, max(h, w))[0][:h, :w])


def test_jpeg_oracle_matches_cv2():
    """oracle/jpeg.py (libjpeg-turbo's baseline decoder restated: Huffman, islow IDCT, h2v2 fancy up-sampling, YCbCr -> BGR)
    against cv2.imdecode -- the decoder the reference itself calls (get_prob_video.py:95) -- bit for bit, on full, partial
    and tiny MCU grids, odd chroma widths, two qualities and 4:4:4."""
    import cv2

    from oracle import jpeg as oj

    for h, w, q in JPEG_CASES:
        ok, buf = cv2.imencode(".jpg", _jpeg_image(h, w), [cv2.IMWRITE_JPEG_QUALITY, q])
        assert ok and np.array_equal(oj.decode(buf.tobytes()), cv2.imdecode(buf, cv2.IMREAD_COLOR)), (h, w, q)
    ok, buf = cv2.imencode(".jpg", _jpeg_image(50, 70), [cv2.IMWRITE_JPEG_SAMPLING_FACTOR, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444])
    assert np.array_equal(oj.decode(buf.tobytes()), cv2.imdecode(buf, cv2.IMREAD_COLOR))


def test_jpeg_host_parser_matches_oracle_header():
    """avcer_b200.jpeg.parse (the product's own marker walk + byte un-stuffing, host side of avcer_jpeg_decode) against the
    oracle's header reader; files outside the GPU decoder's coverage are rejected, not mis-decoded."""
    import cv2

    from avcer_b200 import jpeg
    from oracle import jpeg as oj

    for h, w, q in JPEG_CASES[:6]:
        ok, buf = cv2.imencode(".jpg", _jpeg_image(h, w), [cv2.IMWRITE_JPEG_QUALITY, q])
        b = buf.tobytes()
        p, o = jpeg.parse(b), oj.parse(b)
        assert (p.width, p.height, p.hs) == (o.width, o.height, o.components[0][1])
        assert np.array_equal(p.qt_y, o.qt[o.components[0][3]]) and np.array_equal(p.qt_c, o.qt[o.components[1][3]])
        for t, key in enumerate([(0, 0), (1, 0), (0, 1), (1, 1)]):
            bits, vals = o.huff[key]
            assert np.array_equal(p.huff_bits[t], bits[1:]) and np.array_equal(p.huff_vals[t][: len(vals)], vals)
        assert bytes(p.data) == o.data and jpeg.unstuff(p.data) == o.data.replace(b"\xff\x00", b"\xff")
    img = _jpeg_image(40, 40)
    for flags in ([cv2.IMWRITE_JPEG_PROGRESSIVE, 1], [cv2.IMWRITE_JPEG_RST_INTERVAL, 4]):
        ok, buf = cv2.imencode(".jpg", img, flags)
        with pytest.raises(jpeg.UnsupportedJpeg):
            jpeg.parse(buf.tobytes())
    ok, buf = cv2.imencode(".jpg", img[:, :, 0])
    with pytest.raises(jpeg.UnsupportedJpeg):
        jpeg.parse(buf.tobytes())
    with pytest.raises(jpeg.UnsupportedJpeg):
        jpeg.parse(b"not a jpeg at all")


# ------------------------------------------------------------------------------------------ face detector
def test_face_oracle_matches_reference(golden):
    """oracle/face.py against what the reference's own RetinaFace / RetinaFacePredictor returned (tests/golden/face.npz):
    raw network outputs bit-identical (same torch ops in the same order), detections per frame bit-identical."""
    from oracle import face as ofa

    g = golden["face"]
    sd = syn.make_retinaface_state_dict(5, "spread")
    fr = syn.make_frames(40, 1, 100, 136)[0]
    loc, conf, lm = ofa.forward(ofa.prepare(fr), sd)
    assert np.array_equal(loc[0].numpy(), g["raw_loc"]) and np.array_equal(conf[0].numpy(), g["raw_conf"])
    assert np.array_equal(lm[0].numpy(), g["raw_landms"])
    assert loc.shape[1] == ofa.prior_box(100, 136).shape[0] == 2 * (13 * 17 + 7 * 9 + 4 * 5)
    frames = syn.make_frames(41, 6, 150, 200)
    tracker = ofa.SimpleFaceTracker(0.4, 0.0)
    ids = []
    for i in (0, 3, 5):
        assert np.array_equal(ofa.predict(sd, frames[i]), g[f"dets_{i}"]), i
    for i in range(6):
        ids += tracker(g[f"dets_{i}"])
    assert ids == list(g["ids"])
    assert np.array_equal(ofa.predict(sd, frames[0][..., ::-1].copy(), rgb=True), g["dets_0"])


def test_face_tracker_matches_reference(golden):
    from oracle import face as ofa
    from oracle.make_golden import face_tracker_sequences

    g = golden["face"]
    tracker = ofa.SimpleFaceTracker(0.4, 0.0)
    for s, seq in enumerate(face_tracker_sequences()):
        tracker.reset()
        got = []
        for boxes in seq:
            got += [-1 if v is None else v for v in tracker(boxes)]
        assert got == list(g[f"track_{s}"]), s
    assert -1 in list(g["track_0"])                                   # the zero-area box stays untracked


def test_face_nms_properties():
    """Greedy NMS (py_cpu_nms.py:11-39): survivors are mutually below the IoU threshold, every suppressed box overlaps a
    higher-scoring survivor (tie order is whatever numpy's default argsort yields: the host side calls the same function)."""
    from oracle import face as ofa

    rng = np.random.default_rng(0)
    xy = rng.uniform(0, 100, (300, 2))
    dets = np.concatenate([xy, xy + rng.uniform(5, 40, (300, 2)), rng.uniform(0, 1, (300, 1))], axis=1).astype(np.float32)
    dets[10, 4] = 2.0
    dets[11, :4] = dets[10, :4]                                       # an exact duplicate of the best box is suppressed
    keep = ofa.nms(dets, 0.4, 5000)
    assert keep[0] == 10 and 11 not in keep

    def iou(a, b):
        w = max(0.0, min(a[2], b[2]) - max(a[0], b[0]) + 1)
        h = max(0.0, min(a[3], b[3]) - max(a[1], b[1]) + 1)
        return w * h / ((a[2] - a[0] + 1) * (a[3] - a[1] + 1) + (b[2] - b[0] + 1) * (b[3] - b[1] + 1) - w * h)

    for i, a in enumerate(keep):
        for b in keep[:i]:
            assert iou(dets[a], dets[b]) <= 0.4
    for j in set(range(300)) - set(keep):
        assert any(iou(dets[j], dets[k]) > 0.4 and dets[k, 4] >= dets[j, 4] for k in keep)
    assert len(ofa.nms(dets, 0.4, 5)) <= 5                            # top_k truncates BEFORE suppression
