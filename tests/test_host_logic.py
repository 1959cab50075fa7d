"""CPU: host-side index plans, weight packer and the C-ABI surface (no compute calls)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from avcer_b200 import _lib, pipeline, synthetic as syn, weights
from oracle import audio as oa
from oracle import video as ov

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_plan_video_matches_frame_loop():
    rng = np.random.default_rng(0)
    for _ in range(400):
        n = int(rng.integers(1, 90))
        step = int(rng.integers(1, 13))
        ex = rng.random(n) >= rng.choice([0.0, 0.05, 0.3, 0.8, 1.0])
        a = pipeline.plan_video(ex, step)
        s, w, ss, ds = ov.plan_video(list(ex), step)
        assert np.array_equal(a.samples, s) and np.array_equal(a.windows, w)
        assert np.array_equal(a.stat_src, ss) and np.array_equal(a.dyn_src, ds)


def test_plan_video_edge_cases():
    p = pipeline.plan_video(np.zeros(7, bool), 5)          # no crop at all: all zero rows, no windows
    assert p.windows.shape == (0, 10) and (p.stat_src == -1).all() and (p.dyn_src == -1).all()
    p = pipeline.plan_video(np.ones(11, bool), 5)
    assert p.samples.tolist() == [0, 5, 10] and p.windows[0].tolist() == [0] * 10 and p.windows[2].tolist() == [0] * 8 + [1, 2]
    assert pipeline.vd_step(25) == 5 and pipeline.vd_step(30) == 6 and pipeline.vd_step(24) == 5 and pipeline.vd_step(60) == 12


def test_plan_audio_matches_window_loop():
    for L in [0, 1, 7999, 8000, 64000, 160000, 159840, 960000, 123457]:
        for fps in [25, 30, 24, 29.97, 60]:
            for step in [0.5, 1]:
                ap = pipeline.plan_audio(L, fps, step)
                sch = oa.window_schedule(L, fps, step)
                got = [tuple(int(v) for v in x) for x in zip(ap.starts, ap.ends, ap.f_lo, ap.f_hi)]
                assert got == [tuple(x) for x in sch]
    # C1 of BASELINE.json: 10 s clip -> 21 windows, the last one empty
    ap = pipeline.plan_audio(160000, 25)
    assert len(ap.starts) == 21 and ap.starts[-1] == ap.ends[-1] == 160000


def test_header_symbols_exported():
    """include/avcer_b200.h, the dynamic symbol table of the built library (`nm -D`) and the ctypes table must name
    exactly the same entry points: nothing undeclared is exported, nothing declared is missing."""
    import subprocess

    hdr = open(os.path.join(ROOT, "include", "avcer_b200.h")).read()
    declared = set(re.findall(r"\b(avcer_[a-z0-9_]+)\s*\(", hdr))
    declared.discard("avcer_contract_desc")
    assert declared, "no declarations found"
    for path in (_lib.LIB_PATH, _lib.LIB_PATH_FP16):          # the bf16 build and the half build export the same surface
        nm = subprocess.run(["nm", "-D", "--defined-only", path], check=True, capture_output=True, text=True).stdout
        exported = {line.split()[-1] for line in nm.splitlines() if line.split()[-1].startswith("avcer_") and " T " in line}
        assert exported == declared, f"header vs nm -D of {path}: {sorted(exported ^ declared)}"
    assert _lib.load("bf16").avcer_storage_type() == b"bf16" and _lib.load("fp16").avcer_storage_type() == b"fp16"
    assert declared == set(_lib.exported_symbols()), declared ^ set(_lib.exported_symbols())
    assert _lib.load().avcer_version() >= 100


def test_no_cpu_fallback_without_device():
    if torch.cuda.is_available():
        pytest.skip("a device is visible")
    with pytest.raises(_lib.AvcerError):
        _lib.require_device()
    from avcer_b200 import ops

    with pytest.raises(_lib.AvcerError):
        ops.softmax7(torch.zeros(4, 7))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "avcer_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_bn_fold_and_tap_major_layout():
    sd = syn.make_vs_state_dict(0, "spread")
    w = weights.pack_vs(sd, "cpu", torch.float32)
    x = torch.randn(2, 64, 9, 9)
    blk = w["blocks"][0]
    ref = F.batch_norm(F.conv2d(x, sd["layer1.0.conv2.weight"], padding=1), sd["layer1.0.batch_norm2.running_mean"],
                       sd["layer1.0.batch_norm2.running_var"], sd["layer1.0.batch_norm2.weight"], sd["layer1.0.batch_norm2.bias"],
                       False, eps=1e-3)
    w4 = blk["conv2"].wt.view(64, 3, 3, 64).permute(0, 3, 1, 2)
    got = F.conv2d(x, w4, blk["conv2"].bias, padding=1)
    assert (got - ref).abs().max() < 1e-4
    # stem layout [64][7 rows][8 px][4 ch] with zero pixel 7 / channel 3
    st = w["stem"].wt.view(64, 7, 8, 4)
    assert st[:, :, 7].abs().max() == 0 and st[:, :, :, 3].abs().max() == 0
    assert len(w["blocks"]) == 16 and sum("ds" in b for b in w["blocks"]) == 4


def test_audio_pack_shapes():
    sd = syn.make_audio_state_dict(2, 8, "spread", 2)
    w = weights.pack_audio(sd, "cpu", torch.float32)
    assert w["pos_w"].shape == (1024, 128 * 64) and w["convs"][0][0].shape == (512, 3 * 512)
    assert w["layers"][0]["wqkv"].shape == (3072, 1024) and w["td0_w"].shape == (1024, 5 * 1024) and w["num_classes"] == 8
    # weight-norm fold equals the oracle's
    eff = oa.pos_conv_weight(sd)
    assert (w["pos_w"].view(1024, 128, 64).permute(0, 2, 1) - eff).abs().max() < 1e-6


def test_audio_pack_accepts_every_weight_norm_spelling():
    """The positional conv's weight norm as saved by current torch (parametrizations), by older torch / transformers
    (weight_g / weight_v) and after remove_weight_norm (plain weight); anything else is a descriptive KeyError."""
    sd = syn.make_audio_state_dict(2, 8, "spread", 1)
    pre = "wav2vec2.encoder.pos_conv_embed.conv"
    ref = weights.pack_audio(sd, "cpu", torch.float32)["pos_w"]
    old = dict(sd)
    old[pre + ".weight_g"] = old.pop(pre + ".parametrizations.weight.original0")
    old[pre + ".weight_v"] = old.pop(pre + ".parametrizations.weight.original1")
    assert torch.equal(weights.pack_audio(old, "cpu", torch.float32)["pos_w"], ref)
    plain = {k: v for k, v in sd.items() if "parametrizations" not in k}
    plain[pre + ".weight"] = oa.pos_conv_weight(sd)
    assert (weights.pack_audio(plain, "cpu", torch.float32)["pos_w"] - ref).abs().max() < 2e-6      # fp32 vs fp64 fold
    bad = {k: v for k, v in sd.items() if "parametrizations" not in k}
    with pytest.raises(KeyError, match="weight_g"):
        weights.pack_audio(bad, "cpu", torch.float32)


def test_audio_table_facade_materialises_the_reference_table():
    """tables.AudioTable: compact (window logits + frame ranges) until somebody looks; then exactly the long-format
    DataFrame of get_prob_audio_8_cl.py:94-126 (one row per (window, covered frame), `frames` strings, column order)."""
    import pandas as pd

    from avcer_b200.tables import AudioTable

    ap = pipeline.plan_audio(52923, 25, 0.5)
    rng = np.random.default_rng(0)
    logits = rng.standard_normal((len(ap.starts), 8)).astype(np.float32)
    logits[-1] = np.nan
    cols = pipeline.AUDIO_ORDER
    t = AudioTable(torch.from_numpy(logits), ap.f_lo, ap.f_hi, cols)
    rows, frames = [], []
    for (s0, e0, lo, hi), l in zip(oa.window_schedule(52923, 25, 0.5), logits):          # the reference's loop shape
        for f in range(lo, hi):
            rows.append(l)
            frames.append(str(f).zfill(6) + ".jpg")
    assert len(t) == len(rows) and not t.materialized and "compact" in repr(t)
    assert t.frame_ids().tolist() == sorted({int(f[:-4]) for f in frames})
    assert not t.materialized                                                             # still no string built
    assert list(t.columns) == cols + ["frames"] and t.materialized
    want = pd.DataFrame(np.asarray(rows), columns=cols)
    want["frames"] = frames
    pd.testing.assert_frame_equal(t.materialize(), want)
    assert t["frames"].tolist() == frames and t.shape == want.shape
    pd.testing.assert_frame_equal(t.groupby(["frames"]).mean().reset_index(), want.groupby(["frames"]).mean().reset_index())
    t["image_location"] = 1                                                               # consumers may add columns (run.py:92)
    assert "image_location" in t.materialize().columns


def test_balanced_batches_cover_and_differ_by_one():
    from avcer_b200.pipeline import balanced_batches

    assert balanced_batches(0, 64) == []
    for n, mb in ((968, 64), (12000, 1024), (1, 64), (64, 64), (65, 64), (250, 256), (1500, 256)):
        r = balanced_batches(n, mb)
        assert r[0][0] == 0 and r[-1][1] == n and all(a[1] == b[0] for a, b in zip(r, r[1:]))
        sizes = [e - s for s, e in r]
        assert len(r) == -(-n // mb) and max(sizes) <= mb and max(sizes) - min(sizes) <= 1


def test_contract_descriptor_layout_matches_the_header(tmp_path):
    """The ctypes mirror of avcer_contract_desc must have the size and field offsets the C compiler gives the header's
    struct (a field appended on one side only would silently shift every argument)."""
    import ctypes
    import subprocess

    from avcer_b200._lib import ContractDesc

    fields = [f[0] for f in ContractDesc._fields_]
    src = tmp_path / "layout.c"
    body = "\n".join(f'  printf("{name} %zu\\n", offsetof(avcer_contract_desc, {name}));' for name in fields)
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "avcer_b200.h"\nint main(void) {\n'
                   '  printf("sizeof %zu\\n", sizeof(avcer_contract_desc));\n' + body + "\n  return 0;\n}\n")
    exe = tmp_path / "layout"
    inc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include")
    subprocess.run(["gcc", "-I", inc, str(src), "-o", str(exe)], check=True)
    out = dict(line.split() for line in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    assert int(out["sizeof"]) == ctypes.sizeof(ContractDesc)
    for name in fields:
        assert int(out[name]) == getattr(ContractDesc, name).offset, name


# ------------------------------------------------------------------------------------------ face detector host side
def test_face_tracker_and_nms_product_code_match_reference(golden):
    """The product's host-side SimpleFaceTracker and greedy NMS (avcer_b200/data/face_detection.py) against the track ids
    the unmodified reference produced (tests/golden/face.npz) and against the oracle's restatement of py_cpu_nms."""
    from avcer_b200.data.face_detection import SimpleFaceTracker, greedy_nms
    from oracle import face as ofa
    from oracle.make_golden import face_tracker_sequences

    g = golden["face"]
    t = SimpleFaceTracker(iou_threshold=0.4, minimum_face_size=0.0)
    for s, seq in enumerate(face_tracker_sequences()):
        t.reset()
        got = []
        for boxes in seq:
            got += [-1 if v is None else v for v in t(boxes)]
        assert got == list(g[f"track_{s}"]), s
    t.reset()
    ids = []
    for i in range(6):
        ids += t(g[f"dets_{i}"])
    assert ids == list(g["ids"])
    assert t(np.empty((0, 15), dtype=np.float32)) == [] and t._tracklets == []
    t.iou_threshold, t.minimum_face_size = 0.5, 3.0
    assert (t.iou_threshold, t.minimum_face_size) == (0.5, 3.0)
    rng = np.random.default_rng(1)
    for trial in range(30):
        n = int(rng.integers(1, 700))
        xy = rng.uniform(0, 300, (n, 2))
        d = np.concatenate([xy, xy + rng.uniform(5, 60, (n, 2)), rng.uniform(0, 1, (n, 1))], axis=1).astype(np.float32)
        if trial % 5 == 0:
            d[n // 2, 2] = np.nan
        for k in (5000, 9):
            assert greedy_nms(d, 0.4, k) == ofa.nms(d, 0.4, k), (trial, k)


def test_nearest_source_index_matches_torch_interpolate():
    """FPN top-down merge (retina_face_net.py:88-94): the host computes the source row / column of every output position with
    ATen's nearest rule; checked against F.interpolate itself on an index ramp for every size pair the pyramid can produce."""
    import torch.nn.functional as F

    from avcer_b200 import ops

    for n_in in list(range(1, 40)) + [68, 135, 240]:
        for n_out in range(n_in, min(2 * n_in + 2, 500)):
            ref = F.interpolate(torch.arange(n_in, dtype=torch.float32).view(1, 1, n_in, 1), size=(n_out, 1), mode="nearest").view(-1).long()
            assert torch.equal(ref, ops.nearest_source_index(n_in, n_out).long()), (n_in, n_out)


def test_retinaface_packer_layout():
    """weights.pack_retinaface: BatchNorm folded with eps 1e-5, stem rows ordered (ky, kx, c), heads concatenated
    [class 4 | box 8 | landmarks 20 | zeros], 16 bottlenecks with the stride on conv2."""
    from avcer_b200 import weights

    sd = syn.make_retinaface_state_dict(5, "spread")
    w = weights.pack_retinaface(sd, "cpu", torch.float32)
    s = sd["body.bn1.weight"].double() / torch.sqrt(sd["body.bn1.running_var"].double() + 1e-5)
    want = (sd["body.conv1.weight"].double() * s.view(-1, 1, 1, 1))[5, 2, 3, 4]
    assert abs(float(w["stem_w"][(3 * 7 + 4) * 3 + 2, 5]) - float(want)) < 1e-9
    assert len(w["blocks"]) == 16 and [b["conv2"].stride for b in w["blocks"]] == [1, 1, 1, 2, 1, 1, 1, 2, 1, 1, 1, 1, 1, 2, 1, 1]
    h = w["heads"][1]
    assert h.wt.shape == (64, 256) and float(h.wt[32:].abs().max()) == 0.0
    assert torch.equal(h.wt[0:4], sd["ClassHead.1.conv1x1.weight"].reshape(4, 256))
    assert torch.equal(h.wt[4:12], sd["BboxHead.1.conv1x1.weight"].reshape(8, 256))
    assert torch.equal(h.bias[12:32], sd["LandmarkHead.1.conv1x1.bias"])
    w16 = weights.pack_retinaface(sd, "cpu", torch.bfloat16)
    assert w16["stem_wt"].shape == (64, 224) and w16["stem_packed"].numel() == 64 * 224
