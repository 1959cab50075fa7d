"""Face detector on the GPU against the oracle restatement (oracle/face.py, pinned
bit-identically to the reference's RetinaFace / RetinaFacePredictor / SimpleFaceTracker / VideoPredictor.process) and the
fixtures the unmodified reference produced (tests/golden/face.npz)."""
import hashlib
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from avcer_b200 import synthetic as syn

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _frames_dev(fr):
    return torch.from_numpy(np.ascontiguousarray(fr)).to(DEV)


@pytest.mark.parametrize("h,w,rgb", [(100, 136, False), (33, 47, True), (7, 5, False), (64, 260, False)])
def test_det_stem_matches_torch(cuda_lib, h, w, rgb):
    """uint8 frame -> mean subtraction, conv 7x7/2 pad 3, bias, ReLU (retina_face_predictor.py:61-67 + resnet50 stem)."""
    from avcer_b200 import ops

    g = torch.Generator().manual_seed(h * 1000 + w)
    fr = torch.randint(0, 256, (3, h, w, 3), dtype=torch.uint8, generator=g)
    wt = torch.randn(64, 3, 7, 7, generator=g) * 0.01
    bias = torch.randn(64, generator=g) * 0.1
    bgr = fr.flip(-1) if rgb else fr
    x = (bgr.double() - torch.tensor([104.0, 117.0, 123.0], dtype=torch.float64)).permute(0, 3, 1, 2)
    ref = F.relu(F.conv2d(x, wt.double(), bias.double(), stride=2, padding=3)).permute(0, 2, 3, 1).float()
    wp = wt.permute(2, 3, 1, 0).reshape(147, 64).contiguous().to(DEV)
    got = ops.det_stem(fr.to(DEV), wp, bias.to(DEV), torch.float32, rgb).cpu()
    assert got.shape == ref.shape
    assert float((got - ref).abs().max()) <= 1e-4 * float(ref.abs().max())
    got16 = ops.det_stem(fr.to(DEV), wp, bias.to(DEV), torch.bfloat16, rgb).float().cpu()
    assert float((got16 - ref).abs().max()) <= 2.0 ** -8 * float(ref.abs().max())


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("h,w,rgb", [(100, 136, False), (33, 300, True), (7, 5, False), (61, 515, False)])
def test_det_stem_tensor_core_matches_torch(cuda_lib, dtype, h, w, rgb):
    """The stem as strip-mode tcgen05 contractions over the zero-bordered NHWC4 copy of the frames (avcer_det_prepare +
    avcer_contract a_strip, one call per band of 128 output columns: 300 -> 150 = 128 + 22, 515 -> 258 = 128 + 128 + 2):
    the mean-subtracted pixels are exact in 16 bits, so against a torch conv with the SAME 16-bit-rounded filters only the
    summation order and the output rounding differ."""
    from avcer_b200 import ops, weights

    g = torch.Generator().manual_seed(h * 1000 + w)
    fr = torch.randint(0, 256, (2, h, w, 3), dtype=torch.uint8, generator=g)
    wt = (torch.randn(64, 3, 7, 7, generator=g) * 0.01).to(dtype).float()
    bias = torch.randn(64, generator=g) * 0.1
    bgr = fr.flip(-1) if rgb else fr
    x = (bgr.double() - torch.tensor([104.0, 117.0, 123.0], dtype=torch.float64)).permute(0, 3, 1, 2)
    ref = F.relu(F.conv2d(x, wt.double(), bias.double(), stride=2, padding=3)).permute(0, 2, 3, 1).float()
    stem = torch.zeros(64, 7, 8, 4)
    stem[:, :, :7, :3] = wt.permute(0, 2, 3, 1)
    packed = stem.reshape(8, 8, 7, 4, 8).permute(2, 3, 0, 1, 4).contiguous().reshape(-1)
    got = ops.det_stem_tc(fr.to(DEV), stem.reshape(64, 224).to(DEV, dtype), packed.to(DEV, dtype), bias.to(DEV), dtype, rgb).float().cpu()
    assert got.shape == ref.shape
    ulp = 2.0 ** -8 if dtype == torch.bfloat16 else 2.0 ** -11
    assert bool(((got - ref).abs() <= ulp * ref.abs() + 1e-4 * float(ref.abs().max())).all()), float((got - ref).abs().max())


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
def test_maxpool_pad1_and_upsample_add(cuda_lib, dtype):
    from avcer_b200 import ops

    g = torch.Generator().manual_seed(1)
    for (h, w) in ((50, 68), (13, 17), (1, 1), (2, 7)):
        x = torch.randn(2, h, w, 64, generator=g).to(dtype)
        ref = F.max_pool2d(x.float().permute(0, 3, 1, 2), 3, 2, 1).permute(0, 2, 3, 1)
        got = ops.maxpool3x3s2p1(x.to(DEV)).float().cpu()
        assert torch.equal(got, ref)
    for (ha, wa, hb, wb) in ((13, 17, 7, 9), (7, 9, 4, 5), (10, 10, 5, 5), (19, 25, 10, 13)):
        a = torch.randn(2, ha, wa, 256, generator=g).to(dtype)
        b = torch.randn(2, hb, wb, 256, generator=g).to(dtype)
        up = F.interpolate(b.float().permute(0, 3, 1, 2), size=(ha, wa), mode="nearest").permute(0, 2, 3, 1)
        ref = (a.float() + up).to(dtype).float()
        got = ops.upsample_add(a.to(DEV), b.to(DEV), ops.nearest_source_index(hb, ha).to(DEV), ops.nearest_source_index(wb, wa).to(DEV))
        assert torch.equal(got.float().cpu(), ref)


def test_det_decode_matches_oracle(cuda_lib):
    """Anchors, softmax and decoding (prior_box.py, box_utils.py:210-249, retina_face_predictor.py:75-84) on random head
    outputs: every row within a few ulp of the torch arithmetic of the reference (only expf differs)."""
    from avcer_b200 import ops
    from oracle import face as ofa

    g = torch.Generator().manual_seed(2)
    for (h, w) in ((100, 136), (150, 200), (64, 64), (1080 // 4, 1920 // 4)):
        fhw = [(-(-h // s), -(-w // s)) for s in (8, 16, 32)]
        n = 2
        heads = [torch.randn(n * fh * fw, 64, generator=g) * 1.5 for fh, fw in fhw]
        dets = ops.det_decode([t.to(DEV) for t in heads], n, h, w).cpu()
        priors = ofa.prior_box(h, w)
        for b in range(n):
            cols = [t.view(n, -1, 64)[b] for t in heads]
            cls = torch.cat([c[:, 0:4].reshape(-1, 2) for c in cols])
            loc = torch.cat([c[:, 4:12].reshape(-1, 4) for c in cols])
            lmk = torch.cat([c[:, 12:32].reshape(-1, 10) for c in cols])
            boxes = ofa.decode(loc, priors) * torch.tensor([w, h, w, h], dtype=torch.float32)
            lm = ofa.decode_landm(lmk, priors) * torch.tensor([w, h] * 5, dtype=torch.float32)
            score = F.softmax(cls, dim=-1)[:, 1]
            ref = torch.cat([boxes, score[:, None], lm], dim=1)
            assert dets[b].shape == ref.shape
            # boxes: 1-2 ulp of expf on widths of up to ~1000 px, then x1 = cx - w/2 (cancellation): absolute, in pixels
            err = (dets[b] - ref).abs()
            assert float((err - 2e-6 * ref.abs()).max()) < 5e-4, float(err.max())
            assert float((dets[b][:, 4] - ref[:, 4]).abs().max()) < 3e-7
            assert torch.equal(dets[b][:, 5:], ref[:, 5:])          # no transcendental in the landmark path: bit-identical


def _rel(a, b):
    return float((a - b).abs().max()) / max(float(b.abs().max()), 1e-6)


@pytest.mark.parametrize("precision,tol", [("fp32", 2e-4), ("bf16", 6e-2), ("fp16", 1e-2)])
def test_retinaface_heads_match_oracle(cuda_lib, precision, tol):
    """Whole network (ResNet-50 body, FPN, SSH, heads) on a 100 x 136 frame (odd map sizes 13 x 17, 7 x 9, 4 x 5) against the
    oracle forward, layer taps included.  fp32: SIMT fp32 kernels; bf16 / fp16: tcgen05 path (error relative to the range
    of each tensor; the class / box / landmark regressions of this init span ~ +-10)."""
    from avcer_b200 import nets
    from oracle import face as ofa

    sd = syn.make_retinaface_state_dict(5, "spread")
    fr = syn.make_frames(40, 2, 100, 136)
    taps_ref, taps = {}, {}
    x = torch.cat([ofa.prepare(f) for f in fr])
    loc, conf, lm = ofa.forward(x, sd, taps_ref)
    net = nets.RetinaFaceNet(sd, precision, DEV)
    heads = [t.cpu() for t in net.heads(_frames_dev(fr), False, taps)]
    for name in ("stem", "pool", "layer1", "layer2", "layer3", "layer4", "fpn1", "fpn2", "fpn3", "ssh1", "ssh2", "ssh3"):
        got = taps[name].float().cpu().permute(0, 3, 1, 2)
        assert got.shape == taps_ref[name].shape, name
        assert _rel(got, taps_ref[name]) < tol, (name, _rel(got, taps_ref[name]))
    cls = torch.cat([t.view(2, -1, 64)[:, :, 0:4].reshape(2, -1, 2) for t in heads], dim=1)
    box = torch.cat([t.view(2, -1, 64)[:, :, 4:12].reshape(2, -1, 4) for t in heads], dim=1)
    lmk = torch.cat([t.view(2, -1, 64)[:, :, 12:32].reshape(2, -1, 10) for t in heads], dim=1)
    assert _rel(cls, taps_ref["cls_logits"]) < tol and _rel(box, loc) < tol and _rel(lmk, lm) < tol
    assert all(float(t[:, 32:].abs().max()) == 0.0 for t in heads)


def test_predictor_matches_reference_detections(cuda_lib, golden):
    """RetinaFacePredictor.__call__ / detect_batch in fp32 against what the unmodified reference returned for the same frames
    (same detections in the same order; coordinates to 1e-3 px, scores to 1e-5)."""
    from types import SimpleNamespace

    from avcer_b200.data.face_detection import RetinaFacePredictor, cfg_re50

    g = golden["face"]
    sd = syn.make_retinaface_state_dict(5, "spread")
    pred = RetinaFacePredictor(threshold=0.8, device=DEV, model=SimpleNamespace(weights=sd, config=SimpleNamespace(**cfg_re50)),
                               precision="fp32")
    frames = syn.make_frames(41, 6, 150, 200)
    batch = pred.detect_batch(frames, rgb=False)
    for i in range(6):
        ref = g[f"dets_{i}"]
        assert batch[i].shape == ref.shape and batch[i].dtype == np.float32, (i, batch[i].shape, ref.shape)
        assert np.abs(batch[i][:, 4] - ref[:, 4]).max() < 1e-5
        assert np.abs(batch[i] - ref).max() < 2e-3
    one = pred(frames[0][..., ::-1].copy(), rgb=True)
    assert np.array_equal(one, batch[0])
    none = RetinaFacePredictor(threshold=1.1, device=DEV, model=SimpleNamespace(weights=sd, config=SimpleNamespace(**cfg_re50)),
                               precision="fp32")(frames[0], rgb=False)
    assert none.shape == (0, 15) and none.dtype == np.float32


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_video_predictor_process_matches_reference(cuda_lib, golden, tmp_path, precision):
    """VideoPredictor.process on the MJPG clip the fixture was made from: the same crop files (track directory / frame
    number) as the unmodified reference wrote; in fp32 the JPEG bytes of the crops too."""
    from types import SimpleNamespace

    import cv2

    from avcer_b200.data.face_detection import cfg_re50
    from avcer_b200.data.get_face_images import VideoPredictor
    from oracle.make_golden import write_face_video

    g = golden["face"]
    vpath = str(tmp_path / "faces_clip.avi")
    write_face_video(vpath, syn.make_frames(42, 8, 150, 200))
    cap = cv2.VideoCapture(vpath)
    decoded = []
    while True:
        ok, f = cap.read()
        if not ok:
            break
        decoded.append(f)
    cap.release()
    sha = np.frombuffer(hashlib.sha256(np.stack(decoded).tobytes()).digest(), dtype=np.uint8)
    assert np.array_equal(sha, g["video_frames_sha256"]), "this cv2 build encodes / decodes MJPG differently from the fixture's"
    sd = syn.make_retinaface_state_dict(5, "spread")
    vp = VideoPredictor(batch=3, model=SimpleNamespace(weights=sd, config=SimpleNamespace(**cfg_re50)), precision=precision)
    save = str(tmp_path / "out")
    vp.process(vpath, save)
    assert vp.count_frame == 8 and (vp.w, vp.h) == (200, 150)
    names = []
    for root, _, files in sorted(os.walk(os.path.join(save, "faces_clip"))):
        names += [os.path.relpath(os.path.join(root, fn), save).replace(os.sep, "/") for fn in sorted(files)]
    assert names == list(g["process_files"])
    if precision == "fp32":
        same = sum(np.array_equal(np.frombuffer(hashlib.sha256(open(os.path.join(save, n), "rb").read()).digest(), dtype=np.uint8), d)
                   for n, d in zip(names, g["process_sha256"]))
        assert same == len(names), (same, len(names))


@pytest.mark.parametrize("cfg", [dict(threshold=0.5, top_k=2, nms_top_k=5000, conf_thresh=0.02, nms_thresh=0.4),
                                 dict(threshold=0.6, top_k=750, nms_top_k=3, conf_thresh=0.02, nms_thresh=0.4),
                                 dict(threshold=0.01, top_k=750, nms_top_k=5000, conf_thresh=0.3, nms_thresh=0.1),
                                 dict(threshold=0.9, top_k=750, nms_top_k=5000, conf_thresh=0.02, nms_thresh=0.9)])
def test_predictor_config_variants_match_oracle(cuda_lib, cfg):
    """create_config / threshold variants (retina_face_predictor.py:55-58, 86-109): the device only hands over boxes with
    score > conf_thresh AND >= threshold and the host runs NMS on those -- equivalent to the reference's order (NMS over
    everything above conf_thresh, truncations, then the threshold) because suppression and both truncations act from the
    high-score end.  Checked for truncating top_k / nms_top_k, a threshold BELOW conf_thresh, loose and tight NMS."""
    from types import SimpleNamespace

    from avcer_b200.data.face_detection import RetinaFacePredictor, cfg_re50
    from oracle import face as ofa

    sd = syn.make_retinaface_state_dict(5, "spread")
    frame = syn.make_frames(43, 1, 150, 200)[0]
    loc, conf, lm = ofa.forward(ofa.prepare(frame), sd)
    ref = ofa.postprocess(loc[0], conf[0], lm[0], 150, 200, cfg["threshold"], cfg["conf_thresh"], cfg["nms_thresh"], cfg["nms_top_k"], cfg["top_k"])
    pred = RetinaFacePredictor(threshold=cfg["threshold"], device=DEV, model=SimpleNamespace(weights=sd, config=SimpleNamespace(**cfg_re50)),
                               config=RetinaFacePredictor.create_config(top_k=cfg["top_k"], conf_thresh=cfg["conf_thresh"],
                                                                        nms_thresh=cfg["nms_thresh"], nms_top_k=cfg["nms_top_k"]),
                               precision="fp32")
    got = pred(frame, rgb=False)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    if len(ref):
        assert np.abs(got[:, 4] - ref[:, 4]).max() < 1e-5 and np.abs(got - ref).max() < 2e-3
