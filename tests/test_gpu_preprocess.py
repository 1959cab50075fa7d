"""GPU parity of K1 (face-crop preprocessing) through the C ABI: bit-exact against the oracle
restatement of data/utils.py:19-39 (itself pinned to PIL/torchvision by tests/golden/preprocess.npz)."""
import hashlib

import numpy as np
import pytest
import torch

from oracle import video as ov

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _pack(imgs):
    offs, hs, ws, parts, off = [], [], [], [], 0
    for im in imgs:
        offs.append(off)
        hs.append(im.shape[0])
        ws.append(im.shape[1])
        parts.append(im.reshape(-1))
        off += (im.size + 15) // 16 * 16
    flat = np.zeros(off, np.uint8)
    for p, o in zip(parts, offs):
        flat[o:o + p.size] = p
    t = lambda a, dt: torch.from_numpy(np.asarray(a, dtype=dt)).to(DEV)
    return torch.from_numpy(flat).to(DEV), t(offs, np.int64), t(hs, np.int32), t(ws, np.int32)


def test_k1_ragged_bit_exact_and_digests(cuda_lib, golden):
    from avcer_b200 import ops

    g = golden["preprocess"]
    rng = np.random.default_rng(1)
    imgs = [rng.integers(0, 256, (int(h), int(w), 3), dtype=np.uint8) for (h, w) in g["shapes"]]
    imgs += [rng.integers(0, 256, (s, t, 3), dtype=np.uint8) for (s, t) in ((1, 1), (3, 1500), (1300, 5), (223, 225))]
    flat, offs, hs, ws = _pack(imgs)
    out = torch.empty((len(imgs), 3, 224, 224), device=DEV)
    ops.preprocess(flat, len(imgs), out, 0, offsets=offs, heights=hs, widths=ws)
    out = out.cpu().numpy()
    for i, im in enumerate(imgs):
        assert np.array_equal(out[i], ov.pth_processing(im)[0]), im.shape
    for i, d in enumerate(g["digests"]):
        assert hashlib.sha256(out[i:i + 1].tobytes()).hexdigest() == str(d)


def test_k1_packed_layouts(cuda_lib):
    from avcer_b200 import ops, synthetic as syn

    crops = syn.make_crops(3, 5)
    src = torch.from_numpy(crops).to(DEV)
    nchw = torch.empty((5, 3, 224, 224), device=DEV)
    ops.preprocess(src, 5, nchw, 0)
    ref = np.concatenate([ov.pth_processing(c) for c in crops])
    assert np.array_equal(nchw.cpu().numpy(), ref)
    for layout, dt in ((1, torch.bfloat16), (2, torch.float32)):
        pad = torch.zeros((5, 232, 240, 4), device=DEV, dtype=dt)
        ops.preprocess(src, 5, pad, layout)
        inner = pad[:, 2:226, 2:226, :3].permute(0, 3, 1, 2).float().cpu()
        assert torch.equal(inner, torch.from_numpy(ref).to(dt).float())          # one rounding, nothing else
        border = pad.clone()
        border[:, 2:226, 2:226, :3] = 0
        assert border.abs().max().item() == 0                                     # zero border and zero 4th channel
    # empty batch is a no-op
    ops.preprocess(src[:0], 0, nchw, 0)


def test_pcm16_resample_matches_reference(cuda_lib, golden, tmp_path):
    """Audio decode seam (data/utils.py:49-60): int16 PCM -> mono -> 16 kHz on the GPU against what the reference's own
    convert_mp4_to_mp3 produced from the same .wav payloads (tests/golden/resample.npz), through the drop-in function
    (reads the .wav next to the video path) and against the oracle on a longer seeded signal with three channels."""
    import wave

    from avcer_b200 import ops
    from avcer_b200.data.utils import convert_mp4_to_mp3
    from oracle import audio as oa

    g = golden["resample"]
    for name in ("stereo_44100", "mono_48000"):
        pcm, sr, ref = g[name + "_pcm"], int(g[name + "_sr"]), g[name + "_out"]
        with wave.open(str(tmp_path / (name + ".wav")), "wb") as f:
            f.setnchannels(pcm.shape[1]); f.setsampwidth(2); f.setframerate(sr)
            f.writeframes(pcm.astype("<i2").tobytes())
        got = convert_mp4_to_mp3(str(tmp_path / (name + ".mp4")), 16000)
        assert got.dtype == torch.float32 and tuple(got.shape) == ref.shape
        assert np.abs(got.numpy() - ref).max() < 2e-6, name
    rng = np.random.default_rng(5)
    for sr, n, ch in ((44100, 200003, 3), (22050, 70001, 1), (8000, 33333, 2), (16000, 5000, 2), (44100, 441, 1), (44100, 1, 1)):
        pcm = rng.integers(-20000, 20000, (n, ch)).astype(np.int16)
        want = oa.pcm16_to_mono_16k(pcm, sr, 16000)
        got = ops.pcm16_to_mono(torch.from_numpy(pcm).to(DEV), sr, 16000).cpu().numpy()
        assert got.shape == want.shape, (sr, n, ch)
        assert np.abs(got - want).max() < 3e-6, (sr, n, ch, np.abs(got - want).max())
    with pytest.raises(FileNotFoundError):
        convert_mp4_to_mp3(str(tmp_path / "absent.mp4"), 16000)


def test_jpeg_decode_bit_identical_to_cv2(cuda_lib):
    """avcer_jpeg_decode against cv2.imdecode (the reference's decoder, get_prob_video.py:95): one ragged batch mixing
    sizes (full / partial / single MCUs, odd chroma widths), qualities (different quantisation tables) and 4:4:4 files;
    a second batch of 300 crops 224x224 (more images than one warp, the packed BASELINE shape)."""
    import cv2

    from avcer_b200 import jpeg, synthetic as syn

    cases = [(224, 224, 95), (97, 133, 95), (16, 16, 95), (8, 8, 50), (1, 1, 95), (17, 31, 75), (200, 301, 95), (33, 16, 50),
             (2, 2, 95), (15, 15, 95), (480, 640, 90), (225, 223, 95)]
    files, refs = [], []
    for h, w, q in cases:
        img = np.ascontiguousarray(syn.make_crops(h * 1000 + w, 1, max(h, w))[0][:h, :w])
        for extra in ([], [cv2.IMWRITE_JPEG_SAMPLING_FACTOR, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444]):
            ok, buf = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, q] + extra)
            files.append(buf.tobytes())
            refs.append(cv2.imdecode(buf, cv2.IMREAD_COLOR))
    got = jpeg.decode_images(files, "cuda:0")
    for g, r, c in zip(got, refs, [c for c in cases for _ in range(2)]):
        assert g.shape == r.shape and np.array_equal(g.cpu().numpy(), r), c
    crops = syn.make_crops(5, 300)
    files = [cv2.imencode(".jpg", c)[1].tobytes() for c in crops]
    out, off, hs, ws = jpeg.decode_batch(files, "cuda:0", align_out=1)
    assert (hs == 224).all() and (ws == 224).all() and off[1] == 224 * 224 * 3
    dec = out[: 300 * 224 * 224 * 3].view(300, 224, 224, 3).cpu().numpy()
    for i in (0, 1, 31, 32, 33, 150, 299):
        assert np.array_equal(dec[i], cv2.imdecode(np.frombuffer(files[i], np.uint8), cv2.IMREAD_COLOR)), i
    with pytest.raises(jpeg.UnsupportedJpeg):
        jpeg.decode_batch([cv2.imencode(".jpg", crops[0], [cv2.IMWRITE_JPEG_PROGRESSIVE, 1])[1].tobytes()], "cuda:0")
