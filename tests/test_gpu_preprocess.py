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
