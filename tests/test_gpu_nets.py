"""GPU parity of the VS / VD / A device forwards (all contractions through avcer_contract) against
the oracle on identical seeded weights and inputs, and against the reference golden outputs.

Tolerances (per-class probabilities, absolute):
  fp32 mode : 1e-5   (north star)
  bf16 mode : 2e-3   (north star) on
                * PyTorch-default random init (the init the north star names; measured 2.4e-4), and
                * the "mid" init: He-normal convolutions, perturbed BN / LN statistics, heads scaled to a logit
                  range of ~1 -- input-dependent outputs (asserted below: the reference's own probabilities
                  vary across inputs by more than the tolerance);
              VD (LSTM): 2e-3 on its widest init -- every contraction of the recurrence is bf16x3 (split
              operands, fp32 accumulation), measured ~1e-5;
              the deliberately wide "spread" init (logit range ~5): VS 2e-3 (measured 1.6e-3 on the golden crops, since relu(fc1)
              leaves the network in fp32), A 1.2e-2 (measured 6.4e-3): the probability error is p(1-p) x the logit error, and the
              logit error of bf16 OPERANDS is ~0.65 % of the logit range whatever is done about storage -- bf16 weights alone
              cost 1.1e-3 (VS) / 2.5e-3 (A) at this init, an fp32 residual stream moves the audio total only from 6.8e-3 to
              4.5e-3 (scripts/sim_bf16_budget.py, CPU emulation of every rounding point; profiles/r02_bf16_error_budget.txt).
              What the audio network does assert on this init in bf16: identical arg-max wherever the reference's top-2 margin
              exceeds twice the bound, and >= 99.5 % compound top-1 agreement (test_gpu_dropin.py).
  fp16 mode : 2e-3 on EVERY init, the wide one included (measured VS 9.0e-4, A 1.0e-3): the same kernels built with IEEE half
              storage (libavcer_b200_fp16.so, -DAVCER_HALF) -- 11 mantissa bits instead of 8 at the same tensor-core rate.
"""
import numpy as np
import pytest
import torch

from avcer_b200 import synthetic as syn
from oracle import audio as oa
from oracle import video as ov

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = {("fp32", "spread"): 1e-5, ("fp32", "default"): 1e-5, ("fp32", "mid"): 1e-5,
       # VS in bf16 on the wide init: 1.6e-3 measured on these crops since relu(fc1) leaves the network in fp32 (6.1e-3 with
       # bf16 features in round 1); deterministic kernels make the figure reproducible, so the north-star bar is asserted
       ("bf16", "default"): 2e-3, ("bf16", "mid"): 2e-3, ("bf16", "spread"): 2e-3,
       # precision "fp16": the same kernels built with IEEE half storage (11 mantissa bits): the north-star 2e-3 on EVERY init
       ("fp16", "default"): 2e-3, ("fp16", "mid"): 2e-3, ("fp16", "spread"): 2e-3}


def _vs_probs(sd, prec, crops):
    from avcer_b200 import nets, ops

    net = nets.VSNet(sd, prec, DEV)
    x = net.alloc_input(len(crops))
    ops.preprocess(torch.from_numpy(crops).to(DEV), len(crops), x, net.input_layout)
    probs, feat = net.forward(x)
    return probs.cpu().numpy(), feat.float().cpu().numpy()


@pytest.mark.parametrize("prec", ["fp32", "bf16", "fp16"])
@pytest.mark.parametrize("init", ["spread", "default", "mid"])
def test_vs_matches_reference_golden(cuda_lib, golden, prec, init):
    g = golden["video"]
    crops = syn.make_crops(11, 6)
    probs, feat = _vs_probs(syn.make_vs_state_dict(0, init), prec, crops)
    ref = g[f"vs_{init}_probs"]
    err = np.abs(probs - ref).max()
    assert err < TOL[(prec, init)], err
    if init != "default":
        # not a degenerate case: the reference's own probabilities move with the input by more than the tolerance
        assert (ref.max(0) - ref.min(0)).max() > 2 * TOL[(prec, init)]
        top2 = np.sort(ref, axis=1)[:, -2:]
        sure = (top2[:, 1] - top2[:, 0]) > 2 * TOL[(prec, init)]
        assert np.array_equal(probs.argmax(1)[sure], ref.argmax(1)[sure])
    ref_feat = np.maximum(g[f"vs_{init}_feat"], 0)
    assert np.abs(feat - ref_feat).max() < {"fp32": 2e-4, "bf16": 0.15, "fp16": 0.02}[prec]


def test_vs_batch_invariance_and_tails(cuda_lib):
    """Size-independent property: a crop's output does not depend on its batch position or batch size
    (M-tile boxes, tail tiles and persistent scheduling must not leak between images)."""
    crops = syn.make_crops(5, 37)
    sd = syn.make_vs_state_dict(0, "spread")
    p_all, f_all = _vs_probs(sd, "bf16", crops)
    p_one, f_one = _vs_probs(sd, "bf16", crops[7:8])
    p_rev, _ = _vs_probs(sd, "bf16", crops[::-1].copy())
    assert np.array_equal(p_all[7:8], p_one) and np.array_equal(f_all[7:8], f_one)
    assert np.array_equal(p_all[::-1], p_rev)


def test_vs_config2_full_batch_against_oracle(cuda_lib):
    """BASELINE config 2 at full size: 256 crops through K1 + the bf16 VS forward (one batch, every fused path: stem + pool,
    K-concatenated shortcuts, FLAT residual epilogues) against the fp32 oracle of architectures/video.py on the same crops;
    north-star tolerance 2e-3 on the per-class probabilities with PyTorch-default random init, and identical arg-max
    wherever the oracle's top-2 margin exceeds twice that."""
    crops = syn.make_crops(77, 256)
    sd = syn.make_vs_state_dict(0, "default")
    probs, feat = _vs_probs(sd, "bf16", crops)
    x = torch.from_numpy(np.stack([ov.pth_processing(c)[0] for c in crops]))
    logits, ofeat = ov.resnet50_forward(sd, x)
    ref = torch.softmax(logits, dim=1).numpy()
    assert probs.shape == ref.shape == (256, 7)
    err = np.abs(probs - ref).max()
    assert err < 2e-3, err
    top2 = np.sort(ref, axis=1)[:, -2:]
    sure = (top2[:, 1] - top2[:, 0]) > 4e-3
    assert np.array_equal(probs.argmax(1)[sure], ref.argmax(1)[sure])
    assert np.abs(feat - np.maximum(ofeat.numpy(), 0)).max() < 0.15


@pytest.mark.parametrize("prec", ["fp32", "bf16", "fp16"])
def test_vd_matches_reference_golden(cuda_lib, golden, prec):
    from avcer_b200 import nets

    gen = torch.Generator().manual_seed(5)
    xw = torch.relu(torch.randn(12, 10, 512, generator=gen))
    net = nets.VDNet(syn.make_vd_state_dict(1), prec, DEV)
    feats = xw.reshape(120, 512).to(DEV)                      # fp32 relu(fc1) features, as VSNet hands them over
    wins = torch.arange(120, dtype=torch.int32).view(12, 10).t().contiguous().to(DEV)
    out = net.forward(feats, wins).cpu()
    ref = torch.from_numpy(golden["video"]["vd_logits"])
    perr = (torch.softmax(out, 1) - torch.softmax(ref, 1)).abs().max().item()
    # bf16 mode = bf16x3 contractions (split operands): the north-star 2e-3 holds with two orders of magnitude to spare
    assert perr < (1e-5 if prec == "fp32" else 1e-4), perr
    assert (out - ref).abs().max().item() < (2e-5 if prec == "fp32" else 1e-3)


@pytest.mark.parametrize("prec", ["fp32", "bf16", "fp16"])
@pytest.mark.parametrize("ncls", [8, 7])
def test_audio_matches_reference_golden(cuda_lib, golden, prec, ncls):
    from avcer_b200 import nets, ops, pipeline

    g = golden["audio"]
    L, fps, pad, step = (int(v) for v in g[f"a{ncls}_a_meta"])
    wav = syn.make_wav(31, L)
    net = nets.ANet(syn.make_audio_state_dict(2, ncls, "spread", 12), prec, DEV)
    ap = pipeline.plan_audio(L, fps, step / 1000)
    x = ops.audio_normalize_windows(torch.from_numpy(wav).to(DEV), torch.from_numpy(ap.starts).to(DEV), 64000, "mean")
    out = net.forward(x).cpu().numpy()
    ref = g[f"a{ncls}_a_window_logits"]
    assert out.shape == ref.shape
    p = torch.softmax(torch.from_numpy(out[:, :7]), 1).numpy()
    pr = torch.softmax(torch.from_numpy(ref[:, :7]), 1).numpy()
    # "spread" init (logit range ~5): the bf16 error budget of the module docstring (emulated 6.8e-3 on exactly these
    # windows, measured 8e-3); the north-star 2e-3 is asserted on the default and "mid" inits below
    # fp16 storage: the same wide init within the north-star 2e-3 (emulated 1.2e-3)
    assert np.abs(p - pr).max() < {"fp32": 1e-5, "bf16": 1.2e-2, "fp16": 2e-3}[prec], np.abs(p - pr).max()
    if prec == "bf16":
        top2 = np.sort(pr, axis=1)[:, -2:]
        sure = (top2[:, 1] - top2[:, 0]) > 2.4e-2
        assert np.array_equal(p.argmax(1)[sure], pr.argmax(1)[sure])
    assert np.abs(out - ref).max() < {"fp32": 1e-4, "bf16": 0.08, "fp16": 0.02}[prec]


@pytest.mark.parametrize("prec", ["fp32", "bf16", "fp16"])
@pytest.mark.parametrize("init", ["default", "mid"])
def test_audio_north_star_tolerance(cuda_lib, golden, prec, init):
    """2e-3 (bf16) / 1e-5 (fp32) on the per-class probabilities against the unmodified reference: PyTorch-default init
    and the input-dependent "mid" init (module docstring)."""
    from avcer_b200 import nets, ops, pipeline

    wav = syn.make_wav(31, 52800 + 123)
    net = nets.ANet(syn.make_audio_state_dict(2, 8, init, 12), prec, DEV)
    ap = pipeline.plan_audio(len(wav), 25, 0.5)
    x = ops.audio_normalize_windows(torch.from_numpy(wav).to(DEV), torch.from_numpy(ap.starts).to(DEV), 64000, "mean")
    out = net.forward(x).cpu().numpy()
    ref = golden["audio"][f"a8_{init}_window_logits"]
    p = torch.softmax(torch.from_numpy(out[:, :7]), 1).numpy()
    pr = torch.softmax(torch.from_numpy(ref[:, :7]), 1).numpy()
    tol = 1e-5 if prec == "fp32" else 2e-3
    assert np.abs(p - pr).max() < tol, np.abs(p - pr).max()
    if init == "mid":
        assert (pr.max(0) - pr.min(0)).max() > 10 * tol          # outputs depend on the input window
        top2 = np.sort(pr, axis=1)[:, -2:]
        sure = (top2[:, 1] - top2[:, 0]) > 2 * tol
        assert sure.any() and np.array_equal(p.argmax(1)[sure], pr.argmax(1)[sure])


@pytest.mark.parametrize("prec", ["fp32", "bf16", "fp16"])
def test_audio_v1_gru_variant_matches_reference_golden(cuda_lib, golden, prec):
    """ExprModelV1 (architectures/audio_8_cl.py:18-72: wav2vec2 -> 2-layer GRU(1024 -> 256) -> 256-wide head), the variant
    the reference's alternate model lists use, against the unmodified reference class (tests/golden/audio.npz)."""
    from avcer_b200 import nets, ops, pipeline

    wav = syn.make_wav(31, 52800 + 123)
    net = nets.ANet(syn.make_audio_state_dict(2, 8, "mid", 12, variant="v1"), prec, DEV)
    assert net.w["variant"] == "v1" and net.w["f_size"] == 256
    ap = pipeline.plan_audio(len(wav), 25, 0.5)
    x = ops.audio_normalize_windows(torch.from_numpy(wav).to(DEV), torch.from_numpy(ap.starts[:3]).to(DEV), 64000, "mean")
    out = net.forward(x).cpu().numpy()
    ref = golden["audio"]["a8_v1_window_logits"]
    assert out.shape == ref.shape == (3, 8)
    p = torch.softmax(torch.from_numpy(out[:, :7]), 1).numpy()
    pr = torch.softmax(torch.from_numpy(ref[:, :7]), 1).numpy()
    assert np.abs(p - pr).max() < (1e-5 if prec == "fp32" else 2e-3), np.abs(p - pr).max()
    assert np.abs(out - ref).max() < (1e-4 if prec == "fp32" else 0.02)


def test_audio_padding_modes_and_nan_window(cuda_lib, golden):
    from avcer_b200 import ops, pipeline

    g = golden["audio"]
    L, fps, pad, step = (int(v) for v in g["a8_b_meta"])                    # L multiple of step_a -> trailing empty window
    wav = syn.make_wav(31, L)
    ap = pipeline.plan_audio(L, fps, step / 1000)
    wd = torch.from_numpy(wav).to(DEV)
    st = torch.from_numpy(ap.starts).to(DEV)
    for mode in ("mean", "constant"):
        x = ops.audio_normalize_windows(wd, st, 64000, mode).cpu().numpy()
        for i, (s, e) in enumerate(zip(ap.starts, ap.ends)):
            ref = oa.zero_mean_unit_var(oa.pad_window(wav[s:e], 64000, mode))
            if np.isnan(ref).any():
                assert np.isnan(x[i]).all()
            else:
                assert np.abs(x[i] - ref).max() < 5e-6 * max(1.0, np.abs(ref).max())
    L2 = 40000 - 160
    wav2 = syn.make_wav(31, L2)
    ap2 = pipeline.plan_audio(L2, 25, 1)
    x = ops.audio_normalize_windows(torch.from_numpy(wav2).to(DEV), torch.from_numpy(ap2.starts).to(DEV), 64000, "repeat").cpu().numpy()
    for i, (s, e) in enumerate(zip(ap2.starts, ap2.ends)):
        ref = oa.zero_mean_unit_var(oa.pad_window(wav2[s:e], 64000, "repeat"))
        assert np.abs(x[i] - ref).max() < 2e-5


def test_contract_against_oracle_conv(cuda_lib):
    """avcer_contract unit cases vs torch CPU fp32 convolutions: odd spatial sizes, tail tiles,
    stride-2 views, residual + ReLU, both backends."""
    import torch.nn.functional as F
    from avcer_b200 import ops

    gen = torch.Generator().manual_seed(0)
    for dtype, tol in ((torch.float32, 2e-4), (torch.bfloat16, 0.05)):
        for (n, h, w, cin, cout, k, stride, res) in ((3, 55, 55, 64, 64, 3, 1, False), (2, 28, 28, 128, 512, 1, 1, True),
                                                     (5, 55, 55, 256, 128, 1, 2, False), (7, 7, 7, 512, 512, 3, 1, False),
                                                     (1, 14, 14, 256, 256, 3, 1, True)):
            x = torch.randn(n, h, w, cin, generator=gen).to(dtype)
            w4 = (torch.randn(cout, cin, k, k, generator=gen) / (cin * k * k) ** 0.5).to(dtype)
            b = torch.randn(cout, generator=gen)
            pad = (k - 1) // 2
            ref = F.conv2d(x.float().permute(0, 3, 1, 2), w4.float(), b, stride=stride, padding=pad).permute(0, 2, 3, 1)
            r = torch.randn(ref.shape, generator=gen).to(dtype) if res else None
            if res:
                ref = ref + r.float()
            ref = F.relu(ref)
            wt = w4.permute(0, 2, 3, 1).reshape(cout, -1).contiguous()
            got = ops.conv2d_nhwc(x.to(DEV), wt.to(DEV), b.to(DEV), kh=k, kw=k, stride=stride, pad_h=pad, pad_w=pad,
                                  residual=None if r is None else r.to(DEV), act=ops.ACT_RELU).float().cpu()
            assert (got - ref).abs().max().item() < tol, (dtype, n, h, w, cin, cout, k, stride)


def test_stem_strip_mode_matches_oracle_conv(cuda_lib):
    """The 7x7/2 "TF-same" stem as an implicit GEMM whose A operand is one contiguous padded image row per
    filter row (overlapping windows expressed in the UMMA descriptor, no im2col): against F.conv2d on CPU."""
    import torch.nn.functional as F
    from avcer_b200 import nets, ops

    sd = syn.make_vs_state_dict(0, "spread")
    crops = syn.make_crops(3, 3)
    ref_in = torch.from_numpy(np.concatenate([ov.pth_processing(c) for c in crops]))
    s = sd["batch_norm1.weight"] / torch.sqrt(sd["batch_norm1.running_var"] + 1e-3)
    ref = F.conv2d(F.pad(ref_in, [2, 3, 2, 3]), sd["conv_layer_s2_same.weight"], stride=2)
    ref = F.relu(ref * s.view(1, -1, 1, 1) + (sd["batch_norm1.bias"] - sd["batch_norm1.running_mean"] * s).view(1, -1, 1, 1))
    for prec, tol in (("fp32", 2e-4), ("bf16", 0.06)):
        net = nets.VSNet(sd, prec, DEV)
        x = net.alloc_input(3)
        ops.preprocess(torch.from_numpy(crops).to(DEV), 3, x, net.input_layout)
        y = net.stem(x).float().cpu().permute(0, 3, 1, 2)
        assert y.shape == ref.shape == (3, 64, 112, 112)
        assert (y - ref).abs().max().item() < tol, prec


def test_gemm_tile_flavours_agree(cuda_lib):
    """128x256 / 128x128 / 128x64 tiles, one or two CTAs per SM, fp32-out direct epilogue: same GEMM, same answer."""
    from avcer_b200 import ops

    gen = torch.Generator().manual_seed(1)
    for (m, k, n) in ((40000, 256, 512), (999, 1024, 256), (5000, 64, 64), (70000, 128, 128), (300, 512, 2048)):
        x = torch.randn(m, k, generator=gen).bfloat16()
        w = (torch.randn(n, k, generator=gen) / k ** 0.5).bfloat16()
        b = torch.randn(n, generator=gen)
        r = torch.randn(m, n, generator=gen).bfloat16()
        ref = x.float() @ w.float().t() + b
        got = ops.linear(x.to(DEV), w.to(DEV), b.to(DEV)).float().cpu()
        assert (got - ref).abs().max().item() < 0.05, (m, k, n)
        got = ops.linear(x.to(DEV), w.to(DEV), b.to(DEV), residual=r.to(DEV), act=ops.ACT_GELU).float().cpu()
        assert (got - F_gelu(ref + r.float())).abs().max().item() < 0.06, (m, k, n)
        got = ops.linear(x.to(DEV), w.to(DEV), b.to(DEV), out_dtype=torch.float32).cpu()
        assert (got - ref).abs().max().item() < 2e-3, (m, k, n)


def F_gelu(t):
    return torch.nn.functional.gelu(t)


def test_forwards_are_bit_stable_run_to_run(cuda_lib):
    """Every kernel of the path is deterministic (no atomics, fixed reduction orders): repeated forwards of the audio
    network (64 windows: 12 encoder layers, each with two residual GEMMs on the 4-deep residual ring) and of the VS
    ResNet-50 must be bit-identical at every tap.  Guards against races between pipeline stages of one kernel and
    between consecutive kernels (the failure mode seen with programmatic dependent launch, csrc/common.h)."""
    from avcer_b200 import nets, ops

    g = torch.Generator(device=DEV).manual_seed(11)
    anet = nets.ANet(syn.make_audio_state_dict(2, 8, "spread", 12), "bf16", DEV)
    x = torch.randn((64, 64000), device=DEV, generator=g)
    ref = None
    for _ in range(6):
        taps = {}
        taps["logits"] = anet.forward(x, taps)
        torch.cuda.synchronize()
        cur = {k: v.clone() for k, v in taps.items()}
        if ref is None:
            ref = cur
            continue
        for k in ref:
            assert torch.equal(ref[k], cur[k]), f"audio tap {k} differs between two runs on the same input"
    vnet = nets.VSNet(syn.make_vs_state_dict(0, "default"), "bf16", DEV)
    crops = torch.randint(0, 256, (96, 224, 224, 3), dtype=torch.uint8, device=DEV, generator=g)
    xin = vnet.alloc_input(96)
    ops.preprocess(crops, 96, xin, vnet.input_layout)
    outs = [tuple(t.clone() for t in vnet.forward(xin)) for _ in range(4)]
    for o in outs[1:]:
        assert torch.equal(o[0], outs[0][0]) and torch.equal(o[1], outs[0][1])
