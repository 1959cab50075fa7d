#!/usr/bin/env python
"""CPU emulation of the bf16 rounding points of the VS / VD / A forwards (no GPU needed).

Every tensor the CUDA path stores in bf16 is rounded with .bfloat16().float(); contractions run in fp32 on the rounded
operands (products of bf16 values are exact in fp32, accumulation is fp32 like the TMEM accumulators).  Variants:

  all_bf16      : round-1 scheme -- weights, every activation and the residual stream in bf16
  stream_f32    : residual stream (ResNet block outputs / encoder hidden state) kept in fp32; GEMM operands bf16
  weights_only  : only the weights rounded (what bf16 weights alone cost)
  acts_only     : only activations rounded (fp32 weights)

Prints max |p - p_fp32| over classes for each variant: the error budget behind the tolerances in tests/test_gpu_nets.py
and DESIGN.md section 2.  Usage: python scripts/sim_bf16_budget.py [vs|vd|a] [init] [scale]
"""
import math
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from avcer_b200 import synthetic as syn          # noqa: E402
from avcer_b200 import weights as wpack          # noqa: E402
from oracle import audio as oa                   # noqa: E402
from oracle import video as ov                   # noqa: E402


def rb(t):
    return t.bfloat16().float()


def ident(t):
    return t


# ---------------------------------------------------------------------------------------------- VS
def vs_forward(sd, x, qa, qs, qw):
    """qa: activation (GEMM operand) rounding, qs: residual-stream rounding, qw: weight rounding."""
    eps = 1e-3

    def fold(wname, bn):
        w, b = wpack._fold_bn(sd[wname], sd, bn, eps)
        return qw(w), b

    w, b = fold("conv_layer_s2_same.weight", "batch_norm1")
    y = F.conv2d(F.pad(qa(x), [2, 3, 2, 3]), w, b, stride=2)
    y = qa(F.relu(y))
    y = F.max_pool2d(y, 3, 2)
    for li, blocks in enumerate(ov.VS_BLOCKS, start=1):
        for bi in range(blocks):
            p = f"layer{li}.{bi}"
            stride = 2 if (li > 1 and bi == 0) else 1
            w1, b1 = fold(p + ".conv1.weight", p + ".batch_norm1")
            w2, b2 = fold(p + ".conv2.weight", p + ".batch_norm2")
            w3, b3 = fold(p + ".conv3.weight", p + ".batch_norm3")
            xin = qa(y)
            t = qa(F.relu(F.conv2d(xin, w1, b1, stride=stride)))
            t = qa(F.relu(F.conv2d(t, w2, b2, padding=1)))
            o = F.conv2d(t, w3, b3)
            if bi == 0:
                wd, bd = fold(p + ".i_downsample.0.weight", p + ".i_downsample.1")
                o = o + F.conv2d(xin, wd, bd, stride=stride)          # K-concatenated: shortcut from the bf16 operand
            else:
                o = o + y                                             # residual read from the stream
            y = qs(F.relu(o))
    y = qa(qa(y).mean(dim=(2, 3)))
    feat = qa(F.relu(F.linear(y, qw(sd["fc1.weight"]), sd["fc1.bias"])))
    return F.linear(feat, sd["fc2.weight"], sd["fc2.bias"])


def run_vs(init, scale):
    sd = syn.make_vs_state_dict(0, init)
    if scale != 1.0:
        sd["fc2.weight"] = sd["fc2.weight"] * scale
    crops = syn.make_crops(11, 6)
    x = torch.from_numpy(np.concatenate([ov.pth_processing(c) for c in crops]))
    with torch.no_grad():
        ref = torch.softmax(vs_forward(sd, x, ident, ident, ident), 1)
        lg = vs_forward(sd, x, ident, ident, ident)
        print(f"VS init={init} scale={scale}: logit range {float(lg.max() - lg.min()):.2f}, p max {float(ref.max()):.3f}, p std over crops {float(ref.std(0).mean()):.4f}")
        for name, (qa, qs, qw) in {"all_bf16": (rb, rb, rb), "stream_f32": (rb, ident, rb), "weights_only": (ident, ident, rb),
                                   "acts_only": (rb, rb, ident), "acts_only_stream_f32": (rb, ident, ident)}.items():
            p = torch.softmax(vs_forward(sd, x, qa, qs, qw), 1)
            print(f"  {name:22s} max|dp| = {float((p - ref).abs().max()):.2e}")


# ---------------------------------------------------------------------------------------------- VD
def vd_forward(sd, x, qa, qw, qh=None):
    qh = qh or qa

    def layer(xs, w_ih, w_hh, b):
        bsz, steps, _ = xs.shape
        hid = w_hh.shape[1]
        h = xs.new_zeros(bsz, hid)
        c = xs.new_zeros(bsz, hid)
        outs = []
        for t in range(steps):
            g = qa(xs[:, t]) @ qw(w_ih).t() + qh(h) @ qw(w_hh).t() + b
            i, f, gg, o = g.split(hid, dim=1)
            c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
            h = torch.sigmoid(o) * torch.tanh(c)
            outs.append(h)
        return torch.stack(outs, 1)

    y = layer(x, sd["lstm1.weight_ih_l0"], sd["lstm1.weight_hh_l0"], sd["lstm1.bias_ih_l0"] + sd["lstm1.bias_hh_l0"])
    y = layer(y, sd["lstm2.weight_ih_l0"], sd["lstm2.weight_hh_l0"], sd["lstm2.bias_ih_l0"] + sd["lstm2.bias_hh_l0"])
    return F.linear(qh(y[:, -1]), sd["fc.weight"], sd["fc.bias"])


def split2(t):
    """bf16x2: hi + lo, both bf16 (the [a_hi | a_lo | a_hi] x [w_hi | w_hi | w_lo] K-concatenation keeps 16 mantissa bits)."""
    hi = rb(t)
    return hi + rb(t - hi)


def run_vd():
    sd = syn.make_vd_state_dict(1)
    gen = torch.Generator().manual_seed(5)
    xw = torch.relu(torch.randn(12, 10, 512, generator=gen))
    with torch.no_grad():
        ref = torch.softmax(vd_forward(sd, xw, ident, ident), 1)
        for name, (qa, qw, qh) in {"all_bf16": (rb, rb, rb), "bf16 x, split h+w": (rb, split2, split2), "split all": (split2, split2, split2),
                                   "bf16 x+w, f32 h": (rb, rb, ident)}.items():
            p = torch.softmax(vd_forward(sd, xw, qa, qw, qh), 1)
            print(f"  VD {name:22s} max|dp| = {float((p - ref).abs().max()):.2e}")


# ---------------------------------------------------------------------------------------------- A
class _StageQ:
    """qa given as {stage: fn}: activations are rounded only in the listed stages (per-stage error budget)."""

    def __init__(self, table):
        self.table, self.stage = table, "fe"

    def __call__(self, t):
        return self.table.get(self.stage, ident)(t)


def a_forward(sd, x, qa, qs, qw, n_layers=12):
    p = "wav2vec2."
    if isinstance(qa, dict):
        qa = _StageQ(qa)

    def stage(name):
        if isinstance(qa, _StageQ):
            qa.stage = name

    def ln(t, name, eps=1e-5):
        return F.layer_norm(t, (t.shape[-1],), sd[name + ".weight"], sd[name + ".bias"], eps)

    h = x[:, None, :]
    for i, s in enumerate(oa.CONV_STRIDE):
        q = f"{p}feature_extractor.conv_layers.{i}"
        w = sd[q + ".conv.weight"] if i == 0 else qw(sd[q + ".conv.weight"])
        h = F.conv1d(h, w, sd[q + ".conv.bias"], stride=s)
        if i > 0:
            h = qa(h)                                   # conv output stored, then LN + GELU in place
        h = qa(F.gelu(ln(h.transpose(1, 2), q + ".layer_norm")).transpose(1, 2))
    h = h.transpose(1, 2)
    stage("proj")
    h = qa(ln(h, p + "feature_projection.layer_norm"))
    h = qs(F.linear(h, qw(sd[p + "feature_projection.projection.weight"]), sd[p + "feature_projection.projection.bias"]))
    pc = F.conv1d(qa(h).transpose(1, 2), qw(oa.pos_conv_weight(sd)), sd[p + "encoder.pos_conv_embed.conv.bias"], padding=64, groups=16)
    h = qs(h + F.gelu(pc[:, :, :-1]).transpose(1, 2))

    def mha(a, wq, bq, wk, bk, wv, bv, wo, bo, heads):
        b, t, d = a.shape
        dh = d // heads
        qq = qa(F.linear(a, qw(wq), bq)).view(b, t, heads, dh).transpose(1, 2)
        kk = qa(F.linear(a, qw(wk), bk)).view(b, t, heads, dh).transpose(1, 2)
        vv = qa(F.linear(a, qw(wv), bv)).view(b, t, heads, dh).transpose(1, 2)
        pr = torch.softmax((qq @ kk.transpose(-1, -2)) / math.sqrt(dh), dim=-1)
        o = qa((qa(pr) @ vv).transpose(1, 2).reshape(b, t, d))
        return F.linear(o, qw(wo), bo)

    for i in range(n_layers):
        q = f"{p}encoder.layers.{i}"
        stage("enc")
        a = qa(ln(h, q + ".layer_norm"))
        a = mha(a, sd[q + ".attention.q_proj.weight"], sd[q + ".attention.q_proj.bias"], sd[q + ".attention.k_proj.weight"],
                sd[q + ".attention.k_proj.bias"], sd[q + ".attention.v_proj.weight"], sd[q + ".attention.v_proj.bias"],
                sd[q + ".attention.out_proj.weight"], sd[q + ".attention.out_proj.bias"], 16)
        h = qs(h + a)
        f = qa(ln(h, q + ".final_layer_norm"))
        f = qa(F.gelu(F.linear(f, qw(sd[q + ".feed_forward.intermediate_dense.weight"]), sd[q + ".feed_forward.intermediate_dense.bias"])))
        h = qs(h + F.linear(f, qw(sd[q + ".feed_forward.output_dense.weight"]), sd[q + ".feed_forward.output_dense.bias"]))
    stage("tl")
    h = qa(ln(h, p + "encoder.layer_norm"))
    for name, heads in (("tl1", 32), ("tl2", 16)):
        t = h.shape[1]
        xp = qa(h + sd[f"{name}.positional_encoding.pe"][:, :t])
        a = mha(xp, sd[f"{name}.self_attention.query_w.weight"], None, sd[f"{name}.self_attention.keys_w.weight"], None,
                sd[f"{name}.self_attention.values_w.weight"], None, sd[f"{name}.self_attention.ff_layer_after_concat.weight"], None, heads)
        y = qa(ln(qs(a + xp), f"{name}.add_norm_after_attention.layer_norm"))
        f = qa(F.relu(F.linear(y, qw(sd[f"{name}.feed_forward.layer_1.weight"]), sd[f"{name}.feed_forward.layer_1.bias"])))
        f = F.linear(f, qw(sd[f"{name}.feed_forward.layer_2.weight"]), sd[f"{name}.feed_forward.layer_2.bias"])
        h = qa(ln(qs(f + y), f"{name}.add_norm_after_ff.layer_norm"))
    h = h.permute(0, 2, 1)
    stage("head")
    w, b = wpack._fold_bn(sd["time_downsample.0.weight"], sd, "time_downsample.1", 1e-5, sd["time_downsample.0.bias"])
    h = qa(F.conv1d(h, qw(w), b, stride=3, dilation=2))
    h = qa(F.relu(F.max_pool1d(h, 5)))
    w, b = wpack._fold_bn(sd["time_downsample.4.weight"], sd, "time_downsample.5", 1e-5, sd["time_downsample.4.bias"])
    h = qa(F.conv1d(h, qw(w), b))
    h = qa(F.relu(h.mean(dim=2)))
    return F.linear(h, sd["feature_downsample.weight"], sd["feature_downsample.bias"])


def run_a(init, scale, stages=False):
    """The windows of tests/test_gpu_nets.py::test_audio_matches_reference_golden (52 923 samples, 7 windows incl. the
    mostly padded tail ones)."""
    from avcer_b200 import pipeline

    sd = syn.make_audio_state_dict(2, 8, init, 12)
    if scale != 1.0:
        sd["feature_downsample.weight"] = sd["feature_downsample.weight"] * scale
    L = 52800 + 123
    wav = syn.make_wav(31, L)
    ap = pipeline.plan_audio(L, 25, 0.5)
    xs = np.stack([oa.zero_mean_unit_var(oa.pad_window(wav[s:e], 64000, "mean")) for s, e in zip(ap.starts, ap.ends)])
    x = torch.from_numpy(xs)
    with torch.no_grad():
        lg = a_forward(sd, x, ident, ident, ident)
        ref = torch.softmax(lg[:, :7], 1)
        print(f"A init={init} scale={scale}: logit range {float(lg[:, :7].max() - lg[:, :7].min()):.2f}, p max {float(ref.max()):.3f}")
        variants = {"all_bf16": (rb, rb, rb), "stream_f32": (rb, ident, rb), "weights_only": (ident, ident, rb)}
        if stages:
            variants["residual stream only"] = (ident, rb, ident)
            for st, what in (("fe", "feature-extractor activations only"), ("proj", "projection + positional conv operands only"),
                             ("enc", "12 encoder layers' GEMM operands only"), ("tl", "tl1 / tl2 operands only"),
                             ("head", "time_downsample head activations only")):
                variants[what] = ({st: rb}, ident, ident)
        for name, (qa, qs, qw) in variants.items():
            l2 = a_forward(sd, x, qa, qs, qw)
            p = torch.softmax(l2[:, :7], 1)
            print(f"  {name:44s} max|dp| = {float((p - ref).abs().max()):.2e}   max|dlogit| = {float((l2 - lg).abs().max()):.2e}")


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count() or 1)
    which = sys.argv[1] if len(sys.argv) > 1 else "report"
    init = sys.argv[2] if len(sys.argv) > 2 else "spread"
    scale = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
    if which == "vs":
        run_vs(init, scale)
    elif which == "vd":
        run_vd()
    elif which == "a":
        run_a(init, scale)
    else:
        print("# bf16 error budget by CPU emulation of every rounding point (scripts/sim_bf16_budget.py report)")
        print("# max|dp| = max over inputs and classes of |softmax(bf16 path) - softmax(fp32 path)|; north-star bar 2e-3")
        run_vs("spread", 1.0)
        run_vs("mid", 1.0)
        run_vs("default", 1.0)
        run_vd()
        run_a("spread", 1.0, stages=True)
        run_a("mid", 1.0)
