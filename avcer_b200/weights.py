"""Weight packer: reference state_dicts -> kernel-native layouts.

Input contract = the state_dict keys/shapes of the reference classes:
ResNet50(7) and LSTMPyTorch (src/architectures/video.py:93-185), ExprModelV3/V2
(src/architectures/audio_8_cl.py:131-161, audio_7_cl.py:75-128).

Packing rules
  * BatchNorm (eval, running stats) is folded into the preceding bias-free convolution:
      s = gamma / sqrt(var + eps);  w' = w * s;  b' = beta - mean * s      (eps 1e-3 for VS, 1e-5 for A)
  * convolution weights become [Cout, taps * Cin] (tap-major, channel-minor), the K-contiguous
    operand layout of avcer_contract;
  * the 7x7/2 stem becomes [64, 7 rows, 8 pixels x 4 channels] to match the zero-bordered NHWC4
    input (pixel 7 and channel 3 carry zero weights);
  * weight_norm of the wav2vec2 positional conv is folded: w = g * v / ||v||_(0,1);
  * q/k/v projections are concatenated into one [3*1024, 1024] matrix;
  * storage dtype is bf16 (default) or fp32 ("fp32 mode"); biases and the small heads stay fp32.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

VS_BLOCKS = (3, 4, 6, 3)
VS_PLANES = (64, 128, 256, 512)


def _fold_bn(w: torch.Tensor, sd, p: str, eps: float, conv_bias: Optional[torch.Tensor] = None):
    s = sd[p + ".weight"].double() / torch.sqrt(sd[p + ".running_var"].double() + eps)
    shape = [-1] + [1] * (w.dim() - 1)
    w2 = w.double() * s.view(shape)
    b = sd[p + ".bias"].double() - sd[p + ".running_mean"].double() * s
    if conv_bias is not None:
        b = b + conv_bias.double() * s
    return w2.float(), b.float()


def _tap_major(w: torch.Tensor) -> torch.Tensor:
    """[Cout, Cin, *k] -> [Cout, prod(k) * Cin]."""
    cout, cin = w.shape[:2]
    if w.dim() == 4:
        return w.permute(0, 2, 3, 1).reshape(cout, -1).contiguous()
    if w.dim() == 3:
        return w.permute(0, 2, 1).reshape(cout, -1).contiguous()
    return w.contiguous()


class PackedConv:
    __slots__ = ("wt", "bias", "cin", "cout", "k", "stride")

    def __init__(self, wt, bias, cin, cout, k, stride):
        self.wt, self.bias, self.cin, self.cout, self.k, self.stride = wt, bias, cin, cout, k, stride


def pack_vs(sd: Dict[str, torch.Tensor], device, dtype: torch.dtype) -> dict:
    """ResNet-50 VS: returns {'stem': PackedConv, 'blocks': [dict], 'fc1': PackedConv, 'fc2_w','fc2_b'}."""
    eps = 1e-3

    def dev(t, dt=None):
        return t.to(device=device, dtype=dt or dtype).contiguous()

    out = {}
    w, b = _fold_bn(sd["conv_layer_s2_same.weight"], sd, "batch_norm1", eps)
    stem = torch.zeros(64, 7, 8, 4)
    stem[:, :, :7, :3] = w.permute(0, 2, 3, 1)          # [co, ky, kx, c]
    out["stem"] = PackedConv(dev(stem.reshape(64, 7 * 32)), dev(b, torch.float32), 32, 64, 7, 2)
    # the same weights per filter row in UMMA core-matrix order [ky][k/8][cout/8][8 cout][8 k] (strip-mode B operand)
    out["stem_packed"] = dev(stem.reshape(8, 8, 7, 4, 8).permute(2, 3, 0, 1, 4).contiguous().reshape(-1))
    blocks: List[dict] = []
    cin = 64
    for li, (planes, nblocks) in enumerate(zip(VS_PLANES, VS_BLOCKS), start=1):
        for bi in range(nblocks):
            p = f"layer{li}.{bi}"
            stride = 2 if (li > 1 and bi == 0) else 1
            blk = {}
            for j, (k, ci, co, st) in enumerate([(1, cin, planes, stride), (3, planes, planes, 1), (1, planes, planes * 4, 1)], start=1):
                w, b = _fold_bn(sd[f"{p}.conv{j}.weight"], sd, f"{p}.batch_norm{j}", eps)
                blk[f"conv{j}"] = PackedConv(dev(_tap_major(w)), dev(b, torch.float32), ci, co, k, st)
            if bi == 0:
                wd, bd = _fold_bn(sd[f"{p}.i_downsample.0.weight"], sd, f"{p}.i_downsample.1", eps)
                blk["ds"] = PackedConv(dev(_tap_major(wd)), dev(bd, torch.float32), cin, planes * 4, 1, stride)
                if True:
                    # conv3 + projection shortcut as ONE contraction over K = [(sampled) block input | conv2 output]
                    # (video.py:46-58: relu(bn3(conv3(t)) + bn_ds(conv_ds(x)))): the shortcut never reaches memory
                    blk["conv3_ds"] = PackedConv(dev(torch.cat([_tap_major(wd), _tap_major(w)], dim=1)), dev(bd + b, torch.float32),
                                                 cin + planes, planes * 4, 1, 1)
            blocks.append(blk)
            cin = planes * 4
    out["blocks"] = blocks
    out["fc1"] = PackedConv(dev(sd["fc1.weight"]), dev(sd["fc1.bias"], torch.float32), 2048, 512, 1, 1)
    out["fc2_w"] = dev(sd["fc2.weight"], torch.float32)
    out["fc2_b"] = dev(sd["fc2.bias"], torch.float32)
    return out


def split_bf16x3_weight(w: torch.Tensor, dtype: torch.dtype = torch.bfloat16) -> torch.Tensor:
    """[N, K] fp32 -> [N, 3K] 16-bit = [w_hi | w_hi | w_lo] (w_lo = round(w - w_hi)): the weight side of a bf16x3 (fp16x3)
    contraction against activations laid out [a_hi | a_lo | a_hi] (avcer_split_bf16x3 / avcer_lstm_cell split)."""
    hi = w.float().to(dtype)
    lo = (w.float() - hi.float()).to(dtype)
    return torch.cat([hi, hi, lo], dim=1)


RF_BLOCKS = (3, 4, 6, 3)
RF_PLANES = (64, 128, 256, 512)


def pack_retinaface(sd: Dict[str, torch.Tensor], device, dtype: torch.dtype) -> dict:
    """RetinaFace-ResNet50 (data/face_detection/ibug/face_detection/retina_face/retina_face.py:48-115; state_dict keys
    body.* = torchvision resnet50, fpn.*, ssh{1,2,3}.*, ClassHead / BboxHead / LandmarkHead.{0,1,2}.conv1x1).  BatchNorm
    (eps 1e-5) folded; the stem stays fp32 [147, 64] (row = (ky*7 + kx)*3 + c) for avcer_det_stem; the three 1x1 heads of a
    level become one [64, 256] matrix (rows: 4 class, 8 box, 20 landmark logits, 32 zero rows)."""
    eps = 1e-5

    def dev(t, dt=None):
        return t.to(device=device, dtype=dt or dtype).contiguous()

    def conv_bn(wkey, bnkey, k, stride=1):
        w, b = _fold_bn(sd[wkey], sd, bnkey, eps)
        return PackedConv(dev(_tap_major(w)), dev(b, torch.float32), w.shape[1], w.shape[0], k, stride)

    out = {}
    w, b = _fold_bn(sd["body.conv1.weight"], sd, "body.bn1", eps)
    out["stem_w"] = dev(w.permute(2, 3, 1, 0).reshape(147, 64), torch.float32)
    out["stem_b"] = dev(b, torch.float32)
    if dtype in (torch.bfloat16, torch.float16):
        # tensor-core stem: the VS stem's strip-mode layouts ([64, 7 rows, 8 pixels x 4 channels]; pixel 7, channel 3 zero)
        stem = torch.zeros(64, 7, 8, 4)
        stem[:, :, :7, :3] = w.permute(0, 2, 3, 1)
        out["stem_wt"] = dev(stem.reshape(64, 7 * 32))
        out["stem_packed"] = dev(stem.reshape(8, 8, 7, 4, 8).permute(2, 3, 0, 1, 4).contiguous().reshape(-1))
    blocks: List[dict] = []
    for li, (planes, nblocks) in enumerate(zip(RF_PLANES, RF_BLOCKS), start=1):
        for bi in range(nblocks):
            p = f"body.layer{li}.{bi}"
            stride = 2 if (li > 1 and bi == 0) else 1
            blk = {"conv1": conv_bn(p + ".conv1.weight", p + ".bn1", 1), "conv2": conv_bn(p + ".conv2.weight", p + ".bn2", 3, stride),
                   "conv3": conv_bn(p + ".conv3.weight", p + ".bn3", 1), "last_of_layer": bi == nblocks - 1, "layer": li}
            if bi == 0:
                blk["ds"] = conv_bn(p + ".downsample.0.weight", p + ".downsample.1", 1, stride)
            blocks.append(blk)
    out["blocks"] = blocks
    out["fpn_out"] = [conv_bn(f"fpn.output{i}.0.weight", f"fpn.output{i}.1", 1) for i in (1, 2, 3)]
    out["fpn_merge"] = [conv_bn(f"fpn.merge{i}.0.weight", f"fpn.merge{i}.1", 3) for i in (1, 2)]
    out["ssh"] = [{name: conv_bn(f"ssh{i}.{name}.0.weight", f"ssh{i}.{name}.1", 3)
                   for name in ("conv3X3", "conv5X5_1", "conv5X5_2", "conv7X7_2", "conv7x7_3")} for i in (1, 2, 3)]
    heads = []
    for i in range(3):
        wh = torch.zeros(64, 256)
        bh = torch.zeros(64)
        row = 0
        for name in ("ClassHead", "BboxHead", "LandmarkHead"):
            wi, bi_ = sd[f"{name}.{i}.conv1x1.weight"].float().reshape(-1, 256), sd[f"{name}.{i}.conv1x1.bias"].float()
            wh[row:row + wi.shape[0]] = wi
            bh[row:row + wi.shape[0]] = bi_
            row += wi.shape[0]
        assert row == 32
        heads.append(PackedConv(dev(wh), dev(bh, torch.float32), 256, 64, 1, 1))
    out["heads"] = heads
    return out


def pack_conv0_tc(w: torch.Tensor, b: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """wav2vec2 conv0 filter bank [512, 10] + bias [512] -> the B operand of avcer_w2v_conv0_tc: rows
    [w_hi(10) | w_hi(10) | w_lo(10) | b_hi | b_lo] (K = 32, the weight side of a 16-bit x3 product whose activation side is
    [x_hi | x_lo | x_hi | 1 | 1]) stored as UMMA no-swizzle core matrices [k/8][channel/8][8 channels][8 k]."""
    w, b = w.float().reshape(512, 10), b.float().reshape(512, 1)
    w_hi, b_hi = w.to(dtype), b.to(dtype)
    w_lo, b_lo = (w - w_hi.float()).to(dtype), (b - b_hi.float()).to(dtype)
    # IEEE half: the kernel stores 2^11 x_lo (a quiet sample's low half would be a half subnormal), so this block is 2^-11 w_hi
    w_hi2 = (w_hi.float() / 2048.0).to(dtype) if dtype == torch.float16 else w_hi
    rows = torch.cat([w_hi, w_hi2, w_lo, b_hi, b_lo], dim=1)                 # [512, 32]
    return rows.view(64, 8, 4, 8).permute(2, 0, 1, 3).contiguous()


def pack_vd(sd: Dict[str, torch.Tensor], device, dtype: torch.dtype) -> dict:
    """LSTM(512->512) -> LSTM(512->256) -> Linear(256->7).  Layer 2's input and recurrent matrices
    are concatenated along K so one GEMM over [h1_t | h2_{t-1}] yields its gate pre-activations.
    bf16 mode packs every matrix for bf16x3 contractions (split_bf16x3_weight): the recurrence is a chain of 20
    dependent GEMMs whose rounding errors would otherwise compound, and the network is tiny (57.7 MFLOP / window)."""
    def dev(t, dt=None):
        return t.to(device=device, dtype=dt or dtype).contiguous()

    split = dtype in (torch.bfloat16, torch.float16)
    mat = (lambda w: split_bf16x3_weight(w, dtype)) if split else (lambda w: w)
    return {
        "split": split,
        "w_ih1": dev(mat(sd["lstm1.weight_ih_l0"])),
        "w_hh1": dev(mat(sd["lstm1.weight_hh_l0"])),
        "b1": dev(sd["lstm1.bias_ih_l0"] + sd["lstm1.bias_hh_l0"], torch.float32),
        "w_cat2": dev(torch.cat([mat(sd["lstm2.weight_ih_l0"]), mat(sd["lstm2.weight_hh_l0"])], dim=1)),
        "b2": dev(sd["lstm2.bias_ih_l0"] + sd["lstm2.bias_hh_l0"], torch.float32),
        "fc_w": dev(sd["fc.weight"], torch.float32),
        "fc_b": dev(sd["fc.bias"], torch.float32),
    }


def pos_conv_effective_weight(sd: Dict[str, torch.Tensor], prefix: str) -> torch.Tensor:
    """Effective weight of wav2vec2's weight-normed positional conv (weight_norm(dim=2): w = g * v / ||v||, the norm taken
    over (out, in) per kernel tap).  Accepts the three spellings a checkpoint can carry: the parametrizations API of
    current torch (`parametrizations.weight.original0/1`), the older `weight_g` / `weight_v` pair, or a plain `weight`
    saved after torch.nn.utils.remove_weight_norm."""
    for kg, kv in ((".parametrizations.weight.original0", ".parametrizations.weight.original1"), (".weight_g", ".weight_v")):
        if prefix + kg in sd and prefix + kv in sd:
            g, v = sd[prefix + kg].double(), sd[prefix + kv].double()
            return (v * (g / v.norm(p=2, dim=(0, 1), keepdim=True))).float()
    if prefix + ".weight" in sd:
        return sd[prefix + ".weight"].float()
    raise KeyError(f"{prefix}: expected parametrizations.weight.original0/original1, weight_g/weight_v or weight in the "
                   "audio state_dict (wav2vec2 positional conv)")


def pack_audio(sd: Dict[str, torch.Tensor], device, dtype: torch.dtype) -> dict:
    """ExprModelV3 / V2 (wav2vec2-large-robust 12L + tl1 + tl2 + head)."""
    def dev(t, dt=None):
        return t.to(device=device, dtype=dt or dtype).contiguous()

    f32 = torch.float32
    p = "wav2vec2."
    out = {}
    fe = p + "feature_extractor.conv_layers."
    out["conv0_w"] = dev(sd[fe + "0.conv.weight"].reshape(512, 10), f32)
    out["conv0_b"] = dev(sd[fe + "0.conv.bias"], f32)
    if dtype in (torch.bfloat16, torch.float16):
        out["conv0_tc"] = dev(pack_conv0_tc(sd[fe + "0.conv.weight"], sd[fe + "0.conv.bias"], dtype))
    out["conv_ln"] = [(dev(sd[f"{fe}{i}.layer_norm.weight"], f32), dev(sd[f"{fe}{i}.layer_norm.bias"], f32)) for i in range(7)]
    out["convs"] = [(dev(_tap_major(sd[f"{fe}{i}.conv.weight"])), dev(sd[f"{fe}{i}.conv.bias"], f32)) for i in range(1, 7)]
    out["fp_ln"] = (dev(sd[p + "feature_projection.layer_norm.weight"], f32), dev(sd[p + "feature_projection.layer_norm.bias"], f32))
    out["fp_w"] = dev(sd[p + "feature_projection.projection.weight"])
    out["fp_b"] = dev(sd[p + "feature_projection.projection.bias"], f32)
    w = pos_conv_effective_weight(sd, p + "encoder.pos_conv_embed.conv")    # [1024, 64, 128]
    out["pos_w"] = dev(_tap_major(w))                                       # [1024, 128*64]
    out["pos_b"] = dev(sd[p + "encoder.pos_conv_embed.conv.bias"], f32)
    layers = []
    i = 0
    while f"{p}encoder.layers.{i}.layer_norm.weight" in sd:
        q = f"{p}encoder.layers.{i}."
        layers.append({

            "ln1": (dev(sd[q + "layer_norm.weight"], f32), dev(sd[q + "layer_norm.bias"], f32)),
            "wqkv": dev(torch.cat([sd[q + "attention.q_proj.weight"], sd[q + "attention.k_proj.weight"], sd[q + "attention.v_proj.weight"]], 0)),
            "bqkv": dev(torch.cat([sd[q + "attention.q_proj.bias"], sd[q + "attention.k_proj.bias"], sd[q + "attention.v_proj.bias"]], 0), f32),
            "wo": dev(sd[q + "attention.out_proj.weight"]),
            "bo": dev(sd[q + "attention.out_proj.bias"], f32),
            "ln2": (dev(sd[q + "final_layer_norm.weight"], f32), dev(sd[q + "final_layer_norm.bias"], f32)),
            "w1": dev(sd[q + "feed_forward.intermediate_dense.weight"]),
            "b1": dev(sd[q + "feed_forward.intermediate_dense.bias"], f32),
            "w2": dev(sd[q + "feed_forward.output_dense.weight"]),
            "b2": dev(sd[q + "feed_forward.output_dense.bias"], f32),
        })
        i += 1
    out["layers"] = layers
    out["enc_ln"] = (dev(sd[p + "encoder.layer_norm.weight"], f32), dev(sd[p + "encoder.layer_norm.bias"], f32))
    v1 = "gru.weight_ih_l0" in sd                  # ExprModelV1 (audio_8_cl.py:18-72): 2-layer GRU instead of tl1 / tl2
    out["variant"] = "v1" if v1 else "v3"
    if v1:
        # input projections are plain contractions over every time step at once; the 199-step recurrence runs bf16x3
        # (split_bf16x3_weight) in bf16 mode, like the VD LSTM
        rec = (lambda w: split_bf16x3_weight(w, dtype)) if dtype in (torch.bfloat16, torch.float16) else (lambda w: w)
        out["gru"] = [{"w_ih": dev(sd[f"gru.weight_ih_l{l}"]), "b_ih": dev(sd[f"gru.bias_ih_l{l}"], f32),
                       "w_hh": dev(rec(sd[f"gru.weight_hh_l{l}"])), "b_hh": dev(sd[f"gru.bias_hh_l{l}"], f32)} for l in range(2)]
    for t, heads in (() if v1 else (("tl1", 32), ("tl2", 16))):
        a = f"{t}.self_attention."
        out[t] = {
            "heads": heads,
            "pe": dev(sd[f"{t}.positional_encoding.pe"][0, :512]),          # first 512 positions are plenty (T = 199)
            "wqkv": dev(torch.cat([sd[a + "query_w.weight"], sd[a + "keys_w.weight"], sd[a + "values_w.weight"]], 0)),
            "wo": dev(sd[a + "ff_layer_after_concat.weight"]),
            "ln1": (dev(sd[f"{t}.add_norm_after_attention.layer_norm.weight"], f32), dev(sd[f"{t}.add_norm_after_attention.layer_norm.bias"], f32)),
            "w1": dev(sd[f"{t}.feed_forward.layer_1.weight"]), "b1": dev(sd[f"{t}.feed_forward.layer_1.bias"], f32),
            "w2": dev(sd[f"{t}.feed_forward.layer_2.weight"]), "b2": dev(sd[f"{t}.feed_forward.layer_2.bias"], f32),
            "ln2": (dev(sd[f"{t}.add_norm_after_ff.layer_norm.weight"], f32), dev(sd[f"{t}.add_norm_after_ff.layer_norm.bias"], f32)),
        }
    w, b = _fold_bn(sd["time_downsample.0.weight"], sd, "time_downsample.1", 1e-5, sd["time_downsample.0.bias"])
    out["td0_w"], out["td0_b"] = dev(_tap_major(w)), dev(b, f32)
    w, b = _fold_bn(sd["time_downsample.4.weight"], sd, "time_downsample.5", 1e-5, sd["time_downsample.4.bias"])
    out["td4_w"], out["td4_b"] = dev(_tap_major(w)), dev(b, f32)
    out["fd_w"] = dev(sd["feature_downsample.weight"], f32)
    out["fd_b"] = dev(sd["feature_downsample.bias"], f32)
    out["num_classes"] = int(sd["feature_downsample.weight"].shape[0])
    out["f_size"] = int(sd["feature_downsample.weight"].shape[1])          # 1024 (V2 / V3) or 256 (V1)
    return out
