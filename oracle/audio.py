"""ORACLE (test infrastructure, never shipped on the product path).

CPU restatement (numpy + torch.nn.functional, fp32) of the audio branch of ElenaRyumina/AVCER:
  * window schedule, padding, frame ranges -- src/get_prob_audio_8_cl.py:68-126, src/data/utils.py:63-89
  * HF feature-extractor normalisation     -- third-party `transformers` (pinned 4.36.2 in
        src/requirements.txt; not vendored under /root/reference):
        feature_extraction_wav2vec2.py zero_mean_unit_var_norm: (x - mean) / sqrt(var + 1e-7)
  * wav2vec2-large-robust forward (12 layers, stable layer norm) -- `transformers`
        modeling_wav2vec2.py: Wav2Vec2LayerNormConvLayer, Wav2Vec2FeatureProjection,
        Wav2Vec2PositionalConvEmbedding, Wav2Vec2EncoderLayerStableLayerNorm,
        Wav2Vec2EncoderStableLayerNorm (published algorithm restated below)
  * TransformerLayer x2                    -- src/architectures/attention_layers.py:10-267
  * time_downsample / feature_downsample   -- src/architectures/audio_8_cl.py:146-159,179-190

The reference has no test that pins the third-party arithmetic ("parity unpinned" upstream); this
restatement is pinned here against the installed `transformers` (5.5.0) driven through the
reference's own ExprModelV3 / ExprModelV2 classes by oracle/make_golden.py -> tests/golden/audio_*.npz.
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import numpy as np
import torch
import torch.nn.functional as F

CONV_KERNEL = (10, 3, 3, 3, 3, 2, 2)
CONV_STRIDE = (5, 2, 2, 2, 2, 2, 2)
N_HEADS_W2V = 16


# ------------------------------------------------------------------------------------------ windows
def window_schedule(n_samples: int, fps: float, step: float = 0.5, window: int = 4, sr: int = 16000):
    """[(start, end, frame_lo, frame_hi)] for every window of get_prob_audio_8_cl.py:78-101.
    frame ids covered by a window: range(round(start/sr*fps), round(end/sr*fps + 1))."""
    win = window * sr
    step_a = int(step * sr)
    out = []
    for start in range(0, n_samples + 1, step_a):
        end = min(start + win, n_samples)
        out.append((start, end, round(start / sr * fps), round(end / sr * fps + 1)))
    return out


def pad_window(chunk: np.ndarray, win: int, padding: str) -> np.ndarray:
    """utils.py:63-89.  'mean': pad with the chunk mean (NaN for an empty chunk); 'constant': zeros;
    'repeat': tile and cut (ZeroDivisionError for an empty chunk, like the reference)."""
    chunk = np.asarray(chunk, dtype=np.float32)
    n = len(chunk)
    if padding == "repeat":
        if n < win:
            reps = (win + n - 1) // n
            return np.concatenate([chunk] * reps)[:win]
        return chunk[:win]
    if padding == "mean":
        val = torch.mean(torch.from_numpy(chunk)).item() if n else float("nan")
    elif padding == "constant":
        val = 0.0
    else:
        raise ValueError(padding)
    out = np.full(win, val, dtype=np.float32)
    out[:n] = chunk
    return out


def zero_mean_unit_var(x: np.ndarray) -> np.ndarray:
    """HF Wav2Vec2FeatureExtractor (do_normalize=True, no attention mask): float32 numpy."""
    x = np.asarray(x, dtype=np.float32)
    return (x - x.mean()) / np.sqrt(x.var() + 1e-7)


# ------------------------------------------------------------------------------------------ wav2vec2
def _ln(x, sd, p, eps=1e-5):
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], eps)


def pos_conv_weight(sd) -> torch.Tensor:
    """weight_norm(dim=2): w = g * v / ||v|| with the norm over (out, in) per kernel tap."""
    g = sd["wav2vec2.encoder.pos_conv_embed.conv.parametrizations.weight.original0"]
    v = sd["wav2vec2.encoder.pos_conv_embed.conv.parametrizations.weight.original1"]
    return v * (g / v.norm(p=2, dim=(0, 1), keepdim=True))


def _mha(x, wq, bq, wk, bk, wv, bv, wo, bo, heads):
    b, t, d = x.shape
    dh = d // heads
    q = F.linear(x, wq, bq).view(b, t, heads, dh).transpose(1, 2)
    k = F.linear(x, wk, bk).view(b, t, heads, dh).transpose(1, 2)
    v = F.linear(x, wv, bv).view(b, t, heads, dh).transpose(1, 2)
    a = torch.softmax((q @ k.transpose(-1, -2)) / math.sqrt(dh), dim=-1)
    o = (a @ v).transpose(1, 2).reshape(b, t, d)
    return F.linear(o, wo, bo)


def wav2vec2_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, taps=None) -> torch.Tensor:
    """x: [B, 64000] normalised waveform -> last_hidden_state [B, 199, 1024]."""
    p = "wav2vec2."
    h = x[:, None, :]
    for i, s in enumerate(CONV_STRIDE):
        q = f"{p}feature_extractor.conv_layers.{i}"
        h = F.conv1d(h, sd[q + ".conv.weight"], sd[q + ".conv.bias"], stride=s)
        h = _ln(h.transpose(1, 2), sd, q + ".layer_norm").transpose(1, 2)
        h = F.gelu(h)
        if taps is not None:
            taps[f"conv{i}"] = h.transpose(1, 2)
    h = h.transpose(1, 2)                                                    # [B, T, 512]
    h = _ln(h, sd, p + "feature_projection.layer_norm")
    h = F.linear(h, sd[p + "feature_projection.projection.weight"], sd[p + "feature_projection.projection.bias"])
    if taps is not None:
        taps["proj"] = h
    pc = F.conv1d(h.transpose(1, 2), pos_conv_weight(sd), sd[p + "encoder.pos_conv_embed.conv.bias"], padding=64, groups=16)
    pc = F.gelu(pc[:, :, :-1]).transpose(1, 2)                               # even kernel: drop last frame
    h = h + pc
    if taps is not None:
        taps["posconv"] = h
    n_layers = 0
    while f"{p}encoder.layers.{n_layers}.layer_norm.weight" in sd:
        n_layers += 1
    for i in range(n_layers):
        q = f"{p}encoder.layers.{i}"
        a = _ln(h, sd, q + ".layer_norm")
        a = _mha(a, sd[q + ".attention.q_proj.weight"], sd[q + ".attention.q_proj.bias"],
                 sd[q + ".attention.k_proj.weight"], sd[q + ".attention.k_proj.bias"],
                 sd[q + ".attention.v_proj.weight"], sd[q + ".attention.v_proj.bias"],
                 sd[q + ".attention.out_proj.weight"], sd[q + ".attention.out_proj.bias"], N_HEADS_W2V)
        h = h + a
        f = _ln(h, sd, q + ".final_layer_norm")
        f = F.gelu(F.linear(f, sd[q + ".feed_forward.intermediate_dense.weight"], sd[q + ".feed_forward.intermediate_dense.bias"]))
        f = F.linear(f, sd[q + ".feed_forward.output_dense.weight"], sd[q + ".feed_forward.output_dense.bias"])
        h = h + f
        if taps is not None:
            taps[f"layer{i}"] = h
    return _ln(h, sd, p + "encoder.layer_norm")


def transformer_layer(sd, x: torch.Tensor, name: str, heads: int) -> torch.Tensor:
    """attention_layers.py:221-267: PE added to query, key and value; residual is the PE'd query;
    bias-free projections; post-LN; FFN 1024->1024 ReLU 1024."""
    t = x.shape[1]
    xp = x + sd[f"{name}.positional_encoding.pe"][:, :t]
    a = _mha(xp, sd[f"{name}.self_attention.query_w.weight"], None, sd[f"{name}.self_attention.keys_w.weight"], None,
             sd[f"{name}.self_attention.values_w.weight"], None, sd[f"{name}.self_attention.ff_layer_after_concat.weight"], None, heads)
    y = _ln(a + xp, sd, f"{name}.add_norm_after_attention.layer_norm")
    f = F.linear(F.relu(F.linear(y, sd[f"{name}.feed_forward.layer_1.weight"], sd[f"{name}.feed_forward.layer_1.bias"])),
                 sd[f"{name}.feed_forward.layer_2.weight"], sd[f"{name}.feed_forward.layer_2.bias"])
    return _ln(f + y, sd, f"{name}.add_norm_after_ff.layer_norm")


def _bn1d(x, sd, p):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"], training=False, eps=1e-5)


def gru_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, layers: int = 2) -> torch.Tensor:
    """nn.GRU(1024 -> 256, 2 layers, batch_first) from zero state, eval mode (inter-layer dropout inactive); PyTorch gate
    order r, z, n (architectures/audio_8_cl.py:23-29):  n = tanh(W_in x + b_in + r * (W_hn h + b_hn))."""
    for layer in range(layers):
        w_ih, w_hh = sd[f"gru.weight_ih_l{layer}"], sd[f"gru.weight_hh_l{layer}"]
        b_ih, b_hh = sd[f"gru.bias_ih_l{layer}"], sd[f"gru.bias_hh_l{layer}"]
        hid = w_hh.shape[1]
        h = x.new_zeros(x.shape[0], hid)
        outs = []
        for t in range(x.shape[1]):
            gi = x[:, t] @ w_ih.t() + b_ih
            gh = h @ w_hh.t() + b_hh
            i_r, i_z, i_n = gi.split(hid, dim=1)
            h_r, h_z, h_n = gh.split(hid, dim=1)
            r = torch.sigmoid(i_r + h_r)
            z = torch.sigmoid(i_z + h_z)
            n = torch.tanh(i_n + r * h_n)
            h = (1 - z) * n + z * h
            outs.append(h)
        x = torch.stack(outs, dim=1)
    return x


def audio_model_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, taps=None) -> torch.Tensor:
    """ExprModelV3 / ExprModelV2 forward (audio_8_cl.py:179-190): x [B,64000] -> logits [B, 8 or 7].
    A state_dict with `gru.*` keys is ExprModelV1 (audio_8_cl.py:18-72): wav2vec2 -> 2-layer GRU -> the same head at 256."""
    with torch.no_grad():
        h = wav2vec2_forward(sd, x, taps)
        if taps is not None:
            taps["w2v"] = h
        if "gru.weight_ih_l0" in sd:
            h = gru_forward(sd, h)
            if taps is not None:
                taps["gru"] = h
        else:
            h = transformer_layer(sd, h, "tl1", 32)
            h = transformer_layer(sd, h, "tl2", 16)
            if taps is not None:
                taps["tl2"] = h
        h = h.permute(0, 2, 1)
        h = F.conv1d(h, sd["time_downsample.0.weight"], sd["time_downsample.0.bias"], stride=3, dilation=2)
        h = F.relu(F.max_pool1d(_bn1d(h, sd, "time_downsample.1"), 5))
        h = F.conv1d(h, sd["time_downsample.4.weight"], sd["time_downsample.4.bias"])
        h = F.relu(_bn1d(h, sd, "time_downsample.5").mean(dim=2))
        return F.linear(h, sd["feature_downsample.weight"], sd["feature_downsample.bias"])


def predict_audio(wav: np.ndarray, fps: float, sd, step: float = 0.5, window: int = 4, sr: int = 16000,
                  padding: str = "mean", batch: int = 4):
    """Long-format result of EmotionRecognition.load_audio_features (get_prob_audio_8_cl.py:68-126):
    (logit_rows [R, ncls] float32, frame_ids [R] int) with one row per (window, covered frame),
    plus the per-window logits [Wn, ncls]."""
    sched = window_schedule(len(wav), fps, step, window, sr)
    win = window * sr
    xs = np.stack([zero_mean_unit_var(pad_window(wav[s:e], win, padding)) for (s, e, _, _) in sched])
    logits = []
    for i in range(0, len(xs), batch):
        logits.append(audio_model_forward(sd, torch.from_numpy(xs[i:i + batch])).numpy())
    logits = np.concatenate(logits, axis=0)
    rows, ids = [], []
    for (s, e, lo, hi), l in zip(sched, logits):
        for f in range(lo, hi):
            rows.append(l)
            ids.append(f)
    return np.asarray(rows, dtype=np.float32), np.asarray(ids, dtype=np.int64), logits


# ---------------------------------------------------------------------------------------------- decode seam
def sinc_resample_kernel(orig_freq: int, new_freq: int, lowpass_filter_width: int = 6, rolloff: float = 0.99):
    """Filter bank of torchaudio.transforms.Resample with its defaults ("sinc_interp_hann"), the transform the
    reference applies at src/data/utils.py:53-55.  torchaudio is a third-party dependency (pinned
    torchaudio==2.1.2 in src/requirements.txt, not under /root/reference); its published algorithm
    (torchaudio.functional._get_sinc_resample_kernel) is restated here: float64 table, cast to float32.
    Returns (kernel [new, 2*width + orig] float32, width, orig, new) with orig / new reduced by their gcd."""
    g = math.gcd(int(orig_freq), int(new_freq))
    orig, new = int(orig_freq) // g, int(new_freq) // g
    base_freq = min(orig, new) * rolloff
    width = math.ceil(lowpass_filter_width * orig / base_freq)
    idx = np.arange(-width, width + orig, dtype=np.float64)[None, :] / orig
    t = np.arange(0, -new, -1, dtype=np.float64)[:, None] / new + idx
    t *= base_freq
    t = np.clip(t, -lowpass_filter_width, lowpass_filter_width)
    window = np.cos(t * math.pi / lowpass_filter_width / 2) ** 2
    t *= math.pi
    scale = base_freq / orig
    with np.errstate(invalid="ignore", divide="ignore"):
        kern = np.where(t == 0, 1.0, np.sin(t) / t)
    kern = kern * window * scale
    return kern.astype(np.float32), width, orig, new


def pcm16_to_mono_16k(pcm: np.ndarray, sr: int, sampling_rate: int = 16000) -> np.ndarray:
    """convert_mp4_to_mp3 after the ffmpeg step (src/data/utils.py:49-60): `pcm` is the int16 [n, channels] content
    of the .wav; torchaudio.load scales by 1/32768, the channels are averaged, and the signal is resampled to
    `sampling_rate` (strided conv1d with the filter bank above, torchaudio.functional._apply_sinc_resample_kernel)."""
    wav = torch.from_numpy(pcm.astype(np.float32) / 32768.0).t()           # [channels, n]
    if wav.size(0) > 1:
        wav = wav.mean(dim=0, keepdim=True)
    if sr == sampling_rate:
        return wav.squeeze(0).numpy()
    kern, width, orig, new = sinc_resample_kernel(sr, sampling_rate)
    length = wav.shape[1]
    padded = F.pad(wav, (width, width + orig))
    res = F.conv1d(padded[:, None], torch.from_numpy(kern)[:, None], stride=orig)      # [1, new, frames]
    res = res.transpose(1, 2).reshape(1, -1)
    target = int(math.ceil(new * length / orig))
    return res[0, :target].numpy()
