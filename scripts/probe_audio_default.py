import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from avcer_b200 import ops, synthetic as syn, nets, pipeline
from oracle import audio as oa
dev = "cuda:0"
init = sys.argv[1] if len(sys.argv) > 1 else "default"
sd = syn.make_audio_state_dict(2, 8, init, 12)
wav = syn.make_wav(31, 52800 + 123)
ap = pipeline.plan_audio(len(wav), 25, 0.5)
xs = np.stack([oa.zero_mean_unit_var(oa.pad_window(wav[s:e], 64000, "mean")) for s, e in zip(ap.starts, ap.ends)])
taps_o = {}
ref = oa.audio_model_forward(sd, torch.from_numpy(xs), taps_o)
for prec in ("fp32",):
    net = nets.ANet(sd, prec, dev)
    x = ops.audio_normalize_windows(torch.from_numpy(wav).to(dev), torch.from_numpy(ap.starts).to(dev), 64000, "mean")
    print("normalize per-window err", (x.cpu() - torch.from_numpy(xs)).abs().amax(1).numpy())
    taps = {}
    out = net.forward(x, taps)
    for k in ("conv0", "conv6", "proj", "posconv", "layer0", "layer5", "layer11", "w2v", "tl2"):
        g = taps[k].float().cpu().view(taps_o[k].shape); o = taps_o[k]
        print(k, "per-window max err", (g - o).abs().flatten(1).amax(1).numpy().round(6), "std", float(o.std()))
    print("logits err per window", (out.cpu() - ref).abs().amax(1).numpy())

# layer-by-layer through the feature extractor (fp32), reporting where each window first deviates
import torch.nn.functional as F
net = nets.ANet(sd, "fp32", dev)
w = net.w
h = torch.empty((7, 12799, 512), device=dev)
ops.w2v_conv0_ln_gelu(x, w["conv0_w"], w["conv0_b"], *w["conv_ln"][0], h)
ho = taps_o["conv0"]
t = 12799
for i in range(1, 7):
    wt, bias = w["convs"][i - 1]
    y = net._conv1d_s2(h, t, nets.W2V_KERNELS[i], wt, bias)
    q = f"wav2vec2.feature_extractor.conv_layers.{i}"
    yo = F.conv1d(ho.transpose(1, 2), sd[q + ".conv.weight"], sd[q + ".conv.bias"], stride=2).transpose(1, 2)
    d = (y.cpu() - yo).abs()
    print(f"conv{i} pre-LN per-window err", d.flatten(1).amax(1).numpy().round(5), "ref absmax", float(yo.abs().max()))
    t = y.shape[1]
    y2 = y.view(7 * t, 512)
    g, be = w["conv_ln"][i]
    ops.layernorm(y2, g, be, 1e-5, act=ops.ACT_GELU, out=y2)
    lo = F.gelu(F.layer_norm(yo, (512,), sd[q + ".layer_norm.weight"], sd[q + ".layer_norm.bias"], 1e-5))
    d = (y.cpu() - lo).abs()
    bad = d.flatten(1).argmax(1)
    print(f"conv{i} post-LN per-window err", d.flatten(1).amax(1).numpy().round(5), "argmax (t,c)", [(int(b) // 512, int(b) % 512) for b in bad][:4])
    bw = int(d.flatten(1).amax(1).argmax())
    tt, cc = int(bad[bw]) // 512, int(bad[bw]) % 512
    print("   worst window", bw, "row", tt, "ch", cc, "got", float(y[bw, tt, cc]), "ref", float(lo[bw, tt, cc]), "pre-LN row var", float(yo[bw, tt].var()))
    h, ho = y, lo
