"""GPU probe: VS / VD / A device forwards against the oracle restatements (development aid)."""
import sys, os, time, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from avcer_b200 import ops, synthetic as syn, nets, pipeline
from oracle import video as ov, audio as oa

dev = "cuda:0"
torch.manual_seed(0)
which = sys.argv[1:] or ["vs", "vd", "a"]


def nhwc(t):  # oracle NCHW -> NHWC
    return t.permute(0, 2, 3, 1).contiguous()


def rep(name, got, ref):
    got = got.float().cpu(); ref = ref.float()
    d = (got - ref).abs()
    print(f"   {name:10s} max|d|={d.max().item():.4g} mean|d|={d.mean().item():.3g} ref_std={ref.std().item():.3g}", flush=True)


if "vs" in which:
    for init in ("spread", "default"):
        sd = syn.make_vs_state_dict(0, init)
        crops = syn.make_crops(11, 6)
        x_ref = torch.from_numpy(np.concatenate([ov.pth_processing(c) for c in crops]))
        taps_o = {}
        logits_o, feat_o = ov.resnet50_forward(sd, x_ref, taps_o)
        probs_o = torch.softmax(logits_o, 1)
        for prec in ("fp32", "bf16"):
            try:
                net = nets.VSNet(sd, prec, dev)
                x = net.alloc_input(6)
                ops.preprocess(torch.from_numpy(crops).to(dev), 6, x, net.input_layout)
                # K1 check (fp32 NCHW layout 0 is bit exact)
                if prec == "fp32":
                    nchw = torch.empty((6, 3, 224, 224), device=dev)
                    ops.preprocess(torch.from_numpy(crops).to(dev), 6, nchw, 0)
                    print("K1 layout0 bit-exact:", torch.equal(nchw.cpu(), x_ref), flush=True)
                taps = {}
                probs, feat = net.forward(x, taps)
                torch.cuda.synchronize()
                print(f"VS {init} {prec}: max|dp|={(probs.cpu()-probs_o).abs().max().item():.3g}", flush=True)
                rep("stem", taps["stem"], nhwc(taps_o["stem"]))
                rep("pool", taps["pool"], nhwc(taps_o["pool"]))
                for li, bi in (("layer1", 2), ("layer2", 6), ("layer3", 12), ("layer4", 15)):
                    rep(li, taps[f"block{bi}"], nhwc(taps_o[li]))
                rep("feat", feat, torch.relu(feat_o))
                rep("probs", probs, probs_o)
            except Exception:
                traceback.print_exc(); sys.stdout.flush()

if "vd" in which:
    sd = syn.make_vd_state_dict(1)
    feats = torch.relu(torch.randn(40, 512))
    ex = np.ones(200, dtype=bool); ex[[37, 38, 90]] = False
    plan = pipeline.plan_video(ex, 5)
    wins = plan.windows
    ref = ov.lstm_forward(sd, feats[torch.from_numpy(wins)])
    for prec in ("fp32", "bf16"):
        try:
            net = nets.VDNet(sd, prec, dev)
            f = feats.to(dev).to(net.dtype)
            out = net.forward(f, torch.from_numpy(np.ascontiguousarray(wins.T).astype(np.int32)).to(dev))
            torch.cuda.synchronize()
            print(f"VD {prec}: M={wins.shape[0]}", flush=True)
            rep("logits", out, ref)
            rep("probs", torch.softmax(out, 1), torch.softmax(ref, 1))
        except Exception:
            traceback.print_exc(); sys.stdout.flush()

if "a" in which:
    for ncls in (8,):
        sd = syn.make_audio_state_dict(2, ncls, "spread", 12)
        wav = syn.make_wav(3, 64000 * 2 + 777)
        plan = pipeline.plan_audio(len(wav), 25)
        sel = [0, 3, len(plan.starts) - 2]
        xs = np.stack([oa.zero_mean_unit_var(oa.pad_window(wav[plan.starts[i]:plan.ends[i]], 64000, "mean")) for i in sel])
        taps_o = {}
        t0 = time.time()
        ref = oa.audio_model_forward(sd, torch.from_numpy(xs), taps_o)
        print("oracle audio time", time.time() - t0, flush=True)
        for prec in ("fp32", "bf16"):
            try:
                net = nets.ANet(sd, prec, dev)
                wavd = torch.from_numpy(wav).to(dev)
                x = ops.audio_normalize_windows(wavd, torch.from_numpy(plan.starts[sel]).to(dev), 64000, "mean")
                rep("normalize", x, torch.from_numpy(xs))
                taps = {}
                out = net.forward(x, taps)
                torch.cuda.synchronize()
                print(f"A {ncls}cl {prec}:", flush=True)
                for k in ("conv0", "conv6", "proj", "posconv", "layer0", "layer5", "layer11", "w2v", "tl2"):
                    rep(k, taps[k].view(taps_o[k].shape), taps_o[k])
                rep("logits", out, ref)
                rep("probs7", torch.softmax(out[:, :7], 1), torch.softmax(ref[:, :7], 1))
            except Exception:
                traceback.print_exc(); sys.stdout.flush()
print("probe done", flush=True)
