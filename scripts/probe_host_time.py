import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from avcer_b200 import ops, synthetic as syn, get_weights_matrices as gwm
from avcer_b200.pipeline import Engine
dev = "cuda:0"
eng = Engine(syn.make_vs_state_dict(0, "default"), syn.make_vd_state_dict(1), syn.make_audio_state_dict(2, 8, "spread", 12), device=dev)
c, nf, ns = 4, 1500, 960000
g = torch.Generator(device=dev).manual_seed(0)
crops = torch.randint(0, 256, (c * nf, 224, 224, 3), dtype=torch.uint8, device=dev, generator=g)
wav = (torch.randn(c * ns, device=dev, generator=g) * 0.1)
exists = [np.ones(nf, bool)] * c; fps = [25.0] * c; lens = [ns] * c
w1 = gwm.class_weights(gwm.weights_3)
def t(label, fn):
    t0 = time.perf_counter(); r = fn(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"{label:14s} host {1e3*(t1-t0):8.2f} ms   +sync {1e3*(t2-t1):8.2f} ms", flush=True); return r
for it in range(3):
    print("iter", it)
    probs, feats = t("vs", lambda: eng.vs_forward_u8(crops))
    stat, dyn, plans = t("video_rows", lambda: eng.video_rows(probs, feats, exists, fps))
    a_rows, logits = t("audio_rows", lambda: eng.audio_rows(wav, lens, fps, [nf] * c))
    lab = t("fuse", lambda: eng.fuse(stat, dyn, a_rows, w1, [1, 1, 1], False, True))
