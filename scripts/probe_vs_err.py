"""GPU probe: VS bf16 probability error vs the golden fp32 reference for both inits; run with AVCER_CONV3=0/1 etc."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from avcer_b200 import nets, ops, synthetic as syn
g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "video.npz"))
crops = syn.make_crops(11, 6)
for init in ("spread", "default"):
    net = nets.VSNet(syn.make_vs_state_dict(0, init), "bf16", "cuda:0")
    x = net.alloc_input(6)
    ops.preprocess(torch.from_numpy(crops).to("cuda:0"), 6, x, net.input_layout)
    probs, feat = net.forward(x)
    err = np.abs(probs.cpu().numpy() - g[f"vs_{init}_probs"])
    print(f"{init}: max|dp|={err.max():.5f} mean={err.mean():.5f}  CONV3={os.environ.get('AVCER_CONV3')} FLAT={os.environ.get('AVCER_FLAT')}")
