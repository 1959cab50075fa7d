#!/usr/bin/env python
"""Reduced forward of every hot-path kernel family for compute-sanitizer (scripts/sanitize.sh): K1, the VS ResNet-50 on 2
crops (stem+pool, halo 3x3, one- and two-SM contractions, FLAT / ring residual epilogues), the VD recurrence, the audio
network on 1 window with 2 encoder layers (conv0, conv1-6, positional conv, tcgen05 + mma.sync attention, LayerNorms,
head), alignment glue and K4.  Eager launches (no CUDA graphs: the sanitizer instruments kernels individually)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from avcer_b200 import get_weights_matrices as gwm, synthetic as syn      # noqa: E402
from avcer_b200.pipeline import Engine                                     # noqa: E402

n, fps = 12, 25
exists = np.ones(n, bool)
exists[5] = False
crops = syn.make_crops(3, int(exists.sum()))
wav = syn.make_wav(4, 8000 - 160)
eng = Engine(syn.make_vs_state_dict(0, "default"), syn.make_vd_state_dict(1), syn.make_audio_state_dict(2, 8, "spread", 2),
             precision="bf16", device="cuda:0", use_graphs=False)
w1, w2 = gwm.class_weights(gwm.weights_3), [1, 1, 1]
out = eng.run_clips(torch.from_numpy(crops), [exists], [fps], torch.from_numpy(wav), [len(wav)], w1, w2, False, True)
torch.cuda.synchronize()
print("sanitize target ok: labels", out["labels"][0].tolist())
