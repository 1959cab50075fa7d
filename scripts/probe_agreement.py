"""Compound top-1 agreement of the whole bf16 / fp32 pipeline against the oracle on one synthetic clip."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from avcer_b200 import synthetic as syn, get_weights_matrices as gwm, ops
from avcer_b200.pipeline import Engine
from oracle import audio as oa, fusion as of, video as ov
import pandas as pd

n, fps = int(sys.argv[1]) if len(sys.argv) > 1 else 300, 25
init = sys.argv[2] if len(sys.argv) > 2 else "spread"
exists = np.ones(n, bool); exists[[40, 41, 133]] = False
crops = syn.make_crops(500, n)
# make the clip vary over time: blend towards different base images
wav = syn.make_wav(501, int(n / fps * 16000) - 160)
sd_vs, sd_vd, sd_a = syn.make_vs_state_dict(0, init), syn.make_vd_state_dict(1, init), syn.make_audio_state_dict(2, 8, init, 12)
frames = [crops[i] if exists[i] else None for i in range(n)]
t0 = time.time()
o_dyn, o_stat = ov.predict_video(frames, fps, sd_vs, sd_vd)
rows, ids, o_logits = oa.predict_audio(wav, fps, sd_a)
print("oracle time", time.time() - t0, flush=True)
stat_df = pd.DataFrame(o_stat, columns=of.VIDEO_ORDER); dyn_df = pd.DataFrame(o_dyn, columns=of.VIDEO_ORDER)
audio_df = pd.DataFrame(rows, columns=of.AUDIO_ORDER); audio_df["frames"] = [str(i).zfill(6) + ".jpg" for i in ids]
for prec in ("fp32", "bf16"):
    eng = Engine(sd_vs, sd_vd, sd_a, precision=prec, device="cuda:0", use_graphs=False)
    for tag, w1, w2, cwt, cm in (("w3 rule1", gwm.class_weights(gwm.weights_3), [1, 1, 1], False, True),
                                 ("w3 rule2", gwm.class_weights(gwm.weights_3), [1, 1, 1], True, False),
                                 ("none rule1", None, [1, 1, 1], False, True)):
        out = eng.run_clips(torch.from_numpy(crops[exists]), [exists], [fps], torch.from_numpy(wav), [len(wav)], w1, w2, cwt, cm)
        ref = np.stack(of.get_c_expr_db_pred(stat_df, dyn_df, audio_df, "c", w1, w2, cwt, cm)[:4])
        got = out["labels"].cpu().numpy()
        agree = (got == ref).mean(axis=1)
        print(f"{prec} {tag:10s} agreement AV/VS/VD/A = {agree.round(4).tolist()}  max|dP_vs|={np.abs(out['stat'].cpu().numpy()-o_stat).max():.2e}", flush=True)
