"""Drop-in for the reference's src/get_prob_audio_8_cl.py (8-class ExprModelV3): same class and
function names, same long-format DataFrame / CSV; the per-window loop (:78-101) is replaced by
K5a window gather/normalise + batched audio-network launches.
"""
from __future__ import annotations

import os

import numpy as np
import pandas as pd
import torch

from . import config
from .data.utils import convert_mp4_to_mp3
from .pipeline import AUDIO_ORDER, plan_audio
from .tables import AudioTable

NUM_CLASSES = 8
DEFAULT_MODEL = {"model_name": "FLW-ExprModelV3-2024.03.02-11.42.11", "model_cls": "ExprModelV3", "epoch": 63}


class EmotionRecognition:
    num_classes = NUM_CLASSES

    def __init__(self, step=2, window=4, sr=16000, device="cuda:0", model_params={}, save_path="", padding="",
                 flag_save_prob=True):
        self.model_params = model_params
        self.save_path = save_path
        self.step = step
        self.window = window
        self.sr = sr
        self.device = device
        self.padding = padding
        self.flag_save_prob = flag_save_prob
        self.load_models()

    def predict_emotion(self, path, fps):
        return self.load_audio_features(path, fps)

    def load_models(self):
        self.load_audio_model()

    def load_audio_model(self):
        self.audio_model = config.audio_net(self.model_params["model_name"], self.num_classes,
                                            self.model_params["root_path"], self.model_params["epoch"], self.device)

    def load_audio_features(self, path, fps):
        from . import ops

        wav = convert_mp4_to_mp3(path, self.sr)
        wav_d = torch.as_tensor(np.asarray(wav, dtype=np.float32)).to(self.audio_model.device).contiguous()
        plan = plan_audio(int(wav_d.numel()), fps, self.step, self.window, self.sr)
        if self.padding == "repeat" and bool((plan.ends - plan.starts == 0).any()):
            raise ZeroDivisionError("integer division or modulo by zero")       # data/utils.py:66
        if self.padding not in ("mean", "constant", "repeat"):
            raise UnboundLocalError("cannot access local variable 'a_fss' where it is not associated with a value")
        win = self.window * self.sr
        starts = torch.from_numpy(plan.starts).to(wav_d.device)
        logits = []
        batch = 32
        for s in range(0, len(plan.starts), batch):
            x = ops.audio_normalize_windows(wav_d, starts[s:s + batch], win, self.padding)
            logits.append(self.audio_model.forward(x))
        logits = torch.cat(logits, 0)                                            # [Wn, ncls] on the device
        emo = AUDIO_ORDER[:7] if logits.shape[1] == 7 else AUDIO_ORDER
        # One row per (window, covered frame) with a `frames` string column is what the reference returns (:94-126); the
        # façade holds the windows and their frame ranges and only builds that table if somebody looks at it (tables.py)
        df = AudioTable(logits, plan.f_lo, plan.f_hi, emo)
        if self.flag_save_prob:
            save_path = os.path.join(self.save_path, self.model_params["model_name"])
            os.makedirs(save_path, exist_ok=True)
            name_video = os.path.basename(path[:-4])
            if name_video == "135-24-1920x1080":
                name_video = "135-24-1920x1080_left"
            elif name_video == "6-30-1920x1080":
                name_video = "6-30-1920x1080_right"
            df.to_csv(os.path.join(save_path, "{}.csv".format(name_video)), index=False)
        return df


def preprocess_audio_and_predict(path_video="", path_weights="", save_path="src/pred_results/C-EXPR-DB", fps=25, step=0.5,
                                 padding="mean", flag_save_prob=False, window=4, sr=16000, device="cuda:0"):
    model_params = dict(DEFAULT_MODEL)
    model_params["root_path"] = os.path.join(path_weights, model_params["model_name"])
    audio_ER = EmotionRecognition(step=step, window=window, sr=sr, device=device, model_params=model_params,
                                  save_path=save_path, padding=padding, flag_save_prob=flag_save_prob)
    return audio_ER.predict_emotion(path_video, fps)
