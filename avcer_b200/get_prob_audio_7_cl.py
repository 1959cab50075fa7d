"""Drop-in for the reference's src/get_prob_audio_7_cl.py (7-class ExprModelV2, epoch 51).
Differences to the 8-class driver that are preserved: save_path gets "audio_{padding}_{step}/"
appended (:153) and the default model is 7cl-FLW-ExprModelV2 (:155-159).  The reference's
constructor forgets to store flag_save_prob (:35-42 vs :127 -> AttributeError); here the attribute
is stored, which is the behaviour the reference's own caller expects.
"""
from __future__ import annotations

import os

from . import get_prob_audio_8_cl as _a8

NUM_CLASSES = 7
DEFAULT_MODEL = {"model_name": "7cl-FLW-ExprModelV2-2024.03.04-11.52.11", "model_cls": "ExprModelV2", "epoch": 51}


class EmotionRecognition(_a8.EmotionRecognition):
    num_classes = NUM_CLASSES


def preprocess_audio_and_predict(path_video="", path_weights="", save_path="src/pred_results/C-EXPR-DB", fps=25, step=0.5,
                                 padding="mean", flag_save_prob=False, window=4, sr=16000, device="cuda:0"):
    save_path = os.path.join(save_path, "audio_{}_{}/".format(padding, step))
    model_params = dict(DEFAULT_MODEL)
    model_params["root_path"] = os.path.join(path_weights, model_params["model_name"])
    audio_ER = EmotionRecognition(step=step, window=window, sr=sr, device=device, model_params=model_params,
                                  save_path=save_path, padding=padding, flag_save_prob=flag_save_prob)
    return audio_ER.predict_emotion(path_video, fps)
