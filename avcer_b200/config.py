"""Run-time configuration of the drop-in modules (precision, device, weights).

The reference loads its checkpoints at import time from CWD-relative paths
(src/get_prob_video.py:22-25,51-54; src/get_prob_audio_8_cl.py:52-66).  Here loading is lazy:
`set_state_dicts` injects state_dicts directly (tests, synthetic runs); otherwise the same files
are read on first use and a missing file raises FileNotFoundError like the reference.
"""
from __future__ import annotations

import os
from typing import Dict, Optional

import torch

PATH_STATIC = "src/weights/FER_static_ResNet50_AffectNet.pt"       # get_prob_video.py:22
PATH_DYNAMIC = "src/weights/FER_dinamic_LSTM_Aff-Wild2.pt"         # get_prob_video.py:51

_state: Dict[str, object] = {"precision": "bf16", "device": "cuda:0", "vs": None, "vd": None, "audio": {},
                             "engine": None, "audio_nets": {}, "jpeg": "gpu", "face": None, "face_predictor": None}


def set_jpeg_decoder(which: str) -> None:
    """"gpu" (default): the face crops are decoded by avcer_jpeg_decode, bit-identical to cv2.imread; "cv2": decoded on the
    host with cv2.imread like the reference (get_prob_video.py:95) -- for files outside the GPU decoder's coverage."""
    assert which in ("gpu", "cv2")
    _state["jpeg"] = which


def jpeg_decoder() -> str:
    return _state["jpeg"]


def set_precision(precision: str) -> None:
    assert precision in ("bf16", "fp16", "fp32")
    _state["precision"] = precision
    reset()


def set_device(device: str) -> None:
    _state["device"] = device
    reset()


def set_state_dicts(vs=None, vd=None, audio: Optional[dict] = None, face=None) -> None:
    """audio: {model_name or num_classes: state_dict}; face: the RetinaFace-ResNet50 state_dict of the face detector
    (otherwise read from data/weights/Resnet50_Final.pth on first use, like the reference's ibug package)."""
    if face is not None:
        _state["face"] = face
    if vs is not None:
        _state["vs"] = vs
    if vd is not None:
        _state["vd"] = vd
    if audio is not None:
        _state["audio"].update(audio)
    reset()


def reset() -> None:
    _state["engine"] = None
    _state["audio_nets"] = {}
    _state["face_predictor"] = None


def face_state_dict():
    return _state["face"]


def precision() -> str:
    return _state["precision"]


def device() -> str:
    return _state["device"]


def video_engine():
    """Engine with VS + VD loaded (built once)."""
    from .pipeline import Engine

    if _state["engine"] is None:
        vs = _state["vs"] if _state["vs"] is not None else torch.load(PATH_STATIC, map_location="cpu")
        vd = _state["vd"] if _state["vd"] is not None else torch.load(PATH_DYNAMIC, map_location="cpu")
        _state["engine"] = Engine(vs, vd, None, precision=_state["precision"], device=_state["device"])
    return _state["engine"]


def audio_net(model_name: str, num_classes: int, root_path: str, epoch: int, device: Optional[str] = None):
    from .nets import ANet

    key = (model_name, device or _state["device"])
    if key not in _state["audio_nets"]:
        sd = _state["audio"].get(model_name, _state["audio"].get(num_classes))
        if sd is None:
            ckpt = torch.load(os.path.join(root_path, f"epoch_{epoch}.pth"), map_location="cpu")
            sd = ckpt["model_state_dict"]
        _state["audio_nets"][key] = ANet(sd, _state["precision"], device or _state["device"])
    return _state["audio_nets"][key]
