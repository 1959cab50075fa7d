"""Drop-in for the face detector and tracker the reference crops its faces with:
`data.face_detection.ibug.face_detection.RetinaFacePredictor` (retina_face/retina_face_predictor.py:18-109) and
`...utils.SimpleFaceTracker` (utils/simple_face_tracker.py:9-90), same constructor arguments, call signatures, return
types and error behaviour.  The network (RetinaFace-ResNet50), the anchors, the softmax and the box / landmark decoding
run on the GPU (nets.RetinaFaceNet over libavcer_b200); what stays on the host is what the reference does in numpy /
scipy on a handful of boxes: greedy NMS, the Hungarian assignment of the tracker.

Beyond the reference's one-image call there is `detect_batch` (frames of one size in one forward) -- the reference's
per-frame loop (data/get_face_images.py:45-61) is what `get_face_images.VideoPredictor.process` batches through it.
"""
from __future__ import annotations

import os
from copy import deepcopy
from types import SimpleNamespace
from typing import List, Optional, Sequence, Union

import numpy as np
import torch
from scipy.optimize import linear_sum_assignment

from .. import config, nets

__all__ = ["RetinaFacePredictor", "SimpleFaceTracker", "greedy_nms"]

# retina_face/config.py:21-39 (the fields the inference path reads)
cfg_re50 = {"name": "Resnet50", "min_sizes": [[16, 32], [64, 128], [256, 512]], "steps": [8, 16, 32], "variance": [0.1, 0.2],
            "clip": False, "return_layers": {"layer2": 1, "layer3": 2, "layer4": 3}, "in_channel": 256, "out_channel": 256}


def greedy_nms(dets: np.ndarray, thresh: float, top_k: int) -> List[int]:
    """Indices kept by py_cpu_nms (retina_face/py_cpu_nms.py:11-39): candidates in descending score order (numpy's default
    argsort reversed, cut to top_k first), a candidate is dropped when its IoU with an earlier kept one exceeds `thresh`
    ('+1' pixel widths, float32 arithmetic of the float32 rows).  The overlaps of a block of candidates against all later
    ones are formed in one vectorised pass (the same elementwise float32 operations as the reference's per-box pass), the
    greedy scan then only combines boolean rows."""
    order = dets[:, 4].argsort()[::-1][:top_k]
    b = dets[order]
    n = len(b)
    x1, y1, x2, y2 = b[:, 0], b[:, 1], b[:, 2], b[:, 3]
    area = (x2 - x1 + 1) * (y2 - y1 + 1)
    alive = np.ones(n, dtype=bool)
    keep: List[int] = []
    block = 256
    for lo in range(0, n, block):
        hi = min(n, lo + block)
        if not alive[lo:hi].any():
            continue
        w = np.maximum(0.0, np.minimum(x2[lo:hi, None], x2[None, lo:]) - np.maximum(x1[lo:hi, None], x1[None, lo:]) + 1)
        h = np.maximum(0.0, np.minimum(y2[lo:hi, None], y2[None, lo:]) - np.maximum(y1[lo:hi, None], y1[None, lo:]) + 1)
        inter = w * h
        with np.errstate(invalid="ignore", divide="ignore"):
            ok = inter / (area[lo:hi, None] + area[None, lo:] - inter) <= thresh      # a NaN overlap suppresses, like the reference
        for i in range(lo, hi):
            if alive[i]:
                keep.append(int(order[i]))
                alive[i + 1:] &= ok[i - lo, i + 1 - lo:]
    return keep


class RetinaFacePredictor:
    def __init__(self, threshold: float = 0.8, device: Union[str, torch.device] = "cuda:0", model: Optional[SimpleNamespace] = None,
                 config: Optional[SimpleNamespace] = None, precision: Optional[str] = None) -> None:
        """model.weights: path of the checkpoint (a state_dict, or {'state_dict': ...}, keys optionally prefixed
        'module.': retina_face_predictor.py:28-35) or the state_dict itself."""
        from .. import config as _cfg

        self.threshold = threshold
        self.device = device
        if model is None:
            model = RetinaFacePredictor.get_model()
        if config is None:
            config = RetinaFacePredictor.create_config()
        self.config = SimpleNamespace(**model.config.__dict__, **config.__dict__)
        if self.config.name != "Resnet50":
            raise ValueError("avcer_b200 covers the ResNet-50 detector (what data/get_face_images.py:28-32 instantiates)")
        sd = model.weights
        if isinstance(sd, (str, os.PathLike)):
            sd = torch.load(sd, map_location="cpu")
        if "state_dict" in sd.keys():
            sd = sd["state_dict"]
        sd = {k.split("module.", 1)[-1] if k.startswith("module.") else k: v for k, v in sd.items()}
        self.net = nets.RetinaFaceNet(sd, precision or _cfg.precision(), str(device))
        self._pinned: Optional[torch.Tensor] = None

    @staticmethod
    def get_model(name: str = "resnet50") -> SimpleNamespace:
        name = name.lower().strip()
        if name == "resnet50":
            return SimpleNamespace(weights=os.path.realpath(os.path.join(os.path.dirname(__file__), "weights", "Resnet50_Final.pth")),
                                   config=SimpleNamespace(**deepcopy(cfg_re50)))
        if name == "mobilenet0.25":
            raise ValueError("mobilenet0.25 is not covered by avcer_b200 (the path uses resnet50)")
        raise ValueError("name must be set to either resnet50 or mobilenet0.25")

    @staticmethod
    def create_config(top_k: int = 750, conf_thresh: float = 0.02, nms_thresh: float = 0.4, nms_top_k: int = 5000) -> SimpleNamespace:
        return SimpleNamespace(top_k=top_k, conf_thresh=conf_thresh, nms_thresh=nms_thresh, nms_top_k=nms_top_k)

    # ------------------------------------------------------------------------------------------------------------
    def _select(self, rows: np.ndarray) -> np.ndarray:
        """retina_face_predictor.py:86-109 on the rows of one image that can still matter.  The reference runs NMS over every
        box scoring above conf_thresh and then keeps those at or above `threshold`; a box is only ever suppressed by a
        HIGHER-scoring one, and both truncations (nms_top_k before, top_k after the suppression) cut from the low-score
        end, so the survivors at or above `threshold` are exactly the NMS of the boxes at or above it."""
        if len(rows) == 0:
            return np.empty(shape=(0, 15), dtype=np.float32)
        keep = greedy_nms(rows[:, :5], self.config.nms_thresh, self.config.nms_top_k)
        return rows[keep][:self.config.top_k]

    @torch.no_grad()
    def detect_batch(self, frames: Union[np.ndarray, torch.Tensor], rgb: bool = False) -> List[np.ndarray]:
        """frames: uint8 [n,H,W,3] (numpy, or a tensor already on the device) -> per frame the [k,15] float32 rows the
        reference's call returns (x1, y1, x2, y2, score, 5 landmark points), in its order."""
        if len(frames) == 0:
            return []
        if not isinstance(frames, torch.Tensor):
            frames = self._upload(frames)
        elif not frames.is_cuda:
            frames = frames.to(self.net.device)
        assert frames.dtype == torch.uint8 and frames.dim() == 4 and frames.shape[3] == 3
        n = frames.shape[0]
        dets = self.net.detect(frames.contiguous(), rgb)                                # [n, P, 15] on the device
        score = dets[..., 4]
        hit = torch.nonzero((score > self.config.conf_thresh) & (score >= self.threshold))   # (frame, prior), in that order
        rows = dets[hit[:, 0], hit[:, 1]].cpu().numpy()
        owner = hit[:, 0].cpu().numpy()
        return [self._select(rows[owner == i]) for i in range(n)]

    def _upload(self, frames: Union[np.ndarray, Sequence[np.ndarray]]) -> torch.Tensor:
        """Host frames (one [n,H,W,3] array or a list of [H,W,3] arrays, e.g. straight from cv2.VideoCapture.read) -> device,
        staged through a pinned buffer that is kept between calls."""
        n = len(frames)
        shape = (n,) + tuple(frames[0].shape)
        if self._pinned is None or self._pinned.numel() < int(np.prod(shape)):
            self._pinned = torch.empty(int(np.prod(shape)), dtype=torch.uint8).pin_memory()
        stage = self._pinned[:int(np.prod(shape))].view(shape)
        dst = stage.numpy()
        for i in range(n):
            dst[i] = frames[i]
        return stage.to(self.net.device, non_blocking=True)

    def __call__(self, image: np.ndarray, rgb: bool = True) -> np.ndarray:
        return self.detect_batch(image[None], rgb=rgb)[0]


class SimpleFaceTracker:
    """utils/simple_face_tracker.py:9-90: IoU + Hungarian tracking by detection.  Per call: cost[f, t] = 1 - IoU(face f,
    tracklet t) where that is at most 1 - iou_threshold (and the face is large enough), else 2 * min(#faces, #tracklets);
    scipy's linear_sum_assignment; matched tracklets take the face's box, unmatched ones are dropped immediately, unmatched
    (large enough) faces open tracklets numbered from 1.  An empty frame drops every tracklet."""

    def __init__(self, iou_threshold: float = 0.4, minimum_face_size: float = 0.0) -> None:
        self._iou_threshold = iou_threshold
        self._minimum_face_size = minimum_face_size
        self._tracklets: List[dict] = []
        self._tracklet_counter = 0

    @property
    def iou_threshold(self) -> float:
        return self._iou_threshold

    @iou_threshold.setter
    def iou_threshold(self, threshold: float) -> None:
        self._iou_threshold = threshold

    @property
    def minimum_face_size(self) -> float:
        return self._minimum_face_size

    @minimum_face_size.setter
    def minimum_face_size(self, face_size: float) -> None:
        self._minimum_face_size = face_size

    def __call__(self, face_boxes: np.ndarray) -> List[Optional[int]]:
        if face_boxes.size <= 0:
            self._tracklets = []
            return []
        n, m = face_boxes.shape[0], len(self._tracklets)
        fb = face_boxes[:, :4]
        areas = np.abs((fb[:, 2] - fb[:, 0]) * (fb[:, 3] - fb[:, 1]))
        limit = np.clip(1.0 - self._iou_threshold, 0.0, 1.0)
        big = areas >= max(self._minimum_face_size ** 2, np.finfo(float).eps)
        cost = np.full((n, m), 2.0 * min(n, m), dtype=float)
        if m:
            tb = np.stack([t["bbox"] for t in self._tracklets])
            tarea = np.asarray([t["area"] for t in self._tracklets])
            # intersections of the (corner-order independent) rectangles, in the dtype of the boxes like the reference's max / min
            left = np.maximum(np.minimum(fb[:, None, 0], fb[:, None, 2]), np.minimum(tb[None, :, 0], tb[None, :, 2]))
            top = np.maximum(np.minimum(fb[:, None, 1], fb[:, None, 3]), np.minimum(tb[None, :, 1], tb[None, :, 3]))
            right = np.minimum(np.maximum(fb[:, None, 2], fb[:, None, 0]), np.maximum(tb[None, :, 2], tb[None, :, 0]))
            bottom = np.minimum(np.maximum(fb[:, None, 3], fb[:, None, 1]), np.maximum(tb[None, :, 3], tb[None, :, 1]))
            for r in range(n):
                if not big[r]:
                    continue
                for c in range(m):
                    if right[r, c] <= left[r, c] or bottom[r, c] <= top[r, c]:
                        d = 1.0
                    else:
                        inter = (right[r, c] - left[r, c]) * (bottom[r, c] - top[r, c])
                        d = 1.0 - inter / float(areas[r] + tarea[c] - inter)
                    if d <= limit:
                        cost[r, c] = d
        ids: List[Optional[int]] = [None] * n
        matched = [False] * m
        for r, c in zip(*linear_sum_assignment(cost)):
            if cost[r, c] <= limit:
                ids[r] = self._tracklets[c]["id"]
                self._tracklets[c]["bbox"], self._tracklets[c]["area"] = fb[r].copy(), areas[r]
                matched[c] = True
        self._tracklets = [t for t, ok in zip(self._tracklets, matched) if ok]
        for r in range(n):
            if big[r] and ids[r] is None:
                self._tracklet_counter += 1
                self._tracklets.append({"bbox": fb[r].copy(), "area": areas[r], "id": self._tracklet_counter})
                ids[r] = self._tracklet_counter
        return ids

    def reset(self, reset_tracklet_counter: bool = True) -> None:
        self._tracklets = []
        if reset_tracklet_counter:
            self._tracklet_counter = 0
