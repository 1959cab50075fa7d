"""Drop-in for the reference's face-crop extraction (src/data/get_face_images.py:10-63): `VideoPredictor.process(path,
save_path)` reads a video, detects the faces of every frame, tracks them and writes the crop of track t in frame i to
`<save_path>/<video name without extension>/<t:02d>/<i:06d>.jpg` -- the directory get_prob_video reads.

The reference calls the detector once per frame; here `batch` frames of the video go through one forward of the GPU
detector (face_detection.RetinaFacePredictor.detect_batch), after which tracker and crops run per frame in order, so the
track ids and files are the ones the per-frame loop produces.  Attribute names (`video_stream`, `w`, `h`, `fps`,
`total_frames`, `count_frame`, `model`, `face_tracker`) are the reference's: run.py reads `fps` and `total_frames`.
"""
from __future__ import annotations

import os

import cv2
import numpy as np

from .. import config
from .face_detection import RetinaFacePredictor, SimpleFaceTracker

_STREAM_PROPS = (("w", cv2.CAP_PROP_FRAME_WIDTH), ("h", cv2.CAP_PROP_FRAME_HEIGHT), ("fps", cv2.CAP_PROP_FPS),
                 ("total_frames", cv2.CAP_PROP_FRAME_COUNT))


class VideoPredictor:
    def __init__(self, batch: int = 8, model=None, precision=None):
        self.device = config.device()
        self.batch = max(1, int(batch))
        self.video_stream = self.model = self.count_frame = None
        self._model_spec, self._precision = model, precision
        self.init_predictor()

    def init_predictor(self):
        """Detector (threshold 0.8, ResNet-50) and tracker (IoU 0.4, no minimum size): get_face_images.py:26-32."""
        spec = self._model_spec or RetinaFacePredictor.get_model("resnet50")
        if self._model_spec is None and config.face_state_dict() is not None:       # injected weights (config.set_state_dicts)
            spec.weights = config.face_state_dict()
        self.model = RetinaFacePredictor(threshold=0.8, device=self.device, model=spec, precision=self._precision)
        self.face_tracker = SimpleFaceTracker(iou_threshold=0.4, minimum_face_size=0.0)

    def init_path(self, path):
        """Opens the video; width, height, fps and frame count are the container's values truncated to int (:19-24)."""
        self.video_stream = cv2.VideoCapture(path)
        for name, prop in _STREAM_PROPS:
            setattr(self, name, int(self.video_stream.get(prop)))

    def __del__(self):
        stream = getattr(self, "video_stream", None)
        if stream is not None:
            stream.release()

    def _write_crops(self, frame: np.ndarray, dets: np.ndarray, clip_dir: str) -> None:
        """One frame of the reference's loop (:48-61): track ids for its detections, then per detection the box truncated to
        integers, clamped to [0, w-1] x [0, h-1], cut out of the frame and written as <track - 1:02d>/<frame:06d>.jpg."""
        ids = self.face_tracker(dets)
        if len(ids):
            box = dets[:, :4].astype(int)
            lo = np.maximum(box[:, :2], 0)
            hi = np.minimum(box[:, 2:], [self.w - 1, self.h - 1])
            for (x0, y0), (x1, y1), tid in zip(lo, hi, ids):
                track_dir = os.path.join(clip_dir, f"{tid - 1:02d}")
                os.makedirs(track_dir, exist_ok=True)
                cv2.imwrite(os.path.join(track_dir, f"{self.count_frame:06d}.jpg"), frame[y0:y1, x0:x1])
        self.count_frame += 1

    def process(self, path, save_path):
        self.count_frame = 0
        self.init_path(path)
        clip_dir = os.path.join(save_path, os.path.basename(path)[:-4])
        pending = []
        more = True
        while more:
            more, frame = self.video_stream.read()
            if more:
                pending.append(frame)
            if pending and (len(pending) == self.batch or not more):
                for fr, dets in zip(pending, self.model.detect_batch(pending, rgb=False)):
                    self._write_crops(fr, dets, clip_dir)
                pending = []
        self.face_tracker.reset()
