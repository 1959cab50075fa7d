"""Drop-in for the reference's face-crop extraction (src/data/get_face_images.py:10-63): `VideoPredictor.process(path,
save_path)` reads a video, detects the faces of every frame, tracks them and writes the crop of track t in frame i to
`<save_path>/<video name without extension>/<t:02d>/<i:06d>.jpg` -- the directory get_prob_video reads.

The reference calls the detector once per frame; here `batch` frames of the video go through one forward of the GPU
detector (face_detection.RetinaFacePredictor.detect_batch), after which tracker and crops run per frame in order, so the
track ids and files are the ones the per-frame loop produces.
"""
from __future__ import annotations

import os

import cv2

from .. import config
from .face_detection import RetinaFacePredictor, SimpleFaceTracker


class VideoPredictor:
    def __init__(self, batch: int = 8, model=None, precision=None):
        super().__init__()
        self.video_stream = None
        self.device = config.device()
        self.model = None
        self.count_frame = None
        self.batch = int(batch)
        self._model_spec, self._precision = model, precision
        self.init_predictor()

    def init_path(self, path):
        self.video_stream = cv2.VideoCapture(path)
        self.w = int(self.video_stream.get(cv2.CAP_PROP_FRAME_WIDTH))
        self.h = int(self.video_stream.get(cv2.CAP_PROP_FRAME_HEIGHT))
        self.fps = int(self.video_stream.get(cv2.CAP_PROP_FPS))
        self.total_frames = int(self.video_stream.get(cv2.CAP_PROP_FRAME_COUNT))

    def init_predictor(self):
        spec = self._model_spec or RetinaFacePredictor.get_model("resnet50")
        if self._model_spec is None and config.face_state_dict() is not None:       # injected weights (config.set_state_dicts)
            spec.weights = config.face_state_dict()
        self.model = RetinaFacePredictor(threshold=0.8, device=self.device, model=spec, precision=self._precision)
        self.face_tracker = SimpleFaceTracker(iou_threshold=0.4, minimum_face_size=0.0)

    def __del__(self):
        if getattr(self, "video_stream", None) is not None:
            self.video_stream.release()

    def _emit(self, fr, dets, save_path, name_file):
        """get_face_images.py:50-61 for one frame."""
        n_img = str(self.count_frame).zfill(6)
        tids = self.face_tracker(dets)
        for pred, tid in zip(dets, tids):
            startX, startY, endX, endY = pred[:4].astype(int)
            startX, startY = max(0, startX), max(0, startY)
            endX, endY = min(self.w - 1, endX), min(self.h - 1, endY)
            c_path = os.path.join(save_path, name_file[:-4], str(tid - 1).zfill(2))
            os.makedirs(c_path, exist_ok=True)
            cv2.imwrite(os.path.join(c_path, n_img + ".jpg"), fr[startY:endY, startX:endX])
        self.count_frame += 1

    def process(self, path, save_path):
        self.count_frame = 0
        self.init_path(path)
        name_file = os.path.basename(path)
        done = False
        while not done:
            frames = []
            while len(frames) < self.batch:
                ret, fr = self.video_stream.read()
                if not ret:
                    done = True
                    break
                frames.append(fr)
            if frames:
                for fr, dets in zip(frames, self.model.detect_batch(frames, rgb=False)):
                    self._emit(fr, dets, save_path, name_file)
        self.face_tracker.reset()
