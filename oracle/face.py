"""ORACLE (test infrastructure, never shipped on the product path).

CPU restatement (numpy + torch.nn.functional in fp32) of the face detection / tracking step that produces the face
crops of ElenaRyumina/AVCER:
  * RetinaFace-ResNet50 forward    -- src/data/face_detection/ibug/face_detection/retina_face/retina_face.py:48-115,
                                      retina_face_net.py:43-100 (SSH, FPN) over torchvision's ResNet-50 (v1.5: the stride
                                      of a bottleneck sits on its 3x3 conv; BatchNorm eps 1e-5; max-pool 3x3/2 pad 1)
  * anchors                        -- retina_face/prior_box.py:6-33
  * box / landmark decoding        -- retina_face/box_utils.py:210-249
  * greedy NMS                     -- retina_face/py_cpu_nms.py:11-39
  * the predictor call             -- retina_face/retina_face_predictor.py:60-109
  * IoU + Hungarian tracker        -- utils/simple_face_tracker.py:9-90
  * frame loop / crop boxes        -- src/data/get_face_images.py:38-63

Pinned against the reference's own classes by oracle/make_golden.py -> tests/golden/face.npz.
"""
from __future__ import annotations

from math import ceil
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F
from scipy.optimize import linear_sum_assignment

BN_EPS = 1e-5                              # torch.nn.BatchNorm2d default (torchvision resnet50, conv_bn helpers)
MEAN_BGR = (104, 117, 123)                 # retina_face_predictor.py:65
MIN_SIZES = ((16, 32), (64, 128), (256, 512))   # config.py:22 (cfg_re50)
STEPS = (8, 16, 32)                        # config.py:23
VARIANCE = (0.1, 0.2)                      # config.py:24
BLOCKS = (3, 4, 6, 3)


# ------------------------------------------------------------------------------------------ network
def _bn(x, sd, p):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                        training=False, eps=BN_EPS)


def _conv_bn(x, sd, p, stride=1, pad=1, relu=True):
    """conv_bn / conv_bn_no_relu / conv_bn1X1 of retina_face_net.py:6-27 (LeakyReLU(0) == ReLU for out_channel 256)."""
    y = _bn(F.conv2d(x, sd[p + ".0.weight"], None, stride=stride, padding=pad), sd, p + ".1")
    return F.relu(y) if relu else y


def _bottleneck(x, sd, p, stride, has_ds):
    y = F.relu(_bn(F.conv2d(x, sd[p + ".conv1.weight"]), sd, p + ".bn1"))
    y = F.relu(_bn(F.conv2d(y, sd[p + ".conv2.weight"], stride=stride, padding=1), sd, p + ".bn2"))
    y = _bn(F.conv2d(y, sd[p + ".conv3.weight"]), sd, p + ".bn3")
    idn = _bn(F.conv2d(x, sd[p + ".downsample.0.weight"], stride=stride), sd, p + ".downsample.1") if has_ds else x
    return F.relu(y + idn)


def body(x: torch.Tensor, sd: Dict[str, torch.Tensor], taps: Optional[dict] = None) -> List[torch.Tensor]:
    """torchvision resnet50 up to layer4; returns the layer2 / layer3 / layer4 maps (config.py:34 return_layers)."""
    y = F.relu(_bn(F.conv2d(x, sd["body.conv1.weight"], stride=2, padding=3), sd, "body.bn1"))
    if taps is not None:
        taps["stem"] = y
    y = F.max_pool2d(y, 3, 2, 1)
    if taps is not None:
        taps["pool"] = y
    outs = []
    for li, blocks in enumerate(BLOCKS, start=1):
        for b in range(blocks):
            y = _bottleneck(y, sd, f"body.layer{li}.{b}", 2 if (b == 0 and li > 1) else 1, b == 0)
        if taps is not None:
            taps[f"layer{li}"] = y
        if li >= 2:
            outs.append(y)
    return outs


def fpn(feats: Sequence[torch.Tensor], sd) -> List[torch.Tensor]:
    """retina_face_net.py:65-100."""
    o1 = _conv_bn(feats[0], sd, "fpn.output1", pad=0)
    o2 = _conv_bn(feats[1], sd, "fpn.output2", pad=0)
    o3 = _conv_bn(feats[2], sd, "fpn.output3", pad=0)
    o2 = _conv_bn(o2 + F.interpolate(o3, size=o2.shape[2:], mode="nearest"), sd, "fpn.merge2")
    o1 = _conv_bn(o1 + F.interpolate(o2, size=o1.shape[2:], mode="nearest"), sd, "fpn.merge1")
    return [o1, o2, o3]


def ssh(x, sd, p):
    """retina_face_net.py:43-62."""
    c3 = _conv_bn(x, sd, p + ".conv3X3", relu=False)
    c5_1 = _conv_bn(x, sd, p + ".conv5X5_1")
    c5 = _conv_bn(c5_1, sd, p + ".conv5X5_2", relu=False)
    c7_2 = _conv_bn(c5_1, sd, p + ".conv7X7_2")
    c7 = _conv_bn(c7_2, sd, p + ".conv7x7_3", relu=False)
    return F.relu(torch.cat([c3, c5, c7], dim=1))


def _head(x, sd, p, width):
    y = F.conv2d(x, sd[p + ".conv1x1.weight"], sd[p + ".conv1x1.bias"])
    return y.permute(0, 2, 3, 1).contiguous().view(y.shape[0], -1, width)


def forward(x: torch.Tensor, sd: Dict[str, torch.Tensor], taps: Optional[dict] = None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """RetinaFace.forward in phase 'test' (retina_face.py:93-115): (loc [n,P,4], softmax(conf) [n,P,2], landms [n,P,10]);
    priors are ordered level-major, then row, column, anchor."""
    with torch.no_grad():
        pyramid = fpn(body(x, sd, taps), sd)
        feats = [ssh(pyramid[i], sd, f"ssh{i + 1}") for i in range(3)]
        if taps is not None:
            for i in range(3):
                taps[f"fpn{i + 1}"], taps[f"ssh{i + 1}"] = pyramid[i], feats[i]
        loc = torch.cat([_head(f, sd, f"BboxHead.{i}", 4) for i, f in enumerate(feats)], dim=1)
        cls = torch.cat([_head(f, sd, f"ClassHead.{i}", 2) for i, f in enumerate(feats)], dim=1)
        lmk = torch.cat([_head(f, sd, f"LandmarkHead.{i}", 10) for i, f in enumerate(feats)], dim=1)
        if taps is not None:
            taps["cls_logits"] = cls
        return loc, F.softmax(cls, dim=-1), lmk


# ------------------------------------------------------------------------------------------ post-processing
def prior_box(height: int, width: int) -> torch.Tensor:
    """prior_box.py:17-33: [P, 4] (cx, cy, w, h) in image-relative units, fp32 (built from Python doubles)."""
    anchors = []
    for k, step in enumerate(STEPS):
        fh, fw = ceil(height / step), ceil(width / step)
        for i in range(fh):
            for j in range(fw):
                for m in MIN_SIZES[k]:
                    anchors += [(j + 0.5) * step / width, (i + 0.5) * step / height, m / width, m / height]
    return torch.tensor(anchors, dtype=torch.float32).view(-1, 4)


def decode(loc: torch.Tensor, priors: torch.Tensor) -> torch.Tensor:
    """box_utils.py:210-228 (fp32 torch arithmetic, in this order)."""
    boxes = torch.cat((priors[:, :2] + loc[:, :2] * VARIANCE[0] * priors[:, 2:],
                       priors[:, 2:] * torch.exp(loc[:, 2:] * VARIANCE[1])), 1)
    boxes[:, :2] -= boxes[:, 2:] / 2
    boxes[:, 2:] += boxes[:, :2]
    return boxes


def decode_landm(pre: torch.Tensor, priors: torch.Tensor) -> torch.Tensor:
    """box_utils.py:231-249."""
    return torch.cat([priors[:, :2] + pre[:, 2 * i:2 * i + 2] * VARIANCE[0] * priors[:, 2:] for i in range(5)], dim=1)


def nms(dets: np.ndarray, thresh: float, top_k: int) -> List[int]:
    """py_cpu_nms.py:11-39: greedy, descending score (numpy's default argsort, reversed, truncated to top_k BEFORE the
    suppression loop), '+1' pixel areas, suppression when IoU > thresh."""
    x1, y1, x2, y2, scores = dets[:, 0], dets[:, 1], dets[:, 2], dets[:, 3], dets[:, 4]
    areas = (x2 - x1 + 1) * (y2 - y1 + 1)
    order = scores.argsort()[: -top_k - 1: -1]
    keep = []
    while order.size > 0:
        i = order[0]
        keep.append(int(i))
        rest = order[1:]
        w = np.maximum(0.0, np.minimum(x2[i], x2[rest]) - np.maximum(x1[i], x1[rest]) + 1)
        h = np.maximum(0.0, np.minimum(y2[i], y2[rest]) - np.maximum(y1[i], y1[rest]) + 1)
        inter = w * h
        ovr = inter / (areas[i] + areas[rest] - inter)
        order = rest[np.where(ovr <= thresh)[0]]
    return keep


def postprocess(loc: torch.Tensor, conf: torch.Tensor, landms: torch.Tensor, height: int, width: int, threshold: float = 0.8,
                conf_thresh: float = 0.02, nms_thresh: float = 0.4, nms_top_k: int = 5000, top_k: int = 750) -> np.ndarray:
    """retina_face_predictor.py:69-109 for one image: loc [P,4], conf [P,2] (softmaxed), landms [P,10] -> dets [k,15] f32."""
    priors = prior_box(height, width)
    boxes = (decode(loc, priors) * torch.tensor([width, height, width, height], dtype=torch.float32)).numpy()
    scores = conf.numpy()[:, 1]
    lm = (decode_landm(landms, priors) * torch.tensor([width, height] * 5, dtype=torch.float32)).numpy()
    inds = np.where(scores > conf_thresh)[0]
    if len(inds) == 0:
        return np.empty((0, 15), dtype=np.float32)
    boxes, lm, scores = boxes[inds], lm[inds], scores[inds]
    dets = np.hstack((boxes, scores[:, np.newaxis])).astype(np.float32, copy=False)
    keep = nms(dets, nms_thresh, nms_top_k)
    dets = np.concatenate((dets[keep, :][:top_k], lm[keep][:top_k]), axis=1)
    inds = np.where(dets[:, 4] >= threshold)[0]
    return dets[inds] if len(inds) else np.empty((0, 15), dtype=np.float32)


def prepare(image: np.ndarray, rgb: bool = False) -> torch.Tensor:
    """retina_face_predictor.py:61-67: uint8 HxWx3 -> fp32 [1,3,H,W] in BGR order minus (104, 117, 123)."""
    if rgb:
        image = image[..., ::-1]
    x = image.astype(int) - np.array(MEAN_BGR)
    return torch.from_numpy(np.ascontiguousarray(x.transpose(2, 0, 1))).unsqueeze(0).float()


def predict(sd, image: np.ndarray, rgb: bool = False, threshold: float = 0.8) -> np.ndarray:
    """RetinaFacePredictor.__call__ (retina_face_predictor.py:60-109)."""
    loc, conf, landms = forward(prepare(image, rgb), sd)
    return postprocess(loc[0], conf[0], landms[0], image.shape[0], image.shape[1], threshold)


# ------------------------------------------------------------------------------------------ tracker
class SimpleFaceTracker:
    """utils/simple_face_tracker.py:9-90: per frame, a (faces x tracklets) matrix of 1 - IoU (entries above
    1 - iou_threshold replaced by 2 * min(faces, tracklets)), Hungarian assignment, unmatched tracklets dropped at once,
    unmatched faces start tracklets with the next id (ids start at 1).  A frame without faces clears all tracklets."""

    def __init__(self, iou_threshold: float = 0.4, minimum_face_size: float = 0.0):
        self.iou_threshold, self.minimum_face_size = iou_threshold, minimum_face_size
        self.tracklets, self.counter = [], 0

    def __call__(self, face_boxes: np.ndarray) -> List[Optional[int]]:
        if face_boxes.size <= 0:
            self.tracklets = []
            return []
        areas = np.abs((face_boxes[:, 2] - face_boxes[:, 0]) * (face_boxes[:, 3] - face_boxes[:, 1]))
        for t in self.tracklets:
            t["tracked"] = False
        limit = np.clip(1.0 - self.iou_threshold, 0.0, 1.0)
        min_area = max(self.minimum_face_size ** 2, np.finfo(float).eps)
        n, m = face_boxes.shape[0], len(self.tracklets)
        dist = np.full((n, m), 2.0 * min(n, m), dtype=float)
        for r, fb in enumerate(face_boxes):
            if areas[r] < min_area:
                continue
            for c, t in enumerate(self.tracklets):
                tb = t["bbox"]
                xl, yt = max(min(fb[0], fb[2]), min(tb[0], tb[2])), max(min(fb[1], fb[3]), min(tb[1], tb[3]))
                xr, yb = min(max(fb[2], fb[0]), max(tb[2], tb[0])), min(max(fb[3], fb[1]), max(tb[3], tb[1]))
                if xr <= xl or yb <= yt:
                    d = 1.0
                else:
                    inter = (xr - xl) * (yb - yt)
                    d = 1.0 - inter / float(areas[r] + t["area"] - inter)
                if d <= limit:
                    dist[r, c] = d
        ids: List[Optional[int]] = [None] * n
        for r, c in zip(*linear_sum_assignment(dist)):
            if dist[r, c] <= limit:
                ids[r] = self.tracklets[c]["id"]
                self.tracklets[c].update(bbox=face_boxes[r, :4].copy(), area=areas[r], tracked=True)
        self.tracklets = [t for t in self.tracklets if t["tracked"]]
        for r, fb in enumerate(face_boxes):
            if areas[r] >= min_area and ids[r] is None:
                self.counter += 1
                self.tracklets.append({"bbox": fb[:4].copy(), "area": areas[r], "id": self.counter, "tracked": True})
                ids[r] = self.counter
        return ids

    def reset(self, reset_tracklet_counter: bool = True):
        self.tracklets = []
        if reset_tracklet_counter:
            self.counter = 0


def crop_boxes(dets: np.ndarray, ids: Sequence[Optional[int]], width: int, height: int) -> List[Tuple[int, int, int, int, int]]:
    """get_face_images.py:53-60: (track dir index tid-1, startX, startY, endX, endY) of every detection of one frame;
    coordinates truncated towards zero, clamped to [0, w-1] / [0, h-1]; the crop is frame[startY:endY, startX:endX]."""
    out = []
    for det, tid in zip(dets, ids):
        sx, sy, ex, ey = det[:4].astype(int)
        out.append((tid - 1, max(0, sx), max(0, sy), min(width - 1, ex), min(height - 1, ey)))
    return out
