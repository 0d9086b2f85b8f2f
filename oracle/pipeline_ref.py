"""ORACLE (test infrastructure, never shipped, never on the product path).

CPU restatement of the pre/post-processing the reference runs inside ``ultralytics.YOLO.predict``
(entered from /root/reference/core/model.py:133; consumers of the result layout:
/root/reference/utils/visualization.py:52-68, /root/reference/demos/detection_demo.py:96-132).
Follows SURVEY.md Appendix B: LetterBox + BasePredictor.preprocess (upstream data/augment.py,
engine/predictor.py), non_max_suppression / xywh2xyxy / scale_boxes / clip_boxes (upstream
utils/ops.py), LoadTensor quirks (upstream data/loaders.py).  The two real pieces of the reference
stack that ARE importable here are used directly so the oracle is anchored on them:
``cv2.resize / cv2.copyMakeBorder`` (the reference's letterbox arithmetic) and
``torchvision.ops.nms`` (the reference's NMS).

"parity unpinned" against ultralytics itself (not installable here); pinned by the probes in
SURVEY.md §8c (letterbox shapes / pads, NMS tie and threshold semantics) - tests/test_oracle_kat.py.
"""
from __future__ import annotations

import time
from typing import List, Optional, Sequence, Tuple

import cv2
import numpy as np
import torch
import torchvision


# --------------------------------------------------------------------------- letterbox
def letterbox_params(h0: int, w0: int, new_shape: Tuple[int, int] = (640, 640), auto: bool = False,
                     stride: int = 32, scaleup: bool = True):
    """Geometry of upstream LetterBox.__call__ (center=True, scale_fill=False).

    Returns (new_unpad_w, new_unpad_h, top, bottom, left, right, H, W).
    """
    r = min(new_shape[0] / h0, new_shape[1] / w0)
    if not scaleup:
        r = min(r, 1.0)
    new_w, new_h = int(round(w0 * r)), int(round(h0 * r))
    dw, dh = new_shape[1] - new_w, new_shape[0] - new_h
    if auto:
        dw, dh = np.mod(dw, stride), np.mod(dh, stride)
    dw /= 2
    dh /= 2
    top, bottom = int(round(dh - 0.1)), int(round(dh + 0.1))
    left, right = int(round(dw - 0.1)), int(round(dw + 0.1))
    return new_w, new_h, top, bottom, left, right, new_h + top + bottom, new_w + left + right


def letterbox(img: np.ndarray, new_shape=(640, 640), auto=False, stride=32) -> np.ndarray:
    h0, w0 = img.shape[:2]
    new_w, new_h, top, bottom, left, right, _, _ = letterbox_params(h0, w0, new_shape, auto, stride)
    if (w0, h0) != (new_w, new_h):
        img = cv2.resize(img, (new_w, new_h), interpolation=cv2.INTER_LINEAR)
    return cv2.copyMakeBorder(img, top, bottom, left, right, cv2.BORDER_CONSTANT, value=(114, 114, 114))


def preprocess(imgs: Sequence[np.ndarray], imgsz=(640, 640), rect: bool = True, stride: int = 32) -> torch.Tensor:
    """BGR uint8 HWC list -> float32 [B,3,H,W] RGB in [0,1]  (upstream BasePredictor.preprocess)."""
    same = len({im.shape for im in imgs}) == 1
    auto = bool(rect and same)
    lb = [letterbox(im, imgsz, auto=auto, stride=stride) for im in imgs]
    x = np.stack(lb)
    x = x[..., ::-1].transpose((0, 3, 1, 2))
    x = np.ascontiguousarray(x)
    return torch.from_numpy(x).float() / 255


def preprocess_tensor(x: torch.Tensor) -> torch.Tensor:
    """LoadTensor: 4-D, H,W % 32 == 0, divide by 255 when max > 1 (upstream data/loaders.py)."""
    assert x.ndim == 4 and x.shape[2] % 32 == 0 and x.shape[3] % 32 == 0
    x = x.float()
    if x.max() > 1.0 + torch.finfo(x.dtype).eps:
        x = x / 255.0
    return x


# --------------------------------------------------------------------------- post-processing
def xywh2xyxy(x: torch.Tensor) -> torch.Tensor:
    y = torch.empty_like(x)
    xy = x[..., :2]
    wh = x[..., 2:] / 2
    y[..., :2] = xy - wh
    y[..., 2:] = xy + wh
    return y


def candidates(pred_i: torch.Tensor, conf_thres: float, multi_label: bool, classes=None) -> torch.Tensor:
    """One image of `prediction.transpose(-1,-2)` ([A, 4+nc], xyxy boxes) -> [K,6] in anchor order."""
    nc = pred_i.shape[1] - 4
    xc = pred_i[:, 4:].amax(1) > conf_thres
    x = pred_i[xc]
    box, cls = x[:, :4], x[:, 4:]
    if multi_label and nc > 1:
        i, j = torch.where(cls > conf_thres)
        x = torch.cat((box[i], x[i, 4 + j, None], j[:, None].float()), 1)
    else:
        conf, j = cls.max(1, keepdim=True)
        x = torch.cat((box, conf, j.float()), 1)[conf.view(-1) > conf_thres]
    if classes is not None:
        x = x[(x[:, 5:6] == torch.tensor(classes, device=x.device)).any(1)]
    return x


def nms_sorted_stable(boxes: torch.Tensor, scores: torch.Tensor, iou_thres: float) -> torch.Tensor:
    """torchvision.ops.nms IS the reference's NMS."""
    return torchvision.ops.nms(boxes, scores, iou_thres)


def non_max_suppression(prediction: torch.Tensor, conf_thres=0.25, iou_thres=0.45, classes=None, agnostic=False,
                        multi_label=False, max_det=300, max_nms=30000, max_wh=7680) -> List[torch.Tensor]:
    """upstream utils/ops.py:non_max_suppression for nc-class detect output [B, 4+nc, A].

    Deviations, both documented in SURVEY Appendix B.2: the wall-clock time limit is disabled, and the
    `n > max_nms` truncation uses a STABLE descending sort (upstream argsort is unstable; stable is
    the canonical definition both oracle and CUDA path implement).
    """
    nc = prediction.shape[1] - 4
    multi_label &= nc > 1
    prediction = prediction.transpose(-1, -2)
    prediction = torch.cat((xywh2xyxy(prediction[..., :4]), prediction[..., 4:]), dim=-1)
    output = [torch.zeros((0, 6))] * prediction.shape[0]
    for xi, x in enumerate(prediction):
        x = candidates(x, conf_thres, multi_label, classes)
        n = x.shape[0]
        if not n:
            continue
        if n > max_nms:
            x = x[x[:, 4].argsort(descending=True, stable=True)[:max_nms]]
        c = x[:, 5:6] * (0 if agnostic else max_wh)
        scores = x[:, 4]
        boxes = x[:, :4] + c
        i = nms_sorted_stable(boxes, scores, iou_thres)
        i = i[:max_det]
        output[xi] = x[i]
    return output


def scale_boxes_params(img1_shape, img0_shape):
    gain = min(img1_shape[0] / img0_shape[0], img1_shape[1] / img0_shape[1])
    pad_x = round((img1_shape[1] - img0_shape[1] * gain) / 2 - 0.1)
    pad_y = round((img1_shape[0] - img0_shape[0] * gain) / 2 - 0.1)
    return gain, pad_x, pad_y


def scale_boxes(img1_shape, boxes: torch.Tensor, img0_shape) -> torch.Tensor:
    """upstream utils/ops.py:scale_boxes(padding=True, xywh=False) + clip_boxes; in place on [n,4+]."""
    gain, pad_x, pad_y = scale_boxes_params(img1_shape, img0_shape)
    boxes[..., 0] -= pad_x
    boxes[..., 1] -= pad_y
    boxes[..., 2] -= pad_x
    boxes[..., 3] -= pad_y
    boxes[..., :4] /= gain
    boxes[..., 0].clamp_(0, img0_shape[1])
    boxes[..., 1].clamp_(0, img0_shape[0])
    boxes[..., 2].clamp_(0, img0_shape[1])
    boxes[..., 3].clamp_(0, img0_shape[0])
    return boxes


# --------------------------------------------------------------------------- end to end
@torch.no_grad()
def predict(model, source, conf=0.25, iou=0.7, max_det=300, imgsz=640, rect=True, agnostic_nms=False,
            classes=None, multi_label=False, return_raw=False):
    """Reference `predict` for source in {BGR ndarray, list of ndarrays, float tensor [B,3,H,W]}.

    Returns a list of float32 [n,6] tensors (x1,y1,x2,y2 in original-image pixels, conf, cls).
    `model` is an oracle DetectionModel (fused or not).
    """
    new_shape = (imgsz, imgsz) if isinstance(imgsz, int) else tuple(imgsz)
    if isinstance(source, torch.Tensor):
        x = preprocess_tensor(source)
        orig_shapes = [tuple(x.shape[2:])] * x.shape[0]
    else:
        imgs = [source] if isinstance(source, np.ndarray) else list(source)
        x = preprocess(imgs, new_shape, rect=rect)
        orig_shapes = [im.shape[:2] for im in imgs]
    y, feats = model(x)
    dets = non_max_suppression(y, conf, iou, classes, agnostic_nms, multi_label=multi_label, max_det=max_det)
    out = []
    for d, s0 in zip(dets, orig_shapes):
        d = d.clone()
        if d.shape[0]:
            scale_boxes(x.shape[2:], d[:, :4], s0)
        out.append(d)
    if return_raw:
        return out, x, y, feats
    return out


def time_predict(model, source, warmup=3, iters=10, **kw):
    for _ in range(warmup):
        predict(model, source, **kw)
    ts = []
    for _ in range(iters):
        t0 = time.perf_counter()
        predict(model, source, **kw)
        ts.append(time.perf_counter() - t0)
    return ts
