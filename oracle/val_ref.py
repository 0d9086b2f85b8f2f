"""ORACLE (test infrastructure, never shipped, never on the product path): the detect path of `model.val` restated.

What the reference runs through `YOLO11Model.val` (/root/reference/core/model.py:180-195, called from
/root/reference/core/validator.py:121-141) is ultralytics' DetectionValidator over a RECT dataloader.  Restated here, on the
CPU, with cv2 for the image arithmetic (it is the reference's) - "parity unpinned" against ultralytics itself like the rest of
oracle/ (ultralytics is not installable here); the pieces follow upstream
  data/dataset.py  YOLODataset.set_rectangle        : sort by aspect ratio, per-batch canvas ceil(shape * imgsz / 32 + 0.5) * 32
  data/base.py     BaseDataset.load_image(rect_mode) : long side -> imgsz, (min(ceil(w0 r), imgsz), min(ceil(h0 r), imgsz)), INTER_LINEAR
  data/augment.py  LetterBox(auto=False, scaleup=False, center=True), border 114
  models/yolo/detect/val.py                          : /255, NMS(conf 0.001, iou 0.6, multi_label), scale_boxes(ratio_pad)
The metric maths (match_predictions, ap_per_class) is restated independently of the product in `mean_ap`.
"""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import cv2
import numpy as np
import torch

from . import pipeline_ref as P


def set_rectangle(shapes: Sequence[Tuple[int, int]], batch: int, imgsz: int = 640, stride: int = 32, pad: float = 0.5):
    s = np.array(shapes, dtype=np.float64)
    ar = s[:, 0] / s[:, 1]
    irect = ar.argsort(kind="stable")
    ar = ar[irect]
    bi = np.floor(np.arange(len(ar)) / batch).astype(int)
    nb = bi[-1] + 1
    shp = [[1, 1]] * nb
    for i in range(nb):
        ari = ar[bi == i]
        mini, maxi = ari.min(), ari.max()
        if maxi < 1:
            shp[i] = [maxi, 1]
        elif mini > 1:
            shp[i] = [1, 1 / mini]
    batch_shapes = np.ceil(np.array(shp) * imgsz / stride + pad).astype(int) * stride
    return irect.tolist(), bi.tolist(), batch_shapes.tolist()


def load_image(im: np.ndarray, imgsz: int) -> np.ndarray:
    h0, w0 = im.shape[:2]
    r = imgsz / max(h0, w0)
    if r != 1:
        w, h = (min(math.ceil(w0 * r), imgsz), min(math.ceil(h0 * r), imgsz))
        im = cv2.resize(im, (w, h), interpolation=cv2.INTER_LINEAR)
    return im


def letterbox_val(im: np.ndarray, new_shape: Tuple[int, int]):
    shape = im.shape[:2]
    r = min(min(new_shape[0] / shape[0], new_shape[1] / shape[1]), 1.0)
    new_unpad = int(round(shape[1] * r)), int(round(shape[0] * r))
    dw, dh = (new_shape[1] - new_unpad[0]) / 2, (new_shape[0] - new_unpad[1]) / 2
    if shape[::-1] != new_unpad:
        im = cv2.resize(im, new_unpad, interpolation=cv2.INTER_LINEAR)
    top, bottom = int(round(dh - 0.1)), int(round(dh + 0.1))
    left, right = int(round(dw - 0.1)), int(round(dw + 0.1))
    return cv2.copyMakeBorder(im, top, bottom, left, right, cv2.BORDER_CONSTANT, value=(114, 114, 114)), (left, top)


@torch.no_grad()
def val_predictions(model, images: Sequence[np.ndarray], batch: int = 16, imgsz: int = 640, conf: float = 0.001, iou: float = 0.6,
                    max_det: int = 300) -> List[np.ndarray]:
    """-> per image (in INPUT order) float [n,6] x1,y1,x2,y2 (original pixels), conf, cls."""
    irect, bi, batch_shapes = set_rectangle([im.shape[:2] for im in images], batch, imgsz)
    out: List[np.ndarray] = [None] * len(images)
    for b in range(bi[-1] + 1):
        idx = [irect[k] for k in range(len(irect)) if bi[k] == b]
        canvas = tuple(batch_shapes[b])
        xs, metas = [], []
        for i in idx:
            h0, w0 = images[i].shape[:2]
            im = load_image(images[i], imgsz)
            ratio = (im.shape[0] / h0, im.shape[1] / w0)
            lb, pad = letterbox_val(im, canvas)
            xs.append(lb)
            metas.append((ratio, pad, (h0, w0)))
        x = np.ascontiguousarray(np.stack(xs)[..., ::-1].transpose(0, 3, 1, 2))
        y, _ = model(torch.from_numpy(x).float() / 255)
        dets = P.non_max_suppression(y, conf, iou, multi_label=True, max_det=max_det)
        for i, d, (ratio, pad, (h0, w0)) in zip(idx, dets, metas):
            d = d.clone()
            if len(d):        # scale_boxes(img1_shape, boxes, img0_shape, ratio_pad): gain = ratio_pad[0][0], pad = ratio_pad[1]
                d[:, [0, 2]] -= pad[0]
                d[:, [1, 3]] -= pad[1]
                d[:, :4] /= ratio[0]
                d[:, [0, 2]] = d[:, [0, 2]].clamp(0, w0)
                d[:, [1, 3]] = d[:, [1, 3]].clamp(0, h0)
            out[i] = d.numpy()
    return out


def mean_ap(preds: Sequence[np.ndarray], gts: Sequence[np.ndarray]) -> Tuple[float, float]:
    """(mAP50, mAP50-95) by upstream utils/metrics.py: greedy one-to-one matching per IoU threshold in decreasing IoU, 101-point
    interpolated AP of the precision envelope, mean over the classes present in the ground truth.  gts[i]: [m,5] cls,x1,y1,x2,y2."""
    iouv = np.linspace(0.5, 0.95, 10)
    tp, conf, pcls, tcls = [], [], [], []
    for pr, gt in zip(preds, gts):
        tcls.extend(gt[:, 0].tolist())
        if not len(pr):
            continue
        correct = np.zeros((len(pr), 10), bool)
        if len(gt):
            a, b = gt[:, None, 1:], pr[None, :, :4]
            inter = np.clip(np.minimum(a[..., 2:], b[..., 2:]) - np.maximum(a[..., :2], b[..., :2]), 0, None).prod(-1)
            union = (a[..., 2] - a[..., 0]) * (a[..., 3] - a[..., 1]) + (b[..., 2] - b[..., 0]) * (b[..., 3] - b[..., 1]) - inter
            iou = inter / (union + 1e-7) * (gt[:, 0][:, None] == pr[:, 5][None, :])
            for j, t in enumerate(iouv):
                pairs = sorted(((iou[g, p], g, p) for g, p in zip(*np.nonzero(iou >= t))), reverse=True)
                used_g, used_p = set(), set()
                for _, g, p in pairs:
                    if g not in used_g and p not in used_p:
                        used_g.add(g)
                        used_p.add(p)
                        correct[p, j] = True
        tp.append(correct)
        conf.extend(pr[:, 4].tolist())
        pcls.extend(pr[:, 5].tolist())
    if not tp:
        return 0.0, 0.0
    tp, conf, pcls, tcls = np.concatenate(tp), np.array(conf), np.array(pcls), np.array(tcls)
    order = np.argsort(-conf, kind="stable")
    tp, pcls = tp[order], pcls[order]
    aps = []
    for c in np.unique(tcls):
        m = pcls == c
        n_l = int((tcls == c).sum())
        ap = np.zeros(10)
        if m.sum():
            tpc, fpc = tp[m].cumsum(0), (~tp[m]).cumsum(0)
            rec, prec = tpc / (n_l + 1e-16), tpc / (tpc + fpc)
            for j in range(10):
                mrec = np.concatenate(([0.0], rec[:, j], [1.0]))
                mpre = np.concatenate(([1.0], prec[:, j], [0.0]))
                mpre = np.flip(np.maximum.accumulate(np.flip(mpre)))
                x = np.linspace(0, 1, 101)
                y = np.interp(x, mrec, mpre)
                ap[j] = np.sum((x[1:] - x[:-1]) * (y[1:] + y[:-1]) / 2)
        aps.append(ap)
    aps = np.array(aps)
    return float(aps[:, 0].mean()), float(aps.mean())
