"""Detection validation (mAP) on top of the B200 predict path - SURVEY.md section 8(f) row 1.

What the reference gets from ``ultralytics.YOLO.val`` (``core/model.py:180-195``) and reads back in
``core/validator.py:339-359``: an object with ``.box.{map, map50, map75, mp, mr}`` and a ``.speed`` dict.
The device work (letterbox, forward, multi-label NMS at conf 0.001) is ``YOLO.predict`` - the hot path of this repository;
what lives here is the validator's book-keeping, which in ultralytics is CPU numpy as well (``utils/metrics.py``):

* dataset: a YOLO dataset yaml (``path``, ``val``, ``names``) or a directory of images; labels are ``class cx cy w h``
  (normalised) text files found by replacing ``/images/`` with ``/labels/`` (ultralytics ``img2label_paths``);
* matching: ultralytics ``DetectionValidator.match_predictions`` - IoU matrix restricted to equal classes, for each of
  the 10 thresholds 0.50:0.05:0.95 greedy one-to-one matches in decreasing IoU;
* ``ap_per_class``: per class precision/recall curves over confidence-sorted predictions, AP by 101-point interpolation
  of the precision envelope (COCO), precision/recall reported at the confidence that maximises the smoothed mean F1.

Batching follows ultralytics' RECT val dataloader (``data/dataset.py: set_rectangle``, ``data/base.py: load_image``, the reference
reaches it through ``model.val`` at ``core/validator.py:121-141``): images sorted by aspect ratio h/w, consecutive groups of
``batch`` images share one canvas ``ceil([h, w] * imgsz / 32 + 0.5) * 32`` derived from the group's extreme aspect ratio, every
image is resized so that its long side is ``imgsz`` (``ceil`` of the scaled size, bilinear) and centred on the canvas with
border 114 (``LetterBox(auto=False, scaleup=False)``); predictions go back to original pixels with
``scale_boxes(..., ratio_pad=((h1/h0, w1/w0), (left, top)))``.  ``rect=False`` keeps the square letterbox of ``predict``.
"""
from __future__ import annotations

import time
from pathlib import Path
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

IOUV = np.linspace(0.5, 0.95, 10)
IMG_EXT = {".jpg", ".jpeg", ".png", ".bmp", ".webp", ".tif", ".tiff"}


# ------------------------------------------------------------------------------------------------ metric maths
def box_iou(a: np.ndarray, b: np.ndarray, eps: float = 1e-7) -> np.ndarray:
    """IoU matrix [len(a), len(b)] of xyxy boxes (ultralytics ``metrics.box_iou``)."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    lt = np.maximum(a[:, None, :2], b[None, :, :2])
    rb = np.minimum(a[:, None, 2:], b[None, :, 2:])
    inter = np.clip(rb - lt, 0, None).prod(2)
    area_a = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])
    area_b = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    return inter / (area_a[:, None] + area_b[None, :] - inter + eps)


def match_predictions(pred_cls: np.ndarray, true_cls: np.ndarray, iou: np.ndarray) -> np.ndarray:
    """-> bool [n_pred, 10]: prediction i is a true positive at threshold j.  `iou` is [n_gt, n_pred]."""
    correct = np.zeros((pred_cls.shape[0], IOUV.shape[0]), bool)
    iou = iou * (true_cls[:, None] == pred_cls[None, :])
    for j, thr in enumerate(IOUV):
        g, p = np.nonzero(iou >= thr)
        if g.size:
            m = np.stack((g, p), 1)
            if g.size > 1:
                m = m[iou[g, p].argsort()[::-1]]
                m = m[np.unique(m[:, 1], return_index=True)[1]]   # one ground truth per prediction ...
                m = m[np.unique(m[:, 0], return_index=True)[1]]   # ... and one prediction per ground truth
            correct[m[:, 1].astype(int), j] = True
    return correct


def compute_ap(recall: np.ndarray, precision: np.ndarray) -> float:
    mrec = np.concatenate(([0.0], recall, [1.0]))
    mpre = np.concatenate(([1.0], precision, [0.0]))
    mpre = np.flip(np.maximum.accumulate(np.flip(mpre)))
    x = np.linspace(0, 1, 101)
    y = np.interp(x, mrec, mpre)
    return float(np.sum((x[1:] - x[:-1]) * (y[1:] + y[:-1]) / 2))   # trapezoid


def smooth(y: np.ndarray, f: float = 0.05) -> np.ndarray:
    nf = round(len(y) * f * 2) // 2 + 1
    p = np.ones(nf // 2)
    yp = np.concatenate((p * y[0], y, p * y[-1]), 0)
    return np.convolve(yp, np.ones(nf) / nf, mode="valid")


def ap_per_class(tp: np.ndarray, conf: np.ndarray, pred_cls: np.ndarray, target_cls: np.ndarray, eps: float = 1e-16):
    """-> (p [nc], r [nc], ap [nc, 10], classes [nc]) for the classes present in the ground truth."""
    order = np.argsort(-conf, kind="stable")
    tp, conf, pred_cls = tp[order], conf[order], pred_cls[order]
    classes, nt = np.unique(target_cls, return_counts=True)
    nc = classes.shape[0]
    x = np.linspace(0, 1, 1000)
    ap = np.zeros((nc, tp.shape[1]))
    p_curve, r_curve = np.zeros((nc, 1000)), np.zeros((nc, 1000))
    for ci, c in enumerate(classes):
        i = pred_cls == c
        n_l, n_p = nt[ci], int(i.sum())
        if n_p == 0 or n_l == 0:
            continue
        fpc = (~tp[i]).cumsum(0)
        tpc = tp[i].cumsum(0)
        recall = tpc / (n_l + eps)
        r_curve[ci] = np.interp(-x, -conf[i], recall[:, 0], left=0)
        precision = tpc / (tpc + fpc)
        p_curve[ci] = np.interp(-x, -conf[i], precision[:, 0], left=1)
        for j in range(tp.shape[1]):
            ap[ci, j] = compute_ap(recall[:, j], precision[:, j])
    f1 = 2 * p_curve * r_curve / (p_curve + r_curve + eps)
    i = int(smooth(f1.mean(0), 0.1).argmax()) if nc else 0
    return p_curve[:, i], r_curve[:, i], ap, classes.astype(int)


class Metric:
    """The `.box` object: the attribute names `core/validator.py:339-349` reads."""

    def __init__(self, p, r, ap, classes, nc: int):
        self.p, self.r, self.all_ap, self.ap_class_index, self.nc = p, r, ap, classes, nc

    @property
    def ap50(self):
        return self.all_ap[:, 0] if len(self.all_ap) else np.zeros(0)

    @property
    def ap(self):
        return self.all_ap.mean(1) if len(self.all_ap) else np.zeros(0)

    @property
    def mp(self) -> float:
        return float(self.p.mean()) if len(self.p) else 0.0

    @property
    def mr(self) -> float:
        return float(self.r.mean()) if len(self.r) else 0.0

    @property
    def map50(self) -> float:
        return float(self.all_ap[:, 0].mean()) if len(self.all_ap) else 0.0

    @property
    def map75(self) -> float:
        return float(self.all_ap[:, 5].mean()) if len(self.all_ap) else 0.0

    @property
    def map(self) -> float:
        return float(self.all_ap.mean()) if len(self.all_ap) else 0.0

    @property
    def maps(self) -> np.ndarray:
        m = np.full(self.nc, self.map)
        for i, c in enumerate(self.ap_class_index):
            m[c] = self.ap[i]
        return m

    def mean_results(self):
        return [self.mp, self.mr, self.map50, self.map]


class DetMetrics:
    def __init__(self, box: Metric, speed: Dict[str, float], names: Dict[int, str], n_images: int, n_labels: int):
        self.box, self.speed, self.names = box, speed, names
        self.n_images, self.n_labels = n_images, n_labels

    @property
    def results_dict(self) -> Dict[str, float]:
        return {"metrics/precision(B)": self.box.mp, "metrics/recall(B)": self.box.mr, "metrics/mAP50(B)": self.box.map50,
                "metrics/mAP50-95(B)": self.box.map, "fitness": 0.1 * self.box.map50 + 0.9 * self.box.map}

    @property
    def fitness(self) -> float:
        return self.results_dict["fitness"]

    @property
    def maps(self):
        return self.box.maps

    def __repr__(self) -> str:
        return (f"DetMetrics(images={self.n_images}, labels={self.n_labels}, P={self.box.mp:.3f}, R={self.box.mr:.3f}, "
                f"mAP50={self.box.map50:.3f}, mAP50-95={self.box.map:.3f})")


def evaluate(preds: Sequence[np.ndarray], gts: Sequence[np.ndarray], nc: int) -> Metric:
    """preds[i]: [n,6] x1,y1,x2,y2,conf,cls (original-image pixels); gts[i]: [m,5] cls,x1,y1,x2,y2."""
    tps, confs, pcls, tcls = [], [], [], []
    for pr, gt in zip(preds, gts):
        pr = np.asarray(pr, np.float64).reshape(-1, 6)
        gt = np.asarray(gt, np.float64).reshape(-1, 5)
        tcls.append(gt[:, 0])
        if pr.shape[0] == 0:
            continue
        tp = np.zeros((pr.shape[0], IOUV.shape[0]), bool)
        if gt.shape[0]:
            tp = match_predictions(pr[:, 5], gt[:, 0], box_iou(gt[:, 1:], pr[:, :4]))
        tps.append(tp)
        confs.append(pr[:, 4])
        pcls.append(pr[:, 5])
    tcls_all = np.concatenate(tcls) if tcls else np.zeros(0)
    if not tps or tcls_all.size == 0:
        return Metric(np.zeros(0), np.zeros(0), np.zeros((0, IOUV.shape[0])), np.zeros(0, int), nc)
    p, r, ap, classes = ap_per_class(np.concatenate(tps), np.concatenate(confs), np.concatenate(pcls), tcls_all)
    return Metric(p, r, ap, classes, nc)


# ------------------------------------------------------------------------------------------------ dataset
def _parse_yaml(path: Path) -> dict:
    import yaml
    with open(path) as f:
        return yaml.safe_load(f) or {}


def img2label_path(img: Path) -> Path:
    parts = list(img.parts)
    if "images" in parts:
        idx = len(parts) - 1 - parts[::-1].index("images")
        parts[idx] = "labels"
    return Path(*parts).with_suffix(".txt")


def load_dataset(data) -> Tuple[List[Path], Optional[Dict[int, str]]]:
    p = Path(str(data))
    names = None
    if p.suffix in (".yaml", ".yml"):
        cfg = _parse_yaml(p)
        root = Path(cfg.get("path", p.parent))
        if not root.is_absolute():
            root = (p.parent / root).resolve()
        val = cfg.get("val") or cfg.get("test")
        if val is None:
            raise ValueError(f"{p}: dataset yaml has no 'val' entry")
        src = Path(val) if Path(val).is_absolute() else root / val
        nm = cfg.get("names")
        if isinstance(nm, (list, tuple)):
            names = dict(enumerate(nm))
        elif isinstance(nm, dict):
            names = {int(k): str(v) for k, v in nm.items()}
    else:
        src = p
    if src.is_file() and src.suffix == ".txt":
        files = [Path(l.strip()) if Path(l.strip()).is_absolute() else (src.parent / l.strip()) for l in src.read_text().splitlines() if l.strip()]
    elif src.is_dir():
        files = sorted(f for f in src.rglob("*") if f.suffix.lower() in IMG_EXT)
    else:
        raise FileNotFoundError(f"validation images not found: {src}")
    if not files:
        raise FileNotFoundError(f"no images under {src}")
    return files, names


def read_labels(img: Path, w: int, h: int) -> np.ndarray:
    """-> [m, 5] cls, x1, y1, x2, y2 in pixels of the original image."""
    lp = img2label_path(img)
    if not lp.exists():
        return np.zeros((0, 5))
    rows = [l.split() for l in lp.read_text().splitlines() if l.strip()]
    if not rows:
        return np.zeros((0, 5))
    a = np.asarray([[float(v) for v in r[:5]] for r in rows], np.float64)
    out = np.zeros_like(a)
    out[:, 0] = a[:, 0]
    out[:, 1] = (a[:, 1] - a[:, 3] / 2) * w
    out[:, 2] = (a[:, 2] - a[:, 4] / 2) * h
    out[:, 3] = (a[:, 1] + a[:, 3] / 2) * w
    out[:, 4] = (a[:, 2] + a[:, 4] / 2) * h
    return out


def rect_batches(shapes: Sequence[Tuple[int, int]], batch: int, imgsz: int = 640, stride: int = 32, pad: float = 0.5):
    """ultralytics `YOLODataset.set_rectangle` restated.  shapes[i] = (h0, w0).  -> (order, canvases): `order` = image indices sorted by
    aspect ratio h/w (stable), `canvases[k]` = (H, W) of batch k = images order[k*batch:(k+1)*batch]."""
    s = np.asarray(shapes, np.float64).reshape(-1, 2)
    ar = s[:, 0] / s[:, 1]
    order = np.argsort(ar, kind="stable")
    ar = ar[order]
    n = len(order)
    nb = (n + batch - 1) // batch
    canvases = []
    for k in range(nb):
        ari = ar[k * batch:(k + 1) * batch]
        mini, maxi = ari.min(), ari.max()
        shp = [1.0, 1.0]
        if maxi < 1:
            shp = [maxi, 1.0]
        elif mini > 1:
            shp = [1.0, 1.0 / mini]
        hw = np.ceil(np.asarray(shp) * imgsz / stride + pad).astype(int) * stride
        canvases.append((int(hw[0]), int(hw[1])))
    return [int(i) for i in order], canvases


def rect_geometry(h0: int, w0: int, canvas: Tuple[int, int], imgsz: int = 640):
    """Geometry of one image of a rect batch: ultralytics `load_image(rect_mode=True)` (long side -> imgsz, ceil) followed by
    `LetterBox(canvas, auto=False, scaleup=False, center=True)`.  -> ((new_w, new_h, top, left), [gain, pad_x, pad_y, w0, h0])."""
    import math
    r = imgsz / max(h0, w0)
    if r != 1:
        w1, h1 = min(math.ceil(w0 * r), imgsz), min(math.ceil(h0 * r), imgsz)
    else:
        w1, h1 = w0, h0
    H, W = canvas
    r2 = min(min(H / h1, W / w1), 1.0)                    # scaleup=False
    new_w, new_h = int(round(w1 * r2)), int(round(h1 * r2))
    dw, dh = (W - new_w) / 2, (H - new_h) / 2
    top, left = int(round(dh - 0.1)), int(round(dw - 0.1))
    gain = (h1 / h0) * r2                                 # ratio_pad[0][0]: the HEIGHT ratio is applied to both axes
    return (new_w, new_h, top, left), [gain, float(left), float(top), float(w0), float(h0)]


def validate(engine, data, imgsz: int = 640, batch: int = 16, conf: float = 0.001, iou: float = 0.6, max_det: int = 300,
             verbose: bool = False, rect: bool = True, **_ignored) -> DetMetrics:
    """Run the engine (multi-label NMS, ultralytics val thresholds) over the dataset in rect batches and score it."""
    import cv2
    files, names = load_dataset(data)
    names = names or engine.names
    preds, gts = [], []
    speed = {"preprocess": 0.0, "inference": 0.0, "loss": 0.0, "postprocess": 0.0}
    t_read = 0.0
    canvases = None
    if rect:
        t0 = time.perf_counter()
        shapes = []
        for f in files:                      # header-only read would do; the files are decoded again batch by batch below
            im = cv2.imread(str(f))
            if im is None:
                raise FileNotFoundError(f"cannot read image {f}")
            shapes.append(im.shape[:2])
        t_read += time.perf_counter() - t0
        order, canvases = rect_batches(shapes, batch, imgsz)
        files = [files[i] for i in order]
    for i in range(0, len(files), batch):
        chunk = files[i:i + batch]
        t0 = time.perf_counter()
        imgs = []
        for f in chunk:
            im = cv2.imread(str(f))
            if im is None:
                raise FileNotFoundError(f"cannot read image {f}")
            imgs.append(im)
        t_read += time.perf_counter() - t0
        if rect:
            H, W = canvases[i // batch]
            gr = [rect_geometry(im.shape[0], im.shape[1], (H, W), imgsz) for im in imgs]
            res = engine.predict_letterboxed(imgs, [g[0] for g in gr], H, W, [g[1] for g in gr], paths=[str(f) for f in chunk],
                                             conf=conf, iou=iou, max_det=max_det, multi_label=True)
        else:
            # graph=False: a dataset has many distinct image sizes; a captured pipeline per size would cost more than it saves
            res = engine.predict(imgs, conf=conf, iou=iou, max_det=max_det, imgsz=imgsz, multi_label=True, verbose=False, graph=False)
        for f, im, r in zip(chunk, imgs, res):
            preds.append(r.cpu().boxes.data.numpy())
            gts.append(read_labels(f, im.shape[1], im.shape[0]))
            for k in ("preprocess", "inference", "postprocess"):
                speed[k] += float(r.speed.get(k) or 0.0)
    n = max(len(files), 1)
    speed = {k: v / n for k, v in speed.items()}
    speed["preprocess"] += 1e3 * t_read / n        # image decoding counts as preprocessing, as in ultralytics' dataloader profile
    box = evaluate(preds, gts, len(names))
    m = DetMetrics(box, speed, names, len(files), int(sum(len(g) for g in gts)))
    if verbose:
        print(m)
    return m
