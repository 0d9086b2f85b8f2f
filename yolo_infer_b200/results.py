"""Results / Boxes containers with the layout the reference's consumers read.

The reference receives these objects from ultralytics (engine/results.py, un-vendored) and touches exactly:
``results[0].boxes`` (len, truthiness, iteration -> 1-row Boxes), ``boxes.xyxy[i].cpu().numpy()``,
``boxes.conf[i]``, ``boxes.cls[i]``, ``result.names[class_id]``
(/root/reference/utils/visualization.py:52-74, /root/reference/demos/detection_demo.py:96-132,
SURVEY.md section 8b).  Layout: ``data`` float32 [n, 6] = x1, y1, x2, y2 (original-image pixels, clipped), conf, cls;
rows sorted by confidence descending; n <= max_det.  Tensors stay on the device they were produced on.
"""
from __future__ import annotations

from collections.abc import Sequence
from typing import Dict, Iterator, List, Optional, Tuple

import numpy as np
import torch


class Boxes:
    def __init__(self, data: Optional[torch.Tensor], orig_shape: Tuple[int, int], host: Optional[torch.Tensor] = None,
                 lazy: Optional[tuple] = None):
        # lazy = (device batch [B,max_det,6], host batch [B,max_det,6], image index, row count): the engine hands every image
        # of a batch a view descriptor instead of slicing 2 x B tensors up front (64 images: ~0.3 ms of Python per call);
        # the rows are materialised on first access of `.data` / `.cpu()`.
        self._lazy = lazy
        if data is not None:
            if data.ndim == 1:
                data = data[None, :]
            assert data.shape[-1] == 6, f"expected [n,6] boxes, got {tuple(data.shape)}"
        else:
            assert lazy is not None
        self._data = data
        self.orig_shape = tuple(orig_shape)
        self.is_track = False
        self.id = None
        # Optional host mirror of `data`: the engine fetches the whole batch's results with ONE device->host copy per
        # predict call, so `.cpu()` / `.numpy()` on each Results need no further transfer or stream sync.
        if host is not None and host.ndim == 1:
            host = host[None, :]
        self._host = host

    @property
    def data(self) -> torch.Tensor:
        if self._data is None:
            dev, _, i, n = self._lazy
            self._data = dev[i, :n]
        return self._data

    @data.setter
    def data(self, value) -> None:
        self._data = value

    def _host_rows(self) -> Optional[torch.Tensor]:
        if self._host is None and self._lazy is not None and self._lazy[1] is not None:
            _, host, i, n = self._lazy
            self._host = host[i, :n]
        return self._host

    # ---- the accessors the reference reads ------------------------------------------------------
    @property
    def xyxy(self) -> torch.Tensor:
        return self.data[:, :4]

    @property
    def conf(self) -> torch.Tensor:
        return self.data[:, -2]

    @property
    def cls(self) -> torch.Tensor:
        return self.data[:, -1]

    # ---- derived layouts (ultralytics API) --------------------------------------------------------
    @property
    def xywh(self) -> torch.Tensor:
        b = self.xyxy
        out = torch.empty_like(b)
        out[:, 0] = (b[:, 0] + b[:, 2]) / 2
        out[:, 1] = (b[:, 1] + b[:, 3]) / 2
        out[:, 2] = b[:, 2] - b[:, 0]
        out[:, 3] = b[:, 3] - b[:, 1]
        return out

    def _norm(self, b: torch.Tensor) -> torch.Tensor:
        h, w = self.orig_shape
        out = b.clone()
        out[:, [0, 2]] /= w
        out[:, [1, 3]] /= h
        return out

    @property
    def xyxyn(self) -> torch.Tensor:
        return self._norm(self.xyxy)

    @property
    def xywhn(self) -> torch.Tensor:
        return self._norm(self.xywh)

    @property
    def shape(self):
        return self.data.shape

    def __len__(self) -> int:
        if self._data is None:
            return int(self._lazy[3])
        return int(self._data.shape[0])

    def __getitem__(self, idx) -> "Boxes":
        host = self._host_rows()
        return Boxes(self.data[idx], self.orig_shape, host[idx] if host is not None else None)

    def __iter__(self) -> Iterator["Boxes"]:
        for i in range(len(self)):
            yield self[i]

    def cpu(self) -> "Boxes":
        host = self._host_rows()
        return Boxes(host if host is not None else self.data.cpu(), self.orig_shape)

    def numpy(self) -> "Boxes":
        b = Boxes.__new__(Boxes)
        b._lazy, b._data, b.orig_shape, b.is_track, b.id, b._host = None, self.cpu().data.numpy(), self.orig_shape, False, None, None
        return b

    def cuda(self) -> "Boxes":
        return Boxes(self.data.cuda(), self.orig_shape)

    def to(self, *a, **k) -> "Boxes":
        return Boxes(self.data.to(*a, **k), self.orig_shape)

    def __repr__(self) -> str:
        return f"Boxes(n={len(self)}, orig_shape={self.orig_shape}, device={getattr(self.data, 'device', 'numpy')})"


class Results:
    def __init__(self, orig_img: Optional[np.ndarray], path: str, names: Dict[int, str], boxes: torch.Tensor,
                 orig_shape: Tuple[int, int], speed: Optional[Dict[str, float]] = None,
                 host_boxes: Optional[torch.Tensor] = None, lazy: Optional[tuple] = None):
        self.orig_img = orig_img
        self.orig_shape = tuple(orig_shape)
        self.path = path
        self.names = names
        self.boxes = Boxes(boxes, self.orig_shape, host_boxes, lazy)
        self.masks = None
        self.probs = None
        self.keypoints = None
        self.obb = None
        self.speed = speed or {"preprocess": None, "inference": None, "postprocess": None}

    def __len__(self) -> int:
        return len(self.boxes)

    def cpu(self) -> "Results":
        return Results(self.orig_img, self.path, self.names, self.boxes.cpu().data, self.orig_shape, self.speed)

    def to(self, *a, **k) -> "Results":
        return Results(self.orig_img, self.path, self.names, self.boxes.data.to(*a, **k), self.orig_shape, self.speed)

    def summary(self):
        d = self.boxes.cpu().data.tolist()
        return [{"name": self.names[int(r[5])], "class": int(r[5]), "confidence": r[4],
                 "box": {"x1": r[0], "y1": r[1], "x2": r[2], "y2": r[3]}} for r in d]

    def __repr__(self) -> str:
        return f"Results(path={self.path!r}, orig_shape={self.orig_shape}, boxes={len(self.boxes)})"



class ResultsBatch(Sequence):
    """What `predict` returns for a fixed-shape batch: a read-only sequence of `Results` (``len``, indexing, slicing, iteration -
    everything the reference's callers do with the list ultralytics returns: ``results[0]``, ``for r in results``) whose items are
    built on first access.  The batch's detections live in ONE device tensor and ONE host mirror (the call's single D2H), so a
    throughput loop over hundreds of images does not pay for hundreds of Python objects it never looks at; `det` / `counts`
    give the whole batch at once (rows beyond counts[i] are stale)."""

    def __init__(self, det: Optional[torch.Tensor], det_host: torch.Tensor, counts: List[int], names: Dict[int, str],
                 orig_shapes, paths=None, orig_imgs=None, speed: Optional[Dict[str, float]] = None):
        self.det, self.det_host, self.counts, self.names = det, det_host, counts, names
        self._shapes, self._paths, self._imgs, self.speed = orig_shapes, paths, orig_imgs, speed
        self._items: Dict[int, Results] = {}

    def __len__(self) -> int:
        return len(self.counts)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(len(self)))]
        if i < 0:
            i += len(self)
        if not 0 <= i < len(self):
            raise IndexError(i)
        r = self._items.get(i)
        if r is None:
            shape = self._shapes[i] if isinstance(self._shapes, list) else self._shapes
            path = self._paths[i] if self._paths is not None else f"image{i}.jpg"
            img = self._imgs[i] if self._imgs is not None else None
            n = self.counts[i]
            if self.det is not None:
                r = Results(img, path, self.names, None, shape, self.speed, None, (self.det, self.det_host, i, n))
            else:
                r = Results(img, path, self.names, self.det_host[i, :n], shape, self.speed, self.det_host[i, :n])
            self._items[i] = r
        return r

    def __repr__(self) -> str:
        return f"ResultsBatch(n={len(self)}, detections={sum(self.counts)})"
