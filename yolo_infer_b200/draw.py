"""Result consumers on the B200 path: the reference's `draw_detections` and result serialisers.

`draw_detections(image, results, ...)` has the signature of /root/reference/utils/visualization.py:18-25 and returns the
same pixels (bit-identical for the defaults every call site uses), but the drawing is ONE kernel launch over frames that are
already in device memory (csrc/draw.cu) instead of a Python loop with three device->host syncs and four cv2 calls per box.
`DetectionRasteriser.draw_batch` annotates a whole batch of device frames in place (the video / multi-camera case: frames
were uploaded for inference anyway; they never come back to the host unless the caller asks).
`save_detection_results` writes the txt / json / csv files of visualization.py:341-436 byte for byte from the host mirror
the engine already fetched with the call's single D2H - no per-box `.cpu()`.

No CPU fallback: without the CUDA library / an sm_100 device these raise.
"""
from __future__ import annotations

import csv
import ctypes as C
import json
from pathlib import Path
from typing import Any, Dict, List, Optional, Sequence, Union

import numpy as np
import torch

from . import _cabi as cabi

_ATLAS = Path(__file__).resolve().parent / "glyphs_simplex_0p5.npz"
NAME_STRIDE = 40


class DetectionRasteriser:
    """Per-device state of the rasteriser: glyph atlas, pen advances and the class-name table in device memory."""

    def __init__(self, names: Dict[int, str], device: Union[str, torch.device] = "cuda"):
        self.lib = cabi.load()
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("the rasteriser runs on sm_100 GPUs only (no CPU fallback; the CPU statement of it is oracle/draw_ref.py)")
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        self.device = dev
        a = np.load(_ATLAS)
        self.cell = [int(v) for v in a["cell"]]                                   # cell_h, cell_w, base_y, pad_x
        self.bits = torch.from_numpy(a["bits"].view(np.int32).copy()).to(dev)       # uint32 row masks (bit pattern kept)
        self.adv = torch.from_numpy(a["advance_half_px"].astype(np.int32)).to(dev)
        self.first_char, self.text_h = int(a["first_char"]), int(a["text_height"])
        self.h = C.c_void_p()
        with torch.cuda.device(dev):
            cabi.check(self.lib.y11_create(C.byref(self.h), dev.index), "y11_create")
        self.set_names(names)

    def __del__(self):
        try:
            if self.h:
                self.lib.y11_destroy(self.h)
                self.h = C.c_void_p()
        except Exception:
            pass

    def set_names(self, names: Dict[int, str]) -> None:
        nc = (max(names) + 1) if names else 1
        tab = np.zeros((nc, NAME_STRIDE), np.uint8)
        for k, v in names.items():
            b = str(v).encode("ascii", "replace")[:NAME_STRIDE - 1]
            tab[k, :len(b)] = np.frombuffer(b, np.uint8)
        self.names = dict(names)
        self.name_tab = torch.from_numpy(tab).to(self.device)
        self.font = cabi.Font(self.bits.data_ptr(), self.adv.data_ptr(), self.first_char, self.bits.shape[0], self.cell[0], self.cell[1],
                              self.cell[2], self.cell[3], self.text_h, self.name_tab.data_ptr(), nc, NAME_STRIDE)

    def draw_batch(self, frames: Union[torch.Tensor, Sequence[torch.Tensor]], dets: Sequence[torch.Tensor], counts: Sequence[int],
                   line_thickness: int = 2) -> None:
        """Annotate device uint8 BGR frames IN PLACE.  frames: [B,H,W,3] tensor or a list of [H,W,3] tensors; dets[i]: device fp32
        [>=counts[i], 6] rows x1,y1,x2,y2,conf,cls (original-image pixels); counts[i]: rows to draw."""
        items = (cabi.DrawItem * len(dets))()
        keep = []
        mh = mw = 1
        for i, (d, n) in enumerate(zip(dets, counts)):
            f = frames[i]
            assert f.is_cuda and f.dtype == torch.uint8 and f.ndim == 3 and f.shape[2] == 3 and f.stride(2) == 1 and f.stride(1) == 3
            d = d if (d.is_contiguous() and d.dtype == torch.float32) else d.float().contiguous()
            keep.append(d)
            items[i] = cabi.DrawItem(f.data_ptr(), f.shape[0], f.shape[1], f.stride(0), d.data_ptr() if d.numel() else None, None,
                                     int(n) if d.numel() else 0, int(d.shape[0]))
            mh, mw = max(mh, f.shape[0]), max(mw, f.shape[1])
        with torch.cuda.device(self.device):
            dev_items = torch.frombuffer(bytearray(bytes(items)), dtype=torch.uint8).to(self.device)
            s = torch.cuda.current_stream(self.device).cuda_stream
            cabi.check(self.lib.y11_draw_detections(self.h, dev_items.data_ptr(), len(dets), mh, mw, C.byref(self.font), line_thickness,
                                                    C.c_void_p(s)), "y11_draw_detections")
            self._keep = (dev_items, keep)      # alive until the next call (the launch is asynchronous)


_rasterisers: Dict[Any, DetectionRasteriser] = {}


def _rasteriser_for(names: Dict[int, str], device: torch.device) -> DetectionRasteriser:
    key = (device.index, tuple(sorted(names.items())))
    r = _rasterisers.get(key)
    if r is None:
        if len(_rasterisers) > 8:
            _rasterisers.clear()
        r = _rasterisers[key] = DetectionRasteriser(names, device)
    return r


def draw_detections(image, results: Any, class_names: Optional[Dict[int, str]] = None, line_thickness: int = 2,
                    font_scale: float = 0.5, font_thickness: int = 1):
    """Drop-in for utils/visualization.py:draw_detections.  `image`: BGR uint8 ndarray (returns an annotated COPY as ndarray, like
    the reference) or a device uint8 [H,W,3] tensor (returns an annotated device tensor; nothing crosses PCIe)."""
    is_np = isinstance(image, np.ndarray)
    if results is None or not hasattr(results, "boxes") or results.boxes is None:
        return image.copy() if is_np else image.clone()
    if font_scale != 0.5 or font_thickness != 1:
        raise ValueError("the B200 rasteriser implements the reference's defaults only: font_scale 0.5, font_thickness 1")
    boxes = results.boxes
    data = boxes.data
    if not data.is_cuda:
        data = data.cuda()
    dev = data.device
    names = dict(getattr(results, "names", None) or {})
    if class_names:
        names.update(class_names)          # the reference looks in class_names first, then in results.names
    ids = set(int(c) for c in boxes.cpu().data[:, 5].tolist()) if len(boxes) else set()
    for c in ids:
        names.setdefault(c, "Object")
    r = _rasteriser_for(names, dev)
    frame = (torch.from_numpy(np.ascontiguousarray(image)).to(dev) if is_np else image.to(dev).clone()).contiguous()
    r.draw_batch([frame], [data], [len(boxes)], line_thickness)
    return frame.cpu().numpy() if is_np else frame


# ---- serialisers (visualization.py:341-436), from the host mirror: one D2H per predict call, none here -------------------------
def _rows(results) -> List[List[float]]:
    if not hasattr(results, "boxes") or results.boxes is None:
        return []
    return results.boxes.cpu().data.tolist()


def save_detection_results(results: Any, output_path: str, format: str = "txt") -> None:
    Path(output_path).parent.mkdir(parents=True, exist_ok=True)
    fmt = format.lower()
    rows = _rows(results)
    if fmt == "txt":
        with open(output_path, "w") as f:
            for x1, y1, x2, y2, s, c in rows:
                f.write(f"{int(c)} {s:.6f} {x1:.6f} {y1:.6f} {x2:.6f} {y2:.6f}\n")
    elif fmt == "json":
        data = {"detections": [{"class_id": int(c), "confidence": float(s), "bbox": [float(x1), float(y1), float(x2), float(y2)]}
                               for x1, y1, x2, y2, s, c in rows]}
        with open(output_path, "w") as f:
            json.dump(data, f, indent=2)
    elif fmt == "csv":
        with open(output_path, "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(["class_id", "confidence", "x1", "y1", "x2", "y2"])
            for x1, y1, x2, y2, s, c in rows:
                w.writerow([int(c), float(s), float(x1), float(y1), float(x2), float(y2)])
    else:
        raise ValueError(f"Unsupported format: {format}")
