"""ORACLE (test infrastructure, never shipped, never on the product path): the reference's result consumers restated.

`draw_detections` follows /root/reference/utils/visualization.py:18-106 line by line (per-box `.cpu().numpy()` accessors,
cv2.rectangle / cv2.getTextSize / filled cv2.rectangle / cv2.putText with FONT_HERSHEY_SIMPLEX) and `get_color` follows :109-134;
`save_results_{txt,json,csv}` follow :367-436.  The drawing arithmetic itself is cv2's, which IS the reference's (cv2 is
importable here and on the GPU box), so parity of the B200 rasteriser against this file is parity against the reference.
"""
from __future__ import annotations

import csv
import json

import cv2
import numpy as np

COLORS = [(255, 0, 0), (0, 255, 0), (0, 0, 255), (255, 255, 0), (255, 0, 255), (0, 255, 255), (128, 0, 128), (255, 165, 0),
          (0, 128, 255), (128, 255, 0)]


def get_color(class_id: int):
    return COLORS[class_id % len(COLORS)]


def draw_detections(image: np.ndarray, results, class_names=None, line_thickness: int = 2, font_scale: float = 0.5,
                    font_thickness: int = 1) -> np.ndarray:
    if results is None:
        return image.copy()
    out = image.copy()
    if not hasattr(results, "boxes") or results.boxes is None:
        return out
    boxes = results.boxes
    for i in range(len(boxes)):
        x1, y1, x2, y2 = boxes.xyxy[i].cpu().numpy().astype(int)
        confidence = float(boxes.conf[i].cpu().numpy())
        class_id = int(boxes.cls[i].cpu().numpy())
        name = "Object"
        if class_names and class_id in class_names:
            name = class_names[class_id]
        elif hasattr(results, "names") and class_id in results.names:
            name = results.names[class_id]
        color = get_color(class_id)
        cv2.rectangle(out, (int(x1), int(y1)), (int(x2), int(y2)), color, line_thickness)
        label = f"{name}: {confidence:.2f}"
        size = cv2.getTextSize(label, cv2.FONT_HERSHEY_SIMPLEX, font_scale, font_thickness)[0]
        cv2.rectangle(out, (int(x1), int(y1) - size[1] - 10), (int(x1) + size[0] + 10, int(y1)), color, -1)
        cv2.putText(out, label, (int(x1) + 5, int(y1) - 5), cv2.FONT_HERSHEY_SIMPLEX, font_scale, (255, 255, 255), font_thickness)
    return out


def save_results_txt(results, path):
    with open(path, "w") as f:
        boxes = results.boxes
        for i in range(len(boxes)):
            c = int(boxes.cls[i].cpu().numpy())
            s = float(boxes.conf[i].cpu().numpy())
            x1, y1, x2, y2 = boxes.xyxy[i].cpu().numpy()
            f.write(f"{c} {s:.6f} {x1:.6f} {y1:.6f} {x2:.6f} {y2:.6f}\n")


def save_results_json(results, path):
    data = {"detections": []}
    boxes = results.boxes
    for i in range(len(boxes)):
        x1, y1, x2, y2 = boxes.xyxy[i].cpu().numpy()
        data["detections"].append({"class_id": int(boxes.cls[i].cpu().numpy()), "confidence": float(boxes.conf[i].cpu().numpy()),
                                   "bbox": [float(x1), float(y1), float(x2), float(y2)]})
    with open(path, "w") as f:
        json.dump(data, f, indent=2)


def save_results_csv(results, path):
    with open(path, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["class_id", "confidence", "x1", "y1", "x2", "y2"])
        boxes = results.boxes
        for i in range(len(boxes)):
            x1, y1, x2, y2 = boxes.xyxy[i].cpu().numpy()
            w.writerow([int(boxes.cls[i].cpu().numpy()), float(boxes.conf[i].cpu().numpy()), float(x1), float(y1), float(x2), float(y2)])
