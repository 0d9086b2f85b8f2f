#!/usr/bin/env python
"""Post-processing at validation scale (BASELINE config #4): YOLO11m, 1280x1280, batch 16, `conf=0.001, iou=0.6, multi_label`
(what `model.val` runs, core/validator.py:121-141) next to the predict setting (conf 0.25, single label).  Prints one JSON line:
forward ms, decode ms, sort+NMS ms and candidates per image for both settings, conditioned synthetic weights.

  python tools/post_scale.py [--model m] [--imgsz 1280] [--batch 16]
"""
import argparse
import json
import statistics
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from yolo_infer_b200 import topology as T  # noqa: E402
from yolo_infer_b200.engine import YOLO, letterbox_geometry  # noqa: E402
from yolo_infer_b200.synth import condition_synthetic_weights  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="m")
    ap.add_argument("--imgsz", type=int, default=1280)
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--cls-prior", type=float, default=0.01, help="class-bias prior of the synthetic weights (0.01: nearly every (anchor, class) "
                                                                   "pair clears conf 0.001; 0.0002: ~20 % do)")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    S, B = a.imgsz, a.batch
    eng = YOLO.from_state_dict(T.synthetic_state_dict(a.model, 80, seed=0), a.model).to(dev)
    condition_synthetic_weights(eng, (640, 640), batch=2, seed=0, cls_prior=a.cls_prior)
    net = eng.compiled(B, S, S)
    g = torch.Generator().manual_seed(0)
    frames = torch.randint(0, 256, (B, S, S, 3), generator=g, dtype=torch.uint8).to(dev)
    geoms = [letterbox_geometry(S, S, (S, S), True)] * B
    eng.preprocess_images(net, list(frames), geoms)
    stream = torch.cuda.current_stream(dev)
    fwd = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.forward(net)
        e1.record()
        torch.cuda.synchronize()
        fwd.append(e0.elapsed_time(e1))
    rows = torch.tensor([[1.0, 0.0, 0.0, S, S]] * B, dtype=torch.float32, device=dev)
    out = {"model": f"yolo11{a.model}", "imgsz": S, "batch": B, "anchors": net.A, "forward_ms": statistics.median(fwd), "cls_prior": a.cls_prior}
    for name, conf, iou, ml in (("predict", 0.25, 0.7, False), ("val", 0.001, 0.6, True)):
        dms, nms = [], []
        for _ in range(3):
            d, n = eng.postprocess_timed(net, rows, conf, iou, 300, multi_label=ml)
            dms.append(d)
            nms.append(n)
        _, cnt, ncand = eng.postprocess(net, rows, conf, iou, 300, multi_label=ml)
        torch.cuda.synchronize()
        out[name] = {"conf": conf, "iou": iou, "multi_label": ml, "decode_ms": statistics.median(dms), "sort_nms_ms": statistics.median(nms),
                     "candidates_per_image": float(ncand.float().mean()), "detections_per_image": float(cnt.float().mean()),
                     "post_share_of_step": (statistics.median(dms) + statistics.median(nms)) / (statistics.median(fwd) + statistics.median(dms) + statistics.median(nms))}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
