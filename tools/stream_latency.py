"""BASELINE.json config 5, second half: a paced 30-fps 720p stream through `YOLO11Model.predict(frame, conf=, iou=)` - the
per-frame call of the reference's video loop (demos/detection_demo.py:182-196) - reporting per-frame latency from the BGR
frame on the host to the `Results` (H2D + letterbox resize 720x1280 -> 384x640 + forward + decode + NMS + one D2H).

  python tools/stream_latency.py [--model x] [--frames 150] [--fps 30] [--video file.mp4]"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="x")
    ap.add_argument("--frames", type=int, default=150)
    ap.add_argument("--fps", type=float, default=30.0)
    ap.add_argument("--video", default=None, help="decode frames with cv2.VideoCapture instead of synthetic 720p frames")
    args = ap.parse_args()
    from yolo_infer_b200 import topology as T
    from yolo_infer_b200.engine import YOLO
    eng = YOLO.from_state_dict(T.synthetic_state_dict(args.model, 80, seed=0), args.model).to("cuda:0")
    from yolo_infer_b200.synth import condition_synthetic_weights
    condition_synthetic_weights(eng, (384, 640), batch=2, seed=0)
    cap = None
    if args.video:
        import cv2
        cap = cv2.VideoCapture(args.video)
    rng = np.random.default_rng(0)
    pool = [rng.integers(0, 256, (720, 1280, 3), dtype=np.uint8) for _ in range(8)]

    def next_frame(i):
        if cap is not None:
            ok, f = cap.read()
            return f if ok else None
        return pool[i % len(pool)]

    for i in range(5):                                 # warm-up: builds and captures the pipeline
        eng.predict(next_frame(i), conf=0.25, iou=0.45, verbose=False)
    lat, dec, n_det = [], [], 0
    period = 1.0 / args.fps
    t_start = time.perf_counter()
    for i in range(args.frames):
        due = t_start + i * period
        now = time.perf_counter()
        if now < due:
            time.sleep(due - now)
        t0 = time.perf_counter()
        f = next_frame(i)
        if f is None:
            break
        t1 = time.perf_counter()
        r = eng.predict(f, conf=0.25, iou=0.45, verbose=False)[0]
        n_det += len(r.boxes)
        t2 = time.perf_counter()
        dec.append(1e3 * (t1 - t0))
        lat.append(1e3 * (t2 - t1))
    wall = time.perf_counter() - t_start
    lat.sort()
    pct = lambda q: lat[min(len(lat) - 1, int(q * len(lat)))]  # noqa: E731
    print(json.dumps({"model": f"yolo11{args.model}", "source": args.video or "synthetic 720x1280 BGR frames", "frames": len(lat),
                      "paced_fps": args.fps, "achieved_fps": len(lat) / wall, "frame_to_results_ms_p50": pct(0.5),
                      "frame_to_results_ms_p99": pct(0.99), "frame_to_results_ms_max": lat[-1],
                      "decode_ms_mean": sum(dec) / max(1, len(dec)), "mean_detections": n_det / max(1, len(lat)),
                      "max_sustainable_fps": 1e3 / pct(0.5)}))


if __name__ == "__main__":
    main()
