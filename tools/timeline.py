"""Kernel timeline of the steady-state `value` loop of bench.py (two steps in flight, CUDA-graph replays) captured with
torch.profiler (CUPTI): GPU busy time, idle gaps, concurrency and in-graph kernel durations.

  python tools/timeline.py [--model n] [--steps 6] [--streams 2]"""
import argparse
import json
import os
import sys
import tempfile
from collections import defaultdict

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="n")
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--streams", type=int, default=2)
    ap.add_argument("--batch", type=int, default=64)
    args = ap.parse_args()
    from yolo_infer_b200 import topology as T
    from yolo_infer_b200.engine import YOLO
    dev = torch.device("cuda:0")
    eng = YOLO.from_state_dict(T.synthetic_state_dict(args.model, 80, seed=0), args.model).to(dev)
    from yolo_infer_b200.synth import condition_synthetic_weights
    condition_synthetic_weights(eng, (640, 640), batch=2, seed=0)
    B, S, NROT, NS = args.batch, 640, 4, args.streams
    g = torch.Generator().manual_seed(0)
    frames = [torch.randint(0, 256, (B, S, S, 3), generator=g, dtype=torch.uint8).to(dev) for _ in range(NROT)]
    main_s = torch.cuda.current_stream(dev)
    streams = [main_s] + [torch.cuda.Stream(dev) for _ in range(NS - 1)]
    pipes = []
    for j, f in enumerate(frames):
        with torch.cuda.stream(streams[j % NS]):
            pipes.append(eng.pipeline(B, S, S, S, True, 0.25, 0.7, 300, frames=f, graph=True, replica=j % NS))
    torch.cuda.synchronize()

    def run(n):
        for st in streams[1:]:
            st.wait_stream(main_s)
        for i in range(n):
            with torch.cuda.stream(streams[(i % NROT) % NS]):
                pipes[i % NROT].run()
        for st in streams[1:]:
            main_s.wait_stream(st)
        torch.cuda.synchronize()

    run(8)
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        run(args.steps)
    path = os.path.join(tempfile.mkdtemp(), "trace.json")
    prof.export_chrome_trace(path)
    ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") == "kernel" and e.get("ph") == "X"]
    ev.sort(key=lambda e: e["ts"])
    if not ev:
        print("no kernel events captured")
        return
    t0, t1 = ev[0]["ts"], max(e["ts"] + e["dur"] for e in ev)
    span = t1 - t0
    total = sum(e["dur"] for e in ev)
    # union of busy intervals + gaps
    busy, gaps, cur_end, last = 0.0, [], ev[0]["ts"], ev[0]
    for e in ev:
        if e["ts"] > cur_end:
            gaps.append((e["ts"] - cur_end, last["name"][:40], e["name"][:40]))
            busy_start = e["ts"]
        cur_end_new = max(cur_end, e["ts"] + e["dur"])
        busy += max(0.0, cur_end_new - max(cur_end, e["ts"]))
        if e["ts"] + e["dur"] >= cur_end:
            last = e
        cur_end = cur_end_new
    per = defaultdict(lambda: [0, 0.0])
    for e in ev:
        n = e["name"].replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("void ", "").split("(")[0]
        per[n][0] += 1
        per[n][1] += e["dur"]
    print(f"YOLO11{args.model} B={B}, {args.steps} steps on {NS} stream(s): span {span / 1e3:.3f} ms = {span / args.steps / 1e3:.3f} ms/step; "
          f"{len(ev)} kernels; sum of kernel durations {total / 1e3:.3f} ms ({total / args.steps / 1e3:.3f} ms/step); "
          f"GPU busy (>= 1 kernel running) {busy / 1e3:.3f} ms = {100 * busy / span:.1f} % of the span; mean concurrency while busy "
          f"{total / busy:.2f}")
    print("(kernels launched with programmatic dependent launch start their clock while their predecessor still runs: the sum of "
          "durations and the concurrency include that waiting)")
    print(f"idle gaps: {len(gaps)} totalling {sum(g_[0] for g_ in gaps) / 1e3:.3f} ms; largest:")
    for g_ in sorted(gaps, reverse=True)[:8]:
        print(f"   {g_[0]:7.1f} us  after {g_[1]}  before {g_[2]}")
    print("per kernel (in-graph durations):")
    for k, v in sorted(per.items(), key=lambda kv: -kv[1][1])[:14]:
        print(f"   {k[:48]:48s} {v[0]:5d} launches {v[1] / args.steps:9.1f} us/step  ({100 * v[1] / total:.1f} %)")


if __name__ == "__main__":
    main()
