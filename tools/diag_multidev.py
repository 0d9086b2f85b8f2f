"""Diagnostic: engines on two devices of one process, sequentially and from worker threads."""
import gc, os, sys, traceback
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from oracle import yolo11_ref as R
from yolo_infer_b200.engine import YOLO

sd = R.build("n", init="calibrated", seed=0).state_dict()
mode = sys.argv[1]
g = torch.Generator().manual_seed(4)
batch = torch.randint(0, 256, (6, 320, 320, 3), generator=g, dtype=torch.uint8).pin_memory()


def check(tag):
    for d in range(torch.cuda.device_count()):
        torch.cuda.synchronize(d)
    print("sync ok:", tag, flush=True)


try:
    order = (1, 0) if mode == "dev1first" else (0, 1)
    keep = []
    if mode == "nogc":
        gc.disable()
    outs = []
    for d in order:
        e = YOLO.from_state_dict(sd, "n").to(f"cuda:{d}")
        if mode in ("keep", "nogc", "dev1first", "eager", "steps"):
            keep.append(e)
        print("engine on", d, flush=True)
        if mode == "steps":
            with torch.cuda.device(e.device):
                net = e.compiled(6, 320, 320)
                check(f"compiled {d}")
                from yolo_infer_b200.engine import letterbox_geometry
                fr = list(batch.to(e.device))
                e.preprocess_images(net, fr, [letterbox_geometry(320, 320, (640, 640), True)] * 6)
                check(f"letterbox {d}")
                e.forward(net)
                check(f"forward {d}")
                s = torch.cuda.Stream(e.device)
                with torch.cuda.stream(s):
                    e.forward(net)
                check(f"forward on side stream {d}")
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr, capture_error_mode="thread_local"):
                    e.forward(net)
                check(f"captured {d}")
                gr.replay()
                check(f"replayed {d}")
            continue
        r = e.predict(batch, conf=0.25, verbose=False, graph=(mode != "eager"))
        check(f"predict {d}")
        outs.append([x.boxes.data.cpu() for x in r])
    if outs:
        print("equal:", all(torch.equal(a, b) for a, b in zip(*outs)), flush=True)
except Exception:
    traceback.print_exc()
    sys.exit(1)
