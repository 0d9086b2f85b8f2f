"""Where does YOLO.predict(pinned uint8 batch) spend its time?  (debug helper)"""
import sys, time
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from yolo_infer_b200 import topology as T
from yolo_infer_b200.engine import YOLO

scale = sys.argv[1] if len(sys.argv) > 1 else "n"
B, S = 64, 640
eng = YOLO.from_state_dict(T.synthetic_state_dict(scale, 80, seed=0), scale).to("cuda:0")
from yolo_infer_b200.synth import condition_synthetic_weights
condition_synthetic_weights(eng, (S, S), batch=2, seed=0)
g = torch.Generator().manual_seed(0)
host = [torch.randint(0, 256, (B, S, S, 3), generator=g, dtype=torch.uint8).pin_memory() for _ in range(3)]
for i in range(4):
    eng.predict(host[i % 3], verbose=False)
torch.cuda.synchronize()
pipe = eng.pipeline(B, S, S, S, True, 0.25, 0.7, 300)
def t(fn, n=10):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(n): fn(i)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
print("H2D copy only        %.3f ms" % t(lambda i: pipe.frames.copy_(host[i % 3], non_blocking=True)))
print("graph replay only    %.3f ms" % t(lambda i: [g.replay() for g in pipe.graphs]))
print("copy+replay          %.3f ms" % t(lambda i: pipe.run(host[i % 3])))
def fetch(i):
    with torch.inference_mode():
        eng._fetch_results(*pipe.run(host[i % 3])[:2])
print("copy+replay+fetch    %.3f ms" % t(fetch))
print("predict              %.3f ms" % t(lambda i: eng.predict(host[i % 3], verbose=False)))
print("predict + r.cpu()    %.3f ms" % t(lambda i: [r.cpu().boxes.data for r in eng.predict(host[i % 3], verbose=False)]))

# ---- postprocess alone (heads as left by the last pass) ----
net = pipe.net
import ctypes as C
from yolo_infer_b200 import _cabi as cabi
def post(i):
    with torch.inference_mode():
        eng.postprocess(net, pipe.scale_rows, 0.25, 0.7, 300)
print("postprocess only     %.3f ms  (mean candidates %.0f)" % (t(post, 20), float(pipe.ncand.float().mean())))
lib = eng._lib
hd = net.head_desc()
