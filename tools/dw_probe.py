"""Times single depthwise 3x3 ops through the C ABI (plan of ONE op, CUDA events, rotating buffers larger than L2).
`python tools/dw_probe.py [B H W c ...]`; with a probe build (tools/build_variant.py dwprobe --src dwconv_tma.cu -DY11_DW_PROBE,
Y11_LIB=.../liby11_dwprobe.so) the knobs Y11_DW_DBG (1: loads only, 2: arithmetic + stores only) and Y11_DW_STAGES apply."""
import ctypes as C
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from gpu_utils import Ctx  # noqa: E402
from yolo_infer_b200 import _cabi as cabi  # noqa: E402


def probe(ctx, B, H, W, c, nbuf=3, iters=30):
    dev = ctx.dev
    xs = [torch.randn(B, H, W, c, device=dev).to(torch.bfloat16) for _ in range(nbuf)]
    outs = [torch.zeros(B, H, W, c, device=dev, dtype=torch.bfloat16) for _ in range(nbuf)]
    w = (torch.randn(9, c, device=dev) / 3).to(torch.bfloat16)
    b = torch.randn(c, device=dev)
    plans = []
    for x, o in zip(xs, outs):
        d = cabi.DwConvDesc(cabi.View(x.data_ptr(), c, 0, c), cabi.View(o.data_ptr(), c, 0, c), cabi.NULL_VIEW, w.data_ptr(), b.data_ptr(),
                            B, H, W, 1)
        p = ctx.plan()
        cabi.check(ctx.lib.y11_plan_add_dwconv(p, C.byref(d)))
        plans.append(p)
    s = ctx.stream()
    for p in plans:
        cabi.check(ctx.lib.y11_plan_run(p, s))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        cabi.check(ctx.lib.y11_plan_run(plans[i % nbuf], s))
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1000 / iters
    gb = 2 * B * H * W * c * 2 / 1e9
    return us, gb / (us * 1e-6) / 1e3


if __name__ == "__main__":
    a = [int(v) for v in sys.argv[1:]]
    shapes = [tuple(a[i:i + 4]) for i in range(0, len(a), 4)] or [(64, 80, 80, 128), (64, 40, 40, 256), (64, 40, 40, 128)]
    ctx = Ctx()
    for sh in shapes:
        us, tbs = probe(ctx, *sh)
        print(f"dw {sh} dbg={os.environ.get('Y11_DW_DBG', '0')} stages={os.environ.get('Y11_DW_STAGES', '-')}: {us:.1f} us  {tbs:.2f} TB/s (read+write)", flush=True)
