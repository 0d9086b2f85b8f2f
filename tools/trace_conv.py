"""Debug: per-role timeline of CTA 0 of one conv launch (needs a -DY11_TRACE build: Y11_NVCC_EXTRA=-DY11_TRACE)."""
import ctypes as C
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from yolo_infer_b200 import _cabi as cabi  # noqa: E402
sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "tests"))
from gpu_utils import Ctx  # noqa: E402


def run(B, H, W, cin, cout, k, s, res=False, in_ct=None, in_off=0, label=""):
    ctx = Ctx()
    dev = ctx.dev
    trace = torch.zeros(4 * 4096, dtype=torch.int64, device=dev)
    ctx.lib.y11_debug_set_trace(C.c_void_p(trace.data_ptr()))
    x = torch.randn(B, H, W, in_ct or cin, device=dev).to(torch.bfloat16)
    Ho, Wo = (H + s - 1) // s, (W + s - 1) // s
    out = torch.zeros(B, Ho, Wo, cout, device=dev, dtype=torch.bfloat16)
    r = torch.randn(B, Ho, Wo, cout, device=dev).to(torch.bfloat16)
    w = (torch.randn(cout, k * k * cin, device=dev) / (cin * k * k) ** 0.5).to(torch.bfloat16)
    b = torch.randn(cout, device=dev)
    d = cabi.ConvDesc()
    d.inp = cabi.View(x.data_ptr(), x.shape[-1], in_off, cin)
    d.out = cabi.View(out.data_ptr(), cout, 0, cout)
    if res:
        d.res = cabi.View(r.data_ptr(), cout, 0, cout)
    d.w, d.bias = w.data_ptr(), b.data_ptr()
    d.B, d.Hin, d.Win, d.Hout, d.Wout = B, H, W, Ho, Wo
    d.k, d.stride, d.act, d.out_f32, d.impl = k, s, 1, 0, cabi.IMPL_TCGEN05
    p = ctx.plan()
    cabi.check(ctx.lib.y11_plan_add_conv(p, C.byref(d)), "add")
    for _ in range(3):
        cabi.check(ctx.lib.y11_plan_run(p, ctx.stream()))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    cabi.check(ctx.lib.y11_plan_run(p, ctx.stream()))
    e1.record()
    torch.cuda.synchronize()
    t = trace.cpu().view(4, 2048, 2)
    print(f"=== {label} B{B} {H}x{W} {cin}->{cout} k{k} s{s} res={res}: {e0.elapsed_time(e1) * 1e3:.1f} us")
    names = ["producer", "mma", "epi-leader", "epi-w3"]
    t0 = int(t[t[..., 1] > 0][:, 1].min())
    for role in range(4):
        ev = [(int(a), int(c) - t0) for a, c in t[role] if int(c) > 0]
        print(f"-- {names[role]}: {len(ev)} events")
        # steady-state window: events 40%..40%+36
        lo = int(len(ev) * 0.4)
        prev = None
        line = []
        for a, c in ev[lo:lo + 40]:
            line.append(f"{a}@{c}" + (f"(+{c - prev})" if prev is not None else ""))
            prev = c
        print("   " + " ".join(line))


def timeit(B, H, W, cin, cout, k, s, label="", act=1):
    ctx = Ctx()
    dev = ctx.dev
    x = torch.randn(B, H, W, cin, device=dev).to(torch.bfloat16)
    Ho, Wo = (H + s - 1) // s, (W + s - 1) // s
    out = torch.zeros(B, Ho, Wo, cout, device=dev, dtype=torch.bfloat16)
    w = (torch.randn(cout, k * k * cin, device=dev) / (cin * k * k) ** 0.5).to(torch.bfloat16)
    b = torch.randn(cout, device=dev)
    d = cabi.ConvDesc()
    d.inp = cabi.View(x.data_ptr(), cin, 0, cin)
    d.out = cabi.View(out.data_ptr(), cout, 0, cout)
    d.w, d.bias = w.data_ptr(), b.data_ptr()
    d.B, d.Hin, d.Win, d.Hout, d.Wout = B, H, W, Ho, Wo
    d.k, d.stride, d.act, d.out_f32, d.impl = k, s, act, 0, cabi.IMPL_TCGEN05
    p = ctx.plan()
    cabi.check(ctx.lib.y11_plan_add_conv(p, C.byref(d)), "add")
    junk = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ts = []
    for _ in range(6):
        junk.fill_(1)  # flush L2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        cabi.check(ctx.lib.y11_plan_run(p, ctx.stream()))
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    mb = (x.numel() + out.numel()) * 2 / 1e6
    t = sorted(ts)[len(ts) // 2]
    print(f"{label:28s} B{B} {H}x{W} {cin}->{cout} k{k}s{s}: {t:7.1f} us  {mb / t * 1e-3 * 1e3:7.1f} GB/s(algo)  [{mb:.0f} MB]")


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "time4":
    for cin in (16, 32, 48, 64, 96, 128):
        timeit(64, 160, 160, cin, 32, 1, 1, label="cout 32")
    for cin in (16, 32, 64, 128):
        timeit(64, 160, 160, cin, 64, 1, 1, label="cout 64")
    for cin in (16, 32, 64, 128):
        timeit(64, 160, 160, cin, 16, 1, 1, label="cout 16")
    sys.exit(0)

if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "time3":
    timeit(64, 160, 160, 32, 32, 1, 1, label="160x160 tile 32x4")
    timeit(100, 128, 128, 32, 32, 1, 1, label="128x128 tile 128x1")
    timeit(25, 256, 256, 32, 32, 1, 1, label="256x256 tile 128x1")
    timeit(400, 64, 64, 32, 32, 1, 1, label="64x64 tile 64x2")
    timeit(1600, 32, 32, 32, 32, 1, 1, label="32x32 tile 32x4")
    timeit(64, 160, 160, 32, 64, 1, 1, label="32->64")
    timeit(64, 160, 160, 64, 32, 1, 1, label="64->32")
    sys.exit(0)

if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "time2":
    for B in (64, 32, 16):
        timeit(B, 160, 160, 32, 32, 1, 1, label="1x1 act")
    timeit(64, 160, 160, 32, 32, 1, 1, label="1x1 noact", act=0)
    timeit(64, 160, 160, 16, 16, 1, 1, label="1x1 noact", act=0)
    timeit(64, 160, 160, 16, 16, 3, 1, label="3x3 halo noact", act=0)
    sys.exit(0)

if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "time":
    for (cin, cout) in [(32, 32), (64, 64), (128, 128), (48, 64), (96, 128), (32, 128), (128, 32), (16, 16)]:
        timeit(64, 160, 160, cin, cout, 1, 1, label="1x1")
    for (cin, cout) in [(16, 16), (32, 16), (16, 32), (64, 32), (32, 64), (64, 64)]:
        timeit(64, 160, 160, cin, cout, 3, 1, label="3x3")
    sys.exit(0)

if __name__ == "__main__":
    run(64, 160, 160, 32, 32, 1, 1, label="11n model.2.cv1")
    run(64, 160, 160, 16, 16, 3, 1, in_ct=48, in_off=16, label="11n model.2.m.0.cv1 (halo)")
    run(64, 160, 160, 48, 64, 1, 1, label="11n model.2.cv2")
    run(64, 160, 160, 96, 128, 1, 1, label="11s model.2.cv2")
