"""Experiment: 3x3 stride-1 cin = 64 halo conv with the halo tile in the 128-byte-swizzled layout (Y11_HALO_SW=1/2) vs torch fp32."""
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from gpu_utils import Ctx, conv_case  # noqa: E402

ctx = Ctx()
for (B, H, W, cout) in [(2, 32, 40, 64), (4, 80, 80, 32), (3, 48, 56, 64)]:
    got, want = conv_case(ctx, B, H, W, 64, cout, 3, 1, True, tune=(1, 0, 2, 128))
    err = (got - want).abs().max().item()
    rel = ((got - want).norm() / want.norm()).item()
    print(f"HALO_SW={os.environ.get('Y11_HALO_SW', '0')} B{B} {H}x{W} 64->{cout}: max abs err {err:.4g} rel L2 {rel:.3g}", flush=True)


def timeit(B, H, W, cin, cout, tune, label=""):
    import ctypes as C
    from yolo_infer_b200 import _cabi as cabi
    dev = ctx.dev
    x = torch.randn(B, H, W, cin, device=dev).to(torch.bfloat16)
    out = torch.zeros(B, H, W, cout, device=dev, dtype=torch.bfloat16)
    w = (torch.randn(cout, 9 * cin, device=dev) / (cin * 9) ** 0.5).to(torch.bfloat16)
    b = torch.randn(cout, device=dev)
    d = cabi.ConvDesc()
    d.inp = cabi.View(x.data_ptr(), cin, 0, cin)
    d.out = cabi.View(out.data_ptr(), cout, 0, cout)
    d.w, d.bias = w.data_ptr(), b.data_ptr()
    d.B, d.Hin, d.Win, d.Hout, d.Wout = B, H, W, H, W
    d.k, d.stride, d.act, d.out_f32, d.impl = 3, 1, 1, 0, cabi.IMPL_TCGEN05
    p = ctx.plan()
    cabi.check(ctx.lib.y11_plan_add_conv_tuned(p, C.byref(d), *tune), "add")
    var = (C.c_int32 * 4)()
    cabi.check(ctx.lib.y11_plan_op_variant(p, 0, var), "op_variant")
    junk = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ts = []
    for _ in range(7):
        junk.fill_(1)  # flush L2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        cabi.check(ctx.lib.y11_plan_run(p, ctx.stream()))
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    t = sorted(ts)[len(ts) // 2]
    print(f"TIME HALO_SW={os.environ.get('Y11_HALO_SW', '0')} {label:10s} B{B} {H}x{W} {cin}->{cout} tune={tune} variant={tuple(var)}: {t:7.1f} us", flush=True)


if "time" in sys.argv:
    for tune in [(1, 0, -1, -1), (1, 1, -1, -1), (1, 2, 2, -1), (0, 0, -1, -1), (0, 8, 2, -1)]:
        timeit(64, 80, 80, 64, 64, tune, "P3 64->64")
    for tune in [(1, 0, -1, -1), (1, 2, 2, -1), (0, 0, -1, -1)]:
        timeit(64, 80, 80, 64, 32, tune, "P3 64->32")
        timeit(64, 40, 40, 64, 64, tune, "P4 64->64")

if "stream" in sys.argv:
    for (B, H, W, cin, cout, res) in [(2, 32, 40, 64, 64, False), (2, 40, 36, 128, 64, False), (2, 32, 48, 128, 128, True), (2, 32, 32, 64, 256, False)]:
        for tune in [(3, 0, 2, -1), (3, 1, 2, -1), (3, 2, 2, -1)]:
            got, want = conv_case(ctx, B, H, W, cin, cout, 3, 1, True, res, tune=tune)
            base, _ = conv_case(ctx, B, H, W, cin, cout, 3, 1, True, res, tune=(0, 0, 2, -1))
            print(f"STREAM B{B} {H}x{W} {cin}->{cout} res={res} tune={tune}: rel L2 {((got - want).norm() / want.norm()).item():.3g} bit-identical to tap mode: {torch.equal(got, base)}", flush=True)

    def t3(B, H, W, cin, cout, tunes, label):
        for tune in tunes:
            timeit(B, H, W, cin, cout, tune, label)
    t3(64, 80, 80, 128, 64, [(0, 0, -1, -1), (0, 8, 2, -1), (0, 10, 2, -1), (3, 0, 2, -1), (3, 2, 2, -1), (3, 0, 1, -1), (3, 1, 2, -1)], "cv2.0.0")
    t3(64, 40, 40, 128, 64, [(0, 0, -1, -1), (0, 8, 2, -1), (3, 0, 2, -1), (3, 2, 2, -1)], "13.m.0.cv1")
    t3(64, 40, 40, 64, 128, [(0, 0, -1, -1), (0, 8, 2, -1), (3, 0, 2, -1), (3, 2, 2, -1)], "13.m.0.cv2")
    t3(64, 40, 40, 64, 64, [(0, 0, -1, -1), (2, 2, 2, -1), (3, 0, 2, -1), (3, 2, 2, -1)], "6.m.0.m.x")
