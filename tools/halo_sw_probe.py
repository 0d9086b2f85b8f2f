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
