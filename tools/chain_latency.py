"""Per-layer cost of a dependent chain of small convs inside one CUDA graph (what the 20x20 / 40x40 part of the network looks
like to the GPU): python tools/chain_latency.py   (needs a B200)"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolo_infer_b200 import _cabi as cabi  # noqa: E402

lib = cabi.load()
dev = torch.device("cuda:0")
h = C.c_void_p()
cabi.check(lib.y11_create(C.byref(h), 0), "create")
st = torch.cuda.Stream()
for (B, H, W, c, k, n_layers) in [(64, 20, 20, 128, 1, 60), (64, 20, 20, 128, 3, 60), (64, 40, 40, 64, 1, 60), (64, 40, 40, 64, 3, 60),
                                  (64, 20, 20, 256, 3, 40), (1, 80, 80, 64, 3, 60), (1, 20, 20, 256, 1, 60)]:
    bufs = [torch.zeros((B, H, W, c), dtype=torch.bfloat16, device=dev) for _ in range(2)]
    w = (torch.randn(c, k * k * c, device=dev) / (k * k * c) ** 0.5).to(torch.bfloat16)
    bias = torch.zeros(c, device=dev)
    plan = C.c_void_p()
    cabi.check(lib.y11_plan_create(h, C.byref(plan)), "plan")
    for i in range(n_layers):
        d = cabi.ConvDesc()
        d.inp = cabi.View(bufs[i & 1].data_ptr(), c, 0, c)
        d.out = cabi.View(bufs[(i + 1) & 1].data_ptr(), c, 0, c)
        d.res = cabi.NULL_VIEW
        d.w, d.bias = w.data_ptr(), bias.data_ptr()
        d.B, d.Hin, d.Win, d.Hout, d.Wout = B, H, W, H, W
        d.k, d.stride, d.act, d.out_f32, d.impl, d.res_mode = k, 1, 1, 0, cabi.IMPL_TCGEN05, 0
        cabi.check(lib.y11_plan_add_conv(plan, C.byref(d)), "add")
    with torch.cuda.stream(st):
        cabi.check(lib.y11_plan_autotune(plan, C.c_void_p(st.cuda_stream), 4), "tune")
        cabi.check(lib.y11_plan_run(plan, C.c_void_p(st.cuda_stream)), "run")
        st.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            cabi.check(lib.y11_plan_run(plan, C.c_void_p(st.cuda_stream)), "run")
        for _ in range(3):
            g.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(10):
            g.replay()
        e1.record(st)
        st.synchronize()
    v = (C.c_int32 * 4)()
    lib.y11_plan_op_variant(plan, 0, v)
    flops = 2.0 * B * H * W * c * c * k * k
    us = e0.elapsed_time(e1) * 1e3 / 10 / n_layers
    print(f"B={B} {H}x{W} c={c} k={k}: {us:6.2f} us per layer in a graph chain ({flops / us / 1e6:7.1f} TFLOP/s), variant {list(v)}")
    lib.y11_plan_destroy(plan)
