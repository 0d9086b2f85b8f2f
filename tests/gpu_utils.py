"""Helpers for the -m gpu tests: single-op plans through the C ABI."""
import ctypes as C

import torch

from yolo_infer_b200 import _cabi as cabi


class Ctx:
    def __init__(self, device="cuda:0"):
        self.lib = cabi.load()
        self.dev = torch.device(device)
        self.h = C.c_void_p()
        cabi.check(self.lib.y11_create(C.byref(self.h), self.dev.index or 0), "y11_create")
        self.keep = []

    def stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)

    def plan(self):
        p = C.c_void_p()
        cabi.check(self.lib.y11_plan_create(self.h, C.byref(p)), "plan_create")
        return p

    def run(self, p, repeats: int = 1):
        for _ in range(repeats):
            cabi.check(self.lib.y11_plan_run(p, self.stream()), "plan_run")
        torch.cuda.synchronize(self.dev)
        self.lib.y11_plan_destroy(p)

    def close(self):
        self.lib.y11_destroy(self.h)


def nhwc(x_nchw: torch.Tensor, c_total=None, c_off=0, dtype=torch.bfloat16):
    """NCHW fp32 -> NHWC buffer of c_total channels with the data at c_off (rest filled with junk)."""
    B, Cc, H, W = x_nchw.shape
    c_total = c_total or Cc
    buf = torch.full((B, H, W, c_total), 7.0, dtype=dtype, device=x_nchw.device)
    buf[..., c_off:c_off + Cc] = x_nchw.permute(0, 2, 3, 1).to(dtype)
    return buf


def conv_case(ctx: Ctx, B, H, W, cin, cout, k, stride, act, res=False, out_f32=False, in_off=0, in_extra=0, out_off=0,
              out_extra=0, impl=cabi.IMPL_TCGEN05, seed=0, res_mode=cabi.RES_POST, tune=None, repeats=1, return_variant=False):
    """Runs one conv op; returns (got NCHW fp32, want NCHW fp32 computed by torch fp32 on the same bf16 inputs)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    dev = ctx.dev
    x = torch.randn(B, cin, H, W, generator=g).to(dev).to(torch.bfloat16).float()
    w = (torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).to(dev).to(torch.bfloat16).float()
    bias = torch.randn(cout, generator=g).to(dev)
    Ho, Wo = (H + stride - 1) // stride, (W + stride - 1) // stride
    pre = res and res_mode == cabi.RES_PRE_UP2     # residual = half-resolution map, nearest-upsampled, added BEFORE the activation
    rh, rw = (Ho // 2, Wo // 2) if pre else (Ho, Wo)
    r = torch.randn(B, cout, rh, rw, generator=g).to(dev).to(torch.bfloat16).float() if res else None
    want = torch.nn.functional.conv2d(x, w, bias, stride=stride, padding=k // 2)
    if pre:
        want = want + torch.nn.functional.interpolate(r, scale_factor=2, mode="nearest")
    if act:
        want = torch.nn.functional.silu(want)
    if res and not pre:
        want = want + r
    xin = nhwc(x, cin + in_off + in_extra, in_off)
    out_dtype = torch.float32 if out_f32 else torch.bfloat16
    out = torch.full((B, Ho, Wo, cout + out_off + out_extra), -3.0, dtype=out_dtype, device=dev)
    wp = w.permute(0, 2, 3, 1).reshape(cout, -1).to(torch.bfloat16).contiguous()
    d = cabi.ConvDesc()
    d.inp = cabi.View(xin.data_ptr(), xin.shape[-1], in_off, cin)
    d.out = cabi.View(out.data_ptr(), out.shape[-1], out_off, cout)
    if res:
        rb = nhwc(r)
        d.res = cabi.View(rb.data_ptr(), cout, 0, cout)
    d.w, d.bias = wp.data_ptr(), bias.data_ptr()
    d.B, d.Hin, d.Win, d.Hout, d.Wout = B, H, W, Ho, Wo
    d.k, d.stride, d.act, d.out_f32, d.impl = k, stride, int(act), int(out_f32), impl
    d.res_mode = res_mode
    p = ctx.plan()
    if tune is None:
        cabi.check(ctx.lib.y11_plan_add_conv(p, C.byref(d)), "add_conv")
    else:   # explicit launch variant (lsu, epi_warp, ctas_per_sm, bn_max) instead of the per-layer heuristic
        cabi.check(ctx.lib.y11_plan_add_conv_tuned(p, C.byref(d), *tune), "add_conv_tuned")
    var = (C.c_int32 * 4)()
    cabi.check(ctx.lib.y11_plan_op_variant(p, 0, var), "op_variant")
    ctx.run(p, repeats)
    got = out[..., out_off:out_off + cout].float().permute(0, 3, 1, 2)
    untouched = torch.cat((out[..., :out_off].flatten(), out[..., out_off + cout:].flatten()))
    assert torch.all(untouched == -3.0), "conv wrote outside its channel slice"
    if return_variant:
        return got, tuple(var)       # (lsu, epi_warp bits, ctas_per_sm, bn) actually used
    return got, want
