"""GPU (-m gpu): per-kernel parity of liby11_b200 (through the C ABI) against the oracle / a torch fp32 reference.

Tolerances: bf16 outputs are compared with |got-want| <= 1e-2*|want| + 2e-2 (bf16 has 8 significant bits:
rel 2^-8 = 3.9e-3 rounding on the stored output, fp32 accumulation inside); fp32 outputs with 2e-3 abs/rel
(inputs are bf16-exact, so only summation order differs).  Integer/index results (letterbox u8, NMS keep) are bit-exact.
"""
import ctypes as C

import cv2
import numpy as np
import pytest
import torch
import torchvision

pytestmark = pytest.mark.gpu

from oracle import pipeline_ref as P  # noqa: E402
from oracle import yolo11_ref as R  # noqa: E402
from yolo_infer_b200 import _cabi as cabi  # noqa: E402
from yolo_infer_b200.engine import letterbox_geometry  # noqa: E402

from gpu_utils import Ctx, conv_case, nhwc  # noqa: E402


@pytest.fixture(scope="module")
def ctx():
    c = Ctx()
    yield c
    c.close()


def close_bf16(got, want):
    err = (got - want).abs()
    tol = 1e-2 * want.abs() + 2e-2
    assert torch.all(err <= tol), f"max err {err.max().item():.4g} at want={want.flatten()[err.argmax()].item():.4g}"


# ------------------------------------------------------------------------------------------- letterbox
def run_letterbox(ctx, imgs, new_shape=(640, 640), rect=True, u8=True):
    shapes = [im.shape[:2] for im in imgs]
    auto = rect and len(set(shapes)) == 1
    geoms = [letterbox_geometry(h, w, new_shape, auto) for h, w in shapes]
    H, W = geoms[0][4], geoms[0][5]
    frames = [torch.from_numpy(im).to(ctx.dev) for im in imgs]
    arr = (cabi.Image * len(imgs))()
    for i, (f, g) in enumerate(zip(frames, geoms)):
        arr[i] = cabi.Image(f.data_ptr(), f.shape[0], f.shape[1], f.stride(0), g[1], g[0], g[2], g[3])
    desc = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(ctx.dev)
    if u8:
        out = torch.zeros((len(imgs), H, W, 3), dtype=torch.uint8, device=ctx.dev)
        cabi.check(ctx.lib.y11_letterbox_u8(ctx.h, desc.data_ptr(), len(imgs), H, W, out.data_ptr(), ctx.stream()))
    else:
        out = torch.zeros((len(imgs), H, W, 3), dtype=torch.bfloat16, device=ctx.dev)
        cabi.check(ctx.lib.y11_letterbox(ctx.h, desc.data_ptr(), len(imgs), H, W, out.data_ptr(), ctx.stream()))
    torch.cuda.synchronize()
    return out, auto


@pytest.mark.parametrize("h,w", [(853, 1280), (720, 1280), (1080, 1920), (640, 640), (480, 640), (300, 400), (333, 517),
                                 (1280, 1280), (100, 37), (641, 1283), (1281, 641)])
@pytest.mark.parametrize("rect", [True, False])
def test_letterbox_u8_bit_exact_vs_cv2(ctx, h, w, rect):
    rng = np.random.default_rng(h + w)
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    out, auto = run_letterbox(ctx, [img], rect=rect)
    ref = P.letterbox(img, (640, 640), auto=auto)
    assert out.shape[1:3] == ref.shape[:2]
    assert np.array_equal(out[0].cpu().numpy(), ref)


def test_letterbox_bf16_matches_reference_preprocess(ctx):
    rng = np.random.default_rng(5)
    imgs = [rng.integers(0, 256, (720, 1280, 3), dtype=np.uint8) for _ in range(3)]
    out, _ = run_letterbox(ctx, imgs, u8=False)
    ref = P.preprocess(imgs, (640, 640), rect=True)                      # fp32 NCHW RGB /255
    want = ref.permute(0, 2, 3, 1).to(torch.bfloat16)                    # bf16 RNE of the exact fp32 value
    assert torch.equal(out.cpu(), want)


def test_mixed_shapes_batch_goes_square(ctx):
    rng = np.random.default_rng(6)
    imgs = [rng.integers(0, 256, s + (3,), dtype=np.uint8) for s in [(480, 640), (853, 1280)]]
    out, auto = run_letterbox(ctx, imgs)
    assert not auto and out.shape == (2, 640, 640, 3)
    for i, im in enumerate(imgs):
        assert np.array_equal(out[i].cpu().numpy(), P.letterbox(im, (640, 640), auto=False))


def test_golden_letterbox_fixture(ctx):
    from pathlib import Path
    g = np.load(str(Path(__file__).parent / "golden" / "letterbox_small.npz"))
    out, _ = run_letterbox(ctx, [g["img"]], new_shape=(64, 64), rect=True)
    assert np.array_equal(out[0].cpu().numpy(), g["out_rect"])
    out, _ = run_letterbox(ctx, [g["img"]], new_shape=(64, 64), rect=False)
    assert np.array_equal(out[0].cpu().numpy(), g["out_square"])


def test_tensor_source_conversion(ctx):
    xc = torch.rand(2, 3, 64, 96, generator=torch.Generator().manual_seed(5)) * 255
    x = xc.to(ctx.dev)
    out = torch.zeros((2, 64, 96, 3), dtype=torch.bfloat16, device=ctx.dev)
    cabi.check(ctx.lib.y11_nchw_f32_to_nhwc_bf16(ctx.h, x.data_ptr(), 2, 64, 96, 255.0, out.data_ptr(), ctx.stream()))
    torch.cuda.synchronize()
    # reference on the CPU: the reference path divides exactly (torch's CUDA scalar division multiplies by the reciprocal,
    # 1 ulp off now and then, which made this comparison flaky when it was done on the GPU with unseeded data)
    assert torch.equal(out.cpu(), (xc / 255.0).permute(0, 2, 3, 1).to(torch.bfloat16))


# ------------------------------------------------------------------------------------------- conv (tcgen05)
CONV_CASES = [
    # B, H, W, cin, cout, k, s, act, res
    (2, 16, 16, 64, 64, 1, 1, True, False),      # 1x1, SW128
    (2, 16, 16, 32, 48, 1, 1, True, False),      # SW64, N=48
    (2, 16, 16, 16, 16, 1, 1, False, False),     # SW32, N=16, no act
    (1, 20, 20, 128, 256, 1, 1, True, True),     # 2 K chunks, 2 N tiles, residual, 20x20 ragged tile
    (3, 20, 20, 64, 64, 3, 1, True, True),       # 3x3 s1 + residual (Bottleneck)
    (2, 24, 40, 32, 64, 3, 1, True, False),      # 3x3, SW64, non-square
    (2, 16, 16, 16, 32, 3, 1, True, False),      # 3x3, SW32
    (2, 32, 32, 64, 128, 3, 2, True, False),     # stride 2
    (1, 28, 40, 32, 64, 3, 2, True, False),      # stride 2, rect shape
    (2, 14, 20, 256, 256, 3, 1, True, False),    # deep K (36 stages worth), P5-like rect
    (8, 20, 20, 192, 384, 1, 1, True, False),    # cin 192 = 3 chunks, 3 N tiles, (4,4,8) tiling
    (1, 80, 80, 64, 80, 1, 1, False, False),     # cls logits shape: N=80
    (2, 8, 8, 512, 512, 1, 1, True, False),      # 4 N tiles x 8 K chunks
    # halo mode (3x3 s1, cin in {16,32,64}, maps >= 32x32): cp.async halo tile + un-swizzled shifted descriptors
    (2, 32, 32, 16, 16, 3, 1, True, False),      # exact 8x16 tiles
    (1, 48, 40, 32, 16, 3, 1, True, True),       # residual
    (2, 40, 36, 16, 32, 3, 1, True, False),      # ragged in both directions (clipped stores, zero-filled halo)
    (1, 32, 64, 64, 32, 3, 1, False, False),     # cin 64 (four K steps per tap), no act
    (3, 80, 80, 32, 64, 3, 1, True, True),       # many tiles per CTA: ring wrap-around, accumulator double buffering
]


@pytest.mark.parametrize("case", CONV_CASES, ids=lambda c: "x".join(map(str, c)))
def test_conv_tcgen05_vs_torch(ctx, case):
    B, H, W, cin, cout, k, s, act, res = case
    got, want = conv_case(ctx, B, H, W, cin, cout, k, s, act, res)
    close_bf16(got, want)


def test_conv_tcgen05_channel_slices_and_f32_out(ctx):
    got, want = conv_case(ctx, 2, 16, 24, 64, 64, 3, 1, True, in_off=32, in_extra=16, out_off=16, out_extra=32)
    close_bf16(got, want)
    got, want = conv_case(ctx, 2, 16, 24, 64, 80, 1, 1, False, out_f32=True, out_off=64)
    assert torch.allclose(got, want, rtol=2e-3, atol=2e-3)
    # halo mode reading a channel slice of a wider buffer (C3k2's y buffer) and writing into another slice
    got, want = conv_case(ctx, 2, 32, 40, 32, 16, 3, 1, True, in_off=32, in_extra=16, out_off=16, out_extra=32)
    close_bf16(got, want)


@pytest.mark.parametrize("case", [
    (2, 40, 40, 128, 128, 1, True),     # TMA producer, 40x40 (model.13.cv1#skip of YOLO11n-like sizes)
    (1, 80, 80, 64, 64, 1, True),       # 80x80
    (3, 28, 40, 32, 64, 1, True),       # cp.async producer (cin 32), rect map
    (2, 16, 24, 256, 256, 1, False),    # 2 N tiles, no activation
    (1, 32, 32, 32, 32, 3, True),       # 3x3 halo mode with the pre-activation term
], ids=lambda c: "x".join(map(str, c)))
def test_conv_pre_activation_upsampled_term(ctx, case):
    """Y11_RES_PRE_UP2: out = act(conv(x) + b + up2(r)), r at half resolution - the folded Upsample+Concat form - on the
    tcgen05 kernel, and the CUDA-core cross-check kernel agrees."""
    B, H, W, cin, cout, k, act = case
    got, want = conv_case(ctx, B, H, W, cin, cout, k, 1, act, res=True, res_mode=cabi.RES_PRE_UP2)
    close_bf16(got, want)
    got2, want2 = conv_case(ctx, B, H, W, cin, cout, k, 1, act, res=True, res_mode=cabi.RES_PRE_UP2, impl=cabi.IMPL_SIMT_DEBUG)
    close_bf16(got2, want2)


VARIANT_SHAPES = [
    (1, 80, 80, 64, 80, 1, 1, False, False),     # N = 80: a 64-channel fat chunk + a clipped 16-channel one
    (2, 40, 40, 128, 256, 1, 1, True, True),     # two N tiles of 128 = 2 fat chunks each, residual
    # B, H, W, cin, cout, k, s, act, res
    (2, 32, 32, 32, 64, 1, 1, True, False),      # LSU-eligible 1x1
    (2, 40, 36, 32, 32, 3, 1, True, True),       # LSU-eligible 3x3 (halo), ragged, residual
    (4, 20, 20, 128, 128, 1, 1, True, True),     # TMA 1x1, (4,4,8) tiles
    (2, 16, 16, 64, 128, 3, 1, True, False),     # TMA 3x3
    (2, 32, 32, 64, 128, 3, 2, True, False),     # stride 2
    (2, 40, 36, 64, 64, 3, 1, True, True),       # TMA-halo eligible (3x3 stride 1, cin = 64): ragged tiles, residual
    (3, 32, 32, 64, 32, 3, 1, True, False),      # TMA-halo eligible, N = 32
    (2, 40, 36, 128, 64, 3, 1, True, False),     # halo-stream: two 64-channel halo chunks, ragged tiles
    (2, 32, 48, 128, 128, 3, 1, True, True),     # halo-stream, residual
    (2, 32, 32, 64, 256, 3, 1, False, False),    # halo-stream, one chunk, two N tiles
]


@pytest.mark.parametrize("shape", VARIANT_SHAPES, ids=lambda c: "x".join(map(str, c)))
def test_conv_launch_variants_are_bit_identical(ctx, shape):
    """Every launch variant the plan autotuner may pick (producer TMA | cp.async, epilogue CTA-wide | warp-independent,
    2 | 3 CTAs per SM, narrower N tiles) matches torch and is BIT-identical to the default variant."""
    B, H, W, cin, cout, k, s_, act, res = shape
    base, want = conv_case(ctx, B, H, W, cin, cout, k, s_, act, res)
    close_bf16(base, want)
    n = 0
    halo_tma = k == 3 and s_ == 1 and cin == 64 and cout <= 64 and H >= 32 and W >= 32   # lsu = 2: one TMA box load per halo tile, resident weights
    hstream = k == 3 and s_ == 1 and cin in (64, 128) and H >= 32 and W >= 32             # lsu = 3: TMA halo + streamed weights
    for lsu in (0, 1) + ((2,) if halo_tma else ()) + ((3,) if hstream else ()):
        for ew in (0, 1, 2, 3, 4, 6):     # bit 0: per-warp epilogue, bit 1: fat epilogue (64-channel chunks), bit 2: resident weights
            for cps in (2, 3):
                for bn in (-1, 64, 32):
                    got, _ = conv_case(ctx, B, H, W, cin, cout, k, s_, act, res, tune=(lsu, ew, cps, bn))
                    assert torch.equal(got, base), (lsu, ew, cps, bn, (got - base).abs().max().item())
                    n += 1
    assert n == 72 + 36 * (int(halo_tma) + int(hstream))


PAIR_SHAPES = [
    # B, H, W, cin, cout, k, s, act, res
    (4, 20, 20, 128, 256, 1, 1, True, True),     # 1x1, BN 256 (forced), residual; (4,4,8) tiles, 13 M tiles: odd -> phantom peer tile
    (2, 40, 40, 256, 256, 3, 1, True, False),    # 3x3 long K, BN 256: 25 M tiles
    (2, 40, 40, 128, 128, 3, 2, True, False),    # stride 2, BN 128
    (3, 24, 20, 64, 128, 3, 1, True, True),      # ragged tiles, BN 128 / 64
    (2, 16, 16, 64, 64, 1, 1, False, False),     # BN 64, fp32-free plain
]


@pytest.mark.parametrize("shape", PAIR_SHAPES, ids=lambda c: "x".join(map(str, c)))
def test_conv_cta_pair_variant_is_bit_identical(ctx, shape):
    """tcgen05.mma.cta_group::2 variant (conv_tc_kernel_pair: a cluster of two CTAs computes a 256-row tile, each CTA holding
    half of the weight tile): matches torch and is BIT-identical to the default 1-CTA variant, for every N tile it accepts."""
    B, H, W, cin, cout, k, s_, act, res = shape
    base, want = conv_case(ctx, B, H, W, cin, cout, k, s_, act, res)
    close_bf16(base, want)
    n = 0
    for ew in (8, 10):              # bit 3: pair kernel; bit 1: its fat epilogue (conv_tc_kernel_pair_fat)
        for cps in (1, 2):
            for bn in (-1, 256, 128, 64):
                got, var = conv_case(ctx, B, H, W, cin, cout, k, s_, act, res, tune=(0, ew, cps, bn), repeats=2, return_variant=True)
                assert torch.equal(got, base), (ew, cps, bn, (got - base).abs().max().item())
                n += 1 if (var[1] & 8) else 0
    assert n >= 8, "the pair kernel was never selected"


def test_conv_2x2_space_to_depth_form(ctx):
    """k = 2 (taps {-1,0}^2, top/left zero padding): what a 3x3 stride-2 conv becomes on a space-to-depth input.  Checked
    (a) as a plain 2x2 conv against torch and (b) end to end: stem-style s2d repacking of a 3x3/2 conv == the 3x3/2 conv."""
    g = torch.Generator().manual_seed(3)
    dev = ctx.dev
    for (B, H, W, cin, cout) in [(2, 16, 24, 64, 32), (1, 40, 40, 128, 64)]:
        x = torch.randn(B, cin, H, W, generator=g).to(dev).to(torch.bfloat16).float()
        w = (torch.randn(cout, cin, 2, 2, generator=g) / (cin * 4) ** 0.5).to(dev).to(torch.bfloat16).float()
        bias = torch.randn(cout, generator=g).to(dev)
        want = torch.nn.functional.silu(torch.nn.functional.conv2d(torch.nn.functional.pad(x, (1, 0, 1, 0)), w, bias))
        xin = nhwc(x)
        out = torch.zeros((B, H, W, cout), dtype=torch.bfloat16, device=dev)
        wp = w.permute(0, 2, 3, 1).reshape(cout, -1).to(torch.bfloat16).contiguous()
        d = cabi.ConvDesc()
        d.inp, d.out = cabi.View(xin.data_ptr(), cin, 0, cin), cabi.View(out.data_ptr(), cout, 0, cout)
        d.w, d.bias = wp.data_ptr(), bias.data_ptr()
        d.B, d.Hin, d.Win, d.Hout, d.Wout = B, H, W, H, W
        d.k, d.stride, d.act, d.out_f32, d.impl = 2, 1, 1, 0, cabi.IMPL_TCGEN05
        p = ctx.plan()
        cabi.check(ctx.lib.y11_plan_add_conv(p, C.byref(d)), "add_conv")
        ctx.run(p)
        close_bf16(out.float().permute(0, 3, 1, 2), want)
    # (b) 3x3 stride-2 conv on [2,16,32,48] == 2x2 conv on its space-to-depth form with repacked weights
    B, c, H, W, cout = 2, 16, 32, 48, 32
    x = torch.randn(B, c, H, W, generator=g).to(dev).to(torch.bfloat16).float()
    w3 = (torch.randn(cout, c, 3, 3, generator=g) / (c * 9) ** 0.5).to(dev).to(torch.bfloat16).float()
    bias = torch.randn(cout, generator=g).to(dev)
    want = torch.nn.functional.conv2d(x, w3, bias, stride=2, padding=1)
    xs = x.view(B, c, H // 2, 2, W // 2, 2).permute(0, 2, 4, 3, 5, 1).reshape(B, H // 2, W // 2, 4 * c).to(torch.bfloat16).contiguous()
    wp = torch.zeros(cout, 2, 2, 4, c, device=dev)
    for ty in range(2):
        for tx in range(2):
            for dy in range(2):
                for dx in range(2):
                    kh, kw = 2 * (ty - 1) + dy + 1, 2 * (tx - 1) + dx + 1
                    if 0 <= kh <= 2 and 0 <= kw <= 2:
                        wp[:, ty, tx, dy * 2 + dx, :] = w3[:, :, kh, kw]
    wp = wp.view(cout, -1).to(torch.bfloat16).contiguous()
    out = torch.zeros((B, H // 2, W // 2, cout), dtype=torch.bfloat16, device=dev)
    d = cabi.ConvDesc()
    d.inp, d.out = cabi.View(xs.data_ptr(), 4 * c, 0, 4 * c), cabi.View(out.data_ptr(), cout, 0, cout)
    d.w, d.bias = wp.data_ptr(), bias.data_ptr()
    d.B, d.Hin, d.Win, d.Hout, d.Wout = B, H // 2, W // 2, H // 2, W // 2
    d.k, d.stride, d.act, d.out_f32, d.impl = 2, 1, 0, 0, cabi.IMPL_TCGEN05
    p = ctx.plan()
    cabi.check(ctx.lib.y11_plan_add_conv(p, C.byref(d)), "add_conv")
    ctx.run(p)
    close_bf16(out.float().permute(0, 3, 1, 2), want)


def test_conv_simt_debug_agrees(ctx):
    got, want = conv_case(ctx, 1, 12, 12, 32, 32, 3, 2, True, res=False, impl=cabi.IMPL_SIMT_DEBUG)
    close_bf16(got, want)


# ------------------------------------------------------------------------------------------- CUDA-core ops
@pytest.mark.parametrize("H,W", [(64, 96), (32, 608), (96, 640)])   # Wout 48 (ragged tile), 304 (ragged), 320 (two full tiles)
def test_stem(ctx, H, W):
    for cout in (16, 32, 64, 96):
        g = torch.Generator().manual_seed(cout)
        x = torch.rand(2, 3, H, W, generator=g).to(ctx.dev).to(torch.bfloat16).float()
        w = (torch.randn(cout, 3, 3, 3, generator=g) / 27 ** 0.5).to(ctx.dev).to(torch.bfloat16).float()
        b = torch.randn(cout, generator=g).to(ctx.dev)
        want = torch.nn.functional.silu(torch.nn.functional.conv2d(x, w, b, stride=2, padding=1))
        xin = x.permute(0, 2, 3, 1).to(torch.bfloat16).contiguous()
        out = torch.zeros((2, H // 2, W // 2, cout), dtype=torch.bfloat16, device=ctx.dev)
        wp = w.permute(0, 2, 3, 1).reshape(cout, 27).to(torch.bfloat16).contiguous()
        d = cabi.StemDesc(xin.data_ptr(), cabi.View(out.data_ptr(), cout, 0, cout), wp.data_ptr(), b.data_ptr(), 2, H, W, H // 2, W // 2, 0)
        p = ctx.plan()
        cabi.check(ctx.lib.y11_plan_add_stem(p, C.byref(d)))
        ctx.run(p)
        close_bf16(out.float().permute(0, 3, 1, 2), want)
        # space-to-depth output: [H/4, W/4, 4*cout], block dy*2+dx of pixel (y, x) = output pixel (2y+dy, 2x+dx): identical values
        out2 = torch.zeros((2, H // 4, W // 4, 4 * cout), dtype=torch.bfloat16, device=ctx.dev)
        d = cabi.StemDesc(xin.data_ptr(), cabi.View(out2.data_ptr(), 4 * cout, 0, 4 * cout), wp.data_ptr(), b.data_ptr(), 2, H, W, H // 2, W // 2, 1)
        p = ctx.plan()
        cabi.check(ctx.lib.y11_plan_add_stem(p, C.byref(d)))
        ctx.run(p)
        back = out2.view(2, H // 4, W // 4, 2, 2, cout).permute(0, 1, 3, 2, 4, 5).reshape(2, H // 2, W // 2, cout)
        assert torch.equal(back, out)


@pytest.mark.parametrize("H,W", [(64, 96), (32, 608), (96, 640), (128, 72)])
def test_stem_reading_uint8_frames_equals_letterbox_then_stem(ctx, H, W):
    """u8_src mode (frames already at network resolution): the stem converts BGR uint8 -> RGB bf16 /255 while staging its
    input rows; the result is BIT-identical to y11_letterbox followed by the bf16 stem (plain and space-to-depth output,
    ragged tiles, two images), and switching the source back restores the bf16 path."""
    g = torch.Generator().manual_seed(H + W)
    B = 2
    frames = torch.randint(0, 256, (B, H, W, 3), generator=g, dtype=torch.uint8).to(ctx.dev)
    arr = (cabi.Image * B)()
    for i in range(B):
        arr[i] = cabi.Image(frames[i].data_ptr(), H, W, frames[i].stride(0), H, W, 0, 0)
    desc = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(ctx.dev)
    xin = torch.zeros((B, H, W, 3), dtype=torch.bfloat16, device=ctx.dev)
    cabi.check(ctx.lib.y11_letterbox(ctx.h, desc.data_ptr(), B, H, W, xin.data_ptr(), ctx.stream()))
    for cout, s2d in ((16, 1), (32, 0), (64, 1), (96, 0)):
        w = (torch.randn(cout, 27, generator=g) / 27 ** 0.5).to(ctx.dev).to(torch.bfloat16).contiguous()
        b = torch.randn(cout, generator=g).to(ctx.dev)
        shape = (B, H // 4, W // 4, 4 * cout) if s2d else (B, H // 2, W // 2, cout)
        outs = []
        p = ctx.plan()
        out = torch.zeros(shape, dtype=torch.bfloat16, device=ctx.dev)
        d = cabi.StemDesc(xin.data_ptr(), cabi.View(out.data_ptr(), shape[-1], 0, shape[-1]), w.data_ptr(), b.data_ptr(), B, H, W, H // 2, W // 2, s2d)
        cabi.check(ctx.lib.y11_plan_add_stem(p, C.byref(d)))
        for source in (None, desc.data_ptr(), None):
            cabi.check(ctx.lib.y11_plan_set_stem_source(p, C.c_void_p(source or 0)))
            out.zero_()
            cabi.check(ctx.lib.y11_plan_run(p, ctx.stream()))
            torch.cuda.synchronize()
            outs.append(out.clone())
        ctx.lib.y11_plan_destroy(p)
        assert outs[0].abs().sum() > 0
        assert torch.equal(outs[0], outs[1]), (cout, s2d, (outs[0].float() - outs[1].float()).abs().max().item())
        assert torch.equal(outs[0], outs[2])


@pytest.mark.parametrize("act,res,H,W,c", [(True, False, 20, 28, 64), (False, True, 20, 28, 64), (True, True, 14, 20, 64),
                                           # register kernel (maps narrower than 18 pixels or lower than 8 rows, or narrower than 32 with < 512 channels)
                                           (True, False, 1, 8, 64), (False, True, 7, 16, 64), (True, True, 20, 12, 128),
                                           # TMA-ring kernel (W >= 18, H >= 8: the three cases above and these; ragged tiles): 2/4/8-row groups, 80 channels
                                           # (surplus threads), two 128-channel chunks, residual, many tiles per CTA
                                           (True, False, 16, 32, 64), (False, True, 32, 48, 128), (True, True, 24, 64, 80),
                                           (True, False, 16, 32, 256), (True, False, 8 * 20, 16 * 12, 32), (True, True, 40, 40, 128), (False, False, 20, 36, 64),
                                           (True, True, 20, 20, 256), (True, True, 20, 20, 512)])
def test_dwconv(ctx, act, res, H, W, c):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, c, H, W, generator=g).to(ctx.dev).to(torch.bfloat16).float()
    w = (torch.randn(c, 1, 3, 3, generator=g) / 3).to(ctx.dev).to(torch.bfloat16).float()
    b = torch.randn(c, generator=g).to(ctx.dev)
    r = torch.randn(2, c, H, W, generator=g).to(ctx.dev).to(torch.bfloat16).float()
    want = torch.nn.functional.conv2d(x, w, b, padding=1, groups=c)
    if act:
        want = torch.nn.functional.silu(want)
    if res:
        want = want + r
    xin = nhwc(x, c + 32, 16)
    rb = nhwc(r)
    out = torch.zeros((2, H, W, c), dtype=torch.bfloat16, device=ctx.dev)
    wp = w.view(c, 9).t().to(torch.bfloat16).contiguous()
    d = cabi.DwConvDesc(cabi.View(xin.data_ptr(), c + 32, 16, c), cabi.View(out.data_ptr(), c, 0, c),
                        cabi.View(rb.data_ptr(), c, 0, c) if res else cabi.NULL_VIEW, wp.data_ptr(), b.data_ptr(), 2, H, W, int(act))
    p = ctx.plan()
    cabi.check(ctx.lib.y11_plan_add_dwconv(p, C.byref(d)))
    ctx.run(p)
    close_bf16(out.float().permute(0, 3, 1, 2), want)


@pytest.mark.parametrize("h,w", [(20, 20), (14, 20), (40, 40)])
def test_sppf_pools_exact(ctx, h, w):
    c = 32
    x = torch.randn(2, c, h, w, device=ctx.dev).to(torch.bfloat16).float()
    buf = nhwc(torch.cat((x, torch.zeros(2, 3 * c, h, w, device=ctx.dev)), 1))
    d = cabi.SppfDesc(cabi.View(buf.data_ptr(), 4 * c, 0, 4 * c), 2, h, w, c)
    p = ctx.plan()
    cabi.check(ctx.lib.y11_plan_add_sppf(p, C.byref(d)))
    ctx.run(p)
    mp = torch.nn.MaxPool2d(5, 1, 2)
    y1 = mp(x); y2 = mp(y1); y3 = mp(y2)
    want = torch.cat((x, y1, y2, y3), 1)
    assert torch.equal(buf.float().permute(0, 3, 1, 2), want)        # max is exact in bf16


def test_upsample_exact(ctx):
    x = torch.randn(2, 32, 10, 14, device=ctx.dev).to(torch.bfloat16).float()
    xin = nhwc(x, 48, 16)
    out = torch.zeros((2, 20, 28, 96), dtype=torch.bfloat16, device=ctx.dev)
    d = cabi.UpsampleDesc(cabi.View(xin.data_ptr(), 48, 16, 32), cabi.View(out.data_ptr(), 96, 32, 32), 2, 10, 14)
    p = ctx.plan()
    cabi.check(ctx.lib.y11_plan_add_upsample(p, C.byref(d)))
    ctx.run(p)
    want = torch.nn.functional.interpolate(x, scale_factor=2, mode="nearest")
    assert torch.equal(out[..., 32:64].float().permute(0, 3, 1, 2), want)
    assert torch.all(out[..., :32] == 0) and torch.all(out[..., 64:] == 0)


@pytest.mark.parametrize("n_tok,heads", [(400, 2), (280, 4), (1600, 2), (37, 1)])
def test_attention_vs_torch(ctx, n_tok, heads):
    kd, hd, B = 32, 64, 2
    g = torch.Generator().manual_seed(n_tok)
    q = torch.randn(B, heads, n_tok, kd, generator=g).to(ctx.dev).to(torch.bfloat16).float()
    k = torch.randn(B, heads, n_tok, kd, generator=g).to(ctx.dev).to(torch.bfloat16).float()
    v = torch.randn(B, heads, n_tok, hd, generator=g).to(ctx.dev).to(torch.bfloat16).float()
    want = torch.softmax(q @ k.transpose(-1, -2) * kd ** -0.5, -1) @ v              # [B,h,N,hd]
    want = want.permute(0, 2, 1, 3).reshape(B, n_tok, heads * hd)
    qkv = torch.cat((q.permute(0, 2, 1, 3).reshape(B, n_tok, -1), k.permute(0, 2, 1, 3).reshape(B, n_tok, -1),
                     v.permute(0, 2, 1, 3).reshape(B, n_tok, -1)), -1).to(torch.bfloat16).contiguous()
    out = torch.zeros((B, n_tok, heads * hd), dtype=torch.bfloat16, device=ctx.dev)
    d = cabi.AttnDesc(cabi.View(qkv.data_ptr(), qkv.shape[-1], 0, qkv.shape[-1]), cabi.View(out.data_ptr(), heads * hd, 0, heads * hd),
                      B, n_tok, heads, kd, hd, kd ** -0.5)
    p = ctx.plan()
    cabi.check(ctx.lib.y11_plan_add_attention(p, C.byref(d)))
    ctx.run(p)
    close_bf16(out.float(), want)


# ------------------------------------------------------------------------------------------- decode / NMS
def head_desc(feats_nhwc, nc, B):
    hd = cabi.HeadDesc()
    for l, f in enumerate(feats_nhwc):
        hd.head[l] = f.data_ptr()
        hd.hl[l], hd.wl[l] = f.shape[1], f.shape[2]
        hd.stride[l] = float((8, 16, 32)[l])
    hd.nl, hd.B, hd.nc, hd.row_stride = len(feats_nhwc), B, nc, feats_nhwc[0].shape[-1]
    return hd


def test_decode_dense_matches_oracle_and_golden(ctx):
    from pathlib import Path
    d = np.load(str(Path(__file__).parent / "golden" / "decode_small.npz"))
    feats = [torch.from_numpy(d[f"f{i}"]).to(ctx.dev) for i in range(3)]
    nh = [f.permute(0, 2, 3, 1).contiguous() for f in feats]
    hd = head_desc(nh, 80, 2)
    A = sum(f.shape[2] * f.shape[3] for f in feats)
    y = torch.zeros((2, 84, A), device=ctx.dev)
    cabi.check(ctx.lib.y11_decode_dense(ctx.h, C.byref(hd), y.data_ptr(), ctx.stream()))
    torch.cuda.synchronize()
    want = torch.from_numpy(d["y"]).to(ctx.dev)
    assert torch.allclose(y[:, :4], want[:, :4], rtol=1e-5, atol=2e-3)      # pixels
    assert torch.allclose(y[:, 4:], want[:, 4:], rtol=1e-5, atol=1e-6)      # scores


def nms_gpu(ctx, boxes, scores, cls, n, iou, max_det, agnostic=False, max_nms=30000):
    B, K = scores.shape
    keep = torch.full((B, max_det), -1, dtype=torch.int32, device=ctx.dev)
    cnt = torch.zeros((B,), dtype=torch.int32, device=ctx.dev)
    ws = torch.empty(ctx.lib.y11_nms_workspace(B, K), dtype=torch.uint8, device=ctx.dev)
    p = cabi.NmsParams(0.0, iou, max_det, max_nms, 7680, int(agnostic), 0)
    cabi.check(ctx.lib.y11_nms_batched(ctx.h, boxes.data_ptr(), scores.data_ptr(), cls.data_ptr(), n.data_ptr(), B, K, C.byref(p),
                                       keep.data_ptr(), cnt.data_ptr(), ws.data_ptr(), ws.numel(), ctx.stream()))
    torch.cuda.synchronize()
    return keep.cpu(), cnt.cpu()


def synth_boxes(K, seed, clustered):
    g = torch.Generator().manual_seed(seed)
    if clustered:
        centers = torch.rand(max(K // 8, 1), 2, generator=g) * 1100 + 50
        xy = centers[torch.randint(0, centers.shape[0], (K,), generator=g)] + (torch.rand(K, 2, generator=g) - 0.5) * 4
        scores = (torch.rand(K, generator=g) * 256).round() / 256          # exact ties
    else:
        xy = torch.rand(K, 2, generator=g) * 1200
        scores = torch.rand(K, generator=g)
    wh = torch.rand(K, 2, generator=g) * 200 + 4
    boxes = torch.cat((xy - wh / 2, xy + wh / 2), 1)
    cls = torch.randint(0, 80, (K,), generator=g).float()
    return boxes, scores, cls


@pytest.mark.parametrize("K", [1, 63, 64, 65, 300, 1000, 8400, 20000, 33600])
@pytest.mark.parametrize("clustered", [False, True])
def test_nms_keep_set_bit_exact_vs_torchvision(ctx, K, clustered):
    boxes, scores, cls = synth_boxes(K, K + int(clustered), clustered)
    for iou in (0.45, 0.7, 0.6):
        want = torchvision.ops.nms(boxes + cls[:, None] * 7680, scores, iou)
        keep, cnt = nms_gpu(ctx, boxes[None].to(ctx.dev), scores[None].to(ctx.dev), cls[None].to(ctx.dev),
                            torch.tensor([K], dtype=torch.int32, device=ctx.dev), iou, K, max_nms=1 << 30)
        assert cnt[0].item() == want.numel()
        assert keep[0, :cnt[0]].tolist() == want.tolist()


def test_nms_batched_ragged_empty_maxdet_agnostic(ctx):
    K = 512
    data = [synth_boxes(K, s, True) for s in range(4)]
    boxes = torch.stack([d[0] for d in data]).to(ctx.dev)
    scores = torch.stack([d[1] for d in data]).to(ctx.dev)
    cls = torch.stack([d[2] for d in data]).to(ctx.dev)
    n = torch.tensor([512, 0, 1, 300], dtype=torch.int32, device=ctx.dev)
    for agnostic in (False, True):
        keep, cnt = nms_gpu(ctx, boxes, scores, cls, n, 0.5, 100, agnostic=agnostic)
        for b in range(4):
            nb = int(n[b])
            off = 0 if agnostic else 7680
            want = torchvision.ops.nms(data[b][0][:nb] + data[b][2][:nb, None] * off, data[b][1][:nb], 0.5)[:100] if nb else torch.zeros(0)
            assert cnt[b].item() == want.numel()
            assert keep[b, :cnt[b]].tolist() == want.tolist()


def test_golden_nms_fixture(ctx):
    from pathlib import Path
    g = np.load(str(Path(__file__).parent / "golden" / "nms_small.npz"))
    K = g["scores"].shape[0]
    keep, cnt = nms_gpu(ctx, torch.from_numpy(g["boxes"])[None].to(ctx.dev), torch.from_numpy(g["scores"])[None].to(ctx.dev),
                        torch.from_numpy(g["cls"])[None].to(ctx.dev), torch.tensor([K], dtype=torch.int32, device=ctx.dev), float(g["iou"]), K)
    assert keep[0, :cnt[0]].tolist() == g["keep"].tolist()


def test_multi_label_more_candidates_than_max_nms(ctx):
    """val regime (conf 0.001, multi-label) where almost every (anchor, class) pair passes: 168 000 candidates per image, far
    more than max_nms and than the 131 072-entry list the first version kept (in anchor order, silently dropping the rest).
    The max_nms best BY SCORE must survive (stable on ties), exactly as the oracle's restatement of ultralytics does."""
    B, nc = 2, 80
    g = torch.Generator().manual_seed(11)
    dims = [(40, 40), (20, 20), (10, 10)]
    feats = [(torch.randn(B, h, w, 64 + nc, generator=g) * 2).to(ctx.dev) for (h, w) in dims]
    for f in feats:
        f[..., 64:] -= 3.0
    hd = head_desc(feats, nc, B)
    A = sum(h * w for h, w in dims)
    y = torch.zeros((B, 84, A), device=ctx.dev)
    cabi.check(ctx.lib.y11_decode_dense(ctx.h, C.byref(hd), y.data_ptr(), ctx.stream()))
    conf, iou, max_det, max_nms = 0.001, 0.6, 100, 4000
    det = torch.zeros((B, max_det, 6), device=ctx.dev)
    cnt = torch.zeros((B,), dtype=torch.int32, device=ctx.dev)
    ncand = torch.zeros((B,), dtype=torch.int32, device=ctx.dev)
    ws = torch.empty(ctx.lib.y11_postprocess_workspace(B, A, nc, 1, max_nms), dtype=torch.uint8, device=ctx.dev)
    p = cabi.NmsParams(conf, iou, max_det, max_nms, 7680, 0, 1)
    cabi.check(ctx.lib.y11_detect_postprocess(ctx.h, C.byref(hd), C.byref(p), None, det.data_ptr(), cnt.data_ptr(),
                                              ncand.data_ptr(), ws.data_ptr(), ws.numel(), ctx.stream()))
    torch.cuda.synchronize()
    assert int(ncand.min()) > 131072
    want = P.non_max_suppression(y.cpu(), conf, iou, multi_label=True, max_det=max_det, max_nms=max_nms)
    for b in range(B):
        n = int(cnt[b])
        assert n == want[b].shape[0] == max_det
        got = det[b, :n].cpu()
        assert torch.equal(got[:, 4:], want[b][:, 4:])
        assert (got[:, :4] - want[b][:, :4]).abs().max() <= 1e-3


@pytest.mark.parametrize("case", ["clustered", "ties", "few"])
def test_progressive_topk_stages_equal_sorting_everything(ctx, case):
    """Long multi-label candidate lists are processed as score-ordered prefixes (stage 1: exact top 8192, stage 2: top max_nms,
    then the whole list) - csrc/postprocess.cu `topk_stage`.  Whatever stage finishes an image, the result must be the oracle's
    (= sorting everything): `clustered` = every box overlaps one of a few sites, so stage 1 keeps fewer than max_det boxes and the
    image goes on to stage 2; `ties` = tens of thousands of identical scores at the selection threshold (saturated logits), wider
    than both compact buffers, so the image falls through to the full-list path; `few` = below 8192 rows (stage 1 takes all)."""
    B, nc = 2, 80
    g = torch.Generator().manual_seed({"clustered": 1, "ties": 2, "few": 3}[case])
    dims = [(40, 40), (20, 20), (10, 10)]
    feats = [(torch.randn(B, h, w, 64 + nc, generator=g) * 2) for (h, w) in dims]
    for f in feats:
        if case == "clustered":
            f[..., :64] = 0.0
            f[..., 15:64:16] = 12.0           # every side's DFL expectation = 15 bins: huge boxes, all overlapping their neighbours
            f[..., 64:] -= 3.0
            f[..., 64 + 10:] = -20.0          # ten live classes: ~20 000 rows per image, a handful of survivors per class
        elif case == "ties":
            f[..., 64:] = 30.0                # sigmoid saturates: every (anchor, class) score is exactly 1.0
        else:
            f[..., 64:] -= 10.8               # a few thousand pairs above conf 0.001
    feats = [f.to(ctx.dev) for f in feats]
    hd = head_desc(feats, nc, B)
    A = sum(h * w for h, w in dims)
    y = torch.zeros((B, 84, A), device=ctx.dev)
    cabi.check(ctx.lib.y11_decode_dense(ctx.h, C.byref(hd), y.data_ptr(), ctx.stream()))
    conf, iou, max_det, max_nms = 0.001, (0.1 if case == "clustered" else 0.6), 300, 30000
    det = torch.zeros((B, max_det, 6), device=ctx.dev)
    cnt = torch.zeros((B,), dtype=torch.int32, device=ctx.dev)
    ncand = torch.zeros((B,), dtype=torch.int32, device=ctx.dev)
    ws = torch.empty(ctx.lib.y11_postprocess_workspace(B, A, nc, 1, max_nms), dtype=torch.uint8, device=ctx.dev)
    p = cabi.NmsParams(conf, iou, max_det, max_nms, 7680, 0, 1)
    for _ in range(2):     # twice: the stage buffers and flags are reused from call to call
        cabi.check(ctx.lib.y11_detect_postprocess(ctx.h, C.byref(hd), C.byref(p), None, det.data_ptr(), cnt.data_ptr(),
                                                  ncand.data_ptr(), ws.data_ptr(), ws.numel(), ctx.stream()))
    torch.cuda.synchronize()
    want = P.non_max_suppression(y.cpu(), conf, iou, multi_label=True, max_det=max_det, max_nms=max_nms)
    nmin = int(ncand.min())
    if case == "few":
        assert 0 < nmin and int(ncand.max()) < 8192
    elif case == "clustered":
        assert nmin > 8192
    else:
        assert nmin > 65536
    for b in range(B):
        n = int(cnt[b])
        assert n == want[b].shape[0]
        if case == "clustered":
            assert n < max_det           # stage 1 could not finish this image
        got = det[b, :n].cpu()
        assert torch.equal(got[:, 4:], want[b][:, 4:])
        assert (got[:, :4] - want[b][:, :4]).abs().max() <= 1e-3


@pytest.mark.parametrize("multi_label", [False, True])
def test_fused_postprocess_matches_oracle_nms(ctx, multi_label):
    """decode -> compaction -> sort -> NMS -> scale_boxes, vs oracle non_max_suppression on the GPU's own dense decode
    (identical decoded inputs => keep-set and order bit-exact; boxes within 0.5 px)."""
    B, nc = 3, 80
    g = torch.Generator().manual_seed(7)
    dims = [(16, 24), (8, 12), (4, 6)]
    feats = [(torch.randn(B, h, w, 64 + nc, generator=g) * 2).to(ctx.dev) for (h, w) in dims]
    for f in feats:
        f[..., 64:] -= 3.0 if not multi_label else 4.5
    hd = head_desc(feats, nc, B)
    A = sum(h * w for h, w in dims)
    y = torch.zeros((B, 84, A), device=ctx.dev)
    cabi.check(ctx.lib.y11_decode_dense(ctx.h, C.byref(hd), y.data_ptr(), ctx.stream()))
    conf, iou, max_det = (0.25, 0.45, 30) if not multi_label else (0.05, 0.6, 50)
    orig = (300, 400)
    net_hw = (128, 192)
    gain, px, py = P.scale_boxes_params(net_hw, orig)
    scale = torch.tensor([[gain, px, py, orig[1], orig[0]]] * B, dtype=torch.float32, device=ctx.dev)
    det = torch.zeros((B, max_det, 6), device=ctx.dev)
    cnt = torch.zeros((B,), dtype=torch.int32, device=ctx.dev)
    ncand = torch.zeros((B,), dtype=torch.int32, device=ctx.dev)
    ws = torch.empty(ctx.lib.y11_postprocess_workspace(B, A, nc, int(multi_label), 30000), dtype=torch.uint8, device=ctx.dev)
    p = cabi.NmsParams(conf, iou, max_det, 30000, 7680, 0, int(multi_label))
    cabi.check(ctx.lib.y11_detect_postprocess(ctx.h, C.byref(hd), C.byref(p), scale.data_ptr(), det.data_ptr(), cnt.data_ptr(),
                                              ncand.data_ptr(), ws.data_ptr(), ws.numel(), ctx.stream()))
    torch.cuda.synchronize()
    want = P.non_max_suppression(y.cpu(), conf, iou, multi_label=multi_label, max_det=max_det)
    assert int(ncand.sum()) > 50
    for b in range(B):
        w = want[b].clone()
        if w.shape[0]:
            P.scale_boxes(net_hw, w[:, :4], orig)
        n = int(cnt[b])
        assert n == w.shape[0]
        got = det[b, :n].cpu()
        assert torch.equal(got[:, 4:], w[:, 4:])                        # same candidates, same order, same scores/classes
        assert (got[:, :4] - w[:, :4]).abs().max() <= 0.5 if n else True  # stated tolerance 0.5 px (observed ~1e-4)


def test_conv_dynamic_tile_scheduler_matches_static(ctx, monkeypatch):
    """Y11_DYN_TILES=1 (global atomic tile counter, re-armed by the last CTA of every launch) gives the same bits as the
    static tile walk, launch after launch of the same op."""
    shapes = [(3, 80, 80, 32, 64, 3, 1, True, True), (8, 20, 20, 192, 384, 1, 1, True, False), (2, 32, 32, 64, 128, 3, 2, True, False)]
    base = [conv_case(ctx, *sh)[0] for sh in shapes]
    monkeypatch.setenv("Y11_DYN_TILES", "1")
    for sh, want in zip(shapes, base):
        for repeats in (1, 4):  # repeated launches of the same op exercise the counter re-arm (the output is idempotent)
            assert torch.equal(conv_case(ctx, *sh, repeats=repeats)[0], want)
