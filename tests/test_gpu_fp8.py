"""GPU (-m gpu): FP8 (e4m3) mode - tcgen05.mma.kind::f8f6f4 convs against torch, the quantised network against the quantised
oracle (oracle/quant_ref.py), and the quantizer surface of the reference (optimization/quantization/quantizers.py:860-888)."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import quant_ref as Q  # noqa: E402
from oracle import yolo11_ref as R  # noqa: E402
from yolo_infer_b200 import YOLO11Model, _cabi as cabi  # noqa: E402
from yolo_infer_b200.engine import YOLO  # noqa: E402
from yolo_infer_b200.network import fp8_pairs  # noqa: E402
from yolo_infer_b200.quant import calibrate_activation_scales, create_quantizer  # noqa: E402

from gpu_utils import Ctx  # noqa: E402


@pytest.fixture(scope="module")
def ctx():
    c = Ctx()
    yield c
    c.close()


def e4m3(t):
    return t.clamp(-448, 448).to(torch.float8_e4m3fn)


def fp8_conv_case(ctx, B, H, W, cin, cout, k, stride, act, res, out_kind, in_fp8, seed=0):
    """One conv through the C ABI with e4m3 input/weights (in_fp8) and / or e4m3 output (out_kind == 'fp8'); returns
    (got, want) as fp32 NCHW in REAL units."""
    g = torch.Generator().manual_seed(seed)
    dev = ctx.dev
    s_in, s_out = 0.037, 0.021
    x_real = torch.randn(B, cin, H, W, generator=g) * 2
    w_real = torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5
    bias = torch.randn(cout, generator=g).to(dev)
    Ho, Wo = (H + stride - 1) // stride, (W + stride - 1) // stride
    d = cabi.ConvDesc()
    if in_fp8:
        xq = e4m3(x_real / s_in)
        s_w = (w_real.abs().amax(dim=(1, 2, 3)) / 448).clamp_min(1e-12)
        wq = e4m3(w_real / s_w.view(-1, 1, 1, 1))
        xin = xq.view(torch.uint8).permute(0, 2, 3, 1).contiguous().to(dev)
        wp = wq.view(torch.uint8).permute(0, 2, 3, 1).reshape(cout, -1).contiguous().to(dev)
        cscale = (s_w * s_in).to(dev)
        d.in_fp8, d.cscale = 1, cscale.data_ptr()
        want = torch.nn.functional.conv2d(xq.float().to(dev), wq.float().to(dev), None, stride=stride, padding=k // 2)
        want = want * cscale.view(1, -1, 1, 1) + bias.view(1, -1, 1, 1)
    else:
        xb = x_real.to(torch.bfloat16)
        wb = w_real.to(torch.bfloat16)
        xin = xb.permute(0, 2, 3, 1).contiguous().to(dev)
        wp = wb.permute(0, 2, 3, 1).reshape(cout, -1).contiguous().to(dev)
        want = torch.nn.functional.conv2d(xb.float().to(dev), wb.float().to(dev), bias, stride=stride, padding=k // 2)
    if act:
        want = torch.nn.functional.silu(want)
    r = None
    if res:
        r = torch.randn(B, cout, Ho, Wo, generator=g).to(torch.bfloat16)
        want = want + r.float().to(dev)
        rb = r.permute(0, 2, 3, 1).contiguous().to(dev)
        d.res = cabi.View(rb.data_ptr(), cout, 0, cout)
    odt = {"bf16": torch.bfloat16, "f32": torch.float32, "fp8": torch.uint8}[out_kind]
    out = torch.zeros((B, Ho, Wo, cout), dtype=odt, device=dev)
    d.inp = cabi.View(xin.data_ptr(), cin, 0, cin)
    d.out = cabi.View(out.data_ptr(), cout, 0, cout)
    d.w, d.bias = wp.data_ptr(), bias.data_ptr()
    d.B, d.Hin, d.Win, d.Hout, d.Wout = B, H, W, Ho, Wo
    d.k, d.stride, d.act, d.out_f32, d.impl = k, stride, int(act), int(out_kind == "f32"), cabi.IMPL_TCGEN05
    d.out_fp8, d.out_scale = int(out_kind == "fp8"), 1.0 / s_out
    p = ctx.plan()
    cabi.check(ctx.lib.y11_plan_add_conv(p, C.byref(d)), "add_conv")
    ctx.run(p)
    if out_kind == "fp8":
        got = out.view(torch.float8_e4m3fn).float().permute(0, 3, 1, 2) * s_out
        want_q = e4m3(want / s_out).float() * s_out
        return got, want_q, want
    return out.float().permute(0, 3, 1, 2), want, want


FP8_CASES = [  # B, H, W, cin, cout, k, stride, act, res, out_kind
    (2, 40, 40, 64, 64, 3, 1, True, True, "bf16"), (3, 20, 20, 128, 128, 3, 1, True, True, "bf16"), (2, 24, 20, 32, 64, 3, 1, True, False, "bf16"),
    (2, 80, 80, 64, 64, 3, 1, True, False, "fp8"), (2, 40, 36, 64, 64, 1, 1, False, False, "f32"), (1, 20, 20, 256, 256, 3, 1, True, True, "bf16"),
    (2, 40, 40, 96, 96, 3, 1, True, True, "bf16"), (2, 17, 23, 64, 128, 3, 2, True, False, "bf16"),
]


@pytest.mark.parametrize("case", FP8_CASES, ids=lambda c: "x".join(map(str, c)))
def test_fp8_conv_vs_torch(ctx, case):
    """e4m3 x e4m3 products are exact in fp32, so the only differences to torch are summation order and the SiLU approximation:
    bf16 / fp32 outputs to one bf16 ulp; e4m3 outputs exact except where the pre-rounding value sits on a rounding boundary."""
    B, H, W, cin, cout, k, stride, act, res, out_kind = case
    got, want, _ = fp8_conv_case(ctx, B, H, W, cin, cout, k, stride, act, res, out_kind, in_fp8=True)
    if out_kind == "fp8":
        diff = (got - want).abs()
        frac = float((diff > 0).float().mean())
        assert frac <= 0.01, frac
        assert torch.all(diff <= 0.13 * want.abs() + 1e-3)       # at most one e4m3 step (2^-3 relative)
    else:
        err = (got - want).abs()
        assert torch.all(err <= 1e-2 * want.abs() + 2e-2), float(err.max())


@pytest.mark.parametrize("cout", [32, 64, 128])
def test_bf16_conv_with_fp8_output(ctx, cout):
    """The producer side of an FP8 edge: bf16 operands, e4m3(value / s) stored by the epilogue."""
    got, want_q, want = fp8_conv_case(ctx, 2, 40, 40, 64, cout, 3, 1, True, False, "fp8", in_fp8=False, seed=3)
    diff = (got - want_q).abs()
    assert float((diff > 0).float().mean()) <= 0.01
    assert torch.all(diff <= 0.13 * want_q.abs() + 1e-3)
    assert float((got - want).abs().max()) > 0          # it IS quantised


def heads(eng, x):
    net = eng.compiled(x.shape[0], x.shape[2], x.shape[3])
    eng.preprocess_tensor(net, x.to(eng.device).contiguous(), 1.0)
    eng.forward(net)
    torch.cuda.synchronize()
    return net.raw_head().cpu(), net


@pytest.mark.parametrize("scale", ["n", "s"])
def test_fp8_network_matches_the_quantised_oracle(oracle_models, scale):
    """Whole network in FP8 mode against oracle/quant_ref.emulate_fp8 with the SAME activation scales: relative L2 of the raw head.
    Stated tolerance 3.5e-2 (measured 2.9e-2), or 1.5x the control when that is larger: an e4m3 rounding flip is a 6-12 % change of
    that element (vs 0.4 % for bf16), so the noise floor that summation order / SiLU approximation differences produce is ~4x the
    bf16 network's; the control (printed) measures that floor without any GPU involved.  The FP8 network sits ~3.5 % from the bf16
    network (printed)."""
    _, sd = oracle_models(scale)

    def fused():
        m = R.DetectionModel(scale)
        m.load_state_dict(sd)
        return m.eval().fuse()

    assert sorted(Q.fp8_edges(fused())) == sorted((p, c) for p, c, _ in fp8_pairs(scale))
    x = torch.rand(2, 3, 320, 320, generator=torch.Generator().manual_seed(5))
    emul = R.emulate_bf16_storage(fused())
    scales = Q.calibrate(emul, [x])
    eng = YOLO.from_state_dict(sd, scale).to("cuda:0")
    mine = calibrate_activation_scales(eng, [x])
    assert set(mine) == set(scales)
    worst = max(abs(mine[k] - scales[k]) / scales[k] for k in scales)
    assert worst <= 0.05, worst                       # the bf16 plan measures the same maxima as the oracle (to bf16 noise)
    base, _ = heads(eng, x)
    eng.enable_fp8(scales)
    got, net = heads(eng, x)
    n_q = sum(1 for o in net.ops if o.kind == "conv" and o.name.endswith("#fp8"))
    assert n_q == len(scales) and any(b.dtype == torch.uint8 for b in net.buffers)
    with torch.no_grad():
        _, feats = Q.emulate_fp8(fused(), scales)(x)
        _, feats16 = emul(x)
    want = torch.cat([f.view(2, 144, -1) for f in feats], 2)
    want16 = torch.cat([f.view(2, 144, -1) for f in feats16], 2)
    rel = float((got - want).norm() / want.norm())
    # control: the quantised oracle against itself when 1 % of the input pixels move by one bf16 ulp - the floor that e4m3
    # rounding flips (one flip = a 6-12 % change of that element) put under ANY two evaluations of this network
    xp = x.clone()
    pick = torch.rand(x.shape, generator=torch.Generator().manual_seed(6)) < 0.01
    xp[pick] = xp[pick].to(torch.bfloat16).float() * (1 + 2.0 ** -8)
    with torch.no_grad():
        _, featsp = Q.emulate_fp8(fused(), scales)(xp)
    control = float((torch.cat([f.view(2, 144, -1) for f in featsp], 2) - want).norm() / want.norm())
    print(f"yolo11{scale} FP8: {n_q} e4m3 convs; head rel-L2 vs quantised oracle {rel:.3e} (control oracle-vs-oracle {control:.3e}); FP8 vs bf16 engine "
          f"{float((got - base).norm() / base.norm()):.3e}; quantised oracle vs bf16-storage oracle {float((want - want16).norm() / want16.norm()):.3e}")
    assert rel <= max(3.5e-2, 1.5 * control), (rel, control)
    eng.disable_fp8()
    again, _ = heads(eng, x)
    assert torch.equal(again, base)


def test_quantizer_surface(oracle_models, tmp_path):
    _, sd = oracle_models("n")
    eng = YOLO.from_state_dict(sd, "n").to("cuda:0")
    path = tmp_path / "w.pt"
    eng.save(path)
    model = YOLO11Model(model_path=str(path), device="cuda:0", verbose=False)
    with pytest.raises(ValueError):
        create_quantizer("int4", model)
    q = create_quantizer("fp8", model, {"num_calibration_batches": 2})
    with pytest.raises(ValueError):
        q.optimize()                                   # calibration data is required (quantizers.py:60-61)
    g = torch.Generator().manual_seed(1)
    loader = [torch.randint(0, 256, (2, 320, 320, 3), generator=g, dtype=torch.uint8) for _ in range(3)]
    frames = loader[0]
    before = model.predict(frames, conf=0.25, verbose=False)
    out = q.optimize(calibration_loader=loader)
    assert out is model and q.optimization_metrics["fp8_edges"] == len(fp8_pairs("n")) and model.optimization_history
    after = model.predict(frames, conf=0.25, verbose=False)
    assert len(after) == 2 and all(abs(len(a.boxes) - len(b.boxes)) <= max(5, 0.3 * len(b.boxes)) for a, b in zip(after, before))
    m = q.evaluate(torch.rand(1, 3, 320, 320, generator=g))
    assert 0 < m["head_rel_l2_vs_bf16"] < 0.15
