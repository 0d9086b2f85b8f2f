"""GPU (-m gpu): whole-path parity of the B200 engine against the oracle on the same weights and inputs, through the
reference-facing surface (YOLO11Model.predict) - SURVEY.md section 4 test plan items (3) and (4).

Stated tolerances (north_star): raw head outputs rel <= 1e-2 = relative L2 error per tensor against the oracle evaluated with
the SAME storage format (bf16 weights / bf16 activation tensors, fp32 accumulation: oracle.emulate_bf16_storage), i.e. the
kernels' arithmetic; against the pure-fp32 oracle the bound is 3e-2, because ~40 sequential bf16 tensor roundings alone put
a ~1 % floor under ANY bf16 implementation (measured on CPU: emulated-bf16 oracle vs fp32 oracle = 1.3 % on box logits);
NMS keep-set/order bit-exact on identical decoded inputs (tests/test_gpu_kernels.py);
final boxes within 0.5 px when fed identical head outputs; end-to-end (bf16 network vs fp32 oracle) detections are
matched one-to-one and compared with a looser, stated tolerance because score/threshold ties can flip under bf16.
"""
import ctypes as C
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import pipeline_ref as P  # noqa: E402
from oracle import yolo11_ref as R  # noqa: E402
from yolo_infer_b200 import YOLO11Model, _cabi as cabi  # noqa: E402
from yolo_infer_b200.engine import YOLO  # noqa: E402

ROOT = Path(__file__).resolve().parents[1]


def rel_l2(got: torch.Tensor, want: torch.Tensor) -> float:
    return float((got - want).norm() / want.norm().clamp_min(1e-12))


@pytest.fixture(scope="module")
def engines(oracle_models):
    cache = {}

    def get(scale):
        if scale not in cache:
            m, sd = oracle_models(scale)
            eng = YOLO.from_state_dict(sd, scale).to("cuda:0")
            fused = R.DetectionModel(scale)
            fused.load_state_dict(sd)
            fused.eval().fuse()
            emul = R.DetectionModel(scale)
            emul.load_state_dict(sd)
            R.emulate_bf16_storage(emul.eval().fuse())
            cache[scale] = (eng, fused, emul)
        return cache[scale]

    return get


def oracle_head(fused, x):
    with torch.no_grad():
        y, feats = fused(x)
    B = x.shape[0]
    return y, torch.cat([f.view(B, f.shape[1], -1) for f in feats], 2)


# BASELINE.json configs as parity cases: n @640 / @448x640 (config 0, 1), s @640 (config 2), m @1280 (config 3, A = 33600),
# x @640 batch 1 (config 4)
@pytest.mark.parametrize("scale,B,H,W", [("n", 2, 640, 640), ("n", 1, 448, 640), ("s", 2, 640, 640), ("n", 3, 384, 640),
                                         ("m", 1, 1280, 1280), ("x", 1, 640, 640)])
def test_raw_head_outputs_within_bf16_tolerance(engines, scale, B, H, W):
    eng, fused, emul = engines(scale)
    g = torch.Generator().manual_seed(B * H + W)
    x = torch.rand(B, 3, H, W, generator=g)
    _, want32 = oracle_head(fused, x)                                 # [B,144,A] fp32 everywhere
    _, want = oracle_head(emul, x)                                    # same network, bf16 storage / fp32 accumulate
    net = eng.compiled(B, H, W)
    eng.preprocess_tensor(net, x.to("cuda:0").contiguous(), 1.0)
    eng.forward(net)
    torch.cuda.synchronize()
    got = net.raw_head().cpu()
    assert got.shape == want.shape
    e_box, e_cls = rel_l2(got[:, :64], want[:, :64]), rel_l2(got[:, 64:], want[:, 64:])
    f_box, f_cls = rel_l2(got[:, :64], want32[:, :64]), rel_l2(got[:, 64:], want32[:, 64:])
    # per Detect level = per head tensor [B,144,H_l,W_l] (box and class logits as ultralytics returns them)
    per_level, off = [], 0
    for h in net.head:
        n = h.shape[1] * h.shape[2]
        per_level.append(rel_l2(got[:, :, off:off + n], want[:, :, off:off + n]))
        off += n
    # control: what the WEIGHTS do to any perturbation - the same-storage oracle against itself with 1 % of the input pixels moved
    # by one bf16 ulp (no GPU involved)
    xp = x.clone()
    pick = torch.rand(x.shape, generator=torch.Generator().manual_seed(1)) < 0.01
    xp[pick] = xp[pick].to(torch.bfloat16).float() * (1 + 2.0 ** -8)
    _, wantp = oracle_head(emul, xp)
    c_all, c_box = rel_l2(wantp, want), rel_l2(wantp[:, :64], want[:, :64])
    whole = rel_l2(got, want)
    print(f"yolo11{scale} {B}x{H}x{W}: head rel-L2 vs same-storage oracle {whole:.3e} (per level {[f'{v:.2e}' for v in per_level]}; box logits "
          f"{e_box:.3e}, class logits {e_cls:.3e}); control oracle-vs-oracle {c_all:.3e} (box {c_box:.3e}); vs fp32 oracle: box {f_box:.3e} cls {f_cls:.3e}")
    # north_star tolerance: raw head outputs rel <= 1e-2 against the same-storage oracle (emulate_bf16_storage rounds to bf16
    # exactly where the CUDA path stores a tensor; what is left is summation order and the MUFU tanh / exp2 approximations).
    # The box-logit half alone has no large bias in its norm and sits AT the floor the control measures (a one-ulp change of
    # 1 % of the input pixels moves it by ~0.8 %: roundings flip and every later layer re-rounds): it is bounded by the
    # control, the whole tensor and every level by 1e-2.  Against the pure fp32 oracle the storage format alone costs ~1 %.
    # The per-layer test (test_every_conv_of_the_plan_matches_torch) is the amplification-free one (1 bf16 ulp per layer).
    assert whole <= 1e-2 and max(per_level) <= 1e-2, (whole, per_level)
    assert e_cls <= 1e-2 and e_box <= max(1e-2, 2.0 * c_box), (e_box, e_cls, c_box)
    assert f_box <= 3e-2 and f_cls <= 3e-2, (f_box, f_cls)


def test_tcgen05_and_simt_debug_paths_agree(engines, oracle_models):
    """Bring-up cross-check: the same plan with the naive CUDA-core conv gives the same head (both bf16)."""
    eng = engines("n")[0]
    _, sd = oracle_models("n")
    dbg = YOLO.from_state_dict(sd, "n")
    dbg.conv_impl = cabi.IMPL_SIMT_DEBUG
    dbg.to("cuda:0")
    x = torch.rand(1, 3, 320, 320, generator=torch.Generator().manual_seed(3)).to("cuda:0")
    outs = []
    for e in (eng, dbg):
        net = e.compiled(1, 320, 320)
        e.preprocess_tensor(net, x, 1.0)
        e.forward(net)
        torch.cuda.synchronize()
        outs.append(net.raw_head().cpu())
    assert rel_l2(outs[0], outs[1]) <= 1e-2, rel_l2(outs[0], outs[1])


@pytest.mark.parametrize("scale,B,H,W", [("n", 2, 320, 320), ("s", 1, 224, 320)])
def test_folded_upsample_concat_equals_copy_ops(engines, scale, B, H, W):
    """Upsample -> Concat -> C3k2.cv1 run as `W_up . p` at low resolution + a pre-activation term (no upsampled / concatenated
    tensor) gives the same head as the plan that materialises both (one extra bf16 rounding of the low-resolution term)."""
    eng = engines(scale)[0]
    x = torch.rand(B, 3, H, W, generator=torch.Generator().manual_seed(5)).to("cuda:0")
    outs = []
    for fold in (True, False):
        net = eng.compiled(B, H, W, fold_upsample=fold)
        assert bool(net.folds) == fold
        assert any(o.kind == "upsample" for o in net.ops) != fold
        eng.preprocess_tensor(net, x, 1.0)
        eng.forward(net)
        torch.cuda.synchronize()
        outs.append(net.raw_head().cpu())
    assert rel_l2(outs[0], outs[1]) <= 1e-2, rel_l2(outs[0], outs[1])


def test_autotuned_plan_is_bit_identical_to_heuristic_plan(engines, monkeypatch):
    """The plan autotuner only changes launch variants: the head of an autotuned plan equals the heuristic plan bit for bit."""
    from yolo_infer_b200 import network as N
    eng = engines("n")[0]
    x = torch.rand(2, 3, 320, 320, generator=torch.Generator().manual_seed(9)).to("cuda:0")
    outs, variants = [], []
    for tuned in (True, False):
        monkeypatch.setattr(N, "AUTOTUNE", tuned)
        with torch.cuda.device(eng.device):
            net = N.CompiledNet(eng._engine, eng.scale, eng.nc, eng._packed, 2, 320, 320, eng.device)
        eng.preprocess_tensor(net, x, 1.0)
        eng.forward(net)
        torch.cuda.synchronize()
        outs.append(net.raw_head().cpu())
        variants.append(net.variants())
    assert torch.equal(outs[0], outs[1])
    assert all(v[2] in (1, 2, 3, 4) for v, o in zip(variants[0], net.ops) if o.kind == "conv")


def test_tune_cache_round_trip(engines, monkeypatch, tmp_path):
    """Y11_TUNE_CACHE: the variants an autotune run picked are stored and re-applied verbatim by the next build."""
    import json
    from yolo_infer_b200 import network as N
    eng = engines("n")[0]
    cache = tmp_path / "tune.json"
    monkeypatch.setattr(N, "AUTOTUNE", True)
    monkeypatch.setattr(N, "TUNE_CACHE", str(cache))
    nets = []
    for _ in range(2):
        with torch.cuda.device(eng.device):
            nets.append(N.CompiledNet(eng._engine, eng.scale, eng.nc, eng._packed, 1, 256, 320, eng.device))
    stored = json.loads(cache.read_text())
    assert list(stored) == [nets[0]._tune_key]
    assert nets[0]._cached_variants is None and nets[1]._cached_variants is not None
    by_name = stored[nets[0]._tune_key]          # stored by op name, keyed by device / SM count / plan configuration
    assert nets[0].variants() == nets[1].variants()
    assert all(tuple(by_name[o.name]) == v for o, v in zip(nets[1].ops, nets[1].variants()) if o.kind == "conv")
    assert "NVIDIA" in nets[0]._tune_key and "/sm" in nets[0]._tune_key


def match_detections(got: torch.Tensor, want: torch.Tensor, conf: float, margin: float = 0.03):
    """Every detection of one side whose score clears the threshold by `margin` must exist on the other side
    (same class, IoU >= 0.8, |score delta| <= margin).  Returns (fraction matched both ways, median over matches of the max coordinate delta in px)."""
    from torchvision.ops import box_iou

    def one_way(a, b):
        a = a[a[:, 4] >= conf + margin]
        if not len(a):
            return 1.0, 0.0
        if not len(b):
            return 0.0, 0.0
        iou = box_iou(a[:, :4], b[:, :4])
        iou[a[:, 5, None] != b[None, :, 5]] = 0
        iou[(a[:, 4, None] - b[None, :, 4]).abs() > margin] = 0
        best, j = iou.max(1)
        ok = best >= 0.8
        # median, not max: with random weights some DFL distributions have two near-equal top bins, where a 1 % logit change
        # moves the expectation by whole bins (x stride 32 px) - a property of the weights, not of the kernels
        delta = float((a[ok, :4] - b[j[ok], :4]).abs().max(1).values.median()) if ok.any() else 0.0
        return float(ok.float().mean()), delta

    f1, w1 = one_way(got, want)
    f2, w2 = one_way(want, got)
    return min(f1, f2), max(w1, w2)


def test_predict_image_jpg_like_config1(engines, oracle_models):
    """Config #1 geometry: an 853x1280 BGR frame -> 448x640 under rect=True; demo iou 0.45 (conf 0.3: the calibrated
    random weights put almost no score above the demo's 0.5)."""
    eng, _, fused = engines("n")     # oracle = reference pipeline around the same-storage network
    rng = np.random.default_rng(0)
    yy, xx = np.mgrid[0:853, 0:1280]
    img = np.stack([(xx * 0.2 + yy * 0.1) % 256, (xx * 0.05 + 40 * np.sin(yy / 37.0)) % 256, rng.integers(0, 256, (853, 1280))], -1).astype(np.uint8)
    res = eng.predict(img, conf=0.3, iou=0.45, show=False, save=False, verbose=False)
    assert len(res) == 1 and res[0].orig_shape == (853, 1280)
    want = P.predict(fused, img, conf=0.3, iou=0.45)[0]
    got = res[0].boxes.data.cpu()
    assert got.shape[1] == 6 and got.shape[0] <= 300
    assert torch.all(got[:-1, 4] >= got[1:, 4]) and torch.all(got[:, 4] > 0.3)
    assert torch.all(got[:, 0] >= 0) and torch.all(got[:, 2] <= 1280) and torch.all(got[:, 3] <= 853)
    frac, worst = match_detections(got, want, 0.3)
    print(f"config#1 geometry: {len(got)} detections (oracle {len(want)}), matched {frac:.2f}, median box delta {worst:.2f} px")
    assert frac >= 0.85, (frac, len(got), len(want))
    assert worst <= 4.0, worst      # bf16 network vs fp32 oracle: DFL logits differ by ~1e-2 -> a few px on 32-stride boxes


def test_postprocess_from_identical_head_is_exact(engines):
    """Feed the ORACLE the engine's own head logits: decode + NMS + scale_boxes must then agree to 0.5 px / exact order."""
    eng, fused, _ = engines("n")
    B, H, W = 2, 640, 640
    x = torch.rand(B, 3, H, W, generator=torch.Generator().manual_seed(11))
    net = eng.compiled(B, H, W)
    eng.preprocess_tensor(net, x.to("cuda:0").contiguous(), 1.0)
    eng.forward(net)
    orig = [(H, W)] * B
    rows = torch.tensor([[1.0, 0.0, 0.0, W, H]] * B, dtype=torch.float32, device="cuda:0")
    det, cnt, ncand = eng.postprocess(net, rows, 0.25, 0.7, 300)
    torch.cuda.synchronize()
    head = net.raw_head().cpu()                                       # [B,144,A]
    det_mod = fused.model[-1]
    feats, off = [], 0
    for h in net.head:
        hw = h.shape[1] * h.shape[2]
        feats.append(head[:, :, off:off + hw].reshape(B, 144, h.shape[1], h.shape[2]))
        off += hw
    with torch.no_grad():
        y = det_mod.decode(feats)
    want = P.non_max_suppression(y, 0.25, 0.7, max_det=300)
    for b in range(B):
        w = want[b].clone()
        if len(w):
            P.scale_boxes((H, W), w[:, :4], orig[b])
        n = int(cnt[b])
        got = det[b, :n].cpu()
        # decode arithmetic (expf / reduction order) may differ in the last ulp from torch's softmax, so scores can differ by
        # ~1e-7 and an exact tie could reorder; require same count, classes in order, boxes <= 0.5 px, scores <= 1e-5
        assert n == len(w), (n, len(w))
        assert torch.equal(got[:, 5], w[:, 5])
        assert (got[:, 4] - w[:, 4]).abs().max() <= 1e-5
        assert (got[:, :4] - w[:, :4]).abs().max() <= 0.5


def test_reference_harness_call_patterns(engines, tmp_path):
    """The call sites of the reference: SpeedBenchmark._benchmark_inference (benchmarks/speed_benchmark.py:322-350),
    DetectionDemo.detect_image consumption (demos/detection_demo.py:96-132), draw_detections accessors."""
    eng = engines("n")[0]
    path = tmp_path / "w.pt"
    eng.save(path)
    model = YOLO11Model(model_path=str(path), device="cuda:0", verbose=False)
    assert model.get_model_info()["total_parameters"] == 2624080
    model.model.eval()
    x = torch.randn(4, 3, 320, 320).cuda()
    with torch.no_grad():
        out = model.predict(x, verbose=False)
    assert len(out) == 4
    m = model.benchmark(x, num_runs=3, warmup_runs=1)
    assert set(m) == {"avg_inference_time", "min_inference_time", "max_inference_time", "fps"}
    frame = (np.random.default_rng(1).integers(0, 256, (720, 1280, 3))).astype(np.uint8)
    results = model.predict(frame, conf=0.3, iou=0.45, show=False, save=False)
    result = results[0]
    num = len(result.boxes) if result.boxes else 0
    info = []
    if result.boxes:
        for box in result.boxes:
            info.append({"class_id": int(box.cls[0]), "confidence": float(box.conf[0]), "bbox": box.xyxy[0].tolist()})
        x1, y1, x2, y2 = result.boxes.xyxy[0].cpu().numpy().astype(int)
        assert result.names[int(result.boxes.cls[0].cpu().numpy())] is not None
    assert len(info) == num
    assert set(result.speed) == {"preprocess", "inference", "postprocess"}
    # list-of-images source and file-path source
    import cv2
    p = tmp_path / "f.jpg"
    cv2.imwrite(str(p), frame)
    r2 = model.predict([str(p), frame], conf=0.5, verbose=False)
    assert len(r2) == 2 and r2[0].path == str(p)


def test_repo_image_fixture_if_present(engines):
    """tests/golden/image_small.jpg is a 320x480 crop-scale copy of the reference's image.jpg made by make_golden.py
    (the 1280x853 original cannot travel to the GPU box: /root/reference does not exist there)."""
    p = ROOT / "tests" / "golden" / "image_small.jpg"
    if not p.exists():
        pytest.skip("fixture not generated")
    import cv2
    eng, _, fused = engines("n")     # oracle = reference pipeline around the same-storage network
    img = cv2.imread(str(p))
    res = eng.predict(str(p), conf=0.25, iou=0.45, verbose=False)[0]
    want = P.predict(fused, img, conf=0.25, iou=0.45)[0]
    frac, worst = match_detections(res.boxes.data.cpu(), want, 0.25)
    print(f"image_small.jpg: {len(res.boxes)} detections (oracle {len(want)}), matched {frac:.2f}, median box delta {worst:.2f} px")
    assert frac >= 0.85 and worst <= 4.0, (frac, worst)


@pytest.mark.parametrize("scale,B,H,W", [("n", 2, 320, 320), ("s", 1, 224, 320), ("m", 1, 160, 160)])
def test_every_conv_of_the_plan_matches_torch(engines, oracle_models, scale, B, H, W):
    """The tight, amplification-free check: after a forward pass, every conv / depthwise op of the plan is re-run alone on its
    ACTUAL input buffer and compared with torch fp32 conv2d on that same bf16 input and the same bf16 weights
    (|err| <= 1e-2*|want| + 2e-2, i.e. about one bf16 ulp).  Covers every layer shape the network really uses
    (1x1 / 3x3 / stride 2, SW32/64/128 K chunks, residuals, in-place residuals, channel-slice inputs/outputs, fp32 logits)."""
    if scale in ("n", "s"):
        eng = engines(scale)[0]
    else:
        from yolo_infer_b200 import topology as T
        eng = YOLO.from_state_dict(T.synthetic_state_dict(scale, seed=1), scale).to("cuda:0")
    net = eng.compiled(B, H, W)
    x = torch.rand(B, 3, H, W, generator=torch.Generator().manual_seed(1)).to("cuda:0")
    eng.preprocess_tensor(net, x, 1.0)
    eng.forward(net)
    torch.cuda.synchronize()
    s = torch.cuda.current_stream().cuda_stream
    checked = 0
    for i, op in enumerate(net.ops):
        if op.kind not in ("conv", "dwconv"):
            continue
        pc = eng._packed[op.name]
        vin, vout, vres = op.inp, op.out, op.res
        xin = vin.t[..., vin.off:vin.off + vin.c].float().permute(0, 3, 1, 2).contiguous()
        res = vres.t[..., vres.off:vres.off + vres.c].float().permute(0, 3, 1, 2).clone() if vres is not None else None
        net.run_range(i, i + 1, s)
        torch.cuda.synchronize()
        got = vout.t[..., vout.off:vout.off + vout.c].float().permute(0, 3, 1, 2)
        if pc.depthwise:
            w = pc.w.float().t().reshape(pc.c2, 1, 3, 3)
            want = torch.nn.functional.conv2d(xin, w, pc.b, padding=1, groups=pc.c2)
        else:
            if pc.k == 2:
                from yolo_infer_b200.network import dense_k2_weights
                w = dense_k2_weights(pc).permute(0, 3, 1, 2)     # also from the compact per-stage packing (s, m: s2d_block)
            else:
                w = pc.w.float().view(pc.c2, pc.k, pc.k, pc.c1).permute(0, 3, 1, 2)
            if pc.k == 2:   # space-to-depth form of model.1: taps at {-1, 0}^2 = zero padding on the top/left only
                want = torch.nn.functional.conv2d(torch.nn.functional.pad(xin, (1, 0, 1, 0)), w, pc.b)
            else:
                want = torch.nn.functional.conv2d(xin, w, pc.b, stride=pc.s, padding=pc.k // 2)
        if res is not None and op.res_mode == cabi.RES_PRE_UP2:   # folded Upsample+Concat: up2(W_up . p + b) enters before SiLU
            want = want + torch.nn.functional.interpolate(res, scale_factor=2, mode="nearest")
        if pc.act:
            want = torch.nn.functional.silu(want)
        if res is not None and op.res_mode != cabi.RES_PRE_UP2:
            want = want + res
        err = (got - want).abs()
        tol = 1e-2 * want.abs() + 2e-2
        assert torch.all(err <= tol), f"{op.name}: max err {err.max().item():.4g} (tol {tol.flatten()[err.argmax()].item():.4g})"
        checked += 1
    assert checked >= 80


def test_cuda_graph_pipeline_equals_eager_predict(engines):
    """uint8 batch tensors take the CUDA-graph pipeline (one replay per call); results must be bit-identical to the eager
    list-of-ndarrays path (same kernels, same inputs), call after call, from pinned host and from device memory."""
    eng = engines("n")[0]
    g = torch.Generator().manual_seed(21)
    frames = torch.randint(0, 256, (3, 360, 640, 3), dtype=torch.uint8, generator=g)
    eager = eng.predict([f.numpy() for f in frames], conf=0.3, iou=0.45, verbose=False, graph=False)
    for src in (frames.pin_memory(), frames.cuda(), frames.pin_memory()):
        graphed = eng.predict(src, conf=0.3, iou=0.45, verbose=False)
        assert len(graphed) == 3
        for a, b in zip(eager, graphed):
            assert a.orig_shape == b.orig_shape == (360, 640)
            assert torch.equal(a.boxes.data.cpu(), b.boxes.data.cpu())
            # host mirror fetched with the batch's single D2H == the device rows; .cpu()/.numpy()/indexing use it
            assert b.boxes.data.is_cuda and not b.cpu().boxes.data.is_cuda
            assert torch.equal(b.cpu().boxes.data, b.boxes.data.cpu())
            assert np.array_equal(b.boxes.numpy().data, b.boxes.data.cpu().numpy())
            if len(b.boxes):
                assert torch.equal(b.boxes[0].cpu().data, b.boxes.data[:1].cpu())
    other = torch.randint(0, 256, (3, 360, 640, 3), dtype=torch.uint8, generator=g)
    r2 = eng.predict(other.pin_memory(), conf=0.3, iou=0.45, verbose=False)
    e2 = eng.predict([f.numpy() for f in other], conf=0.3, iou=0.45, verbose=False, graph=False)
    assert all(torch.equal(a.boxes.data.cpu(), b.boxes.data.cpu()) for a, b in zip(e2, r2))


def test_array_and_file_sources_use_the_graph_pipeline(engines, tmp_path):
    """BGR arrays / image files of one shape (the demo's per-frame `predict(frame, conf=, iou=)` loop, reference
    demos/detection_demo.py:190-196) run through the CUDA-graph pipeline: same rows as the eager launches, `orig_img`, `path`
    and `orig_shape` as the reference's Results carry them; mixed shapes fall back to the eager path."""
    import cv2
    eng = engines("n")[0]
    rng = np.random.default_rng(12)
    frames = [rng.integers(0, 256, (360, 640, 3), dtype=np.uint8) for _ in range(2)]
    for src in (frames[0], frames):
        lst = src if isinstance(src, list) else [src]
        want = eng.predict(src, conf=0.3, iou=0.45, verbose=False, graph=False)
        n_pipes = len(getattr(eng, "_pipes", {}))
        got = eng.predict(src, conf=0.3, iou=0.45, verbose=False)
        assert len(got) == len(want) == len(lst)
        for a, b, im in zip(want, got, lst):
            assert torch.equal(a.boxes.data.cpu(), b.boxes.data.cpu())
            assert b.orig_img is im and b.orig_shape == (360, 640) and b.path == a.path
        assert len(eng._pipes) >= max(n_pipes, 1)
    # a video-like loop reuses one pipeline, frame after frame
    n_pipes = len(eng._pipes)
    for _ in range(3):
        f = rng.integers(0, 256, (360, 640, 3), dtype=np.uint8)
        a = eng.predict(f, conf=0.3, iou=0.45, verbose=False, graph=False)[0]
        b = eng.predict(f, conf=0.3, iou=0.45, verbose=False)[0]
        assert torch.equal(a.boxes.data.cpu(), b.boxes.data.cpu())
    assert len(eng._pipes) == n_pipes
    # files keep their paths; mixed shapes take the eager path
    p0, p1 = tmp_path / "a.png", tmp_path / "b.png"
    cv2.imwrite(str(p0), frames[0])
    cv2.imwrite(str(p1), frames[1][:200, :320])
    r = eng.predict(str(p0), conf=0.3, iou=0.45, verbose=False)[0]
    assert r.path == str(p0) and torch.equal(r.boxes.data.cpu(), eng.predict(frames[0], conf=0.3, iou=0.45, verbose=False, graph=False)[0].boxes.data.cpu())
    mixed = eng.predict([str(p0), str(p1)], conf=0.3, iou=0.45, verbose=False)
    assert [m.path for m in mixed] == [str(p0), str(p1)] and mixed[1].orig_shape == (200, 320)


def test_fused_uint8_stem_pipeline_equals_letterbox_pipeline(engines, monkeypatch):
    """Frames already at network resolution: the pipeline skips the letterbox launch (the stem reads the uint8 frames) - for
    device-resident, host-fed and chunked host-fed batches the detections are bit-identical to the eager path, which runs
    y11_letterbox + the bf16 stem, and the two kinds of pipeline can share one plan without disturbing each other."""
    from yolo_infer_b200 import engine as E
    eng = engines("n")[0]
    g = torch.Generator().manual_seed(77)
    for B, hw in ((3, 320), (16, 256)):
        frames = torch.randint(0, 256, (B, hw, hw, 3), dtype=torch.uint8, generator=g)
        eager = eng.predict([f.numpy() for f in frames], conf=0.3, iou=0.45, imgsz=hw, verbose=False, graph=False)
        for src in (frames.cuda(), frames.pin_memory()):
            got = eng.predict(src, conf=0.3, iou=0.45, imgsz=hw, verbose=False)
            pipe = eng.pipeline(B, hw, hw, hw, True, 0.3, 0.45, 300)
            assert pipe.fused_stem and pipe.launches == pipe.net.n_launches + 2
            assert len(got) == B and sum(len(r.boxes) for r in got) > 0
            for a, b in zip(eager, got):
                assert torch.equal(a.boxes.data.cpu(), b.boxes.data.cpu())
        again = eng.predict([f.numpy() for f in frames], conf=0.3, iou=0.45, imgsz=hw, verbose=False, graph=False)   # eager after fused
        assert all(torch.equal(a.boxes.data.cpu(), b.boxes.data.cpu()) for a, b in zip(eager, again))
    # padded / resized frames keep the letterbox launch
    assert not eng.pipeline(2, 360, 640, 640, True, 0.3, 0.45, 300).fused_stem
    monkeypatch.setattr(E, "FUSE_U8_STEM", False)
    eng._pipes.clear()
    assert not eng.pipeline(3, 320, 320, 320, True, 0.3, 0.45, 300).fused_stem
    eng._pipes.clear()


def test_chunked_host_pipeline_equals_eager_predict(engines):
    """Host-fed batches with B % 4 == 0, B >= 16 cross PCIe in four chunks while layers 0-4 of the previous chunk run
    (one graph per chunk + one for the rest): results must be bit-identical to the eager path, call after call."""
    eng = engines("n")[0]
    g = torch.Generator().manual_seed(33)
    for rep in range(2):
        frames = torch.randint(0, 256, (16, 192, 320, 3), dtype=torch.uint8, generator=g)
        eager = eng.predict([f.numpy() for f in frames], conf=0.3, iou=0.45, verbose=False, graph=False)
        pipe = eng.pipeline(16, 192, 320, 640, True, 0.3, 0.45, 300)
        assert pipe.chunks == 4 and len(pipe.graphs) == 5 and len(pipe.net.prefix_ranges) == 4, (pipe.chunks, len(pipe.graphs))
        for src in (frames.pin_memory(), frames.cuda()):
            got = eng.predict(src, conf=0.3, iou=0.45, verbose=False)
            assert len(got) == 16
            for a, b in zip(eager, got):
                assert torch.equal(a.boxes.data.cpu(), b.boxes.data.cpu())


def test_split_host_batch_equals_eager_predict(engines, monkeypatch):
    """Optional mode (Y11_SPLIT_HOST=1): host-fed batches of >= 32 frames as two half-batch pipelines on two streams (second
    half on the PCIe bus while the first half computes): same rows, same order as the eager path, call after call."""
    from yolo_infer_b200 import engine as E
    monkeypatch.setattr(E, "SPLIT_HOST_BATCH", True)
    eng = engines("n")[0]
    g = torch.Generator().manual_seed(44)
    for rep in range(3):
        frames = torch.randint(0, 256, (32, 160, 256, 3), dtype=torch.uint8, generator=g)
        eager = eng.predict([f.numpy() for f in frames], conf=0.3, iou=0.45, verbose=False, graph=False)
        got = eng.predict(frames.pin_memory(), conf=0.3, iou=0.45, verbose=False)
        assert len(got) == 32
        for a, b in zip(eager, got):
            assert torch.equal(a.boxes.data.cpu(), b.boxes.data.cpu())
            assert torch.equal(b.cpu().boxes.data, b.boxes.data.cpu())
    assert any(p.B == 16 and getattr(p.net, "replica", 0) == 1 for p in eng._pipes.values()), "the half-batch replica pipeline was not used"


def test_val_on_a_synthetic_dataset(engines, tmp_path):
    """`YOLO11Model.val(data)` (reference core/model.py:180-195, read back at core/validator.py:339-359): a dataset whose
    labels are the engine's own confident detections must score a high mAP50; shuffled class labels must not."""
    import cv2
    eng = engines("n")[0]
    rng = np.random.default_rng(5)
    root = tmp_path / "ds"
    (root / "images" / "val").mkdir(parents=True)
    (root / "labels" / "val").mkdir(parents=True)
    n_lab = 0
    for i in range(6):
        img = rng.integers(0, 256, (360, 640, 3), dtype=np.uint8)
        path = root / "images" / "val" / f"{i}.png"
        cv2.imwrite(str(path), img)                                  # png: lossless, the engine sees the same pixels
        # labels = everything the val-mode pipeline itself detects above conf 0.4 (same NMS settings as val uses)
        det = eng.predict(cv2.imread(str(path)), conf=0.4, iou=0.6, multi_label=True, verbose=False)[0].cpu().boxes.data.numpy()
        rows = []
        for x1, y1, x2, y2, _, c in det:
            rows.append(f"{int(c)} {(x1 + x2) / 2 / 640:.6f} {(y1 + y2) / 2 / 360:.6f} {(x2 - x1) / 640:.6f} {(y2 - y1) / 360:.6f}")
        n_lab += len(rows)
        (root / "labels" / "val" / f"{i}.txt").write_text("\n".join(rows) + "\n")
    assert n_lab >= 6, "calibrated synthetic weights should give confident detections"
    (root / "data.yaml").write_text(f"path: {root}\nval: images/val\nnames:\n" + "".join(f"  {k}: c{k}\n" for k in range(80)))
    model = YOLO11Model(model_path="yolo11n.yaml", device="cuda:0", verbose=False)
    model.model = eng                                                  # same weights as the labels were made with
    # rect=False: the labels above were made by `predict` (its letterbox geometry); a random-weight network is not translation
    # robust, so scoring them under the rect-batch canvases of val would measure the weights, not the path.  The rect path is
    # checked against the oracle's val path in tests/test_gpu_round2.py::test_val_rect_path_matches_the_oracle_val_path.
    m = model.val(data=str(root / "data.yaml"), batch=4, rect=False)
    assert m.n_images == 6 and m.n_labels == n_lab
    assert m.box.map50 > 0.9 and m.box.mr > 0.9 and 0.0 <= m.box.map <= 1.0 and 0.0 <= m.box.map75 <= 1.0
    assert set(m.speed) >= {"preprocess", "inference", "postprocess"} and m.speed["inference"] > 0
    assert "metrics/mAP50-95(B)" in m.results_dict
