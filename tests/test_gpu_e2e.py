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


@pytest.mark.parametrize("scale,B,H,W", [("n", 2, 640, 640), ("n", 1, 448, 640), ("s", 2, 640, 640), ("n", 3, 384, 640)])
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
    print(f"head rel-L2 vs bf16-storage oracle: box {e_box:.3e} cls {e_cls:.3e}; vs fp32 oracle: box {f_box:.3e} cls {f_cls:.3e}")
    assert e_box <= 1e-2 and e_cls <= 1e-2, (e_box, e_cls)
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
    assert rel_l2(outs[0], outs[1]) <= 5e-3


def match_detections(got: torch.Tensor, want: torch.Tensor):
    """Greedy one-to-one match on (class equal, IoU); returns matched pairs count and max box delta among matches."""
    from torchvision.ops import box_iou
    if not len(got) or not len(want):
        return 0, 0.0
    iou = box_iou(got[:, :4], want[:, :4])
    iou[got[:, 5, None] != want[None, :, 5]] = 0
    n, worst = 0, 0.0
    used = set()
    for i in range(len(got)):
        j = int(iou[i].argmax())
        if iou[i, j] > 0.9 and j not in used:
            used.add(j)
            n += 1
            worst = max(worst, float((got[i, :4] - want[j, :4]).abs().max()))
    return n, worst


def test_predict_image_jpg_like_config1(engines, oracle_models):
    """Config #1 geometry: an 853x1280 BGR frame -> 448x640 under rect=True; demo iou 0.45 (conf 0.3: the calibrated
    random weights put almost no score above the demo's 0.5)."""
    eng, fused, _ = engines("n")
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
    n, worst = match_detections(got, want)
    assert n >= 0.8 * max(len(want), 1) and n >= 0.8 * len(got), (n, len(got), len(want))
    assert worst <= 8.0, worst      # bf16 network vs fp32 oracle: DFL logits differ by ~1e-2 -> a few px on 32-stride boxes


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
    eng, fused, _ = engines("n")
    img = cv2.imread(str(p))
    res = eng.predict(str(p), conf=0.25, iou=0.45, verbose=False)[0]
    want = P.predict(fused, img, conf=0.25, iou=0.45)[0]
    n, worst = match_detections(res.boxes.data.cpu(), want)
    assert n >= 0.8 * max(len(want), 1)
