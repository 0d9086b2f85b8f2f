"""CPU: pin the oracle (oracle/*.py) against every known answer available for this path."""
import json
import math
from pathlib import Path

import cv2
import numpy as np
import pytest
import torch
import torchvision

from oracle import letterbox_ref as LB
from oracle import pipeline_ref as P
from oracle import yolo11_ref as R

GOLD = Path(__file__).parent / "golden"

PARAMS = {"n": 2624080, "s": 9458752, "m": 20114688, "l": 25372160, "x": 56966176}  # upstream yolo11.yaml comments
GFLOPS = {"n": 6.5, "s": 21.5, "m": 68.0, "l": 86.9, "x": 194.9}                   # upstream model card (fused)


@pytest.mark.parametrize("scale", list(PARAMS))
def test_param_counts_match_published(scale):
    assert R.count_params(R.DetectionModel(scale)) == PARAMS[scale]


@pytest.mark.parametrize("scale", ["n", "s", "m"])
def test_conv_gflops_match_published(scale):
    g = R.conv_flops(R.DetectionModel(scale).fuse()) / 1e9
    assert abs(g - GFLOPS[scale]) < 0.06, g


def test_output_shapes_and_anchor_counts():
    m = R.DetectionModel("n").eval()
    with torch.no_grad():
        for (h, w, a) in [(640, 640, 8400), (448, 640, 5880), (384, 640, 5040)]:
            y, feats = m(torch.zeros(1, 3, h, w))
            assert y.shape == (1, 84, a)
            assert [f.shape[1] for f in feats] == [144] * 3


def test_default_init_is_bias_regime():
    # with ultralytics-style init the cls score is set by bias_init: 5/80/(640/s)^2
    m = R.DetectionModel("n")
    det = m.model[-1]
    for seq, s in zip(det.cv3, (8, 16, 32)):
        p = torch.sigmoid(seq[-1].bias[0]).item()
        assert abs(p - 5 / 80 / (640 / s) ** 2) / p < 1e-3


def test_fuse_is_equivalent(oracle_models):
    m, sd = oracle_models("n")
    x = torch.rand(1, 3, 64, 64)
    with torch.no_grad():
        y0, _ = m(x)
        f = R.DetectionModel("n")
        f.load_state_dict(sd)
        f.eval().fuse()
        y1, _ = f(x)
    assert torch.allclose(y0, y1, rtol=1e-3, atol=1e-3)


# ----------------------------------------------------------------------------- letterbox
@pytest.mark.parametrize("h0,w0,rect,exp", [
    (853, 1280, True, (640, 426, 11, 11, 0, 0, 448, 640)),     # image.jpg under rect=True
    (720, 1280, True, (640, 360, 12, 12, 0, 0, 384, 640)),     # 720p video frame
    (853, 1280, False, (640, 426, 107, 107, 0, 0, 640, 640)),  # square mode
    (1080, 1920, True, (640, 360, 12, 12, 0, 0, 384, 640)),
])
def test_letterbox_geometry_kat(h0, w0, rect, exp):
    assert P.letterbox_params(h0, w0, (640, 640), auto=rect) == exp


def test_scale_boxes_kat():
    gain, px, py = P.scale_boxes_params((448, 640), (853, 1280))
    assert (gain, px, py) == (0.5, 0, 11)


@pytest.mark.parametrize("h,w,dh,dw", [(853, 1280, 426, 640), (720, 1280, 360, 640), (1080, 1920, 360, 640),
                                       (300, 400, 480, 640), (333, 517, 412, 640), (1280, 1280, 640, 640),
                                       (100, 37, 640, 237), (641, 1283, 320, 640), (64, 64, 64, 64), (7, 5, 3, 2)])
def test_fixed_point_resize_restatement_matches_cv2_bit_for_bit(h, w, dh, dw):
    rng = np.random.default_rng(h * 7 + w)
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    ref = cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR)
    assert np.array_equal(LB.resize_linear_u8(img, dw, dh), ref)


def test_letterbox_restatement_matches_cv2_pipeline():
    rng = np.random.default_rng(1)
    img = rng.integers(0, 256, (853, 1280, 3), dtype=np.uint8)
    nw, nh, top, bottom, left, right, H, W = P.letterbox_params(853, 1280, (640, 640), auto=True)
    assert np.array_equal(LB.letterbox_u8(img, nw, nh, top, left, H, W), P.letterbox(img, (640, 640), auto=True))


# ----------------------------------------------------------------------------- NMS semantics
def test_nms_semantics_probes():
    # IoU == thr is kept (strict >); ties -> lower index first; output score-descending (SURVEY 8c iv)
    boxes = torch.tensor([[0, 0, 10, 10], [0, 0, 10, 5], [20, 20, 30, 30], [20, 20, 30, 30]], dtype=torch.float32)
    scores = torch.tensor([0.9, 0.8, 0.7, 0.7])
    keep = torchvision.ops.nms(boxes, scores, 0.5)          # IoU(0,1) = 0.5 exactly -> kept
    assert keep.tolist() == [0, 1, 2]
    keep = torchvision.ops.nms(boxes, scores, 0.49)
    assert keep.tolist() == [0, 2]


def test_non_max_suppression_layout():
    g = torch.Generator().manual_seed(0)
    A = 200
    y = torch.zeros(2, 84, A)
    y[:, 0:2] = torch.rand(2, 2, A, generator=g) * 600
    y[:, 2:4] = torch.rand(2, 2, A, generator=g) * 100 + 4
    y[:, 4:] = torch.rand(2, 80, A, generator=g) * 0.6
    out = P.non_max_suppression(y, 0.25, 0.45, max_det=50)
    assert len(out) == 2
    for o in out:
        assert o.shape[1] == 6 and o.shape[0] <= 50
        assert torch.all(o[:-1, 4] >= o[1:, 4])      # score-descending
        assert torch.all(o[:, 4] > 0.25)
    ml = P.non_max_suppression(y, 0.25, 0.45, multi_label=True, max_det=300)
    assert ml[0].shape[0] >= out[0].shape[0]


# ----------------------------------------------------------------------------- golden fixtures
def test_golden_fixtures_reproduce():
    """tests/golden/*.npz were produced by tests/golden/make_golden.py from the importable parts of the real
    reference stack (cv2, torchvision.ops.nms) + this oracle; they must reproduce exactly."""
    g = np.load(GOLD / "letterbox_small.npz")
    out = P.letterbox(g["img"], (64, 64), auto=True)
    assert np.array_equal(out, g["out_rect"])
    n = np.load(GOLD / "nms_small.npz")
    keep = torchvision.ops.nms(torch.from_numpy(n["boxes"]) + torch.from_numpy(n["cls"])[:, None] * 7680,
                               torch.from_numpy(n["scores"]), float(n["iou"]))
    assert keep.tolist() == n["keep"].tolist()
    d = np.load(GOLD / "decode_small.npz")
    det = R.Detect(nc=80, ch=(16, 16, 16))
    feats = [torch.from_numpy(d[f"f{i}"]) for i in range(3)]
    y = det.decode(feats)
    assert torch.allclose(y, torch.from_numpy(d["y"]), rtol=1e-6, atol=1e-6)
