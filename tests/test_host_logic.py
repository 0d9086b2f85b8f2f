"""CPU: host-side logic of the B200 path (geometry, topology, weight packing, result containers, boundary errors)."""
import math

import numpy as np
import pytest
import torch

from oracle import pipeline_ref as P
from oracle import yolo11_ref as R
from yolo_infer_b200 import topology as T
from yolo_infer_b200.engine import DetectionNet, infer_scale, letterbox_geometry, scale_geometry
from yolo_infer_b200.network import fold, pack_weights, pad16, qkv_permutation
from yolo_infer_b200.results import Boxes, Results


def test_letterbox_and_scale_geometry_match_oracle():
    rng = np.random.default_rng(0)
    shapes = [(853, 1280), (720, 1280), (1080, 1920), (640, 640), (480, 640), (333, 517), (100, 37), (1281, 641)]
    shapes += [tuple(int(v) for v in rng.integers(20, 2000, 2)) for _ in range(200)]
    for (h, w) in shapes:
        for auto in (True, False):
            nw, nh, top, bottom, left, right, H, W = P.letterbox_params(h, w, (640, 640), auto=auto)
            assert letterbox_geometry(h, w, (640, 640), auto) == (nw, nh, top, left, H, W)
            assert H % 32 == 0 and W % 32 == 0 if auto else (H, W) == (640, 640)
            assert scale_geometry((H, W), (h, w)) == P.scale_boxes_params((H, W), (h, w))


@pytest.mark.parametrize("scale", list(T.SCALES))
def test_topology_matches_oracle_state_dict(scale):
    sd = R.DetectionModel(scale).state_dict()
    shapes = T.param_shapes(scale)
    ref = {k: tuple(v.shape) for k, v in sd.items() if not k.endswith("num_batches_tracked")}
    assert shapes == ref
    assert T.count_parameters(scale) == R.count_params(R.DetectionModel(scale))
    assert infer_scale(sd) == scale


def test_default_state_dict_loads_into_oracle():
    sd = T.default_state_dict("n", seed=3)
    m = R.DetectionModel("n")
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected and all(k.endswith("num_batches_tracked") for k in missing)
    assert T.anchors_for(640, 640) == 8400 and T.anchors_for(448, 640) == 5880 and T.anchors_for(1280, 1280) == 33600


def test_fold_matches_oracle_fuse(oracle_models):
    m, sd = oracle_models("n")
    fused = R.DetectionModel("n")
    fused.load_state_dict(sd)
    fused.eval().fuse()
    fsd = fused.state_dict()
    for cp in T.conv_params("n"):
        w, b = fold(sd, cp)
        key = f"{cp.prefix}.conv" if cp.bn else cp.prefix
        assert torch.allclose(w, fsd[f"{key}.weight"], rtol=1e-6, atol=1e-7)
        assert torch.allclose(b, fsd[f"{key}.bias"], rtol=1e-5, atol=1e-6)


def test_pack_weights_layout(oracle_models):
    _, sd = oracle_models("n")
    packed = pack_weights("n", 80, sd, torch.device("cpu"))
    cps = {cp.prefix: cp for cp in T.conv_params("n")}
    # dense 3x3 with padded 8-channel hidden tensor (yolo11n layer 2 bottleneck)
    pc = packed["model.2.m.0.cv1"]
    cp = cps["model.2.m.0.cv1"]
    assert (cp.c1, cp.c2) == (16, 8) and pc.w.shape == (16, 9 * 16) and pc.c2 == 16
    w, b = fold(sd, cp)
    wk = pc.w.float().view(16, 3, 3, 16)
    assert torch.allclose(wk[:8].permute(0, 3, 1, 2), w, rtol=1e-2, atol=1e-3)  # bf16 rounding
    assert torch.all(wk[8:] == 0) and torch.all(pc.b[8:] == 0)
    # depthwise is tap-major
    pc = packed["model.23.cv3.0.0.0"]
    w, _ = fold(sd, cps["model.23.cv3.0.0.0"])
    assert pc.depthwise and pc.w.shape == (9, w.shape[0])
    assert torch.allclose(pc.w.float().t().reshape(-1, 1, 3, 3), w, rtol=1e-2, atol=1e-3)
    # stem K order (kh, kw, c)
    pc = packed["model.0"]
    w, _ = fold(sd, cps["model.0"])
    assert torch.allclose(pc.w.float().view(-1, 3, 3, 3).permute(0, 3, 1, 2), w, rtol=1e-2, atol=1e-3)


def test_upsample_concat_fold_split(oracle_models):
    """Upsample -> Concat -> C3k2.cv1 folding: the two neck chains are found for every scale, and the split weights
    reproduce cv1 on the concatenated input: W . cat(up2(p), skip) + b == up2(W_up . p + b) + W_skip . skip."""
    from yolo_infer_b200.network import upsample_folds
    for scale in "nsmlx":
        assert upsample_folds(scale) == {13: (10, 6, 11, 12), 16: (13, 4, 14, 15)}
    _, sd = oracle_models("n")
    packed = pack_weights("n", 80, sd, torch.device("cpu"))
    full, up, skip = packed["model.13.cv1"], packed["model.13.cv1#up"], packed["model.13.cv1#skip"]
    assert up.c1 + skip.c1 == full.c1 and up.c2 == skip.c2 == full.c2 and up.act == 0 and skip.act == full.act
    assert torch.equal(torch.cat((up.w, skip.w), 1), full.w) and torch.equal(up.b, full.b) and torch.all(skip.b == 0)
    g = torch.Generator().manual_seed(0)
    p, s_ = torch.randn(1, up.c1, 4, 6, generator=g), torch.randn(1, skip.c1, 8, 12, generator=g)
    up2 = lambda t: torch.nn.functional.interpolate(t, scale_factor=2, mode="nearest")  # noqa: E731
    conv = lambda x, pc: torch.nn.functional.conv2d(x, pc.w.float().view(pc.c2, pc.c1, 1, 1), pc.b)  # noqa: E731
    assert torch.allclose(conv(torch.cat((up2(p), s_), 1), full), up2(conv(p, up)) + conv(s_, skip), atol=1e-5)


def test_qkv_permutation_groups_heads():
    perm = qkv_permutation(128, 2, 32, 64)
    assert sorted(perm.tolist()) == list(range(256))
    assert perm[:32].tolist() == list(range(0, 32)) and perm[32:64].tolist() == list(range(128, 160))   # Q h0, Q h1
    assert perm[64:96].tolist() == list(range(32, 64))                                                    # K h0
    assert perm[128:192].tolist() == list(range(64, 128))                                                 # V h0
    assert pad16(8) == 16 and pad16(80) == 80 and pad16(1) == 16


def test_boxes_results_contract():
    data = torch.tensor([[1., 2., 30., 40., .9, 3.], [5., 6., 70., 80., .8, 1.]])
    r = Results(None, "x.jpg", {i: str(i) for i in range(80)}, data, (100, 200))
    assert len(r.boxes) == 2 and bool(r.boxes)
    assert r.boxes.xyxy.shape == (2, 4) and r.boxes.conf.tolist() == pytest.approx([.9, .8]) and r.boxes.cls.tolist() == [3., 1.]
    # the exact access pattern of reference utils/visualization.py:52-68
    x1, y1, x2, y2 = r.boxes.xyxy[0].cpu().numpy().astype(int)
    assert (x1, y1, x2, y2) == (1, 2, 30, 40)
    assert float(r.boxes.conf[0].cpu().numpy()) == pytest.approx(.9) and int(r.boxes.cls[0].cpu().numpy()) == 3
    rows = [b for b in r.boxes]                     # demos/detection_demo.py:123-132 iterates 1-row Boxes
    assert len(rows) == 2 and rows[1].xyxy.shape == (1, 4) and float(rows[1].conf[0]) == pytest.approx(.8)
    assert r.boxes.xywh[0].tolist() == [15.5, 21., 29., 38.]
    assert r.boxes.xyxyn[0].tolist() == pytest.approx([1 / 200, 2 / 100, 30 / 200, 40 / 100])
    empty = Results(None, "x.jpg", {}, torch.zeros((0, 6)), (10, 10))
    assert len(empty.boxes) == 0 and not empty.boxes
    assert r.names[3] == "3" and r.summary()[0]["class"] == 3


def test_detection_net_parameter_surface():
    sd = T.default_state_dict("n")
    net = DetectionNet("n", 80, sd, {})
    assert sum(p.numel() for p in net.parameters()) == 2624080          # get_model_info contract
    assert net.eval() is net
    bad = dict(sd)
    bad.pop("model.0.conv.weight")
    with pytest.raises(KeyError):
        DetectionNet("n", 80, bad, {})


def test_model_wrapper_validation_errors_match_reference():
    from yolo_infer_b200 import YOLO11Model
    with pytest.raises(ValueError, match="Unsupported task"):
        YOLO11Model(task="banana", device="cuda")
    with pytest.raises(ValueError, match="Unsupported size"):
        YOLO11Model(size="q", device="cuda")
    with pytest.raises(NotImplementedError):
        YOLO11Model(task="segment", device="cuda")
