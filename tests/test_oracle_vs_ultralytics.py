"""Pins the oracle against ultralytics ITSELF - the package the reference delegates the whole detect path to
(/root/reference/core/model.py:18,110,118-133; requirements.txt:4 `ultralytics>=8.0.0`).

ultralytics is not installed in the build image or on the GPU boxes and cannot be fetched (no network), so every test
here is SKIPPED there and the oracle's header keeps saying "parity unpinned".  On any machine where `import ultralytics`
works (CPU is enough) the module un-skips by itself and checks, on the same weights and inputs:

  * the network: oracle `DetectionModel` loaded with the state_dict of `ultralytics.nn.tasks.DetectionModel('yolo11{n,s}.yaml')`
    -> identical parameter names/shapes and raw head + decoded outputs equal to fp32 round-off (same op order, so <= 1e-5);
  * `LetterBox` (cv2 resize + border) bit for bit, `non_max_suppression` (single- and multi-label) and `scale_boxes`.

Nothing in this file touches the GPU path; it validates test infrastructure (oracle/ is imported only by tests, smoke and
bench.py's CPU legs).
"""
import numpy as np
import pytest
import torch

ultralytics = pytest.importorskip("ultralytics", reason="ultralytics not installed: the oracle stays 'parity unpinned' here")

from oracle import pipeline_ref as P  # noqa: E402
from oracle import yolo11_ref as R  # noqa: E402


def _ul_model(scale: str):
    from ultralytics.nn.tasks import DetectionModel
    torch.manual_seed(0)
    m = DetectionModel(f"yolo11{scale}.yaml", ch=3, nc=80, verbose=False)
    # default init gives bias-determined class scores; randomise BN statistics so that every layer matters
    g = torch.Generator().manual_seed(1)
    for mod in m.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.weight.data = torch.rand(mod.weight.shape, generator=g) * 0.3 + 0.2
            mod.bias.data = torch.randn(mod.bias.shape, generator=g) * 0.3 + 1.5
            mod.running_mean.data = torch.randn(mod.running_mean.shape, generator=g) * 0.1
            mod.running_var.data = torch.rand(mod.running_var.shape, generator=g) * 0.5 + 0.75
    return m.eval()


@pytest.mark.parametrize("scale", ["n", "s"])
def test_network_outputs_equal_ultralytics(scale):
    ul = _ul_model(scale)
    sd = {k: v.clone() for k, v in ul.state_dict().items()}
    mine = R.DetectionModel(scale)
    missing, unexpected = mine.load_state_dict(sd, strict=False)
    assert not [k for k in missing if "num_batches_tracked" not in k], missing
    assert not [k for k in unexpected if "num_batches_tracked" not in k and "dfl" not in k], unexpected
    assert R.count_params(mine) == sum(p.numel() for p in ul.parameters())
    mine.eval()
    x = torch.rand(2, 3, 320, 448, generator=torch.Generator().manual_seed(2))
    with torch.no_grad():
        y_ul, feats_ul = ul(x)
        y, feats = mine(x)
    for a, b in zip(feats, feats_ul):
        assert a.shape == b.shape
        assert float((a - b).abs().max()) <= 1e-4 * float(b.abs().max())
    assert y.shape == y_ul.shape
    assert float((y - y_ul).abs().max()) <= 1e-3
    # fused (BN folded) model: what predict/val run
    ul.fuse()
    mine.fuse()
    with torch.no_grad():
        y2_ul, _ = ul(x)
        y2, _ = mine(x)
    assert float((y2 - y2_ul).abs().max()) <= 1e-3


def test_letterbox_nms_scale_boxes_equal_ultralytics():
    from ultralytics.data.augment import LetterBox
    from ultralytics.utils import ops
    try:
        from ultralytics.utils.nms import non_max_suppression as ul_nms      # newer releases
    except Exception:
        ul_nms = ops.non_max_suppression
    rng = np.random.default_rng(3)
    for (h0, w0), auto in [((720, 1280), True), ((853, 1280), True), ((480, 640), False), ((1080, 1920), True), ((333, 500), False)]:
        img = rng.integers(0, 256, (h0, w0, 3), dtype=np.uint8)
        want = LetterBox((640, 640), auto=auto, stride=32)(image=img)
        got = P.letterbox(img, (640, 640), auto=auto, stride=32)
        assert got.shape == want.shape and np.array_equal(got, want)
    g = torch.Generator().manual_seed(4)
    A = 2100
    pred = torch.zeros(2, 84, A)
    pred[:, 0:2] = torch.rand(2, 2, A, generator=g) * 600 + 20
    pred[:, 2:4] = torch.rand(2, 2, A, generator=g) * 120 + 4
    pred[:, 4:] = torch.rand(2, 80, A, generator=g) ** 6
    for multi, conf, iou in [(False, 0.25, 0.7), (False, 0.25, 0.45), (True, 0.05, 0.6)]:
        want = ul_nms(pred.clone(), conf, iou, multi_label=multi, max_det=300)
        got = P.non_max_suppression(pred.clone(), conf, iou, multi_label=multi, max_det=300)
        for a, b in zip(got, want):
            assert a.shape == b.shape and torch.equal(a, b)
    boxes = torch.rand(50, 4, generator=g) * 600
    for shape0 in [(720, 1280), (333, 500), (1080, 1920)]:
        want = ops.scale_boxes((384, 640), boxes.clone(), shape0)
        got = P.scale_boxes((384, 640), boxes.clone(), shape0)
        assert torch.equal(got, want)
