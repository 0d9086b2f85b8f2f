"""GPU (-m gpu): the device rasteriser and the serialisers against the reference's cv2 loop (oracle/draw_ref.py =
/root/reference/utils/visualization.py:18-106, :367-436 restated; cv2 itself is the reference's arithmetic)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import draw_ref as D  # noqa: E402
from yolo_infer_b200.draw import DetectionRasteriser, draw_detections, save_detection_results  # noqa: E402
from yolo_infer_b200.results import Results  # noqa: E402

COCO_LIKE = {i: n for i, n in enumerate(["person", "bicycle", "car", "motorcycle", "airplane", "bus", "train", "truck", "boat",
                                         "traffic light", "fire hydrant", "stop sign", "dog", "cell phone", "teddy bear"])}


def random_results(rng, n, h, w, names, margin):
    x1 = rng.uniform(margin, w - 80, n)
    y1 = rng.uniform(margin, h - 40, n)
    bw, bh = rng.uniform(3, 300, n), rng.uniform(3, 200, n)
    conf = np.sort(rng.uniform(0.0, 1.0, n))[::-1]
    conf[: min(3, n)] = [1.0, 0.125, 0.995][: min(3, n)]          # exact ties of the :.2f rounding and the 1.00 case
    cls = rng.integers(0, len(names), n)
    det = np.stack([x1, y1, np.minimum(x1 + bw, w), np.minimum(y1 + bh, h), conf, cls], 1).astype(np.float32)
    t = torch.from_numpy(det).cuda()
    return Results(None, "x.jpg", names, t, (h, w), None, torch.from_numpy(det))


@pytest.mark.parametrize("thickness", [2, 1])
@pytest.mark.parametrize("h,w,n", [(720, 1280, 60), (480, 640, 300), (853, 1280, 7), (96, 352, 3)])
def test_draw_detections_is_bit_identical_to_cv2_when_labels_are_inside_the_image(h, w, n, thickness):
    """Boxes anywhere, labels fully inside the frame (top margin >= 23 px, right margin): every pixel equals the reference loop."""
    rng = np.random.default_rng(h * 7 + n)
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    res = random_results(rng, n, h, w, COCO_LIKE, margin=24)
    # keep labels inside: shift boxes whose label would cross the right border
    d = res.boxes.data.cpu().numpy().copy()
    d[:, 0] = np.minimum(d[:, 0], w - 200)
    d[:, 2] = np.maximum(d[:, 2], d[:, 0] + 2)
    res = Results(None, "x.jpg", COCO_LIKE, torch.from_numpy(d).cuda(), (h, w), None, torch.from_numpy(d))
    want = D.draw_detections(img, res, line_thickness=thickness)
    got = draw_detections(img, res, line_thickness=thickness)
    assert got.shape == want.shape and got.dtype == np.uint8
    bad = np.argwhere((got != want).any(-1))
    assert len(bad) == 0, f"{len(bad)} pixels differ, first at {bad[:5].tolist()}"
    assert (got != img).any()
    # device tensor in -> device tensor out, same pixels, input untouched
    dev_img = torch.from_numpy(img).cuda()
    out = draw_detections(dev_img, res, line_thickness=thickness)
    assert out.is_cuda and torch.equal(out.cpu(), torch.from_numpy(want)) and torch.equal(dev_img.cpu(), torch.from_numpy(img))


def test_draw_detections_clipped_labels_differ_only_inside_cut_glyphs():
    """Boxes at the frame border: outlines and label backgrounds stay exact; only glyphs cut by the border may differ (cv2 clips each
    stroke segment before rasterising it), and only inside the label's text cell rows/columns."""
    rng = np.random.default_rng(3)
    h, w, n = 360, 640, 80
    img = np.zeros((h, w, 3), np.uint8)
    res = random_results(rng, n, h, w, COCO_LIKE, margin=0)
    want = D.draw_detections(img, res)
    got = draw_detections(img, res)
    diff = (got != want).any(-1)
    frac = diff.mean()
    ys, xs = np.nonzero(diff)
    print(f"clipped labels: {diff.sum()} of {h * w} pixels differ ({frac:.2e})")
    assert frac < 2e-3
    assert len(ys) == 0 or ys.max() < 24 or xs.max() >= w - 24     # only next to the top / right border
    white_or_color = np.isin(got[diff].reshape(-1, 3), [0, 128, 165, 255]).all()
    assert white_or_color


def test_names_conf_rounding_and_painters_order():
    """`{name}: {conf:.2f}` formatted on the device equals Python's formatting for every 2-decimal tie a float32 can hit, and later
    detections are drawn over earlier ones."""
    names = {0: "a", 1: "Object"}
    img = np.zeros((2200, 200, 3), np.uint8)
    confs = np.array([0.005, 0.015, 0.025, 0.125, 0.375, 0.625, 0.875, 0.995, 1.0, 0.5, 0.25, 0.75, 0.999, 0.9949999, 1e-9, 0.3, 0.7, 0.045,
                      0.055, 0.105], np.float32)
    det = np.zeros((len(confs), 6), np.float32)
    for i, c in enumerate(confs):
        det[i] = [10, 30 + 100 * i, 150, 90 + 100 * i, c, i % 2]
    det = np.concatenate([det, det[:1] + np.array([5, 3, 5, 3, 0, 1], np.float32)])       # overlaps detection 0: drawn last, on top
    res = Results(None, "x.jpg", names, torch.from_numpy(det).cuda(), img.shape[:2], None, torch.from_numpy(det))
    want = D.draw_detections(img, res)
    got = draw_detections(img, res)
    assert np.array_equal(got, want)
    # class_names overrides results.names; unknown ids fall back to 'Object' (visualization.py:69-74)
    want = D.draw_detections(img, res, class_names={0: "zebra"})
    got = draw_detections(img, res, class_names={0: "zebra"})
    assert np.array_equal(got, want)


def test_batched_in_place_draw_and_unsupported_arguments():
    r = DetectionRasteriser(COCO_LIKE, "cuda:0")
    rng = np.random.default_rng(9)
    frames = torch.from_numpy(rng.integers(0, 256, (3, 240, 320, 3), dtype=np.uint8)).cuda()
    before = frames.cpu().numpy().copy()
    results = [random_results(rng, k, 240, 320, COCO_LIKE, margin=24) for k in (5, 0, 12)]
    dets = [x.boxes.data if len(x.boxes) else torch.zeros((0, 6), device="cuda") for x in results]
    r.draw_batch(frames, dets, [len(x.boxes) for x in results])
    torch.cuda.synchronize()
    for i, x in enumerate(results):
        d = x.boxes.data.cpu().numpy().copy()
        ok = d[:, 0] <= 320 - 200 if len(d) else np.zeros(0, bool)
        if len(d) and ok.all():
            assert np.array_equal(frames[i].cpu().numpy(), D.draw_detections(before[i], x))
    assert np.array_equal(frames[1].cpu().numpy(), before[1])
    with pytest.raises(ValueError):
        draw_detections(before[0], results[0], font_scale=1.0)
    with pytest.raises(Exception):
        draw_detections(before[0], results[0], line_thickness=5)


def test_serialisers_are_byte_identical(tmp_path):
    rng = np.random.default_rng(4)
    res = random_results(rng, 25, 480, 640, COCO_LIKE, margin=0)
    for fmt, ref in (("txt", D.save_results_txt), ("json", D.save_results_json), ("csv", D.save_results_csv)):
        a, b = tmp_path / f"a.{fmt}", tmp_path / f"b.{fmt}"
        ref(res, a)
        save_detection_results(res, str(b), fmt)
        assert a.read_bytes() == b.read_bytes(), fmt
    with pytest.raises(ValueError):
        save_detection_results(res, str(tmp_path / "x"), "xml")
