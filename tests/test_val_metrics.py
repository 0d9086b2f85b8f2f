"""Host-side validation maths (yolo_infer_b200/val.py): ultralytics match_predictions / ap_per_class restated."""
import numpy as np
import pytest

from yolo_infer_b200 import val as V


def _boxes(n, rng, size=640):
    xy = rng.uniform(0, size - 120, (n, 2))
    wh = rng.uniform(20, 100, (n, 2))
    return np.concatenate((xy, xy + wh), 1)


def test_box_iou_and_matching_are_one_to_one():
    gt = np.array([[0, 0, 10, 10], [20, 20, 30, 30]], float)
    pr = np.array([[0, 0, 10, 10], [1, 1, 11, 11], [20, 20, 30, 31]], float)
    iou = V.box_iou(gt, pr)
    assert iou.shape == (2, 3) and abs(iou[0, 0] - 1) < 1e-6 and iou[0, 2] == 0
    tp = V.match_predictions(np.zeros(3), np.zeros(2), iou)
    assert tp[0].all()                       # exact box: true positive at every threshold
    assert not tp[1].any()                   # second prediction on the same ground truth: the better one keeps it
    assert tp[2, :9].all() and not tp[2, 9]  # IoU 10/11 = 0.909: passes 0.5 .. 0.90, fails 0.95
    tp = V.match_predictions(np.array([1.0, 0, 0]), np.zeros(2), iou)
    assert not tp[0].any() and tp[1].any()   # wrong class never matches; the runner-up now gets the ground truth


def test_perfect_and_empty_predictions():
    rng = np.random.default_rng(0)
    gts, preds = [], []
    for _ in range(8):
        b = _boxes(6, rng)
        c = rng.integers(0, 3, 6).astype(float)
        gts.append(np.concatenate((c[:, None], b), 1))
        preds.append(np.concatenate((b, rng.uniform(0.3, 0.9, (6, 1)), c[:, None]), 1))
    m = V.evaluate(preds, gts, nc=3)
    assert m.map > 0.99 and m.map50 > 0.99 and m.map75 > 0.99 and m.mp > 0.99 and m.mr > 0.99   # 101-point interp: 0.995
    assert list(m.ap_class_index) == [0, 1, 2] and m.maps.shape == (3,)
    empty = V.evaluate([np.zeros((0, 6))] * 8, gts, nc=3)
    assert empty.map == 0 and empty.mp == 0 and empty.mr == 0


def test_iou_thresholds_and_false_positives():
    gt = [np.array([[0, 0, 0, 100, 100]], float)]
    # IoU = 82/100 = 0.82: true positive at 0.50..0.80 (7 thresholds), false positive above
    pr = [np.array([[0, 0, 100, 82, 0.9, 0]], float)]
    m = V.evaluate(pr, gt, nc=1)
    assert abs(m.map50 - 0.995) < 1e-3 and abs(m.map75 - 0.995) < 1e-3
    assert abs(m.map - 0.7 * 0.995) < 1e-3
    # a higher-confidence false positive in front of the true positive halves the precision: AP50 = 0.5 (envelope)
    pr = [np.array([[300, 300, 400, 400, 0.95, 0], [0, 0, 100, 100, 0.9, 0]], float)]
    m = V.evaluate(pr, gt, nc=1)
    assert abs(m.map50 - 0.5) < 1e-2
    # ... behind it, it costs nothing
    pr = [np.array([[0, 0, 100, 100, 0.9, 0], [300, 300, 400, 400, 0.5, 0]], float)]
    m = V.evaluate(pr, gt, nc=1)
    assert abs(m.map50 - 0.995) < 1e-3


def test_dataset_loading_and_label_conversion(tmp_path):
    cv2 = pytest.importorskip("cv2")
    (tmp_path / "images" / "val").mkdir(parents=True)
    (tmp_path / "labels" / "val").mkdir(parents=True)
    img = np.zeros((100, 200, 3), np.uint8)
    cv2.imwrite(str(tmp_path / "images" / "val" / "a.jpg"), img)
    (tmp_path / "labels" / "val" / "a.txt").write_text("2 0.5 0.5 0.25 0.5\n")
    (tmp_path / "d.yaml").write_text(f"path: {tmp_path}\nval: images/val\nnames:\n  0: a\n  1: b\n  2: c\n")
    files, names = V.load_dataset(tmp_path / "d.yaml")
    assert [f.name for f in files] == ["a.jpg"] and names == {0: "a", 1: "b", 2: "c"}
    lab = V.read_labels(files[0], 200, 100)
    assert np.allclose(lab, [[2, 75, 25, 125, 75]])
    assert V.read_labels(tmp_path / "images" / "val" / "missing.jpg", 10, 10).shape == (0, 5)
