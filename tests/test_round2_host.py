"""CPU (not gpu): host logic added in round 2 against the oracle / cv2 - rect val batching and mAP, the glyph atlas the device
rasteriser blits, label formatting, serialisers, the result-exchange layout, the harness' system probes."""
import json
import struct

import cv2
import numpy as np
import pytest
import torch

from oracle import draw_ref as D
from oracle import val_ref as VR
from yolo_infer_b200 import val as V
from yolo_infer_b200.results import Results


def test_rect_batches_and_geometry_match_the_oracle_restatement():
    rng = np.random.default_rng(0)
    shapes = [(int(rng.integers(200, 1300)), int(rng.integers(200, 1300))) for _ in range(37)] + [(640, 640), (480, 640), (640, 480)]
    for batch in (1, 4, 16):
        order, canvases = V.rect_batches(shapes, batch)
        ir, bi, bs = VR.set_rectangle(shapes, batch)
        assert order == ir and canvases == [tuple(x) for x in bs]
        for k, i in enumerate(order):
            h0, w0 = shapes[i]
            H, W = canvases[k // batch]
            (new_w, new_h, top, left), row = V.rect_geometry(h0, w0, (H, W))
            li = VR.load_image(np.zeros((h0, w0, 3), np.uint8), 640)
            lb, (l2, t2) = VR.letterbox_val(li, (H, W))
            assert lb.shape[:2] == (H, W) and (new_w, new_h) == (li.shape[1], li.shape[0]) and (top, left) == (t2, l2)
            assert row[0] == pytest.approx(li.shape[0] / h0, rel=1e-12) and row[1:] == [float(left), float(top), float(w0), float(h0)]
    # known answer: COCO-style 480x640 images validate on a 512x672 canvas
    assert V.rect_batches([(480, 640)] * 3, 16)[1] == [(512, 672)]


def test_map_matches_the_oracle_restatement():
    rng = np.random.default_rng(1)
    preds, gts = [], []
    for _ in range(20):
        m = int(rng.integers(0, 6))
        gt = np.zeros((m, 5))
        gt[:, 0] = rng.integers(0, 4, m)
        xy, wh = rng.uniform(0, 400, (m, 2)), rng.uniform(20, 120, (m, 2))
        gt[:, 1:3], gt[:, 3:5] = xy, xy + wh
        pr = [[*(g[1:] + rng.normal(0, 5, 4)), rng.random(), g[0]] for g in gt if rng.random() < 0.8]
        for _ in range(int(rng.integers(0, 4))):
            xy = rng.uniform(0, 400, 2)
            pr.append([*xy, *(xy + rng.uniform(20, 100, 2)), rng.random(), rng.integers(0, 4)])
        preds.append(np.array(pr).reshape(-1, 6))
        gts.append(gt)
    m = V.evaluate(preds, gts, 4)
    m50, m5095 = VR.mean_ap(preds, gts)
    assert m.map50 == pytest.approx(m50, abs=1e-12) and m.map == pytest.approx(m5095, abs=1e-12)


# ---- the rasteriser's host-side model: atlas + pen arithmetic + label formatting --------------------------------------------------
ATLAS = np.load(V.Path(__file__).resolve().parents[1] / "yolo_infer_b200" / "glyphs_simplex_0p5.npz")


def compose(label, org, H, W):
    bits, adv = ATLAS["bits"], ATLAS["advance_half_px"]
    ch_, cw_, by_, px_ = [int(v) for v in ATLAS["cell"]]
    img = np.zeros((H, W), bool)
    pen = 0
    for ch in label:
        i = ord(ch) - 32
        X = org[0] + pen // 2 - px_
        for gy in range(ch_):
            y = org[1] - by_ + gy
            row = int(bits[i, pen % 2, gy])
            for gx in range(cw_):
                if (row >> gx) & 1 and 0 <= y < H and 0 <= X + gx < W:
                    img[y, X + gx] = True
        pen += int(adv[i])
    return img, pen


def test_glyph_atlas_composition_equals_cv2_puttext_and_gettextsize():
    rng = np.random.default_rng(2)
    alphabet = [chr(c) for c in range(32, 127)]
    for _ in range(150):
        label = "".join(rng.choice(alphabet, int(rng.integers(1, 24))))
        org = (int(rng.integers(5, 40)), int(rng.integers(18, 50)))
        ref = np.zeros((64, 420), np.uint8)
        cv2.putText(ref, label, org, cv2.FONT_HERSHEY_SIMPLEX, 0.5, 255, 1)
        got, pen = compose(label, org, 64, 420)
        assert np.array_equal(got, ref > 0), label
        (w, h), _ = cv2.getTextSize(label, cv2.FONT_HERSHEY_SIMPLEX, 0.5, 1)
        i = (pen >> 1) + 1
        assert w == i + ((i & 1) if pen & 1 else 0) and h == int(ATLAS["text_height"])     # csrc/draw.cu: bg_w


def conf_times_100(v):       # csrc/draw.cu: conf_times_100, in Python integers
    u = struct.unpack("<I", struct.pack("<f", float(v)))[0]
    e = (u >> 23) & 0xFF
    m = ((u & 0x7FFFFF) | (0x800000 if e else 0)) * 100
    sh = 150 - (e if e else 1)
    if sh <= 0:
        return m << (-sh)
    if sh >= 64:
        return 0
    q, rem, half = m >> sh, m & ((1 << sh) - 1), 1 << (sh - 1)
    return q + (1 if (rem > half or (rem == half and q & 1)) else 0)


def test_device_label_formatting_rule_equals_python_format():
    rng = np.random.default_rng(3)
    vals = np.concatenate([rng.random(20000).astype(np.float32), np.arange(0, 1.0001, 0.005, dtype=np.float32),
                           np.float32([1.0, 0.0, 1e-9, 0.125, 0.375, 0.995, 0.9949999, 0.005])])
    for v in vals:
        n = conf_times_100(v)
        assert f"{n // 100}.{n // 10 % 10}{n % 10}" == f"{float(v):.2f}", float(v)


def test_serialisers_byte_identical_to_the_reference_writers(tmp_path):
    from yolo_infer_b200.draw import save_detection_results
    rng = np.random.default_rng(4)
    det = np.concatenate([rng.uniform(0, 640, (30, 4)), rng.random((30, 1)), rng.integers(0, 80, (30, 1))], 1).astype(np.float32)
    res = Results(None, "x.jpg", {i: str(i) for i in range(80)}, torch.from_numpy(det), (480, 640))
    for fmt, ref in (("txt", D.save_results_txt), ("json", D.save_results_json), ("csv", D.save_results_csv)):
        a, b = tmp_path / f"a.{fmt}", tmp_path / f"b.{fmt}"
        ref(res, a)
        save_detection_results(res, str(b), fmt)
        assert a.read_bytes() == b.read_bytes(), fmt
    empty = Results(None, "x.jpg", {}, torch.zeros((0, 6)), (480, 640))
    save_detection_results(empty, str(tmp_path / "e.json"), "json")
    assert json.loads((tmp_path / "e.json").read_text()) == {"detections": []}


def test_result_exchange_layout_single_rank():
    """world 1: the exchange degenerates to a local buffer; slots, signal and free words must not overlap."""
    from yolo_infer_b200.parallel import ResultExchange, split_flat
    B, MD = 3, 5
    n_flat = B * MD * 6 + B
    x = ResultExchange(None, None, n_flat, 2, torch.device("cpu"))
    assert x.mode == "local" and x.world == 1 and x.push_ptrs(0) is None
    a, b = x.out_flat(0), x.out_flat(1)
    assert a.numel() == b.numel() == n_flat and a.data_ptr() + 4 * n_flat == b.data_ptr()
    assert x.o_sig == 2 * n_flat and x.o_free == x.o_sig + 2 and x.o_done == x.o_free + 2 and x.buf.numel() >= x.o_done + 2
    a[: B * MD * 6] = torch.arange(B * MD * 6, dtype=torch.float32)
    a[B * MD * 6:] = torch.tensor([1, 2, 3], dtype=torch.int32).view(torch.float32)
    det, cnt = split_flat(x.slot_results(0), B, MD)
    assert det.shape == (B, MD, 6) and cnt.tolist() == [1, 2, 3]


def test_harness_system_probes_keep_the_reference_keys():
    from yolo_infer_b200.harness import ResourceMonitor, SpeedBenchmark, get_device_info
    info = get_device_info()
    for k in ("platform", "processor", "architecture", "python_version", "cpu_count", "memory_total_gb", "memory_available_gb",
              "torch_version", "cuda_available", "mps_available"):     # utils/helpers.py:27-38
        assert k in info
    m = ResourceMonitor(interval=0.05)
    m.start_monitoring()
    import time
    time.sleep(0.5)
    m.stop_monitoring()
    avg = m.get_average_usage()
    assert {"avg_cpu_percent", "avg_memory_percent"} <= set(avg) and m.get_current_usage()["memory_total"] > 0
    s = SpeedBenchmark._calculate_summary([{"fps": 10.0, "throughput": 40.0}, {"fps": 20.0, "throughput": 20.0}])
    assert s == {"best_fps": 20.0, "worst_fps": 10.0, "avg_fps": 15.0, "best_throughput": 40.0, "worst_throughput": 20.0,
                 "avg_throughput": 30.0, "total_configurations": 2}


def test_results_batch_is_a_lazy_sequence_of_results():
    from yolo_infer_b200.results import ResultsBatch
    det = torch.arange(3 * 4 * 6, dtype=torch.float32).view(3, 4, 6)
    rb = ResultsBatch(None, det, [2, 0, 4], {0: "a"}, (480, 640), paths=["p0", "p1", "p2"])
    assert len(rb) == 3 and not rb._items                        # nothing built yet
    assert len(rb[0].boxes) == 2 and rb[0].path == "p0" and rb[0].orig_shape == (480, 640)
    assert not rb[1].boxes and len(rb[-1].boxes) == 4 and rb[0] is rb[0]
    assert [len(r.boxes) for r in rb] == [2, 0, 4] and [len(r.boxes) for r in rb[1:]] == [0, 4]
    assert torch.equal(rb[2].boxes.xyxy, det[2, :, :4]) and torch.equal(rb[2].cpu().boxes.conf, det[2, :, 4])
    with pytest.raises(IndexError):
        rb[3]


def test_compact_space_to_depth_packing_equals_the_3x3_stride2_conv():
    """model.1 on the permuted space-to-depth stem output (network.S2D_PERM, y11_conv_desc.s2d_block): the per-stage packed weights,
    expanded back to a dense 2x2 kernel (network.dense_k2_weights), applied to the permuted space-to-depth tensor with top/left zero
    padding, must equal the original 3x3 stride-2 conv - pure torch on the CPU, for both supported block sizes."""
    import torch
    from yolo_infer_b200 import network as N
    from yolo_infer_b200 import topology as T
    for scale, c in (("s", 32), ("m", 64)):
        sd = T.synthetic_state_dict(scale, seed=3)
        packed = N.pack_weights(scale, 80, sd, torch.device("cpu"))
        pc = packed["model.1"]
        assert pc.k == 2 and pc.s2d_block == c and pc.c1 == 4 * c
        stages = N.compact_k2_stages(c)
        assert len(stages) == (5 if c == 32 else 9) and pc.w.shape == (pc.c2, 64 * len(stages))
        cp = next(p for p in T.conv_params(scale, 80) if p.prefix == "model.1")
        w3, b3 = N.fold(sd, cp)
        w3 = w3.to(torch.bfloat16).float()                                  # the packing rounds the folded weights to bf16
        g = torch.Generator().manual_seed(4)
        x = torch.randn(2, c, 24, 32, generator=g)                          # a stem output [B, c, H, W]
        want = torch.nn.functional.conv2d(x, w3, b3, stride=2, padding=1)
        # permuted space-to-depth form: [B, 4c, H/2, W/2], block i holds pixels (2y+dy, 2x+dx) with (dy, dx) = S2D_PERM[i]
        xs = torch.cat([x[:, :, dy::2, dx::2] for dy, dx in N.S2D_PERM], 1)
        w2 = N.dense_k2_weights(pc).permute(0, 3, 1, 2)                       # [cout, 4c, 2, 2]
        got = torch.nn.functional.conv2d(torch.nn.functional.pad(xs, (1, 0, 1, 0)), w2, pc.b)
        assert got.shape == want.shape
        assert float((got - want).abs().max()) <= 1e-4 * float(want.abs().max())
        # only the blocks a tap can touch carry weights: 9 of the 16 (tap, block) pairs
        nz = (w2.view(pc.c2, 4, c, 2, 2).abs().sum((0, 2)) > 0).sum().item()
        assert nz == 9


def test_uint8_to_unit_float_through_the_mantissa_trick_is_the_rounded_quotient():
    """The stem's uint8 path (csrc/conv_simt.cu, stem_kernel<., true>) converts a byte b to b/255 as one fused multiply-add on the
    float 2^23 + b (PRMT puts the byte into the mantissa of 0x4B000000): fma(2^23 + b, r, -(2^23 * r)) with r = fp32(1/255).  The
    product of two 24-bit significands is exact in float64, so the fused result is RN_fp32(b * r) - the value letterbox_kernel's
    __fmul_rn((float)b, r) produces - and its bf16 rounding equals the bf16 rounding of the exact fp32 quotient b / 255."""
    import numpy as np
    r = np.float32(1.0) / np.float32(255.0)
    c = np.float32(-8388608.0) * r                                   # exact: a power-of-two multiple of r
    assert float(c) == -8388608.0 * float(r)
    b = np.arange(256, dtype=np.uint32)
    v = (np.uint32(0x4B000000) | b).view(np.float32)                 # what PRMT builds: 2^23 + b
    assert np.array_equal(v, (8388608.0 + b).astype(np.float32))
    fused = (v.astype(np.float64) * np.float64(r) + np.float64(c)).astype(np.float32)   # one rounding, like fma.rn
    plain = b.astype(np.float32) * r                                 # __fmul_rn((float)b, r)
    assert np.array_equal(fused.view(np.uint32), plain.view(np.uint32))

    def bf16_bits(x):                                                # round-to-nearest-even truncation of fp32 to bf16
        u = x.view(np.uint32).astype(np.uint64)
        return ((u + 0x7FFF + ((u >> 16) & 1)) >> 16).astype(np.uint32)
    quotient = b.astype(np.float32) / np.float32(255.0)
    assert np.array_equal(bf16_bits(plain), bf16_bits(quotient))
