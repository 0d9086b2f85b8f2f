"""Generates the small committed fixtures under tests/golden/ (run once, in the build container):
    python tests/golden/make_golden.py
Sources of truth: cv2 (the reference's letterbox arithmetic), torchvision.ops.nms (the reference's NMS) and the
oracle's Detect.decode restatement.  Nothing here reads /root/reference."""
import sys
from pathlib import Path

import numpy as np
import torch
import torchvision

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import pipeline_ref as P  # noqa: E402
from oracle import yolo11_ref as R  # noqa: E402

OUT = Path(__file__).parent
rng = np.random.default_rng(0)

img = rng.integers(0, 256, (45, 80, 3), dtype=np.uint8)
np.savez_compressed(OUT / "letterbox_small.npz", img=img, out_rect=P.letterbox(img, (64, 64), auto=True),
                    out_square=P.letterbox(img, (64, 64), auto=False))

g = torch.Generator().manual_seed(0)
K = 300
centers = torch.rand(40, 2, generator=g) * 500 + 50
idx = torch.randint(0, 40, (K,), generator=g)
xy = centers[idx] + (torch.rand(K, 2, generator=g) - 0.5) * 4        # clustered: jitter +-2 px
wh = torch.rand(K, 2, generator=g) * 60 + 20
boxes = torch.cat((xy - wh / 2, xy + wh / 2), 1)
scores = (torch.rand(K, generator=g) * 64).round() / 64               # many exact ties
cls = torch.randint(0, 3, (K,), generator=g).float()
iou = 0.45
keep = torchvision.ops.nms(boxes + cls[:, None] * 7680, scores, iou)
np.savez_compressed(OUT / "nms_small.npz", boxes=boxes.numpy(), scores=scores.numpy(), cls=cls.numpy(), iou=np.float64(iou),
                    keep=keep.numpy())

det = R.Detect(nc=80, ch=(16, 16, 16))
feats = [torch.randn(2, 144, h, w, generator=g) * 2 for (h, w) in [(8, 12), (4, 6), (2, 3)]]
y = det.decode(feats)
np.savez_compressed(OUT / "decode_small.npz", f0=feats[0].numpy(), f1=feats[1].numpy(), f2=feats[2].numpy(), y=y.numpy())
ref_img = Path("/root/reference/image.jpg")
if ref_img.exists():  # only in the build container; the committed small copy is what travels
    import cv2
    im = cv2.imread(str(ref_img))
    small = cv2.resize(im, (480, 320), interpolation=cv2.INTER_AREA)
    cv2.imwrite(str(OUT / "image_small.jpg"), small, [cv2.IMWRITE_JPEG_QUALITY, 90])
print("golden fixtures written to", OUT)
