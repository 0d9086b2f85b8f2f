"""ORACLE (test infrastructure, never shipped, never on the product path).

Integer restatement (numpy) of OpenCV's uint8 ``cv2.resize(INTER_LINEAR)`` + constant border, i.e. the
arithmetic behind the reference's letterbox (reached from
/root/reference/core/model.py:133 -> ultralytics LetterBox; the in-tree statement of intent is
/root/reference/utils/data_loader.py:281-337).  cv2 itself is importable, so this restatement is
PINNED: tests/test_oracle_kat.py checks it bit-for-bit against ``cv2.resize`` on many shapes.  It
exists so that the integer algorithm the CUDA letterbox kernel implements is written down once in
plain numpy:

  * horizontal taps: fx = float32((dx+.5)*scale-.5); sx=floor(fx); clamp (sx<0 -> 0,fx=0;
    sx>=w-1 -> w-1,fx=0); a0 = rint((1-fx)*2048), a1 = rint(fx*2048)  (int16 coefficients)
  * vertical taps: same fy but NOT clamped; the two source rows are clamped to [0,h-1] instead
  * value = ((b0*(S0>>4))>>16) + ((b1*(S1>>4))>>16) + 2) >> 2, S = x0*a0 + x1*a1
  * exact 2x downscale in both axes takes OpenCV's INTER_AREA fast path: (a+b+c+d+2)>>2
"""
from __future__ import annotations

import numpy as np


def _taps(src: int, dst: int, vertical: bool):
    scale = src / dst
    d = np.arange(dst)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int32)
    f = (f - s).astype(np.float32)
    if not vertical:
        lo = s < 0
        f[lo] = 0
        s[lo] = 0
        hi = s >= src - 1
        f[hi] = 0
        s[hi] = src - 1
    c0 = np.rint((np.float32(1.0) - f) * np.float32(2048)).astype(np.int32)
    c1 = np.rint(f * np.float32(2048)).astype(np.int32)
    return np.clip(s, 0, src - 1), np.clip(s + 1, 0, src - 1), c0, c1


def resize_linear_u8(img: np.ndarray, dw: int, dh: int) -> np.ndarray:
    h, w = img.shape[:2]
    if (w, h) == (dw, dh):
        return img.copy()
    if w == 2 * dw and h == 2 * dh:
        a = img.astype(np.int32)
        return ((a[0::2, 0::2] + a[0::2, 1::2] + a[1::2, 0::2] + a[1::2, 1::2] + 2) >> 2).astype(np.uint8)
    x0, x1, a0, a1 = _taps(w, dw, False)
    y0, y1, b0, b1 = _taps(h, dh, True)
    s = img.astype(np.int32)
    rows = s[:, x0] * a0[None, :, None] + s[:, x1] * a1[None, :, None]
    out = (((b0[:, None, None] * (rows[y0] >> 4)) >> 16) + ((b1[:, None, None] * (rows[y1] >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def letterbox_u8(img: np.ndarray, new_w: int, new_h: int, top: int, left: int, H: int, W: int) -> np.ndarray:
    """Resize to (new_w,new_h), place at (top,left) on an HxW canvas of 114."""
    out = np.full((H, W, 3), 114, np.uint8)
    out[top:top + new_h, left:left + new_w] = resize_linear_u8(img, new_w, new_h)
    return out


def normalize_rgb(img_bgr_u8: np.ndarray) -> np.ndarray:
    """BGR->RGB, /255 in fp32 (HWC kept): what BasePredictor.preprocess feeds the network."""
    return img_bgr_u8[..., ::-1].astype(np.float32) / np.float32(255.0)
