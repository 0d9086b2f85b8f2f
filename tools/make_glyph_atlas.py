#!/usr/bin/env python
"""Generates yolo_infer_b200/glyphs_simplex_0p5.npz: the FONT_HERSHEY_SIMPLEX glyphs at fontScale 0.5, thickness 1, as cv2.putText
rasterises them, for both sub-pixel phases of the pen position (advances are multiples of half a pixel at this scale).

The reference draws its labels with cv2.putText(..., FONT_HERSHEY_SIMPLEX, 0.5, (255,255,255), 1)
(/root/reference/utils/visualization.py:95-104).  The B200 rasteriser (csrc/draw.cu) blits these bitmaps with the same pen
arithmetic (OpenCV: view_x advances by (right - left) * round(fontScale * 65536) in 16.16 fixed point), which makes its
output bit-identical to cv2's - tests/test_draw.py composes random labels from the atlas and compares with cv2.putText.

Glyph (c, phase) is recovered by rendering `prefix + c` and `prefix` alone and keeping the pixels the character added; the
prefix sets the pen phase (a character of odd advance -> phase 1).  Two different prefixes per phase must give the same bitmap
(a stroke overlapping the prefix's pixels would be lost otherwise); the script asserts that.
"""
import sys
from pathlib import Path

import cv2
import numpy as np

FONT, SCALE, THICK = cv2.FONT_HERSHEY_SIMPLEX, 0.5, 1
GH, GW, BASE_Y, PAD_X = 24, 24, 17, 4      # glyph cell; baseline row inside the cell; columns left of the pen position


def advance_half_px(ch: str) -> int:
    (w, _), _ = cv2.getTextSize(ch, FONT, 1.0, 1)      # width = (right - left) + thickness at scale 1
    return w - 1                                        # in half pixels at scale 0.5


def render(s: str, org_x: int, width: int) -> np.ndarray:
    img = np.zeros((GH, width), np.uint8)
    cv2.putText(img, s, (org_x, BASE_Y), FONT, SCALE, 255, THICK)
    return img > 0


def glyph(ch: str, prefix: str) -> np.ndarray:
    adv = sum(advance_half_px(c) for c in prefix)
    org = 40
    width = org + adv // 2 + 60
    both, alone = render(prefix + ch, org, width), (render(prefix, org, width) if prefix else np.zeros((GH, width), bool))
    x0 = org + adv // 2 - PAD_X                          # cell origin = integer pen column - PAD_X
    cell = (both & ~alone)[:, x0:x0 + GW]
    assert not (both & ~alone)[:, :x0].any() and not (both & ~alone)[:, x0 + GW:].any(), f"glyph {ch!r} leaves its cell"
    return cell


def main():
    chars = [chr(c) for c in range(32, 127)]
    adv = np.array([advance_half_px(c) for c in chars], np.int32)
    odd = [c for c, a in zip(chars, adv) if a % 2]
    atlas = np.zeros((len(chars), 2, GH, GW), bool)
    for i, ch in enumerate(chars):
        for phase, prefixes in ((0, ["", "  ", "00"]), (1, odd[:1] + odd[5:6] + [odd[10] + "  "])):
            cells = [glyph(ch, p) for p in prefixes]
            for c in cells[1:]:
                assert np.array_equal(cells[0], c), f"glyph {ch!r} phase {phase}: prefixes disagree"
            atlas[i, phase] = cells[0]
    # pack rows into uint32 bit masks (bit x of row y)
    bits = np.zeros((len(chars), 2, GH), np.uint32)
    for x in range(GW):
        bits |= atlas[..., x].astype(np.uint32) << np.uint32(x)
    (tw, th), base = cv2.getTextSize("Ag", FONT, SCALE, THICK)
    out = Path(__file__).resolve().parents[1] / "yolo_infer_b200" / "glyphs_simplex_0p5.npz"
    np.savez_compressed(out, bits=bits, advance_half_px=adv, first_char=np.int32(32), cell=np.array([GH, GW, BASE_Y, PAD_X], np.int32),
                        text_height=np.int32(th), baseline=np.int32(base))
    print(out, bits.shape, "text height", th, "baseline", base)


if __name__ == "__main__":
    main()
