"""GPU JPEG decode for file sources: `cv2.imread` of the reference's loaders
(/root/reference/utils/data_loader.py:42; ultralytics LoadImagesAndVideos) replaced by nvJPEG writing the BGR frame straight
into device memory, where the letterbox kernel (or the stem) reads it - the compressed file crosses PCIe, not the 3.3 MB frame.

Opt-in: `YOLO.predict(path, decode="nvjpeg")` / `GpuJpegDecoder(device).decode(path_or_bytes)`.  The default stays cv2.imread,
which is what the reference runs and therefore the bit-exact parity path; nvJPEG frames differ from it by the decoder's IDCT /
chroma-upsampling arithmetic (tolerance measured in tests/test_gpu_decode.py).  Video (NVDEC) is not built: the image has no
demuxer / libnvcuvid headers, so the video loop keeps cv2.VideoCapture on the host and uploads frames.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path
from typing import Union

import numpy as np
import torch

from . import _cabi as cabi


class GpuJpegDecoder:
    def __init__(self, device: Union[str, torch.device] = "cuda", engine_handle=None):
        self.lib = cabi.load()
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("GpuJpegDecoder needs a CUDA device (the CPU decode of the reference is cv2.imread)")
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        self.device = dev
        self._own = engine_handle is None
        self.h = engine_handle if engine_handle is not None else C.c_void_p()
        self.j = C.c_void_p()
        with torch.cuda.device(dev):
            if self._own:
                cabi.check(self.lib.y11_create(C.byref(self.h), dev.index), "y11_create")
            cabi.check(self.lib.y11_jpeg_create(self.h, C.byref(self.j)), "y11_jpeg_create")

    def __del__(self):
        try:
            if self.j:
                self.lib.y11_jpeg_destroy(self.j)
                self.j = C.c_void_p()
            if self._own and self.h:
                self.lib.y11_destroy(self.h)
                self.h = C.c_void_p()
        except Exception:
            pass

    def decode(self, src: Union[str, Path, bytes, np.ndarray]) -> torch.Tensor:
        """-> device uint8 [H,W,3] BGR (the layout cv2.imread returns).  Asynchronous on the current stream."""
        if isinstance(src, (str, Path)):
            data = Path(src).read_bytes()
        elif isinstance(src, np.ndarray):
            data = src.tobytes()
        else:
            data = bytes(src)
        buf = (C.c_uint8 * len(data)).from_buffer_copy(data)
        h, w, nc = C.c_int32(), C.c_int32(), C.c_int32()
        cabi.check(self.lib.y11_jpeg_info(self.j, buf, len(data), C.byref(h), C.byref(w), C.byref(nc)), "y11_jpeg_info")
        with torch.cuda.device(self.device):
            out = torch.empty((h.value, w.value, 3), dtype=torch.uint8, device=self.device)
            s = torch.cuda.current_stream(self.device).cuda_stream
            cabi.check(self.lib.y11_jpeg_decode(self.j, buf, len(data), out.data_ptr(), out.stride(0), h.value, w.value, C.c_void_p(s)),
                       "y11_jpeg_decode")
        return out


def is_jpeg_path(p) -> bool:
    return isinstance(p, (str, Path)) and str(p).lower().endswith((".jpg", ".jpeg"))
