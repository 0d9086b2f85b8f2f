"""ctypes binding of include/y11.h (liby11_b200.so).  No compute happens in Python; there is NO CPU fallback:
if the library is missing or no sm_100 GPU is present, every entry point raises."""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "_lib" / "liby11_b200.so"

ABI_VERSION = 6
ACT_NONE, ACT_SILU = 0, 1
IMPL_TCGEN05, IMPL_SIMT_DEBUG = 0, 1
RES_POST, RES_PRE_UP2 = 0, 1


class Y11Error(RuntimeError):
    pass


class Image(C.Structure):
    _fields_ = [("src", C.c_void_p), ("h0", C.c_int32), ("w0", C.c_int32), ("pitch", C.c_int32),
                ("new_h", C.c_int32), ("new_w", C.c_int32), ("top", C.c_int32), ("left", C.c_int32)]


class View(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("c_total", C.c_int32), ("c_off", C.c_int32), ("c", C.c_int32)]


class ConvDesc(C.Structure):
    _fields_ = [("inp", View), ("out", View), ("res", View), ("w", C.c_void_p), ("bias", C.c_void_p),
                ("B", C.c_int32), ("Hin", C.c_int32), ("Win", C.c_int32), ("Hout", C.c_int32), ("Wout", C.c_int32),
                ("k", C.c_int32), ("stride", C.c_int32), ("act", C.c_int32), ("out_f32", C.c_int32), ("impl", C.c_int32), ("res_mode", C.c_int32),
                ("in_fp8", C.c_int32), ("out_fp8", C.c_int32), ("cscale", C.c_void_p), ("out_scale", C.c_float), ("s2d_block", C.c_int32)]


class StemDesc(C.Structure):
    _fields_ = [("inp", C.c_void_p), ("out", View), ("w", C.c_void_p), ("bias", C.c_void_p),
                ("B", C.c_int32), ("Hin", C.c_int32), ("Win", C.c_int32), ("Hout", C.c_int32), ("Wout", C.c_int32),
                ("s2d", C.c_int32), ("images", C.c_void_p), ("u8_src", C.c_int32)]


class DwConvDesc(C.Structure):
    _fields_ = [("inp", View), ("out", View), ("res", View), ("w", C.c_void_p), ("bias", C.c_void_p),
                ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("act", C.c_int32)]


class SppfDesc(C.Structure):
    _fields_ = [("io", View), ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("c", C.c_int32)]


class UpsampleDesc(C.Structure):
    _fields_ = [("inp", View), ("out", View), ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32)]


class AttnDesc(C.Structure):
    _fields_ = [("qkv", View), ("out", View), ("B", C.c_int32), ("N", C.c_int32), ("heads", C.c_int32),
                ("kd", C.c_int32), ("hd", C.c_int32), ("scale", C.c_float)]


class HeadDesc(C.Structure):
    _fields_ = [("head", C.c_void_p * 3), ("hl", C.c_int32 * 3), ("wl", C.c_int32 * 3), ("stride", C.c_float * 3),
                ("nl", C.c_int32), ("B", C.c_int32), ("nc", C.c_int32), ("row_stride", C.c_int32)]


class NmsParams(C.Structure):
    _fields_ = [("conf", C.c_float), ("iou", C.c_double), ("max_det", C.c_int32), ("max_nms", C.c_int32),
                ("max_wh", C.c_int32), ("agnostic", C.c_int32), ("multi_label", C.c_int32)]


class DrawItem(C.Structure):
    _fields_ = [("img", C.c_void_p), ("h", C.c_int32), ("w", C.c_int32), ("pitch", C.c_int32), ("det", C.c_void_p),
                ("count", C.c_void_p), ("n", C.c_int32), ("max_det", C.c_int32)]


class Font(C.Structure):
    _fields_ = [("glyph_bits", C.c_void_p), ("advance", C.c_void_p), ("first_char", C.c_int32), ("n_chars", C.c_int32),
                ("cell_h", C.c_int32), ("cell_w", C.c_int32), ("base_y", C.c_int32), ("pad_x", C.c_int32), ("text_h", C.c_int32),
                ("names", C.c_void_p), ("nc", C.c_int32), ("name_stride", C.c_int32)]


class Push(C.Structure):
    _fields_ = [("done_counter", C.c_void_p), ("signal", C.c_void_p)]


class ClsEmit(C.Structure):
    _fields_ = [("list", C.c_void_p), ("count", C.c_void_p), ("cap", C.c_int32), ("nc", C.c_int32),
                ("anchor_offset", C.c_int32), ("logit_threshold", C.c_float)]


# name -> (restype, argtypes); mirrors include/y11.h one to one (tests/test_cabi_symbols.py checks the header)
_P = C.c_void_p
SIGNATURES = {
    "y11_abi_version": (C.c_int, []),
    "y11_last_error": (C.c_char_p, []),
    "y11_create": (C.c_int, [C.POINTER(_P), C.c_int]),
    "y11_destroy": (None, [_P]),
    "y11_engine_error_code": (C.c_int, [_P]),
    "y11_letterbox": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, _P, _P]),
    "y11_letterbox_u8": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, _P, _P]),
    "y11_nchw_f32_to_nhwc_bf16": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_float, _P, _P]),
    "y11_nchw_f32_to_nhwc_bf16_auto": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, _P, _P, _P]),
    "y11_plan_create": (C.c_int, [_P, C.POINTER(_P)]),
    "y11_plan_destroy": (None, [_P]),
    "y11_plan_add_conv": (C.c_int, [_P, C.POINTER(ConvDesc)]),
    "y11_plan_add_conv_tuned": (C.c_int, [_P, C.POINTER(ConvDesc), C.c_int, C.c_int, C.c_int, C.c_int]),
    "y11_plan_autotune": (C.c_int, [_P, _P, C.c_int]),
    "y11_plan_op_variant": (C.c_int, [_P, C.c_int, C.POINTER(C.c_int32)]),
    "y11_plan_add_stem": (C.c_int, [_P, C.POINTER(StemDesc)]),
    "y11_plan_set_stem_source": (C.c_int, [_P, _P]),
    "y11_plan_add_dwconv": (C.c_int, [_P, C.POINTER(DwConvDesc)]),
    "y11_plan_add_sppf": (C.c_int, [_P, C.POINTER(SppfDesc)]),
    "y11_plan_add_upsample": (C.c_int, [_P, C.POINTER(UpsampleDesc)]),
    "y11_plan_add_attention": (C.c_int, [_P, C.POINTER(AttnDesc)]),
    "y11_plan_fork": (C.c_int, [_P, C.c_int]),
    "y11_plan_set_lane": (C.c_int, [_P, C.c_int]),
    "y11_plan_join": (C.c_int, [_P, C.c_int]),
    "y11_plan_num_ops": (C.c_int, [_P]),
    "y11_plan_num_launches": (C.c_int, [_P]),
    "y11_plan_run": (C.c_int, [_P, _P]),
    "y11_plan_run_ops": (C.c_int, [_P, C.c_int, C.c_int, _P]),
    "y11_plan_run_range": (C.c_int, [_P, C.c_int, C.c_int, _P]),
    "y11_plan_run_timed": (C.c_int, [_P, _P, C.POINTER(C.c_float)]),
    "y11_plan_op_flops": (C.c_double, [_P, C.c_int]),
    "y11_decode_dense": (C.c_int, [_P, C.POINTER(HeadDesc), _P, _P]),
    "y11_postprocess_workspace": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "y11_detect_postprocess": (C.c_int, [_P, C.POINTER(HeadDesc), C.POINTER(NmsParams), _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "y11_detect_postprocess_push": (C.c_int, [_P, C.POINTER(HeadDesc), C.POINTER(NmsParams), _P, _P, _P, _P, _P, C.c_size_t,
                                              C.POINTER(Push), _P]),
    "y11_detect_postprocess_list": (C.c_int, [_P, C.POINTER(HeadDesc), C.POINTER(NmsParams), _P, _P, C.c_int32, _P, _P, _P, _P, _P,
                                              C.c_size_t, C.POINTER(Push), _P]),
    "y11_detect_postprocess_list_timed": (C.c_int, [_P, C.POINTER(HeadDesc), C.POINTER(NmsParams), _P, _P, C.c_int32, _P, _P, _P, _P, _P,
                                                    C.c_size_t, C.POINTER(C.c_float), _P]),
    "y11_plan_set_cls_emit": (C.c_int, [_P, C.c_int, C.POINTER(ClsEmit)]),
    "y11_wait_signals": (C.c_int, [_P, _P, C.c_int, C.c_uint32, _P]),
    "y11_detect_postprocess_timed": (C.c_int, [_P, C.POINTER(HeadDesc), C.POINTER(NmsParams), _P, _P, _P, _P, _P, C.c_size_t,
                                               C.POINTER(C.c_float), _P]),
    "y11_nms_batched": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, C.c_int, C.POINTER(NmsParams), _P, _P, _P, C.c_size_t, _P]),
    "y11_nms_workspace": (C.c_size_t, [C.c_int, C.c_int]),
    "y11_jpeg_create": (C.c_int, [_P, C.POINTER(_P)]),
    "y11_jpeg_destroy": (None, [_P]),
    "y11_jpeg_info": (C.c_int, [_P, _P, C.c_size_t, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "y11_jpeg_decode": (C.c_int, [_P, _P, C.c_size_t, _P, C.c_int32, C.c_int32, C.c_int32, _P]),
    "y11_draw_detections": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.POINTER(Font), C.c_int, _P]),
}

_lib = None


def load(build_if_missing: bool = True) -> C.CDLL:
    """dlopen liby11_b200.so (building it in-tree with nvcc if absent).  Raises if it cannot be had."""
    global _lib
    if _lib is not None:
        return _lib
    path = Path(os.environ.get("Y11_LIB", LIB_PATH))
    if not path.exists():
        if not build_if_missing:
            raise Y11Error(f"{path} missing: build it with `python -m yolo_infer_b200.build` (no CPU fallback exists)")
        from . import build as _build
        path = _build.build()
    lib = C.CDLL(str(path))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here == header/library mismatch
        fn.restype = res
        fn.argtypes = args
    if lib.y11_abi_version() != ABI_VERSION:
        raise Y11Error(f"ABI mismatch: library {lib.y11_abi_version()} vs binding {ABI_VERSION}")
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().y11_last_error()
        raise Y11Error(f"{what or 'liby11_b200'} failed (rc={rc}): {msg.decode() if msg else '?'}")


def view(t, c_off: int = 0, c: int | None = None) -> View:
    """NHWC torch tensor [B,H,W,C] (contiguous) -> channel-slice view."""
    ct = t.shape[-1]
    return View(t.data_ptr(), ct, c_off, ct - c_off if c is None else c)


NULL_VIEW = View(None, 0, 0, 0)
