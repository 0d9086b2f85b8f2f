"""The engine object that sits where ``ultralytics.YOLO`` sits in the reference (core/model.py:110: ``self.model``).

``YOLO(path).predict(source, **kw)`` reproduces the reference's detect path: source loaders ->
letterbox preprocess -> fused network -> Detect decode -> NMS -> scale_boxes -> ``Results`` - with every
arithmetic step executed by liby11_b200.so on an sm_100 GPU.  There is no CPU fallback: constructing the
engine without a B200-class device raises.
"""
from __future__ import annotations

import ctypes as C
import logging
import os
import math
import threading
import time
from pathlib import Path
from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import _cabi as cabi
from . import topology as T
from .network import CompiledNet, pack_weights
from .results import Results, ResultsBatch

logger = logging.getLogger(__name__)

# Y11_SPLIT_HOST=1: host-fed uint8 batches of >= 32 frames run as two half-batch pipelines on two streams (see
# YOLO.predict).  Off by default: measured on B200 it does not pay (YOLO11n 18.1 k vs 18.3 k img/s, YOLO11s 13.6 k vs 14.1 k) -
# the low-resolution layers are latency bound, so two half batches cost almost as much as two whole ones.
SPLIT_HOST_BATCH = os.environ.get("Y11_SPLIT_HOST", "0") != "0"
# Y11_FUSE_STEM=0: always run the letterbox kernel, even for frames that need neither resizing nor padding
FUSE_U8_STEM = os.environ.get("Y11_FUSE_STEM", "1") != "0"
MAX_CACHED_PIPELINES = 16   # CUDA-graph pipeline instances kept per engine (one per source shape x thresholds)
# Plans and CUDA graphs are BUILT by one thread at a time, process-wide: a stream capture must not overlap another thread's
# allocations / launches (the engines of a multi-device predict build their pipelines from one worker thread per device).
_BUILD_LOCK = threading.RLock()
PREDICT_DEFAULTS = dict(conf=0.25, iou=0.7, max_det=300, imgsz=640, rect=True, agnostic_nms=False, classes=None,
                        half=False, verbose=True, save=False, show=False, stream=False, batch=1, device=None,
                        multi_label=False, max_nms=30000,
                        devices=None,   # extension: list of CUDA device indices - the batch is image-sharded over them (one process)
                        decode="cv2",   # extension: "nvjpeg" decodes .jpg/.jpeg file sources on the GPU (decode.py; not bit-identical to cv2.imread)
                        graph=True)   # extension: False = launch the kernels one by one (no CUDA-graph pipeline) for file/array sources


def letterbox_geometry(h0: int, w0: int, new_shape: Tuple[int, int], auto: bool, stride: int = 32):
    """ultralytics LetterBox geometry (center, scaleup): returns new_w, new_h, top, left, H, W."""
    r = min(new_shape[0] / h0, new_shape[1] / w0)
    new_w, new_h = int(round(w0 * r)), int(round(h0 * r))
    dw, dh = new_shape[1] - new_w, new_shape[0] - new_h
    if auto:
        dw, dh = dw % stride, dh % stride
    dw, dh = dw / 2, dh / 2
    top, bottom = int(round(dh - 0.1)), int(round(dh + 0.1))
    left, right = int(round(dw - 0.1)), int(round(dw + 0.1))
    return new_w, new_h, top, left, new_h + top + bottom, new_w + left + right


def scale_geometry(net_hw: Tuple[int, int], orig_hw: Tuple[int, int]):
    """ultralytics scale_boxes geometry: gain, pad_x, pad_y."""
    gain = min(net_hw[0] / orig_hw[0], net_hw[1] / orig_hw[1])
    pad_x = round((net_hw[1] - orig_hw[1] * gain) / 2 - 0.1)
    pad_y = round((net_hw[0] - orig_hw[0] * gain) / 2 - 0.1)
    return gain, pad_x, pad_y


class DetectionNet:
    """What ``YOLO.model`` is in the reference (an nn.Module there): parameter access for get_model_info
    (/root/reference/core/model.py:237-247) and eval() for the benchmark (benchmarks/speed_benchmark.py:323)."""

    def __init__(self, scale: str, nc: int, state_dict: Dict[str, torch.Tensor], names: Dict[int, str]):
        self.scale, self.nc, self.names = scale, nc, names
        self._sd = {k: v for k, v in state_dict.items() if not k.endswith("num_batches_tracked")}
        shapes = T.param_shapes(scale, nc)
        missing = sorted(set(shapes) - set(self._sd))
        if missing:
            raise KeyError(f"state_dict is missing {len(missing)} tensors, e.g. {missing[:3]}")
        for k, s in shapes.items():
            if tuple(self._sd[k].shape) != s:
                raise ValueError(f"{k}: shape {tuple(self._sd[k].shape)} != expected {s}")
        self.training = False
        self.stride = torch.tensor([float(s) for s in T.STRIDES])
        self._params = None

    def state_dict(self) -> Dict[str, torch.Tensor]:
        return dict(self._sd)

    def named_parameters(self):
        if self._params is None:
            self._params = {k: torch.nn.Parameter(v, requires_grad=not k.startswith("model.23.dfl"))
                            for k, v in self._sd.items() if T.is_learnable(k)}
        return iter(self._params.items())

    def parameters(self):
        return (p for _, p in self.named_parameters())

    def eval(self):
        self.training = False
        return self

    def train(self, mode: bool = True):
        if mode:
            raise NotImplementedError("yolo_infer_b200 is an inference path; training is out of scope (SURVEY.md section 2 #8)")
        return self

    def to(self, *a, **k):
        return self

    def fuse(self):
        return self


def _load_state_dict(path: Path):
    import pickle
    try:
        obj = torch.load(str(path), map_location="cpu", weights_only=True)
    except (pickle.UnpicklingError, AttributeError, ModuleNotFoundError, RuntimeError) as e:
        # an ultralytics checkpoint is a pickle of the whole DetectionModel object: unreadable without ultralytics
        raise ValueError(f"{path}: not a plain state_dict file ({type(e).__name__}); pickled ultralytics model objects cannot be "
                         "read without ultralytics - re-save the weights as a state_dict (ultralytics key names, SURVEY.md A.4), "
                         "e.g. torch.save(YOLO(path).model.state_dict(), out)") from e
    if isinstance(obj, dict) and "state_dict" in obj:
        return obj["state_dict"], obj.get("scale"), obj.get("names"), obj.get("nc")
    if isinstance(obj, dict) and all(isinstance(v, torch.Tensor) for v in obj.values()):
        return obj, None, None, None
    raise ValueError(f"{path}: expected a plain state_dict (ultralytics key names, see SURVEY.md A.4); pickled "
                     "ultralytics model objects cannot be read without ultralytics")


def infer_scale(sd: Dict[str, torch.Tensor]) -> str:
    c0 = sd["model.0.conv.weight"].shape[0]
    deep = "model.2.m.1.cv1.conv.weight" in sd  # depth 1.0 -> two inner blocks
    return {16: "n", 32: "s", 96: "x"}.get(c0) or ("l" if deep else "m")


class YOLO:
    """Drop-in for the subset of ``ultralytics.YOLO`` the reference uses on the detect path."""

    def __init__(self, model: Union[str, Path] = "yolo11n.yaml", task: Optional[str] = "detect", verbose: bool = False,
                 init: str = "default", seed: int = 0):
        if task not in (None, "detect"):
            raise NotImplementedError(f"task={task!r}: only 'detect' is on the B200 hot path (SURVEY.md section 8)")
        self.task = "detect"
        self.ckpt_path = str(model)
        p = Path(str(model))
        stem = p.stem.lower()
        names = None
        nc = 80
        if p.suffix in (".yaml", ".yml"):
            if not (stem.startswith("yolo11") and len(stem) >= 7 and stem[6] in T.SCALES):
                raise ValueError(f"{model}: expected yolo11{{n,s,m,l,x}}.yaml")
            scale = stem[6]
            sd = T.default_state_dict(scale, nc, seed)
        elif p.suffix in (".pt", ".pth"):
            if p.exists():
                sd, scale, names, nc_ = _load_state_dict(p)
                scale = scale or infer_scale(sd)
                nc = nc_ or sd["model.23.cv3.0.2.weight"].shape[0]
            elif (stem.startswith("yolo11") and len(stem) >= 7 and stem[6] in T.SCALES
                  and (init == "random" or os.environ.get("Y11_ALLOW_RANDOM_INIT", "0") != "0")):
                # the reference would download pretrained weights here (core/model.py:106-110); there is no network, and
                # silently answering with random weights would return garbage detections: explicit opt-in only
                logger.warning("%s not found and cannot be downloaded offline: using random-init %s.yaml weights", model, stem[:7])
                scale = stem[6]
                sd = T.default_state_dict(scale, nc, seed)
            else:
                raise FileNotFoundError(f"{model}: checkpoint not found (the reference would download it, core/model.py:106-110; "
                                        "this path has no network access).  Pass an existing state_dict .pt, a yolo11{n,s,m,l,x}.yaml "
                                        "for a random-init model, or opt in with init='random' / Y11_ALLOW_RANDOM_INIT=1")
        else:
            raise ValueError(f"unsupported model file {model!r}")
        self.scale = scale
        self.nc = nc
        self.names: Dict[int, str] = names or {i: f"{i}" for i in range(nc)}
        self.model = DetectionNet(scale, nc, sd, self.names)
        self.overrides: Dict[str, object] = {}
        self.device: Optional[torch.device] = None
        self._engine = C.c_void_p()
        self._packed = None
        self._nets: Dict[Tuple[int, int, int], CompiledNet] = {}
        self._lock = threading.Lock()
        self._ws: Dict[Tuple, torch.Tensor] = {}
        self._lib = None
        self.conv_impl = cabi.IMPL_TCGEN05
        self.last_speed: Dict[str, float] = {}

    # ---- construction from an in-memory state_dict (tests, bench) -----------------------------------
    @classmethod
    def from_state_dict(cls, sd: Dict[str, torch.Tensor], scale: Optional[str] = None, nc: int = 80) -> "YOLO":
        self = cls.__new__(cls)
        YOLO.__init__(self, f"yolo11{scale or infer_scale(sd)}.yaml")
        self.nc = nc
        if len(self.names) != nc:
            self.names = {i: f"{i}" for i in range(nc)}
        self.model = DetectionNet(self.scale, nc, {k: v.detach().clone() for k, v in sd.items()}, self.names)
        return self

    # ---- device ---------------------------------------------------------------------------------
    def to(self, device) -> "YOLO":
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError(f"device={device!r}: yolo_infer_b200 runs on sm_100 GPUs only - there is no CPU fallback "
                               "(the reference's CPU path is kept only as the test oracle under oracle/)")
        if not torch.cuda.is_available():
            raise RuntimeError("CUDA device requested but none is visible; yolo_infer_b200 has no CPU fallback")
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        if self.device == dev and self._engine:
            return self
        self._release()
        self.device = dev
        self._lib = cabi.load()
        with torch.cuda.device(dev):
            cabi.check(self._lib.y11_create(C.byref(self._engine), dev.index), "y11_create")
            self._packed = pack_weights(self.scale, self.nc, self.model.state_dict(), dev)
            if getattr(self, "fp8_scales", None):
                from .network import pack_fp8
                pack_fp8(self.scale, self.nc, self.model.state_dict(), self._packed, self.fp8_scales, dev)
        return self

    def cuda(self) -> "YOLO":
        return self.to("cuda")

    def eval(self) -> "YOLO":
        return self

    def fuse(self) -> "YOLO":
        return self

    def _release(self):
        getattr(self, "_pipes", {}).clear()
        self._nets.clear()
        self._ws.clear()
        if self._engine and self._lib is not None:
            self._lib.y11_destroy(self._engine)
        self._engine = C.c_void_p()

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def _ensure_device(self, device=None):
        if device is not None:
            self.to(device)
        elif not self._engine:
            self.to("cuda")

    def compiled(self, B: int, H: int, W: int, chunks: int = 1, replica: int = 0, fold_upsample: Optional[bool] = None) -> CompiledNet:
        """replica > 0: an independent instance (own activation buffers / plan) for a pipeline that runs concurrently
        with another one on a second stream.  fold_upsample=False: keep Upsample/Concat as copy ops (one op per reference
        conv, which is what the weight conditioning walks)."""
        key = (B, H, W) if (chunks == 1 and replica == 0) else (B, H, W, chunks, replica)
        if fold_upsample is not None:
            key = key + ("fold", fold_upsample)
        net = self._nets.get(key)
        if net is None:
            with _BUILD_LOCK, torch.cuda.device(self.device):
                net = CompiledNet(self._engine, self.scale, self.nc, self._packed, B, H, W, self.device, self.conv_impl, chunks,
                                  fold_upsample, getattr(self, "fp8_scales", None))
                net.replica = replica
            self._nets[key] = net
        return net

    # ---- FP8 (e4m3) mode: the Blackwell counterpart of the reference's int8 post-training quantizers (quant.py) -----------------
    def enable_fp8(self, act_scales: Dict[str, float]) -> None:
        """act_scales: {producer conv name -> activation scale of its output tensor} for the edges of `network.fp8_pairs`: those
        tensors are stored as e4m3(value / scale) and their consumers run with e4m3 weights on tcgen05.mma.kind::f8f6f4.
        Plans and pipelines are rebuilt on next use."""
        from .network import pack_fp8
        self._ensure_device()
        self.fp8_scales = {k: float(v) for k, v in act_scales.items()}
        with torch.cuda.device(self.device):
            for k in [k for k in self._packed if k.endswith("#fp8")]:
                del self._packed[k]
            pack_fp8(self.scale, self.nc, self.model.state_dict(), self._packed, self.fp8_scales, self.device)
        getattr(self, "_pipes", {}).clear()
        self._nets.clear()

    def disable_fp8(self) -> None:
        self.fp8_scales = None
        getattr(self, "_pipes", {}).clear()
        self._nets.clear()

    def _workspace(self, key, nbytes: int) -> torch.Tensor:
        t = self._ws.get(key)
        if t is None or t.numel() < nbytes:
            t = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=self.device)
            self._ws[key] = t
        return t

    def _side_stream(self) -> torch.cuda.Stream:
        if getattr(self, "_side", None) is None or self._side.device != self.device:
            self._side = torch.cuda.Stream(self.device)
        return self._side

    def capture_stream(self) -> torch.cuda.Stream:
        """The stream CUDA-graph captures of this engine run on (one per engine, on the engine's device)."""
        if getattr(self, "_capture", None) is None or self._capture.device != self.device:
            self._capture = torch.cuda.Stream(self.device)
        return self._capture

    def shared_copy_stream(self) -> torch.cuda.Stream:
        """One upload stream per engine: host->device copies of concurrently running pipelines stay in submission order."""
        if getattr(self, "_copy", None) is None or self._copy.device != self.device:
            self._copy = torch.cuda.Stream(self.device)
        return self._copy

    def _fetch_results(self, det, count):
        """ONE device->host transfer of the whole batch's results (det fp32 [B,max_det,6] + count int32 [B]) into a pinned
        staging buffer; also the sync point of the call.  Returns (device copy of det, host copy of det, counts list): the
        Results objects keep the device rows (reference semantics) plus a host mirror, so `.cpu()` costs nothing more.
        det / count may be lists (the half-batch pipelines of a host-fed call): their rows are concatenated in order."""
        if isinstance(det, (list, tuple)):
            n = [d.shape[0] for d in det]
            shape = (sum(n),) + tuple(det[0].shape[1:])
            key = ("pinned_out_parts", shape, len(det))
            pin = self._ws.get(key)
            if pin is None:
                with torch.inference_mode(False):
                    pin = (torch.empty(shape, dtype=det[0].dtype).pin_memory(), torch.empty((shape[0],), dtype=count[0].dtype).pin_memory())
                self._ws[key] = pin
            det_dev = torch.empty(shape, dtype=det[0].dtype, device=self.device)
            o = 0
            for d, c, k in zip(det, count, n):
                pin[0][o:o + k].copy_(d, non_blocking=True)
                pin[1][o:o + k].copy_(c, non_blocking=True)
                det_dev[o:o + k].copy_(d, non_blocking=True)
                o += k
            torch.cuda.current_stream(self.device).synchronize()
            return det_dev, pin[0].clone(), pin[1].tolist()
        key = ("pinned_out", tuple(det.shape))
        pin = self._ws.get(key)
        if pin is None:
            with torch.inference_mode(False):  # staging buffers must stay writable from any mode
                pin = (torch.empty(det.shape, dtype=det.dtype).pin_memory(), torch.empty(count.shape, dtype=count.dtype).pin_memory())
            self._ws[key] = pin
        pin[0].copy_(det, non_blocking=True)
        pin[1].copy_(count, non_blocking=True)
        det_dev = det.clone()
        torch.cuda.current_stream(self.device).synchronize()
        return det_dev, pin[0].clone(), pin[1].tolist()

    # ---- stages (each is one or a few kernel launches through the C ABI) -----------------------------
    def preprocess_images(self, net: CompiledNet, frames: Sequence[torch.Tensor], geoms) -> None:
        """frames: device uint8 HWC BGR tensors; writes net.input (bf16 NHWC RGB /255)."""
        arr = (cabi.Image * len(frames))()
        for i, (f, g) in enumerate(zip(frames, geoms)):
            new_w, new_h, top, left, _, _ = g
            arr[i] = cabi.Image(f.data_ptr(), f.shape[0], f.shape[1], f.stride(0), new_h, new_w, top, left)
        host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
        dev = self._workspace(("img_desc", len(frames)), host.numel())
        dev[: host.numel()].copy_(host, non_blocking=False)
        s = torch.cuda.current_stream(self.device).cuda_stream
        net.set_stem_source(None)    # a pipeline sharing this plan may have pointed the stem at its uint8 frames
        cabi.check(self._lib.y11_letterbox(self._engine, dev.data_ptr(), len(frames), net.H, net.W, net.input.data_ptr(),
                                           C.c_void_p(s)), "y11_letterbox")

    def preprocess_tensor(self, net: CompiledNet, x: torch.Tensor, divisor: float) -> None:
        s = torch.cuda.current_stream(self.device).cuda_stream
        net.set_stem_source(None)
        cabi.check(self._lib.y11_nchw_f32_to_nhwc_bf16(self._engine, x.data_ptr(), net.B, net.H, net.W, divisor,
                                                       net.input.data_ptr(), C.c_void_p(s)), "y11_nchw_f32_to_nhwc_bf16")

    def forward(self, net: CompiledNet, cls_emit: Optional[float] = None) -> None:
        """cls_emit = conf of a single-label post-processing that follows: the class-logit convs then list (anchor, class,
        max logit) of the anchors that can pass conf instead of storing logits (CompiledNet.set_cls_emit; `postprocess` picks the
        lists up).  None (default): logits are stored - `raw_head`, multi-label and dense decode need them."""
        net.set_cls_emit(cls_emit)
        net.run(torch.cuda.current_stream(self.device).cuda_stream)

    def result_buffers(self, B: int, max_det: int, rep: int = 0, flat: Optional[torch.Tensor] = None):
        """(det [B,max_det,6] fp32, count [B] int32, ncand [B] int32, flat) where det and count are views of ONE flat buffer
        ([B*max_det*6] fp32 followed by [B] int32) so that a multi-GPU caller moves a step's results as one piece.  `flat`:
        use this buffer instead of an engine-owned one - it may live on ANOTHER GPU (an NVLink peer mapping): the NMS kernel
        then writes its results straight into the gathering rank's memory (parallel.ResultExchange)."""
        okey = ("post_out", B, max_det, rep, flat.data_ptr() if flat is not None else None)
        out = self._ws.get(okey)
        if out is None:
            n = B * max_det * 6 + B
            if flat is None:
                flat = torch.zeros((n,), dtype=torch.float32, device=self.device)
            assert flat.numel() == n and flat.dtype == torch.float32 and flat.is_contiguous()
            out = (flat[: B * max_det * 6].view(B, max_det, 6), flat[B * max_det * 6:].view(torch.int32),
                   torch.zeros((B,), dtype=torch.int32, device=self.device), flat)
            self._ws[okey] = out
        return out

    def postprocess(self, net: CompiledNet, scale_rows: Optional[torch.Tensor], conf: float, iou: float, max_det: int,
                    agnostic: bool = False, multi_label: bool = False, max_nms: int = 30000,
                    out_flat: Optional[torch.Tensor] = None, push: Optional[Tuple[int, int]] = None):
        """-> (det fp32 [B,max_det,6], count int32 [B], ncand int32 [B]) device tensors.
        push = (done_counter device pointer, signal device pointer): see y11_detect_postprocess_push."""
        B = net.B
        nbytes = self._lib.y11_postprocess_workspace(B, net.A, self.nc, int(multi_label), max_nms)
        rep = getattr(net, "replica", 0)
        ws = self._workspace(("post", B, net.A, multi_label, rep), nbytes)
        det, count, ncand, _flat = self.result_buffers(B, max_det, rep, out_flat)
        hd = net.head_desc()
        p = cabi.NmsParams(conf, iou, max_det, max_nms, 7680, int(agnostic), int(multi_label))
        s = torch.cuda.current_stream(self.device).cuda_stream
        srows = scale_rows.data_ptr() if scale_rows is not None else None
        if net.emit_conf is not None:
            # the plan ran in class-emit mode (CompiledNet.set_cls_emit): the class logits were reduced in the conv epilogues
            assert not multi_label and net.emit_conf == conf, "class-emit plan used with different post-processing parameters"
            pp = cabi.Push(push[0], push[1]) if push is not None else None
            cabi.check(self._lib.y11_detect_postprocess_list(self._engine, C.byref(hd), C.byref(p), net.emit_list.data_ptr(),
                                                             net.emit_count.data_ptr(), net.A, srows, det.data_ptr(),
                                                             count.data_ptr(), ncand.data_ptr(), ws.data_ptr(), ws.numel(),
                                                             C.byref(pp) if pp is not None else None, C.c_void_p(s)),
                       "y11_detect_postprocess_list")
        elif push is not None:
            pp = cabi.Push(push[0], push[1])
            cabi.check(self._lib.y11_detect_postprocess_push(self._engine, C.byref(hd), C.byref(p), srows, det.data_ptr(),
                                                             count.data_ptr(), ncand.data_ptr(), ws.data_ptr(), ws.numel(),
                                                             C.byref(pp), C.c_void_p(s)), "y11_detect_postprocess_push")
        else:
            cabi.check(self._lib.y11_detect_postprocess(self._engine, C.byref(hd), C.byref(p), srows, det.data_ptr(),
                                                        count.data_ptr(), ncand.data_ptr(), ws.data_ptr(), ws.numel(),
                                                        C.c_void_p(s)), "y11_detect_postprocess")
        return det, count, ncand

    def postprocess_timed(self, net: CompiledNet, scale_rows, conf: float, iou: float, max_det: int, multi_label: bool = False,
                          max_nms: int = 30000) -> Tuple[float, float]:
        """Measurement only: (ms of decode + compaction, ms of sort + NMS) of one post-processing pass on net's current head."""
        B = net.B
        nbytes = self._lib.y11_postprocess_workspace(B, net.A, self.nc, int(multi_label), max_nms)
        ws = self._workspace(("post", B, net.A, multi_label, getattr(net, "replica", 0)), nbytes)
        det, count, ncand, _ = self.result_buffers(B, max_det, getattr(net, "replica", 0))
        hd = net.head_desc()
        p = cabi.NmsParams(conf, iou, max_det, max_nms, 7680, 0, int(multi_label))
        ms = (C.c_float * 2)()
        s = torch.cuda.current_stream(self.device).cuda_stream
        if net.emit_conf is not None:      # the plan last ran in class-emit mode: time the list path on its lists
            assert not multi_label and net.emit_conf == conf
            cabi.check(self._lib.y11_detect_postprocess_list_timed(self._engine, C.byref(hd), C.byref(p), net.emit_list.data_ptr(),
                                                                   net.emit_count.data_ptr(), net.A,
                                                                   scale_rows.data_ptr() if scale_rows is not None else None,
                                                                   det.data_ptr(), count.data_ptr(), ncand.data_ptr(), ws.data_ptr(),
                                                                   ws.numel(), ms, C.c_void_p(s)), "y11_detect_postprocess_list_timed")
            return float(ms[0]), float(ms[1])
        cabi.check(self._lib.y11_detect_postprocess_timed(self._engine, C.byref(hd), C.byref(p),
                                                          scale_rows.data_ptr() if scale_rows is not None else None, det.data_ptr(),
                                                          count.data_ptr(), ncand.data_ptr(), ws.data_ptr(), ws.numel(), ms,
                                                          C.c_void_p(s)), "y11_detect_postprocess_timed")
        return float(ms[0]), float(ms[1])

    # ---- CUDA-graph pipeline for fixed-shape uint8 batches (throughput and batch-1 latency modes) ----------------
    def result_flat(self, net: CompiledNet, max_det: int) -> torch.Tensor:
        """The engine-owned flat [B*max_det*6 fp32 | B int32] buffer behind the (det, count) tensors `postprocess` returns."""
        return self.result_buffers(net.B, max_det, getattr(net, "replica", 0))[3]

    def pipeline(self, B: int, h0: int, w0: int, imgsz=640, rect: bool = True, conf: float = 0.25, iou: float = 0.7,
                 max_det: int = 300, agnostic: bool = False, multi_label: bool = False, frames: Optional[torch.Tensor] = None,
                 graph: bool = True, replica: int = 0, max_nms: int = 30000, kind: str = "u8",
                 out_flat: Optional[torch.Tensor] = None, push: Optional[Tuple[int, int]] = None) -> "GraphedPipeline":
        """preprocess -> forward -> decode/NMS for a fixed-shape batch, captured once as a CUDA graph (one replay per call).
        kind "u8": B uint8 BGR frames of h0 x w0 (letterbox, or read directly by the stem when no resize is needed);
        kind "f32": a float [B,3,h0,w0] tensor (LoadTensor semantics: /255 iff max > 1, decided on the device).
        `frames`: optional device tensor to bind as the graph's input (zero-copy); otherwise the pipeline owns a static input
        buffer that `run(src)` fills with one async copy (H2D from pinned memory or D2D).
        `out_flat` / `push`: result buffer override and push signal (multi-GPU result push, see `postprocess`)."""
        self._ensure_device()
        key = (kind, B, h0, w0, imgsz if isinstance(imgsz, int) else tuple(imgsz), rect, conf, iou, max_det, agnostic, multi_label,
               frames.data_ptr() if frames is not None else None, graph, replica, max_nms,
               out_flat.data_ptr() if out_flat is not None else None, push)
        p = self._pipes.get(key) if hasattr(self, "_pipes") else None
        if p is None:
            if not hasattr(self, "_pipes"):
                self._pipes = {}
            with _BUILD_LOCK, torch.inference_mode(False):  # static buffers must stay writable from any mode
                p = GraphedPipeline(self, B, h0, w0, imgsz, rect, conf, iou, max_det, agnostic, multi_label, frames, graph, replica,
                                    max_nms, kind, out_flat, push)
            if len(self._pipes) >= MAX_CACHED_PIPELINES:   # dicts keep insertion order: drop the oldest instance
                self._pipes.pop(next(iter(self._pipes)))
            self._pipes[key] = p
        return p

    # ---- sources ----------------------------------------------------------------------------------
    @staticmethod
    def _load_sources(source) -> Tuple[List[np.ndarray], List[str]]:
        import cv2
        items = source if isinstance(source, (list, tuple)) else [source]
        imgs, paths = [], []
        for i, it in enumerate(items):
            if isinstance(it, (str, Path)):
                im = cv2.imread(str(it))  # BGR, as ultralytics LoadImagesAndVideos does
                if im is None:
                    raise FileNotFoundError(f"cannot read image {it}")
                imgs.append(im)
                paths.append(str(it))
            elif isinstance(it, np.ndarray):
                if it.ndim != 3 or it.shape[2] != 3 or it.dtype != np.uint8:
                    raise ValueError(f"ndarray source must be HxWx3 uint8 BGR, got {it.shape} {it.dtype}")
                imgs.append(np.ascontiguousarray(it))
                paths.append(f"image{i}.jpg")
            else:
                raise TypeError(f"unsupported source element {type(it)}")
        return imgs, paths

    # ---- predict ----------------------------------------------------------------------------------
    def predict(self, source=None, stream: bool = False, **kwargs) -> List[Results]:
        args = {**PREDICT_DEFAULTS, **self.overrides, **kwargs}
        unknown = set(kwargs) - set(PREDICT_DEFAULTS)
        if unknown:
            logger.debug("predict: ignoring unsupported kwargs %s", sorted(unknown))
        if args["half"]:
            logger.debug("half=True: the B200 path always computes in bf16 with fp32 accumulation")
        if source is None:
            raise ValueError("predict: source is required")
        if stream:
            return self._predict_stream(source, kwargs)
        if args["devices"] is not None and len(list(args["devices"])) > 1:
            return self._predict_multi_device(source, list(args["devices"]), kwargs)
        if args["devices"] is not None and len(list(args["devices"])) == 1 and args["device"] is None:
            args["device"] = f"cuda:{int(list(args['devices'])[0])}"
        self._ensure_device(args["device"])
        imgsz = args["imgsz"]
        new_shape = (imgsz, imgsz) if isinstance(imgsz, int) else tuple(imgsz)
        pargs = dict(imgsz=imgsz, rect=bool(args["rect"]), conf=float(args["conf"]), iou=float(args["iou"]),
                     max_det=int(args["max_det"]), agnostic=bool(args["agnostic_nms"]), multi_label=bool(args["multi_label"]),
                     max_nms=int(args["max_nms"]), graph=bool(args["graph"]))
        classes = torch.as_tensor(list(args["classes"])) if args["classes"] is not None else None

        def finish(det, det_h, counts, speed, orig_imgs, paths, shapes):
            """One Results per image, built on first access (ResultsBatch); rows are sliced lazily too."""
            if classes is None:
                return ResultsBatch(det, det_h, counts, self.names, shapes, paths, orig_imgs, speed)
            results = []
            for i, n in enumerate(counts):
                img_i = orig_imgs[i] if orig_imgs is not None else None
                path_i = paths[i] if paths is not None else f"image{i}.jpg"
                if classes is None:
                    results.append(Results(img_i, path_i, self.names, None, shapes[i], speed, None, (det, det_h, i, n)))
                    continue
                d, dh = det[i, :n], det_h[i, :n]
                keep = torch.isin(dh[:, 5].long(), classes)
                results.append(Results(img_i, path_i, self.names, d[keep.to(d.device)], shapes[i], speed, dh[keep]))
            return results

        # ---- float tensor sources [B,3,H,W] (the reference harness's input, benchmarks/speed_benchmark.py:100-102) ------------
        if isinstance(source, torch.Tensor) and source.is_floating_point():
            x = source[None] if source.ndim == 3 else source
            if x.ndim != 4 or x.shape[1] != 3 or x.shape[2] % 32 or x.shape[3] % 32:
                raise ValueError(f"tensor source must be [B,3,H,W] with H,W % 32 == 0, got {tuple(x.shape)}")
            B, _, H, W = x.shape
            with self._lock, torch.cuda.device(self.device), torch.inference_mode():
                if x.dtype != torch.float32 or not x.is_contiguous():
                    x = x.to(torch.float32).contiguous()
                pipe = self.pipeline(B, H, W, kind="f32", **pargs)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                det, count, _ = pipe.run(x)
                e1.record()
                det, det_h, counts = self._fetch_results(det, count)
                speed = pipe.speed(e0.elapsed_time(e1) / B)
                self.last_speed = speed
                results = finish(det, det_h, counts, speed, None, None, [(H, W)] * B)
            if args["verbose"]:
                logger.info("%d image(s) %dx%d (tensor): %.2f ms per image", B, H, W, sum(speed.values()))
            return results

        # ---- uint8 sources: files / BGR arrays / uint8 batch tensors ------------------------------------------------------
        orig_imgs = paths = None
        if isinstance(source, torch.Tensor):
            if source.dtype != torch.uint8 or source.ndim != 4 or source.shape[-1] != 3:
                raise ValueError("uint8 tensor source must be [B,H,W,3] BGR")
        elif args["decode"] == "nvjpeg":
            # .jpg files decoded on the GPU: the frames are born in device memory; equal-sized ones form a device batch
            from .decode import GpuJpegDecoder, is_jpeg_path
            items = list(source) if isinstance(source, (list, tuple)) else [source]
            if not all(is_jpeg_path(it) for it in items):
                raise ValueError("decode='nvjpeg' takes .jpg/.jpeg file paths")
            with self._lock, torch.cuda.device(self.device):
                if getattr(self, "_jpeg", None) is None or self._jpeg.device != self.device:
                    self._jpeg = GpuJpegDecoder(self.device, self._engine)
                frames = [self._jpeg.decode(it) for it in items]
            if len({tuple(f.shape) for f in frames}) != 1:
                raise ValueError("decode='nvjpeg': the files of one call must have one size (group them by size, as a video / camera "
                                 "folder is)")
            res = self.predict(torch.stack(frames), **{**kwargs, "decode": "cv2"})
            for r, it, f in zip(res, items, frames):
                r.path, r.orig_img = str(it), f          # the decoded frame stays on the device (draw.draw_detections takes it there)
            return res
        elif args["decode"] != "cv2":
            raise ValueError(f"decode={args['decode']!r}: expected 'cv2' or 'nvjpeg'")
        else:
            orig_imgs, paths = self._load_sources(source)
        one_shape = orig_imgs is None or len({im.shape for im in orig_imgs}) == 1
        if one_shape:
            # a fixed-shape uint8 batch [B,H,W,3] BGR (pinned host or device; files / arrays of ONE shape - a video stream, a demo
            # frame, a folder of equal-sized images - are staged into a pinned batch): the whole path is one CUDA-graph replay
            # (reference call sites: demos/detection_demo.py:87-93, 190-196, 280-286); graph=False launches kernel by kernel
            with self._lock, torch.cuda.device(self.device), torch.inference_mode():
                if orig_imgs is not None:   # stage the frames in a pinned batch (under the lock: the buffer is per engine)
                    key = ("pinned_frames", len(orig_imgs)) + tuple(orig_imgs[0].shape)
                    source = self._ws.get(key)
                    if source is None:
                        with torch.inference_mode(False):
                            source = torch.empty((len(orig_imgs),) + tuple(orig_imgs[0].shape), dtype=torch.uint8).pin_memory()
                        self._ws[key] = source
                    stage_np = source.numpy()
                    for i, im in enumerate(orig_imgs):
                        stage_np[i] = im
                B, h0, w0, _ = source.shape
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                if (not source.is_cuda) and B >= 32 and B % 8 == 0 and SPLIT_HOST_BATCH and pargs["graph"]:
                    # Host-fed batch as two half-batch pipelines on two streams (off by default: measured no gain)
                    Bh = B // 2
                    cur = torch.cuda.current_stream(self.device)
                    side = self._side_stream()
                    pipes = [self.pipeline(Bh, h0, w0, replica=r, **pargs) for r in (0, 1)]
                    side.wait_stream(cur)
                    det0, count0, _ = pipes[0].run(source[:Bh])
                    with torch.cuda.stream(side):   # its uploads queue behind the first half's on the shared copy stream
                        det1, count1, _ = pipes[1].run(source[Bh:])
                    cur.wait_stream(side)
                    e1.record()
                    det, det_h, counts = self._fetch_results([det0, det1], [count0, count1])
                    pipe = pipes[0]
                else:
                    pipe = self.pipeline(B, h0, w0, **pargs)
                    det, count, _ = pipe.run(source)
                    e1.record()
                    det, det_h, counts = self._fetch_results(det, count)
                # the stages of one graph replay are not separable call by call: the call's time is split by the stage shares
                # measured once when the pipeline was built (GraphedPipeline.speed)
                speed = pipe.speed(e0.elapsed_time(e1) / B)
                self.last_speed = speed
                results = finish(det, det_h, counts, speed, orig_imgs, paths, [(h0, w0)] * B)
            if args["verbose"]:
                logger.info("%d image(s) %dx%d: %.2f ms per image (preprocess + forward + decode + NMS)", B, h0, w0, sum(speed.values()))
            return results

        # ---- mixed shapes: kernel-by-kernel launches, square letterbox (ultralytics: auto=False when shapes differ) ---------------
        orig_shapes = [(int(im.shape[0]), int(im.shape[1])) for im in orig_imgs]
        geoms = [letterbox_geometry(h, w, new_shape, False) for h, w in orig_shapes]
        H, W = geoms[0][4], geoms[0][5]
        rows = []
        for (h0, w0) in orig_shapes:
            gain, px, py = scale_geometry((H, W), (h0, w0))
            rows.append([gain, float(px), float(py), float(w0), float(h0)])
        results = self.predict_letterboxed(orig_imgs, [g[:4] for g in geoms], H, W, rows, paths=paths, classes=classes, **pargs)
        if args["verbose"]:
            sp = self.last_speed
            logger.info("%d image(s) %dx%d: %.2f ms pre, %.2f ms inference, %.2f ms post per image", len(orig_imgs), H, W,
                        sp["preprocess"], sp["inference"], sp["postprocess"])
        return results

    def predict_letterboxed(self, imgs: Sequence[np.ndarray], geoms: Sequence[Tuple[int, int, int, int]], H: int, W: int,
                            scale_rows: Sequence[Sequence[float]], paths: Optional[Sequence[str]] = None, classes=None,
                            conf: float = 0.25, iou: float = 0.7, max_det: int = 300, agnostic: bool = False,
                            multi_label: bool = False, max_nms: int = 30000, **_unused) -> List[Results]:
        """Kernel-by-kernel pass over BGR uint8 arrays with an EXPLICIT letterbox geometry per image - (new_w, new_h, top, left) on an
        H x W canvas - and explicit un-letterbox rows [gain, pad_x, pad_y, w0, h0].  `predict` uses it for batches of mixed shapes;
        the val path uses it for ultralytics' rect batches (val.rect_batches), whose geometry differs from predict's."""
        self._ensure_device()
        with self._lock, torch.cuda.device(self.device), torch.inference_mode():
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            ev[0].record()
            frames = [torch.from_numpy(np.ascontiguousarray(im)).pin_memory().to(self.device, non_blocking=True) for im in imgs]
            B = len(frames)
            net = self.compiled(B, H, W)
            self.preprocess_images(net, frames, [(g[0], g[1], g[2], g[3], H, W) for g in geoms])
            ev[1].record()
            self.forward(net, None if multi_label else conf)
            ev[2].record()
            rows = torch.tensor([list(map(float, r)) for r in scale_rows], dtype=torch.float32).to(self.device, non_blocking=True)
            det, count, ncand = self.postprocess(net, rows, conf, iou, max_det, agnostic, multi_label, max_nms)
            ev[3].record()
            det, det_h, counts = self._fetch_results(det, count)  # one D2H; also the sync point of the call
            speed = {"preprocess": ev[0].elapsed_time(ev[1]) / B, "inference": ev[1].elapsed_time(ev[2]) / B,
                     "postprocess": ev[2].elapsed_time(ev[3]) / B}
            self.last_speed = speed
            results = []
            for i, n in enumerate(counts):
                shape_i = (int(imgs[i].shape[0]), int(imgs[i].shape[1]))
                path_i = paths[i] if paths is not None else f"image{i}.jpg"
                if classes is None:
                    results.append(Results(imgs[i], path_i, self.names, None, shape_i, speed, None, (det, det_h, i, n)))
                    continue
                d, dh = det[i, :n], det_h[i, :n]
                keep = torch.isin(dh[:, 5].long(), classes)
                results.append(Results(imgs[i], path_i, self.names, d[keep.to(d.device)], shape_i, speed, dh[keep]))
        return results

    __call__ = predict

    # ---- stream mode: a generator over batches, two batches in flight ---------------------------------------------------------------
    def _predict_stream(self, batches, kwargs):
        """`predict(batches, stream=True)` (ultralytics' stream mode returns a generator too): `batches` is an iterable of fixed-shape
        uint8 batch tensors [B,H,W,3] (pinned host or device).  Yields one ResultsBatch per input batch, in order.  Two pipeline
        replicas on two streams: while batch i computes, the frames of batch i+1 cross PCIe and the results of batch i-1 are read
        back - a synchronous `predict` call cannot overlap its own upload with anything."""
        args = {**PREDICT_DEFAULTS, **self.overrides, **kwargs}
        self._ensure_device(args["device"])
        pargs = dict(imgsz=args["imgsz"], rect=bool(args["rect"]), conf=float(args["conf"]), iou=float(args["iou"]),
                     max_det=int(args["max_det"]), agnostic=bool(args["agnostic_nms"]), multi_label=bool(args["multi_label"]),
                     max_nms=int(args["max_nms"]), graph=bool(args["graph"]))
        dev = self.device
        with torch.cuda.device(dev):
            if getattr(self, "_stream_lanes", None) is None or self._stream_lanes[0].device != dev:
                self._stream_lanes = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]
        lanes = self._stream_lanes

        def finish(p):
            det_dev, pin, ev, shape, t0 = p
            ev.synchronize()
            B = det_dev.shape[0]
            speed = {"preprocess": 0.0, "inference": 1e3 * (time.perf_counter() - t0) / B, "postprocess": 0.0}   # latency of the batch
            # the pinned staging buffers are reused two batches later: hand out a copy (461 KB for 64 images)
            return ResultsBatch(det_dev, pin[0].clone(), pin[1].tolist(), self.names, shape, speed=speed)

        pending = None
        for i, src in enumerate(batches):
            if not (isinstance(src, torch.Tensor) and src.dtype == torch.uint8 and src.ndim == 4 and src.shape[-1] == 3):
                raise ValueError("stream=True takes an iterable of uint8 [B,H,W,3] BGR batch tensors")
            r = i & 1
            B, h0, w0, _ = src.shape
            with self._lock, torch.cuda.device(dev), torch.inference_mode():
                pipe = self.pipeline(B, h0, w0, replica=r, **pargs)
                key = ("pinned_stream_out", B, pargs["max_det"], r)
                pin = self._ws.get(key)
                if pin is None:
                    with torch.inference_mode(False):
                        pin = (torch.empty((B, pargs["max_det"], 6), dtype=torch.float32).pin_memory(),
                               torch.empty((B,), dtype=torch.int32).pin_memory())
                    self._ws[key] = pin
                t0 = time.perf_counter()
                lanes[r].wait_stream(torch.cuda.current_stream(dev))
                with torch.cuda.stream(lanes[r]):
                    det, count, _ = pipe.run(src)
                    det_dev = det.clone()                       # the pipeline's static result buffer is rewritten two batches later
                    pin[0].copy_(det, non_blocking=True)
                    pin[1].copy_(count, non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(lanes[r])
            if pending is not None:
                yield finish(pending)       # the batch before this one: its lane has had a whole batch time to finish
            pending = (det_dev, (pin[0], pin[1]), ev, (h0, w0), t0)
        if pending is not None:
            yield finish(pending)

    # ---- multi-GPU in ONE process: the README's "Multi-GPU ... inference" (/root/reference/README.md:13) as a library call -----
    def _replica_on(self, index: int) -> "YOLO":
        """An engine with the same weights on device `index` (weights are replicated once)."""
        if self.device is not None and self.device.index == index and self._engine:
            return self
        reps = self.__dict__.setdefault("_replicas", {})
        rep = reps.get(index)
        if rep is None:
            with _BUILD_LOCK:
                rep = reps.get(index)
                if rep is None:
                    rep = YOLO.from_state_dict(self.model.state_dict(), self.scale, self.nc)
                    rep.names, rep.overrides = self.names, dict(self.overrides)
                    rep.to(f"cuda:{index}")
                    reps[index] = rep
        return rep

    def _predict_multi_device(self, source, devices: List[int], kwargs) -> List[Results]:
        """Image-sharded predict over several GPUs of one box: the batch is split contiguously (parallel.shard_range), every
        device runs the full single-device path on its slice from its own worker thread (own pinned staging buffers, own
        copy / compute streams, own CUDA graphs), and the Results come back in source order.  No collective: the only
        exchange is each device's one result D2H."""
        from concurrent.futures import ThreadPoolExecutor
        from .parallel import shard_range
        devices = [int(d) for d in devices]
        if len(set(devices)) != len(devices):
            raise ValueError(f"devices must be distinct, got {devices}")
        if isinstance(source, torch.Tensor):
            n = source.shape[0] if source.ndim == 4 else 1
            take = (lambda lo, hi: source[lo:hi]) if source.ndim == 4 else (lambda lo, hi: source)
        else:
            items = list(source) if isinstance(source, (list, tuple)) else [source]
            n = len(items)
            take = lambda lo, hi: items[lo:hi]
        world = min(len(devices), n)
        kw = {k: v for k, v in kwargs.items() if k not in ("devices", "device")}
        pool = self.__dict__.get("_pool")
        if pool is None or pool._max_workers < world:
            pool = self.__dict__["_pool"] = ThreadPoolExecutor(max_workers=max(world, 8), thread_name_prefix="y11dev")

        def work(r: int):
            lo, hi = shard_range(n, r, world)
            rep = self._replica_on(devices[r])
            with torch.cuda.device(rep.device):
                return rep.predict(take(lo, hi), **kw)

        parts = list(pool.map(work, range(world)))
        return [res for part in parts for res in part]

    # ---- out-of-scope surface: fail loudly, never silently fall back -----------------------------------
    def val(self, data=None, **kwargs):
        """Detection validation (ultralytics `YOLO.val` as the reference uses it, `core/model.py:180-195`): runs `predict`
        with multi-label NMS at conf 0.001 / iou 0.6 over the dataset and scores it on the host.  Returns an object with
        `.box.{map, map50, map75, mp, mr}` and `.speed` (`core/validator.py:339-359`)."""
        if data is None:
            raise ValueError("val: `data` (dataset yaml or image directory with YOLO labels) is required - no dataset is bundled")
        from .val import validate
        self._ensure_device(kwargs.pop("device", None))
        return validate(self, data, **kwargs)

    def train(self, *a, **k):
        raise NotImplementedError("training is out of scope for the B200 inference path (SURVEY.md section 2 #8)")

    def export(self, *a, **k):
        raise NotImplementedError("export is out of scope: the B200 path is not a TensorRT/ONNX export")

    def save(self, path: Union[str, Path]):
        torch.save({"state_dict": self.model.state_dict(), "scale": self.scale, "nc": self.nc, "names": self.names}, str(path))

    def info(self, *a, **k):
        n = T.count_parameters(self.scale, self.nc)
        return {"scale": self.scale, "parameters": n}


class GraphedPipeline:
    """One fixed-shape instance of the whole hot path, replayed as CUDA graphs.

    All device buffers (input frames, letterbox descriptors, activations, post-processing workspace, results) are static,
    so a call is: an async copy of the input (skipped when the caller binds its own device tensor) + graph launches.
    The launches inside are exactly the ones the kernel-by-kernel path issues; CUDA graphs only remove the per-launch CPU
    cost (~96 launches for YOLO11n/s), which dominates at batch 1.

    kind "u8"  : uint8 BGR frames [B,h0,w0,3] -> letterbox kernel (or read directly by the stem when no resize/pad is needed).
    kind "f32" : float [B,3,H,W] tensors (ultralytics LoadTensor semantics; what the reference's SpeedBenchmark feeds):
                 device-side max -> /255 rule -> bf16 NHWC, no letterbox, no host sync.

    Host-fed uint8 batches (`frames is None`, B a multiple of 4, B >= 16) are CHUNKED: the frames cross PCIe in four pieces on
    a copy stream, and preprocess + layers 0-4 of chunk c (one graph per chunk) run while chunk c+1 is still in flight; the
    rest of the network, decode and NMS run once on the whole batch.  The 78.6 MB upload of a 64-frame batch takes 1.42 ms
    against 2.6-3.9 ms of compute: unchunked it is simply added to every call.
    """

    def __init__(self, eng: YOLO, B: int, h0: int, w0: int, imgsz, rect: bool, conf: float, iou: float, max_det: int,
                 agnostic: bool, multi_label: bool, frames: Optional[torch.Tensor], graph: bool, replica: int = 0,
                 max_nms: int = 30000, kind: str = "u8", out_flat: Optional[torch.Tensor] = None,
                 push: Optional[Tuple[int, int]] = None):
        assert kind in ("u8", "f32")
        self.eng, self.B, self.h0, self.w0, self.kind = eng, B, h0, w0, kind
        self.conf, self.iou, self.max_det, self.agnostic, self.multi_label = conf, iou, max_det, agnostic, multi_label
        self.max_nms, self.out_flat, self.push = max_nms, out_flat, push
        self._emit = None if multi_label else conf     # class-emit mode of the class-logit convs (CompiledNet.set_cls_emit)
        dev = eng.device
        new_shape = (imgsz, imgsz) if isinstance(imgsz, int) else tuple(imgsz)
        if kind == "u8":
            geom = letterbox_geometry(h0, w0, new_shape, bool(rect))
            self.H, self.W = geom[4], geom[5]
        else:
            geom = None
            self.H, self.W = h0, w0          # tensor sources are used as they are (LoadTensor: H, W % 32 == 0)
        self.owns_input = frames is None
        self.chunks = 4 if (kind == "u8" and self.owns_input and graph and B % 4 == 0 and B >= 16) else 1
        if os.environ.get("Y11_CHUNKS") and kind == "u8":   # experiment knob: chunk-major prefix also for device-resident frames
            k = int(os.environ["Y11_CHUNKS"])
            self.chunks = k if (k >= 1 and B % k == 0 and graph) else self.chunks
        with torch.cuda.device(dev):
            self.net = eng.compiled(B, self.H, self.W, self.chunks, replica)
            self.fused_stem = False
            if kind == "u8":
                self.frames = frames if frames is not None else torch.zeros((B, h0, w0, 3), dtype=torch.uint8, device=dev)
                assert self.frames.shape == (B, h0, w0, 3) and self.frames.dtype == torch.uint8 and self.frames.is_cuda
                arr = (cabi.Image * B)()
                for i in range(B):
                    f = self.frames[i]
                    arr[i] = cabi.Image(f.data_ptr(), h0, w0, f.stride(0), geom[1], geom[0], geom[2], geom[3])
                self.desc = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(dev)
                self.desc_stride = C.sizeof(cabi.Image)
                # Frames already at network resolution (no resize, no padding): the stem reads the uint8 frames itself - bit
                # for bit the letterbox kernel's conversion - and the letterbox launch with its bf16 round trip is skipped.
                new_w, new_h, top, left = geom[0], geom[1], geom[2], geom[3]
                self.fused_stem = bool(FUSE_U8_STEM and new_w == w0 and new_h == h0 and top == 0 and left == 0 and self.H == h0
                                       and self.W == w0 and self.frames.data_ptr() % 4 == 0 and self.frames.stride(1) % 4 == 0
                                       and self.frames.stride(0) % 4 == 0)
            else:
                self.frames = frames if frames is not None else torch.zeros((B, 3, h0, w0), dtype=torch.float32, device=dev)
                assert self.frames.shape == (B, 3, h0, w0) and self.frames.dtype == torch.float32 and self.frames.is_cuda
                self.max_scratch = torch.zeros((1,), dtype=torch.int32, device=dev)
            gain, px, py = scale_geometry((self.H, self.W), (h0, w0))
            self.scale_rows = torch.tensor([[gain, float(px), float(py), float(w0), float(h0)]] * B, dtype=torch.float32, device=dev)
            self.graphs: List[torch.cuda.CUDAGraph] = []
            self._stage_fns = self._stages()
            for fn in self._stage_fns:           # warm-up: allocates workspaces, sets function attributes
                fn()
            torch.cuda.synchronize(dev)
            self._measure_stage_shares()
            if graph:
                for fn in self._stage_fns:
                    g = torch.cuda.CUDAGraph()
                    # an explicit capture stream ON THIS DEVICE: torch.cuda.graph's default capture stream is created once per
                    # process, on whichever device captured first - a capture for an engine on another GPU would switch the
                    # current device to that one and launch this device's kernels outside the capture
                    with torch.cuda.graph(g, stream=eng.capture_stream(), capture_error_mode="thread_local"):
                        fn()
                    self.graphs.append(g)
            if self.chunks > 1:
                self.copy_stream = eng.shared_copy_stream()
                self.copy_events = [torch.cuda.Event() for _ in range(self.chunks)]
        post_launches = 4 if (multi_label or self.net.A >= 65536) else 2   # (count, scan, write | one-pass decode) + sort/NMS
        pre_launches = 2 if kind == "f32" else (0 if self.fused_stem else self.chunks)   # max + convert | letterbox per chunk
        self.launches = pre_launches + self.net.n_launches + post_launches

    # ---- the stage closures -------------------------------------------------------------------------------------------------
    def _pre(self, b0: int, nb: int):
        eng, net = self.eng, self.net
        s = torch.cuda.current_stream(eng.device).cuda_stream
        # single-label: the class-logit convs emit (max logit, class) lists instead of logits (plans are shared: set it per launch)
        net.set_cls_emit(self._emit)
        if self.kind == "f32":
            net.set_stem_source(None)
            cabi.check(eng._lib.y11_nchw_f32_to_nhwc_bf16_auto(eng._engine, self.frames.data_ptr(), self.B, self.H, self.W,
                                                               self.max_scratch.data_ptr(), net.input.data_ptr(), C.c_void_p(s)),
                       "y11_nchw_f32_to_nhwc_bf16_auto")
            return
        if self.fused_stem:      # (re-)point the plan's stem op(s) at this pipeline's frames; no letterbox launch
            net.set_stem_source(self.desc.data_ptr())
            return
        net.set_stem_source(None)
        cabi.check(eng._lib.y11_letterbox(eng._engine, self.desc.data_ptr() + b0 * self.desc_stride, nb, self.H, self.W,
                                          net.input[b0:b0 + nb].data_ptr(), C.c_void_p(s)), "y11_letterbox")

    def _post(self):
        self.det, self.count, self.ncand = self.eng.postprocess(self.net, self.scale_rows, self.conf, self.iou, self.max_det,
                                                                self.agnostic, self.multi_label, self.max_nms, self.out_flat,
                                                                self.push)

    # the enqueue functions: [chunk 0 prefix, ..., chunk K-1 prefix, rest]  (chunks == 1: a single stage)
    def _stages(self):
        eng, net = self.eng, self.net
        if self.chunks == 1:
            def whole():
                self._pre(0, self.B)
                eng.forward(net, self._emit)
                self._post()
            return [whole]
        Bc = self.B // self.chunks
        fns = []
        for c, (first, last) in enumerate(net.prefix_ranges):
            def prefix(c=c, first=first, last=last):
                self._pre(c * Bc, Bc)
                net.run_ops(first, last, torch.cuda.current_stream(eng.device).cuda_stream)
            fns.append(prefix)

        def rest():
            net.set_cls_emit(self._emit)
            net.run_ops(net.rest_first, net.n_ops, torch.cuda.current_stream(eng.device).cuda_stream)
            self._post()
        fns.append(rest)
        return fns

    def _measure_stage_shares(self):
        """One kernel-by-kernel pass with CUDA events between preprocess / network / decode+NMS: the shares by which
        `speed()` splits the measured time of a graph replay into the reference's `Results.speed` keys
        (/root/reference/core/validator.py:355-359 reads them)."""
        dev = self.eng.device
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record()
        self._pre(0, self.B)
        ev[1].record()
        self.eng.forward(self.net, self._emit)
        ev[2].record()
        self._post()
        ev[3].record()
        torch.cuda.synchronize(dev)
        t = [ev[i].elapsed_time(ev[i + 1]) for i in range(3)]
        tot = sum(t) or 1.0
        self.stage_share = {"preprocess": t[0] / tot, "inference": t[1] / tot, "postprocess": t[2] / tot}

    def speed(self, ms_per_image: float) -> Dict[str, float]:
        """ms per image of one call, split by the stage shares measured when the pipeline was built (with frames read
        directly by the stem there is no separate preprocess launch: its share is 0 by construction)."""
        return {k: ms_per_image * v for k, v in self.stage_share.items()}

    def _launch(self, i: int):
        if self.graphs:
            self.graphs[i].replay()
        else:
            self._stage_fns[i]()

    def run(self, src: Optional[torch.Tensor] = None):
        """Enqueue one pass on the current stream; returns the static (det [B,max_det,6], count [B], ncand [B]) tensors."""
        if self.chunks > 1 and src is not None and not src.is_cuda:
            cur = torch.cuda.current_stream(self.eng.device)
            cs = self.copy_stream
            cs.wait_stream(cur)                  # the previous pass has finished reading the static frame buffer
            Bc = self.B // self.chunks
            for c in range(self.chunks):
                with torch.cuda.stream(cs):
                    self.frames[c * Bc:(c + 1) * Bc].copy_(src[c * Bc:(c + 1) * Bc], non_blocking=True)
                    self.copy_events[c].record(cs)
                cur.wait_event(self.copy_events[c])
                self._launch(c)
            self._launch(self.chunks)
            return self.det, self.count, self.ncand
        if src is not None:
            self.frames.copy_(src, non_blocking=True)
        for i in range(len(self._stage_fns)):
            self._launch(i)
        return self.det, self.count, self.ncand
