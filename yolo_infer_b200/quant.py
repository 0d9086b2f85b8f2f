"""FP8 post-training quantization behind the reference's quantizer surface.

The reference quantizes with `create_quantizer(kind, model, config).optimize(calibration_loader)` to CPU int8
(/root/reference/optimization/quantization/quantizers.py:24-310 PostTrainingQuantizer, :860-888 create_quantizer; base class
/root/reference/optimization/base.py:18-262).  int8 eager-mode CPU backends (qnnpack / fbgemm) have no meaning on a B200; the
Blackwell counterpart is e4m3 on the 5th-generation tensor cores: `create_quantizer("fp8", model).optimize(calibration_loader)`
returns the same `YOLO11Model` with its engine switched to FP8 mode -

  * which tensors: every conv -> conv edge whose intermediate tensor has exactly one reader (`network.fp8_pairs`): the hidden
    tensor of every Bottleneck and the two hidden tensors of each Detect box tower.  The producer's epilogue stores
    e4m3(value / s_act) instead of bf16 (half the bytes through HBM and L2), the consumer's weights are e4m3 with one scale
    per output channel (amax / 448), its MMA is tcgen05.mma.kind::f8f6f4 (K = 32 per instruction, fp32 accumulation in TMEM)
    and its epilogue multiplies the accumulator by s_act * s_w[n] before bias / SiLU / residual.  Everything else stays bf16.
  * calibration = static per-tensor activation scales s_act = amax / 448 over the calibration batches (PTQ as in the reference:
    a loader of inputs, no labels), measured by running the bf16 plan of the SAME engine.
  * parity: oracle/quant_ref.py applies the same quantisation points to the CPU restatement; tests/test_gpu_fp8.py.

Opt-in and separate from the headline numbers: bench.py reports it as `value_fp8`.
"""
from __future__ import annotations

import logging
from typing import Any, Dict, Iterable, List, Optional

import torch

from .network import E4M3_MAX, fp8_pairs

logger = logging.getLogger(__name__)


def calibrate_activation_scales(engine, batches: Iterable[torch.Tensor], margin: float = 1.0) -> Dict[str, float]:
    """batches: uint8 [B,H,W,3] BGR frame batches (host or device) or float [B,3,H,W] tensors in [0,1] / [0,255].  Runs the bf16
    plan of `engine` on each and returns {producer conv -> amax of its output / 448 * margin} for the fp8 edges."""
    from .engine import letterbox_geometry
    engine._ensure_device()
    saved = getattr(engine, "fp8_scales", None)
    if saved:
        engine.disable_fp8()
    prods = sorted({p for p, _, _ in fp8_pairs(engine.scale, engine.nc)})
    amax = {p: 0.0 for p in prods}
    with torch.cuda.device(engine.device), torch.inference_mode():
        for x in batches:
            if x.dtype == torch.uint8:
                B, h0, w0, _ = x.shape
                g = letterbox_geometry(h0, w0, (640, 640), True)
                net = engine.compiled(B, g[4], g[5])
                engine.preprocess_images(net, list(x.to(engine.device)), [g] * B)
            else:
                x = x.to(engine.device, torch.float32).contiguous()
                net = engine.compiled(x.shape[0], x.shape[2], x.shape[3])
                engine.preprocess_tensor(net, x, 255.0 if float(x.max()) > 1.0 + 1e-6 else 1.0)
            engine.forward(net)
            torch.cuda.synchronize(engine.device)
            for op in net.ops:
                if op.kind == "conv" and op.name in amax:
                    v = op.out
                    amax[op.name] = max(amax[op.name], float(v.t[..., v.off:v.off + v.c].float().abs().amax()))
    if saved:
        engine.enable_fp8(saved)
    return {p: max(a, 1e-6) / E4M3_MAX * margin for p, a in amax.items()}


class Fp8PostTrainingQuantizer:
    """Same surface as the reference's PostTrainingQuantizer (optimize / evaluate / optimization_metrics / optimization_history)."""

    def __init__(self, model: Any, config: Optional[Dict[str, Any]] = None, device: Optional[str] = None):
        self.original_model = model
        self.optimized_model = None
        self.config = config or {}
        self.device = device or "cuda"
        self.optimization_metrics: Dict[str, Any] = {}
        self.optimization_history: List[Dict[str, Any]] = []
        self.calibration_data = None
        self.num_calibration_batches = self.config.get("num_calibration_batches", 8)
        self.margin = float(self.config.get("margin", 1.0))

    def set_calibration_data(self, calibration_data: Any) -> None:
        self.calibration_data = calibration_data

    @staticmethod
    def _engine_of(model):
        return model.model if hasattr(model, "model") and hasattr(model.model, "enable_fp8") else model

    def optimize(self, calibration_loader: Any = None, **kwargs) -> Any:
        loader = calibration_loader if calibration_loader is not None else self.calibration_data
        if loader is None:
            raise ValueError("Calibration data is required for post-training quantization")   # quantizers.py:60-61
        eng = self._engine_of(self.original_model)
        batches = []
        for i, b in enumerate(loader):
            if i >= self.num_calibration_batches:
                break
            batches.append(b[0] if isinstance(b, (list, tuple)) else b)
        scales = calibrate_activation_scales(eng, batches, self.margin)
        eng.enable_fp8(scales)
        self.optimized_model = self.original_model
        self.optimization_metrics = {"quantization": "fp8_e4m3", "fp8_edges": len(scales), "calibration_batches": len(batches),
                                     "activation_scales": scales}
        rec = {"type": "fp8_post_training_quantization", "edges": len(scales)}
        self.optimization_history.append(rec)
        if hasattr(self.original_model, "optimization_history"):
            self.original_model.optimization_history.append(rec)
        return self.optimized_model

    def evaluate(self, test_data: Any, metrics: Optional[List[str]] = None) -> Dict[str, float]:
        """Head drift of the FP8 engine against the bf16 engine on `test_data` (a float [B,3,H,W] tensor): relative L2 error."""
        eng = self._engine_of(self.optimized_model or self.original_model)
        scales = getattr(eng, "fp8_scales", None)
        x = test_data.to(eng.device, torch.float32).contiguous()
        heads = []
        for mode in (None, scales):
            eng.disable_fp8() if mode is None else eng.enable_fp8(mode)
            with torch.cuda.device(eng.device), torch.inference_mode():
                net = eng.compiled(x.shape[0], x.shape[2], x.shape[3])
                eng.preprocess_tensor(net, x, 1.0)
                eng.forward(net)
                torch.cuda.synchronize(eng.device)
                heads.append(net.raw_head().clone())
        rel = float((heads[1] - heads[0]).norm() / heads[0].norm())
        return {"head_rel_l2_vs_bf16": rel}


def create_quantizer(quantization_type: str, model: Any, config: Optional[Dict[str, Any]] = None, **kwargs):
    """quantizers.py:860-888.  Supported on this path: 'fp8' (alias 'ptq': post-training, static scales)."""
    table = {"fp8": Fp8PostTrainingQuantizer, "ptq": Fp8PostTrainingQuantizer}
    if quantization_type not in table:
        raise ValueError(f"Unsupported quantization type: {quantization_type}. Supported types: {list(table.keys())}")
    return table[quantization_type](model, config, **kwargs)
