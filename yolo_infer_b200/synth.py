"""Benchmark utility (NOT on the predict path): variance conditioning of random-init weights.

Random-init YOLO11 weights either collapse or explode through ~90 layers, which would make the decode / NMS stages of a
benchmark meaningless (zero or 8400 candidates per image).  `condition_synthetic_weights(engine)` rescales every conv so that
activations stay O(1) on random frames and a realistic number of anchors clears the confidence threshold; the factors are
written back into the engine's state_dict, so the CPU oracle can load the very same weights.  Used by bench.py and tools/.
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import torch

from . import topology as T
from .engine import DetectionNet, YOLO, letterbox_geometry
from .network import pack_weights


def condition_synthetic_weights(eng, hw: Tuple[int, int] = (640, 640), batch: int = 2, seed: int = 0, iters: int = 3,
                                act_rms: float = 1.0, box_std: float = 3.0, cls_std: float = 1.1,
                                cls_prior: float = 0.01) -> Dict[str, float]:
    """Rescale each conv (its BN gamma, or the plain conv weight) so activations stay O(1) on random frames.

    Random-init YOLO11 weights either collapse or explode through ~90 layers, which would make the decode/NMS stages
    of a benchmark meaningless (zero or 8400 candidates per image).  This walks the plan in execution order; for conv
    i it runs ops [0, i] on seeded uint8 frames, measures the rms (std for the two logit convs) of the op's output
    buffer and scales that conv's packed weights towards the target, `iters` times.  The per-conv factors are then
    written into the model's state_dict (so the oracle can load the very same weights) and everything is re-packed.
    Uses the engine's own kernels plus torch reductions at init time only - nothing here is on the predict path.
    """
    eng._ensure_device()
    H, W = hw
    with torch.cuda.device(eng.device), torch.inference_mode():
        net = eng.compiled(batch, H, W, fold_upsample=False)
        g = torch.Generator().manual_seed(seed)
        frames = torch.randint(0, 256, (batch, H, W, 3), generator=g, dtype=torch.uint8).to(eng.device)
        geoms = [letterbox_geometry(H, W, (H, W), False)] * batch
        eng.preprocess_images(net, list(frames), geoms)
        s = torch.cuda.current_stream(eng.device).cuda_stream
        factors: Dict[str, float] = {}
        bias_shift: Dict[str, torch.Tensor] = {}
        for i, op in enumerate(net.ops):
            in_place = op.name.endswith(("attn.proj", "ffn.1"))  # out aliases the residual: must run exactly once
            if op.kind not in ("conv", "dwconv", "stem") or op.name not in eng._packed or in_place:
                net.run_range(i, i + 1, s)
                continue
            pc = eng._packed[op.name]
            v = op.out
            is_logit = op.name.endswith((".cv2.0.2", ".cv2.1.2", ".cv2.2.2", ".cv3.0.2", ".cv3.1.2", ".cv3.2.2"))
            target = act_rms if not is_logit else (cls_std if ".cv3." in op.name else box_std)
            total = 1.0
            for _ in range(iters):
                net.run_range(i, i + 1, s)
                out = v.t[..., v.off:v.off + v.c].float()
                if is_logit:
                    cur = float((out - out.mean((0, 1, 2), keepdim=True)).std())
                else:  # residual adds are part of the signal the next layer sees, so they stay in
                    cur = float(out.pow(2).mean().sqrt())
                if not math.isfinite(cur) or cur <= 0:
                    break
                f = min(max(target / cur, 0.05), 20.0)
                pc.w.mul_(f)
                if not is_logit:
                    pc.b.mul_(f)
                total *= f
            net.run_range(i, i + 1, s)
            if is_logit and ".cv3." in op.name:
                # class logits: centre every channel on logit(prior) (random weights on positive-mean inputs give each
                # class its own offset of ~+-2, which would push most anchors over conf 0.25)
                out = v.t[..., v.off:v.off + eng.nc].float()
                shift = math.log(cls_prior / (1 - cls_prior)) - out.mean((0, 1, 2))
                pc.b[: eng.nc].add_(shift)
                bias_shift[op.name] = shift.cpu()
                net.run_range(i, i + 1, s)
            factors[op.name] = total
        torch.cuda.synchronize(eng.device)
    sd = eng.model.state_dict()
    for cp in T.conv_params(eng.scale, eng.nc):
        f = factors.get(cp.prefix, 1.0)
        if cp.bn:
            sd[f"{cp.prefix}.bn.weight"] = sd[f"{cp.prefix}.bn.weight"] * f
            sd[f"{cp.prefix}.bn.bias"] = sd[f"{cp.prefix}.bn.bias"] * f
        else:
            sd[f"{cp.prefix}.weight"] = sd[f"{cp.prefix}.weight"] * f
            if cp.prefix in bias_shift:
                sd[f"{cp.prefix}.bias"] = sd[f"{cp.prefix}.bias"] + bias_shift[cp.prefix]
    eng.model = DetectionNet(eng.scale, eng.nc, sd, eng.names)
    getattr(eng, "_pipes", {}).clear()
    eng._nets.clear()
    with torch.cuda.device(eng.device):
        eng._packed = pack_weights(eng.scale, eng.nc, sd, eng.device)
    return factors

