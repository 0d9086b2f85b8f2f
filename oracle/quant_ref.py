"""ORACLE (test infrastructure, never shipped, never on the product path): the FP8 (e4m3) mode restated on the CPU.

The B200 path's FP8 mode (yolo_infer_b200/quant.py, the counterpart of the reference's post-training quantizers,
/root/reference/optimization/quantization/quantizers.py:24-310) stores the hidden tensor of every Bottleneck and of the Detect box
towers as e4m3 and runs their consumer convs with e4m3 weights.  `emulate_fp8` applies exactly those quantisation points to the
fused oracle network (on top of `yolo11_ref.emulate_bf16_storage`, which models where bf16 tensors are stored):

    hidden  : q_h = rn_e4m3(clamp(act(conv1(x)) / s_act, +-448))                 (from the fp32 value, as the GPU epilogue does)
    weights : s_w[n] = amax(|W[n]|) / 448 over the BN-folded fp32 weights, q_w = rn_e4m3(W / s_w)
    consumer: act(conv(q_h, q_w) * (s_act * s_w[n]) + b)                          (fp32 accumulation; products of e4m3 are exact in fp32)

torch's float8_e4m3fn conversion is round-to-nearest-even like the GPU's cvt.rn.satfinite.e4m3x2.f32; saturation is the clamp.
Which edges are quantised is decided HERE from the module tree (Bottleneck.cv1 -> cv2 with a hidden width that is a multiple of
32; Detect.cv2[l][0] -> [1] -> [2]), independently of the product's `network.fp8_pairs`; the test compares the two lists.
"""
from __future__ import annotations

import types
from typing import Dict, List, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import yolo11_ref as R

E4M3_MAX = 448.0


def q8(t: torch.Tensor, s) -> torch.Tensor:
    return (t / s).clamp(-E4M3_MAX, E4M3_MAX).to(torch.float8_e4m3fn).float()


def fp8_edges(model: R.DetectionModel) -> List[Tuple[str, str]]:
    out = []
    for name, m in model.named_modules():
        if isinstance(m, R.Bottleneck) and m.cv1.conv.out_channels % 32 == 0 and m.cv1.conv.kernel_size == (3, 3):
            out.append((f"{name}.cv1", f"{name}.cv2"))
    det = model.model[-1]
    for l, seq in enumerate(det.cv2):
        c = seq[0].conv.out_channels
        if c % 32 == 0:
            out.append((f"model.23.cv2.{l}.0", f"model.23.cv2.{l}.1"))
            out.append((f"model.23.cv2.{l}.1", f"model.23.cv2.{l}.2"))
    return out


def _quant_weights(conv: nn.Conv2d):
    w = conv.weight.detach().clone()
    s_w = (w.abs().amax(dim=(1, 2, 3)) / E4M3_MAX).clamp_min(1e-12)
    return q8(w, s_w.view(-1, 1, 1, 1)), s_w


def emulate_fp8(fused: R.DetectionModel, act_scales: Dict[str, float]) -> R.DetectionModel:
    """`fused`: a FUSED fp32 oracle model (not yet storage-emulated).  act_scales: {producer name -> s_act}."""
    edges = [(p, c) for p, c in fp8_edges(fused) if p in act_scales]
    mods = dict(fused.named_modules())
    qw = {}
    for _, c in edges:           # quantise from the fp32 folded weights, BEFORE the bf16 rounding of emulate_bf16_storage
        conv = mods[c].conv if isinstance(mods[c], R.Conv) else mods[c]
        qw[c] = _quant_weights(conv)
    R.emulate_bf16_storage(fused)

    def rq(t):
        return t.to(torch.bfloat16).float()

    def consumer(name, xq, s_in):
        m = mods[name]
        conv = m.conv if isinstance(m, R.Conv) else m
        wq, s_w = qw[name]
        y = F.conv2d(xq, wq, None, conv.stride, conv.padding) * (s_w * s_in).view(1, -1, 1, 1) + conv.bias.view(1, -1, 1, 1)
        return m.act(y) if isinstance(m, R.Conv) else y

    for p, c in edges:
        if p.endswith(".cv1"):
            bn = mods[p[:-4]]
            s = float(act_scales[p])
            bn.cv1._keep_fp32 = True

            def fwd(self, x, c=c, s=s):
                y = consumer(c, q8(self.cv1(x), s), s)
                return rq(x + y) if self.add else rq(y)
            bn.forward = types.MethodType(fwd, bn)
    det = fused.model[-1]
    for l, seq in enumerate(det.cv2):
        n0, n1, n2 = (f"model.23.cv2.{l}.{j}" for j in range(3))
        if n0 not in act_scales:
            continue
        s0, s1 = float(act_scales[n0]), float(act_scales.get(n1, 0.0))
        seq[0]._keep_fp32 = True

        class Tower(nn.Module):
            def __init__(self, seq, n1=n1, n2=n2, s0=s0, s1=s1):
                super().__init__()
                self.seq, self.n1, self.n2, self.s0, self.s1 = seq, n1, n2, s0, s1

            def forward(self, x):
                h1 = consumer(self.n1, q8(self.seq[0](x), self.s0), self.s0)
                if self.s1 > 0:
                    return consumer(self.n2, q8(h1, self.s1), self.s1)
                return self.seq[2](rq(h1))
        det.cv2[l] = Tower(seq)
    return fused


def calibrate(fused_emul: R.DetectionModel, xs: List[torch.Tensor]) -> Dict[str, float]:
    """amax / 448 of every fp8-edge producer's output on the bf16-storage oracle (the CPU statement of quant.calibrate_activation_scales)."""
    prods = sorted({p for p, _ in fp8_edges(fused_emul)})
    amax = {p: 0.0 for p in prods}
    mods = dict(fused_emul.named_modules())
    hooks = [mods[p].register_forward_hook(lambda m, i, o, p=p: amax.__setitem__(p, max(amax[p], float(o.abs().amax())))) for p in prods]
    with torch.no_grad():
        for x in xs:
            fused_emul(x)
    for h in hooks:
        h.remove()
    return {p: max(a, 1e-6) / E4M3_MAX for p, a in amax.items()}
