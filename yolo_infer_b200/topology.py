"""YOLO11-detect topology as data (no torch.nn): layer table, channel arithmetic, parameter inventory.

This is the host-side statement of the network the reference builds with ``ultralytics.YOLO(path)``
(/root/reference/core/model.py:110): yolo11.yaml + parse_model scaling rules.
It yields, in ultralytics module order, every convolution with its state_dict prefix (Appendix A.4) so that
weights exchange with ultralytics-style ``state_dict`` files by key name, and it is what
yolo_infer_b200/network.py walks to emit the op list for liby11_b200.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Iterator, List, Tuple

import torch

SCALES: Dict[str, Tuple[float, float, int]] = {
    "n": (0.50, 0.25, 1024), "s": (0.50, 0.50, 1024), "m": (0.50, 1.00, 512), "l": (1.00, 1.00, 512), "x": (1.00, 1.50, 512),
}
REG_MAX = 16
STRIDES = (8, 16, 32)

# (from, repeats, module, args)
LAYERS = (
    (-1, 1, "Conv", (64, 3, 2)), (-1, 1, "Conv", (128, 3, 2)), (-1, 2, "C3k2", (256, False, 0.25)),
    (-1, 1, "Conv", (256, 3, 2)), (-1, 2, "C3k2", (512, False, 0.25)), (-1, 1, "Conv", (512, 3, 2)),
    (-1, 2, "C3k2", (512, True)), (-1, 1, "Conv", (1024, 3, 2)), (-1, 2, "C3k2", (1024, True)),
    (-1, 1, "SPPF", (1024, 5)), (-1, 2, "C2PSA", (1024,)), (-1, 1, "Upsample", ()), ((-1, 6), 1, "Concat", ()),
    (-1, 2, "C3k2", (512, False)), (-1, 1, "Upsample", ()), ((-1, 4), 1, "Concat", ()), (-1, 2, "C3k2", (256, False)),
    (-1, 1, "Conv", (256, 3, 2)), ((-1, 13), 1, "Concat", ()), (-1, 2, "C3k2", (512, False)),
    (-1, 1, "Conv", (512, 3, 2)), ((-1, 10), 1, "Concat", ()), (-1, 2, "C3k2", (1024, True)),
    ((16, 19, 22), 1, "Detect", ()),
)


def make_divisible(x: float, d: int = 8) -> int:
    return int(math.ceil(x / d) * d)


@dataclass(frozen=True)
class LayerSpec:
    index: int
    kind: str            # Conv | C3k2 | SPPF | C2PSA | Upsample | Concat | Detect
    frm: tuple           # absolute source layer indices
    c1: int              # input channels (sum for Concat; tuple-sum irrelevant for Detect)
    c2: int              # output channels
    k: int = 1
    s: int = 1
    n: int = 1           # inner repeats (C3k2 / C2PSA)
    c3k: bool = False
    e: float = 0.5
    ch_in: tuple = ()    # Detect: per-level input channels


def layer_specs(scale: str) -> List[LayerSpec]:
    depth, width, max_ch = SCALES[scale]
    ch: List[int] = []
    out: List[LayerSpec] = []
    for i, (f, n, m, args) in enumerate(LAYERS):
        frm = tuple((i + x if x < 0 else x) for x in ((f,) if isinstance(f, int) else f))
        n = max(round(n * depth), 1) if n > 1 else n
        c1 = 3 if i == 0 else ch[frm[0]]
        if m in ("Conv", "C3k2", "SPPF", "C2PSA"):
            c2 = make_divisible(min(args[0], max_ch) * width, 8)
            if m == "Conv":
                spec = LayerSpec(i, m, frm, c1, c2, k=args[1], s=args[2])
            elif m == "C3k2":
                c3k = bool(args[1]) or scale in "mlx"
                spec = LayerSpec(i, m, frm, c1, c2, n=n, c3k=c3k, e=args[2] if len(args) > 2 else 0.5)
            elif m == "SPPF":
                spec = LayerSpec(i, m, frm, c1, c2, k=args[1])
            else:
                spec = LayerSpec(i, m, frm, c1, c2, n=n)
        elif m == "Upsample":
            spec = LayerSpec(i, m, frm, c1, c1)
        elif m == "Concat":
            c = sum(ch[x] for x in frm)
            spec = LayerSpec(i, m, frm, c, c)
        else:
            spec = LayerSpec(i, m, frm, 0, 0, ch_in=tuple(ch[x] for x in frm))
        out.append(spec)
        ch.append(spec.c2)
    return out


@dataclass(frozen=True)
class ConvParam:
    """One convolution of the network in ultralytics naming.  bn=True: `{p}.conv.weight` + `{p}.bn.*`;
    bn=False: plain nn.Conv2d `{p}.weight`, `{p}.bias`."""
    prefix: str
    c1: int
    c2: int
    k: int = 1
    s: int = 1
    g: int = 1
    act: bool = True
    bn: bool = True


def _bottleneck(p: str, c1: int, c2: int, e: float) -> Iterator[ConvParam]:
    c_ = int(c2 * e)
    yield ConvParam(f"{p}.cv1", c1, c_, 3)
    yield ConvParam(f"{p}.cv2", c_, c2, 3)


def _c3k(p: str, c1: int, c2: int, n: int = 2) -> Iterator[ConvParam]:
    c_ = int(c2 * 0.5)
    yield ConvParam(f"{p}.cv1", c1, c_, 1)
    yield ConvParam(f"{p}.cv2", c1, c_, 1)
    yield ConvParam(f"{p}.cv3", 2 * c_, c2, 1)
    for j in range(n):
        yield from _bottleneck(f"{p}.m.{j}", c_, c_, 1.0)


def detect_dims(ch_in: tuple, nc: int) -> Tuple[int, int]:
    return max(16, ch_in[0] // 4, REG_MAX * 4), max(ch_in[0], min(nc, 100))


def conv_params(scale: str, nc: int = 80) -> Iterator[ConvParam]:
    for sp in layer_specs(scale):
        p = f"model.{sp.index}"
        if sp.kind == "Conv":
            yield ConvParam(p, sp.c1, sp.c2, sp.k, sp.s)
        elif sp.kind == "C3k2":
            c = int(sp.c2 * sp.e)
            yield ConvParam(f"{p}.cv1", sp.c1, 2 * c, 1)
            yield ConvParam(f"{p}.cv2", (2 + sp.n) * c, sp.c2, 1)
            for j in range(sp.n):
                if sp.c3k:
                    yield from _c3k(f"{p}.m.{j}", c, c, 2)
                else:
                    yield from _bottleneck(f"{p}.m.{j}", c, c, 0.5)
        elif sp.kind == "SPPF":
            c_ = sp.c1 // 2
            yield ConvParam(f"{p}.cv1", sp.c1, c_, 1)
            yield ConvParam(f"{p}.cv2", 4 * c_, sp.c2, 1)
        elif sp.kind == "C2PSA":
            c = int(sp.c1 * 0.5)
            yield ConvParam(f"{p}.cv1", sp.c1, 2 * c, 1)
            yield ConvParam(f"{p}.cv2", 2 * c, sp.c1, 1)
            heads = c // 64
            kd = int((c // heads) * 0.5)
            for j in range(sp.n):
                yield ConvParam(f"{p}.m.{j}.attn.qkv", c, c + 2 * heads * kd, 1, act=False)
                yield ConvParam(f"{p}.m.{j}.attn.proj", c, c, 1, act=False)
                yield ConvParam(f"{p}.m.{j}.attn.pe", c, c, 3, g=c, act=False)
                yield ConvParam(f"{p}.m.{j}.ffn.0", c, 2 * c, 1)
                yield ConvParam(f"{p}.m.{j}.ffn.1", 2 * c, c, 1, act=False)
        elif sp.kind == "Detect":
            c2, c3 = detect_dims(sp.ch_in, nc)
            for l, x in enumerate(sp.ch_in):
                yield ConvParam(f"{p}.cv2.{l}.0", x, c2, 3)
                yield ConvParam(f"{p}.cv2.{l}.1", c2, c2, 3)
                yield ConvParam(f"{p}.cv2.{l}.2", c2, 4 * REG_MAX, 1, act=False, bn=False)
            for l, x in enumerate(sp.ch_in):
                yield ConvParam(f"{p}.cv3.{l}.0.0", x, x, 3, g=x)
                yield ConvParam(f"{p}.cv3.{l}.0.1", x, c3, 1)
                yield ConvParam(f"{p}.cv3.{l}.1.0", c3, c3, 3, g=c3)
                yield ConvParam(f"{p}.cv3.{l}.1.1", c3, c3, 1)
                yield ConvParam(f"{p}.cv3.{l}.2", c3, nc, 1, act=False, bn=False)


def param_shapes(scale: str, nc: int = 80) -> Dict[str, Tuple[int, ...]]:
    """name -> shape of every learnable tensor + BN buffers, ultralytics state_dict naming."""
    out: Dict[str, Tuple[int, ...]] = {}
    for cp in conv_params(scale, nc):
        wshape = (cp.c2, cp.c1 // cp.g, cp.k, cp.k)
        if cp.bn:
            out[f"{cp.prefix}.conv.weight"] = wshape
            for nm in ("weight", "bias", "running_mean", "running_var"):
                out[f"{cp.prefix}.bn.{nm}"] = (cp.c2,)
        else:
            out[f"{cp.prefix}.weight"] = wshape
            out[f"{cp.prefix}.bias"] = (cp.c2,)
    out["model.23.dfl.conv.weight"] = (1, REG_MAX, 1, 1)
    return out


LEARNABLE_SUFFIXES = (".conv.weight", ".bn.weight", ".bn.bias", ".weight", ".bias")


def is_learnable(name: str) -> bool:
    return not name.endswith(("running_mean", "running_var", "num_batches_tracked"))


def count_parameters(scale: str, nc: int = 80) -> int:
    """== sum(p.numel() for p in ultralytics_model.parameters()) (includes the frozen DFL conv)."""
    return sum(math.prod(s) for n, s in param_shapes(scale, nc).items() if is_learnable(n))


def default_state_dict(scale: str, nc: int = 80, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Random init in the style ultralytics applies to a model built from yolo11*.yaml:
    conv weights kaiming-uniform(a=sqrt 5) == U(+-1/sqrt(fan_in)), BN identity, Detect.bias_init."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    for cp in conv_params(scale, nc):
        fan_in = cp.c1 // cp.g * cp.k * cp.k
        bound = 1.0 / math.sqrt(fan_in)
        w = (torch.rand((cp.c2, cp.c1 // cp.g, cp.k, cp.k), generator=g) * 2 - 1) * bound
        if cp.bn:
            sd[f"{cp.prefix}.conv.weight"] = w
            sd[f"{cp.prefix}.bn.weight"] = torch.ones(cp.c2)
            sd[f"{cp.prefix}.bn.bias"] = torch.zeros(cp.c2)
            sd[f"{cp.prefix}.bn.running_mean"] = torch.zeros(cp.c2)
            sd[f"{cp.prefix}.bn.running_var"] = torch.ones(cp.c2)
        else:
            sd[f"{cp.prefix}.weight"] = w
            sd[f"{cp.prefix}.bias"] = (torch.rand(cp.c2, generator=g) * 2 - 1) * bound
    for l, s in enumerate(STRIDES):
        sd[f"model.23.cv2.{l}.2.bias"] = torch.ones(4 * REG_MAX)
        sd[f"model.23.cv3.{l}.2.bias"] = torch.full((nc,), math.log(5 / nc / (640 / s) ** 2))
    sd["model.23.dfl.conv.weight"] = torch.arange(REG_MAX, dtype=torch.float32).view(1, REG_MAX, 1, 1)
    return sd


def anchors_for(h: int, w: int) -> int:
    return sum((h // s) * (w // s) for s in STRIDES)


def synthetic_state_dict(scale: str, nc: int = 80, seed: int = 0, gain: float = 2.4, cls_prior: float = 0.01,
                         cls_gain: float = 1.6, box_gain: float = 3.0) -> Dict[str, torch.Tensor]:
    """Data-free, variance-preserving random weights for benchmarks (bench.py, smoke()).

    ultralytics-style default init collapses activations to ~1e-9 and produces zero detections, which
    would make the decode/NMS stages of a benchmark trivially cheap.  Here conv W ~ N(0, gain/fan_in) with `gain` chosen
    so the second moment survives SiLU, BN is the identity, the class bias is logit(cls_prior) and the last-layer gains
    make a few hundred to a few thousand of the 8400 anchors clear conf 0.25 with peaky DFL bins.
    """
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    for cp in conv_params(scale, nc):
        fan_in = cp.c1 // cp.g * cp.k * cp.k
        std = math.sqrt((gain if cp.act else 1.0) / fan_in)
        w = torch.randn((cp.c2, cp.c1 // cp.g, cp.k, cp.k), generator=g) * std
        if cp.bn:
            sd[f"{cp.prefix}.conv.weight"] = w
            sd[f"{cp.prefix}.bn.weight"] = torch.ones(cp.c2)
            sd[f"{cp.prefix}.bn.bias"] = torch.zeros(cp.c2)
            sd[f"{cp.prefix}.bn.running_mean"] = torch.zeros(cp.c2)
            sd[f"{cp.prefix}.bn.running_var"] = torch.ones(cp.c2)
        else:
            is_cls = ".cv3." in cp.prefix
            sd[f"{cp.prefix}.weight"] = w * (cls_gain if is_cls else box_gain)
            sd[f"{cp.prefix}.bias"] = (torch.full((cp.c2,), math.log(cls_prior / (1 - cls_prior))) if is_cls
                                       else torch.randn(cp.c2, generator=g) * 0.5)
    sd["model.23.dfl.conv.weight"] = torch.arange(REG_MAX, dtype=torch.float32).view(1, REG_MAX, 1, 1)
    return sd
