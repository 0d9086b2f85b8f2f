"""ORACLE (test infrastructure, never shipped, never on the product path).

CPU fp32 restatement of the YOLO11-detect network that the reference executes through
``ultralytics.YOLO`` (reference call sites: /root/reference/core/model.py:110 builds it,
core/model.py:133 runs ``predict``).  ultralytics itself is an un-vendored dependency
(/root/reference/requirements.txt:4, ``ultralytics>=8.0.0``, no lockfile) that is not
installed and not installable in this environment, so this file restates its published
architecture (upstream ``cfg/models/11/yolo11.yaml``, ``nn/tasks.py:parse_model``,
``nn/modules/{conv,block,head}.py``, ``utils/tal.py``, ``utils/torch_utils.py``) as
specified in SURVEY.md Appendix A.

PARITY PINNING: the reference holds no golden vectors.  This restatement is
pinned by known-answer tests only: exact published parameter counts and conv GFLOPs for the five
scales, anchor counts, and state_dict key names (tests/test_oracle_kat.py).  Numerical
parity against ultralytics itself is therefore "parity unpinned" (see DESIGN.md).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this module.
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

# scale -> (depth, width, max_channels)   [upstream yolo11.yaml `scales:`]
SCALES: Dict[str, Tuple[float, float, int]] = {
    "n": (0.50, 0.25, 1024),
    "s": (0.50, 0.50, 1024),
    "m": (0.50, 1.00, 512),
    "l": (1.00, 1.00, 512),
    "x": (1.00, 1.50, 512),
}

# [from, repeats, module, args]  [upstream yolo11.yaml backbone + head]
YAML_LAYERS = [
    [-1, 1, "Conv", [64, 3, 2]],
    [-1, 1, "Conv", [128, 3, 2]],
    [-1, 2, "C3k2", [256, False, 0.25]],
    [-1, 1, "Conv", [256, 3, 2]],
    [-1, 2, "C3k2", [512, False, 0.25]],
    [-1, 1, "Conv", [512, 3, 2]],
    [-1, 2, "C3k2", [512, True]],
    [-1, 1, "Conv", [1024, 3, 2]],
    [-1, 2, "C3k2", [1024, True]],
    [-1, 1, "SPPF", [1024, 5]],
    [-1, 2, "C2PSA", [1024]],
    [-1, 1, "Upsample", [None, 2, "nearest"]],
    [[-1, 6], 1, "Concat", [1]],
    [-1, 2, "C3k2", [512, False]],
    [-1, 1, "Upsample", [None, 2, "nearest"]],
    [[-1, 4], 1, "Concat", [1]],
    [-1, 2, "C3k2", [256, False]],
    [-1, 1, "Conv", [256, 3, 2]],
    [[-1, 13], 1, "Concat", [1]],
    [-1, 2, "C3k2", [512, False]],
    [-1, 1, "Conv", [512, 3, 2]],
    [[-1, 10], 1, "Concat", [1]],
    [-1, 2, "C3k2", [1024, True]],
    [[16, 19, 22], 1, "Detect", ["nc"]],
]


def make_divisible(x: float, divisor: int) -> int:
    return int(math.ceil(x / divisor) * divisor)


def autopad(k: int, p=None) -> int:
    return k // 2 if p is None else p


class Conv(nn.Module):
    """conv(no bias) -> BN(eps 1e-3) -> SiLU | identity.  [upstream nn/modules/conv.py:Conv]"""

    def __init__(self, c1, c2, k=1, s=1, p=None, g=1, act=True):
        super().__init__()
        self.conv = nn.Conv2d(c1, c2, k, s, autopad(k, p), groups=g, bias=False)
        self.bn = nn.BatchNorm2d(c2, eps=1e-3, momentum=0.03)
        self.act = nn.SiLU() if act is True else nn.Identity()

    def forward(self, x):
        if hasattr(self, "bn"):
            return self.act(self.bn(self.conv(x)))
        return self.act(self.conv(x))  # fused (forward_fuse upstream)


class DWConv(Conv):
    def __init__(self, c1, c2, k=1, s=1, act=True):
        super().__init__(c1, c2, k, s, g=math.gcd(c1, c2), act=act)


class Bottleneck(nn.Module):
    def __init__(self, c1, c2, shortcut=True, g=1, k=(3, 3), e=0.5):
        super().__init__()
        c_ = int(c2 * e)
        self.cv1 = Conv(c1, c_, k[0], 1)
        self.cv2 = Conv(c_, c2, k[1], 1, g=g)
        self.add = shortcut and c1 == c2

    def forward(self, x):
        return x + self.cv2(self.cv1(x)) if self.add else self.cv2(self.cv1(x))


class C3k(nn.Module):
    def __init__(self, c1, c2, n=1, shortcut=True, g=1, e=0.5, k=3):
        super().__init__()
        c_ = int(c2 * e)
        self.cv1 = Conv(c1, c_, 1, 1)
        self.cv2 = Conv(c1, c_, 1, 1)
        self.cv3 = Conv(2 * c_, c2, 1)
        self.m = nn.Sequential(*(Bottleneck(c_, c_, shortcut, g, k=(k, k), e=1.0) for _ in range(n)))

    def forward(self, x):
        return self.cv3(torch.cat((self.m(self.cv1(x)), self.cv2(x)), 1))


class C3k2(nn.Module):
    def __init__(self, c1, c2, n=1, c3k=False, e=0.5, g=1, shortcut=True):
        super().__init__()
        self.c = int(c2 * e)
        self.cv1 = Conv(c1, 2 * self.c, 1, 1)
        self.cv2 = Conv((2 + n) * self.c, c2, 1)
        self.m = nn.ModuleList(
            C3k(self.c, self.c, 2, shortcut, g) if c3k else Bottleneck(self.c, self.c, shortcut, g)
            for _ in range(n)
        )

    def forward(self, x):
        y = list(self.cv1(x).chunk(2, 1))
        y.extend(m(y[-1]) for m in self.m)
        return self.cv2(torch.cat(y, 1))


class SPPF(nn.Module):
    def __init__(self, c1, c2, k=5):
        super().__init__()
        c_ = c1 // 2
        self.cv1 = Conv(c1, c_, 1, 1)
        self.cv2 = Conv(c_ * 4, c2, 1, 1)
        self.m = nn.MaxPool2d(kernel_size=k, stride=1, padding=k // 2)

    def forward(self, x):
        y = [self.cv1(x)]
        y.extend(self.m(y[-1]) for _ in range(3))
        return self.cv2(torch.cat(y, 1))


class Attention(nn.Module):
    def __init__(self, dim, num_heads=8, attn_ratio=0.5):
        super().__init__()
        self.num_heads = num_heads
        self.head_dim = dim // num_heads
        self.key_dim = int(self.head_dim * attn_ratio)
        self.scale = self.key_dim ** -0.5
        nh_kd = self.key_dim * num_heads
        h = dim + nh_kd * 2
        self.qkv = Conv(dim, h, 1, act=False)
        self.proj = Conv(dim, dim, 1, act=False)
        self.pe = Conv(dim, dim, 3, 1, g=dim, act=False)

    def forward(self, x):
        B, C, H, W = x.shape
        N = H * W
        qkv = self.qkv(x)
        q, k, v = qkv.view(B, self.num_heads, self.key_dim * 2 + self.head_dim, N).split(
            [self.key_dim, self.key_dim, self.head_dim], dim=2
        )
        attn = (q.transpose(-2, -1) @ k) * self.scale
        attn = attn.softmax(dim=-1)
        x = (v @ attn.transpose(-2, -1)).view(B, C, H, W) + self.pe(v.reshape(B, C, H, W))
        return self.proj(x)


class PSABlock(nn.Module):
    def __init__(self, c, attn_ratio=0.5, num_heads=4, shortcut=True):
        super().__init__()
        self.attn = Attention(c, attn_ratio=attn_ratio, num_heads=num_heads)
        self.ffn = nn.Sequential(Conv(c, c * 2, 1), Conv(c * 2, c, 1, act=False))
        self.add = shortcut

    def forward(self, x):
        x = x + self.attn(x) if self.add else self.attn(x)
        x = x + self.ffn(x) if self.add else self.ffn(x)
        return x


class C2PSA(nn.Module):
    def __init__(self, c1, c2, n=1, e=0.5):
        super().__init__()
        assert c1 == c2
        self.c = int(c1 * e)
        self.cv1 = Conv(c1, 2 * self.c, 1, 1)
        self.cv2 = Conv(2 * self.c, c1, 1)
        self.m = nn.Sequential(*(PSABlock(self.c, attn_ratio=0.5, num_heads=self.c // 64) for _ in range(n)))

    def forward(self, x):
        a, b = self.cv1(x).split((self.c, self.c), dim=1)
        b = self.m(b)
        return self.cv2(torch.cat((a, b), 1))


class Concat(nn.Module):
    def __init__(self, dimension=1):
        super().__init__()
        self.d = dimension

    def forward(self, x):
        return torch.cat(x, self.d)


class DFL(nn.Module):
    def __init__(self, c1=16):
        super().__init__()
        self.conv = nn.Conv2d(c1, 1, 1, bias=False).requires_grad_(False)
        self.conv.weight.data[:] = torch.arange(c1, dtype=torch.float).view(1, c1, 1, 1)
        self.c1 = c1

    def forward(self, x):
        b, _, a = x.shape
        return self.conv(x.view(b, 4, self.c1, a).transpose(2, 1).softmax(1)).view(b, 4, a)


def make_anchors(feats, strides, grid_cell_offset=0.5):
    """[upstream utils/tal.py:make_anchors]"""
    anchor_points, stride_tensor = [], []
    dtype, device = feats[0].dtype, feats[0].device
    for i, stride in enumerate(strides):
        h, w = feats[i].shape[2:]
        sx = torch.arange(end=w, device=device, dtype=dtype) + grid_cell_offset
        sy = torch.arange(end=h, device=device, dtype=dtype) + grid_cell_offset
        sy, sx = torch.meshgrid(sy, sx, indexing="ij")
        anchor_points.append(torch.stack((sx, sy), -1).view(-1, 2))
        stride_tensor.append(torch.full((h * w, 1), stride, dtype=dtype, device=device))
    return torch.cat(anchor_points), torch.cat(stride_tensor)


def dist2bbox(distance, anchor_points, xywh=True, dim=-1):
    """[upstream utils/tal.py:dist2bbox]"""
    lt, rb = distance.chunk(2, dim)
    x1y1 = anchor_points - lt
    x2y2 = anchor_points + rb
    if xywh:
        c_xy = (x1y1 + x2y2) / 2
        wh = x2y2 - x1y1
        return torch.cat((c_xy, wh), dim)
    return torch.cat((x1y1, x2y2), dim)


class Detect(nn.Module):
    """Non-legacy (YOLO11) Detect head.  [upstream nn/modules/head.py:Detect]"""

    def __init__(self, nc=80, ch=()):
        super().__init__()
        self.nc = nc
        self.nl = len(ch)
        self.reg_max = 16
        self.no = nc + self.reg_max * 4
        self.stride = torch.tensor([8.0, 16.0, 32.0])
        c2, c3 = max((16, ch[0] // 4, self.reg_max * 4)), max(ch[0], min(self.nc, 100))
        self.cv2 = nn.ModuleList(
            nn.Sequential(Conv(x, c2, 3), Conv(c2, c2, 3), nn.Conv2d(c2, 4 * self.reg_max, 1)) for x in ch
        )
        self.cv3 = nn.ModuleList(
            nn.Sequential(
                nn.Sequential(DWConv(x, x, 3), Conv(x, c3, 1)),
                nn.Sequential(DWConv(c3, c3, 3), Conv(c3, c3, 1)),
                nn.Conv2d(c3, self.nc, 1),
            )
            for x in ch
        )
        self.dfl = DFL(self.reg_max)

    def forward_feats(self, x: List[torch.Tensor]) -> List[torch.Tensor]:
        return [torch.cat((self.cv2[i](x[i]), self.cv3[i](x[i])), 1) for i in range(self.nl)]

    def decode(self, feats: List[torch.Tensor]) -> torch.Tensor:
        shape = feats[0].shape
        x_cat = torch.cat([xi.view(shape[0], self.no, -1) for xi in feats], 2)
        anchors, strides = (t.transpose(0, 1) for t in make_anchors(feats, self.stride, 0.5))
        box, cls = x_cat.split((self.reg_max * 4, self.nc), 1)
        dbox = dist2bbox(self.dfl(box), anchors.unsqueeze(0), xywh=True, dim=1) * strides
        return torch.cat((dbox, cls.sigmoid()), 1)

    def forward(self, x: List[torch.Tensor]):
        feats = self.forward_feats(x)
        if self.training:
            return feats
        return self.decode(feats), feats

    def bias_init(self):
        for a, b, s in zip(self.cv2, self.cv3, self.stride):
            a[-1].bias.data[:] = 1.0
            b[-1].bias.data[: self.nc] = math.log(5 / self.nc / (640 / float(s)) ** 2)


class DetectionModel(nn.Module):
    """parse_model + _predict_once restated.  [upstream nn/tasks.py]"""

    def __init__(self, scale: str = "n", nc: int = 80):
        super().__init__()
        depth, width, max_ch = SCALES[scale]
        self.scale, self.nc = scale, nc
        ch = [3]
        layers, self.froms, self.save = [], [], set()
        for i, (f, n, m, args) in enumerate(YAML_LAYERS):
            args = list(args)
            n = max(round(n * depth), 1) if n > 1 else n
            if m in ("Conv", "C3k2", "SPPF", "C2PSA"):
                c1, c2 = ch[f], args[0]
                c2 = make_divisible(min(c2, max_ch) * width, 8)
                args = [c1, c2, *args[1:]]
                if m in ("C3k2", "C2PSA"):
                    args.insert(2, n)
                    n = 1
                if m == "C3k2" and scale in "mlx":
                    args[3] = True
                mod = {"Conv": Conv, "C3k2": C3k2, "SPPF": SPPF, "C2PSA": C2PSA}[m](*args)
            elif m == "Upsample":
                c2 = ch[f]
                mod = nn.Upsample(None, 2, "nearest")
            elif m == "Concat":
                c2 = sum(ch[x] for x in f)
                mod = Concat(1)
            elif m == "Detect":
                c2 = None
                mod = Detect(nc, [ch[x] for x in f])
            else:
                raise ValueError(m)
            layers.append(mod)
            self.froms.append(f)
            for x in ([f] if isinstance(f, int) else f):
                if x != -1:
                    self.save.add(x % i if x >= 0 else x)
            if i == 0:
                ch = []
            ch.append(c2)
        self.model = nn.ModuleList(layers)
        self.stride = torch.tensor([8.0, 16.0, 32.0])
        self.names = {i: f"{i}" for i in range(nc)}
        # initialize_weights (upstream utils/torch_utils.py): BN eps/momentum already set above
        self.model[-1].bias_init()

    def forward(self, x):
        y = []
        for f, m in zip(self.froms, self.model):
            if f != -1:
                x = y[f] if isinstance(f, int) else [x if j == -1 else y[j] for j in f]
            x = m(x)
            y.append(x)
        return x

    def fuse(self) -> "DetectionModel":
        """AutoBackend(fuse=True) -> fuse_conv_and_bn on every Conv.  [upstream utils/torch_utils.py]"""
        for m in self.modules():
            if isinstance(m, Conv) and hasattr(m, "bn"):
                w, b = fold_bn(m.conv.weight, m.bn.weight, m.bn.bias, m.bn.running_mean, m.bn.running_var, m.bn.eps)
                conv = nn.Conv2d(
                    m.conv.in_channels, m.conv.out_channels, m.conv.kernel_size, m.conv.stride, m.conv.padding,
                    groups=m.conv.groups, bias=True,
                ).requires_grad_(False)
                conv.weight.copy_(w)
                conv.bias.copy_(b)
                m.conv = conv
                delattr(m, "bn")
        return self


def fold_bn(w, gamma, beta, mean, var, eps=1e-3):
    """W' = W*gamma/sqrt(var+eps); b' = beta - mean*gamma/sqrt(var+eps)  (SURVEY §8 a5)."""
    scale = gamma / torch.sqrt(var + eps)
    return (w * scale.view(-1, 1, 1, 1)).detach(), (beta - mean * scale).detach()


def count_params(model: nn.Module) -> int:
    return sum(p.numel() for p in model.parameters())


def conv_flops(model: DetectionModel, h: int = 640, w: int = 640) -> int:
    """2*MAC over every nn.Conv2d for one h x w image."""
    total = 0
    hooks = []

    def hook(m, inp, out):
        nonlocal total
        cin_g = m.in_channels // m.groups
        total += 2 * out.shape[1] * out.shape[2] * out.shape[3] * cin_g * m.kernel_size[0] * m.kernel_size[1]

    for m in model.modules():
        if isinstance(m, nn.Conv2d) and not isinstance(m, DFL) and m.out_channels != 1:
            hooks.append(m.register_forward_hook(hook))
    was = model.training
    model.eval()
    with torch.no_grad():
        model(torch.zeros(1, 3, h, w))
    model.train(was)
    for hk in hooks:
        hk.remove()
    return total


@torch.no_grad()
def calibrated_init(model: DetectionModel, seed: int = 0, calib_hw: Tuple[int, int] = (640, 640),
                    cls_prior: float = 0.01, cls_gain: float = 1.0, box_gain: float = 3.0,
                    gamma: Tuple[float, float] = (0.2, 0.5), beta: Tuple[float, float] = (1.5, 0.3)) -> DetectionModel:
    """Deterministic, well-conditioned random weights shared by oracle and CUDA path.

    ultralytics' default init collapses activations (~1e-9 by P5) and yields zero detections, so parity tests use:
    conv W ~ N(0, 1/fan_in); BN gamma ~ U(gamma), beta ~ N(beta); BN running statistics := the statistics the layer
    actually sees on a seeded random batch (one train-mode pass with momentum 1), which keeps every activation O(1);
    Detect cls bias = logit(cls_prior); last-conv gains chosen so that hundreds of anchors clear conf 0.25 and DFL bins
    are peaky.

    Why gamma is small and beta positive: with gamma ~ U(.5,1.5), beta ~ 0 a random SiLU network is in the chaotic phase
    - a 2^-9 perturbation (one bf16 rounding) is amplified ~50x by the time it reaches the head, so an fp32 run and ANY
    16-bit run of the same network differ by 15-25 % (measured: bf16-emulating oracle vs fp32 oracle, box logits), which
    says nothing about kernel correctness.  gamma ~ U(.2,.5), beta ~ N(1.5,.3) keeps the perturbation growth per layer
    ~1, so the remaining fp32-vs-bf16 difference is the accumulated storage rounding of ~40 sequential layers (~1 %).
    """
    g = torch.Generator().manual_seed(seed)
    for m in model.modules():
        if isinstance(m, nn.Conv2d) and not (m.out_channels == 1 and m.in_channels == 16):
            fan_in = m.in_channels // m.groups * m.kernel_size[0] * m.kernel_size[1]
            m.weight.copy_(torch.randn(m.weight.shape, generator=g) * (1.0 / fan_in) ** 0.5)
        elif isinstance(m, nn.BatchNorm2d):
            m.weight.copy_(torch.rand(m.weight.shape, generator=g) * (gamma[1] - gamma[0]) + gamma[0])
            m.bias.copy_(torch.randn(m.bias.shape, generator=g) * beta[1] + beta[0])
    det: Detect = model.model[-1]
    for a, b in zip(det.cv2, det.cv3):
        a[-1].weight.mul_(box_gain)
        a[-1].bias.copy_(torch.randn(a[-1].bias.shape, generator=g) * 0.5)
        b[-1].weight.mul_(cls_gain)
        b[-1].bias.fill_(math.log(cls_prior / (1 - cls_prior)))
    # one train-mode pass: BN running stats := batch stats of a seeded image-like batch
    moms = {}
    for m in model.modules():
        if isinstance(m, nn.BatchNorm2d):
            moms[m] = m.momentum
            m.momentum = 1.0
    x = torch.rand((2, 3, *calib_hw), generator=g)
    model.train()
    model(x)
    model.eval()
    for m, mom in moms.items():
        m.momentum = mom
    return model


def emulate_bf16_storage(fused: DetectionModel, fold_upsample: bool = True) -> DetectionModel:
    """Turn a FUSED oracle model into the fp32-accumulate / bf16-storage network the B200 path computes, with a rounding to
    bf16 at EXACTLY the places where the CUDA path writes an activation tensor to HBM (and nowhere else):

      * conv weights rounded to bf16 once; the network input rounded to bf16;
      * every Conv output, EXCEPT a conv whose epilogue adds a residual before the store (Bottleneck.cv2, Attention.proj,
        PSABlock.ffn[1]): there the fp32 sum `act(conv) + x` is rounded once, as the fused epilogue does;
      * Attention: the softmax(QK^T)V product is written as a bf16 tensor (with bf16 probabilities feeding the PV product, fp32
        row sums - the flash kernel's operand format), the positional dwconv adds it in fp32 and stores one bf16 tensor;
      * Upsample -> Concat -> C3k2.cv1 (yaml layers 11-13, 14-16; `fold_upsample`): the low-resolution half `W_up . p + b` is
        stored in bf16 before it enters the skip half's accumulator (network.upsample_folds);
      * the Detect logits stay fp32 (the last 1x1 convs of the head write fp32 on the GPU).

    Used by tests to separate kernel arithmetic (must agree to <= 1e-2) from the storage-format noise floor that any bf16
    implementation has against an fp32 run.  What it cannot reproduce: fp32 summation order inside a conv, the MUFU
    tanh / exp2 approximations (2^-11 relative)."""
    import types

    def rq(t):
        return t.to(torch.bfloat16).float()

    for m in fused.modules():
        if isinstance(m, nn.Conv2d) and not (m.out_channels == 1 and m.in_channels == 16):
            m.weight.data = rq(m.weight.data)

    def conv_forward(self, x):
        y = self.act(self.conv(x))
        return y if getattr(self, "_keep_fp32", False) else rq(y)

    def bottleneck_forward(self, x):
        return rq(x + self.cv2(self.cv1(x))) if self.add else self.cv2(self.cv1(x))

    def attention_forward(self, x):
        B, C, H, W = x.shape
        N = H * W
        qkv = self.qkv(x)
        q, k, v = qkv.view(B, self.num_heads, self.key_dim * 2 + self.head_dim, N).split(
            [self.key_dim, self.key_dim, self.head_dim], dim=2)
        s = (q.transpose(-2, -1) @ k) * self.scale
        p = torch.exp(s - s.amax(-1, keepdim=True))
        o = rq((v @ rq(p).transpose(-2, -1)) / p.sum(-1).unsqueeze(-2)).view(B, C, H, W)   # bf16 P, fp32 row sums, bf16 store
        t = rq(o + self.pe(v.reshape(B, C, H, W)))                                         # dwconv epilogue: + residual, one store
        return self.proj(t)                                                                # fp32: PSABlock adds x, then stores

    def psa_forward(self, x):
        x = rq(x + self.attn(x)) if self.add else self.attn(x)
        x = rq(x + self.ffn(x)) if self.add else self.ffn(x)
        return x

    for m in fused.modules():
        if isinstance(m, Conv):
            m.forward = types.MethodType(conv_forward, m)
        if isinstance(m, Bottleneck):
            m.forward = types.MethodType(bottleneck_forward, m)
            m.cv2._keep_fp32 = bool(m.add)
        elif isinstance(m, Attention):
            m.forward = types.MethodType(attention_forward, m)
            m.pe._keep_fp32 = True
        elif isinstance(m, PSABlock):
            m.forward = types.MethodType(psa_forward, m)
            m.attn.proj._keep_fp32 = bool(m.add)
            m.ffn[1]._keep_fp32 = bool(m.add)
    if fold_upsample:
        for i, m in enumerate(fused.model):
            f = fused.froms[i]
            if not (isinstance(m, C3k2) and f == -1 and isinstance(fused.model[i - 1], Concat)):
                continue
            cat_from = fused.froms[i - 1]
            if not (isinstance(cat_from, list) and cat_from[0] == -1 and isinstance(fused.model[i - 2], nn.Upsample)):
                continue

            def folded_cv1(self, x):
                # x = cat(up2(p), skip); c_up = channels of p = all input channels minus the skip tensor's
                w, bias = self.conv.weight, self.conv.bias
                low = rq(torch.nn.functional.conv2d(x[:, :self._c_up], w[:, :self._c_up], bias))     # == up2(W_up . p + b), stored bf16
                return rq(self.act(torch.nn.functional.conv2d(x[:, self._c_up:], w[:, self._c_up:]) + low))

            skip_layer = cat_from[1]
            c_skip = _out_channels(fused.model[skip_layer])
            m.cv1._c_up = m.cv1.conv.in_channels - c_skip
            m.cv1.forward = types.MethodType(folded_cv1, m.cv1)
    fused.register_forward_pre_hook(lambda mod, inp: (rq(inp[0]),))
    return fused


def _out_channels(layer: nn.Module) -> int:
    """Output channels of a backbone/neck layer (Conv, C3k2, SPPF, C2PSA)."""
    if isinstance(layer, Conv):
        return layer.conv.out_channels
    return layer.cv2.conv.out_channels


def build(scale: str = "n", nc: int = 80, init: str = "calibrated", seed: int = 0) -> DetectionModel:
    torch.manual_seed(seed)
    model = DetectionModel(scale, nc)
    if init == "calibrated":
        calibrated_init(model, seed)
    elif init == "survey_b":
        # SURVEY.md section 8(d) init (B) as written: gamma ~ U(0.5, 1.5), beta ~ N(0, 0.1) (running statistics from a seeded batch as
        # above).  A random SiLU network with these BN parameters is in the chaotic phase (see calibrated_init's docstring):
        # kept as a parity case to SHOW the amplification, with the control experiment next to it (tests/test_gpu_e2e.py).
        calibrated_init(model, seed, gamma=(0.5, 1.5), beta=(0.0, 0.1))
    model.eval()
    return model
