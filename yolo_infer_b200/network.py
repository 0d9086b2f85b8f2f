"""Host-side compiler: YOLO11 topology + state_dict -> BN-folded bf16 weights + an op list for liby11_b200.

Replaces what the reference gets from ultralytics ``AutoBackend(fuse=True)`` + ``DetectionModel._predict_once``
(SURVEY.md section 8a rows a5, a6): Conv+BN are folded offline (W' = W*g/sqrt(v+eps), b' = beta - mu*g/sqrt(v+eps)),
weights are packed K-major (kh, kw, cin) in bf16, and the 24-layer graph is flattened into one launch per fused op.
Concat / chunk / split never copy: producers write straight into channel slices of the consumer's buffer.

PyTorch is used for device memory only (buffers are torch tensors kept alive by the CompiledNet); every
computation is a kernel of liby11_b200.so reached through the C ABI in include/y11.h.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch

from . import _cabi as cabi
from . import topology as T

BN_EPS = 1e-3
S2D_STEM = True   # stem output in space-to-depth form, model.1 as a 2x2 stride-1 conv (see pack_weights)
# model.1 in the compact 2x2 form for 32- / 64-channel stems (s, m, l): permuted block order, only the blocks a tap touches are loaded
S2D_COMPACT = os.environ.get("Y11_S2D_COMPACT", "1") != "0"
# Upsample -> Concat -> C3k2.cv1 (yaml layers 11-13 and 14-16) without the upsampled / concatenated tensors: a 1x1 conv
# commutes with nearest upsampling, so cv1's weights are split by input channel, W_up . p (+bias) runs at LOW resolution and
# enters the conv over the skip tensor as a pre-activation term (Y11_RES_PRE_UP2).  Y11_FOLD_UP=0 restores the copy ops.
FOLD_UPSAMPLE = os.environ.get("Y11_FOLD_UP", "1") != "0"
# Plan autotuner: time every tcgen05 conv in each feasible launch variant at plan-build time and keep the fastest
# (y11_plan_autotune; variants are bit-identical in their results).  Y11_AUTOTUNE=0 keeps the built-in heuristics.
AUTOTUNE = os.environ.get("Y11_AUTOTUNE", "1") != "0"
# C3k.cv2 on a side lane (see CompiledNet._c3k): "auto" = scales m/l/x only.  Measured (batch 64, two steps in flight): YOLO11m
# 8.67 -> 8.98 k img/s, but YOLO11n 34.3 -> 33.6 k and YOLO11s 20.9 -> 20.8 k (their three C3k blocks sit on small maps where
# the extra fork/join edges cost more than the overlapped launch saves).
C3K_LANES = os.environ.get("Y11_C3K_LANES", "auto")
C3K_LANES_MAX_SMALL_B = int(os.environ.get("Y11_C3K_LANES_SMALL_B", "8"))
# C3k.cv1 and C3k.cv2 are two 1x1 convs over the SAME input: run them as ONE GEMM with concatenated output channels
# (weights [cv2 | cv1] stacked at pack time, one launch instead of two, the input tile read once).  Y11_C3K_MERGE=0 restores
# the two launches (and the side lane above).
C3K_MERGE = os.environ.get("Y11_C3K_MERGE", "1") != "0"   # "auto" also enables the lane for batches <= this (batch 1: n 0.505 -> 0.498 ms, s 0.691 -> 0.680 ms)
AUTOTUNE_REPS = int(os.environ.get("Y11_AUTOTUNE_REPS", "4"))
# Y11_TUNE_CACHE=<file.json>: tuned variants are stored per (scale, nc, B, H, W, chunks, fold) and re-applied on the next
# build instead of re-timing (a service restarts with the same plans; ncu sees the tuned plan without the tuning launches).
TUNE_CACHE = os.environ.get("Y11_TUNE_CACHE")
# Single-label pipelines: max-class reduction + conf pre-filter in the epilogue of the class-logit convs instead of an fp32
# [B,A,nc] logit tensor and a scan kernel (CompiledNet.set_cls_emit).  Y11_FUSE_CLS=0 keeps the stored-logits path everywhere.
FUSE_CLS_DECODE = os.environ.get("Y11_FUSE_CLS", "1") != "0"


def _tune_cache_load() -> Dict[str, dict]:
    import json
    try:
        with open(TUNE_CACHE) as f:
            return json.load(f)
    except (OSError, ValueError, TypeError):
        return {}


def _tune_cache_store(key: str, variants: dict) -> None:
    """Read-modify-write of the JSON file under an exclusive lock (torchrun ranks build their plans at the same time)."""
    import fcntl
    import json
    with open(f"{TUNE_CACHE}.lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            data = _tune_cache_load()
            data[key] = variants
            tmp = f"{TUNE_CACHE}.{os.getpid()}.tmp"
            with open(tmp, "w") as f:
                json.dump(data, f)
            os.replace(tmp, TUNE_CACHE)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def upsample_folds(scale: str) -> Dict[int, Tuple[int, int, int, int]]:
    """C3k2 layer index -> (low-res source layer, skip layer, upsample layer, concat layer) for every
    `Upsample(2) -> Concat([up, skip]) -> C3k2` chain whose intermediate tensors have no other consumer."""
    specs = T.layer_specs(scale)
    users: Dict[int, List[int]] = {}
    for sp in specs:
        for f in sp.frm:
            users.setdefault(f, []).append(sp.index)
    folds = {}
    for sp in specs:
        if sp.kind != "C3k2" or len(sp.frm) != 1:
            continue
        cat = specs[sp.frm[0]]
        if cat.kind != "Concat" or len(cat.frm) != 2 or users.get(cat.index) != [sp.index]:
            continue
        up = specs[cat.frm[0]]
        if up.kind != "Upsample" or users.get(up.index) != [cat.index]:
            continue
        folds[sp.index] = (up.frm[0], cat.frm[1], up.index, cat.index)
    return folds


def pad16(c: int) -> int:
    return (c + 15) // 16 * 16


@dataclass
class PackedConv:
    w: torch.Tensor          # bf16 [cout_p, k*k*cin_p]  (dense) | [9, c] (depthwise) | [cout, 27] (stem)
    b: torch.Tensor          # fp32 [cout_p]
    c1: int
    c2: int
    k: int
    s: int
    act: int
    depthwise: bool = False
    alg_k: int = 0           # algorithmic K per output (e.g. 9*cin of the 3x3 conv a repacked 2x2 space-to-depth conv stands for)
    in_fp8: bool = False     # w holds e4m3 bytes (uint8 tensor); the input tensor is e4m3 as well
    cscale: Optional[torch.Tensor] = None   # fp32 [cout]: input activation scale x per-channel weight scale (fp8 dequantisation)
    s2d_block: int = 0       # k == 2, compact form: space-to-depth input with blocks of this many channels in the permuted order


S2D_PERM = [(1, 0), (1, 1), (0, 1), (0, 0)]     # permuted space-to-depth block order of the compact 2x2 form: block index -> (dy, dx)


def compact_k2_stages(c: int) -> List[Tuple[int, int]]:
    """K stages (tap = ty*2+tx, first channel) of the compact 2x2 form for blocks of c channels (y11_conv_desc.s2d_block)."""
    lo, hi = [1, 0, 1, 0], [2, 2, 3, 4]
    return [(tap, lo[tap] * c + 64 * j) for tap in range(4) for j in range(((hi[tap] - lo[tap]) * c + 63) // 64)]


def dense_k2_weights(pc: "PackedConv") -> torch.Tensor:
    """[cout, 2, 2, 4c] fp32 weights of a k = 2 conv over its (possibly permuted-block) space-to-depth input, from either packing."""
    if not pc.s2d_block:
        return pc.w.float().view(pc.c2, 2, 2, pc.c1)
    w = torch.zeros(pc.c2, 4, pc.c1, device=pc.w.device)
    wc = pc.w.float().view(pc.c2, -1, 64)
    for k, (tap, c0) in enumerate(compact_k2_stages(pc.s2d_block)):
        w[:, tap, c0:c0 + 64] += wc[:, k]          # stages of one tap never overlap
    return w.view(pc.c2, 2, 2, pc.c1)


def fold(sd: Dict[str, torch.Tensor], cp: T.ConvParam) -> Tuple[torch.Tensor, torch.Tensor]:
    """fp32 folded (weight [c2, c1/g, k, k], bias [c2])."""
    if cp.bn:
        w = sd[f"{cp.prefix}.conv.weight"].float()
        scale = sd[f"{cp.prefix}.bn.weight"].float() / torch.sqrt(sd[f"{cp.prefix}.bn.running_var"].float() + BN_EPS)
        b = sd[f"{cp.prefix}.bn.bias"].float() - sd[f"{cp.prefix}.bn.running_mean"].float() * scale
        return w * scale.view(-1, 1, 1, 1), b
    return sd[f"{cp.prefix}.weight"].float(), sd[f"{cp.prefix}.bias"].float()


def qkv_permutation(c: int, heads: int, kd: int, hd: int) -> torch.Tensor:
    """Row order that turns ultralytics' per-head [q|k|v] qkv channels into [Q all heads | K all heads | V all heads]."""
    per = 2 * kd + hd
    q = [h * per + i for h in range(heads) for i in range(kd)]
    k = [h * per + kd + i for h in range(heads) for i in range(kd)]
    v = [h * per + 2 * kd + i for h in range(heads) for i in range(hd)]
    return torch.tensor(q + k + v, dtype=torch.long)


def pack_weights(scale: str, nc: int, sd: Dict[str, torch.Tensor], device) -> Dict[str, PackedConv]:
    packed: Dict[str, PackedConv] = {}
    for cp in T.conv_params(scale, nc):
        w, b = fold(sd, cp)
        if cp.prefix.endswith("attn.qkv"):
            c = cp.c1
            heads = c // 64
            hd = c // heads
            perm = qkv_permutation(c, heads, int(hd * 0.5), hd)
            w, b = w[perm], b[perm]
        act = cabi.ACT_SILU if cp.act else cabi.ACT_NONE
        if cp.g > 1:  # depthwise
            assert cp.g == cp.c1 == cp.c2 and cp.k == 3
            wp = w.view(cp.c2, 9).t().contiguous()  # [9, c] tap-major
            packed[cp.prefix] = PackedConv(wp.to(device, torch.bfloat16), b.to(device).contiguous(), cp.c1, cp.c2, 3, 1, act, True)
        elif cp.c1 == 3:  # stem
            wp = w.permute(0, 2, 3, 1).reshape(cp.c2, 27).contiguous()
            packed[cp.prefix] = PackedConv(wp.to(device, torch.bfloat16), b.to(device).contiguous(), 3, cp.c2, 3, 2, act)
        elif cp.prefix == "model.1" and S2D_STEM:
            # The stem writes its output in space-to-depth form ([H/2, W/2, 4*c] blocks (dy*2+dx)), on which this 3x3
            # stride-2 conv is a 2x2 stride-1 conv with taps at block offsets {-1, 0}^2: input row 2*(oy+by)+dy = 2*oy+kh-1
            # => kh = 2*by + dy + 1 (by in {-1, 0}); combinations outside 0..2 get zero weights.
            assert cp.k == 3 and cp.s == 2 and cp.c1 % 16 == 0 and cp.c2 % 16 == 0
            wp = torch.zeros(cp.c2, 2, 2, 4, cp.c1)                       # [cout][tap_y][tap_x][dy*2+dx][c]
            for ty in range(2):
                for tx in range(2):
                    for dy in range(2):
                        for dx in range(2):
                            kh, kw = 2 * (ty - 1) + dy + 1, 2 * (tx - 1) + dx + 1
                            if 0 <= kh <= 2 and 0 <= kw <= 2:
                                wp[:, ty, tx, dy * 2 + dx, :] = w[:, :, kh, kw]
            if S2D_COMPACT and cp.c1 in (32, 64):
                # Compact form (y11_conv_desc.s2d_block): the stem writes its blocks in the order [(1,0), (1,1), (0,1), (0,0)], in
                # which the blocks each tap can touch - 1, 2, 2 and 4 of the 4 - are contiguous channel ranges; only those are loaded.
                c = cp.c1
                stages = []
                for tap, c0 in compact_k2_stages(c):
                    ty, tx = tap // 2, tap % 2
                    st = torch.zeros(cp.c2, 64)
                    for i in range(64):
                        ch = c0 + i
                        dy, dx = S2D_PERM[ch // c]
                        kh, kw = 2 * (ty - 1) + dy + 1, 2 * (tx - 1) + dx + 1
                        if 0 <= kh <= 2 and 0 <= kw <= 2:
                            st[:, i] = w[:, ch % c, kh, kw]
                    stages.append(st)
                wc = torch.cat(stages, 1)
                packed[cp.prefix] = PackedConv(wc.to(device, torch.bfloat16).contiguous(), b.to(device).contiguous(), 4 * c, cp.c2, 2, 1,
                                               act, alg_k=9 * c, s2d_block=c)
                continue
            packed[cp.prefix] = PackedConv(wp.view(cp.c2, -1).to(device, torch.bfloat16).contiguous(), b.to(device).contiguous(),
                                           4 * cp.c1, cp.c2, 2, 1, act, alg_k=9 * cp.c1)
        else:
            c1p, c2p = pad16(cp.c1), pad16(cp.c2)
            wp = torch.zeros(c2p, cp.k, cp.k, c1p)
            wp[: cp.c2, :, :, : cp.c1] = w.permute(0, 2, 3, 1)
            bp = torch.zeros(c2p)
            bp[: cp.c2] = b
            packed[cp.prefix] = PackedConv(wp.view(c2p, -1).to(device, torch.bfloat16).contiguous(), bp.to(device), c1p, c2p,
                                           cp.k, cp.s, act)
    specs = T.layer_specs(scale)
    for sp in specs:
        if sp.kind == "C3k2" and sp.c3k:
            for j in range(sp.n):
                pre = f"model.{sp.index}.m.{j}"
                a, b = packed[f"{pre}.cv2"], packed[f"{pre}.cv1"]      # output order [cv2 | cv1] (see CompiledNet._c3k)
                if a.c1 == b.c1 and a.act == b.act and a.k == b.k == 1:
                    packed[f"{pre}.cv12"] = PackedConv(torch.cat((a.w, b.w), 0).contiguous(), torch.cat((a.b, b.b), 0).contiguous(),
                                                       a.c1, a.c2 + b.c2, 1, 1, a.act)
    for idx, (low, skip, _up, _cat) in upsample_folds(scale).items():
        # cv1 over Concat([up2(p), skip]) = up2(W[:, :c_up] . p + b)  +  W[:, c_up:] . skip      (then SiLU)
        full = packed[f"model.{idx}.cv1"]
        c_up, c_skip = specs[low].c2, specs[skip].c2
        if full.k != 1 or c_up % 16 or c_skip % 16 or full.c1 != c_up + c_skip:
            continue
        packed[f"model.{idx}.cv1#up"] = PackedConv(full.w[:, :c_up].contiguous(), full.b.clone(), c_up, full.c2, 1, 1, cabi.ACT_NONE)
        packed[f"model.{idx}.cv1#skip"] = PackedConv(full.w[:, c_up:].contiguous(), torch.zeros_like(full.b), c_skip, full.c2, 1, 1,
                                                     full.act)
    return packed


E4M3_MAX = 448.0


def fp8_pairs(scale: str, nc: int = 80) -> List[Tuple[str, str, bool]]:
    """(producer, consumer, consumer_writes_fp8) for every conv -> conv edge whose intermediate tensor has NO other reader, so that it
    can live in HBM as e4m3: the hidden tensor of every Bottleneck (cv1 -> cv2) and the two hidden tensors of the Detect box
    tower (cv2.l.0 -> cv2.l.1 -> cv2.l.2).  Edges whose tensor has fewer than 32 channels (or not a multiple of 32) are left in
    bf16 (tcgen05.mma.kind::f8f6f4 consumes 32 channels per instruction)."""
    cps = {c.prefix: c for c in T.conv_params(scale, nc)}
    out = []
    for name, c in cps.items():
        if name.endswith(".cv1") and c.k == 3 and c.g == 1 and ".m." in name:
            cons = name[:-1] + "2"
            if cons in cps and cps[cons].k == 3 and c.c2 % 32 == 0:
                out.append((name, cons, False))
    for l in range(3):
        a, b, c = (f"model.23.cv2.{l}.{j}" for j in range(3))
        if cps[a].c2 % 32 == 0:
            out.append((a, b, cps[b].c2 % 32 == 0))
            if cps[b].c2 % 32 == 0:
                out.append((b, c, False))
    return out


def pack_fp8(scale: str, nc: int, sd: Dict[str, torch.Tensor], packed: Dict[str, PackedConv], act_scales: Dict[str, float], device) -> None:
    """Adds `<consumer>#fp8` entries: BN-folded weights quantised to e4m3 with one scale per output channel (amax / 448),
    K-major like the bf16 packing; cscale[n] = activation scale of the input tensor x weight scale of channel n."""
    cps = {c.prefix: c for c in T.conv_params(scale, nc)}
    for prod, cons, _ in fp8_pairs(scale, nc):
        if prod not in act_scales:
            continue
        cp = cps[cons]
        w, b = fold(sd, cp)
        s_w = (w.abs().amax(dim=(1, 2, 3)) / E4M3_MAX).clamp_min(1e-12)
        q = (w / s_w.view(-1, 1, 1, 1)).clamp(-E4M3_MAX, E4M3_MAX).to(torch.float8_e4m3fn)
        c2p = pad16(cp.c2)
        wq = torch.zeros((c2p, cp.k, cp.k, cp.c1), dtype=torch.uint8)
        wq[: cp.c2] = q.permute(0, 2, 3, 1).contiguous().view(torch.uint8)
        bp, cs = torch.zeros(c2p), torch.zeros(c2p)
        bp[: cp.c2] = b
        cs[: cp.c2] = s_w * float(act_scales[prod])
        packed[f"{cons}#fp8"] = PackedConv(wq.view(c2p, -1).to(device).contiguous(), bp.to(device), cp.c1, c2p, cp.k, cp.s,
                                           cabi.ACT_SILU if cp.act else cabi.ACT_NONE, in_fp8=True, cscale=cs.to(device))


@dataclass
class V:
    """A channel-slice view of an NHWC buffer."""
    t: torch.Tensor
    off: int
    c: int

    @property
    def H(self):
        return self.t.shape[1]

    @property
    def W(self):
        return self.t.shape[2]

    def sub(self, off: int, c: int) -> "V":
        assert off + c <= self.c
        return V(self.t, self.off + off, c)

    @property
    def B(self):
        return self.t.shape[0]

    def bslice(self, b0: int, nb: int) -> "V":
        """The same channel view over images [b0, b0+nb) (batch is the outermost dimension: a pointer offset)."""
        return V(self.t[b0:b0 + nb], self.off, self.c)

    def cview(self) -> cabi.View:
        return cabi.View(self.t.data_ptr(), self.t.shape[-1], self.off, self.c)


@dataclass
class OpRecord:
    kind: str
    name: str
    flops: float = 0.0
    bytes_algo: float = 0.0   # algorithmic HBM bytes: unique input + weights + output
    out: object = None  # where the op writes (used by weight conditioning and the per-layer parity test)
    inp: object = None
    res: object = None
    res_mode: int = 0


class CompiledNet:
    """One (scale, B, H, W) instance: owns activation buffers, the plan handle and the head tensors."""

    PREFIX_LAYERS = 5   # yaml layers 0-4 (stem .. first stride-8 C3k2): the high-resolution, memory-heavy part

    def __init__(self, engine, scale: str, nc: int, packed: Dict[str, PackedConv], B: int, H: int, W: int, device,
                 conv_impl: int = cabi.IMPL_TCGEN05, chunks: int = 1, fold_upsample: Optional[bool] = None,
                 fp8: Optional[Dict[str, float]] = None):
        """chunks > 1: layers 0-4 are emitted once per batch chunk (chunk-major), so that a host-fed pipeline can run chunk c
        while chunk c+1 is still on the PCIe bus (engine.GraphedPipeline); the rest of the network runs on the whole batch."""
        assert H % 32 == 0 and W % 32 == 0, "network input must be a multiple of 32"
        assert chunks >= 1 and B % chunks == 0, (B, chunks)
        self.chunks = chunks
        # fp8: {producer conv -> activation scale of its output}: those tensors are stored as e4m3 and their consumers run on
        # tcgen05.mma.kind::f8f6f4 (see fp8_pairs / quant.py); None = the bf16 network
        self.fp8 = dict(fp8) if (fp8 and conv_impl == cabi.IMPL_TCGEN05 and fold_upsample is not False) else {}
        self.fp8_out = {p: c_out for p, _, c_out in fp8_pairs(scale, nc)} if self.fp8 else {}
        # fold_upsample=False asks for ONE OP PER REFERENCE CONV (the weight conditioning walks the plan by conv name)
        self.merge_c3k = C3K_MERGE and fold_upsample is not False
        self.folds = upsample_folds(scale) if (FOLD_UPSAMPLE if fold_upsample is None else fold_upsample) else {}
        self.folds = {k: v for k, v in self.folds.items() if f"model.{k}.cv1#up" in packed}
        self._cur_B = B
        self.prefix_ranges: List[Tuple[int, int]] = []
        self.lib = cabi.load()
        self.engine = engine
        self.scale, self.nc, self.B, self.H, self.W, self.device = scale, nc, B, H, W, device
        self.packed = packed
        self.conv_impl = conv_impl
        self.buffers: List[torch.Tensor] = []
        self.ops: List[OpRecord] = []
        self.plan = C.c_void_p()
        cabi.check(self.lib.y11_plan_create(engine, C.byref(self.plan)), "y11_plan_create")
        self.input = self._alloc(H, W, 3)                      # bf16 NHWC, written by the letterbox kernel
        self.no = 64 + pad16(nc)
        self.head: List[torch.Tensor] = []
        self.cls_ops: List[Tuple[int, int]] = []          # (op index, anchor offset) of the class-logit convs cv3.l.2
        self.emit_list: Optional[torch.Tensor] = None     # class-emit mode (set_cls_emit): int32 [B, A, 4] pre-candidate lists
        self.emit_count: Optional[torch.Tensor] = None
        self.emit_conf: Optional[float] = None
        tune = AUTOTUNE and conv_impl == cabi.IMPL_TCGEN05
        props = torch.cuda.get_device_properties(device)
        self._tune_key = (f"{props.name}/sm{props.multi_processor_count}/{scale}/nc{nc}/B{B}/{H}x{W}/chunks{chunks}/"
                          f"fold{int(bool(self.folds))}/merge{int(self.merge_c3k)}/lanes{C3K_LANES}/impl{conv_impl}/fp8{len(self.fp8)}")
        cached = _tune_cache_load().get(self._tune_key) if (tune and TUNE_CACHE) else None
        # stored BY OP NAME (a chunk-major plan has one op of a name per chunk: same variant); a plan whose op names differ
        # from the stored ones is re-tuned
        self._cached_variants = cached if isinstance(cached, dict) else None
        self._build()
        self.A = sum(h.shape[1] * h.shape[2] for h in self.head)
        if tune and self._cached_variants is not None and not all(o.name in self._cached_variants for o in self.ops if o.kind == "conv"):
            self._cached_variants = None     # stale entry (different op list): tune again
        if tune and self._cached_variants is None:
            s = torch.cuda.current_stream(device).cuda_stream
            cabi.check(self.lib.y11_plan_autotune(self.plan, C.c_void_p(s), AUTOTUNE_REPS), "y11_plan_autotune")
            for t in self.buffers:      # the timing runs left garbage (in-place residual ops accumulate): start from zeros again
                t.zero_()
            if TUNE_CACHE:
                _tune_cache_store(self._tune_key, {o.name: list(v) for o, v in zip(self.ops, self.variants()) if o.kind == "conv"})

    def __del__(self):
        try:
            if self.plan:
                self.lib.y11_plan_destroy(self.plan)
                self.plan = C.c_void_p()
        except Exception:
            pass

    # ---- buffers ------------------------------------------------------------------------------
    def _alloc(self, h: int, w: int, c: int, dtype=torch.bfloat16) -> torch.Tensor:
        t = torch.zeros((self._cur_B, h, w, c), dtype=dtype, device=self.device)
        self.buffers.append(t)
        return t

    def _new(self, h: int, w: int, c: int) -> V:
        cp = pad16(c)
        return V(self._alloc(h, w, cp), 0, cp)

    def _new8(self, h: int, w: int, c: int) -> V:
        """An e4m3 tensor (one byte per channel; c % 32 == 0)."""
        assert c % 32 == 0
        return V(self._alloc(h, w, c, torch.uint8), 0, c)

    def _fp8_edge(self, prod: str, cons: str) -> bool:
        return prod in self.fp8 and f"{cons}#fp8" in self.packed

    # ---- op emitters --------------------------------------------------------------------------
    def _conv(self, name: str, x: V, out: V, res: Optional[V] = None, out_f32: bool = False, res_mode: int = cabi.RES_POST,
              out_fp8_scale: Optional[float] = None):
        """out_fp8_scale: store the result as e4m3(value / out_fp8_scale) (out is a uint8 buffer)."""
        pc = self.packed[name]
        assert not pc.depthwise
        assert x.c == pc.c1, (name, x.c, pc.c1)
        assert out.c == pc.c2, (name, out.c, pc.c2)
        assert pc.in_fp8 == (x.t.dtype == torch.uint8) and (out_fp8_scale is not None) == (out.t.dtype == torch.uint8), name
        d = cabi.ConvDesc()
        d.inp, d.out = x.cview(), out.cview()
        d.res = res.cview() if res is not None else cabi.NULL_VIEW
        d.w, d.bias = pc.w.data_ptr(), pc.b.data_ptr()
        assert x.B == out.B
        d.B, d.Hin, d.Win, d.Hout, d.Wout = x.B, x.H, x.W, out.H, out.W
        d.k, d.stride, d.act, d.out_f32, d.impl = pc.k, pc.s, pc.act, int(out_f32), self.conv_impl
        d.res_mode = res_mode
        d.in_fp8 = int(pc.in_fp8)
        d.cscale = pc.cscale.data_ptr() if pc.cscale is not None else None
        d.out_fp8 = int(out_fp8_scale is not None)
        d.out_scale = 1.0 / out_fp8_scale if out_fp8_scale is not None else 1.0
        d.s2d_block = pc.s2d_block
        var = self._cached_variants.get(name) if self._cached_variants is not None else None
        if var is not None and var[2] > 0 and self.conv_impl == cabi.IMPL_TCGEN05:   # variant chosen by an earlier autotune run
            cabi.check(self.lib.y11_plan_add_conv_tuned(self.plan, C.byref(d), *[int(v) for v in var]), f"plan_add_conv_tuned({name})")
        else:
            cabi.check(self.lib.y11_plan_add_conv(self.plan, C.byref(d)), f"plan_add_conv({name})")
        px = x.B * out.H * out.W
        res_bytes = 0 if res is None else res.B * res.H * res.W * pc.c2 * 2
        self.ops.append(OpRecord("conv", name, 2.0 * px * pc.c2 * (pc.alg_k or pc.c1 * pc.k * pc.k),
                                 x.B * x.H * x.W * pc.c1 * x.t.element_size() + pc.w.numel() * pc.w.element_size()
                                 + px * pc.c2 * out.t.element_size() + res_bytes, out, x, res, res_mode))

    def _dw(self, name: str, x: V, out: V, res: Optional[V] = None):
        pc = self.packed[name]
        assert pc.depthwise and x.c == pc.c1 == out.c
        d = cabi.DwConvDesc()
        d.inp, d.out = x.cview(), out.cview()
        d.res = res.cview() if res is not None else cabi.NULL_VIEW
        d.w, d.bias = pc.w.data_ptr(), pc.b.data_ptr()
        d.B, d.H, d.W, d.act = x.B, x.H, x.W, pc.act
        cabi.check(self.lib.y11_plan_add_dwconv(self.plan, C.byref(d)), f"plan_add_dwconv({name})")
        px = x.B * x.H * x.W
        self.ops.append(OpRecord("dwconv", name, 2.0 * px * pc.c1 * 9, px * pc.c1 * 2 * (3 if res is not None else 2), out, x, res))

    # ---- modules ------------------------------------------------------------------------------
    def _bottleneck(self, p: str, x: V, out: V, e: float):
        if self._fp8_edge(f"{p}.cv1", f"{p}.cv2"):
            # the hidden tensor has one writer and one reader: it lives in HBM as e4m3 and cv2 runs on the fp8 tensor-core path
            hidden = self._new8(x.H, x.W, int(out.c * e))
            self._conv(f"{p}.cv1", x, hidden, out_fp8_scale=self.fp8[f"{p}.cv1"])
            self._conv(f"{p}.cv2#fp8", hidden, out, res=x)
            return
        hidden = self._new(x.H, x.W, int(out.c * e))
        self._conv(f"{p}.cv1", x, hidden)
        self._conv(f"{p}.cv2", hidden, out, res=x)  # shortcut: c1 == c2 everywhere in YOLO11

    C3K_SIDE_LANE = 7   # lanes 1-6 belong to the Detect towers

    def _c3k(self, p: str, x: V, out: V):
        """C3k = cv3(cat(m(cv1(x)), cv2(x))).  cv2 depends only on x, so it runs on a side lane (a parallel branch of the CUDA
        graph) next to the cv1 -> Bottleneck x2 chain: on the 20x20 / 40x40 maps every launch is latency bound and one of
        the block's seven is taken off the critical path."""
        c_ = int(out.c * 0.5)
        if self.merge_c3k and f"{p}.cv12" in self.packed and c_ % 16 == 0:
            # z = [m(cv1 x) | cv2 x | cv1 x]: ONE conv writes channels [c_, 3c_), cv3 reads [0, 2c_), the Bottleneck chain
            # starts from the [2c_, 3c_) slice
            z = self._new(x.H, x.W, 3 * c_)
            self._conv(f"{p}.cv12", x, z.sub(c_, 2 * c_))
            t1 = self._new(x.H, x.W, c_)
            self._bottleneck(f"{p}.m.0", z.sub(2 * c_, c_), t1, 1.0)
            self._bottleneck(f"{p}.m.1", t1, z.sub(0, c_), 1.0)
            self._conv(f"{p}.cv3", z.sub(0, 2 * c_), out)
            return
        z = self._new(x.H, x.W, 2 * c_)
        side = C3K_LANES == "1" or (C3K_LANES == "auto" and (self.scale in "mlx" or self.B <= C3K_LANES_MAX_SMALL_B))
        if side:
            cabi.check(self.lib.y11_plan_fork(self.plan, self.C3K_SIDE_LANE), "plan_fork")
            cabi.check(self.lib.y11_plan_set_lane(self.plan, self.C3K_SIDE_LANE), "plan_set_lane")
            self._conv(f"{p}.cv2", x, z.sub(c_, c_))
            cabi.check(self.lib.y11_plan_set_lane(self.plan, 0), "plan_set_lane")
        t0 = self._new(x.H, x.W, c_)
        self._conv(f"{p}.cv1", x, t0)
        t1 = self._new(x.H, x.W, c_)
        self._bottleneck(f"{p}.m.0", t0, t1, 1.0)
        self._bottleneck(f"{p}.m.1", t1, z.sub(0, c_), 1.0)
        if side:
            cabi.check(self.lib.y11_plan_join(self.plan, self.C3K_SIDE_LANE), "plan_join")
        else:
            self._conv(f"{p}.cv2", x, z.sub(c_, c_))
        self._conv(f"{p}.cv3", z, out)

    def _c3k2(self, sp: T.LayerSpec, x: V, out: V, low: Optional[V] = None):
        """low != None: x is only the skip half of a folded Upsample+Concat input and `low` the low-resolution half."""
        p = f"model.{sp.index}"
        c = int(sp.c2 * sp.e)
        y = self._new(x.H, x.W, (2 + sp.n) * c)
        if low is not None:
            pre = self._new(low.H, low.W, 2 * c)                   # W_up . low + b at low resolution (bf16, no activation)
            self._conv(f"{p}.cv1#up", low, pre)
            self.ops[-1].flops *= 4.0     # ALGORITHMIC FLOPs: the reference runs these MACs on the 4x larger upsampled map
            self._conv(f"{p}.cv1#skip", x, y.sub(0, 2 * c), res=pre, res_mode=cabi.RES_PRE_UP2)
        else:
            self._conv(f"{p}.cv1", x, y.sub(0, 2 * c))
        for j in range(sp.n):
            src, dst = y.sub((1 + j) * c, c), y.sub((2 + j) * c, c)
            if sp.c3k:
                self._c3k(f"{p}.m.{j}", src, dst)
            else:
                self._bottleneck(f"{p}.m.{j}", src, dst, 0.5)
        self._conv(f"{p}.cv2", y, out)

    def _sppf(self, sp: T.LayerSpec, x: V, out: V):
        p = f"model.{sp.index}"
        c_ = sp.c1 // 2
        s = self._new(x.H, x.W, 4 * c_)
        self._conv(f"{p}.cv1", x, s.sub(0, c_))
        d = cabi.SppfDesc(s.cview(), self.B, x.H, x.W, c_)
        cabi.check(self.lib.y11_plan_add_sppf(self.plan, C.byref(d)), "plan_add_sppf")
        self.ops.append(OpRecord("sppf", p + ".pool", 0.0, self.B * x.H * x.W * c_ * 2 * 4))
        self._conv(f"{p}.cv2", s, out)

    def _c2psa(self, sp: T.LayerSpec, x: V, out: V):
        p = f"model.{sp.index}"
        c = int(sp.c1 * 0.5)
        heads = c // 64
        hd = c // heads
        kd = int(hd * 0.5)
        pbuf = self._new(x.H, x.W, 2 * c)
        self._conv(f"{p}.cv1", x, pbuf)
        b = pbuf.sub(c, c)
        n_tok = x.H * x.W
        for j in range(sp.n):
            q = self._new(x.H, x.W, c + 2 * heads * kd)
            self._conv(f"{p}.m.{j}.attn.qkv", b, q)
            o = self._new(x.H, x.W, c)
            d = cabi.AttnDesc(q.cview(), o.cview(), self.B, n_tok, heads, kd, hd, float(kd ** -0.5))
            cabi.check(self.lib.y11_plan_add_attention(self.plan, C.byref(d)), "plan_add_attention")
            self.ops.append(OpRecord("attention", f"{p}.m.{j}.attn", 2.0 * self.B * heads * n_tok * n_tok * (kd + hd),
                                     self.B * n_tok * (q.c + c) * 2))
            t = self._new(x.H, x.W, c)
            self._dw(f"{p}.m.{j}.attn.pe", q.sub(2 * heads * kd, c), t, res=o)
            self._conv(f"{p}.m.{j}.attn.proj", t, b, res=b)        # x = x + attn(x), in place on the b slice
            f = self._new(x.H, x.W, 2 * c)
            self._conv(f"{p}.m.{j}.ffn.0", b, f)
            self._conv(f"{p}.m.{j}.ffn.1", f, b, res=b)            # x = x + ffn(x)
        self._conv(f"{p}.cv2", pbuf, out)

    def _upsample(self, x: V, out: V):
        d = cabi.UpsampleDesc(x.cview(), out.cview(), self.B, x.H, x.W)
        cabi.check(self.lib.y11_plan_add_upsample(self.plan, C.byref(d)), "plan_add_upsample")
        self.ops.append(OpRecord("upsample", "upsample", 0.0, self.B * x.H * x.W * x.c * 2 * 5))

    def _detect_level(self, sp: T.LayerSpec, l: int, x: V):
        """One level of the Detect head.  Its box tower and class tower are independent chains: each gets its own lane
        (parallel branch of the CUDA graph), forked as soon as the level's feature map exists."""
        p = f"model.{sp.index}"
        c2, c3 = T.detect_dims(sp.ch_in, self.nc)
        lane_box, lane_cls = 1 + 2 * l, 2 + 2 * l
        for lane in (lane_box, lane_cls):
            cabi.check(self.lib.y11_plan_fork(self.plan, lane), "plan_fork")
        head = self._alloc(x.H, x.W, self.no, torch.float32)
        self.head.append(head)
        cabi.check(self.lib.y11_plan_set_lane(self.plan, lane_box), "plan_set_lane")
        n0, n1, n2 = (f"{p}.cv2.{l}.{j}" for j in range(3))
        q01, q12 = self._fp8_edge(n0, n1), self._fp8_edge(n1, n2)
        t1 = self._new8(x.H, x.W, c2) if q01 else self._new(x.H, x.W, c2)
        t2 = self._new8(x.H, x.W, c2) if q12 else self._new(x.H, x.W, c2)
        self._conv(n0, x, t1, out_fp8_scale=self.fp8[n0] if q01 else None)
        self._conv(n1 + "#fp8" if q01 else n1, t1, t2, out_fp8_scale=self.fp8[n1] if q12 else None)
        self._conv(n2 + "#fp8" if q12 else n2, t2, V(head, 0, 64), out_f32=True)
        cabi.check(self.lib.y11_plan_set_lane(self.plan, lane_cls), "plan_set_lane")
        u1 = self._new(x.H, x.W, x.c)
        u2 = self._new(x.H, x.W, c3)
        u3 = self._new(x.H, x.W, c3)
        u4 = self._new(x.H, x.W, c3)
        self._dw(f"{p}.cv3.{l}.0.0", x, u1)
        self._conv(f"{p}.cv3.{l}.0.1", u1, u2)
        self._dw(f"{p}.cv3.{l}.1.0", u2, u3)
        self._conv(f"{p}.cv3.{l}.1.1", u3, u4)
        self._conv(f"{p}.cv3.{l}.2", u4, V(head, 64, pad16(self.nc)), out_f32=True)
        self.cls_ops.append((len(self.ops) - 1, sum(h.shape[1] * h.shape[2] for h in self.head[:-1])))   # (op index, anchor offset)
        cabi.check(self.lib.y11_plan_set_lane(self.plan, 0), "plan_set_lane")
        self.head_lanes += [lane_box, lane_cls]

    # ---- graph ----------------------------------------------------------------------------------
    def _build(self):
        specs = T.layer_specs(self.scale)
        H, W = self.H, self.W
        # where each layer's output lives: concat consumers own the storage, producers write into slices
        concat_of: Dict[int, Tuple[int, int]] = {}   # producer layer -> (concat layer, channel offset)
        folded_cats = {v[3] for v in self.folds.values()}
        folded_ups = {v[2] for v in self.folds.values()}
        for sp in specs:
            if sp.kind == "Concat" and sp.index not in folded_cats:
                off = 0
                for src in sp.frm:
                    concat_of[src] = (sp.index, off)
                    off += specs[src].c2
        cat_buf: Dict[int, V] = {}
        outs: Dict[int, V] = {}
        hw: Dict[int, Tuple[int, int]] = {}

        def out_view(sp: T.LayerSpec, h: int, w: int) -> V:
            if sp.index in concat_of:
                ci, off = concat_of[sp.index]
                if ci not in cat_buf:
                    cat_buf[ci] = self._new(h, w, specs[ci].c2)
                return cat_buf[ci].sub(off, sp.c2)
            return self._new(h, w, sp.c2)

        detect = next(sp for sp in specs if sp.kind == "Detect")
        self.head_lanes: List[int] = []

        s2d = S2D_STEM and self.packed["model.1"].k == 2 and 0 not in concat_of

        def out_hw(sp: T.LayerSpec, src_hw: Dict[int, Tuple[int, int]]) -> Tuple[int, int]:
            """LOGICAL output size of a layer (the stem's space-to-depth buffer is stored at half of it)."""
            if sp.index == 0:
                return H // 2, W // 2
            ih, iw = src_hw[sp.frm[0]]
            return ((ih + 1) // 2, (iw + 1) // 2) if sp.kind == "Conv" else (ih, iw)

        plain_out_view = out_view

        def out_view(sp: T.LayerSpec, h: int, w: int) -> V:  # noqa: F811
            if sp.index == 0 and s2d:
                return self._new(h // 2, w // 2, 4 * sp.c2)
            return plain_out_view(sp, h, w)

        def emit(sp: T.LayerSpec, x: Optional[V], o: V, inp: torch.Tensor):
            """Ops of one backbone/neck layer: x -> o (stem: the letterboxed frames `inp` -> o)."""
            if sp.kind == "Conv" and sp.index == 0:
                pc = self.packed["model.0"]
                nb = inp.shape[0]
                s2d = int(o.c == 4 * sp.c2)          # o is then the [H/4, W/4, 4*c2] space-to-depth tensor
                if s2d and self.packed["model.1"].s2d_block:
                    s2d = 2                          # permuted block order (compact 2x2 form of model.1)
                d = cabi.StemDesc(inp.data_ptr(), o.cview(), pc.w.data_ptr(), pc.b.data_ptr(), nb, H, W, H // 2, W // 2, s2d)
                cabi.check(self.lib.y11_plan_add_stem(self.plan, C.byref(d)), "plan_add_stem")
                self.ops.append(OpRecord("stem", "model.0", 2.0 * nb * (H // 2) * (W // 2) * sp.c2 * 27,
                                         nb * (H * W * 3 * 2 + (H // 2) * (W // 2) * sp.c2 * 2), o))
            elif sp.kind == "Conv":
                self._conv(f"model.{sp.index}", x, o)
            else:
                {"C3k2": self._c3k2, "SPPF": self._sppf, "C2PSA": self._c2psa}[sp.kind](sp, x, o)

        n_prefix = 0
        if self.chunks > 1:
            # Chunk-major prefix: full-batch output buffers first, then for every batch chunk the ops of layers
            # 0..PREFIX_LAYERS-1 on that chunk's slice (intermediates are chunk-sized and private to the chunk).
            n_prefix = self.PREFIX_LAYERS
            assert all(sp.kind in ("Conv", "C3k2") and tuple(sp.frm) == (sp.index - 1,) for sp in specs[1:n_prefix])
            for sp in specs[:n_prefix]:
                hw[sp.index] = out_hw(sp, hw)
                outs[sp.index] = out_view(sp, *hw[sp.index])
            Bc = self.B // self.chunks
            for c in range(self.chunks):
                first = len(self.ops)
                self._cur_B = Bc
                x = None
                for sp in specs[:n_prefix]:
                    o = outs[sp.index].bslice(c * Bc, Bc)
                    emit(sp, x, o, self.input[c * Bc:(c + 1) * Bc])
                    x = o
                self._cur_B = self.B
                self.prefix_ranges.append((first, len(self.ops)))
        for sp in specs[n_prefix:]:
            if sp.index in self.folds:
                low_i, skip_i, _, cat_i = self.folds[sp.index]
                hw[sp.index] = hw[cat_i]
                o = out_view(sp, *hw[sp.index])
                self._c3k2(sp, outs[skip_i], o, low=outs[low_i])
            elif sp.kind in ("Conv", "C3k2", "SPPF", "C2PSA"):
                x = outs[sp.frm[0]] if sp.index else None
                hw[sp.index] = out_hw(sp, hw)
                o = out_view(sp, *hw[sp.index])
                emit(sp, x, o, self.input)
            elif sp.index in folded_ups or sp.index in folded_cats:
                src = outs[sp.frm[0]] if sp.index in folded_ups else None
                hw[sp.index] = (2 * src.H, 2 * src.W) if src is not None else hw[sp.frm[0]]
                o = None                                         # never materialised
            elif sp.kind == "Upsample":
                x = outs[sp.frm[0]]
                hw[sp.index] = (2 * x.H, 2 * x.W)
                o = out_view(sp, 2 * x.H, 2 * x.W)
                self._upsample(x, o)
            elif sp.kind == "Concat":
                o = cat_buf[sp.index]
                hw[sp.index] = (o.H, o.W)
            elif sp.kind == "Detect":
                assert len(self.head) == len(sp.frm)     # every level was emitted right after its feature map
                for lane in self.head_lanes:
                    cabi.check(self.lib.y11_plan_join(self.plan, lane), "plan_join")
                o = None
            outs[sp.index] = o
            if sp.index in detect.frm:
                self._detect_level(detect, detect.frm.index(sp.index), o)
        self.rest_first = self.prefix_ranges[-1][1] if self.prefix_ranges else 0
        self.layer_out = outs
        self.n_ops = self.lib.y11_plan_num_ops(self.plan)
        self.n_launches = self.lib.y11_plan_num_launches(self.plan)
        assert self.n_ops == len(self.ops)

    # ---- execution ------------------------------------------------------------------------------
    def run(self, stream: int) -> None:
        cabi.check(self.lib.y11_plan_run(self.plan, C.c_void_p(stream)), "y11_plan_run")

    def run_ops(self, first: int, last: int, stream: int) -> None:
        """Ops [first, last) with their lanes (parallel Detect towers)."""
        cabi.check(self.lib.y11_plan_run_ops(self.plan, first, last, C.c_void_p(stream)), "y11_plan_run_ops")

    def run_range(self, first: int, last: int, stream: int) -> None:
        cabi.check(self.lib.y11_plan_run_range(self.plan, first, last, C.c_void_p(stream)), "y11_plan_run_range")

    def set_cls_emit(self, conf: Optional[float]) -> bool:
        """Single-label prediction at threshold `conf`: the three class-logit convs (cv3.l.2) reduce every anchor to (maximum
        logit, class) in their epilogue and list the anchors that can pass `conf` instead of storing [B,A,nc] fp32 logits
        (y11_plan_set_cls_emit); `engine.postprocess` then consumes `self.emit_list` / `self.emit_count`.  None: back to stored
        logits (multi-label, raw-head consumers).  Returns whether emit mode is on.  Like `set_stem_source` it edits the plan,
        i.e. it holds for launches and graph captures made after the call."""
        if (conf is None or self.A >= 65536 or self.conv_impl != cabi.IMPL_TCGEN05 or not FUSE_CLS_DECODE
                or pad16(self.nc) > 128):        # more than one N tile of class logits: a row is split over CTAs, keep the scan kernel
            if self.emit_conf is not None:
                for op, _ in self.cls_ops:
                    cabi.check(self.lib.y11_plan_set_cls_emit(self.plan, op, None), "y11_plan_set_cls_emit")
                self.emit_conf = None
            return False
        if self.emit_conf == conf:
            return True
        if self.emit_list is None:
            self.emit_list = torch.zeros((self.B, self.A, 4), dtype=torch.int32, device=self.device)
            self.emit_count = torch.zeros((self.B,), dtype=torch.int32, device=self.device)
        cc = min(max(float(conf), 1e-30), 1.0 - 1e-9)
        thr = math.log(cc / (1.0 - cc)) - 1e-2      # sigmoid(x) > conf  =>  x > logit(conf) - slack (as y11_detect_postprocess)
        for op, aoff in self.cls_ops:
            e = cabi.ClsEmit(self.emit_list.data_ptr(), self.emit_count.data_ptr(), self.A, self.nc, aoff, thr)
            cabi.check(self.lib.y11_plan_set_cls_emit(self.plan, op, C.byref(e)), "y11_plan_set_cls_emit")
        self.emit_conf = conf
        return True

    def set_stem_source(self, images_dev_ptr: Optional[int]) -> None:
        """Point the stem at a device array of `y11_image` descriptors of frames ALREADY at network resolution (the stem then
        reads the uint8 frames itself and no letterbox launch is needed), or back at `self.input` (None)."""
        cabi.check(self.lib.y11_plan_set_stem_source(self.plan, C.c_void_p(images_dev_ptr or 0)), "y11_plan_set_stem_source")

    def run_timed(self, stream: int) -> List[float]:
        ms = (C.c_float * self.n_ops)()
        cabi.check(self.lib.y11_plan_run_timed(self.plan, C.c_void_p(stream), ms), "y11_plan_run_timed")
        return list(ms)

    def variants(self) -> List[Tuple[int, int, int, int]]:
        """Per op: (lsu, epi_warp, ctas_per_sm, bn) of the tcgen05 launch variant in use; (-1,)*4 for other kernels."""
        out = []
        v = (C.c_int32 * 4)()
        for i in range(self.n_ops):
            cabi.check(self.lib.y11_plan_op_variant(self.plan, i, v), "y11_plan_op_variant")
            out.append(tuple(v))
        return out

    def head_desc(self) -> cabi.HeadDesc:
        hd = cabi.HeadDesc()
        for l, h in enumerate(self.head):
            hd.head[l] = h.data_ptr()
            hd.hl[l], hd.wl[l] = h.shape[1], h.shape[2]
            hd.stride[l] = float(T.STRIDES[l])
        hd.nl, hd.B, hd.nc, hd.row_stride = len(self.head), self.B, self.nc, self.no
        return hd

    def raw_head(self) -> torch.Tensor:
        """[B, 64+nc, A] fp32 in ultralytics' layout (test helper; a torch view/permute, not a kernel)."""
        assert self.emit_conf is None, "the plan last ran in class-emit mode: no class logits were stored (forward(net) first)"
        parts = []
        for h in self.head:
            x = torch.cat((h[..., :64], h[..., 64:64 + self.nc]), -1)
            parts.append(x.reshape(self.B, -1, 64 + self.nc))
        return torch.cat(parts, 1).permute(0, 2, 1).contiguous()

    @property
    def conv_flops(self) -> float:
        return sum(o.flops for o in self.ops if o.kind in ("conv", "stem", "dwconv"))
