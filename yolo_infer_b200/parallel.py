"""Image-sharded multi-GPU execution: one process per GPU, weights replicated, each rank owns a
contiguous slice of the global batch, and NO collective on the data path - the only exchange is gathering the
fixed-shape results (`det` fp32 [B_local, max_det, 6] + `count` int32 [B_local]).  The reference has no
multi-GPU inference at all (README.md:13 advertises it; no implementation), so this mirrors nothing
and simply keeps the single-call `predict` contract: rank 0 ends up with every image's detections.

Works with backend "nccl" (GPU tensors, NVLink/NVSwitch) and "gloo" (CPU tensors; used by the CPU tests).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of `n_items` owned by `rank`; earlier ranks take the remainder."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_detections(det: torch.Tensor, count: torch.Tensor, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """All ranks contribute equally-shaped (det [b,max_det,6], count [b]); every rank gets the global
    tensors in rank order (== global image order under shard_range with equal shards)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return det, count
    world = dist.get_world_size(group)
    dets = [torch.empty_like(det) for _ in range(world)]
    counts = [torch.empty_like(count) for _ in range(world)]
    dist.all_gather(dets, det.contiguous(), group=group)
    dist.all_gather(counts, count.contiguous(), group=group)
    return torch.cat(dets, 0), torch.cat(counts, 0)


def gather_flat(flat: torch.Tensor, out: Optional[torch.Tensor] = None, group=None) -> torch.Tensor:
    """ONE collective per step: every rank contributes its flat result buffer ([b*max_det*6] fp32 det followed by [b] int32
    count bit-cast into the same fp32 tensor, see `engine.YOLO.postprocess`) and receives [world, len(flat)] in rank order.
    `out` is a persistent [world * len(flat)] buffer (no allocation, no concatenation on the hot path)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return flat.view(1, -1)
    world = dist.get_world_size(group)
    if out is None:
        out = torch.empty((world * flat.numel(),), dtype=flat.dtype, device=flat.device)
    dist.all_gather_into_tensor(out, flat, group=group)
    return out.view(world, -1)


def split_flat(gathered: torch.Tensor, b: int, max_det: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """[world, b*max_det*6 + b] -> (det [world*b, max_det, 6] fp32, count [world*b] int32) in global image order."""
    world = gathered.shape[0]
    det = gathered[:, : b * max_det * 6].reshape(world * b, max_det, 6)
    count = gathered[:, b * max_det * 6:].contiguous().view(torch.int32).reshape(world * b)
    return det, count


def pad_shard(det: torch.Tensor, count: torch.Tensor, b_max: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Pad a ragged last shard to b_max images so all_gather sees equal shapes (count 0 for padding)."""
    b = det.shape[0]
    if b == b_max:
        return det, count
    pd = det.new_zeros((b_max - b,) + tuple(det.shape[1:]))
    pc = count.new_zeros((b_max - b,))
    return torch.cat((det, pd), 0), torch.cat((count, pc), 0)


def unpad_gathered(det: torch.Tensor, count: torch.Tensor, n_items: int, world: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Inverse of pad_shard after gather_detections: drop padding rows, restoring global image order."""
    b_max = det.shape[0] // world
    keep: List[int] = []
    for r in range(world):
        lo, hi = shard_range(n_items, r, world)
        keep.extend(range(r * b_max, r * b_max + (hi - lo)))
    idx = torch.as_tensor(keep, device=det.device)
    return det.index_select(0, idx), count.index_select(0, idx)
