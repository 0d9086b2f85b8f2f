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


# ------------------------------------------------------------------------------------------------------------------------
# Result push: the gather without a collective.
#
# Every rank's NMS kernel writes its kept boxes STRAIGHT INTO RANK 0's result buffer through an NVLink peer mapping
# (torch symmetric memory = CUDA VMM allocations exchanged between the ranks of one node), then bumps a signal word on
# rank 0 (y11_detect_postprocess_push).  No NCCL kernel competes for SM slots with the persistent conv grids, no extra
# stream, no copy: the 461 KB per rank ride on the stores the kernel makes anyway.  NCCL all-gather stays as the tested
# fallback (`mode == "nccl"`) for boxes without peer access.
# ------------------------------------------------------------------------------------------------------------------------
class ResultExchange:
    """Symmetric buffer layout (fp32 words), identical on every rank; only rank 0's copy receives results:
         results [slots][world][n_flat] | signals uint32 [slots][world] | free uint32 [slots] | done int32 [slots]
    `signals[s][r]` counts rank r's completed passes into slot s (bumped by rank r's NMS kernel, lives on rank 0);
    `free[s]` counts how often rank 0 has consumed slot s (bumped by rank 0 on every rank's copy); `done[s]` is the
    device-local last-CTA counter of the push."""

    def __init__(self, lib, engine_handle, n_flat: int, slots: int, device, group=None, mode: str = "auto"):
        self.lib, self.h = lib, engine_handle
        self.n_flat, self.slots, self.device = n_flat, slots, device
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.group = group
        self.mode = "local"
        self.why = ""
        self.produced = [0] * slots          # passes this rank has enqueued into each slot
        self.consumed = [0] * slots          # rank 0: slots consumed so far
        self.base = None
        w, s = self.world, slots
        self.o_sig = s * w * n_flat
        self.o_free = self.o_sig + s * w
        self.o_done = self.o_free + s
        total = (self.o_done + s + 63) // 64 * 64
        if self.world > 1 and mode in ("auto", "push"):
            try:
                import torch.distributed._symmetric_memory as symm
                self.buf = symm.empty(total, dtype=torch.float32, device=device)
                self.buf.zero_()
                torch.cuda.synchronize(device)
                self.hdl = symm.rendezvous(self.buf, group=group if group is not None else dist.group.WORLD)
                self.peer = [self.hdl.get_buffer(r, (total,), torch.float32) for r in range(w)]
                self.mode = "push"
            except Exception as e:           # no peer access / no VMM support: fall back to the collective
                if mode == "push":
                    raise
                self.why = f"{type(e).__name__}: {e}"
        if self.mode != "push":
            self.buf = torch.zeros((total,), dtype=torch.float32, device=device)
            self.peer = [self.buf] * w
            if self.world > 1:
                self.mode = "nccl"
                self.local_flat = [torch.zeros((n_flat,), dtype=torch.float32, device=device) for _ in range(s)]
        self.root = self.peer[0]             # rank 0's buffer as seen from this rank

    # ---- producer side -------------------------------------------------------------------------------------------------
    def out_flat(self, slot: int) -> torch.Tensor:
        """Where this rank's pipeline for `slot` writes its flat (det | count) results."""
        if self.mode == "nccl":
            return self.local_flat[slot]
        o = (slot * self.world + self.rank) * self.n_flat
        return self.root[o:o + self.n_flat]

    def push_ptrs(self, slot: int):
        """(done_counter, signal) device pointers for y11_detect_postprocess_push; None when nothing has to be signalled."""
        if self.mode != "push":
            return None
        done = self.buf.data_ptr() + 4 * (self.o_done + slot)
        sig = self.root.data_ptr() + 4 * (self.o_sig + slot * self.world + self.rank)
        return (done, sig)

    def arm(self) -> None:
        """After every pipeline has been built (their warm-up passes bumped the signals): remember the signal values."""
        cuda = torch.device(self.device).type == "cuda"
        if cuda:
            torch.cuda.synchronize(self.device)
        if self.world > 1:
            dist.barrier(group=self.group)
            if cuda:
                torch.cuda.synchronize(self.device)
        words = self.buf.view(torch.int32)
        self.base = words[self.o_sig:self.o_sig + self.slots * self.world].cpu().tolist()
        self.base_free = words[self.o_free:self.o_free + self.slots].cpu().tolist()
        if self.world > 1:
            dist.barrier(group=self.group)

    def before_produce(self, slot: int, stream, backpressure: bool) -> None:
        """Call before enqueueing the pass that overwrites `slot`: with back-pressure the stream parks until rank 0 has
        consumed the previous contents of the slot."""
        k = self.produced[slot]
        if backpressure and self.mode == "push" and k >= 1 and self.rank != 0:
            ptr = self.buf.data_ptr() + 4 * (self.o_free + slot)
            rc = self.lib.y11_wait_signals(self.h, ptr, 1, (self.base_free[slot] + k) & 0xFFFFFFFF, stream)
            if rc:
                raise RuntimeError("y11_wait_signals failed")
        self.produced[slot] = k + 1

    def after_produce(self, slot: int, comm_stream=None, gathered: Optional[torch.Tensor] = None) -> None:
        """nccl mode: the collective that replaces the push (on `comm_stream`, into `gathered` [world * n_flat])."""
        if self.mode == "nccl":
            dist.all_gather_into_tensor(gathered, self.local_flat[slot], group=self.group)

    # ---- consumer side (rank 0) ----------------------------------------------------------------------------------------
    def wait_all(self, slot: int, stream) -> None:
        """Rank 0: park `stream` until every rank's pass number produced[slot] into `slot` has landed."""
        if self.mode != "push" or self.rank != 0:
            return
        k = self.produced[slot]
        base = self.base[slot * self.world:(slot + 1) * self.world]
        # all ranks call the exchange in lockstep order, so every rank's target is its base + k; the wait kernel takes ONE
        # target, so the per-rank bases (equal in practice: every rank builds the same pipelines) must agree
        assert len(set(base)) == 1, f"ranks built different numbers of passes into slot {slot}: {base}"
        ptr = self.buf.data_ptr() + 4 * (self.o_sig + slot * self.world)
        rc = self.lib.y11_wait_signals(self.h, ptr, self.world, (base[0] + k) & 0xFFFFFFFF, stream)
        if rc:
            raise RuntimeError("y11_wait_signals failed")

    def release(self, slot: int) -> None:
        """Rank 0, on the current stream, after it has read `slot`: tell every rank the slot may be overwritten."""
        if self.mode != "push" or self.rank != 0:
            return
        self.consumed[slot] += 1
        for r in range(1, self.world):
            self.peer[r].view(torch.int32)[self.o_free + slot:self.o_free + slot + 1].add_(1)

    def slot_results(self, slot: int) -> torch.Tensor:
        """Rank 0: [world, n_flat] view of a slot's gathered results."""
        o = slot * self.world * self.n_flat
        return self.root[o:o + self.world * self.n_flat].view(self.world, self.n_flat)


class ShardedPredictor:
    """Multi-process (torchrun, one rank per GPU) form of `YOLO.predict` for fixed-shape uint8 batches: every rank calls
    `predict(local_frames)` with its slice of the global batch (shard_range order); rank 0 returns the `Results` of the GLOBAL
    batch, the other ranks return `[]`.  Results travel by result push (see ResultExchange); back-pressure keeps a fast rank
    from overwriting a slot rank 0 has not read yet."""

    def __init__(self, eng, B_local: int, h0: int, w0: int, imgsz=640, rect: bool = True, conf: float = 0.25, iou: float = 0.7,
                 max_det: int = 300, agnostic: bool = False, multi_label: bool = False, slots: int = 2, mode: str = "auto"):
        from . import _cabi as cabi
        self.eng, self.B, self.h0, self.w0, self.max_det = eng, B_local, h0, w0, max_det
        eng._ensure_device()
        n_flat = B_local * max_det * 6 + B_local
        self.x = ResultExchange(cabi.load(), eng._engine, n_flat, slots, eng.device, mode=mode)
        self.pipes = []
        with torch.cuda.device(eng.device):
            for s in range(slots):
                self.pipes.append(eng.pipeline(B_local, h0, w0, imgsz, rect, conf, iou, max_det, agnostic, multi_label, replica=s,
                                               out_flat=self.x.out_flat(s), push=self.x.push_ptrs(s)))
            self.gathered = (torch.empty((self.x.world * n_flat,), dtype=torch.float32, device=eng.device)
                             if self.x.mode == "nccl" else None)
            self.pins = ([torch.empty((self.x.world, n_flat), dtype=torch.float32).pin_memory() for _ in range(slots)]
                         if self.x.rank == 0 else None)
        self.x.arm()
        self.step = 0

    def predict(self, frames: torch.Tensor):
        from .results import Results
        eng, x = self.eng, self.x
        slot = self.step % len(self.pipes)
        self.step += 1
        with eng._lock, torch.cuda.device(eng.device), torch.inference_mode():
            st = torch.cuda.current_stream(eng.device)
            x.before_produce(slot, C_void(st.cuda_stream), backpressure=True)
            self.pipes[slot].run(frames)
            if x.mode == "nccl":
                x.after_produce(slot, None, self.gathered)
                src = self.gathered.view(x.world, -1)
            else:
                x.wait_all(slot, C_void(st.cuda_stream))
                src = x.slot_results(slot)
            if x.rank != 0:
                st.synchronize()
                return []
            pin = self.pins[slot]
            pin.copy_(src, non_blocking=True)
            x.release(slot)
            st.synchronize()
            det, count = split_flat(pin, self.B, self.max_det)     # views of the slot's pinned buffer: valid until the slot's next use
            from .results import ResultsBatch
            return ResultsBatch(None, det, count.tolist(), eng.names, (self.h0, self.w0),
                                speed={"preprocess": 0.0, "inference": 0.0, "postprocess": 0.0})


def C_void(v: int):
    import ctypes
    return ctypes.c_void_p(v)
