"""CPU: the N>1 path (image sharding + result gather) on world_size-2 gloo."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from yolo_infer_b200.parallel import gather_detections, gather_flat, pad_shard, shard_range, split_flat, unpad_gathered


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 64, 65, 512):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _worker(rank, world, port, n_items, max_det, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    det_all = torch.rand(n_items, max_det, 6, generator=g)
    cnt_all = torch.randint(0, max_det + 1, (n_items,), generator=g, dtype=torch.int32)
    lo, hi = shard_range(n_items, rank, world)
    b_max = shard_range(n_items, 0, world)[1]
    det, cnt = pad_shard(det_all[lo:hi], cnt_all[lo:hi], b_max)
    gd, gc = gather_detections(det, cnt)
    gd, gc = unpad_gathered(gd, gc, n_items, world)
    ok = torch.equal(gd, det_all) and torch.equal(gc, cnt_all)
    # single-collective form used by bench.py: det and count share one flat fp32 buffer (count bit-cast)
    flat = torch.cat((det.reshape(-1), cnt.view(torch.float32)))
    out = torch.empty(world * flat.numel())
    fd, fc = split_flat(gather_flat(flat, out), b_max, max_det)
    fd, fc = unpad_gathered(fd, fc, n_items, world)
    ok = ok and torch.equal(fd, det_all) and torch.equal(fc, cnt_all)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, bool(ok)))


@pytest.mark.parametrize("n_items", [8, 7])
def test_gather_equals_single_process_world2(n_items):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_items, 5, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def _exchange_worker(rank, world, port, q):
    """ResultExchange in its collective (fallback) mode on gloo/CPU: slots, out_flat, the gather and split_flat - the same host logic
    bench.py and parallel.ShardedPredictor drive on NCCL when the NVLink result push is unavailable."""
    from yolo_infer_b200.parallel import ResultExchange
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    B, MD, slots = 3, 4, 2
    n_flat = B * MD * 6 + B
    x = ResultExchange(None, None, n_flat, slots, torch.device("cpu"), mode="nccl")
    ok = x.mode == "nccl" and x.world == world and x.push_ptrs(0) is None
    x.arm()
    gathered = torch.empty(world * n_flat)
    for step in range(4):
        slot = step % slots
        g = torch.Generator().manual_seed(100 * step + rank)
        det = torch.rand(B, MD, 6, generator=g)
        cnt = torch.randint(0, MD + 1, (B,), generator=g, dtype=torch.int32)
        x.before_produce(slot, None, backpressure=True)
        x.out_flat(slot).copy_(torch.cat((det.reshape(-1), cnt.view(torch.float32))))
        x.after_produce(slot, None, gathered)
        x.wait_all(slot, None)
        gd, gc = split_flat(gathered.view(world, -1), B, MD)
        for r in range(world):
            g = torch.Generator().manual_seed(100 * step + r)
            d = torch.rand(B, MD, 6, generator=g)
            c = torch.randint(0, MD + 1, (B,), generator=g, dtype=torch.int32)
            ok = ok and torch.equal(gd[r * B:(r + 1) * B], d) and torch.equal(gc[r * B:(r + 1) * B], c)
        x.release(slot)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, bool(ok)))


def test_result_exchange_collective_mode_world2():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_exchange_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]
