#!/usr/bin/env python
"""bench.py - end-to-end detection throughput of the B200 YOLO11 path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--model s|n|m|l|x] [--batch 64] [--imgsz 640]
  python bench.py --impl reference ...     # the reference's CPU path (oracle restatement), same metric/config
  torchrun --nproc-per-node N bench.py --gpus N ...   # one rank per GPU, image-sharded (weak scaling)

Headline workload = the north-star target, BASELINE.json configs[2]: YOLO11s, 640x640, batch 64 per GPU (configs[1], YOLO11n,
and YOLO11m ride along as `value_n` / `value_m`, a few steps each, 1 GPU only).

A "step" = one pass of the hot path over one batch of synthetic frames per GPU:
  preprocess (u8 HWC -> bf16 NHWC) -> fused YOLO11 forward -> decode -> NMS -> max_det results (gathered on rank 0 at N > 1).
`value`   : images/s with the uint8 frames already resident in HBM.  The timed block is EXACTLY --steps steps between
            barrier + synchronize, CUDA-event timed, max over ranks; the block is repeated (--repeats) and the MEDIAN block is
            reported, all blocks listed in `ms_per_step_blocks`.
`e2e`     : images/s through the public API from pinned HOST frames, H2D + result D2H inside the timing (at N > 1 through
            parallel.ShardedPredictor: rank 0 receives the global batch's results inside the timed region).
`roofline`: conv FLOPs of one step / ms_per_step of the timed loop (frac against the sustained AND the burst bf16 peak); the
            serialised per-op figure (events around every launch of one extra pass) is kept as `roofline.serialised`.
`roofline_pre` / `roofline_post`: letterbox, decode and sort+NMS kernels alone, algorithmic bytes / time vs the measured HBM peak.
`cpu_baseline`: the oracle (reference CPU path restated) on the host cores, bounded sample (rank 0, N=1 only).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))


def _baseline_metric() -> str:
    """BASELINE.json's metric string when the file is there (the driver compares lines by it); both arms use the same one.
    `value` is its throughput half (preprocess + forward + decode + NMS, images/s); `latency_b1` its batch-1 latency half."""
    try:
        return json.loads((ROOT / "BASELINE.json").read_text())["metric"]
    except Exception:
        return "end-to-end images/sec @640 (letterbox+forward+decode+NMS)"


METRIC = _baseline_metric()
UNIT = "images/s"
CONF, IOU, MAX_DET = 0.25, 0.7, 300   # the reference benchmark's thresholds (ultralytics predict defaults)
CONV_GFLOP = {"n": 6.481, "s": 21.467, "m": 67.982, "l": 86.910, "x": 194.904}   # SURVEY 8(d), per 640x640 image


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="s", help="YOLO11 scale; default = the north-star target, BASELINE.json configs[2] (YOLO11s)")
    ap.add_argument("--batch", type=int, default=64, help="images per GPU per step")
    ap.add_argument("--imgsz", type=int, default=640)
    ap.add_argument("--repeats", type=int, default=5, help="timed blocks of --steps steps; the median block is reported")
    ap.add_argument("--extras", default="n,m", help="other scales measured (value only, few steps) at 1 GPU: value_<scale>")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-batch", type=int, default=16, help="images per step of the CPU reference arm (bounded sample)")
    ap.add_argument("--per-op", action="store_true", help="also print a per-op time table to stderr")
    ap.add_argument("--skip-condition", action="store_true",
                    help="profiling only: skip the init-time weight conditioning pass (keeps the launch list short under ncu)")
    ap.add_argument("--skip-e2e", action="store_true", help="profiling only: value loop only")
    ap.add_argument("--weights-cache", default=None,
                    help="profiling only: file to keep the conditioned state_dict in, so that the ncu passes start from the conditioned "
                         "weights (realistic candidate counts in decode / NMS) without the ~1000 launches of the conditioning pass")
    ap.add_argument("--no-graph", action="store_true", help="launch kernels one by one instead of replaying CUDA graphs")
    ap.add_argument("--latency-iters", type=int, default=200, help="batch-1 latency samples (0 = skip)")
    ap.add_argument("--dump-ops", default=None, help="write the plan's op list (kind, name, algorithmic flops/bytes, launch variant) "
                                                     "to this JSON file (joined with ncu launch lists by tools/ncu_join.py)")
    ap.add_argument("--no-gather", action="store_true", help="experiment only (N > 1): skip the per-step result gather")
    ap.add_argument("--gather", default="auto", choices=["auto", "push", "nccl"],
                    help="N > 1: result push over NVLink peer memory (default when available) or the NCCL all-gather fallback")
    ap.add_argument("--streams", type=int, default=2,
                    help="steps in flight: consecutive steps alternate between this many streams (own buffers each), so the "
                         "under-filled tail of one step (NMS: one CTA per image) overlaps the head of the next")
    return ap.parse_args()


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"hbm_gbs": d.get("hbm_gbs", 6650.0), "tflops_burst": d.get("bf16_tflops", 1590.0),
                "tflops_sustained": d.get("bf16_tflops_sustained", 1400.0), "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "tflops_burst": 1590.0, "tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


def synth_frames(batch: int, imgsz: int, seed: int, h: int = 0, w: int = 0):
    import torch
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, 256, (batch, h or imgsz, w or imgsz, 3), generator=g, dtype=torch.uint8)


class ClockSampler:
    """Samples SM clock / throttle reasons with NVML while the timed region runs."""

    def __init__(self, index: int):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.02)

    def __enter__(self):
        if self.nv:
            self._t = threading.Thread(target=self._loop, daemon=True)
            self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._t:
            self._t.join(timeout=2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# ----------------------------------------------------------------------------------------------- CPU arms
def cpu_reference_arm(args, sd, steps: int, warmup: int, batch: int):
    """The reference's own CPU implementation of the path = oracle/ (ultralytics-on-torch-CPU restated; ultralytics itself
    is not installable here).  All host threads; returns (images/s, ms/step, cores, sample description)."""
    import numpy as np
    import torch
    from oracle import pipeline_ref as P
    from oracle import yolo11_ref as R
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model = R.DetectionModel(args.model)
    model.load_state_dict(sd, strict=False)
    model.eval().fuse()
    frames = synth_frames(batch, args.imgsz, 1234).numpy()
    imgs = [np.ascontiguousarray(f) for f in frames]
    kw = dict(conf=CONF, iou=IOU, max_det=MAX_DET, imgsz=args.imgsz)
    for _ in range(warmup):
        P.predict(model, imgs, **kw)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        P.predict(model, imgs, **kw)
        ts.append(time.perf_counter() - t0)
    ms = 1e3 * sum(ts) / len(ts)
    return batch / (ms / 1e3), ms, cores, (f"{steps} x predict({batch} frames {args.imgsz}x{args.imgsz}) after {warmup} warm-up, fused fp32, "
                                          f"torch {torch.__version__} CPU, {cores} threads")


def run_reference(args):
    """--impl reference: the reference's CPU path on the box's host cores, the driver's --steps / --warmup honoured, every step
    a bounded sample (--ref-batch frames) of the headline workload.  Rank 0 only under torchrun."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import yolo11_ref as R
    ref = R.build(args.model, init="calibrated", seed=0)
    sd = ref.state_dict()
    steps, warmup = max(1, args.steps), max(1, args.warmup)
    ips, ms, cores, sample = cpu_reference_arm(args, sd, steps, warmup, args.ref_batch)
    line = {"impl": "reference", "metric": METRIC, "value": ips, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"YOLO11{args.model} detect, {args.imgsz}x{args.imgsz}, batch {args.batch}/GPU, synthetic uint8 frames "
                                   f"(reference CPU path: {args.ref_batch}-frame steps, a bounded sample of that workload)",
                       "conf": CONF, "iou": IOU, "max_det": MAX_DET},
            "cpu_baseline": {"value": ips, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": ips, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------- GPU arm
class ResidentLoop:
    """The `value` loop of one engine: NROT device-resident uint8 batches (> L2 together with the activations) rotated between
    steps, one CUDA-graph pipeline per batch, consecutive steps alternating between NS streams (own buffers each)."""
    NROT = 4

    def __init__(self, eng, args, B, S, dev, rank, world, exchange_mode):
        import torch
        self.torch, self.eng, self.B, self.S, self.dev, self.world, self.rank = torch, eng, B, S, dev, world, rank
        self.stream = torch.cuda.current_stream(dev)
        self.host_batches = [synth_frames(B, S, 1000 * rank + i).pin_memory() for i in range(self.NROT)]
        self.dev_batches = [hb.to(dev) for hb in self.host_batches]
        self.NS = max(1, min(args.streams, self.NROT))
        self.streams = [self.stream] + [torch.cuda.Stream(dev) for _ in range(self.NS - 1)]
        self.x = None
        self.comm = None
        if world > 1 and exchange_mode != "none":
            from yolo_infer_b200 import _cabi as cabi
            from yolo_infer_b200.parallel import ResultExchange
            self.x = ResultExchange(cabi.load(), eng._engine, B * MAX_DET * 6 + B, self.NS, dev, mode=exchange_mode)
            self.comm = torch.cuda.Stream(dev)
            if self.x.mode == "nccl":
                self.gathered = [torch.empty((world * self.x.n_flat,), dtype=torch.float32, device=dev) for _ in range(self.NS)]
                self.ready = [torch.cuda.Event() for _ in range(self.NS)]
                self.drained = [torch.cuda.Event() for _ in range(self.NS)]
                for ev in self.drained:
                    ev.record(self.stream)
        self.pipes = []
        for j, db in enumerate(self.dev_batches):
            r = j % self.NS
            with torch.cuda.stream(self.streams[r]):
                kw = {}
                if self.x is not None:
                    kw = dict(out_flat=self.x.out_flat(r), push=self.x.push_ptrs(r))
                self.pipes.append(eng.pipeline(B, S, S, S, True, CONF, IOU, MAX_DET, frames=db, graph=not args.no_graph, replica=r, **kw))
        torch.cuda.synchronize(dev)
        if self.x is not None:
            self.x.arm()
        self.launches = self.pipes[0].launches
        self.last = None

    def step(self, i: int):
        import ctypes as C
        torch = self.torch
        j = i % self.NROT
        r = j % self.NS
        st = self.streams[r]
        x = self.x
        with torch.cuda.stream(st):
            if x is not None and x.mode == "nccl":
                st.wait_event(self.drained[r])     # the previous gather of this replica's result buffer has read it
            if x is not None:
                x.before_produce(r, C.c_void_p(st.cuda_stream), backpressure=False)
            self.last = self.pipes[j].run()
            if x is not None and x.mode == "nccl":
                self.ready[r].record(st)
        if x is not None and x.mode == "nccl":
            # fallback: ONE all-gather of the fixed-shape results per step on its own stream (the compute streams never wait
            # for the other ranks, only for the gather that last READ the result buffer they overwrite)
            self.comm.wait_event(self.ready[r])
            with torch.cuda.stream(self.comm):
                x.after_produce(r, None, self.gathered[r])
                self.drained[r].record(self.comm)
        elif x is not None and x.rank == 0:
            # result push: every rank's NMS kernel wrote its results into rank 0's slot and bumped a signal; rank 0 observes
            # the arrival of step i from ALL ranks on a side stream (a one-warp wait kernel) - the gather is complete when it ends
            with torch.cuda.stream(self.comm):
                x.wait_all(r, C.c_void_p(self.comm.cuda_stream))

    def block(self, first: int, steps: int):
        """`steps` steps bracketed by events on the main stream; side streams forked before / joined after."""
        torch = self.torch
        side = self.streams[1:] + ([self.comm] if self.comm is not None else [])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(self.stream)
        for st in side:
            st.wait_stream(self.stream)
        for i in range(first, first + steps):
            self.step(i)
        for st in side:
            self.stream.wait_stream(st)
        e1.record(self.stream)
        return e0, e1


def run_ours(args):
    import torch
    import torch.distributed as dist
    from yolo_infer_b200 import topology as T
    from yolo_infer_b200.engine import YOLO, letterbox_geometry, scale_geometry
    from yolo_infer_b200.synth import condition_synthetic_weights

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 and world == 1:
        raise SystemExit("launch with torchrun for --gpus > 1 (one rank per GPU)")
    assert torch.cuda.is_available(), "bench.py needs a B200; there is no CPU fallback (use --impl reference for the CPU arm)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        try:   # bind this rank to the CPU cores / NUMA node next to its GPU: pinned staging buffers are then allocated locally
            import pynvml
            pynvml.nvmlInit()
            pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local))
        except Exception:
            pass
        # NCCL prints its version banner on STDOUT when NCCL_DEBUG is VERSION/INFO; stdout carries exactly one JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def make_engine(scale: str):
        cache = Path(f"{args.weights_cache}.{scale}.pt") if args.weights_cache else None
        if cache is not None and cache.exists():
            return YOLO.from_state_dict(torch.load(cache, weights_only=True), scale).to(dev)
        eng = YOLO.from_state_dict(T.synthetic_state_dict(scale, 80, seed=0), scale).to(dev)
        if not args.skip_condition:
            condition_synthetic_weights(eng, (S, S), batch=2, seed=0)
            if cache is not None and rank == 0:
                torch.save({k: v.cpu() for k, v in eng.model.state_dict().items()}, cache)
        return eng

    def measure_value(loop: ResidentLoop, steps: int, warmup: int, repeats: int):
        """-> (median block ms/step [max over ranks], all blocks, per-rank ms/step of the median block, clocks)."""
        for i in range(warmup):
            loop.step(i)
        barrier()
        blocks, per_rank, block_mhz = [], [], []
        with ClockSampler(local) as clocks:
            for rep in range(repeats):
                barrier()
                n0 = len(clocks.samples)
                e0, e1 = loop.block(rep * steps, steps)
                barrier()
                block_mhz.append(statistics.median(clocks.samples[n0:]) if len(clocks.samples) > n0 else None)
                mine = e0.elapsed_time(e1) / steps
                if world > 1:
                    allms = [torch.zeros(1, device=dev) for _ in range(world)]
                    dist.all_gather(allms, torch.tensor([mine], device=dev))
                    ranks_ms = [float(t) for t in allms]
                else:
                    ranks_ms = [mine]
                blocks.append(max(ranks_ms))
                per_rank.append(ranks_ms)
        med = sorted(range(len(blocks)), key=lambda k: blocks[k])[len(blocks) // 2]
        summary = clocks.summary()
        summary["sm_mhz_per_block"] = block_mhz      # back-to-back blocks heat the part up: the power cap shows here
        return blocks[med], blocks, per_rank[med], summary

    B, S = args.batch, args.imgsz
    pk = peaks()
    eng = make_engine(args.model)
    net = eng.compiled(B, S, S)
    stream = torch.cuda.current_stream(dev)
    if args.dump_ops and rank == 0:
        Path(args.dump_ops).write_text(json.dumps([{"kind": o.kind, "name": o.name, "flops": o.flops, "bytes_algo": o.bytes_algo,
                                                    "variant": list(v)} for o, v in zip(net.ops, net.variants())]))
    exchange_mode = "none" if (world == 1 or args.no_gather) else args.gather
    loop = ResidentLoop(eng, args, B, S, dev, rank, world, exchange_mode)
    ms_step, blocks, ranks_ms, clocks = measure_value(loop, args.steps, args.warmup, max(1, args.repeats))
    value = world * B / (ms_step / 1e3)
    det, cnt, ncand = loop.last
    mean_cand = float(ncand.float().mean())
    mean_det = float(cnt.float().mean())
    gather_info = None
    if world > 1:
        gather_info = {"mode": loop.x.mode if loop.x is not None else "none",
                       "what": {"push": "each rank's NMS kernel stores its results into rank 0's buffer over NVLink peer memory and bumps a "
                                        "signal; rank 0 waits for all signals per step (no NCCL on the data path)",
                                "nccl": "one all_gather_into_tensor of the flat (det|count) buffer per step on its own stream",
                                "none": "no gather (--no-gather control)"}[loop.x.mode if loop.x is not None else "none"],
                       "fallback_reason": getattr(loop.x, "why", "") if loop.x is not None else ""}
        if loop.x is not None and not args.skip_e2e:
            # control: the same loop without any gather (what the exchange costs)
            ctl = ResidentLoop(eng, args, B, S, dev, rank, world, "none")
            ctl_ms, _, _, _ = measure_value(ctl, args.steps, min(args.warmup, 3), 3)
            gather_info["no_gather_ms_per_step"] = ctl_ms
            del ctl
    if args.skip_e2e:
        if rank == 0 and args.per_op:
            geoms = [letterbox_geometry(S, S, (S, S), True)] * B
            eng.preprocess_images(net, list(loop.dev_batches[0]), geoms)
            per_op = net.run_timed(stream.cuda_stream)
            per_op = net.run_timed(stream.cuda_stream)
            for m, o, v in sorted(zip(per_op, net.ops, net.variants()), key=lambda r: -r[0]):
                tf = o.flops / (m / 1e3) / 1e12 if m > 0 else 0
                gb = o.bytes_algo / (m / 1e3) / 1e9 if m > 0 else 0
                var = f"  [{('tma', 'lsu', 'halo-tma', 'halo-stream')[v[0]]} {'epiW' if v[1] & 1 else 'epiC'}{'-fat' if v[1] & 2 else ''}{'-bres' if v[1] & 4 else ''}{'-pair' if v[1] & 8 else ''} {v[2]}cta/SM BN{v[3]}]" if v[2] > 0 else ""
                print(f"{m:8.4f} ms  {o.kind:9s} {o.name:28s} {tf:8.1f} TFLOP/s {gb:8.1f} GB/s(algo){var}", file=sys.stderr)
            print(f"network total {sum(per_op):.3f} ms; step {ms_step:.3f} ms", file=sys.stderr)
        if rank == 0:
            print(json.dumps({"profiling_only": True, "value": value, "ms_per_step": ms_step, "ms_per_step_blocks": blocks,
                              "launches_per_step": loop.launches, "mean_candidates_per_image": mean_cand}))
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- e2e through the public API from pinned host frames (H2D + D2H inside the timed region) ----
    host_batches = loop.host_batches
    if world == 1:
        def step_e2e(i: int):
            res = eng.predict(host_batches[i % loop.NROT], conf=CONF, iou=IOU, max_det=MAX_DET, imgsz=S, verbose=False)
            first = res[0].cpu().boxes.data                 # a Results as the reference's callers take it ...
            return res.det_host, res.counts, first          # ... and the whole batch's host rows (the call's one D2H)
        e2e_api = ("YOLO.predict(pinned uint8 [B,H,W,3] host tensor) -> sequence of Results (built on access; the batch's host rows "
                   "are read as one tensor); H2D in 4 chunks overlapped with layers 0-4 of the previous chunk, one D2H of all results")
    else:
        from yolo_infer_b200.parallel import ShardedPredictor
        sp = ShardedPredictor(eng, B, S, S, S, True, CONF, IOU, MAX_DET, mode=args.gather)

        def step_e2e(i: int):
            res = sp.predict(host_batches[i % loop.NROT])   # rank 0: Results of the GLOBAL batch (world * B images)
            return (res.det_host, res.counts, res[0].boxes.data) if len(res) else ()
        e2e_api = (f"parallel.ShardedPredictor.predict(pinned uint8 [B,H,W,3] local shard) on every rank; rank 0 returns the Results of "
                   f"all {world * B} images (gather = {sp.x.mode}) inside the timed region")
    for i in range(min(args.warmup, 3)):
        step_e2e(i)
    barrier()
    e2e_steps = max(3, min(args.steps, 10))
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        out = step_e2e(i)
    barrier()
    e2e_ms = 1e3 * (time.perf_counter() - t0) / e2e_steps
    e2e_sync_value = None
    if world == 1:
        # The same API in stream mode - `for results in model.predict(batches, stream=True)` - keeps two batches in flight: the
        # upload of batch i+1 and the read-back of batch i-1 overlap the compute of batch i.  Every step's H2D and D2H are still
        # inside the timed region; the one-call-at-a-time figure above is kept as `sync_value`.
        e2e_sync_value = B / (e2e_ms / 1e3)
        list(eng.predict((host_batches[i % loop.NROT] for i in range(3)), stream=True, conf=CONF, iou=IOU, max_det=MAX_DET, imgsz=S, verbose=False))
        barrier()
        n_stream = max(args.steps, 10)     # exactly --steps batches (pipeline fill and drain included in the timed region)
        t0 = time.perf_counter()
        for res in eng.predict((host_batches[i % loop.NROT] for i in range(n_stream)), stream=True, conf=CONF, iou=IOU, max_det=MAX_DET,
                               imgsz=S, verbose=False):
            out = (res.det_host, res.counts, res[0].cpu().boxes.data)
        barrier()
        e2e_ms = 1e3 * (time.perf_counter() - t0) / n_stream
        e2e_api = ("for results in YOLO.predict(iterable of pinned uint8 [B,H,W,3] host batches, stream=True): two batches in flight "
                   "(chunked H2D of batch i+1 and the result D2H of batch i-1 under the compute of batch i); per batch one Results taken + "
                   "the batch's host rows read")
        e2e_steps = n_stream
    t2 = torch.tensor([e2e_ms], device=dev)
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_value = world * B / (float(t2) / 1e3)
    d2h = (out[0].numel() * 4 + len(out[1]) * 4) if out else 0
    if world > 1:
        t3 = torch.tensor([float(d2h)], device=dev)
        dist.all_reduce(t3, op=dist.ReduceOp.MAX)
        d2h = int(t3)

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": f"YOLO11{args.model} detect, {S}x{S}, batch {B}/GPU, synthetic uint8 frames, random-init "
                                   f"variance-conditioned weights (BASELINE.json configs[{ {'n': 1, 's': 2}.get(args.model, '-') }])",
                       "global_batch": B * world, "parallelism": f"image-sharded x{world}",
                       "conf": CONF, "iou": IOU, "max_det": MAX_DET, "mean_candidates_per_image": mean_cand,
                       "mean_detections_per_image": mean_det, "steps_in_flight": loop.NS,
                       "l2_policy": f"{loop.NROT} rotating input batches ({loop.NROT * B * S * S * 3 / 1e6:.0f} MB) + "
                                    f"{sum(b.numel() * b.element_size() for b in net.buffers) / 1e9:.2f} GB of activations per step (> 126 MB L2)"},
            "ms_per_step_blocks": blocks, "ms_per_step_spread": (max(blocks) - min(blocks)) / ms_step,
            "value_best_block": world * B / (min(blocks) / 1e3),
            "ms_per_step_per_rank": ranks_ms,
            "timing": f"median of {len(blocks)} blocks of exactly {args.steps} steps (barrier + synchronize on both sides, CUDA events, max over ranks)",
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * S * S * 3, "d2h_bytes_per_step": d2h,
                    "api": e2e_api, "steps": e2e_steps, "sync_value": e2e_sync_value,
                    "sync_api": "one YOLO.predict(batch) call at a time (no overlap between calls)" if e2e_sync_value else None},
            "gpu_launches": loop.launches * args.steps,
            "cuda_graph": not args.no_graph,
            "gather": gather_info}

    if world == 1 or rank == 0:
        # ---- roofline of the dominant kernel (conv_tc_kernel): from the timed loop, serialised per-op pass as a secondary key ----
        geoms = [letterbox_geometry(S, S, (S, S), True)] * B
        per_op = None
        for _ in range(3):
            eng.preprocess_images(net, list(loop.dev_batches[0]), geoms)
            ms = net.run_timed(stream.cuda_stream)
            per_op = ms if per_op is None else [a + b for a, b in zip(per_op, ms)]
        per_op = [m / 3 for m in per_op]
        conv_ms = sum(m for m, o in zip(per_op, net.ops) if o.kind == "conv")
        conv_flops = sum(o.flops for o in net.ops if o.kind == "conv")
        conv_bytes = sum(o.bytes_algo for o in net.ops if o.kind == "conv")
        all_ms = sum(per_op)
        n_conv = sum(1 for o in net.ops if o.kind == "conv")
        step_ms_rank0 = ranks_ms[0]
        achieved = conv_flops / (step_ms_rank0 / 1e3) / 1e12        # conv FLOPs of ONE step / the timed loop's ms per step
        ser = conv_flops / (conv_ms / 1e3) / 1e12 if conv_ms > 0 else 0.0
        ser_gbs = conv_bytes / (conv_ms / 1e3) / 1e9 if conv_ms > 0 else 0.0
        traffic = traffic_src = None
        for tf in sorted((ROOT / "profiles").glob("r*_traffic.json"), reverse=True):   # newest capture whose launch count matches
            if B == 64 and S == 640 and traffic is None:
                t = json.loads(tf.read_text()).get(args.model)
                if t and t.get("conv_tc_launches") == n_conv:
                    traffic, traffic_src = t["dram_read_bytes"] + t["dram_write_bytes"], tf.name
        line["roofline"] = {
            "bound": "tensor", "achieved": achieved, "peak": pk["tflops_sustained"], "unit": "TFLOP/s",
            "frac": achieved / pk["tflops_sustained"], "frac_of_burst": achieved / pk["tflops_burst"], "peak_burst": pk["tflops_burst"],
            "traffic": traffic, "traffic_source": f"static, from profiles/{traffic_src} (ncu dram bytes of the conv launches of one step)"
                                                  if traffic_src else None,
            "kernel": f"conv_tc_kernel ({n_conv} launches/step); achieved = {conv_flops / 1e9:.1f} GFLOP of dense convs per step / "
                      f"{step_ms_rank0:.3f} ms per step of the timed loop (whole path: preprocess + network + decode + NMS)",
            "peak_source": f"{pk['source']}: sustained bf16 {pk['tflops_sustained']}, burst {pk['tflops_burst']} TFLOP/s",
            "algorithmic_flops_per_image": conv_flops / B, "algorithmic_bytes": conv_bytes,
            "hbm_side": {"achieved": conv_bytes / (step_ms_rank0 / 1e3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                         "frac": conv_bytes / (step_ms_rank0 / 1e3) / 1e9 / pk["hbm_gbs"],
                         "note": "unfused in+weights+out bytes of the same launches per step of the timed loop"},
            "serialised": {"achieved": ser, "frac": ser / pk["tflops_sustained"], "conv_ms": conv_ms, "network_ms": all_ms,
                           "hbm_gbs": ser_gbs, "note": "CUDA events around every launch of one extra pass (cold, no overlap between ops)"}}
        # ---- pre / post-processing kernels alone, against the HBM roof (north_star: achieved GB/s of letterbox, decode, NMS) ----
        line["roofline_pre"], line["roofline_post"] = pre_post_rooflines(eng, net, loop, B, S, pk, mean_cand, mean_det)
        if args.per_op:
            for l, h in enumerate(net.head):
                cls_l, box_l = h[..., 64:64 + 80].float(), h[..., :64].float()
                print(f"head level {l}: cls mean {cls_l.mean():.3f} std {cls_l.std():.3f} max {cls_l.max():.3f}; box std {box_l.std():.3f}",
                      file=sys.stderr)
            rows = sorted(zip(per_op, net.ops, net.variants()), key=lambda r: -r[0])
            for m, o, v in rows[:70]:
                tf = o.flops / (m / 1e3) / 1e12 if m > 0 else 0
                gb = o.bytes_algo / (m / 1e3) / 1e9 if m > 0 else 0
                var = f"  [{('tma', 'lsu', 'halo-tma', 'halo-stream')[v[0]]} {'epiW' if v[1] & 1 else 'epiC'}{'-fat' if v[1] & 2 else ''}{'-bres' if v[1] & 4 else ''}{'-pair' if v[1] & 8 else ''} {v[2]}cta/SM BN{v[3]}]" if v[2] > 0 else ""
                print(f"{m:8.4f} ms  {o.kind:9s} {o.name:28s} {tf:8.1f} TFLOP/s {gb:8.1f} GB/s(algo){var}", file=sys.stderr)
            print(f"network total {all_ms:.3f} ms; conv {conv_ms:.3f} ms; step {ms_step:.3f} ms", file=sys.stderr)

    # ---- batch-1 latency (BASELINE metric's second half): one frame, CUDA-graph replay, p50/p99 over N samples ----
    if args.latency_iters > 0 and rank == 0:
        line["latency_b1"] = latency_b1(eng, S, stream, dev, args.latency_iters)
    # ---- the reference harness's own input: float [B,3,S,S] tensors through predict (LoadTensor semantics) ----
    if rank == 0 and world == 1:
        xt = torch.rand((B, 3, S, S), generator=torch.Generator().manual_seed(7)).to(dev)
        for _ in range(3):
            eng.predict(xt, conf=CONF, iou=IOU, max_det=MAX_DET, verbose=False)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for _ in range(5):
            eng.predict(xt, conf=CONF, iou=IOU, max_det=MAX_DET, verbose=False)
        torch.cuda.synchronize(dev)
        line["tensor_source"] = {"value": B * 5 / (time.perf_counter() - t0), "unit": UNIT,
                                 "what": "YOLO.predict(device float [B,3,S,S]) as benchmarks/speed_benchmark.py:100-102 feeds it: device-side "
                                         "max -> /255 rule, one CUDA graph, one D2H of the results"}

    # ---- the other scales of the metric (value only, bounded) ----
    if world == 1 and args.extras and not args.no_graph:
        del loop
        for sc in [x for x in args.extras.split(",") if x and x != args.model]:
            try:
                e2 = make_engine(sc)
                l2 = ResidentLoop(e2, args, B, S, dev, rank, world, "none")
                ms2, bl2, _, _ = measure_value(l2, min(args.steps, 10), 3, 3)
                line[f"value_{sc}"] = {"value": B / (ms2 / 1e3), "unit": UNIT, "ms_per_step": ms2, "ms_per_step_blocks": bl2,
                                       "conv_frac_of_sustained": (B / (ms2 / 1e3)) * CONV_GFLOP[sc] * 1e9 / (pk["tflops_sustained"] * 1e12),
                                       "workload": f"YOLO11{sc} detect, {S}x{S}, batch {B}, same loop as `value` (3 blocks of {min(args.steps, 10)} steps)"}
                del l2, e2
                torch.cuda.empty_cache()
            except Exception as e:   # never lose the headline line to an extra
                line[f"value_{sc}"] = {"error": f"{type(e).__name__}: {e}"}

    # ---- opt-in FP8 (e4m3) mode of the same engine (quant.py): value only, NOT the headline (north_star numerics are bf16) ----
    if world == 1 and args.extras and not args.no_graph:
        try:
            from yolo_infer_b200.quant import calibrate_activation_scales
            hb = synth_frames(B, S, 4321).to(dev)
            scales = calibrate_activation_scales(eng, [hb[:8]])
            eng.enable_fp8(scales)
            l8 = ResidentLoop(eng, args, B, S, dev, rank, world, "none")
            ms8, bl8, _, _ = measure_value(l8, min(args.steps, 10), 3, 3)
            line["value_fp8"] = {"value": B / (ms8 / 1e3), "unit": UNIT, "ms_per_step": ms8, "ms_per_step_blocks": bl8, "e4m3_convs": len(scales),
                                 "workload": f"YOLO11{args.model}, same loop as `value`, with the {len(scales)} single-reader hidden tensors "
                                             "(Bottleneck, Detect box towers) stored as e4m3 and their consumers on tcgen05.mma.kind::f8f6f4; "
                                             "opt-in (create_quantizer('fp8', model).optimize(loader)), parity vs the quantised oracle in tests/test_gpu_fp8.py"}
            del l8
            eng.disable_fp8()
        except Exception as e:
            line["value_fp8"] = {"error": f"{type(e).__name__}: {e}"}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sd = {k: v.cpu() for k, v in eng.model.state_dict().items()}
        ips, ms, cores, sample = cpu_reference_arm(args, sd, steps=2, warmup=1, batch=min(16, B))
        line["cpu_baseline"] = {"value": ips, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
    else:
        line["cpu_baseline"] = None
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def latency_b1(eng, S, stream, dev, iters):
    import torch
    frame_h = synth_frames(1, S, 4242).pin_memory()
    p1 = eng.pipeline(1, S, S, S, True, CONF, IOU, MAX_DET)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for _ in range(10):
        p1.run(frame_h)
    torch.cuda.synchronize(dev)
    wall = []
    for a, b in evs:
        t0 = time.perf_counter()
        a.record(stream)
        det1, cnt1, _ = p1.run(frame_h)            # H2D of the frame + graph replay
        b.record(stream)
        int(cnt1.cpu())                            # D2H of the count = sync point, as a caller would do
        wall.append(1e3 * (time.perf_counter() - t0))
    dev_ms = sorted(a.elapsed_time(b) for a, b in evs)
    wall.sort()
    pct = lambda v, q: v[min(len(v) - 1, int(q * len(v)))]
    return {"batch": 1, "device_ms_p50": pct(dev_ms, 0.5), "device_ms_p99": pct(dev_ms, 0.99),
            "host_wall_ms_p50": pct(wall, 0.5), "host_wall_ms_p99": pct(wall, 0.99), "samples": len(evs),
            "what": "pinned uint8 frame -> H2D -> preprocess+forward+decode+NMS (one CUDA graph) -> count D2H"}


def pre_post_rooflines(eng, net, loop, B, S, pk, mean_cand, mean_det):
    """Letterbox (720p -> 384x640 resize, and the no-resize format conversion), decode and sort+NMS timed alone with CUDA events;
    algorithmic bytes per SURVEY.md 8(d)."""
    import ctypes as C
    import torch
    from yolo_infer_b200 import _cabi as cabi
    from yolo_infer_b200.engine import letterbox_geometry, scale_geometry
    dev = eng.device
    stream = torch.cuda.current_stream(dev)

    def time_letterbox(h0, w0):
        frames = synth_frames(B, S, 99, h0, w0).to(dev)
        g = letterbox_geometry(h0, w0, (S, S), True)
        H, W = g[4], g[5]
        out = torch.empty((B, H, W, 3), dtype=torch.bfloat16, device=dev)
        arr = (cabi.Image * B)()
        for i in range(B):
            f = frames[i]
            arr[i] = cabi.Image(f.data_ptr(), h0, w0, f.stride(0), g[1], g[0], g[2], g[3])
        desc = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(dev)
        flush = torch.empty((256 << 20,), dtype=torch.uint8, device=dev)
        best = []
        for _ in range(5):
            flush.zero_()                      # > L2: the frames come from HBM
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            cabi.check(eng._lib.y11_letterbox(eng._engine, desc.data_ptr(), B, H, W, out.data_ptr(), C.c_void_p(stream.cuda_stream)), "y11_letterbox")
            e1.record(stream)
            torch.cuda.synchronize(dev)
            best.append(e0.elapsed_time(e1))
        ms = statistics.median(best)
        nbytes = B * (h0 * w0 * 3 + H * W * 3 * 2)
        return {"kernel": "letterbox_kernel", "src": f"{h0}x{w0}", "dst": f"{H}x{W}", "ms": ms, "bytes": nbytes,
                "achieved": nbytes / (ms / 1e3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": nbytes / (ms / 1e3) / 1e9 / pk["hbm_gbs"]}

    pre = {"bound": "hbm", "resize_720p": time_letterbox(720, 1280), "same_size": time_letterbox(S, S),
           "note": "in the headline loop frames are already at network resolution and the stem reads them itself (no letterbox launch)"}
    # post: the head of the last value step is still in the net of pipeline 0's replica
    pnet = loop.pipes[0].net
    gain, px, py = scale_geometry((S, S), (S, S))
    rows = torch.tensor([[gain, px, py, S, S]] * B, dtype=torch.float32, device=dev)
    dms, nms_ = [], []
    for _ in range(5):
        a, b = eng.postprocess_timed(pnet, rows, CONF, IOU, MAX_DET)
        dms.append(a)
        nms_.append(b)
    d_ms, n_ms = statistics.median(dms), statistics.median(nms_)
    A = pnet.A
    K = mean_cand
    fused_cls = pnet.emit_conf is not None
    if fused_cls:   # class-emit conv epilogues: the decode kernel reads one 16-byte list entry + the DFL logits per listed anchor
        dec_bytes = B * K * (16 + 64 * 4 + 28)
    else:
        dec_bytes = B * (80 * A * 4 + K * (64 * 4 + 28))      # class logits of every anchor once + DFL logits and one row per candidate
    nms_bytes = B * (K * 32 + mean_det * (24 + 28))             # candidate rows + kept rows (the bit masks live in shared memory)
    post = {"bound": "hbm",
            "decode": {"kernel": "decode_list_kernel (class maximum + conf pre-filter run in the cv3.l.2 conv epilogues)" if fused_cls
                       else "decode_onepass_kernel", "ms": d_ms, "bytes": dec_bytes, "achieved": dec_bytes / (d_ms / 1e3) / 1e9,
                       "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": dec_bytes / (d_ms / 1e3) / 1e9 / pk["hbm_gbs"]},
            "sort_nms": {"kernel": "sort_nms_kernel", "ms": n_ms, "bytes": nms_bytes, "achieved": nms_bytes / (n_ms / 1e3) / 1e9,
                         "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": nms_bytes / (n_ms / 1e3) / 1e9 / pk["hbm_gbs"],
                         "note": f"latency bound by design: one CTA per image, ~{K:.0f} candidates -> bitonic sort + greedy sweep; "
                                 "its bytes are negligible, the number to read is the ms"},
            "candidates_per_image": K}
    return pre, post


def main():
    args = parse_args()
    # stdout carries exactly ONE JSON line: everything else that lands on file descriptor 1 while the run lasts (NCCL's version
    # banner is printed by the C library, past sys.stdout) is sent to stderr, and the descriptor is restored for the line itself
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    buf = []
    import builtins
    real_print = builtins.print

    def capture(*a, **k):
        if k.get("file") in (None, sys.stdout):
            buf.append(" ".join(str(x) for x in a))
        else:
            real_print(*a, **k)
    builtins.print = capture
    try:
        if args.impl == "reference":
            run_reference(args)
        else:
            run_ours(args)
    finally:
        builtins.print = real_print
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)
        for line in buf:
            real_print(line, flush=True)


if __name__ == "__main__":
    main()
