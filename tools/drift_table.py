#!/usr/bin/env python
"""Per-layer drift of the B200 path against the same-storage oracle (GPU box only; writes a markdown table).

For every conv / depthwise op of the plan, after ONE forward pass of both implementations on the same input:
  drift      = rel-L2(GPU tensor, oracle tensor of the same layer)            - accumulated difference, amplified by the weights
  per-layer  = rel-L2(GPU op re-run alone on ITS OWN input, torch fp32 conv on that same input) - the kernel's own error
  control    = drift of the ORACLE AGAINST ITSELF when 1 % of the input pixels move by one bf16 ulp - what the weights do to ANY
               perturbation, i.e. the floor no implementation can beat
A kernel bug shows as a jump in `per-layer`; amplification shows as drift ~ control with flat per-layer.

  python tools/drift_table.py --scale n --init calibrated --out profiles/r02_drift_n_calibrated.md
"""
import argparse
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle import yolo11_ref as R  # noqa: E402
from yolo_infer_b200 import _cabi as cabi  # noqa: E402
from yolo_infer_b200.engine import YOLO  # noqa: E402


def rel(a, b):
    return float((a - b).norm() / b.norm().clamp_min(1e-12))


def capture(model, x):
    outs = {}
    hooks = []
    for name, m in model.named_modules():
        if isinstance(m, (R.Conv, torch.nn.Conv2d)) and name and not name.endswith(".conv") and "dfl" not in name:
            hooks.append(m.register_forward_hook(lambda mod, i, o, name=name: outs.__setitem__(name, o.detach())))
    with torch.no_grad():
        model(x)
    for h in hooks:
        h.remove()
    return outs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", default="n")
    ap.add_argument("--init", default="calibrated", choices=["calibrated", "survey_b"])
    ap.add_argument("--hw", type=int, default=640)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    ref = R.build(a.scale, init=a.init, seed=0)
    sd = {k: v.clone() for k, v in ref.state_dict().items()}

    def emul():
        m = R.DetectionModel(a.scale)
        m.load_state_dict(sd)
        return R.emulate_bf16_storage(m.eval().fuse(), fold_upsample=False)

    B, H, W = 2, a.hw, a.hw
    x = torch.rand(B, 3, H, W, generator=torch.Generator().manual_seed(17))
    xp = x.clone()
    pick = torch.rand(x.shape, generator=torch.Generator().manual_seed(18)) < 0.01
    xp[pick] = xp[pick].to(torch.bfloat16).float() * (1 + 2.0 ** -8)
    want = capture(emul(), x)
    ctrl = capture(emul(), xp)
    eng = YOLO.from_state_dict(sd, a.scale).to("cuda:0")
    net = eng.compiled(B, H, W, fold_upsample=False)    # one op per reference conv
    eng.preprocess_tensor(net, x.cuda().contiguous(), 1.0)
    eng.forward(net)
    torch.cuda.synchronize()
    s = torch.cuda.current_stream().cuda_stream
    rows = []
    for i, op in enumerate(net.ops):
        if op.kind not in ("conv", "dwconv", "stem") or op.name not in want:
            continue
        v = op.out
        got = v.t[..., v.off:v.off + v.c].float().permute(0, 3, 1, 2).cpu()
        if op.kind == "stem" and got.shape[1] != want[op.name].shape[1]:      # space-to-depth stem buffer
            c = want[op.name].shape[1]
            got = got.view(B, 4, c, got.shape[2], got.shape[3])
            if eng._packed["model.1"].s2d_block:       # permuted block order of the compact 2x2 form: back to dy*2+dx
                from yolo_infer_b200.network import S2D_PERM
                got = got[:, [S2D_PERM.index((dy, dx)) for dy in (0, 1) for dx in (0, 1)]]
            got = got.reshape(B, 2, 2, c, got.shape[3], got.shape[4]).permute(0, 3, 4, 1, 5, 2).reshape(B, c, 2 * got.shape[3], 2 * got.shape[4])
        w = want[op.name]
        overwritten = any(o2.out is not None and o2.out.t.data_ptr() == v.t.data_ptr() and o2.out.off < v.off + v.c and v.off < o2.out.off + o2.out.c
                          for o2 in net.ops[i + 1:])
        if op.res is not None and op.res_mode == cabi.RES_POST:
            drift = None      # the buffer holds conv + residual; the oracle's conv output does not: compared at the next layer
        elif overwritten or op.name.endswith("attn.qkv"):
            drift = None      # updated in place by a later residual op (C2PSA's b half) / channels stored in [Q|K|V] order
        else:
            drift = rel(got[:, :w.shape[1]], w)
        own = None
        if op.kind in ("conv", "dwconv") and op.inp is not None and not (op.res is not None and op.res.t.data_ptr() == v.t.data_ptr() and op.res.off == v.off):
            pc = eng._packed[op.name]
            vin = op.inp
            xin = vin.t[..., vin.off:vin.off + vin.c].float().permute(0, 3, 1, 2).contiguous()
            res = op.res.t[..., op.res.off:op.res.off + op.res.c].float().permute(0, 3, 1, 2).clone() if op.res is not None else None
            net.run_range(i, i + 1, s)
            torch.cuda.synchronize()
            g2 = v.t[..., v.off:v.off + v.c].float().permute(0, 3, 1, 2)
            if pc.depthwise:
                t = torch.nn.functional.conv2d(xin, pc.w.float().t().reshape(pc.c2, 1, 3, 3), pc.b, padding=1, groups=pc.c2)
            else:
                wt = pc.w.float().view(pc.c2, pc.k, pc.k, pc.c1).permute(0, 3, 1, 2)
                t = (torch.nn.functional.conv2d(torch.nn.functional.pad(xin, (1, 0, 1, 0)), wt, pc.b) if pc.k == 2
                     else torch.nn.functional.conv2d(xin, wt, pc.b, stride=pc.s, padding=pc.k // 2))
            if pc.act:
                t = torch.nn.functional.silu(t)
            if res is not None:
                t = t + res
            own = rel(g2, t)
        rows.append((op.name, drift, own, rel(ctrl[op.name], w)))
    lines = [f"# Drift of the B200 path vs the same-storage oracle: YOLO11{a.scale}, init `{a.init}`, {B}x{H}x{W} (tools/drift_table.py)", "",
             "| layer | drift GPU vs oracle | kernel alone (own input vs torch fp32) | control: oracle vs oracle, 1 % of pixels + 1 ulp |", "|---|---|---|---|"]
    for name, d, o, c in rows:
        lines.append(f"| {name} | {'-' if d is None else f'{d:.2e}'} | {'-' if o is None else f'{o:.2e}'} | {c:.2e} |")
    txt = "\n".join(lines) + "\n"
    if a.out:
        Path(a.out).write_text(txt)
    print(txt)


if __name__ == "__main__":
    main()
