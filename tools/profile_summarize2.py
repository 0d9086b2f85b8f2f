"""Turn the raw ncu output of tools/profile_round2.sh (gpurun_out/<TAG>_*) into the committed evidence under profiles/:
launch-list shares under load (conditioned weights: ~1300 candidates per image in decode / NMS), per-layer DRAM traffic
vs algorithmic bytes + traffic totals (read by bench.py), and the `--set full` per-layer table with the TENSOR-PIPE utilisation,
L2 and L2->SM crossbar utilisation of every conv of one YOLO11s and one YOLO11m step.

  python tools/profile_summarize2.py r02a"""
import csv
import io
import json
import subprocess
import sys
from collections import defaultdict
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
TAG = sys.argv[1]
G, P = ROOT / "gpurun_out", ROOT / "profiles"


def ncu_csv(path):
    text = Path(path).read_text()
    return list(csv.DictReader(io.StringIO(text[text.index('"ID"'):])))


out = [f"# {TAG}: ncu summary (commands: tools/profile_round2.sh / profile_round2b.sh; B200, 1 GPU, batch 64, 640x640, `--clock-control none`, conditioned "
       "synthetic weights, kernels launched one by one in plan order)\n"]
traffic = {"what": "dram__bytes_read.sum + dram__bytes_write.sum summed over the conv_tc_kernel launches of ONE step, per model; raw rows "
                   f"profiles/{TAG}_conv_dram_M.csv, per-layer join profiles/{TAG}_layers_M.md"}
for m in "sn":
    if not (G / f"{TAG}_launches_yolo11{m}_b64.csv").exists():     # tools/profile_round2b.sh profiles YOLO11s only
        continue
    rows = ncu_csv(G / f"{TAG}_launches_yolo11{m}_b64.csv")
    first = next(i for i, r in enumerate(rows) if "letterbox" in r["Kernel Name"] or "stem_kernel" in r["Kernel Name"])
    per, n_pass = defaultdict(lambda: [0, 0.0]), 0
    for r in rows[first:]:
        name = r["Kernel Name"].split("(")[0].replace("<unnamed>::", "").replace("void ", "")
        if "at::native" in name or "elementwise" in name or "fill" in name.lower():
            name = "(torch fill/copy)"
        t = float(r["Metric Value"].replace(",", ""))
        t_us = t / 1e3 if r["Metric Unit"].startswith("n") else t
        per[name][0] += 1
        per[name][1] += t_us
        n_pass += 1
        if "sort_nms" in name:
            break
    tot = sum(v[1] for v in per.values())
    out.append(f"## YOLO11{m}: launch list of one pass UNDER LOAD ({n_pass} launches, {tot:.1f} us serialised, cold caches)\n")
    out.append("| kernel | launches | us | share |\n|---|---|---|---|")
    for k, v in sorted(per.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| {k} | {v[0]} | {v[1]:.1f} | {100 * v[1] / tot:.1f} % |")
    out.append("")
    for f in (f"{TAG}_launches_yolo11{m}_b64.csv", f"{TAG}_ops_{m}.json"):
        (P / f).write_bytes((G / f).read_bytes())
    j = subprocess.run([sys.executable, str(ROOT / "tools" / "ncu_join.py"), str(G / f"{TAG}_ops_{m}.json"),
                        str(G / f"{TAG}_conv_dram_{m}.csv")], capture_output=True, text=True, check=True).stdout
    (P / f"{TAG}_layers_{m}.md").write_text(f"# {TAG}: YOLO11{m} conv_tc_kernel launches of one step - DRAM traffic (ncu) vs algorithmic bytes\n\n" + j)
    tot_line = json.loads(j.strip().splitlines()[-1])
    traffic[m] = tot_line
    out.append(f"conv_tc_kernel DRAM traffic of the step: read {tot_line['dram_read_bytes'] / 1e9:.2f} GB + write "
               f"{tot_line['dram_write_bytes'] / 1e9:.2f} GB vs {tot_line['algorithmic_bytes'] / 1e9:.2f} GB algorithmic "
               f"(unfused in+weights+out) - per layer in `{TAG}_layers_{m}.md`.\n")
(P / f"{TAG}_traffic.json").write_text(json.dumps(traffic, indent=1))

COLS = [("gpu__time_duration.sum", "us"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
        ("l1tex__m_xbar2l1tex_read_bytes.sum.pct_of_peak_sustained_elapsed", "L2->SM rd %"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps act %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid")]


def full_table(model, title, only=None):
    f = G / f"{TAG}_conv_{model}_full_raw.csv"
    if not f.exists():
        return
    rows = list(csv.reader(open(f)))
    hdr, data = rows[0], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    ops = [o for o in json.loads((G / f"{TAG}_ops_{model}.json").read_text()) if o["kind"] == "conv"][: len(data)]
    cols = [(c, t) for c, t in COLS if c in idx]
    out.append(f"## {title}\n")
    out.append("| layer | variant | TFLOP/s | " + " | ".join(t for _, t in cols) + " |\n|---|---|---|" + "---|" * len(cols))
    keep_cols = [0] + [idx[c] for c, _ in cols]
    slim = [["layer"] + [c for c, _ in cols]]
    for o, d in zip(ops, data):
        v = o["variant"]
        var = f"{('tma', 'lsu', 'halo-tma', 'halo-stream')[v[0]]} {'epiW' if v[1] & 1 else 'epiC'}{'-fat' if v[1] & 2 else ''}{'-bres' if v[1] & 4 else ''}{'-pair' if v[1] & 8 else ''} {v[2]}cta BN{v[3]}"
        us = float(d[idx["gpu__time_duration.sum"]])
        tf = o["flops"] / (us * 1e-6) / 1e12
        slim.append([o["name"]] + [d[idx[c]] for c, _ in cols])
        if only is None or only(o, us):
            out.append(f"| {o['name']} | {var} | {tf:.0f} | " + " | ".join(f"{float(d[idx[c]]):.1f}" if "." in d[idx[c]] else d[idx[c]] for c, _ in cols) + " |")
    out.append("")
    with open(P / f"{TAG}_conv_{model}_full.csv", "w", newline="") as fh:
        csv.writer(fh).writerows(slim)


full_table("s", "conv_tc_kernel, `ncu --set full`, EVERY conv launch of one YOLO11s step (tensor-pipe, L2 and L2->SM crossbar utilisation)")
full_table("m", "conv_tc_kernel, `ncu --set full`, YOLO11m step: the launches above 60 us (the compute-bound 3x3 layers)", only=lambda o, us: us >= 60.0)

f = G / f"{TAG}_other_s_full_raw.csv"
if f.exists():
    rows = list(csv.reader(open(f)))
    hdr, data = rows[0], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [(c, t) for c, t in COLS if c in idx]
    name_i = idx.get("Kernel Name", 4)
    out.append("## The other kernels of a YOLO11s step under load (`--set full`): stem, depthwise, SPPF, attention, decode, sort+NMS\n")
    out.append("| kernel | " + " | ".join(t for _, t in cols) + " | DRAM rd MB | DRAM wr MB |\n|---|" + "---|" * (len(cols) + 2))
    for d in data:
        nm = d[name_i].split("(")[0].replace("<unnamed>::", "").replace("void ", "")
        out.append(f"| {nm} | " + " | ".join(f"{float(d[idx[c]]):.1f}" if "." in d[idx[c]] else d[idx[c]] for c, _ in cols)
                   + f" | {float(d[idx['dram__bytes_read.sum']]):.1f} | {float(d[idx['dram__bytes_write.sum']]):.1f} |")
    out.append("")
(P / f"{TAG}_summary.md").write_text("\n".join(out) + "\n")
print("\n".join(out))
