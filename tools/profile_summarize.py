"""Turn the raw ncu output of tools/profile_round.sh (gpurun_out/<TAG>_*) into the committed evidence under profiles/:
launch-list shares, per-layer DRAM-vs-algorithmic table, traffic totals (read by bench.py), --set full highlights.

  python tools/profile_summarize.py r01e"""
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


out = [f"# {TAG}: ncu summary (commands: tools/profile_round.sh; B200, 1 GPU, batch 64, 640x640, `--clock-control none`)\n"]
traffic = {"what": "dram__bytes_read.sum + dram__bytes_write.sum summed over the conv_tc_kernel launches of ONE step (first eager "
                   f"pass), per model; raw rows profiles/{TAG}_conv_dram_M.csv, per-layer join profiles/{TAG}_layers_M.md"}
for m in "ns":
    rows = ncu_csv(G / f"{TAG}_launches_yolo11{m}_b64.csv")
    # one pass = from the first letterbox_kernel / stem_kernel (frames at network resolution need no letterbox launch) up to and
    # including the first sort_nms_kernel after it (what precedes is set-up: torch.zeros fills of the buffers, weight packing)
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
    out.append(f"## YOLO11{m}: launch list of one pass ({n_pass} launches, {tot:.1f} us serialised, cold caches)\n")
    out.append("| kernel | launches | us | share |\n|---|---|---|---|")
    for k, v in sorted(per.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| {k} | {v[0]} | {v[1]:.1f} | {100 * v[1] / tot:.1f} % |")
    out.append("")
    for f in (f"{TAG}_launches_yolo11{m}_b64.csv", f"{TAG}_conv_dram_{m}.csv", f"{TAG}_ops_{m}.json"):
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

rep = G / f"{TAG}_conv_s_full.ncu-rep"
if rep.exists():
    raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    (P / f"{TAG}_conv_tc_full_raw_yolo11s.csv").write_text(raw)
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, data = rows[0], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    ops = [o for o in json.loads((G / f"{TAG}_ops_s.json").read_text()) if o["kind"] == "conv"][: len(data)]
    cols = [("gpu__time_duration.sum", "time us"), ("dram__bytes_read.sum", "DRAM rd MB"), ("dram__bytes_write.sum", "DRAM wr MB"),
            ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % peak"),
            ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
            ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
            ("smsp__pcsamp_warps_issue_stalled_long_scoreboard", "stall long_sb"), ("smsp__pcsamp_warps_issue_stalled_wait", "stall wait"),
            ("smsp__pcsamp_warps_issue_stalled_short_scoreboard", "stall short_sb"), ("smsp__pcsamp_warps_issue_stalled_barrier", "stall barrier"),
            ("smsp__pcsamp_warps_issue_stalled_mio_throttle", "stall mio"), ("smsp__pcsamp_warps_issue_stalled_selected", "selected"),
            ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid")]
    cols = [(c, t) for c, t in cols if c in idx]
    out.append(f"## conv_tc_kernel, `ncu --set full`, first {len(data)} conv launches of a YOLO11s step\n")
    out.append("| layer | " + " | ".join(t for _, t in cols) + " |\n|---|" + "---|" * len(cols))
    for o, d in zip(ops, data):
        out.append(f"| {o['name']} | " + " | ".join(d[idx[c]][:8] for c, _ in cols) + " |")
    out.append("")
(P / f"{TAG}_summary.md").write_text("\n".join(out) + "\n")
print("\n".join(out))
