"""Join an ncu per-launch CSV (dram__bytes_read.sum, dram__bytes_write.sum, gpu__time_duration.sum of the conv_tc_kernel
launches of ONE pass) with the plan's op list (bench.py --dump-ops): per-layer DRAM traffic next to the algorithmic bytes.

  python tools/ncu_join.py ops.json ncu.csv [first_launch_index] > table.md
Also prints the totals bench.py reads from profiles/*_traffic.json."""
import csv
import io
import json
import sys

ops = [o for o in json.load(open(sys.argv[1])) if o["kind"] == "conv"]
text = open(sys.argv[2]).read()
text = text[text.index('"ID"'):]
rows = list(csv.DictReader(io.StringIO(text)))
skip = int(sys.argv[3]) if len(sys.argv) > 3 else 0
per = {}
for r in rows:
    if "conv_tc_kernel" not in r["Kernel Name"]:
        continue
    per.setdefault(int(r["ID"]), {})[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
ids = sorted(per)[skip:skip + len(ops)]
assert len(ids) == len(ops), (len(ids), len(ops))
print("| layer | variant | time us | DRAM read MB | DRAM write MB | algorithmic MB | DRAM/algo |")
print("|---|---|---|---|---|---|---|")
tr = tw = ta = tt = 0.0
for o, i in zip(ops, ids):
    m = per[i]
    rd, wr = m.get("dram__bytes_read.sum", 0.0), m.get("dram__bytes_write.sum", 0.0)
    t = m.get("gpu__time_duration.sum", 0.0)
    t_us = t / 1e3 if t > 1e3 else t
    v = o["variant"]
    var = f"{'lsu' if v[0] else 'tma'} {'epiW' if v[1] & 1 else 'epiC'}{'-fat' if v[1] & 2 else ''}{'-bres' if v[1] & 4 else ''}{'-pair' if v[1] & 8 else ''} {v[2]}cta BN{v[3]}"
    tr += rd; tw += wr; ta += o["bytes_algo"]; tt += t_us
    print(f"| {o['name']} | {var} | {t_us:.1f} | {rd / 1e6:.1f} | {wr / 1e6:.1f} | {o['bytes_algo'] / 1e6:.1f} | {(rd + wr) / o['bytes_algo']:.2f} |")
print(f"\ntotals: {len(ops)} conv_tc launches, {tt:.1f} us, DRAM read {tr / 1e9:.3f} GB + write {tw / 1e9:.3f} GB, algorithmic {ta / 1e9:.3f} GB")
print(json.dumps({"conv_tc_launches": len(ops), "dram_read_bytes": tr, "dram_write_bytes": tw, "algorithmic_bytes": ta}))
