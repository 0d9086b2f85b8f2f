"""Compare two `bench.py --per-op` stderr tables op by op: python tools/perop_diff.py A.err B.err"""
import re
import sys


def load(path):
    rows = {}
    for line in open(path):
        m = re.match(r"\s+([\d.]+) ms\s+(\w+)\s+(\S+)(.*)", line)
        if m:
            v = re.search(r"\[(.*)\]", m.group(4))
            rows[m.group(3) + ":" + m.group(2)] = (float(m.group(1)), v.group(1) if v else "")
    return rows


a, b = load(sys.argv[1]), load(sys.argv[2])
ta = tb = 0.0
for k in sorted(set(a) & set(b), key=lambda k: -(a[k][0])):
    ta += a[k][0]
    tb += b[k][0]
    d = b[k][0] - a[k][0]
    if abs(d) > 0.003:
        print(f"{k:40s} {a[k][0]:.4f} -> {b[k][0]:.4f} ({d:+.4f})  {a[k][1]} -> {b[k][1]}")
print(f"common ops: {ta:.3f} -> {tb:.3f} ms")
