"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, mean time, share."""
import csv
import re
import sys
from collections import OrderedDict

path = sys.argv[1]
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if l.startswith('"')]
r = csv.reader(lines)
hdr = next(r)
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
for row in r:
    rows.append((row[ki], float(row[vi].replace(",", ""))))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = rows[skip:]
agg = OrderedDict()
for name, ns in rows:
    short = re.sub(r"\(.*", "", name).replace("void ", "")
    short = re.sub(r"at::native::|at::", "", short)[:90]
    a = agg.setdefault(short, [0, 0.0])
    a[0] += 1; a[1] += ns
total = sum(a[1] for a in agg.values())
print(f"| kernel | launches | mean us | total us | share |\n|---|---:|---:|---:|---:|")
for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {n} | {ns / n / 1e3:.1f} | {ns / 1e3:.1f} | {100 * ns / total:.1f}% |")
print(f"\ntotal {total / 1e6:.3f} ms over {len(rows)} launches")
