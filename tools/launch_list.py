"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table."""
import csv
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 10]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
ui = hdr.index("Metric Unit")
agg = OrderedDict()
for r in rows[1:]:
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    if r[ui] in ("ns", "nsecond"):
        v /= 1000.0
    elif r[ui] in ("ms", "msecond"):
        v *= 1000.0
    a = agg.setdefault(r[ki][:70], [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print("| kernel | launches | total us | share |\n|---|---:|---:|---:|")
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| {k} | {n} | {us:.1f} | {us / tot * 100:.1f}% |")
