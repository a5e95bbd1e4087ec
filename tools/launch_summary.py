#!/usr/bin/env python
"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel.

    python tools/launch_summary.py gpurun_out/<tag>/launches.csv [skip_kernels_regex] > profiles/<name>.md

Times are cold-cache and serialised (ncu replays): compare SHARES, not absolutes.
"""
import csv
import re
import sys
from collections import OrderedDict

path = sys.argv[1]
skip = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = OrderedDict()
for r in rows[1:]:
    name = r[ik].split("(")[0].replace("void ", "")
    if skip and skip.search(name):
        continue
    t = float(r[iv]) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[iu], 1.0)
    c = agg.setdefault(name, [0, 0.0])
    c[0] += 1
    c[1] += t
tot = sum(v[1] for v in agg.values())
print(f"launch list: {path}  ({sum(v[0] for v in agg.values())} launches, {tot / 1e3:.3f} ms summed device time)\n")
print("| kernel | launches | total us | mean us | share |")
print("|---|---|---|---|---|")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {n} | {t:.1f} | {t / n:.2f} | {100 * t / tot:.1f}% |")
