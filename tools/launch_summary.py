#!/usr/bin/env python
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): per-kernel launches, total time, share.
usage: python tools/launch_summary.py gpurun_out/launches.csv [> profiles/rN_launches_<what>.txt]"""
import csv
import re
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0] != "ID"]
agg = OrderedDict()
for r in rows:
    name = re.sub(r"\(.*", "", r[4]).replace("void ", "").strip()
    if name.startswith("at::") or "at::native" in name:
        name = "(torch fill/copy helpers)"
    t = float(r[-1].replace(",", ""))
    unit = r[-2]
    t_us = t / 1e3 if unit in ("ns", "nsecond") else (t if unit in ("us", "usecond") else t * 1e3)
    a = agg.setdefault(name, [0, 0.0, r[7], r[8]])
    a[0] += 1
    a[1] += t_us
tot = sum(a[1] for a in agg.values())
print(f"{'kernel':60s} {'launches':>8s} {'total_us':>12s} {'avg_us':>10s} {'share':>7s}  block grid(last)")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:60]:60s} {a[0]:8d} {a[1]:12.1f} {a[1] / a[0]:10.1f} {100 * a[1] / tot:6.1f}%  {a[2]} {a[3]}")
print(f"{'total':60s} {sum(a[0] for a in agg.values()):8d} {tot:12.1f}")
