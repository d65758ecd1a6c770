#!/usr/bin/env python
"""Condense qball_b200/build.log (nvcc -Xptxas -v): registers / spills / stack per kernel and per __noinline__ device function.
usage: python tools/ptxas_report.py [pattern]"""
import re, subprocess, sys
log = open("qball_b200/build.log").read().splitlines()
pat = sys.argv[1] if len(sys.argv) > 1 else ""
cur = None
out = []
for i, l in enumerate(log):
    m = re.search(r"Function properties for (\S+)", l)
    if m:
        cur = m.group(1)
        spill = log[i + 1].strip() if i + 1 < len(log) else ""
        used = ""
        for j in range(i + 1, min(i + 4, len(log))):
            if "Used" in log[j]:
                used = log[j].split(":", 1)[-1].strip(); break
            if "Function properties" in log[j] and j > i: break
        out.append((cur, spill, used))
names = subprocess.run(["c++filt"], input="\n".join(o[0] for o in out), capture_output=True, text=True).stdout.splitlines()
for (n, s, u), d in zip(out, names):
    d = re.sub(r"\(.*", "", d)
    if pat in d:
        sp = re.findall(r"(\d+) bytes", s)
        print(f"{d[:70]:70s} stack {sp[0]:>4s} spill st/ld {sp[1]:>4s}/{sp[2]:>4s}  {u[:60]}")
