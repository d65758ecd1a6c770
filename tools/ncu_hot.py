#!/usr/bin/env python
"""Top sampled SASS lines per function of one kernel in an ncu report.
usage: python tools/ncu_hot.py rep.ncu-rep <kernel-index> [topN]"""
import csv, io, subprocess, sys
rep, ki = sys.argv[1], int(sys.argv[2]); topn = int(sys.argv[3]) if len(sys.argv) > 3 else 12
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(sass)))
i = 0; blocks = []
while i < len(rows):
    if rows[i] and rows[i][0] == 'Kernel Name':
        name = rows[i][1]; h = rows[i + 1]; j = i + 2; body = []
        while j < len(rows) and not (rows[j] and rows[j][0] == 'Kernel Name'):
            if len(rows[j]) == len(h): body.append(rows[j])
            j += 1
        blocks.append((name, h, body)); i = j
    else: i += 1
name, h, body = blocks[ki]
print(name[:100])
iS = h.index('# Samples'); iI = h.index('Instructions Executed')
tot = sum(int(r[iS] or 0) for r in body)
funcs = []; cur = []
for r in body:
    cur.append(r); t = r[1].strip()
    if t.startswith('RET') or t.startswith('EXIT'): funcs.append(cur); cur = []
for fi, fn in enumerate(funcs):
    s = sum(int(r[iS] or 0) for r in fn)
    if s * 50 < tot: continue
    print(f"--- function {fi}: {len(fn)} sass lines, {100*s/tot:.1f}% of samples, {sum(int(r[iI] or 0) for r in fn)} warp-instr")
    top = sorted(range(len(fn)), key=lambda k: -int(fn[k][iS] or 0))[:topn]
    for k in sorted(top):
        r = fn[k]
        print(f"   {k:5d} {100*int(r[iS])/tot:5.2f}%  x{r[iI]:>9s}  {r[1].strip()[:70]:70s} <- prev: {fn[k-1][1].strip()[:40] if k else ''}")
