#!/usr/bin/env python
"""Sampling share per barrier-delimited region of one kernel (phases of the plane kernels).
usage: python tools/ncu_phases.py rep.ncu-rep <kernel-index>"""
import csv, io, subprocess, sys
rep, ki = sys.argv[1], int(sys.argv[2])
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
print(name[:110])
iS = h.index('# Samples'); iI = h.index('Instructions Executed')
tot = sum(int(r[iS] or 0) for r in body)
reg = []; cur = dict(start=0, s=0, n=0, fp=0, lds=0, sts=0, ldg=0, stg=0, bar='')
for k, r in enumerate(body):
    t = r[1].strip(); s = int(r[iS] or 0); n = int(r[iI] or 0)
    cur['s'] += s; cur['n'] += n
    op = t.split()[0] if t else ''
    if op.startswith('@'): op = t.split()[1] if len(t.split()) > 1 else ''
    if op.startswith(('DADD', 'DMUL', 'DFMA')): cur['fp'] += n
    elif op.startswith('LDS'): cur['lds'] += n
    elif op.startswith('STS'): cur['sts'] += n
    elif op.startswith(('LDG', 'LD.')): cur['ldg'] += n
    elif op.startswith(('STG', 'ST.')): cur['stg'] += n
    if op.startswith('BAR'):
        cur['bar'] = t[:40]; cur['end'] = k; reg.append(cur)
        cur = dict(start=k + 1, s=0, n=0, fp=0, lds=0, sts=0, ldg=0, stg=0, bar='')
cur['end'] = len(body) - 1; reg.append(cur)
print(f"{'lines':>11s} {'samples%':>8s} {'warp-instr':>11s} {'fp64':>10s} {'LDS':>9s} {'STS':>9s} {'LDG':>8s} {'STG':>8s}  ends with")
for r in reg:
    if r['s'] * 200 < tot and r['n'] == 0: continue
    print(f"{r['start']:5d}-{r['end']:5d} {100*r['s']/tot:8.2f} {r['n']:11d} {r['fp']:10d} {r['lds']:9d} {r['sts']:9d} {r['ldg']:8d} {r['stg']:8d}  {r['bar']}")
