#!/usr/bin/env python
"""Summarise an ncu report (read on the CPU box): per-kernel headline metrics + per-function stall breakdown.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [> profiles/xyz.txt]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "sm__pipe_tensor_op_dmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor_op_dmma.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
for r in rows[2:]:
    print("=" * 100)
    for w in want:
        if w in hdr:
            print(f"{w:75s} {r[hdr.index(w)]} {rows[1][hdr.index(w)]}")
    st = sorted(((float(r[hdr.index(s)] or 0), s.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")) for s in stall), reverse=True)
    print("stall cycles per issued instruction:", ", ".join(f"{n}={v:.2f}" for v, n in st[:8]))

sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(sass)))
# several kernels: each starts with a "Kernel Name" row followed by a header row
i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == "Kernel Name":
        name = rows[i][1]
        h = rows[i + 1]
        j = i + 2
        body = []
        while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
            if len(rows[j]) == len(h):
                body.append(rows[j])
            j += 1
        iS, iI = h.index("# Samples"), h.index("Instructions Executed")
        st = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
        segs, cur = [], []
        for r in body:
            cur.append(r)
            t = r[1].strip()
            if t.startswith("RET") or t.startswith("EXIT"):
                segs.append(cur)
                cur = []
        if cur:
            segs.append(cur)
        tot = max(1, sum(int(r[iS] or 0) for r in body))
        print("-" * 100)
        print("per-function sampling for", name[:90])
        for s in segs:
            smp = sum(int(r[iS] or 0) for r in s)
            if smp < tot * 0.01:
                continue
            ops = {}
            for r in s:
                t = r[1].split()
                op = t[1] if t[0].startswith("@") and len(t) > 1 else t[0]
                ops[op] = ops.get(op, 0) + int(r[iI] or 0)
            fp64 = sum(v for k, v in ops.items() if k[:4] in ("DFMA", "DADD", "DMUL"))
            ins = sum(ops.values())
            stl = sorted(((sum(int(r[h.index(c)] or 0) for r in s), c[6:]) for c in st), reverse=True)[:5]
            print(f"  sass={len(s):5d} samples={100.0 * smp / tot:5.1f}% warp-instr={ins:10d} fp64={fp64:10d} LDS={ops.get('LDS.128', 0) + ops.get('LDS.64', 0):9d} "
                  f"STS={ops.get('STS.128', 0) + ops.get('STS.64', 0):9d} LDG={sum(v for k, v in ops.items() if k.startswith('LDG')):8d} "
                  f"local={sum(v for k, v in ops.items() if k.startswith('LDL') or k.startswith('STL')):8d} DMMA={sum(v for k, v in ops.items() if k.startswith('DMMA')):8d} | "
                  + ", ".join(f"{n}={100.0 * v / max(smp, 1):.0f}%" for v, n in stl))
        i = j
    else:
        i += 1
