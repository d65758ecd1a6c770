#!/usr/bin/env python
"""DRAM traffic per launch of the hot kernels from ncu --set full reports -> profiles/ncu_traffic.json (read by bench.py for
roofline.traffic).  usage: python tools/ncu_traffic.py rep1.ncu-rep [rep2.ncu-rep ...] > profiles/ncu_traffic.json"""
import csv, io, json, re, subprocess, sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
out = {}
for rep in sys.argv[1:]:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[h.index("Kernel Name")]
        m = re.match(r"(?:void )?(?:qb200::)?(\w+)(?:<\(?(?:int\))?(-?\d+))?", name)
        key = m.group(1) + (f"<{m.group(2)}>" if m.group(2) is not None else "")
        tot = 0.0
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            tot += float(r[h.index(k)].replace(",", "")) * UNIT[units[h.index(k)]]
        e = out.setdefault(key, {"dram_bytes_per_launch": 0.0, "launches": 0, "grid": int(r[h.index("launch__grid_size")].replace(",", "")),
                                 "source": [], "fp64_pipe_active_pct": 0.0, "tensor_pipe_active_pct": 0.0, "smem_wavefronts_per_cycle_per_sm": 0.0,
                                 "duration_ms": 0.0})
        e["dram_bytes_per_launch"] += tot
        e["launches"] += 1

        def num(k):
            return float(r[h.index(k)].replace(",", "") or 0) if k in h else 0.0

        e["fp64_pipe_active_pct"] += num("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active")
        e["tensor_pipe_active_pct"] += num("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
        cyc, nsm = num("sm__cycles_elapsed.max"), 148.0
        if cyc > 0:
            e["smem_wavefronts_per_cycle_per_sm"] += num("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum") / (cyc * nsm)
        dur = num("gpu__time_duration.sum")
        du = units[h.index("gpu__time_duration.sum")] if "gpu__time_duration.sum" in h else "ms"
        e["duration_ms"] += dur * {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}.get(du, 1.0)
        if rep not in e["source"]:
            e["source"].append(rep)
for e in out.values():
    for k in ("dram_bytes_per_launch", "fp64_pipe_active_pct", "tensor_pipe_active_pct", "smem_wavefronts_per_cycle_per_sm", "duration_ms"):
        e[k] /= e["launches"]
print(json.dumps(out, indent=1))
