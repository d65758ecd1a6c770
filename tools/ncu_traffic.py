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
                                 "source": []})
        e["dram_bytes_per_launch"] += tot
        e["launches"] += 1
        if rep not in e["source"]:
            e["source"].append(rep)
for e in out.values():
    e["dram_bytes_per_launch"] /= e["launches"]
print(json.dumps(out, indent=1))
