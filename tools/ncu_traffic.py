"""Writes profiles/r02_kernel_traffic.json from `ncu --page raw --csv` exports: dram__bytes_read.sum + dram__bytes_write.sum of
one launch per named kernel.  usage: ncu_traffic.py key=raw.csv:kernel_regex[:launch_index] ..."""
import csv
import json
import os
import re
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
out_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r02_kernel_traffic.json")
out = json.load(open(out_path)) if os.path.exists(out_path) else {}
for arg in sys.argv[1:]:
    key, rest = arg.split("=", 1)
    parts = rest.split(":")
    path, rx, idx = parts[0], parts[1], int(parts[2]) if len(parts) > 2 else 0
    rows = list(csv.reader(open(path)))
    h = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    names, units = rows[h], rows[h + 1]
    hits = [r for r in rows[h + 2:] if re.search(rx, r[names.index("Kernel Name")])]
    r = hits[idx]
    tot = 0.0
    for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        i = names.index(m)
        tot += float(r[i].replace(",", "")) * UNIT[units[i]]
    out[key] = {"dram_bytes": tot, "kernel": r[names.index("Kernel Name")][:80], "duration_us_under_ncu": r[names.index("gpu__time_duration.sum")],
                "source": f"ncu --set full --clock-control none, {os.path.basename(path)} (profiles/), launch {idx} of /{rx}/"}
json.dump(out, open(out_path, "w"), indent=1)
print(json.dumps(out, indent=1))
