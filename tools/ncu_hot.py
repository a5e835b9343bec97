"""Hottest SASS lines of one kernel from `ncu -i X.ncu-rep --page source --csv` (first kernel block in the file)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
ix = {k: i for i, k in enumerate(hdr)}
data = []
for r in rows[2:]:
    if len(r) <= ix['stall_wait'] or not r[ix['# Samples']].isdigit():
        if r and r[0] == "Kernel Name":
            break
        continue
    data.append(r)
tot = sum(int(r[ix['# Samples']]) for r in data)
print("total samples", tot, "instructions", len(data))
top = sorted(enumerate(data), key=lambda t: -int(t[1][ix['# Samples']]))[:n]
for i, r in sorted(top):
    print(f"{i:5d} {r[ix['Source']][:60]:60s} smp {r[ix['# Samples']]:>5s} long {r[ix['stall_long_sb']]:>4s} short {r[ix['stall_short_sb']]:>4s} "
          f"wait {r[ix['stall_wait']]:>4s} br {r[ix['stall_branch_resolving']]:>4s} noi {r[ix['stall_no_inst']]:>4s} mio {r[ix['stall_mio']]:>3s} "
          f"math {r[ix['stall_math']]:>3s} exec {r[ix['Instructions Executed']]:>8s}")
