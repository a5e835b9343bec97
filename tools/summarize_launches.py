"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total time, share."""
import collections
import csv
import re
import sys


def main(path, out=None, top=60):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except Exception:
            continue
        u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v
        name = re.sub(r"\(.*", "", re.sub(r"<.*", "", row["Kernel Name"]))[:90]
        agg[name][0] += 1
        agg[name][1] += v
        tot += v
    w = open(out, "w") if out else sys.stdout
    ours = sum(t for k, (n, t) in agg.items() if "pdb::" in k)
    w.write(f"# {path}: {sum(a[0] for a in agg.values())} launches, {tot / 1e3:.2f} ms of kernel time "
            f"(cold-cache, serialised under ncu); libpdb200 kernels: {ours / 1e3:.2f} ms = {100 * ours / tot:.1f}%\n")
    w.write(f"{'total_us':>12} {'share':>7} {'count':>6}  kernel\n")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        w.write(f"{t:12.1f} {100 * t / tot:6.1f}% {n:6d}  {k}\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
