"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel.  python -m tools.launch_summary in.csv [header lines...]"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
ik, iv, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
tot = collections.OrderedDict()
for r in rows[1:]:
    if r[im] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"^void ", "", r[ik])
    name = re.sub(r"s3od::", "", name)[:100]
    us = float(r[iv].replace(",", "")) / 1e3
    n, t = tot.get(name, (0, 0.0))
    tot[name] = (n + 1, t + us)
total = sum(t for _, t in tot.values())
for line in sys.argv[2:]:
    print("# " + line)
print(f"# total {total / 1e3:.2f} ms over {sum(n for n, _ in tot.values())} launches")
print("kernel\tlaunches\ttotal_us\tshare\tavg_us")
for name, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{name}\t{n}\t{t:.1f}\t{t / total:.4f}\t{t / n:.1f}")
