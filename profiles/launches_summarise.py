"""ncu launch list (ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X.csv <cmd>) -> per-kernel totals and shares.
Usage: python profiles/launches_summarise.py gpurun_out/r01zg_launches.csv "<command that was profiled>" > profiles/r01zg_launches_summary.csv"""
import csv
import re
import sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = {}
for r in rows[1:]:
    if len(r) <= vi:
        continue
    name = re.sub(r"\(.*", "", r[ki]).split("::")[-1]
    v = float(r[vi].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1.0)
    n, t = tot.get(name, (0, 0.0))
    tot[name] = (n + 1, t + v)
allt = sum(t for _, t in tot.values())
print("# ncu launch list of: %s" % (sys.argv[2] if len(sys.argv) > 2 else "?"))
print("# ncu --metrics gpu__time_duration.sum --clock-control none --csv ; per-launch times are cold-cache and serialised: compare SHARES with bench.py kernels{}.share")
print("kernel,launches,total_us,share")
for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("%s,%d,%.1f,%.4f" % (k, n, t, t / allt))
