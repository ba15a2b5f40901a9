#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
usage: python scripts/launch_breakdown.py gpurun_out/launches.csv [out.txt]"""
import csv, sys, re, collections
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
r = csv.reader(lines)
hdr = next(r)
ni, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for row in r:
    if len(row) <= vi or row[mi] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", row[ni]); name = re.sub(r"^void ", "", name); name = name.replace("sarssl::", "")
    v = float(row[vi].replace(",", "")); u = row[ui]
    us = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1e-3)
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += us
tot = sum(a[1] for a in agg.values())
out = ["%-90s %8s %12s %7s" % ("kernel", "launches", "total_us", "share")]
for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append("%-90s %8d %12.1f %6.1f%%" % (name[:90], n, us, 100 * us / tot))
out.append("%-90s %8d %12.1f" % ("TOTAL", sum(a[0] for a in agg.values()), tot))
txt = "\n".join(out)
print(txt)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(txt + "\n")
