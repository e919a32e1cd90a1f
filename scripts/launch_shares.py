#!/usr/bin/env python3
"""Per-kernel average duration from an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections
import csv
import sys

for f in sys.argv[1:]:
    rows = [r for r in csv.reader(l for l in open(f) if l.startswith('"'))]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1.0)
        agg.setdefault(r[ki].split("(")[0], []).append(v)
    print(f)
    tot = sum(sum(v) / len(v) for v in agg.values())
    for k, v in agg.items():
        print("  %-28s launches=%2d avg_us=%10.1f share=%5.1f%%" % (k, len(v), sum(v) / len(v), 100 * sum(v) / len(v) / tot))
    print("  %-28s avg_us=%10.1f" % ("step total (serialised)", tot))
