#!/usr/bin/env python
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): mean time and share per kernel."""
import csv
import sys
from collections import OrderedDict

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = OrderedDict()
for r in rows[1:]:
    agg.setdefault(r[ki][:64], []).append(float(r[vi].replace(",", "")))
tot = sum(sum(v) for v in agg.values())
for k, v in agg.items():
    print("%-66s n=%2d mean=%10.1f us share=%5.1f%%" % (k, len(v), sum(v) / len(v) / 1e3, 100 * sum(v) / tot))
