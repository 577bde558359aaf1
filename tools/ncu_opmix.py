#!/usr/bin/env python
"""Dynamic SASS opcode mix per named source-line region from an `ncu --page source --csv
--print-source cuda,sass` dump.  Usage: ncu_opmix.py dump.csv name:lo-hi[,lo-hi] ..."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
regions = []
for spec in sys.argv[2:]:
    name, rng = spec.split(":")
    regions.append((name, [tuple(int(v) for v in r.split("-")) for r in rng.split(",")]))
hdr = None
cur = "other"
mix = collections.defaultdict(collections.Counter)
for r in rows:
    if r and r[0] == "Line No":
        hdr = r
        iex = hdr.index("Instructions Executed")
        continue
    if hdr is None or not r:
        continue
    if r[0] != "":
        try:
            line = int(r[0])
        except ValueError:
            continue
        cur = "other"
        for n, rr in regions:
            if any(lo <= line <= hi for lo, hi in rr):
                cur = n
                break
        continue
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[3])
    if not m:
        continue
    try:
        n = float(r[iex].replace(",", ""))
    except ValueError:
        continue
    mix[cur][m.group(2)] += n
tot = sum(sum(c.values()) for c in mix.values())
for name, c in sorted(mix.items(), key=lambda kv: -sum(kv[1].values())):
    s = sum(c.values())
    print("%-10s %6.2f%%  " % (name, 100 * s / tot) + "  ".join("%s %.1f" % (k, 100 * v / s) for k, v in c.most_common(14)))
