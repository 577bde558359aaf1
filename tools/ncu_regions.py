#!/usr/bin/env python
"""Aggregate an ncu source-page CSV dump by named line ranges.
Usage: ncu_regions.py dump.csv name:lo-hi[,lo-hi] ..."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
regions = []
for spec in sys.argv[2:]:
    name, rng = spec.split(":")
    regions.append((name, [tuple(int(v) for v in r.split("-")) for r in rng.split(",")]))
hdr = None
tot = {}
for r in rows:
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or not r:
        continue
    try:
        line = int(r[0])
    except ValueError:
        continue
    d = dict(zip(hdr[2:], r[2:]))

    def num(k):
        try:
            return float(d.get(k, "0").replace(",", ""))
        except ValueError:
            return 0.0
    name = "other"
    for n, rr in regions:
        if any(lo <= line <= hi for lo, hi in rr):
            name = n
            break
    t = tot.setdefault(name, [0.0, 0.0, 0.0])
    t[0] += num("Instructions Executed")
    t[1] += num("# Samples")
    t[2] += num("L1 Wavefronts Shared")
ti = sum(t[0] for t in tot.values()) or 1
ts = sum(t[1] for t in tot.values()) or 1
for n, t in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("%-14s inst %6.2f%%  samples %6.2f%%  shared wavefronts %.3g" % (n, 100 * t[0] / ti, 100 * t[1] / ts, t[2]))
