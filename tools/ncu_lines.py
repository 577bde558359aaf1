#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump by CUDA source line:
instructions executed, stall samples, shared wavefronts.  Usage: ncu_lines.py dump.csv [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = None
out = []
for r in rows:
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or not r or r[0] in ("", "File Path", "Function Name"):
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
    out.append((line, r[1].strip()[:90], num("Instructions Executed"), num("# Samples"),
                num("L1 Wavefronts Shared"), num("L1 Wavefronts Shared Excessive"), num("stall_barrier"),
                num("stall_long_sb"), num("stall_short_sb"), num("stall_math"), num("stall_wait"), num("stall_mio")))
ti = sum(o[2] for o in out) or 1
ts = sum(o[3] for o in out) or 1
print("total inst %.3g  samples %d" % (ti, ts))
print("%5s %6s %6s %9s %6s | bar longsb shortsb math wait mio | source" % ("line", "inst%", "smp%", "shwave", "excess"))
for o in sorted(out, key=lambda o: -o[3])[:top]:
    print("%5d %6.2f %6.2f %9.3g %6.3g | %4d %4d %4d %4d %4d %4d | %s" % (
        o[0], 100 * o[2] / ti, 100 * o[3] / ts, o[4], o[5], o[6], o[7], o[8], o[9], o[10], o[11], o[1]))
