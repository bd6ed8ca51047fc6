#!/usr/bin/env python3
"""Per-CUDA-line summary of an ncu report's source page (stall samples, instructions executed).
usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > src.csv ; python tools/ncu_lines.py src.csv [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = None
out = []
fname = ""
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    if len(r) > 10 and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < 10:
        continue
    if r[0] != "" and r[2] == "-":
        def g(name):
            try:
                return float(r[hdr.index(name)])
            except Exception:
                return 0.0
        out.append((g("Warp Stall Sampling (All Samples)"), g("Instructions Executed"), fname, r[0], r[1][:110],
                    g("stall_long_sb"), g("stall_short_sb"), g("stall_lg"), g("stall_wait") if "stall_wait" in hdr else 0))
tot = sum(o[0] for o in out) or 1
toti = sum(o[1] for o in out) or 1
print("total samples %d, total warp-instructions %.3g" % (tot, toti))
for o in sorted(out, reverse=True)[:top]:
    print("%5.1f%% smp %5.1f%% inst  long_sb %5d short %5d lg %5d | %s:%s  %s" % (100 * o[0] / tot, 100 * o[1] / toti, o[5], o[6], o[7], o[2], o[3], o[4]))
