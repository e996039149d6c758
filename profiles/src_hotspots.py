"""Aggregate an ncu source page (cuda,sass view) per CUDA source line.
usage: ncu -i X.ncu-rep --page source --print-source cuda,sass --csv > src.csv
       python profiles/src_hotspots.py src.csv [top]"""
import csv
import sys


def num(x):
    try:
        return int(x)
    except Exception:
        return 0


rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
hdr = rows[2]
iI = hdr.index("Instructions Executed")
iT = hdr.index("Thread Instructions Executed")
iS = hdr.index("# Samples")
lines = []
for r in rows[3:]:
    if r and r[0].isdigit():
        lines.append((int(r[0]), r[1], num(r[iI]), num(r[iT]), num(r[iS])))
tot = sum(l[2] for l in lines) or 1
tots = sum(l[4] for l in lines) or 1
print("total warp-instructions %d, stall samples %d" % (tot, tots))
lines.sort(key=lambda x: -x[4])
for l in lines[:top]:
    print("%5d  inst %5.2f%%  samples %5.2f%%  lanes %5.1f  %s" % (
        l[0], 100 * l[2] / tot, 100 * l[4] / tots, l[3] / max(l[2], 1), l[1][:100]))
