"""Share of executed warp instructions and stall samples per REGION of the kernel source,
from an ncu source page (see src_hotspots.py).  Regions are line ranges of the source version
the capture was taken from (default: the round-2 kernel as captured in profiles/r02/step_*).
usage: src_regions.py src.csv events_in_launch"""
import csv
import sys

REGIONS = [  # (first line, last line, name) in hmc_kernels.cu at commit bb2d6cd
    (102, 349, "helpers: min-image, bead load/store, group shuffles"),
    (350, 432, "predict_pair (exact FP64 prediction of one pair)"),
    (433, 495, "predict_cell / cell adjacency"),
    (496, 611, "RNG, trace, hamiltonian"),
    (612, 687, "list merge / commit of new predictions"),
    (688, 870, "constraint checks + initialize_system (once per MD interval)"),
    (871, 935, "next_event: delete + compact + argmin sweep"),
    (936, 1065, "event body: advance, collide, bond logic, write-back"),
    (1066, 1162, "cell-wall times of the event beads + probe masks"),
    (1163, 1274, "pass 1: packed-FP32 pre-filter over all partners"),
    (1275, 1332, "pass 2: control of the exact pass (calls predict_pair)"),
    (1333, 1431, "advance_all + sampling (once per MD interval)"),
    (1432, 1597, "crankshaft moves"),
    (1598, 1853, "kernel: schedule state machine, replica load/store"),
]


def num(x):
    try:
        return int(x)
    except Exception:
        return 0


rows = list(csv.reader(open(sys.argv[1])))
events = float(sys.argv[2]) if len(sys.argv) > 2 else None
hdr = rows[2]
iI, iS = hdr.index("Instructions Executed"), hdr.index("# Samples")
acc = {r[2]: [0, 0] for r in REGIONS}
other = [0, 0]
for r in rows[3:]:
    if r and r[0].isdigit():
        ln = int(r[0])
        tgt = other
        for a, b, name in REGIONS:
            if a <= ln <= b:
                tgt = acc[name]
                break
        tgt[0] += num(r[iI])
        tgt[1] += num(r[iS])
tot = sum(v[0] for v in acc.values()) + other[0]
tots = sum(v[1] for v in acc.values()) + other[1]
print("total warp-instructions %d%s" % (tot, ", %.0f per collision event" % (tot / events) if events else ""))
print("%-62s %8s %8s%s" % ("region", "instr %", "stall %", "  instr/event" if events else ""))
for a, b, name in REGIONS:
    v = acc[name]
    print("%-62s %8.2f %8.2f%s" % (name, 100.0 * v[0] / tot, 100.0 * v[1] / tots,
                                   "  %8.1f" % (v[0] / events) if events else ""))
print("%-62s %8.2f %8.2f" % ("(library headers: intrinsics, sqrt/div subroutines)", 100.0 * other[0] / tot,
                             100.0 * other[1] / tots))
