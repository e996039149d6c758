"""mfpt.csv (published header) from the histograms packed by collect_hist.py.
usage: fit_hist_npz.py IN.npz OUT.csv [procs]"""
import csv
import multiprocessing
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hybridmc_b200 import mfpt  # noqa: E402

z = np.load(sys.argv[1])
RMAX, RH, RC = float(z["rmax"]), float(z["rh"]), float(z["rc"])


def one(args):
    name, h = args
    _, layer, bi, bo = name.split("_")
    try:
        inner = mfpt.fpt_from_hist(h[1], RMAX, RH, RC, True)
        outer = mfpt.fpt_from_hist(h[0], RMAX, RC, None, False)
    except Exception as ex:
        sys.stderr.write("%s: %s\n" % (name, ex))
        inner = outer = float("nan")
    return [bi, bo, int(layer), inner, outer]


if __name__ == "__main__":
    with multiprocessing.Pool(int(sys.argv[3]) if len(sys.argv) > 3 else None) as pool:
        rows = pool.map(one, list(zip(z["names"].tolist(), z["hist"])), chunksize=16)
    rows.sort(key=lambda r: r[2])
    with open(sys.argv[2], "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["state i bits", "state j bits", "layer", "inner fpt", "outer fpt"])
        w.writerows(rows)
    print(len(rows), "rows ->", sys.argv[2])
