"""First-passage times of every transition of a campaign directory — the counterpart of the
reference's `tools/run_mfpt.py` (one `mfpt.py` job per transition in a multiprocessing pool)
followed by `tools/print_mfpt_to_file.py:11-33` (collect into `mfpt.csv`):

  python -m hybridmc_b200.run_mfpt DIR [--procs N]

reads DIR/hybridmc_<layer>_<bits_i>_<bits_j>.{json,h5} as run_campaign wrote them, fits the
transient bond's inner / outer first-passage time from the pooled `dist_hist` histograms
(hybridmc_b200/mfpt.py, the estimator of tools/mfpt.py) and writes DIR/mfpt.csv with the
published header `state i bits,state j bits,layer,inner fpt,outer fpt`."""
import argparse
import csv
import json
import multiprocessing
import os
import re
import sys

if __package__ in (None, ""):
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    __package__ = "hybridmc_b200"

from . import mfpt
from .output import File
from .params import params_from_json

NAME = re.compile(r"hybridmc_([0-9]+)_([01]+)_([01]+)\.h5$")


def _one(path):
    m = NAME.search(path)
    with open(path[:-3] + ".json") as f:
        p = params_from_json(json.load(f))
    try:
        with File(path) as f:
            inner, outer = mfpt.transition_fpts(f["dist_hist"], float(f.attrs["dist_hist_rmax"]), p)
    except Exception as ex:  # a state too thinly sampled to fit
        sys.stderr.write("%s: %s\n" % (os.path.basename(path), ex))
        inner = outer = float("nan")
    return [m.group(2), m.group(3), int(m.group(1)), inner, outer]


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("dir")
    ap.add_argument("--procs", type=int, default=None)
    args = ap.parse_args(argv)
    paths = sorted(os.path.join(args.dir, n) for n in os.listdir(args.dir) if NAME.search(n))
    with multiprocessing.Pool(args.procs) as pool:
        rows = pool.map(_one, paths, chunksize=8)
    rows.sort(key=lambda r: r[2])
    with open(os.path.join(args.dir, "mfpt.csv"), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["state i bits", "state j bits", "layer", "inner fpt", "outer fpt"])
        w.writerows(rows)
    print("%d transitions -> %s" % (len(rows), os.path.join(args.dir, "mfpt.csv")))


if __name__ == "__main__":
    main()
