"""Long-loop outer first-passage times of the crambin sweep, fitted (a) to the maximum of the
likelihood and (b) with the reference tool's stopping rule (flat start, rel. tolerances 5e-5,
tools/minimize.py:26-30 — hybridmc_b200.mfpt.fit_pmf(reference_tolerance=True)), against the
published mfpt.csv.  Build container only (reads /root/reference/crambin_s_bias_mfpt)."""
import csv
import multiprocessing
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hybridmc_b200 import mfpt  # noqa: E402

PUB = "/root/reference/crambin_s_bias_mfpt/mfpt.csv"
pub = {(r[0], r[1]): (float(r[3]), float(r[4])) for r in list(csv.reader(open(PUB)))[1:]}
packs = [np.load(p) for p in sys.argv[1:]]
RMAX, RC = float(packs[0]["rmax"]), float(packs[0]["rc"])
items = []
for z in packs:
    for name, h in zip(z["names"].tolist(), z["hist"]):
        _, layer, bi, bo = name.split("_")
        if (bi, bo) in pub and pub[(bi, bo)][1] >= 30:  # long loops
            items.append(((bi, bo), h[0]))


def one(item):
    key, h = item
    try:
        return key, mfpt.fpt_from_hist(h, RMAX, RC, None, False), \
            mfpt.fpt_from_hist(h, RMAX, RC, None, False, reference_tolerance=True)
    except Exception:
        return key, float("nan"), float("nan")


if __name__ == "__main__":
    os.environ["OMP_NUM_THREADS"] = "1"
    with multiprocessing.Pool(7) as pool:
        res = pool.map(one, items, chunksize=8)
    a = np.array([[r[1], r[2], pub[r[0]][1]] for r in res if np.isfinite(r[1]) and np.isfinite(r[2])])
    print("%d long-loop transitions (published outer fpt >= 30)" % len(a))
    for col, name in ((0, "maximum-likelihood fit"), (1, "reference stopping rule")):
        r = np.log(a[:, col] / a[:, 2])
        print("%-26s outer fpt / published: median %.3f, 16-84 %% %.3f .. %.3f, corr(log) %.4f" % (
            name, np.exp(np.median(r)), np.exp(np.percentile(r, 16)), np.exp(np.percentile(r, 84)),
            np.corrcoef(np.log(a[:, col]), np.log(a[:, 2]))[0, 1]))
