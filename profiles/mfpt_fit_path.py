"""How the outer first-passage time of a long-loop bond depends on how far the spline
maximum-likelihood fit has converged.  Input: distance samples of one per-transition run
made on the CPU with the oracle (profiles/cpu_reference_dist.py, bit-identical to the
reference executable).  Prints the outer fpt at successive L-BFGS iterates from the flat
start y = 0 that the reference's tools/mfpt.py uses (its stopping rule: nlopt
ftol_rel = xtol_rel = 5e-5, tools/minimize.py:27-30).
usage: mfpt_fit_path.py samples.npz [column]"""
import sys

import numpy as np
from scipy import optimize

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from hybridmc_b200 import mfpt  # noqa: E402

z = np.load(sys.argv[1])
col = int(sys.argv[2]) if len(sys.argv) > 2 else z["nonlocal_bonds"].tolist().index(z["transient"].tolist()[0])
rc = 1.5
d = z["dist"][:, col]
off = d[d >= rc]
print("rows", len(d), "bond open in", len(off), "max distance %.2f" % off.max())
x_knot = np.linspace(rc, off.max(), mfpt.NKNOTS)
bw = mfpt._basis(x_knot, off).mean(axis=0)
qx, qw = mfpt._quadrature(x_knot)
bq = mfpt._basis(x_knot, qx)


def f_df(y):
    s = bq @ y
    m = (-s).max()
    e = qw * np.exp(-s - m)
    zz = e.sum()
    return bw @ y + np.log(zz) + m, bw - (bq.T @ e) / zz


path = []
optimize.minimize(f_df, np.zeros(mfpt.NKNOTS), jac=True, method="L-BFGS-B",
                  callback=lambda y: path.append((y.copy(), f_df(y)[0])),
                  options=dict(maxiter=2000, ftol=1e-15, gtol=1e-10))
prev = None
for it, (y, f) in enumerate(path):
    rel = abs(prev - f) / abs(f) if prev is not None else float("nan")
    if it < 20 or it == len(path) - 1 or it % 10 == 0:
        print("iterate %3d  f %.6f  rel. change %.1e  outer fpt %8.1f" % (
            it + 1, f, rel, mfpt.fpt_from_fit(x_knot, y, False)))
    prev = f
