"""First-passage times from bond-distance statistics — a self-contained restatement of
the reference's `tools/mfpt.py:153-381` (which needs nlopt, h5py and
`scipy.integrate.quadrature`, none of which exist in this image; versions unpinned in
the reference, so this boundary is "parity unpinned" and validated statistically against
the published `crambin_s_bias_mfpt/mfpt.csv`).

Model (mfpt.py:372-381, 211-222, 340-369): the potential of mean force -ln p(r) on
[lo, hi] is a NATURAL cubic spline through `nknots` = 14 equally spaced knots whose
ordinates maximise the likelihood of the sampled distances,
        f(y) = mean_i S(r_i; y) + ln Integral_lo^hi exp(-S(r; y)) dr,
and with pdf = exp(-S)/Z, cdf its integral,
        inner fpt = Integral cdf^2 / pdf dr        (bond formed, r in [rh, rc])
        outer fpt = Integral (1 - cdf)^2 / pdf dr  (bond open, r in [rc, r_max]).
Because S is linear in the knot ordinates (cardinal natural splines), f is convex; it is
minimised here with scipy's L-BFGS (the reference: nlopt LD_LBFGS, rel. tol 5e-5) using
Gauss-Legendre quadrature per knot interval.  Works on raw samples (the `dist` dataset)
or on the engine's pooled histograms (`hmc_get_dist_hist`), i.e. a binned likelihood.
"""
import numpy as np
from scipy import interpolate, optimize

NKNOTS = 14  # mfpt.py:62


def _basis(x_knot, x):
    """B[i, k] = value at x[i] of the natural cubic spline that is 1 at knot k, 0 at the
    others (scipy CubicSpline(bc_type='natural'), as mfpt.py:172)."""
    eye = np.eye(len(x_knot))
    spl = interpolate.CubicSpline(x_knot, eye, bc_type="natural", axis=0)
    return spl(np.clip(x, x_knot[0], x_knot[-1]))


def _quadrature(x_knot, order=24):
    """Gauss-Legendre nodes / weights on every knot interval"""
    t, w = np.polynomial.legendre.leggauss(order)
    a, b = x_knot[:-1, None], x_knot[1:, None]
    nodes = 0.5 * (b - a) * t[None, :] + 0.5 * (b + a)
    weights = 0.5 * (b - a) * w[None, :]
    return nodes.ravel(), weights.ravel()


REFERENCE_TOL = 5e-5  # tools/minimize.py:26-30: ftol_rel = xtol_rel = 5e-5


def fit_pmf(x, weights, lo, hi, nknots=NKNOTS, reference_tolerance=False):
    """Maximum-likelihood knot ordinates y of -ln p(r) on [lo, hi] for data points x
    with weights (sample: weights = 1; histogram: bin centres and counts).
    Returns (x_knot, y).

    reference_tolerance: an experiment switch — L-BFGS from the flat start y = 0
    (tools/mfpt.py:376) ended by nlopt's relative tolerances as tools/minimize.py:26-30 sets
    them: the first iterate whose step changed the objective by less than 5e-5 |f| or every
    ordinate by less than 5e-5 |y_i|.  With scipy's L-BFGS path this rule stops close to the
    maximum-likelihood fit (profiles/long_loop_reference_tolerance.py: the long-loop outer
    times stay 2.2x the published ones), so the stopping rule alone does not explain the
    published long-loop values; nlopt's own iterates cannot be reproduced here."""
    x_knot = np.linspace(lo, hi, nknots)
    w = np.asarray(weights, dtype=float)
    w = w / w.sum()
    bw = _basis(x_knot, np.asarray(x, dtype=float)).T @ w  # mean_i B_k(r_i)
    qx, qw = _quadrature(x_knot)
    bq = _basis(x_knot, qx)

    def f_df(y):
        s = bq @ y
        m = (-s).max()
        e = qw * np.exp(-s - m)
        z = e.sum()
        return bw @ y + np.log(z) + m, bw - (bq.T @ e) / z

    if reference_tolerance:
        path = [np.zeros(nknots)]
        optimize.minimize(f_df, np.zeros(nknots), jac=True, method="L-BFGS-B",
                          callback=lambda xk: path.append(np.array(xk)),
                          options=dict(maxiter=200, ftol=1e-13, gtol=1e-9))
        f_prev = f_df(path[0])[0]
        for k in range(1, len(path)):
            f_k = f_df(path[k])[0]
            dy = np.abs(path[k] - path[k - 1])
            if abs(f_k - f_prev) < REFERENCE_TOL * abs(f_k) or np.all(dy < REFERENCE_TOL * np.abs(path[k])):
                return x_knot, path[k]
            f_prev = f_k
        return x_knot, path[-1]
    res = optimize.minimize(f_df, np.zeros(nknots), jac=True, method="L-BFGS-B",
                            options=dict(maxiter=2000, ftol=1e-13, gtol=1e-9))
    return x_knot, res.x


def fpt_from_fit(x_knot, y, inner, order=24):
    """mfpt.py:340-369 with pdf/cdf from the fitted spline."""
    qx, qw = _quadrature(x_knot, order)
    s = _basis(x_knot, qx) @ y
    e = np.exp(-(s - s.min()))
    z = (qw * e).sum()
    pdf = e / z
    # cdf at the quadrature nodes: full intervals to the left + partial interval by a
    # second Gauss-Legendre rule on [knot, node]
    n_int = len(x_knot) - 1
    per = pdf.reshape(n_int, order) * qw.reshape(n_int, order)
    left = np.concatenate([[0.0], np.cumsum(per.sum(axis=1))[:-1]])
    t, w = np.polynomial.legendre.leggauss(order)
    cdf = np.empty_like(pdf)
    for i in range(n_int):
        a = x_knot[i]
        for j in range(order):
            b = qx[i * order + j]
            xx = 0.5 * (b - a) * t + 0.5 * (b + a)
            ss = _basis(x_knot, xx) @ y
            cdf[i * order + j] = left[i] + (0.5 * (b - a) * w * np.exp(np.clip(-(ss - s.min()), -745.0, 700.0)) / z).sum()
    integrand = cdf * cdf / pdf if inner else (1.0 - cdf) * (1.0 - cdf) / pdf
    return float((qw * integrand).sum())


def fpt_from_samples(dist, lo, hi, inner, nknots=NKNOTS):
    """fpt_per_bead_pair, mfpt.py:372-381, on raw samples restricted to [lo, hi]."""
    d = np.asarray(dist, dtype=float)
    d = d[(d >= lo) & (d <= hi)]
    if len(d) < nknots:
        raise ValueError("too few samples in [%g, %g]" % (lo, hi))
    xk, y = fit_pmf(d, np.ones(len(d)), lo, hi, nknots)
    return fpt_from_fit(xk, y, inner)


def fpt_from_hist(counts, rmax, lo, hi, inner, nknots=NKNOTS, reference_tolerance=False):
    """The same from a histogram of `len(counts)` equal bins over [0, rmax) (binned
    likelihood).  A bin contributes the AVERAGE of the spline over its part inside [lo, hi]
    (3-point Gauss-Legendre; the samples of a state-split histogram all lie inside the state's
    range, so a bin that straddles rh or rc counts in full) — with bin centres instead, the
    15 bins a 2 048-bin histogram over the half box diagonal has inside [rh, rc] bias the
    inner time by -1 %.  For the outer time hi=None means "up to the last occupied bin"
    (mfpt.py:137: max(t_off))."""
    counts = np.asarray(counts, dtype=float)
    n = len(counts)
    width = rmax / n
    left = np.arange(n) * width
    if hi is None:
        nz = np.nonzero(counts)[0]
        hi = (nz[-1] + 1) * width
    a = np.maximum(left, lo)
    b = np.minimum(left + width, hi)
    sel = (b > a + 1e-12 * width) & (counts > 0)
    if sel.sum() < 4:
        raise ValueError("too few occupied bins in [%g, %g]" % (lo, hi))
    t, w = np.polynomial.legendre.leggauss(3)
    a, b = a[sel, None], b[sel, None]
    nodes = 0.5 * (b - a) * t[None, :] + 0.5 * (b + a)
    weights = counts[sel, None] * (0.5 * w)[None, :]
    xk, y = fit_pmf(nodes.ravel(), weights.ravel(), lo, hi, nknots, reference_tolerance)
    return fpt_from_fit(xk, y, inner)


def fpt_hist_bootstrap(counts, rmax, lo, hi, inner, nboot, seed=0, nknots=NKNOTS):
    """Variance of the histogram estimate by the reference's bootstrap (tools/mfpt.py:383-395:
    `nboot` resamples of HALF the sample size, with replacement): resampling n/2 of the n
    binned distances is a multinomial draw over the bins."""
    c = np.asarray(counts, dtype=np.int64)
    n = int(c.sum())
    rng = np.random.default_rng(seed)
    vals = [fpt_from_hist(rng.multinomial(n // 2, c / n), rmax, lo, hi, inner, nknots)
            for _ in range(nboot)]
    return float(np.var(vals, ddof=1))


def transition_fpts(hist, rmax, params, bond_column=None):
    """(inner fpt, outer fpt) of a transition's transient bond from the engine's histograms
    hist[n_nonlocal][2][nbins] (index 1 = transient bond formed): the two numbers a row of
    the reference's mfpt.csv holds."""
    if bond_column is None:
        t = params.bond_list("transient")[0]
        cols = sorted(params.bond_list("nonlocal"))  # dist columns are in (i,j) scan order
        bond_column = cols.index(t)
    h = np.asarray(hist)[bond_column]
    inner = fpt_from_hist(h[1], rmax, params.rh, params.rc, True)
    outer = fpt_from_hist(h[0], rmax, params.rc, None, False)
    return inner, outer


# ---- the reference tool's command line (tools/mfpt.py:32-150) -----------------------------

def fpt_rows(data, dist=None, hist=None, rmax=None, nboot=0, seed=0):
    """Rows of the reference's per-transition csv: for every nonlocal bond i that is not
    permanent, [i, fpt_on, fpt_off (, var_on, var_off)] — for the transient bond fpt_on is
    the INNER time of the formed bond and fpt_off the outer time of the open bond; for any
    other bond both are outer times, with the transient bond formed / open
    (tools/mfpt.py:99-146).  Input: the `dist` samples [rows][n_nonlocal] of the executable,
    or the engine's state-split histograms `hist[n_nonlocal][2][nbins]` over [0, rmax)."""
    rc, rh = data["rc"], data["rh"]
    nl = [sorted(b) for b in data["nonlocal_bonds"]]
    t_ind = nl.index(sorted(data["transient_bonds"][0]))
    p_ind = [i for i, b in enumerate(nl) if b in [sorted(p) for p in data.get("permanent_bonds", [])]]
    rng = np.random.default_rng(seed)
    rows = []
    for i in range(len(nl)):
        if i in p_ind:
            continue
        if dist is not None:
            d_i, d_t = np.asarray(dist)[:, i], np.asarray(dist)[:, t_ind]
            if i == t_ind:
                on, off = d_i[d_i < rc], d_i[d_i >= rc]
            else:
                on, off = d_i[(d_t < rc) & (d_i >= rc)], d_i[(d_t >= rc) & (d_i >= rc)]

            def est(x, inner):
                return fpt_from_samples(x, rh, rc, True) if inner else \
                    fpt_from_samples(x, rc, x.max(), False)

            def var(x, inner):  # bootstrap over half-size resamples (tools/mfpt.py:383-395)
                vals = [est(rng.choice(x, size=len(x) // 2, replace=True), inner) for _ in range(nboot)]
                return float(np.var(vals, ddof=1))
            row = [i, est(on, i == t_ind), est(off, False)]
            if nboot > 1:
                row += [var(on, i == t_ind), var(off, False)]
        else:
            h = np.asarray(hist)[i]
            row = [i, fpt_from_hist(h[1], rmax, rh, rc, True) if i == t_ind else
                   fpt_from_hist(h[1], rmax, rc, None, False),
                   fpt_from_hist(h[0], rmax, rc, None, False)]
            if nboot > 1:
                row += [fpt_hist_bootstrap(h[1], rmax, rh if i == t_ind else rc,
                                           rc if i == t_ind else None, i == t_ind, nboot, seed),
                        fpt_hist_bootstrap(h[0], rmax, rc, None, False, nboot, seed)]
        rows.append(row)
    return rows


def main(argv=None):
    """python -m hybridmc_b200.mfpt json hdf5 nboot csv — the arguments of tools/mfpt.py"""
    import argparse
    import csv
    import json

    from .output import File
    ap = argparse.ArgumentParser()
    ap.add_argument("json", help="json input file")
    ap.add_argument("hdf5", help="hdf5 input file (executable output: dist samples; campaign "
                                 "output: dist_hist histograms)")
    ap.add_argument("nboot", type=int, help="number of bootstrap samples to run (samples only)")
    ap.add_argument("csv", help="csv output file")
    ap.add_argument("--layers", help="accepted for compatibility")
    args = ap.parse_args(argv)
    with open(args.json) as f:
        data = json.load(f)
    with File(args.hdf5) as f:
        if "dist" in f.keys() and len(f["dist"]):
            print("number of distances =", len(f["dist"]))
            rows = fpt_rows(data, dist=f["dist"], nboot=args.nboot)
        else:
            rows = fpt_rows(data, hist=f["dist_hist"], rmax=float(f.attrs["dist_hist_rmax"]),
                            nboot=args.nboot)
    with open(args.csv, "w", newline="") as f:
        csv.writer(f).writerows(rows)


if __name__ == "__main__":
    main()
