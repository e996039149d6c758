"""CPU: the restated first-passage-time estimator (hybridmc_b200/mfpt.py, after
tools/mfpt.py) on distributions with known answers, and samples vs histogram input."""
import ctypes as C

import numpy as np
import pytest

from common import load_config
from hybridmc_b200 import mfpt
from hybridmc_b200.params import params_from_json


def test_uniform_density_has_analytic_fpt():
    # p = 1/(b-a): inner = Int ((x-a)/(b-a))^2 (b-a) dx = (b-a)^2 / 3, outer the same
    rng = np.random.default_rng(0)
    a, b = 1.25, 1.5
    x = rng.uniform(a, b, 200000)
    assert abs(mfpt.fpt_from_samples(x, a, b, True) / ((b - a) ** 2 / 3) - 1) < 0.02
    assert abs(mfpt.fpt_from_samples(x, a, b, False) / ((b - a) ** 2 / 3) - 1) < 0.02


def test_r2_density_matches_numeric_integral():
    # p(r) ~ r^2 on [1.5, 9] (free relative diffusion in 3-D)
    rng = np.random.default_rng(1)
    lo, hi = 1.5, 9.0
    u = rng.uniform(0, 1, 400000)
    x = (lo ** 3 + u * (hi ** 3 - lo ** 3)) ** (1 / 3)
    r = np.linspace(lo, hi, 200001)
    pdf = 3 * r ** 2 / (hi ** 3 - lo ** 3)
    cdf = (r ** 3 - lo ** 3) / (hi ** 3 - lo ** 3)
    exact = np.trapezoid((1 - cdf) ** 2 / pdf, r)
    est = mfpt.fpt_from_samples(x, lo, hi, False)
    assert abs(est / exact - 1) < 0.03
    counts, edges = np.histogram(x, bins=4096, range=(0, 16.0))
    est_h = mfpt.fpt_from_hist(counts, 16.0, lo, hi, False)
    assert abs(est_h / est - 1) < 0.01


def test_samples_and_histogram_agree_on_simulated_distances(oracle):
    """dist rows of a short reference-exact 8bead_1bond program (oracle)"""
    from oracle_lib import OrcMainResult
    p = params_from_json(load_config("8bead_1bond", total_iter=150, total_iter_eq=20))
    r = OrcMainResult()
    cap = 150 * 8 * 3
    d = np.zeros((cap, 1))
    n = C.c_uint(0)
    assert oracle.orc_run_main(C.byref(p), None, 3, cap, d.ctypes.data_as(C.c_void_p), C.byref(n),
                               C.byref(r)) == 0
    x = d[:n.value, 0]
    on, off = x[x < p.rc], x[x >= p.rc]
    if len(on) < 200 or len(off) < 200:
        pytest.skip("too few samples on one side")
    rmax = 18.0 * 0.5 * 3 ** 0.5
    inner_s = mfpt.fpt_from_samples(on, p.rh, p.rc, True)
    outer_s = mfpt.fpt_from_samples(off, p.rc, off.max(), False)
    h_on, _ = np.histogram(on, bins=16384, range=(0, rmax))
    h_off, _ = np.histogram(off, bins=16384, range=(0, rmax))
    inner_h = mfpt.fpt_from_hist(h_on, rmax, p.rh, p.rc, True)
    outer_h = mfpt.fpt_from_hist(h_off, rmax, p.rc, None, False)
    assert abs(inner_h / inner_s - 1) < 0.05 and abs(outer_h / outer_s - 1) < 0.05
    assert 0.005 < inner_s < 0.05 and 0.5 < outer_s < 50  # same order as mfpt.csv short loops


def test_command_line_matches_the_reference_tools_layout(tmp_path):
    """python -m hybridmc_b200.mfpt json hdf5 nboot csv (tools/mfpt.py:32-150) on an output
    file with `dist` samples: one row per non-permanent bond, transient bond -> inner/outer,
    other bonds -> outer with the transient bond formed / open"""
    import csv
    import json

    from hybridmc_b200 import mfpt, output
    rng = np.random.default_rng(3)
    rc, rh = 1.5, 1.25
    n = 6000
    formed = rng.random(n) < 0.4
    d_t = np.where(formed, rng.uniform(rh, rc, n), rc + rng.gamma(2.0, 0.8, n))
    d_o = rc + rng.gamma(2.0, 1.0, n) + 0.3 * formed
    d_p = rng.uniform(rh, rc, n)
    data = dict(rc=rc, rh=rh, nonlocal_bonds=[[2, 6], [6, 10], [10, 14]],
                transient_bonds=[[6, 10]], permanent_bonds=[[10, 14]])
    (tmp_path / "t.json").write_text(json.dumps(data))
    output.write_hdf5(str(tmp_path / "t.h5"), {"dist": np.stack([d_o, d_t, d_p], axis=1)})
    mfpt.main([str(tmp_path / "t.json"), str(tmp_path / "t.h5"), "0", str(tmp_path / "t.csv")])
    rows = [[float(x) for x in r] for r in csv.reader(open(tmp_path / "t.csv"))]
    assert [int(r[0]) for r in rows] == [0, 1]  # the permanent bond (index 2) has no row
    inner = mfpt.fpt_from_samples(d_t[d_t < rc], rh, rc, True)
    outer = mfpt.fpt_from_samples(d_t[d_t >= rc], rc, d_t.max(), False)
    assert rows[1][1] == pytest.approx(inner, rel=1e-9) and rows[1][2] == pytest.approx(outer, rel=1e-9)
    on = d_o[(d_t < rc) & (d_o >= rc)]
    assert rows[0][1] == pytest.approx(mfpt.fpt_from_samples(on, rc, on.max(), False), rel=1e-9)
    assert all(np.isfinite(r[1]) and r[1] > 0 and r[2] > 0 for r in rows)


def test_histogram_bootstrap_variance_matches_repeated_sampling():
    """fpt_hist_bootstrap (the reference's half-sample bootstrap, tools/mfpt.py:383-395, as a
    multinomial draw over the bins): its variance is that of an estimate from n/2 samples,
    i.e. about twice the variance seen over independent full-size histograms"""
    rng = np.random.default_rng(11)
    lo, hi, rmax, n = 1.25, 1.5, 8.0, 40000

    def sample():
        x = rng.uniform(lo, hi, 3 * n)
        return x[rng.uniform(0, hi * hi, x.size) < x * x][:n]
    hists = [np.histogram(sample(), bins=2048, range=(0, rmax))[0] for _ in range(10)]
    vals = [mfpt.fpt_from_hist(h, rmax, lo, hi, True) for h in hists]
    boot = mfpt.fpt_hist_bootstrap(hists[0], rmax, lo, hi, True, nboot=12, seed=1)
    assert boot > 0
    ratio = boot / (2 * np.var(vals, ddof=1))
    assert 0.2 < ratio < 5.0, ratio


def test_run_mfpt_collects_a_campaign_directory(tmp_path):
    """python -m hybridmc_b200.run_mfpt DIR (tools/run_mfpt.py + tools/print_mfpt_to_file.py): one
    row per hybridmc_<layer>_<bits_i>_<bits_j>.{json,h5} pair, published header, layers ascending"""
    import csv
    import json

    from common import load_config
    from hybridmc_b200 import output, run_mfpt
    from hybridmc_b200.params import transition_jsons
    rng = np.random.default_rng(9)
    master = load_config("test")
    rmax, nbins = 15.0, 2048
    picked = [t for t in transition_jsons(master) if t[0] in (0, 1)][:4]
    expect = {}
    for layer, bits_in, bits_out, d in picked:
        fmt = lambda b: "".join("1" if x else "0" for x in b)  # noqa: E731
        name = "hybridmc_%d_%s_%s" % (layer, fmt(bits_in), fmt(bits_out))
        nl = sorted(sorted(b) for b in d["nonlocal_bonds"])
        col = nl.index(sorted(d["transient_bonds"][0]))
        hist = np.zeros((len(nl), 2, nbins), np.uint64)
        inner = rng.uniform(d["rh"], d["rc"], 30000)
        outer = d["rc"] + rng.gamma(2.0, 0.7 + 0.1 * layer, 30000)
        outer = outer[outer < rmax]
        hist[col, 1] = np.histogram(inner, bins=nbins, range=(0, rmax))[0]
        hist[col, 0] = np.histogram(outer, bins=nbins, range=(0, rmax))[0]
        (tmp_path / (name + ".json")).write_text(json.dumps(d))
        output.write_hdf5(str(tmp_path / (name + ".h5")), {"dist_hist": hist},
                          {"dist_hist_rmax": np.float64(rmax)})
        expect[(fmt(bits_in), fmt(bits_out))] = (
            layer, mfpt.fpt_from_samples(inner, d["rh"], d["rc"], True),
            mfpt.fpt_from_samples(outer, d["rc"], outer.max(), False))
    run_mfpt.main([str(tmp_path), "--procs", "2"])
    rows = list(csv.reader(open(tmp_path / "mfpt.csv")))
    assert rows[0] == ["state i bits", "state j bits", "layer", "inner fpt", "outer fpt"]
    assert len(rows) == 5 and [int(r[2]) for r in rows[1:]] == sorted(int(r[2]) for r in rows[1:])
    for bi, bo, layer, inner, outer in rows[1:]:
        e = expect[(bi, bo)]
        assert int(layer) == e[0]
        assert float(inner) == pytest.approx(e[1], rel=0.02) and float(outer) == pytest.approx(e[2], rel=0.03)
