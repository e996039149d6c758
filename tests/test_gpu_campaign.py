"""Batched campaign launcher end to end on the GPU: every transition
of a lattice layer as one ensemble, structures handed from layer to layer, one HDF5 file per
transition with the reference's names (tools/init_json.py:109-125)."""
import csv
import os

import numpy as np
import pytest

from common import load_config
from hybridmc_b200 import h5lite, mfpt
from hybridmc_b200.output import File
from hybridmc_b200.params import params_from_json
from hybridmc_b200.run_campaign import run_campaign

pytestmark = pytest.mark.gpu


def test_two_layers_of_test_json(tmp_path):
    master = load_config("test")
    out = str(tmp_path / "camp")
    summary, seconds, events = run_campaign(master, out, replicas=48, iters_per_round=2, max_rounds=2,
                                    max_layers=2, hist_bins=2048, wl=False, eq_iters=2,
                                    log=lambda *a: None)
    names = sorted(f for f in os.listdir(out) if f.endswith(".h5"))
    assert len(names) == 3 + 6 and len(summary) == 9  # layers 0 and 1 of the 3-bond lattice
    assert names[0] == "hybridmc_0_000_001.h5" or names[0].startswith("hybridmc_0_000_")
    nb = master["nbeads"]
    for n in names:
        path = os.path.join(out, n)
        assert h5lite.is_hdf5(path) and os.path.exists(path[:-3] + ".json")
        with File(path) as f:
            assert f["s_bias"].dtype == np.float32 and f["s_bias"].shape == (2,)
            assert np.isfinite(f["s_bias"]).all()
            cc = f["config_count"]
            # a zero row, then one row per G-test round (main.cc:394,488)
            assert cc.dtype == np.uint64 and cc.shape in ((2, 2), (3, 2))
            assert not cc[0].any() and cc[1:].all()
            # what tools/mfpt.py:76-77 and tools/get_error_s_bias.py read
            nl = len(f.attrs["nonlocal_bonds"])
            assert f["dist"].ndim == 2 and f["dist"].shape[1] == nl and len(f["dist"]) > 0
            assert (f["dist"] > 0).all() and (f["dist"] < master["length"]).all()
            ci = f["config"]["int"]
            assert ci.dtype == np.uint64 and len(ci) > 0 and set(np.unique(ci)) <= {0, 1}
            assert f["unique_config"]["vel"].shape == f["unique_config"]["pos"].shape
            assert len(f["unique_config"]["config"]) == len(f["unique_config"]["pos"])
            assert int(f.attrs["err_flags"]) & ~32 == 0 and int(f.attrs["replicas"]) == 48
            assert f["dist_hist"].shape[-2:] == (2, 2048)
            assert f["unique_config"]["pos"].shape[1:] == (nb, 3)
            assert f.attrs["nsteps"] == master["nsteps"]
            assert f.attrs["transient_bonds"].shape == (1, 2)
            layer = int(n.split("_")[1])
            assert len(f.attrs["permanent_bonds"]) == layer
    # the estimator of row f3 runs on what the campaign wrote
    with File(os.path.join(out, names[0])) as f:
        with open(os.path.join(out, names[0][:-3] + ".json")) as jf:
            import json
            p = params_from_json(json.load(jf))
        inner, outer = mfpt.transition_fpts(f["dist_hist"], float(f.attrs["dist_hist_rmax"]), p)
        assert np.isfinite(inner) and 0 < inner < 1 and outer > 0
        # ... and, like tools/mfpt.py, on the raw `dist` rows of the same file
        t = p.bond_list("transient")[0]
        col = [tuple(b) for b in p.bond_list("nonlocal")].index(tuple(t))
        d = f["dist"][:, col]
        assert ((d > p.rh * 0.99) & (d < master["length"])).all()
    assert events > 0
    with open(os.path.join(out, "diff_s_bias.csv")) as f:
        assert len(list(csv.reader(f))) == 9
    # finished transitions are skipped on a second pass (tools/init_json.py:141-142)
    summary2, _, _ = run_campaign(master, out, replicas=48, iters_per_round=2, max_rounds=2,
                               max_layers=2, hist_bins=2048, wl=False, eq_iters=2,
                               log=lambda *a: None)
    assert summary2 == []


def test_staircase_sweep_layer0(tmp_path):
    """init_json.py --bp/--rc (tools/init_json.py:93-120): the transition that forms the
    staircase bond runs in stages rc = 3.0 -> 2.2 -> 1.5, each stage started from the previous
    stage's bonded structure, intermediate files carry their rc in the name"""
    master = load_config("test")
    bonds = sorted(sorted(b) for b in master["nonlocal_bonds"])
    bp = bonds[0]
    out = str(tmp_path / "stair")
    summary, _, _ = run_campaign(master, out, replicas=48, iters_per_round=2, max_rounds=2,
                              max_layers=1, hist_bins=512, wl=False, eq_iters=2,
                              log=lambda *a: None, stair_bp=bp, stair_rc=[3.0, 2.2, 1.5])
    names = sorted(f[:-3] for f in os.listdir(out) if f.endswith(".h5"))
    stair = [n for n in names if n.startswith("hybridmc_0_000_100")]
    assert stair == ["hybridmc_0_000_100", "hybridmc_0_000_100_2.2", "hybridmc_0_000_100_3.0"]
    assert len(names) == 3 + 2 and len(summary) == 5
    import json
    for n, rc, st in (("hybridmc_0_000_100_3.0", 3.0, None), ("hybridmc_0_000_100_2.2", 2.2, 3.0),
                      ("hybridmc_0_000_100", 1.5, 2.2)):
        with open(os.path.join(out, n + ".json")) as f:
            d = json.load(f)
        assert d["rc"] == rc and d.get("stair") == st and "p_rc" not in d
        assert (d.get("stair_bonds") == [bp]) == (st is not None)
        with File(os.path.join(out, n + ".h5")) as f:
            assert np.isfinite(f["s_bias"]).all() and len(f["unique_config"]["pos"]) == 1
            # the structure handed on has the bond inside this stage's rc
            pos = f["unique_config"]["pos"][0]
            dd = pos[bp[1]] - pos[bp[0]]
            dd -= master["length"] * np.round(dd / master["length"])
            assert np.linalg.norm(dd) < rc


def test_restart_pack_continues_without_the_earlier_files(tmp_path):
    """restart.npz (written after every layer) carries the hand-off structures and s_bias
    values: a second run continues with the next layer although the per-transition files of
    the finished layer are gone (a campaign interrupted on a machine whose disk is lost)"""
    master = load_config("test")
    out = str(tmp_path / "camp")
    kw = dict(replicas=48, iters_per_round=2, max_rounds=2, hist_bins=256, wl=False, eq_iters=2,
              log=lambda *a: None, warm_start=True)
    s1, _, _ = run_campaign(master, out, max_layers=1, **kw)
    assert len(s1) == 3 and os.path.exists(os.path.join(out, "restart.npz"))
    with open(os.path.join(out, "diff_s_bias.csv")) as f:
        assert len(list(csv.reader(f))) == 3
    out2 = str(tmp_path / "camp2")
    os.makedirs(out2)
    s2, _, _ = run_campaign(master, out2, max_layers=2, restart=os.path.join(out, "restart.npz"), **kw)
    assert len(s2) == 6 and all(r[0].count("1") == 1 for r in s2)
    names = sorted(f for f in os.listdir(out2) if f.endswith(".h5"))
    assert len(names) == 6 and all(n.startswith("hybridmc_1_") for n in names)


def test_list_capacity_overflow_is_flagged_and_curable():
    """HMC_ERR_LIST_OVERFLOW: a list of live nonlocal predictions that is too small raises the
    flag (the replica's trajectory is then void); the default capacity and a larger one do
    not, and give identical trajectories (the capacity is not part of the arithmetic)"""
    from common import transition_params
    from hybridmc_b200.engine import ERR_LIST_OVERFLOW, PHASE_MAIN, Engine, init_pos_chain
    p = transition_params("crambin", 0)
    pos = init_pos_chain(p)
    outs = []
    for cap in (8, 0, 1000):
        with Engine(p, 6) as e:
            e.set_list_capacity(cap)
            assert e.get_list_capacity() >= 8
            e.set_positions(np.broadcast_to(pos, (6,) + pos.shape))
            e.init_config_from_positions()
            e.init_s()
            e.seed(5, 6)
            e.reset_clocks()
            e.run(PHASE_MAIN, 0, 1)
            err = int(np.bitwise_or.reduce(e.get_counts()[3]))
            outs.append((err, e.get_positions().copy()))
    assert outs[0][0] & ERR_LIST_OVERFLOW
    assert outs[1][0] == 0 and outs[2][0] == 0
    assert np.array_equal(outs[1][1].view(np.uint64), outs[2][1].view(np.uint64))
