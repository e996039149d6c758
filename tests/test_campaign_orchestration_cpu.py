"""CPU: the campaign orchestration of hybridmc_b200/run_campaign.py (layer order, structure and
s_bias hand-off, per-transition failure + retry with a larger prediction list, sticky list
capacity, restart pack, progressive diff_s_bias.csv, staircase stages, output files) with the
GPU ensemble program replaced by a stand-in that records how it was called.  The arithmetic is
covered by the GPU tests; this covers the control flow around it."""
import csv
import os

import numpy as np
import pytest

from common import load_config
from hybridmc_b200 import run_campaign as rc
from hybridmc_b200.engine import ERR_LIST_OVERFLOW
from hybridmc_b200.output import File


class FakeEnsemble:
    """stands in for campaign.run_transition_ensemble: deterministic outputs, optional
    list-overflow failure of chosen transitions while the list capacity is too small"""

    def __init__(self, overflow_until_cap=None, overflow_layers=()):
        self.calls = []
        self.overflow_until_cap = overflow_until_cap
        self.overflow_layers = set(overflow_layers)

    def __call__(self, params, start_pos, R, **kw):
        T = len(params)
        nb = params[0].nbeads
        layer = params[0].n_permanent
        self.calls.append(dict(T=T, R=R, layer=layer, list_capacity=kw["list_capacity"],
                               start_s_bias=None if kw["start_s_bias"] is None else kw["start_s_bias"].copy(),
                               iters=kw["iters_per_round"], seed=kw["seed"], stair=bool(params[0].has_stair),
                               rc=params[0].rc))
        err = np.zeros(T, np.uint32)
        if layer in self.overflow_layers and kw["list_capacity"] < (self.overflow_until_cap or 0):
            err[0] = ERR_LIST_OVERFLOW | 32
        s = np.zeros((T, 2))
        # a recognisable value: 3 + layer + 0.01 * (index of the transient bond in the bond list)
        for t, p in enumerate(params):
            tb = p.bond_list("transient")[0]
            s[t, 0] = 3.0 + layer + 0.01 * sorted(p.bond_list("nonlocal")).index(tb)
        pos = [np.full((nb, 3), float(layer + 1)) + t for t in range(T)]
        return dict(s_bias=s, counts=np.full((T, 2), 50, np.uint64), rounds=np.ones(T, int),
                    accepted=np.ones(T, bool), err=err, failed=(err & ~np.uint32(32)) != 0,
                    formed=np.ones(T, np.uint64), broken=np.ones(T, np.uint64),
                    unique_pos=pos, unique_vel=[np.zeros((nb, 3))] * T,
                    counts_rounds=[np.full((1, 2), 50, np.uint64)] * T,
                    events=np.array([1000 * T, 10], np.uint64), seconds=0.1, seconds_device=0.1,
                    collectives=0, dist=[np.ones((4, params[0].n_nonlocal))] * T,
                    config_int=[np.zeros(3, np.uint64)] * T,
                    dist_hist=np.ones((T, max(q.n_nonlocal for q in params), 2, kw["hist_bins"]), np.uint64))


@pytest.fixture
def fake(monkeypatch):
    f = FakeEnsemble()
    monkeypatch.setattr(rc, "run_transition_ensemble", f)
    return f


def test_layers_run_in_order_with_hand_off_and_warm_start(tmp_path, fake):
    master = load_config("test")
    out = str(tmp_path / "c")
    summary, _, events = rc.run_campaign(master, out, replicas=64, hist_bins=16, wl=False,
                                         warm_start=True, fill=1000, log=lambda *a: None)
    assert [c["layer"] for c in fake.calls] == [0, 1, 2] and [c["T"] for c in fake.calls] == [3, 6, 3]
    assert len(summary) == 12 and events == 12000
    # thin layers are topped up to `fill` replicas, the rounds keep the nominal length
    assert [c["R"] for c in fake.calls] == [334, 167, 334]
    assert all(c["iters"] == -(-master["total_iter"] // 64) for c in fake.calls)
    # layer 0 starts from init_s, deeper layers from the value of the same bond one layer earlier
    assert (fake.calls[0]["start_s_bias"][:, 0] == 3.0).all()
    assert np.allclose(np.sort(fake.calls[1]["start_s_bias"][:, 0]), [3.0, 3.0, 3.01, 3.01, 3.02, 3.02])
    assert np.allclose(np.sort(fake.calls[2]["start_s_bias"][:, 0]), [4.0, 4.01, 4.02])
    rows = list(csv.reader(open(os.path.join(out, "diff_s_bias.csv"))))
    assert len(rows) == 12 and rows[0][0] == "000"
    names = sorted(n for n in os.listdir(out) if n.endswith(".h5"))
    assert len(names) == 12
    with File(os.path.join(out, names[-1])) as f:  # a layer-2 transition: two permanent bonds
        assert f["config_count"].shape == (2, 2) and not f["config_count"][0].any()
        assert len(f.attrs["permanent_bonds"]) == 2 and int(f.attrs["replicas"]) == 334
        assert f["unique_config"]["pos"].shape == (1, master["nbeads"], 3)
    # a second pass finds every file and runs nothing
    n = len(fake.calls)
    assert rc.run_campaign(master, out, replicas=64, hist_bins=16, wl=False, log=lambda *a: None)[0] == []
    assert len(fake.calls) == n


def test_failed_transition_is_rerun_with_a_larger_list_and_the_capacity_sticks(tmp_path, monkeypatch):
    master = load_config("test")
    nb = master["nbeads"]
    f = FakeEnsemble(overflow_until_cap=8 * nb, overflow_layers=[1])
    monkeypatch.setattr(rc, "run_transition_ensemble", f)
    out = str(tmp_path / "c")
    summary, _, _ = rc.run_campaign(master, out, replicas=32, hist_bins=8, wl=False, log=lambda *a: None)
    assert len(summary) == 12
    calls = [(c["layer"], c["T"], c["list_capacity"]) for c in f.calls]
    # layer 1: 6 transitions, one overflows -> that one alone again with 8 nb; layer 2 starts there
    assert calls == [(0, 3, 0), (1, 6, 0), (1, 1, 8 * nb), (2, 3, 8 * nb)]
    assert f.calls[2]["seed"] != f.calls[1]["seed"]  # the rerun draws other streams
    assert len([n for n in os.listdir(out) if n.endswith(".h5")]) == 12


def test_transition_that_keeps_failing_stops_the_campaign(tmp_path, monkeypatch):
    f = FakeEnsemble(overflow_until_cap=10 ** 9, overflow_layers=[0])
    monkeypatch.setattr(rc, "run_transition_ensemble", f)
    with pytest.raises(RuntimeError, match="still raise replica error flags after 3 attempts"):
        rc.run_campaign(load_config("test"), str(tmp_path / "c"), replicas=32, hist_bins=8, wl=False,
                        log=lambda *a: None)
    nb = load_config("test")["nbeads"]
    assert [c["list_capacity"] for c in f.calls] == [0, 8 * nb, nb * nb]
    # the two transitions that were fine have their files, the failed one has none
    assert len([n for n in os.listdir(str(tmp_path / "c")) if n.endswith(".h5")]) == 2


def test_restart_pack_skips_finished_layers(tmp_path, fake):
    master = load_config("test")
    a, b = str(tmp_path / "a"), str(tmp_path / "b")
    rc.run_campaign(master, a, replicas=32, hist_bins=8, wl=False, max_layers=2, warm_start=True,
                    log=lambda *a_: None)
    n = len(fake.calls)
    os.makedirs(b)
    s2, _, _ = rc.run_campaign(master, b, replicas=32, hist_bins=8, wl=False, warm_start=True,
                               restart=os.path.join(a, "restart.npz"), log=lambda *a_: None)
    assert [c["layer"] for c in fake.calls[n:]] == [2] and len(s2) == 3
    # ... with the s_bias values of layer 1 as start values and its structures as start structures
    assert np.allclose(np.sort(fake.calls[-1]["start_s_bias"][:, 0]), [4.0, 4.01, 4.02])


def test_staircase_stages_run_as_separate_ensembles(tmp_path, fake):
    master = load_config("test")
    bp = sorted(sorted(b) for b in master["nonlocal_bonds"])[0]
    summary, _, _ = rc.run_campaign(master, str(tmp_path / "s"), replicas=32, hist_bins=8, wl=False,
                                    max_layers=1, stair_bp=bp, stair_rc=[3.0, 2.2, 1.5],
                                    log=lambda *a: None)
    # stage 0: all three transitions (the staircase one at rc = 3.0), stages 1 and 2: it alone
    assert [(c["T"], c["stair"]) for c in fake.calls] == [(3, False), (1, True), (1, True)]
    assert [c["rc"] for c in fake.calls[1:]] == [2.2, 1.5]
    assert len(summary) == 5
    from hybridmc_b200.campaign import collect_stair
    assert len(collect_stair(summary)) == 3


def test_non_zero_rank_computes_everything_and_writes_nothing(tmp_path, fake):
    """multi-rank campaigns: every rank runs all transitions with replicas / world replicas and
    keeps the full summary and hand-off state; only rank 0 writes files"""
    class FakeComm:
        def bcast(self, arr, root=0):
            return arr
    master = load_config("test")
    out = str(tmp_path / "r1")
    summary, _, _ = rc.run_campaign(master, out, replicas=64, hist_bins=8, wl=False, rank=1, world=4,
                                    comm=FakeComm(), log=lambda *a: None)
    assert len(summary) == 12 and [c["R"] for c in fake.calls] == [16, 16, 16]
    assert os.listdir(out) == []
