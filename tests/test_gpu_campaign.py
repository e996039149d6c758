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
    summary, seconds = run_campaign(master, out, replicas=48, iters_per_round=2, max_rounds=2,
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
            assert cc.dtype == np.uint64 and cc.shape == (2, 2) and not cc[0].any() and cc[1].all()
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
    with open(os.path.join(out, "diff_s_bias.csv")) as f:
        assert len(list(csv.reader(f))) == 9
    # finished transitions are skipped on a second pass (tools/init_json.py:141-142)
    summary2, _ = run_campaign(master, out, replicas=48, iters_per_round=2, max_rounds=2,
                               max_layers=2, hist_bins=2048, wl=False, eq_iters=2,
                               log=lambda *a: None)
    assert summary2 == []
