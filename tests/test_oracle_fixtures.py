"""CPU: the C oracle against fixtures generated from the UNMODIFIED reference
(tests/golden/ref_fixtures.npz, made by tests/golden/make_fixtures.py): RNG restatement,
init_pos, event traces (bit-exact), crankshaft, whole program."""
import ctypes as C
import os

import numpy as np
import pytest

from common import OracleRun, load_config, transition_params
from hybridmc_b200.params import params_from_json
from oracle_lib import OrcMainResult, OrcMt

HERE = os.path.dirname(os.path.abspath(__file__))
FX = np.load(os.path.join(HERE, "golden", "ref_fixtures.npz"))


def params_of(name):
    return params_from_json(load_config(name)) if name == "8bead_1bond" else transition_params(name, 0)


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def test_mt19937_seed_seq_and_distributions(oracle):
    p = params_of("8bead_1bond")
    o = OracleRun(p, oracle)
    words = np.array([oracle.orc_mt_next(C.byref(o.g)) & 0xffffffff for _ in range(64)], dtype=np.uint32)
    assert np.array_equal(words, FX["rng_words_seed_3_1"])
    o = OracleRun(p, oracle)
    n = np.zeros(48)
    oracle.orc_mt_normals(C.byref(o.g), C.c_double(0.5), n.ctypes.data_as(C.c_void_p), 48)
    assert np.array_equal(bits(n), bits(FX["rng_normals_seed_3_1_sigma_0p5"]))
    w, st, ct = o.tape(64)
    assert np.array_equal(w, FX["rng_tape_words"])
    assert np.array_equal(bits(st), bits(FX["rng_tape_sin"]))
    assert np.array_equal(bits(ct), bits(FX["rng_tape_cos"]))


@pytest.mark.parametrize("name", ["8bead_1bond", "test", "crambin", "50bead_9bond"])
def test_init_pos_trace_crankshaft(oracle, name):
    p = params_of(name)
    o = OracleRun(p, oracle)
    pos = o.init_pos()
    assert np.array_equal(bits(pos), bits(FX["initpos_" + name]))
    o.sim.init_s()
    o.sim.reset_clocks()
    normals = FX["trace_%s_normals" % name]
    o.sim.enable_trace(len(FX["trace_%s_type" % name]) + 16)
    mine = o.normals(len(normals))
    assert np.array_equal(bits(mine), bits(normals))
    for k in range(len(normals)):
        assert o.sim.md_step(0, k, normals[k]) == 0
    tr, ntot = o.sim.get_trace()
    assert ntot == int(FX["trace_%s_total" % name][0])
    assert np.array_equal(tr["type"].astype(np.uint8), FX["trace_%s_type" % name])
    assert np.array_equal(tr["i"].astype(np.uint8), FX["trace_%s_i" % name])
    assert np.array_equal(tr["j"].astype(np.uint8), FX["trace_%s_j" % name])
    assert np.array_equal(tr["t"].view(np.uint64), FX["trace_%s_tbits" % name])
    assert np.array_equal(bits(o.sim.get_pos()), bits(FX["trace_%s_pos_end" % name]))
    if name == "8bead_1bond":
        assert ntot >= 10000  # north_star prefix length
    # 40 crankshaft moves from the end state, tape from the same engine state
    words, st, ct = o.tape(4096)
    err, used = o.crankshaft(0, 40, words, st, ct)
    assert err == 0
    cfg, ref_used = [int(x) for x in FX["crank_%s_cfg_used" % name]]
    assert used == ref_used
    assert o.sim.get_config() == cfg
    assert np.array_equal(bits(o.sim.get_pos()), bits(FX["crank_%s_pos" % name]))


@pytest.mark.parametrize("name", ["8bead_1bond", "test"])
def test_whole_program(oracle, name):
    gi, acc, ti, te, cap = [int(x) for x in FX["main_%s_meta" % name]]
    p = params_from_json(load_config(name, total_iter=ti, total_iter_eq=te))
    r = OrcMainResult()
    assert oracle.orc_run_main(C.byref(p), None, cap, 0, None, None, C.byref(r)) == 0
    ns = p.nstates
    assert np.array_equal(bits(np.array(list(r.s_bias)[:ns])), bits(FX["main_%s_s_bias" % name]))
    assert list(r.config_count)[:ns] == [int(x) for x in FX["main_%s_counts" % name]]
    assert (r.gtest_iters, r.accepted) == (gi, acc)


def test_empty_and_edge_cases(oracle):
    """zero-length step (step index 0: only t <= 0 events), no nonlocal bonds, tape underrun"""
    p = params_from_json(load_config("8bead_1bond", nonlocal_bonds=[], transient_bonds=[]))
    o = OracleRun(p, oracle)
    o.init_pos()
    o.sim.init_s()
    o.sim.reset_clocks()
    o.sim.enable_trace(64)
    nrm = o.normals(1)
    assert o.sim.md_step(2, 0, nrm[0]) == 0
    tr, n = o.sim.get_trace()
    assert all(t <= 0.0 for t in tr["t"])  # bead 0 starts on a cell wall (pos 0,0,0)
    words, st, ct = o.tape(2)
    err, used = o.crankshaft(0, 1, words, st, ct)
    assert err & 16 and used == 0  # HMC_ERR_TAPE_UNDERRUN, move rolled back
