"""GPU parity proper, beyond the MD event loop: crankshaft moves, sampling
(config_count / config_int / dist / unique_config), Wang-Landau and the WHOLE reference
program in reference-exact mode — all bit-exact against the CPU oracle — plus the
committed fixtures generated from the unmodified reference."""
import ctypes as C
import os

import numpy as np
import pytest

from common import OracleRun, first_trace_mismatch, load_config, transition_params
from hybridmc_b200.engine import (PHASE_EQ, PHASE_MAIN, PHASE_WL, Engine, run_program_exact)
from hybridmc_b200.params import params_from_json
from oracle_lib import OrcMainResult

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
FX = np.load(os.path.join(HERE, "golden", "ref_fixtures.npz"))


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def params_of(name):
    return params_from_json(load_config(name)) if name == "8bead_1bond" else transition_params(name, 0)


@pytest.mark.parametrize("name", ["8bead_1bond", "test", "crambin", "50bead_9bond"])
def test_trace_and_crankshaft_match_reference_fixture(name):
    """engine vs the trace the UNMODIFIED reference produced (tests/golden/ref_fixtures.npz):
    event type, pair and event-time bits; then 40 crankshaft moves from the end state."""
    p = params_of(name)
    normals = FX["trace_%s_normals" % name]
    cap = len(FX["trace_%s_type" % name]) + 16
    o = OracleRun(p)  # only for the mt19937 stream that follows the velocity draws
    o.init_pos()
    assert np.array_equal(bits(o.normals(len(normals))), bits(normals))
    with Engine(p, 1) as e:
        e.set_positions(FX["initpos_" + name][None])
        e.init_s()
        e.reset_clocks()
        e.enable_trace(0, cap)
        e.md_steps_injected(PHASE_EQ, 0, normals[None])
        tr, n = e.get_trace()
        assert n == int(FX["trace_%s_total" % name][0])
        assert np.array_equal(tr["type"].astype(np.uint8), FX["trace_%s_type" % name])
        assert np.array_equal(tr["i"].astype(np.uint8), FX["trace_%s_i" % name])
        assert np.array_equal(tr["j"].astype(np.uint8), FX["trace_%s_j" % name])
        assert np.array_equal(tr["t"].view(np.uint64), FX["trace_%s_tbits" % name])
        assert np.array_equal(bits(e.get_positions()[0]), bits(FX["trace_%s_pos_end" % name]))
        words, st, ct = o.tape(4096)
        used = e.crankshaft_injected(0, 40, words, st, ct)
        cfg, ref_used = [int(x) for x in FX["crank_%s_cfg_used" % name]]
        assert int(used[0]) == ref_used
        assert int(e.get_config()[0]) == cfg
        assert np.array_equal(bits(e.get_positions()[0]), bits(FX["crank_%s_pos" % name]))
        assert e.get_counts()[3][0] == 0


def test_main_phase_sampling_matches_oracle():
    """run_trajectory body: dist rows, config_count, config_int order (MD + crankshaft
    samples), unique_config, bond counters — 6 iterations with crankshaft tapes."""
    p = params_from_json(load_config("8bead_1bond"))
    o = OracleRun(p)
    pos = o.init_pos()
    o.sim.init_s()
    o.sim.set_s_bias(np.array([1.0, 0.0]))  # shallow bias so that both states are visited
    o.sim.reset_clocks()
    o.lib.orc_enable_samples(o.sim.h, C.c_uint(256), C.c_uint(256))
    with Engine(p, 1) as e:
        e.set_positions(pos[None])
        e.set_s_bias(np.array([1.0, 0.0]))
        e.reset_clocks()
        e.enable_samples(256, 256)
        for it in range(6):
            normals = o.normals(p.nsteps)
            for k in range(p.nsteps):
                assert o.sim.md_step(PHASE_MAIN, it * p.nsteps + k, normals[k]) == 0
            e.md_steps_injected(PHASE_MAIN, it * p.nsteps, normals[None])
            words, st, ct = o.tape(2048)
            err, used_o = o.crankshaft(0, p.mc_moves, words, st, ct)
            assert err == 0
            used = e.crankshaft_injected(0, p.mc_moves, words, st, ct)
            assert int(used[0]) == used_o
            o.discard(used_o)
            assert np.array_equal(bits(e.get_positions()[0]), bits(o.sim.get_pos()))
            assert int(e.get_config()[0]) == o.sim.get_config()
        cc, formed, broken, err = e.get_counts()
        occ = np.zeros(2, dtype=np.uint64)
        of, ob = C.c_uint64(), C.c_uint64()
        o.lib.orc_get_counts(o.sim.h, occ.ctypes.data_as(C.c_void_p), 2, C.byref(of), C.byref(ob))
        assert list(cc[0]) == list(occ) and int(formed[0]) == of.value and int(broken[0]) == ob.value
        assert err[0] == 0
        dist, nrows = e.get_dist()
        od = np.zeros((256, 1))
        on = o.lib.orc_get_dist(o.sim.h, od.ctypes.data_as(C.c_void_p))
        assert int(nrows[0]) == on == 6 * p.nsteps
        assert np.array_equal(bits(dist[0, :on]), bits(od[:on]))
        ci, nci = e.get_config_int()
        oc = np.zeros(256, dtype=np.uint64)
        onc = o.lib.orc_get_config_int(o.sim.h, oc.ctypes.data_as(C.c_void_p))
        assert int(nci[0]) == onc and np.array_equal(ci[0, :onc], oc[:onc])
        upos, uvel, found = e.get_unique_config()
        op_, ov_ = np.zeros((p.nbeads, 3)), np.zeros((p.nbeads, 3))
        ofound = o.lib.orc_get_unique_config(o.sim.h, op_.ctypes.data_as(C.c_void_p),
                                             ov_.ctypes.data_as(C.c_void_p))
        assert bool(found[0]) == bool(ofound)
        if ofound:
            assert np.array_equal(bits(upos[0]), bits(op_)) and np.array_equal(bits(uvel[0]), bits(ov_))


def test_wang_landau_matches_oracle():
    p = params_from_json(load_config("8bead_1bond"))
    o = OracleRun(p)
    pos = o.init_pos()
    o.sim.init_s()
    o.sim.reset_clocks()
    normals = o.normals(60)
    for k in range(60):
        assert o.sim.md_step(PHASE_WL, k, normals[k]) == 0
    with Engine(p, 1) as e:
        e.set_positions(pos[None])
        e.init_s()
        e.reset_clocks()
        e.md_steps_injected(PHASE_WL, 0, normals[None])
        assert np.array_equal(bits(e.get_s_bias()[0]), bits(o.sim.get_s_bias()))
        assert np.array_equal(bits(e.get_positions()[0]), bits(o.sim.get_pos()))
        assert int(e.get_config()[0]) == o.sim.get_config()


@pytest.mark.parametrize("name", ["8bead_1bond", "test"])
def test_whole_program_exact_mode_matches_reference(name):
    """EQ -> WL -> G-test loop with std::mt19937(seed_seq(seeds)) on the host and every
    trajectory on the GPU: final s_bias BITS and counts equal the unmodified reference's
    (fixture) and the oracle's."""
    gi, acc, ti, te, cap = [int(x) for x in FX["main_%s_meta" % name]]
    p = params_from_json(load_config(name, total_iter=ti, total_iter_eq=te))
    res, _ = run_program_exact(p, 0, None, cap)
    ns = p.nstates
    assert np.array_equal(bits(np.array(list(res.s_bias)[:ns])), bits(FX["main_%s_s_bias" % name]))
    assert list(res.config_count)[:ns] == [int(x) for x in FX["main_%s_counts" % name]]
    assert (res.gtest_iters, res.accepted, res.err_flags) == (gi, acc, 0)


def test_whole_program_exact_mode_with_dist_matches_oracle():
    p = params_from_json(load_config("8bead_1bond", total_iter=24, total_iter_eq=4, seeds=[7, 9]))
    res, dist = run_program_exact(p, 0, None, 2, dist_rows_cap=1024)
    o = OracleRun(p)
    r = OrcMainResult()
    od = np.zeros((1024, 1))
    on = C.c_uint(0)
    assert o.lib.orc_run_main(C.byref(p), None, 2, 1024, od.ctypes.data_as(C.c_void_p),
                              C.byref(on), C.byref(r)) == 0
    assert list(res.s_bias)[:2] == list(r.s_bias)[:2]
    assert res.gtest_iters == r.gtest_iters and res.collisions == r.collisions
    # the oracle keeps only the last G-test iteration's rows (cursor reset), the program
    # appends all of them like DistWriter: compare the tail
    assert np.array_equal(bits(dist[-on.value:]), bits(od[:on.value]))


def test_tape_underrun_rolls_back():
    p = params_from_json(load_config("8bead_1bond"))
    o = OracleRun(p)
    pos = o.init_pos()
    with Engine(p, 1) as e:
        e.set_positions(pos[None])
        e.init_s()
        words, st, ct = o.tape(8)
        used = e.crankshaft_injected(0, 20, words, st, ct)
        moves = np.zeros(1, dtype=np.uint32)
        e._ck(e.lib.hmc_get_moves_done(e.h, moves.ctypes.data_as(C.c_void_p)))
        o.sim.init_s()
        err, used_o = o.crankshaft(0, 20, words, st, ct)
        assert err & 16 and int(used[0]) == used_o and int(moves[0]) == o.lib.orc_last_moves_done(o.sim.h)
        assert e.get_counts()[3][0] & 16
        assert np.array_equal(bits(e.get_positions()[0]), bits(o.sim.get_pos()))
