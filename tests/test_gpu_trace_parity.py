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
