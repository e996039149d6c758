"""CPU, build container only: the C oracle against the UNMODIFIED reference compiled
into oracle/_ref (skipped when it is absent, e.g. on the GPU box before build)."""
import ctypes as C

import numpy as np
import pytest

from common import OracleRun, load_config, transition_params
from hybridmc_b200.params import params_from_json
from oracle_lib import OrcMainResult, RefMainResult, Sim


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


@pytest.mark.parametrize("name,index,nsteps,phase", [
    ("8bead_1bond", None, 12, 2), ("test", 0, 4, 2), ("test", 2, 4, 0), ("crambin", 0, 2, 2),
    ("crambin", 7, 2, 0), ("50bead_9bond", 3, 2, 2)])
def test_trace_matches_reference(oracle, reflib, name, index, nsteps, phase):
    p = params_from_json(load_config(name)) if index is None else transition_params(name, index)
    o = OracleRun(p, oracle)
    pos = o.init_pos()
    r = Sim(reflib, "ref", p)
    r.set_pos(pos)
    r.init_s()
    o.sim.init_s()
    r.reset_clocks()
    o.sim.reset_clocks()
    r.enable_trace(100000)
    o.sim.enable_trace(100000)
    normals = o.normals(nsteps)
    for k in range(nsteps):
        assert r.md_step(phase, k, normals[k]) == o.sim.md_step(phase, k, normals[k])
    a, na = r.get_trace()
    b, nb = o.sim.get_trace()
    assert na == nb and a.tobytes() == b.tobytes()
    assert np.array_equal(bits(r.get_pos()), bits(o.sim.get_pos()))
    assert r.event_counts()[4] == o.sim.event_counts()[4]  # same number of heap pops


def test_wang_landau_steps_match_reference(oracle, reflib):
    p = params_from_json(load_config("8bead_1bond"))
    o = OracleRun(p, oracle)
    pos = o.init_pos()
    r = Sim(reflib, "ref", p)
    for s in (r, o.sim):
        s.set_pos(pos)
        s.init_s()
        s.reset_clocks()
    normals = o.normals(40)
    for k in range(40):
        assert r.md_step(1, k, normals[k]) == o.sim.md_step(1, k, normals[k])
    assert np.array_equal(bits(r.get_s_bias()), bits(o.sim.get_s_bias()))
    assert np.array_equal(bits(r.get_pos()), bits(o.sim.get_pos()))


def test_staircase_and_permanent_bonds_match_reference(oracle, reflib):
    """stair / p_rc / permanent-bond branches of if_coll and get_rc2_* (diff_s_bias_stair path)"""
    d = load_config("test")
    d.update(transient_bonds=[[2, 6]], permanent_bonds=[], rc=2.2, stair=3.0, p_rc=1.5,
             stair_bonds=[[2, 6]])
    p = params_from_json(d)
    o = OracleRun(p, oracle)
    pos = o.init_pos()
    r = Sim(reflib, "ref", p)
    for s in (r, o.sim):
        s.set_pos(pos)
        s.init_s()
        s.init_config_from_pos()
        s.reset_clocks()
        s.enable_trace(50000)
    normals = o.normals(4)
    for k in range(4):
        assert r.md_step(2, k, normals[k]) == o.sim.md_step(2, k, normals[k])
    a, na = r.get_trace()
    b, nb = o.sim.get_trace()
    assert na == nb and a.tobytes() == b.tobytes()


def test_whole_program_matches_reference(oracle, reflib):
    p = params_from_json(load_config("8bead_1bond", total_iter=30, total_iter_eq=6, seeds=[11, 5]))
    rr, ro = RefMainResult(), OrcMainResult()
    assert reflib.ref_run_main(C.byref(p), None, 2, C.byref(rr)) == 0
    assert oracle.orc_run_main(C.byref(p), None, 2, 0, None, None, C.byref(ro)) == 0
    assert list(rr.s_bias)[:2] == list(ro.s_bias)[:2]
    assert list(rr.config_count)[:2] == list(ro.config_count)[:2]
    assert rr.gtest_iters == ro.gtest_iters


def test_reference_unit_tests_pass():
    """the reference's own 22 Boost.Test cases, compiled against oracle/shim"""
    import os
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle",
                       "_ref", "hybridmc_test")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/hybridmc_test not built")
    res = subprocess.run([exe], capture_output=True, text=True, cwd="/tmp")
    assert res.returncode == 0, res.stdout[-2000:]
    assert "cases 22" in res.stdout and "failures 0" in res.stdout


@pytest.mark.parametrize("index", [1, 4, 9])
def test_bench_main_phase_replay_is_the_reference_run(oracle, reflib, index):
    """bench.py's CPU leg times the reference's own run_trajectory calls
    (ref_bench_main_phase) and takes the event count of that run from the oracle's replay
    (orc_bench_main_phase): the two must be the same trajectory, bit for bit, from the
    committed start structures of every lattice layer."""
    import os
    from hybridmc_b200.params import transition_jsons
    master = load_config("test")
    layer, bits_in, bits_out, d = transition_jsons(master)[index]
    st = np.load(os.path.join(os.path.dirname(__file__), "golden", "start_structures_test.npz"))
    start = np.ascontiguousarray(st["".join("1" if b else "0" for b in bits_in)])
    p = params_from_json(dict(d, seeds=[77 + index, 1]))
    out = []
    for lib, fn in ((reflib, "ref_bench_main_phase"), (oracle, "orc_bench_main_phase")):
        sec, cfg = C.c_double(), C.c_uint64()
        pos = np.zeros((p.nbeads, 3))
        extra = []
        if fn.startswith("orc"):
            coll, cells = C.c_uint64(), C.c_uint64()
            extra = [C.byref(coll), C.byref(cells)]
        rc = getattr(lib, fn)(C.byref(p), start.ctypes.data_as(C.c_void_p), C.c_uint(1), C.c_uint(2),
                              C.c_uint(2), C.byref(sec), pos.ctypes.data_as(C.c_void_p),
                              C.byref(cfg), *extra)
        assert rc == 0
        out.append((pos, cfg.value))
    assert np.array_equal(bits(out[0][0]), bits(out[1][0])) and out[0][1] == out[1][1]
    assert coll.value > 4 * 8 * 1000 and cells.value > 0
