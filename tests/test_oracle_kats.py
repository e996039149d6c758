"""CPU: pin the C oracle against the known-answer vectors of the reference's own unit
tests (tests/golden/reference_kats.json, transcribed from src/hybridmc_test.cc)."""
import ctypes as C
import json
import math
import os

import numpy as np
import pytest

from common import load_config
from hybridmc_b200.params import HmcParams, params_from_json

HERE = os.path.dirname(os.path.abspath(__file__))
KATS = json.load(open(os.path.join(HERE, "golden", "reference_kats.json")))


def close(a, b, tol=1e-10):
    if b == 0:
        return abs(a) <= tol
    return abs(a - b) <= tol * abs(a) and abs(a - b) <= tol * abs(b) or a == b


def test_inner_collision_time(oracle):
    k = KATS["inner_collision_time"]
    oracle.orc_t_until_inner_coll.argtypes = [C.c_double] * 7 + [C.POINTER(C.c_double)]
    for c in k["cases"]:
        t = C.c_double(0.0)
        hit = oracle.orc_t_until_inner_coll(*c["r"], *c["v"], k["sigma2"], C.byref(t))
        assert bool(hit) == c["hit"]
        if c["hit"]:
            assert close(t.value, c["t"])


def test_outer_collision_time(oracle):
    k = KATS["outer_collision_time"]
    oracle.orc_t_until_outer_coll.argtypes = [C.c_double] * 7 + [C.POINTER(C.c_double)]
    for c in k["cases"]:
        t = C.c_double(0.0)
        oracle.orc_t_until_outer_coll(*c["r"], *c["v"], k["sigma"] * k["sigma"], C.byref(t))
        assert close(t.value, c["t"])


def test_collision_velocities(oracle):
    k = KATS["collision_velocities"]
    oracle.orc_v_after_coll.argtypes = [C.c_double] * 4 + [C.c_void_p, C.c_void_p, C.c_int,
                                                           C.c_double, C.c_double]
    for c in k["cases"]:
        vi = np.array(c["vi"], dtype=float)
        vj = np.array(c["vj"], dtype=float)
        has = c["dS"] is not None
        oracle.orc_v_after_coll(*c["r"], c["sigma2"], vi.ctypes.data, vj.ctypes.data, int(has),
                                c["dS"] if has else 0.0, k["m"])
        for a, b in zip(list(vi) + list(vj), c["vi_out"] + c["vj_out"]):
            assert close(a, b)


def test_boundary_conditions(oracle):
    k = KATS["boundary_conditions"]
    oracle.orc_minpos.argtypes = [C.c_double, C.c_void_p]
    oracle.orc_mindist.argtypes = [C.c_double, C.c_void_p]
    for c in k["minpos"]:
        x = np.array(c["in"], dtype=float)
        oracle.orc_minpos(k["l"], x.ctypes.data)
        assert all(close(a, b) for a, b in zip(x, c["out"]))
    for c in k["mindist"]:
        x = np.array(c["in"], dtype=float)
        oracle.orc_mindist(k["l"], x.ctypes.data)
        assert all(close(a, b) for a, b in zip(x, c["out"]))


def test_cell_crossing_time(oracle):
    k = KATS["cell_crossing_time"]
    oracle.orc_t_until_cell.argtypes = ([C.c_double] * 4 + [C.c_uint, C.c_double] + [C.c_double] * 3 +
                                        [C.c_uint] * 3 + [C.c_void_p] * 3)
    for c in k["cases"]:
        dt, wall = C.c_double(0.0), C.c_uint(0)
        new = (C.c_uint * 3)()
        cr = oracle.orc_t_until_cell(*c["pos"], k["l"], k["ncell"], k["lcell"], *c["v"], *c["cell"],
                                     C.addressof(dt), C.addressof(wall), C.addressof(new))
        assert bool(cr) == c["crossing"]
        if c["crossing"]:
            assert dt.value == c["dt"]  # BOOST_CHECK_EQUAL in the reference
            assert wall.value == c["wall"]
            assert list(new) == c["new"]


def _sim_for(oracle, nb, l):
    d = load_config("8bead_1bond", nbeads=nb, length=l, nonlocal_bonds=[], transient_bonds=[])
    p = params_from_json(d)
    from oracle_lib import Sim
    return Sim(oracle, "orc", p), p


def test_rodrigues(oracle):
    k = KATS["rodrigues"]
    pos = np.array(k["pos"], dtype=float)
    sim, p = _sim_for(oracle, len(pos), k["l"])
    sim.set_pos(pos)
    out = np.zeros_like(pos)
    oracle.orc_rodrigues.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_uint, C.c_void_p]
    oracle.orc_rodrigues(sim.h, math.sin(math.pi), math.cos(math.pi), k["ind"], out.ctypes.data)
    exp = np.array(k["out"], dtype=float)
    assert np.allclose(out, exp, rtol=1e-10, atol=1e-10)
    sim.close()


def test_compute_entropy(oracle):
    k = KATS["compute_entropy"]
    n = 1 << k["nbonds"]
    sb = None
    oracle.orc_compute_entropy.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_double]
    for c in k["cases"]:
        if not c["chain"]:
            sb = np.array([3.0 * (k["nbonds"] - bin(i).count("1")) for i in range(n)])  # init_s
        cc = np.bincount(c["config_int"], minlength=n).astype(np.uint64)
        oracle.orc_compute_entropy(cc.ctypes.data, sb.ctypes.data, n, k["pos_scale"], k["neg_scale"])
        for a, b in zip(sb, c["s_bias"]):
            assert close(a, b)


def test_g_val_and_chi2(oracle):
    k = KATS["g_val"]
    cc = np.array(k["config_count"], dtype=np.uint64)
    oracle.orc_compute_g_val.argtypes = [C.c_void_p, C.c_int]
    assert close(oracle.orc_compute_g_val(cc.ctypes.data, len(cc)), k["g"])
    q = KATS["chi_squared"]
    oracle.orc_g_test.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
    crit = C.c_double()
    two = np.array([50, 50], dtype=np.uint64)
    assert oracle.orc_g_test(two.ctypes.data, 2, q["sig"], None, C.addressof(crit), None) == 1
    assert abs(crit.value - q["g_crit"]) < 1e-9


def test_rc2_inner_outer(oracle):
    k = KATS["rc2_outer"]
    oracle.orc_rc2_inner.argtypes = [C.POINTER(HmcParams), C.c_int]
    oracle.orc_rc2_outer.argtypes = [C.POINTER(HmcParams), C.c_int, C.c_int]
    for c in k["cases"]:
        p = HmcParams()
        p.rc = c["rc"]
        if c["stair"] is not None:
            p.stair, p.has_stair = c["stair"], 1
        if c["p_rc"] is not None:
            p.p_rc, p.has_p_rc = c["p_rc"], 1
        assert oracle.orc_rc2_outer(C.byref(p), int(c["t_nonbonded"]), int(c["p_bond"])) == c["outer"]
        if "inner" in c:
            assert oracle.orc_rc2_inner(C.byref(p), 0) == c["inner"]
        if "inner_p" in c:
            assert oracle.orc_rc2_inner(C.byref(p), 1) == c["inner_p"]
