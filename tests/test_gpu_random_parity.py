"""GPU parity over the parameter space: randomly drawn chains / bond sets / potentials /
schedules (seeded), and a multi-transition ensemble where every replica has its own
injected stream — all bit-exact against the oracle through the C-ABI."""
import numpy as np
import pytest

from common import OracleRun, first_trace_mismatch, load_config, transition_params
from hybridmc_b200.engine import PHASE_MAIN, Engine
from hybridmc_b200.params import params_from_json, transition_jsons

pytestmark = pytest.mark.gpu


def random_config(seed):
    rng = np.random.default_rng(seed)
    nb = int(rng.integers(5, 27))
    pairs = [(i, j) for i in range(nb) for j in range(i + 3, nb)]
    nbonds = int(rng.integers(1, min(4, len(pairs)) + 1))
    idx = rng.choice(len(pairs), size=nbonds, replace=False)
    bonds = [list(map(int, pairs[k])) for k in idx]
    ntrans = int(rng.integers(1, nbonds + 1))
    d = load_config("8bead_1bond")
    ncell = int(rng.integers(4, 8))
    rc = float(rng.choice([1.5, 1.8, 2.2]))
    use_stair = bool(rng.integers(0, 2)) and ntrans == 1
    stair = rc + float(rng.choice([0.5, 1.0])) if use_stair else None
    lcell_min = (stair or rc) + 0.05
    length = max(float(rng.uniform(8.0, 22.0)), lcell_min * ncell, 0.7 * nb * 1.17)
    d.update(nbeads=nb, nonlocal_bonds=bonds, transient_bonds=bonds[:ntrans], permanent_bonds=[],
             ncell=ncell, length=length, rc=rc, m=float(rng.choice([1.0, 6.0, 2.5])),
             temp=float(rng.choice([1.0, 0.7, 1.6])), del_t=float(rng.choice([3.0, 7.5, 15.0])),
             write_step=int(rng.integers(1, 4)), max_nbonds=int(rng.integers(1, ntrans + 1)),
             seeds=[int(rng.integers(1, 10 ** 6)), int(rng.integers(1, 100))],
             mc_moves=int(rng.integers(1, 12)), mc_write=int(rng.integers(1, 6)))
    if stair is not None:
        d.update(stair=stair, stair_bonds=bonds[:1])
    return d


def stair_start(d, pos):
    """A start structure inside a staircase potential's domain, made the way the campaign
    makes it (tools/init_json.py:93-120): the previous stage — the same transition with
    rc = stair and no outer wall — runs until the bond has formed; its configuration is
    then inside the stair wall of the next stage."""
    prev = {k: v for k, v in d.items() if k not in ("stair", "stair_bonds")}
    prev["rc"] = d["stair"]
    o = OracleRun(params_from_json(prev))
    o.sim.set_pos(pos)
    o.sim.init_config_from_pos()
    o.sim.init_s()
    o.sim.reset_clocks()
    for step in range(20000):
        assert o.sim.md_step(PHASE_MAIN, step, o.normals(1)[0]) == 0
        if o.sim.get_config() == 1:
            return o.sim.get_pos().copy()
    raise AssertionError("the previous stair stage never formed the bond")


@pytest.mark.parametrize("seed", list(range(14)))
def test_random_configuration_parity(seed):
    d = random_config(seed)
    p = params_from_json(d)
    o = OracleRun(p)
    pos = o.init_pos()
    if "stair" in d:
        pos = stair_start(d, pos)
        o.sim.set_pos(pos)
    ns = p.nstates
    sb = np.random.default_rng(seed + 99).normal(size=ns) * 1.5  # arbitrary bias table
    o.sim.set_s_bias(sb)
    o.sim.init_config_from_pos()
    o.sim.reset_clocks()
    o.sim.enable_trace(60000)
    nsteps = 3
    normals = o.normals(nsteps)
    errs = 0
    for s in range(nsteps):
        errs |= o.sim.md_step(PHASE_MAIN, s, normals[s])
    assert errs & 3 == 0  # (a start outside the potential's domain: main.cc:48-56)
    words, st, ct = o.tape(4096)
    cerr, used_o = o.crankshaft(0, p.mc_moves, words, st, ct)
    otr, on = o.sim.get_trace()
    with Engine(p, 1) as e:
        e.set_positions(pos[None])
        e.set_s_bias(sb)
        e.init_config_from_positions()
        e.reset_clocks()
        e.enable_trace(0, 60000)
        e.md_steps_injected(PHASE_MAIN, 0, normals[None])
        gtr, gn = e.get_trace()
        assert gn == on
        k = first_trace_mismatch(otr, gtr)
        assert k is None, "config %s: first mismatch at event %s" % (d, k)
        used = e.crankshaft_injected(0, p.mc_moves, words, st, ct)
        assert int(used[0]) == used_o
        assert np.array_equal(e.get_positions()[0].view(np.uint64), o.sim.get_pos().view(np.uint64))
        assert int(e.get_config()[0]) == o.sim.get_config()
        cc, formed, broken, err = e.get_counts()
        assert int(err[0]) == (errs | cerr)
        assert [int(x) for x in cc[0]] == [int(x) for x in _counts(o, ns)]


def _counts(o, ns):
    import ctypes as C
    cc = np.zeros(ns, dtype=np.uint64)
    f, b = C.c_uint64(), C.c_uint64()
    o.lib.orc_get_counts(o.sim.h, cc.ctypes.data_as(C.c_void_p), ns, C.byref(f), C.byref(b))
    return cc


def test_multi_transition_ensemble_parity():
    """3 transitions x 5 replicas in ONE engine (15 is not a multiple of the groups per
    warp): replica r uses the pair tables of transition r // 5 and its own injected
    velocities; each must equal an oracle run with that transition's parameters."""
    trans = [t for t in transition_jsons(load_config("test")) if t[0] == 0]
    params = [params_from_json(t[3]) for t in trans]
    R, nsteps = 5, 3
    runs, all_pos, all_normals = [], [], []
    for t, p in enumerate(params):
        for r in range(R):
            q = params_from_json(dict(trans[t][3], seeds=[1000 + 10 * t + r, 3]))
            o = OracleRun(q)
            pos = o.init_pos()
            o.sim.init_s()
            o.sim.set_s_bias(np.array([0.5, 0.0]))
            o.sim.reset_clocks()
            nrm = o.normals(nsteps)
            for s in range(nsteps):
                assert o.sim.md_step(PHASE_MAIN, s, nrm[s]) == 0
            runs.append(o)
            all_pos.append(pos)
            all_normals.append(nrm)
    with Engine(params, R) as e:
        e.set_positions(np.stack(all_pos))
        e.set_s_bias(np.array([0.5, 0.0]))
        e.reset_clocks()
        e.md_steps_injected(PHASE_MAIN, 0, np.stack(all_normals))
        gp, gc = e.get_positions(), e.get_config()
        cc, formed, broken, err = e.get_counts()
        assert not err.any()
        for k, o in enumerate(runs):
            assert np.array_equal(gp[k].view(np.uint64), o.sim.get_pos().view(np.uint64)), k
            assert int(gc[k]) == o.sim.get_config()
        for t in range(len(params)):
            tot = sum(_counts(runs[t * R + r], 2) for r in range(R))
            assert [int(x) for x in cc[t]] == [int(x) for x in tot]
