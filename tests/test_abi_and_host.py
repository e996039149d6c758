"""CPU: the C-ABI library loads and exports every symbol include/hybridmc_b200.h declares
(no compute calls), the host-side mirrors (params, campaign enumeration, entropy) behave
like the reference's, and the product never falls back to the CPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from common import load_config
from hybridmc_b200 import engine as eng
from hybridmc_b200.params import HmcParams, params_from_json, transition_jsons

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lib():
    from hybridmc_b200.build import build_library
    build_library()
    return eng.load_library()


def test_header_symbols_exported():
    lib = _lib()
    hdr = open(os.path.join(ROOT, "include", "hybridmc_b200.h")).read()
    declared = set(re.findall(r"\b(hmc_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"hmc_params", "hmc_event", "hmc_engine"}
    assert declared == set(eng.EXPORTS), declared ^ set(eng.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.hmc_abi_version() == 2


def test_params_struct_layout_matches_c():
    # sizeof(hmc_params) as the C compiler sees it
    import subprocess
    import tempfile
    src = '#include "hybridmc_b200.h"\n#include <stdio.h>\nint main(){printf("%zu", sizeof(hmc_params));}'
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "s.c"), "w").write(src)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "s.c"), "-o",
                        os.path.join(d, "s")], check=True)
        out = subprocess.run([os.path.join(d, "s")], capture_output=True, text=True).stdout
    assert int(out) == C.sizeof(HmcParams)


def test_ensemble_struct_layouts_match_c():
    """the ctypes mirrors of the ensemble-program structs (engine.py) against the header"""
    import subprocess
    import tempfile
    names = ["hmc_ensemble_options", "hmc_ensemble_result", "hmc_ensemble_totals", "hmc_ensemble_sink"]
    src = ('#include "hybridmc_b200.h"\n#include <stdio.h>\n#include <stddef.h>\nint main(){printf("%zu %zu %zu %zu %zu %zu", '
           + ", ".join("sizeof(%s)" % n for n in names)
           + ', offsetof(hmc_ensemble_options, comm), offsetof(hmc_ensemble_result, unique_pos));}')
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "s.c"), "w").write(src)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "s.c"), "-o",
                        os.path.join(d, "s")], check=True)
        out = [int(x) for x in subprocess.run([os.path.join(d, "s")], capture_output=True,
                                              text=True).stdout.split()]
    assert out[:4] == [C.sizeof(eng.EnsembleOptions), C.sizeof(eng.EnsembleResult),
                       C.sizeof(eng.EnsembleTotals), C.sizeof(eng.EnsembleSink)]
    assert out[4] == eng.EnsembleOptions.comm.offset
    assert out[5] == eng.EnsembleResult.unique_pos.offset


def test_create_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    _lib()
    p = params_from_json(load_config("8bead_1bond"))
    with pytest.raises(eng.EngineError) as ei:
        eng.Engine(p, 1)
    assert "no CPU fallback" in str(ei.value) or "CUDA" in str(ei.value)


def test_create_validation_messages():
    lib = _lib()
    h = C.c_void_p()
    p = params_from_json(load_config("8bead_1bond", ncell=3))
    assert lib.hmc_create(C.byref(p), 1, 1, 0, C.byref(h)) != 0
    assert b"ncell must be at least 4" in lib.hmc_last_error(None)  # main.cc:58-60
    p = params_from_json(load_config("8bead_1bond", ncell=20))
    assert lib.hmc_create(C.byref(p), 1, 1, 0, C.byref(h)) != 0
    assert b"lcell must be at least rc" in lib.hmc_last_error(None)  # main.cc:62-69
    with pytest.raises(RuntimeError):
        params_from_json(load_config("8bead_1bond", stair=1.2))  # main.cc:535-537


def test_params_from_json_mirrors_from_json():
    d = load_config("crambin")
    p = params_from_json(d)
    assert (p.nbeads, p.ncell, p.n_nonlocal, p.n_transient, p.nstates) == (46, 9, 10, 10, 1024)
    assert p.bond_list("nonlocal")[0] == (1, 33)
    d2 = dict(d, nonlocal_bonds=[[33, 1]], transient_bonds=[[33, 1]])
    assert params_from_json(d2).bond_list("transient") == [(1, 33)]  # normalised i<j
    assert not p.has_stair and not p.has_p_rc


def test_transition_enumeration_matches_init_json():
    ts = transition_jsons(load_config("test"))
    assert len(ts) == 12  # 3 bonds: 3 + 6 + 3 transitions
    assert [t[0] for t in ts] == [0] * 3 + [1] * 6 + [2] * 3
    assert ts[0][3]["seeds"] == [1, 1] and ts[0][3]["transient_bonds"] == [[2, 6]]
    assert len(transition_jsons(load_config("crambin"))) == 5120
    assert len(transition_jsons(load_config("50bead_9bond"))) == 2304


def test_entropy_host_functions_match_oracle(oracle):
    _lib()
    rng = np.random.default_rng(0)
    oracle.orc_compute_entropy.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_double]
    oracle.orc_compute_g_val.argtypes = [C.c_void_p, C.c_int]
    oracle.orc_compute_g_val.restype = C.c_double
    oracle.orc_g_test.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
    for n in (2, 2, 4, 8, 1024):
        cc = rng.integers(0, 500, size=n).astype(np.uint64)
        if n == 2:
            cc[rng.integers(0, 2)] = 0  # unvisited state branch
        sb = rng.normal(size=n)
        mine = eng.compute_entropy(cc, sb, 0.5, 0.1)
        ref = sb.copy()
        oracle.orc_compute_entropy(cc.ctypes.data, ref.ctypes.data, n, 0.5, 0.1)
        assert np.array_equal(mine.view(np.uint64), ref.view(np.uint64))
        assert eng.compute_g_val(cc) == oracle.orc_compute_g_val(cc.ctypes.data, n)
        ok, g, crit, pv = eng.g_test(cc, 0.05)
        g2, c2, p2 = C.c_double(), C.c_double(), C.c_double()
        ok2 = oracle.orc_g_test(cc.ctypes.data, n, 0.05, C.addressof(g2), C.addressof(c2), C.addressof(p2))
        assert (ok, g, crit, pv) == (bool(ok2), g2.value, c2.value, p2.value)


def test_init_pos_chain_matches_oracle(oracle):
    _lib()
    from common import OracleRun
    for name in ("8bead_1bond", "crambin"):
        p = params_from_json(load_config(name))
        pos = eng.init_pos_chain(p)
        o = OracleRun(p, oracle)
        assert np.array_equal(pos.view(np.uint64), o.init_pos().view(np.uint64))


def test_product_does_not_reference_oracle():
    """the shipped package / C sources never import, link or dlopen oracle/"""
    for base, _, files in os.walk(os.path.join(ROOT, "hybridmc_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cpp")):
                txt = open(os.path.join(base, f)).read()
                assert "oracle_lib" not in txt and "libhmc_oracle" not in txt and "libhmc_ref" not in txt, f


def test_stair_stages_follow_init_json():
    """tools/init_json.py:93-120: stage j runs with rc = rc[j], the previous radius as
    `stair` wall on the transient bond, p_rc only above layer 0, rc in every name but the last"""
    from hybridmc_b200.run_campaign import stair_stages
    d = dict(rc=1.5, transient_bonds=[[4, 16]], permanent_bonds=[])
    st = stair_stages(d, 0, 1.5, [3.0, 2.2, 1.5])
    assert [s for _, s in st] == ["_3.0", "_2.2", ""]
    assert [x["rc"] for x, _ in st] == [3.0, 2.2, 1.5]
    assert "stair" not in st[0][0] and st[1][0]["stair"] == 3.0 and st[2][0]["stair"] == 2.2
    assert st[1][0]["stair_bonds"] == [[4, 16]] and all("p_rc" not in x for x, _ in st)
    st1 = stair_stages(dict(d, permanent_bonds=[[2, 6]]), 1, 1.5, [3.0, 1.5])
    assert all(x["p_rc"] == 1.5 for x, _ in st1) and d["rc"] == 1.5 and "stair" not in d


def test_stair_entropy_combination():
    """tools/diff_s_bias_stair.py:47-58: all non-empty subsets of the stage values"""
    from itertools import combinations
    from hybridmc_b200.campaign import collect_stair, stair_entropy
    v = [2.1, 0.7, 1.3]
    ref = np.log(sum(np.exp(sum(c)) for i in range(1, 4) for c in combinations(v, i)))
    assert abs(stair_entropy(v) - ref) < 1e-12
    rows = [("000", "100_3.0", 2.1), ("000", "100_2.2", 0.7), ("000", "100", 1.3), ("000", "010", 3.5)]
    out = collect_stair(rows)
    assert abs(out[("000", "100")] - ref) < 1e-12 and out[("000", "010")] == 3.5


def test_campaign_host_logic():
    """layer_replicas (replica split + topping thin layers up) and the warm-start guess"""
    from hybridmc_b200.run_campaign import layer_replicas, warm_start_guess
    assert layer_replicas(64, 1260, world=8) == 8
    assert layer_replicas(64, 10, world=1, fill=4736) == 474      # thin layer topped up
    assert layer_replicas(64, 10, world=2, fill=4736) == 474      # per rank
    assert layer_replicas(64, 1260, world=2, fill=4736) == 32     # fat layer: the nominal split
    assert layer_replicas(65, 4, world=8) == 9                    # rounded up
    T, F = True, False
    est = {((F, F, F), (T, F, F)): 9.4, ((F, T, F), (T, T, F)): 2.0, ((F, F, T), (T, F, T)): 8.8}
    assert warm_start_guess(est, (F, F, F), (T, F, F)) is None                 # layer 0
    assert warm_start_guess(est, (F, T, F), (T, T, F)) == 9.4                  # one predecessor
    # bond 0 on top of {1, 2}: predecessors remove bond 1 (-> 8.8) or bond 2 (-> 2.0): the smaller
    assert warm_start_guess(est, (F, T, T), (T, T, T)) == 2.0


def test_state_entropies_match_avg_s_bias_semantics():
    """tools/avg_s_bias.py on a 2-bond lattice worked by hand: S(11) = 0; S(01) = d(01->11);
    S(10) = d(10->11); S(00) = mean(S(01) + d(00->01), S(10) + d(00->10))"""
    from hybridmc_b200.campaign import state_entropies
    d = {("01", "11"): 2.0, ("10", "11"): 3.0, ("00", "01"): 4.0, ("00", "10"): 3.5}
    S = state_entropies(2, d)
    assert S == {"11": 0.0, "01": 2.0, "10": 3.0, "00": (6.0 + 6.5) / 2}


def test_comm_argument_validation():
    """the communicator entry points refuse bad arguments before touching NCCL or CUDA"""
    lib = _lib()
    c = C.c_void_p()
    assert lib.hmc_comm_wrap_nccl(None, 0, 1, 0, C.byref(c)) == -1
    assert lib.hmc_comm_create(None, 0, 1, 0, C.byref(c)) == -1
    buf = (C.c_uint8 * 128)()
    assert lib.hmc_comm_create(buf, 3, 2, 0, C.byref(c)) == -1  # rank >= world
    assert lib.hmc_comm_world(None) == 1 and lib.hmc_comm_rank(None) == 0
    assert lib.hmc_allreduce_counts(None, None) != 0            # null handle


def test_ensemble_program_fails_loudly_without_gpu():
    """the production entry point through its ctypes mirror (structs, callbacks): without a
    CUDA device it reports the library's message instead of falling back or crashing"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    _lib()
    p = params_from_json(load_config("8bead_1bond"))
    with pytest.raises(eng.EnsembleError) as ei:
        eng.run_program_ensemble([p], [eng.init_pos_chain(p)], 4, hist_bins=8, hist_rmax=4.0, dist_rows=8)
    assert "no CPU fallback" in str(ei.value) and ei.value.code < 0
    with pytest.raises(eng.EngineError):
        eng.resident_replicas(20)
