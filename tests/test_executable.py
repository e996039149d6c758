"""The drop-in executable (hybridmc_b200/hybridmc): CLI contract of src/main.cc:307-356 on
the CPU, and on the GPU its outputs against the oracle (bit-identical program)."""
import ctypes as C
import json
import os
import subprocess

import numpy as np
import pytest

from common import load_config
from hybridmc_b200.params import params_from_json, transition_jsons

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def exe():
    from hybridmc_b200.build import build_executable, build_library
    build_library()
    return build_executable()


def run(args, cwd=None):
    return subprocess.run([exe()] + args, capture_output=True, text=True, cwd=cwd)


def test_cli_contract():
    r = run(["--help"])
    assert r.returncode == 0 and "usage: hybridmc [options]... json-file output-file" in r.stdout
    r = run([])
    assert r.returncode == 1 and "try hybridmc --help" in r.stderr  # main.cc:337-341
    r = run(["a.json"])
    assert r.returncode == 1
    r = run(["a.json", "b.h5", "c.h5"])
    assert r.returncode == 1
    r = run(["a.json", "b.h5", "--bogus"])
    assert r.returncode == 1


def test_bad_config_fails_nonzero(tmp_path):
    bad = dict(load_config("8bead_1bond"), stair=1.2)
    p = tmp_path / "bad.json"
    p.write_text(json.dumps(bad))
    r = run([str(p), str(tmp_path / "o.h5")])
    assert r.returncode != 0 and "stair boundary is smaller than rc" in r.stderr
    r = run([str(tmp_path / "missing.json"), str(tmp_path / "o.h5")])
    assert r.returncode != 0
    assert not (tmp_path / "o.h5").exists()


def test_fails_loudly_without_gpu(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    p = tmp_path / "c.json"
    p.write_text(json.dumps(load_config("8bead_1bond")))
    r = run([str(p), str(tmp_path / "o.h5")])
    assert r.returncode != 0 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_outputs_match_oracle_and_layer_handoff(tmp_path, oracle):
    from hybridmc_b200.output import File
    from oracle_lib import OrcMainResult
    trans = transition_jsons(load_config("test"))
    # layer 0: bond [2,6] transient
    d0 = dict(trans[0][3], total_iter=40, total_iter_eq=8, max_g_test_count=3)
    j0 = tmp_path / "hybridmc_0_000_100.json"
    j0.write_text(json.dumps(d0))
    out0 = tmp_path / "hybridmc_0_000_100.h5"
    r = run([str(j0), str(out0)])
    assert r.returncode == 0, r.stderr
    assert "ACCEPTED" in r.stdout or "REJECTED" in r.stdout
    assert not os.path.exists(str(out0) + ".tmp")
    p0 = params_from_json(d0)
    ro = OrcMainResult()
    dist_o = np.zeros((40 * 8 * 3, p0.n_nonlocal))
    n_o = C.c_uint(0)
    assert oracle.orc_run_main(C.byref(p0), None, 0, len(dist_o), dist_o.ctypes.data_as(C.c_void_p),
                               C.byref(n_o), C.byref(ro)) == 0
    from hybridmc_b200 import h5lite
    assert h5lite.is_hdf5(str(out0))
    with File(str(out0)) as f:
        assert np.array_equal(f["s_bias"], np.array(list(ro.s_bias)[:2], dtype=np.float32))
        cc = f["config_count"]
        assert cc.shape == (1 + ro.gtest_iters, 2) and not cc[0].any()
        assert list(cc[-1]) == list(ro.config_count)[:2]
        assert f["dist"].shape == (ro.gtest_iters * 40 * 8, p0.n_nonlocal)
        assert np.array_equal(f["dist"][-n_o.value:].view(np.uint64), dist_o[:n_o.value].view(np.uint64))
        assert f.attrs["nsteps"] == 8 and f.attrs["length"] == 18.0
        assert f.attrs["transient_bonds"].tolist() == [[2, 6]]
        assert f["config"]["int"].dtype == np.uint64 and len(f["config"]["int"]) > 0
        upos = f["unique_config"]["pos"]
    if len(upos) == 0:
        pytest.skip("bond did not form in the shortened run")
    # layer 1: [2,6] permanent, next bond transient, started from the harvested structure
    idx = next(k for k, t in enumerate(trans) if t[0] == 1 and t[3]["permanent_bonds"] == [[2, 6]])
    d1 = dict(trans[idx][3], total_iter=30, total_iter_eq=6, max_g_test_count=2)
    j1 = tmp_path / "l1.json"
    j1.write_text(json.dumps(d1))
    out1 = tmp_path / "l1.h5"
    r = run([str(j1), str(out1), "--input-file", str(out0)])
    assert r.returncode == 0, r.stderr
    p1 = params_from_json(d1)
    r1 = OrcMainResult()
    pin = np.ascontiguousarray(upos[0])
    assert oracle.orc_run_main(C.byref(p1), pin.ctypes.data_as(C.c_void_p), 0, 0, None, None,
                               C.byref(r1)) == 0
    with File(str(out1)) as f:
        assert np.array_equal(f["s_bias"], np.array(list(r1.s_bias)[:2], dtype=np.float32))


@pytest.mark.gpu
def test_snapshot_resume(tmp_path):
    from hybridmc_b200.output import File
    d = dict(load_config("8bead_1bond"), total_iter=20, total_iter_eq=4, max_g_test_count=2)
    j = tmp_path / "c.json"
    j.write_text(json.dumps(d))
    snap = tmp_path / "snap.h5"
    r = run([str(j), str(tmp_path / "a.h5"), "--snapshot-file", str(snap)])
    assert r.returncode == 0, r.stderr
    with File(str(snap)) as s:
        assert s["pos"].shape == (8, 3) and s["s_bias"].shape == (2, 1)
        assert len(s["mt"].split()) == 625  # 624 words + index, as operator<< writes
        assert int(s.attrs["config"]) in (0, 1)
    r = run([str(j), str(tmp_path / "b.h5"), "--snapshot-file", str(snap)])
    assert r.returncode == 0, r.stderr
    with File(str(tmp_path / "b.h5")) as f:
        assert np.isfinite(f["s_bias"]).all()


@pytest.mark.gpu
def test_hmcb_and_hdf5_outputs_hold_the_same_data(tmp_path):
    from hybridmc_b200 import h5lite
    from hybridmc_b200.output import File
    d = dict(load_config("8bead_1bond"), total_iter=12, total_iter_eq=3, max_g_test_count=2)
    j = tmp_path / "c.json"
    j.write_text(json.dumps(d))
    assert run([str(j), str(tmp_path / "a.h5")]).returncode == 0
    assert run([str(j), str(tmp_path / "a.hmcb"), "--output-format", "hmcb"]).returncode == 0
    assert h5lite.is_hdf5(str(tmp_path / "a.h5")) and not h5lite.is_hdf5(str(tmp_path / "a.hmcb"))
    a, b = File(str(tmp_path / "a.h5")), File(str(tmp_path / "a.hmcb"))
    assert sorted(a) == sorted(b) and sorted(a.attrs) == sorted(b.attrs)
    for k in a:
        assert np.asarray(a[k]).dtype == np.asarray(b[k]).dtype and np.array_equal(a[k], b[k]), k
    for k in a.attrs:
        assert np.array_equal(a.attrs[k], b.attrs[k]), k
