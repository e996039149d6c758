"""Generate golden fixtures FROM THE UNMODIFIED REFERENCE (oracle/_ref/libhmc_ref.so,
i.e. /root/reference/src compiled against oracle/shim).  Run in the build container:

    make -C oracle ref && python tests/golden/make_fixtures.py

Outputs (committed): tests/golden/ref_fixtures.npz
  rng_*        std::mt19937(seed_seq) words, init_vel normals, crankshaft tape sin/cos
  initpos_*    init_pos chains of the four shipped configs
  trace_<cfg>  valid-event trace (type,i,j,time bits) of the reference fed injected
               velocities; 8bead_1bond covers the first 10^4 events (north_star)
  crank_*      positions / config / words consumed after 40 reference crankshaft moves
  main_*       whole-program results of reduced-length runs (s_bias bits, counts)
"""
import ctypes as C
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import load_config, transition_params  # noqa: E402
from hybridmc_b200.params import params_from_json  # noqa: E402
from oracle_lib import RefMainResult, Sim, load_ref  # noqa: E402

ref = load_ref()
assert ref is not None, "build oracle/_ref first"
out = {}


def params_of(name):
    if name == "8bead_1bond":
        return params_from_json(load_config(name))
    return transition_params(name, 0)


def seeded(p):
    s = Sim(ref, "ref", p)
    seeds = (C.c_uint32 * p.n_seeds)(*list(p.seeds)[:p.n_seeds])
    ref.ref_mt_seed_seq(s.h, seeds, p.n_seeds)
    return s


# --- RNG -------------------------------------------------------------------
p = params_of("8bead_1bond")
s = seeded(p)
out["rng_words_seed_3_1"] = np.array([ref.ref_mt_next(s.h) for _ in range(64)], dtype=np.uint32)
s = seeded(p)
n = np.zeros(48)
ref.ref_mt_normals(s.h, C.c_double(0.5), n.ctypes.data_as(C.c_void_p), 48)
out["rng_normals_seed_3_1_sigma_0p5"] = n
w, st, ct = np.zeros(64, dtype=np.uint32), np.zeros(64), np.zeros(64)
ref.ref_mt_tape(s.h, w.ctypes.data_as(C.c_void_p), st.ctypes.data_as(C.c_void_p),
                ct.ctypes.data_as(C.c_void_p), 64)
out["rng_tape_words"], out["rng_tape_sin"], out["rng_tape_cos"] = w, st, ct

# --- init_pos + traces -------------------------------------------------------
for name, nsteps, cap in (("8bead_1bond", 24, 12000), ("test", 3, 12000), ("crambin", 2, 20000),
                          ("50bead_9bond", 2, 20000)):
    p = params_of(name)
    s = seeded(p)
    assert ref.ref_init_pos(s.h) == 0
    pos = s.get_pos()
    out["initpos_" + name] = pos
    s.init_s()
    s.reset_clocks()
    s.enable_trace(cap)
    normals = np.zeros((nsteps, p.nbeads, 3))
    for k in range(nsteps):
        ref.ref_mt_normals(s.h, C.c_double(float(np.sqrt(p.temp / p.m))),
                           normals[k].ctypes.data_as(C.c_void_p), 3 * p.nbeads)
        assert s.md_step(0, k, normals[k]) == 0
    tr, ntot = s.get_trace()
    out["trace_%s_normals" % name] = normals
    out["trace_%s_type" % name] = tr["type"].astype(np.uint8)
    out["trace_%s_i" % name] = tr["i"].astype(np.uint8)
    out["trace_%s_j" % name] = tr["j"].astype(np.uint8)
    out["trace_%s_tbits" % name] = tr["t"].view(np.uint64)
    out["trace_%s_total" % name] = np.array([ntot], dtype=np.uint64)
    out["trace_%s_pos_end" % name] = s.get_pos()
    print(name, "events", ntot, "kept", len(tr))
    # --- crankshaft from the end state ---
    used = ref.ref_crankshaft(s.h, C.c_uint(40))
    out["crank_%s_pos" % name] = s.get_pos()
    out["crank_%s_cfg_used" % name] = np.array([s.get_config(), used], dtype=np.uint64)

# --- whole program ---------------------------------------------------------
for name, over in (("8bead_1bond", dict(total_iter=40, total_iter_eq=10)),
                   ("test", dict(total_iter=20, total_iter_eq=5))):
    p = params_from_json(load_config(name, **over))
    r = RefMainResult()
    assert ref.ref_run_main(C.byref(p), None, 3, C.byref(r)) == 0
    ns = p.nstates
    out["main_%s_s_bias" % name] = np.array(list(r.s_bias)[:ns])
    out["main_%s_counts" % name] = np.array(list(r.config_count)[:ns], dtype=np.uint64)
    out["main_%s_meta" % name] = np.array([r.gtest_iters, r.accepted, over["total_iter"],
                                           over["total_iter_eq"], 3], dtype=np.int64)
    print(name, "main", list(r.s_bias)[:ns], r.gtest_iters, r.accepted)

np.savez_compressed(os.path.join(HERE, "ref_fixtures.npz"), **out)
print("wrote", os.path.join(HERE, "ref_fixtures.npz"),
      os.path.getsize(os.path.join(HERE, "ref_fixtures.npz")), "bytes")
