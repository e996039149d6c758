"""Worker of tests/test_gpu_debug_builds.py: runs ensembles of the shipped and of random
configurations under the library named by HMC_LIB (a checking build) and prints one JSON line
{events, err_or, bounds_violations}."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
from common import load_config  # noqa: E402
from hybridmc_b200.engine import PHASE_EQ, PHASE_MAIN, PHASE_WL, Engine, init_pos_chain, load_library  # noqa: E402
from hybridmc_b200.params import params_from_json, transition_jsons  # noqa: E402
from test_gpu_random_parity import random_config, stair_start  # noqa: E402

events, err_or = 0, 0


def run(params, pos, R, iters, wl=False, label="", allow=0):
    global events, err_or
    with Engine(params, R, 0) as e:
        e.set_positions(pos)
        e.init_config_from_positions()
        e.init_s()
        e.seed(11, 13)
        e.reset_clocks()
        e.run(PHASE_EQ, 0, iters)
        e.init_config_from_positions()
        e.reset_clocks()
        if wl:
            e.run(PHASE_WL, 0, 20)
            e.reset_clocks()
        e.run(PHASE_MAIN, 0, iters)
        err = e.get_counts()[3]
        err_or |= int(np.bitwise_or.reduce(err)) & ~allow
        events += int(e.get_event_counts()[0])
        if np.any(err & ~np.uint32(32 | allow)):
            sys.stderr.write("%s: err flags %s in replicas %s\n" % (
                label, sorted(set(int(x) for x in err if x)), np.nonzero(err)[0][:8].tolist()))


# every transition of test.json from the committed start structures (all lattice layers)
master = load_config("test")
st = np.load(os.path.join(HERE, "golden", "start_structures_test.npz"))
trans = transition_jsons(master)
params = [params_from_json(t[3]) for t in trans]
R = 96
pos = np.stack([np.broadcast_to(st["".join("1" if b else "0" for b in t[1])], (R, master["nbeads"], 3))
                for t in trans]).reshape(len(trans) * R, master["nbeads"], 3)
run(params, pos, R, 2, wl=True, label="test.json")
# long chains (16-lane groups, general minimum image), layer 0
for name, R in (("crambin", 24), ("50bead_9bond", 24), ("8bead_1bond", 64)):
    m = load_config(name)
    tj = [t for t in transition_jsons(m) if t[0] == 0] if len(m["nonlocal_bonds"]) > 1 else None
    ps = [params_from_json(t[3]) for t in tj] if tj else [params_from_json(m)]
    p0 = init_pos_chain(ps[0])
    run(ps, np.broadcast_to(p0, (len(ps) * R,) + p0.shape).copy(), R, 1, label=name)
# random configurations (staircase walls, several transient bonds, odd cell counts)
for seed in range(10):
    d = random_config(seed)
    p = params_from_json(d)
    p0 = init_pos_chain(p)
    if "stair" in d:
        p0 = stair_start(d, p0)
    # (draw 9 — three transient bonds of rc 2.2 on 14 beads with max_nbonds 2 — reaches states in
    # which no crankshaft rotation passes the reference's checks: the reference's do-while,
    # crankshaft.cc:238-247, does not terminate there and the oracle stops with the same
    # condition in 14 of 400 seeds; the engine raises HMC_ERR_CRANK_STUCK for the replica)
    run([p], np.broadcast_to(p0, (48,) + p0.shape).copy(), 48, 2, label="random %d" % seed, allow=128)
print(json.dumps({"events": events, "err_or": err_or,
                  "bounds_violations": int(load_library().hmc_debug_bounds_violations())}))
