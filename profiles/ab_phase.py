"""A/B helper: the bench workload (test.json, all 12 transitions) timed phase by phase:
EQ with nsteps_eq=8 (MD only), MAIN without crankshaft (mc_moves=0), MAIN as configured.
Usage: [HMC_LIB=...] python profiles/ab_phase.py [R] [config]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import load_master  # noqa: E402
from hybridmc_b200.campaign import harvest_start_structures, layers_of  # noqa: E402
from hybridmc_b200.engine import PHASE_EQ, PHASE_MAIN, Engine  # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
config = sys.argv[2] if len(sys.argv) > 2 else "test"
base = load_master(config)
structures = harvest_start_structures(base, replicas=128, seed=99)
for label, over, phase in (("EQ 8 steps", dict(nsteps_eq=8), PHASE_EQ),
                           ("MAIN no crank", dict(mc_moves=0), PHASE_MAIN),
                           ("MAIN", {}, PHASE_MAIN)):
    master = dict(base, **over)
    trans = [t for _, ts in layers_of(master) for t in ts]
    e = Engine([t[2] for t in trans], R, 0)
    pos = np.stack([np.broadcast_to(structures[t[0]], (R,) + structures[t[0]].shape)
                    for t in trans]).reshape(e.R, e.nb, 3)
    e.set_positions(pos)
    e.init_config_from_positions()
    e.init_s()
    e.seed(5, 6)
    e.reset_clocks()
    e.run(PHASE_EQ, 0, 1)
    e.init_config_from_positions()
    e.reset_clocks()
    e.sync()
    tot_ms, tot_ev = 0.0, 0
    for it in range(3):
        ev0 = e.get_event_counts()
        e.run(phase, it, 1)
        e.sync()
        ev1 = e.get_event_counts()
        if it:
            tot_ms += e.launch_stats()[1]
            tot_ev += int(ev1[0] - ev0[0])
    print("%-14s replicas %d  %.1f ms/launch  %.4g events/s  err %d" % (
        label, e.R, tot_ms / 2, tot_ev / (tot_ms * 1e-3), int(np.bitwise_or.reduce(e.get_counts()[3]))))
    e.close()
