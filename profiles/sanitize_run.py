"""Tiny workload for compute-sanitizer: injected MD steps + crankshaft + a Philox MAIN
iteration on 8bead_1bond and a 20-bead transition (few replicas)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import load_config, transition_params  # noqa: E402
from hybridmc_b200.engine import PHASE_EQ, PHASE_MAIN, PHASE_WL, Engine, init_pos_chain  # noqa: E402
from hybridmc_b200.params import params_from_json  # noqa: E402

for p, R in ((params_from_json(load_config("8bead_1bond")), 7), (transition_params("test", 0), 5),
             (transition_params("crambin", 0), 2)):
    pos = init_pos_chain(p)
    with Engine(p, R) as e:
        e.set_positions(np.broadcast_to(pos, (R,) + pos.shape))
        e.init_config_from_positions()
        e.init_s()
        e.seed(1, 2)
        e.enable_samples(64, 64)
        e.enable_dist_hist(256, p.length)
        e.reset_clocks()
        e.run(PHASE_EQ, 0, 1)
        e.reset_clocks()
        e.run(PHASE_WL, 0, 3)
        e.reset_clocks()
        e.run(PHASE_MAIN, 0, 1)
        e.sync()
        print(p.nbeads, e.get_event_counts(), e.get_counts()[3])
