"""Small profiling target: ONE hybrid-MC interval (initialize_system + run_step, EQ phase
with nsteps_eq=1) of R replicas per layer-0 transition.  Short enough for
`ncu --set full` (tens of replay passes).  Usage: prof_step.py [config] [R] [launches]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import load_master  # noqa: E402
from hybridmc_b200.campaign import layers_of  # noqa: E402
from hybridmc_b200.engine import PHASE_EQ, Engine, init_pos_chain  # noqa: E402

config = sys.argv[1] if len(sys.argv) > 1 else "test"
R = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
launches = int(sys.argv[3]) if len(sys.argv) > 3 else 1
master = dict(load_master(config), nsteps_eq=1)
trans = layers_of(master)[0][1][:3]
pos0 = init_pos_chain(trans[0][2])
e = Engine([t[2] for t in trans], R, 0)
e.set_positions(np.broadcast_to(pos0, (e.R,) + pos0.shape))
e.init_config_from_positions()
e.init_s()
e.seed(5, 6)
e.reset_clocks()
e.run(PHASE_EQ, 0, 3)  # warm-up launch: steps 0..2 decorrelate the replicas
e.sync()
torch.cuda.synchronize()
ev0 = e.get_event_counts()
torch.cuda.profiler.start()
tot = 0.0
for k in range(launches):
    e.run(PHASE_EQ, 3 + k, 1)
    e.sync()
    tot += e.launch_stats()[1]
torch.cuda.profiler.stop()
ev1 = e.get_event_counts()
print("config %s replicas %d launches %d events %d ms %.3f -> %.4g events/s err %d" % (
    config, e.R, launches, int(ev1[0] - ev0[0]), tot, float(ev1[0] - ev0[0]) / (tot * 1e-3),
    int(np.bitwise_or.reduce(e.get_counts()[3]))))
