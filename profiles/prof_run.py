"""Profiling driver: one launch of the ensemble kernel on the bench workload
(test.json, 12 transitions x R replicas, 1 run_trajectory iteration) between
cudaProfilerStart/Stop, so `ncu --profile-from-start off` captures exactly it.
Usage: python profiles/prof_run.py [config] [replicas_per_transition] [iters]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import load_master  # noqa: E402
from hybridmc_b200.campaign import harvest_start_structures, layers_of  # noqa: E402
from hybridmc_b200.engine import PHASE_EQ, PHASE_MAIN, Engine  # noqa: E402

config = sys.argv[1] if len(sys.argv) > 1 else "test"
R = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 1
max_layers = int(sys.argv[4]) if len(sys.argv) > 4 else 99
master = load_master(config)
layers = [l for l in layers_of(master) if l[0] < max_layers]
if max_layers >= 99:
    structures = harvest_start_structures(master, replicas=128, seed=99)
else:
    from hybridmc_b200.engine import init_pos_chain
    nb_bonds = len(master["nonlocal_bonds"])
    structures = {tuple([False] * nb_bonds): init_pos_chain(layers[0][1][0][2])}
trans = [t for _, ts in layers for t in ts]
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
torch.cuda.synchronize()
ev0 = e.get_event_counts()
torch.cuda.profiler.start()
e.run(PHASE_MAIN, 0, iters)
e.sync()
torch.cuda.profiler.stop()
ev1 = e.get_event_counts()
n, ms = e.launch_stats()
print("config %s replicas %d events %d ms %.3f -> %.4g events/s  err %d" % (
    config, e.R, int(ev1[0] - ev0[0]), ms, float(ev1[0] - ev0[0]) / (ms * 1e-3),
    int(np.bitwise_or.reduce(e.get_counts()[3]))))
