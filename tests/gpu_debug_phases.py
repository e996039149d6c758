import sys, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
from common import load_config
from hybridmc_b200.engine import *
from hybridmc_b200.params import params_from_json
what = sys.argv[1]; R = int(sys.argv[2])
p = params_from_json(load_config("8bead_1bond", mc_moves=int(sys.argv[3]) if len(sys.argv) > 3 else 20))
e = Engine(p, R)
pos = init_pos_chain(p)
e.set_positions(np.broadcast_to(pos, (R,) + pos.shape)); e.init_config_from_positions(); e.init_s(); e.seed(3, 4); e.reset_clocks()
t0 = time.time()
if what == "eq":
    e.run(PHASE_EQ, 0, 2)
elif what == "main":
    e.run(PHASE_MAIN, 0, 2)
elif what == "wl":
    e.run(PHASE_WL, 0, 5)
e.sync()
print(what, R, "ok", time.time() - t0, e.get_event_counts(), e.get_counts()[3][:4])
