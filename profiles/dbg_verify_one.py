"""verification build on one random configuration: which pair did the pre-filter drop?"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from hybridmc_b200.engine import PHASE_EQ, PHASE_MAIN, Engine, init_pos_chain
from hybridmc_b200.params import params_from_json
from test_gpu_random_parity import random_config
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 4
d = random_config(seed); p = params_from_json(d)
print({k: d[k] for k in ("nbeads", "ncell", "length", "rc", "m", "temp", "del_t", "nonlocal_bonds", "transient_bonds", "mc_moves", "write_step")})
p0 = init_pos_chain(p)
R = 512
with Engine([p], R, 0) as e:
    e.set_positions(np.ascontiguousarray(np.broadcast_to(p0, (R,) + p0.shape)))
    e.init_config_from_positions(); e.init_s(); e.seed(11, 13); e.reset_clocks()
    e.enable_trace(int(sys.argv[2]) if len(sys.argv) > 2 else 509, 400000)
    for phase, name in ((PHASE_EQ, "EQ"), (PHASE_MAIN, "MAIN")):
        if phase == PHASE_MAIN:
            e.init_config_from_positions(); e.reset_clocks()
        e.run(phase, 0, 3)
        err = e.get_counts()[3]
        for r in np.nonzero(err)[0]:
            v = int(err[r])
            print(name, "replica", r, "err", hex(v), "flag" if v & 0x40000000 else "", "b", (v >> 8) & 63, "k", (v >> 14) & 63, "crossing", (v >> 20) & 1)
    tr, n = e.get_trace()
    print("trace records", n)
    sp = tr[tr["type"] >= 200]
    names = "t tb tk tt pty xb yb zb vxb vyb vzb cellb xk yk zk vxk vyk vzk cellk wall".split()
    for r in sp[:40]:
        print(names[int(r["type"]) - 200], repr(float(r["t"])), "b", int(r["i"]), "k", int(r["j"]))
