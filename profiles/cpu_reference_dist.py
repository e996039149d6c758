"""Distance samples (`dist` dataset) of one per-transition run on the CPU with the C oracle
(bit-identical to the reference's sources), at the reference's own settings, for checking
the MFPT estimator at the reference's sample size.
usage: cpu_reference_dist.py config transition_index total_iter seed max_rounds out.npz"""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

if __name__ == "__main__":
    from common import load_config
    from hybridmc_b200.params import params_from_json, transition_jsons
    from oracle_lib import OrcMainResult, load_oracle
    config, index, total_iter, seed, max_rounds, out = sys.argv[1], *map(int, sys.argv[2:6]), sys.argv[6]
    d = transition_jsons(load_config(config))[index][3]
    d.update(total_iter=total_iter, total_iter_eq=max(1, total_iter // 4), seeds=[seed, 1])
    p = params_from_json(d)
    cap = max_rounds * (total_iter * p.nsteps // p.write_step + 8)
    dist = np.zeros((cap, p.n_nonlocal))
    n = C.c_uint(0)
    r = OrcMainResult()
    rc = load_oracle().orc_run_main(C.byref(p), None, max_rounds, cap, dist.ctypes.data_as(C.c_void_p),
                                    C.byref(n), C.byref(r))
    np.savez_compressed(out, dist=dist[:n.value], s_bias=np.array(list(r.s_bias)[:2]),
                        transient=np.array(d["transient_bonds"]), nonlocal_bonds=np.array(d["nonlocal_bonds"]))
    print(json.dumps(dict(seed=seed, rc=rc, rows=int(n.value), s_bias0=r.s_bias[0], rounds=r.gtest_iters,
                          seconds=r.seconds_md)), flush=True)
