"""CPU reference estimate of s_bias[0] for one per-transition run, several seeds in
parallel (one process per seed), with the C oracle (bit-identical to the reference's
sources).  usage: cpu_reference_s_bias.py config transition_index total_iter nseeds max_rounds"""
import ctypes as C
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def one(config, index, total_iter, seed, max_rounds, sig_level=None):
    from common import load_config
    from hybridmc_b200.params import params_from_json, transition_jsons
    from oracle_lib import OrcMainResult, load_oracle
    d = transition_jsons(load_config(config))[index][3]
    d.update(total_iter=total_iter, total_iter_eq=max(1, total_iter // 4), seeds=[seed, 1])
    if sig_level is not None:
        d["sig_level"] = sig_level  # ~1: the G-test never accepts -> exactly max_rounds rounds
    p = params_from_json(d)
    r = OrcMainResult()
    rc = load_oracle().orc_run_main(C.byref(p), None, max_rounds, 0, None, None, C.byref(r))
    print(json.dumps(dict(seed=seed, rc=rc, s_bias0=r.s_bias[0], rounds=r.gtest_iters,
                          accepted=r.accepted, counts=list(r.config_count)[:2],
                          collisions=r.collisions, seconds=r.seconds_md)), flush=True)


if __name__ == "__main__":
    if sys.argv[1] == "--one":
        one(sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6]),
            float(sys.argv[7]) if len(sys.argv) > 7 else None)
    else:
        config, index, total_iter, nseeds, max_rounds = sys.argv[1], *map(int, sys.argv[2:6])
        extra = sys.argv[6:7]
        procs = [subprocess.Popen([sys.executable, __file__, "--one", config, str(index),
                                   str(total_iter), str(500 + s), str(max_rounds)] + extra,
                                  stdout=subprocess.PIPE, text=True) for s in range(nseeds)]
        for p in procs:
            print(p.communicate()[0].strip(), flush=True)
