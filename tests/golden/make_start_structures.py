"""Generates tests/golden/start_structures_<config>.npz: one valid start structure per bond
state of the master config (the empty state from init_pos, every other state the first
step-end configuration with the transient bond formed, as the reference hands
unique_config/pos from layer to layer, src/main.cc:211-216).  Generated on the CPU with
the oracle (bit-identical to the compiled reference); used by bench.py so that both arms
(GPU and reference CPU) start from the same structures.

  python tests/golden/make_start_structures.py test
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from common import OracleRun, load_config  # noqa: E402
from hybridmc_b200.params import params_from_json, transition_jsons  # noqa: E402

PHASE_MAIN = 2


def key(bits):
    return "".join("1" if b else "0" for b in bits)


def main(name):
    master = load_config(name)
    structures = {}
    for layer, bits_in, bits_out, d in transition_jsons(master):
        p = params_from_json(d)
        run = OracleRun(p)
        if not structures:
            structures[key(bits_in)] = run.init_pos().copy()
        if key(bits_out) in structures:
            continue
        run.sim.set_pos(structures[key(bits_in)])
        run.sim.init_config_from_pos()
        run.sim.init_s()
        run.sim.reset_clocks()
        for step in range(100000):
            assert run.sim.md_step(PHASE_MAIN, step, run.normals(1)[0]) == 0
            if run.sim.get_config() == 1:
                structures[key(bits_out)] = run.sim.get_pos().copy()
                print("layer", layer, key(bits_in), "->", key(bits_out), "after", step + 1, "steps")
                break
        else:
            raise SystemExit("no bonded structure for " + key(bits_out))
    np.savez(os.path.join(HERE, "start_structures_%s.npz" % name), **structures)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "test")
