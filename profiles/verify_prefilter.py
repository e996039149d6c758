"""Run ensembles under the verification build of the partner pre-filter
(`python -m hybridmc_b200.build --out=hybridmc_b200/libhmc_verify.so -DHMC_VERIFY_PREFILTER`,
then HMC_LIB=.../libhmc_verify.so): every pair the single-precision filter would drop is
predicted exactly anyway and error bit 0x40000000 is raised if one of them has an event.
Covers the shipped configurations (incl. deeper layers with permanent bonds and
crambin's staircase variant) and the random configurations of tests/test_gpu_random_parity.py."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import load_config  # noqa: E402
from hybridmc_b200.campaign import harvest_start_structures, layers_of  # noqa: E402
from hybridmc_b200.engine import PHASE_EQ, PHASE_MAIN, Engine, init_pos_chain  # noqa: E402
from hybridmc_b200.params import params_from_json  # noqa: E402
from test_gpu_random_parity import random_config  # noqa: E402

FLAG = 0x40000000
assert "verify" in os.environ.get("HMC_LIB", ""), "set HMC_LIB to the verification build"
total = 0
bad = 0


def run(params, pos, R, iters, label):
    global total, bad
    with Engine(params, R, 0) as e:
        e.set_positions(pos)
        e.init_config_from_positions()
        e.init_s()
        e.seed(11, 13)
        e.reset_clocks()
        e.run(PHASE_EQ, 0, iters)
        e.init_config_from_positions()
        e.reset_clocks()
        e.run(PHASE_MAIN, 0, iters)
        err = int(np.bitwise_or.reduce(e.get_counts()[3]))
        ev = int(e.get_event_counts()[0])
    total += ev
    bad += 1 if err & FLAG else 0
    print("%-40s events %12d err 0x%x %s" % (label, ev, err, "PREFILTER VIOLATION" if err & FLAG else "ok"),
          flush=True)


for name, R, iters, max_layers in (("test", 512, 2, 3), ("crambin", 32, 2, 1), ("50bead_9bond", 32, 2, 1)):
    master = load_config(name)
    layers = layers_of(master)[:max_layers]
    sub = dict(master)
    structures = None
    try:
        structures = harvest_start_structures(master if max_layers >= len(layers_of(master)) else master,
                                              replicas=32, max_rounds=40) if name == "test" else None
    except RuntimeError:
        structures = None
    for layer, trans in layers:
        trans = trans[:24]
        params = [t[2] for t in trans]
        if structures is not None:
            pos = np.stack([np.broadcast_to(structures[t[0]], (R,) + structures[t[0]].shape) for t in trans])
        elif layer == 0:
            p0 = init_pos_chain(params[0])
            pos = np.broadcast_to(p0, (len(trans) * R,) + p0.shape)
        else:
            continue  # deeper layers need harvested structures
        run(params, np.ascontiguousarray(pos).reshape(len(trans) * R, params[0].nbeads, 3), R, iters,
            "%s layer %d (%d transitions)" % (name, layer, len(trans)))

# (random configurations 0, 3, 5 and 12 start outside their potential's domain — the
# reference refuses them — and are left out)
for seed in (1, 2, 4, 6, 7, 8, 9, 10, 11, 13):
    d = random_config(seed)
    p = params_from_json(d)
    try:
        p0 = init_pos_chain(p)
    except Exception as exc:  # noqa: BLE001
        print("random %d: no start structure (%s)" % (seed, exc))
        continue
    run([p], np.ascontiguousarray(np.broadcast_to(p0, (512,) + p0.shape)), 512, 3,
        "random config %d (nb %d ncell %d stair %s)" % (seed, p.nbeads, p.ncell, bool(p.has_stair)))
print("total events %d, configurations with a dropped real event: %d" % (total, bad))
sys.exit(1 if bad else 0)
