"""Pack the transient bond's state-split distance histograms of every transition of a
campaign directory into one compressed npz (so that a 5 120-transition sweep fits the
64 MiB that travel back from the GPU box); `fit_hist_npz.py` turns it into mfpt.csv.
usage: collect_hist.py DIR OUT.npz"""
import json
import os
import re
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hybridmc_b200.output import File  # noqa: E402
from hybridmc_b200.params import params_from_json  # noqa: E402

NAME = re.compile(r"hybridmc_([0-9]+)_([01]+)_([01]+)\.h5$")
d, out = sys.argv[1], sys.argv[2]
names, hists, rounds, replicas, sbias = [], [], [], [], []
rmax = rh = rc = None
for n in sorted(os.listdir(d)):
    m = NAME.search(n)
    if not m:
        continue
    with open(os.path.join(d, n[:-3] + ".json")) as f:
        p = params_from_json(json.load(f))
    t = p.bond_list("transient")[0]
    col = sorted(p.bond_list("nonlocal")).index(t)
    with File(os.path.join(d, n)) as f:
        hists.append(np.asarray(f["dist_hist"][col], dtype=np.uint32))
        rounds.append(len(f["config_count"]) - 1)
        sbias.append(float(f["s_bias"][0]))
        replicas.append(int(f.attrs["replicas"]))
        rmax = float(f.attrs["dist_hist_rmax"])
    rh, rc = p.rh, p.rc
    names.append(n[:-3])
np.savez_compressed(out, names=np.array(names), hist=np.stack(hists), rounds=np.array(rounds),
                    replicas=np.array(replicas), s_bias0=np.array(sbias), rmax=rmax, rh=rh, rc=rc)
print(len(names), "transitions ->", out, os.path.getsize(out) >> 20, "MiB")
