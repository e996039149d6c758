"""Spread of the biased-entropy estimate of one transition — the counterpart of the
reference's `tools/get_error_s_bias.py` (which starts the executable `nboot` times with seeds
[i, seed_increment] in a multiprocessing pool and reports mean, variance and "percent error"
of the `nboot` values of s_bias[0]): here the `nboot` independent programs are the replicas of
ONE ensemble launch in independent-program mode (every replica runs the reference's whole
program — equilibration, Wang-Landau, G-test rounds of total_iter iterations — on its own
counts and its own Philox stream).

  python -m hybridmc_b200.get_error_s_bias json [--hdf5 previous_layer.h5] nboot [--seed_increment S]

writes `<json name>_s_bias_error.csv` with the reference's row [mean, var, percent_rel_e]
(get_error_s_bias.py:52-66) and prints the three numbers."""
import argparse
import csv
import json
import os
import sys

import numpy as np

if __package__ in (None, ""):
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    __package__ = "hybridmc_b200"

from .engine import init_pos_chain, run_program_ensemble
from .output import File
from .params import params_from_json


def s_bias_samples(data, nboot, start=None, seed_increment=2, device=0, max_rounds=0):
    """s_bias[0] of `nboot` independent programs of the transition described by `data`."""
    p = params_from_json(data)
    if start is None:
        start = init_pos_chain(p)
    out = run_program_ensemble([p], [start], nboot, device=device, seed=(p.seeds[0], seed_increment),
                               independent=True, iters_per_round=p.total_iter,
                               eq_iters=p.total_iter_eq, max_rounds=max_rounds)
    return out["replicas"][0][0][:, 0], out


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("json", help="json input file")
    ap.add_argument("exe", nargs="?", default=None, help="(compatibility) hybridmc executable; unused")
    ap.add_argument("--hdf5", help="hdf5 input file (unique_config/pos of the previous layer)")
    ap.add_argument("nboot", type=int, help="number of independent programs")
    ap.add_argument("--seed_increment", type=int, default=2)
    ap.add_argument("--device", type=int, default=0)
    args = ap.parse_args(argv)
    with open(args.json) as f:
        data = json.load(f)
    start = None
    if args.hdf5:
        with File(args.hdf5) as f:
            start = np.array(f["unique_config"]["pos"][0])
    s, _ = s_bias_samples(data, args.nboot, start, args.seed_increment, args.device)
    mean, var = float(np.mean(s)), float(np.var(s))
    percent = var / mean * 100  # get_error_s_bias.py:58-59
    print("mean of entropy samples:", mean)
    print("var of entropy samples:", var)
    print("percent error:", percent)
    name = os.path.splitext(os.path.basename(args.json))[0]
    with open("%s_s_bias_error.csv" % name, "w", newline="") as f:
        csv.writer(f).writerows([[mean, var, percent]])
    return mean, var, percent


if __name__ == "__main__":
    main()
