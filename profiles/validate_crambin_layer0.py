"""Statistical parity against the reference's PUBLISHED crambin results: the ten layer-0
transitions (empty state -> one bond) run as one GPU ensemble; s_bias[0] of each is
compared with S(empty) - S(one-bond state) from crambin_s_bias_mfpt/s_bias.csv (subset
committed as tests/golden/crambin_published_s_bias_subset.json).
usage: python profiles/validate_crambin_layer0.py [replicas] [iters_per_round] [rounds]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import load_master  # noqa: E402
from hybridmc_b200.campaign import layers_of, run_transition_ensemble  # noqa: E402
from hybridmc_b200.engine import init_pos_chain  # noqa: E402
from hybridmc_b200 import mfpt  # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 else 512
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 8
rounds = int(sys.argv[3]) if len(sys.argv) > 3 else 10
nseeds = int(sys.argv[4]) if len(sys.argv) > 4 else 2
select = [int(x) for x in sys.argv[5].split(",")] if len(sys.argv) > 5 and sys.argv[5] != "all" else None
eq_iters = int(sys.argv[6]) if len(sys.argv) > 6 else 8
pub = json.load(open(os.path.join(ROOT, "tests", "golden", "crambin_published_s_bias_subset.json")))
master = load_master("crambin")
trans = layers_of(master)[0][1]
if select is not None:
    trans = [trans[k] for k in select]
params = [t[2] for t in trans]
pos0 = init_pos_chain(params[0])
t0 = time.time()
runs = []
for seed in range(nseeds):
    out = run_transition_ensemble(params, [pos0] * len(trans), R, seed=(900 + seed, 5),
                                  iters_per_round=iters, eq_iters=eq_iters, max_rounds=rounds,
                                  hist_bins=16384, hist_rmax=master["length"] * 0.5 * 3 ** 0.5)
    runs.append(out)
    print("run", seed, "rounds", out["rounds"].tolist(), "accepted", out["accepted"].tolist(),
          "err", int(np.bitwise_or.reduce(out["err"])), "events", out["events"].tolist(),
          "wall %.1f s" % (time.time() - t0), flush=True)
fmt = lambda b: "".join("1" if x else "0" for x in b)
print("%-12s %-10s %10s %10s %10s %8s" % ("bits_out", "bond", "gpu_mean", "gpu_|d|/2", "published", "diff"))
worst = 0.0
for k, (bits_in, bits_out, p) in enumerate(trans):
    vals = [r["s_bias"][k][0] for r in runs]
    mean, half = float(np.mean(vals)), (max(vals) - min(vals)) / 2
    ref = pub["empty"] - pub["one_bond"][fmt(bits_out)]
    worst = max(worst, abs(mean - ref))
    print("%-12s %-10s %10.4f %10.4f %10.4f %8.4f" % (fmt(bits_out), p.bond_list("transient")[0],
                                                       mean, half, ref, mean - ref))
print("largest |gpu - published| = %.4f" % worst)

print()
print("%-12s %-10s %12s %12s %12s %12s" % ("bits_out", "bond", "inner_gpu", "inner_pub", "outer_gpu", "outer_pub"))
rmax = master["length"] * 0.5 * 3 ** 0.5
for k, (bits_in, bits_out, p) in enumerate(trans):
    hist = sum(r["dist_hist"][k].astype(np.float64) for r in runs)
    try:
        inner, outer = mfpt.transition_fpts(hist, rmax, p)
    except Exception as exc:
        inner = outer = float("nan")
        print("fit failed:", exc)
    ref = pub["mfpt_layer0"][fmt(bits_out)]
    print("%-12s %-10s %12.5f %12.5f %12.3f %12.3f" % (fmt(bits_out), p.bond_list("transient")[0],
                                                       inner, ref["inner"], outer, ref["outer"]))
