"""Compare a campaign's diff_s_bias.csv (bits_in, bits_out, s_bias[0]) with the reference's
published per-state entropies: s_bias[0] of a transition estimates S(bits_in) - S(bits_out)
(tools/diff_s_bias.py:53-54, tools/avg_s_bias.py).  usage: compare_campaign_to_published.py CSV"""
import csv
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pub = json.load(open(os.path.join(ROOT, "tests", "golden", "crambin_published_s_bias_subset.json")))
S = pub["upto_three_bonds"]
rows = [r for r in csv.reader(open(sys.argv[1])) if r]
by_layer = {}
for bits_in, bits_out, s in rows:
    if bits_in in S and bits_out in S:
        by_layer.setdefault(bits_in.count("1"), []).append((float(s), S[bits_in] - S[bits_out]))
print("%5s %6s %10s %10s %10s %10s" % ("layer", "n", "mean diff", "sd diff", "max |diff|", "corr"))
for layer in sorted(by_layer):
    a = np.array(by_layer[layer])
    d = a[:, 0] - a[:, 1]
    print("%5d %6d %10.4f %10.4f %10.4f %10.4f" % (layer, len(a), d.mean(), d.std(ddof=1),
                                                   np.abs(d).max(), np.corrcoef(a[:, 0], a[:, 1])[0, 1]))
