"""Compare a crambin campaign (diff_s_bias.csv: bits_in, bits_out, s_bias[0]; mfpt.csv) with
the reference's published results /root/reference/crambin_s_bias_mfpt/{s_bias,mfpt}.csv
(build container only).  s_bias[0] of a transition estimates S(bits_in) - S(bits_out)
(tools/diff_s_bias.py:53-54, tools/avg_s_bias.py).
usage: compare_sweep_to_published.py DIFF_S_BIAS_CSV [MFPT_CSV]"""
import csv
import sys

import numpy as np

PUB = "/root/reference/crambin_s_bias_mfpt/"
S = {r[0]: float(r[1]) for r in list(csv.reader(open(PUB + "s_bias.csv")))[1:]}
rows = [r for r in csv.reader(open(sys.argv[1])) if r]
by_layer = {}
for bits_in, bits_out, s in rows:
    if bits_in in S and bits_out in S:
        by_layer.setdefault(bits_in.count("1"), []).append((float(s), S[bits_in] - S[bits_out]))
print("s_bias[0] vs published S(in) - S(out)")
print("%5s %6s %10s %10s %10s %10s" % ("layer", "n", "mean diff", "sd diff", "max |diff|", "corr"))
alld = []
for layer in sorted(by_layer):
    a = np.array(by_layer[layer])
    d = a[:, 0] - a[:, 1]
    alld.append(d)
    print("%5d %6d %10.4f %10.4f %10.4f %10.4f" % (layer, len(a), d.mean(), d.std(ddof=1) if len(a) > 1 else 0,
                                                   np.abs(d).max(), np.corrcoef(a[:, 0], a[:, 1])[0, 1] if len(a) > 2 else 1))
d = np.concatenate(alld)
print("%5s %6d %10.4f %10.4f %10.4f" % ("all", len(d), d.mean(), d.std(ddof=1), np.abs(d).max()))
if len(rows) == 5120:
    sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
    from hybridmc_b200.campaign import state_entropies
    mine_S = state_entropies(10, {(r[0], r[1]): float(r[2]) for r in rows})
    keys = sorted(S)
    a = np.array([mine_S[k] for k in keys])
    b = np.array([S[k] for k in keys])
    nb_ = np.array([k.count("1") for k in keys])
    print("\nper-state entropies (tools/avg_s_bias.py path average) vs published s_bias.csv, %d states" % len(keys))
    print("%7s %5s %10s %10s %10s" % ("bonds", "n", "mean diff", "sd diff", "max |diff|"))
    for k_ in range(11):
        m = nb_ == k_
        dd = (a - b)[m]
        print("%7d %5d %10.4f %10.4f %10.4f" % (k_, m.sum(), dd.mean(), dd.std(ddof=1) if m.sum() > 1 else 0, np.abs(dd).max()))
    print("S(no bonds): %.3f vs published %.3f; corr %.5f; rms diff %.4f" % (
        mine_S["0" * 10], S["0" * 10], np.corrcoef(a, b)[0, 1], np.sqrt(np.mean((a - b) ** 2))))
if len(sys.argv) > 2:
    pub = {(r[0], r[1]): (float(r[3]), float(r[4])) for r in list(csv.reader(open(PUB + "mfpt.csv")))[1:]}
    mine = {(r[0], r[1]): (float(r[3]), float(r[4])) for r in list(csv.reader(open(sys.argv[2])))[1:]}
    keys = [k for k in mine if k in pub and np.isfinite(mine[k]).all()]
    a = np.array([mine[k] for k in keys])
    b = np.array([pub[k] for k in keys])
    print("\nfirst-passage times vs published mfpt.csv (%d transitions, %d not fitted)" % (
        len(keys), len(mine) - len(keys)))
    for col, name in ((0, "inner fpt"), (1, "outer fpt")):
        r = np.log(a[:, col] / b[:, col])
        short = b[:, 1] < 30  # short loops: well-sampled outer integrand
        print("%s: median ratio %.3f, 16-84%% of ratio %.3f..%.3f; short loops (%d): median %.3f, "
              "16-84%% %.3f..%.3f; corr(log) %.4f" % (
                  name, np.exp(np.median(r)), np.exp(np.percentile(r, 16)), np.exp(np.percentile(r, 84)),
                  short.sum(), np.exp(np.median(r[short])), np.exp(np.percentile(r[short], 16)),
                  np.exp(np.percentile(r[short], 84)), np.corrcoef(np.log(a[:, col]), np.log(b[:, col]))[0, 1]))
