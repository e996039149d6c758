"""Batched campaign launcher — the resident-process counterpart of the reference's
`tools/run.py` + `tools/init_json.py` (one OS process per transition): every transition of
a lattice layer runs as ONE GPU ensemble, layers run in order, and each transition leaves
the files the reference leaves, with the reference's names
(`hybridmc_<layer>_<bits_in>_<bits_out>.{json,h5,log}`, tools/init_json.py:109-125), so
`diff_s_bias.py`-style collection works on the result.  `.h5` files are HDF5 with the
reference's dataset names (hybridmc_b200/output.py, written by the library's own writer).

  python -m hybridmc_b200.run_campaign master.json [--replicas R] [--out DIR] [--device D]
                                       [--iters-per-round K] [--max-rounds N] [--max-layers L]

Multi-GPU: launch one process per GPU (torchrun); the transitions of a layer are dealt
round-robin to the ranks (campaign.shard) and the harvested start structures of the next
layer are exchanged with one all-gather per layer.
"""
import argparse
import csv
import json
import os
import sys
import time

import numpy as np

from .campaign import layers_of, run_transition_ensemble, shard
from .engine import init_pos_chain
from .output import File, write_hdf5
from .params import transition_jsons

def write_output(path, datasets, attrs):
    """HDF5 through the library's writer, temporary file + rename as main.cc:383,495"""
    write_hdf5(path + ".tmp", datasets, attrs)
    os.replace(path + ".tmp", path)


def fmt(bits):
    return "".join("1" if b else "0" for b in bits)


def run_campaign(master, out_dir, replicas=512, device=0, iters_per_round=None, max_rounds=40,
                 max_layers=None, hist_bins=2048, rank=0, world=1, log=print, wl=True,
                 eq_iters=None):
    os.makedirs(out_dir, exist_ok=True)
    layers = layers_of(master)
    jsons = {(fmt(t[1]), fmt(t[2])): t[3] for t in transition_jsons(master)}
    nb_bonds = len(master["nonlocal_bonds"])
    structures = {tuple([False] * nb_bonds): init_pos_chain(layers[0][1][0][2])}
    rmax = master["length"] * 0.5 * 3 ** 0.5
    summary = []
    t_start = time.time()
    for layer, trans in layers:
        if max_layers is not None and layer >= max_layers:
            break
        mine = shard(len(trans), rank, world)
        todo = [trans[k] for k in mine
                if not os.path.exists(os.path.join(out_dir, "hybridmc_%d_%s_%s.h5" % (
                    layer, fmt(trans[k][0]), fmt(trans[k][1]))))]  # init_json.py:141-142
        new_structs = {}
        if todo:
            params = [t[2] for t in todo]
            t0 = time.time()
            out = run_transition_ensemble(params, [structures[t[0]] for t in todo], replicas,
                                          device=device, seed=(params[0].seeds[0], 1 + layer),
                                          iters_per_round=iters_per_round, max_rounds=max_rounds,
                                          hist_bins=hist_bins, hist_rmax=rmax, wl=wl,
                                          eq_iters=eq_iters)
            dt = time.time() - t0
            for k, (bits_in, bits_out, p) in enumerate(todo):
                name = "hybridmc_%d_%s_%s" % (layer, fmt(bits_in), fmt(bits_out))
                with open(os.path.join(out_dir, name + ".json"), "w") as f:
                    json.dump(jsons[(fmt(bits_in), fmt(bits_out))], f)
                up = out["unique_pos"][k]
                ds = {"s_bias": out["s_bias"][k].astype(np.float32),  # f32 as main.cc:445-446
                      "config_count": np.vstack([np.zeros(2, np.uint64), out["counts"][k]]),
                      "dist_hist": out["dist_hist"][k],
                      "unique_config/pos": (up[None] if up is not None
                                            else np.zeros((0, p.nbeads, 3)))}
                at = {"sigma": np.float64(p.sigma), "length": np.float64(p.length),
                      "nsteps": np.uint32(p.nsteps), "dist_hist_rmax": np.float64(rmax),
                      "nonlocal_bonds": np.array(p.bond_list("nonlocal"), np.uint32).reshape(-1, 2),
                      "transient_bonds": np.array(p.bond_list("transient"), np.uint32).reshape(-1, 2),
                      "permanent_bonds": np.array(p.bond_list("permanent"), np.uint32).reshape(-1, 2)}
                write_output(os.path.join(out_dir, name + ".h5"), ds, at)
                with open(os.path.join(out_dir, name + ".log"), "w") as f:
                    f.write("replicas %d rounds %d accepted %s s_bias %s counts %s\n" % (
                        replicas, out["rounds"][k], bool(out["accepted"][k]),
                        out["s_bias"][k].tolist(), out["counts"][k].tolist()))
                summary.append((fmt(bits_in), fmt(bits_out), float(out["s_bias"][k][0])))
                if up is not None:
                    new_structs.setdefault(bits_out, up)
            log("layer %d: %d transitions, %d replicas each, rounds %s, %.1f s, events %d, err %d" % (
                layer, len(todo), replicas, out["rounds"].tolist(), dt, int(out["events"][0]),
                int(np.bitwise_or.reduce(out["err"]))))
        # finished transitions of earlier runs still hand their structure on
        for k in mine:
            bits_in, bits_out, p = trans[k]
            if bits_out in new_structs:
                continue
            path = os.path.join(out_dir, "hybridmc_%d_%s_%s.h5" % (layer, fmt(bits_in), fmt(bits_out)))
            if os.path.exists(path):
                up = File(path)["unique_config"]["pos"]
                if len(up):
                    new_structs.setdefault(bits_out, np.array(up[0]))
        if world > 1:
            import torch.distributed as dist
            gathered = [None] * world
            dist.all_gather_object(gathered, new_structs)
            for g in gathered:
                for b, v in g.items():
                    new_structs.setdefault(b, v)
        for b, v in new_structs.items():
            structures.setdefault(b, v)
        missing = [t[1] for lt in layers[layer + 1:layer + 2] for t in lt[1] if t[0] not in structures]
        if missing:
            raise RuntimeError("layer %d produced no structure for %d states" % (layer, len(set(missing))))
    if rank == 0:
        with open(os.path.join(out_dir, "diff_s_bias.csv"), "a", newline="") as f:
            csv.writer(f).writerows(summary)  # tools/diff_s_bias.py:72-74
    return summary, time.time() - t_start


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("json", help="master json input file")
    ap.add_argument("--out", default=None, help="output directory (default: <json name>)")
    ap.add_argument("--replicas", type=int, default=512)
    ap.add_argument("--device", type=int, default=None)
    ap.add_argument("--iters-per-round", type=int, default=None)
    ap.add_argument("--max-rounds", type=int, default=40)
    ap.add_argument("--max-layers", type=int, default=None)
    ap.add_argument("--no-wl", action="store_true", help="skip the per-replica Wang-Landau pre-estimate")
    ap.add_argument("--eq-iters", type=int, default=None)
    ap.add_argument("--hist-bins", type=int, default=2048)
    args = ap.parse_args(argv)
    with open(args.json) as f:
        master = json.load(f)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    device = args.device if args.device is not None else int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("gloo")  # host-side hand-off of structures only
    out_dir = args.out or os.path.splitext(os.path.basename(args.json))[0]
    summary, wall = run_campaign(master, out_dir, args.replicas, device, args.iters_per_round,
                                 args.max_rounds, args.max_layers, hist_bins=args.hist_bins, rank=rank,
                                 world=world, wl=not args.no_wl, eq_iters=args.eq_iters)
    print("rank %d: %d transitions in %.1f s -> %.0f transitions/hour" % (
        rank, len(summary), wall, 3600.0 * len(summary) / max(wall, 1e-9)))


if __name__ == "__main__":
    main()
