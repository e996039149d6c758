"""Batched campaign launcher — the resident-process counterpart of the reference's
`tools/run.py` + `tools/init_json.py` (one OS process per transition): every transition of
a lattice layer runs as ONE GPU ensemble, layers run in order, and each transition leaves
the files the reference leaves, with the reference's names
(`hybridmc_<layer>_<bits_in>_<bits_out>.{json,h5,log}`, tools/init_json.py:109-125), so
`diff_s_bias.py`-style collection works on the result.  `.h5` files are HDF5 with the
reference's dataset names (hybridmc_b200/output.py, written by the library's own writer).

  python -m hybridmc_b200.run_campaign master.json [--replicas R] [--out DIR] [--device D]
                                       [--iters-per-round K] [--max-rounds N] [--max-layers L]

It also accepts init_json.py's own command line (`json exe seed_increment [--nproc N]`,
tools/init_json.py:174-183), writing into the current directory as init_json.py does, so
the reference's `tools/run.py master.json <exe> <this file>` drives it unchanged (`exe`
and `--nproc` are accepted and unused: the transitions run inside this process).

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

if __package__ in (None, ""):  # started by path, the way tools/run.py starts init_json.py
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import hybridmc_b200  # noqa: F401
    __package__ = "hybridmc_b200"

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


def stair_stages(d, layer, master_rc, stair_rc):
    """The run sequence of a staircase transition (tools/init_json.py:93-120): run j uses
    rc = stair_rc[j] with the previous radius as outer `stair` wall; permanent bonds keep the
    master's rc through `p_rc`; every run but the last carries its rc in the file name."""
    out = []
    for j, rc in enumerate(stair_rc):
        dj = dict(d)
        if layer > 0:
            dj["p_rc"] = master_rc
        if j > 0:
            dj["stair_bonds"] = [d["transient_bonds"][0]]
            dj["stair"] = stair_rc[j - 1]
        dj["rc"] = rc
        out.append((dj, "_%s" % rc if j < len(stair_rc) - 1 else ""))
    return out


def run_campaign(master, out_dir, replicas=512, device=0, iters_per_round=None, max_rounds=40,
                 max_layers=None, hist_bins=2048, rank=0, world=1, log=print, wl=True,
                 eq_iters=None, seed_increment=1, stair_bp=None, stair_rc=None):
    """stair_bp / stair_rc: init_json.py's --bp / --rc — the transitions that form bond
    `stair_bp` run as a sequence of stages with shrinking rc, each started from the previous
    stage's structure; stage j of all of a layer's staircase transitions is one ensemble."""
    from .params import params_from_json
    os.makedirs(out_dir, exist_ok=True)
    layers = layers_of(master, seed_increment)
    jsons = {(fmt(t[1]), fmt(t[2])): t[3] for t in transition_jsons(master, seed_increment)}
    nb_bonds = len(master["nonlocal_bonds"])
    structures = {tuple([False] * nb_bonds): init_pos_chain(layers[0][1][0][2])}
    rmax = master["length"] * 0.5 * 3 ** 0.5
    stair_bp = sorted(stair_bp) if stair_bp is not None else None
    summary = []
    t_start = time.time()
    for layer, trans in layers:
        if max_layers is not None and layer >= max_layers:
            break
        mine = shard(len(trans), rank, world)
        # jobs[k] = the stages of transition k: [(params, json, output name)]
        jobs = {}
        for k in mine:
            bits_in, bits_out, p = trans[k]
            d = jsons[(fmt(bits_in), fmt(bits_out))]
            base = "hybridmc_%d_%s_%s" % (layer, fmt(bits_in), fmt(bits_out))
            if stair_bp is not None and sorted(d["transient_bonds"][0]) == stair_bp:
                jobs[k] = [(params_from_json(dj), dj, base + suffix)
                           for dj, suffix in stair_stages(d, layer, master["rc"], stair_rc)]
            else:
                jobs[k] = [(p, d, base)]
        new_structs = {}
        stage_pos = {k: structures[trans[k][0]] for k in mine}
        for stage in range(max(len(j) for j in jobs.values()) if jobs else 0):
            # finished runs are skipped (init_json.py:141-142) but still hand their structure on
            todo = []
            for k in mine:
                if stage >= len(jobs[k]):
                    continue
                path = os.path.join(out_dir, jobs[k][stage][2] + ".h5")
                if os.path.exists(path):
                    up = File(path)["unique_config"]["pos"]
                    if len(up):
                        stage_pos[k] = np.array(up[0])
                        if stage == len(jobs[k]) - 1:
                            new_structs.setdefault(trans[k][1], stage_pos[k])
                else:
                    todo.append(k)
            if not todo:
                continue
            params = [jobs[k][stage][0] for k in todo]
            t0 = time.time()
            out = run_transition_ensemble(params, [stage_pos[k] for k in todo], replicas,
                                          device=device, seed=(params[0].seeds[0], 1 + layer + 64 * stage),
                                          iters_per_round=iters_per_round, max_rounds=max_rounds,
                                          hist_bins=hist_bins, hist_rmax=rmax, wl=wl,
                                          eq_iters=eq_iters)
            dt = time.time() - t0
            for n, k in enumerate(todo):
                bits_in, bits_out, _ = trans[k]
                p, dj, name = jobs[k][stage]
                with open(os.path.join(out_dir, name + ".json"), "w") as f:
                    json.dump(dj, f)
                up = out["unique_pos"][n]
                ds = {"s_bias": out["s_bias"][n].astype(np.float32),  # f32 as main.cc:445-446
                      "config_count": np.vstack([np.zeros(2, np.uint64), out["counts"][n]]),
                      "dist_hist": out["dist_hist"][n],
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
                        replicas, out["rounds"][n], bool(out["accepted"][n]),
                        out["s_bias"][n].tolist(), out["counts"][n].tolist()))
                # one row per run; a stair stage carries its rc like its file name does
                summary.append((fmt(bits_in), name.split("_", 3)[3], float(out["s_bias"][n][0])))
                if up is not None:
                    stage_pos[k] = up
                    if stage == len(jobs[k]) - 1:
                        new_structs.setdefault(bits_out, up)
                elif stage < len(jobs[k]) - 1:
                    raise RuntimeError("%s: no bonded structure to start the next stair stage from" % name)
            log("layer %d%s: %d transitions, %d replicas each, rounds %s, %.1f s, events %d, err %d" % (
                layer, " stage %d" % stage if stage else "", len(todo), replicas,
                out["rounds"].tolist(), dt, int(out["events"][0]),
                int(np.bitwise_or.reduce(out["err"]))))
        if world > 1:
            import torch.distributed as dist
            gathered = [None] * world
            dist.all_gather_object(gathered, new_structs)
            for g in gathered:
                for b_, v in g.items():
                    new_structs.setdefault(b_, v)
        for b_, v in new_structs.items():
            structures.setdefault(b_, v)
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
    ap.add_argument("exe", nargs="?", default=None,
                    help="(init_json.py compatibility) hybridmc executable; unused")
    ap.add_argument("seed_increment", nargs="?", type=int, default=1,
                    help="(init_json.py compatibility) amount to increment the seed with every run")
    ap.add_argument("--nproc", type=int, default=None, help="(init_json.py compatibility) unused")
    ap.add_argument("--rc", nargs="+", type=float, help="list of rc if potential is step")
    ap.add_argument("--bp", nargs=2, type=int, help="bead pair whose potential is step")
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
    if bool(args.rc) != bool(args.bp):
        ap.error("--rc and --bp go together (tools/init_json.py:180-183)")
    with open(args.json) as f:
        master = json.load(f)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    device = args.device if args.device is not None else int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("gloo")  # host-side hand-off of structures only
    # invoked the way tools/run.py invokes init_json.py: outputs go to the current directory
    out_dir = args.out or ("." if args.exe is not None else
                           os.path.splitext(os.path.basename(args.json))[0])
    summary, wall = run_campaign(master, out_dir, args.replicas, device, args.iters_per_round,
                                 args.max_rounds, args.max_layers, hist_bins=args.hist_bins, rank=rank,
                                 world=world, wl=not args.no_wl, eq_iters=args.eq_iters,
                                 seed_increment=args.seed_increment, stair_bp=args.bp, stair_rc=args.rc)
    print("rank %d: %d transitions in %.1f s -> %.0f transitions/hour" % (
        rank, len(summary), wall, 3600.0 * len(summary) / max(wall, 1e-9)))


if __name__ == "__main__":
    main()
