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

Multi-GPU: launch one process per GPU (torchrun).  Every rank runs ALL transitions of the
layer with `replicas / world` replicas each (the replicas of one transition span the
GPUs, so a 10-transition layer loads 8 GPUs as evenly as a 1260-transition one); the ranks
meet in ONE NCCL all-reduce of the per-state counts per G-test round (library
communicator, `hmc_allreduce_counts`), compute the identical s_bias update redundantly, and
rank 0 writes the files.  The start structures need no exchange: every rank receives the
same `unique_config` row from the program.
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

from .campaign import layers_of, run_transition_ensemble
from .engine import ERR_LIST_OVERFLOW, Comm, init_pos_chain
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


def layer_replicas(replicas, n_transitions, world=1, fill=0):
    """Replicas per transition on ONE rank.  `replicas` is the total per transition over all
    ranks; with fill > 0 thin layers are topped up until about `fill` replicas run per GPU
    (a 10-transition layer would otherwise leave most of the 148 SMs idle; the extra replicas
    cost no time and add samples)."""
    per_rank = -(-replicas // world)
    if fill and n_transitions * per_rank < fill:
        per_rank = -(-fill // n_transitions)
    return per_rank


def warm_start_guess(s_est, bits_in, bits_out):
    """Start value of s_bias[0] for the transition bits_in -> bits_out (it forms bond b):
    the smallest converged value of "b on top of bits_in minus one bond" among the
    transitions of the previous layer found in s_est {(bits_in, bits_out): s_bias[0]};
    None if there is none (layer 0).  The smallest, because the flat-histogram iteration
    climbs quickly from an estimate that is too low (+0.5 ln(c0/c1) per round) but descends
    slowly from one that is too high (-neg_scale ln(norm) while the open state is never
    visited, entropy.cc:46-63)."""
    new = [i for i in range(len(bits_in)) if bits_in[i] != bits_out[i]][0]
    guesses = []
    for c in range(len(bits_in)):
        if bits_in[c]:
            pi = tuple(bool(x) and i != c for i, x in enumerate(bits_in))
            po = tuple(x or i == new for i, x in enumerate(pi))
            if (pi, po) in s_est:
                guesses.append(s_est[(pi, po)])
    return min(guesses) if guesses else None


def run_campaign(master, out_dir, replicas=512, device=0, iters_per_round=None, max_rounds=40,
                 max_layers=None, hist_bins=2048, rank=0, world=1, log=print, wl=True,
                 eq_iters=None, seed_increment=1, stair_bp=None, stair_rc=None, comm=None,
                 fill=0, dist_rows=None, config_int_rows=None, keep_rounds=4, wl_replicas=0, warm_start=False, restart=None):
    """stair_bp / stair_rc: init_json.py's --bp / --rc — the transitions that form bond
    `stair_bp` run as a sequence of stages with shrinking rc, each started from the previous
    stage's structure; stage j of all of a layer's staircase transitions is one ensemble.
    `replicas`: per transition, summed over the ranks of `comm`.
    restart: path of a `restart.npz` written by an earlier (interrupted) run of the same
    campaign: its finished layers are skipped and their hand-off structures / s_bias values
    are taken from the file (the per-transition .h5 files of those layers need not exist).
    warm_start: a transition that forms bond b on top of the bonds B starts its
    flat-histogram iteration from the converged s_bias of "b on top of B minus one bond" of
    the previous layer (the smallest of those values: an estimate that is too low climbs
    faster than one that is too high descends, entropy.cc:46-63) instead of init_s.
    Returns (summary rows, seconds, collision events)."""
    from .params import params_from_json
    os.makedirs(out_dir, exist_ok=True)
    layers = layers_of(master, seed_increment)
    jsons = {(fmt(t[1]), fmt(t[2])): t[3] for t in transition_jsons(master, seed_increment)}
    nb_bonds = len(master["nonlocal_bonds"])
    structures = {tuple([False] * nb_bonds): init_pos_chain(layers[0][1][0][2])}
    rmax = master["length"] * 0.5 * 3 ** 0.5
    stair_bp = sorted(stair_bp) if stair_bp is not None else None
    summary = []
    n_written = 0
    events = 0
    s_est = {}  # (bits_in, bits_out) -> converged s_bias[0], for warm_start
    list_cap_floor = 0
    layers_done = -1
    if restart:
        z = np.load(restart)
        layers_done = int(z["layers_done"])
        unbits = lambda t: tuple(c == "1" for c in t)  # noqa: E731
        for key, pos in zip(z["state_bits"].tolist(), z["state_pos"]):
            structures[unbits(key)] = pos
        for bi, bo, v in zip(z["est_in"].tolist(), z["est_out"].tolist(), z["est_val"].tolist()):
            s_est[(unbits(bi), unbits(bo))] = v
        log("restart: layers 0..%d taken from %s (%d structures)" % (layers_done, restart, len(structures)))
    t_start = time.time()
    for layer, trans in layers:
        if max_layers is not None and layer >= max_layers:
            break
        if layer <= layers_done:
            continue
        mine = list(range(len(trans)))
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
                    fdone = File(path)
                    if stage == len(jobs[k]) - 1:
                        s_est[(trans[k][0], trans[k][1])] = float(fdone["s_bias"][0])
                    up = fdone["unique_config"]["pos"]
                    if len(up):
                        stage_pos[k] = np.array(up[0])
                        if stage == len(jobs[k]) - 1:
                            new_structs.setdefault(trans[k][1], stage_pos[k])
                else:
                    todo.append(k)
            if world > 1:  # every rank must see the same list (rank 0 may already have written)
                todo = [int(k) for k in comm.bcast(np.array(
                    todo + [-1] * (len(mine) - len(todo)), dtype=np.int64)) if k >= 0]
            if not todo:
                continue
            # (sticky: once a layer needed a larger list the deeper, more compact ones start
            # with it instead of running their overflowing transitions twice)
            list_cap = list_cap_floor
            for attempt in range(3):
                params = [jobs[k][stage][0] for k in todo]
                p0 = params[0]
                R = layer_replicas(replicas, len(todo), world, fill)
                # iterations per replica and round follow the NOMINAL replica count: topping a thin
                # layer up adds samples, it must not shorten the trajectories (a slowly mixing
                # transition needs every replica to run several relaxation times per round)
                iters = iters_per_round or max(1, -(-p0.total_iter // max(replicas, 1)))
                # raw rows kept per transition and round (the reference: total_iter * nsteps per
                # round, all rounds; here bounded — the pooled dist_hist is the full statistic)
                rows = min(p0.total_iter * p0.nsteps, 4096) if dist_rows is None else dist_rows
                cfgs = (min(p0.total_iter * (p0.nsteps // max(p0.write_step, 1) + p0.mc_write + 1), 4096)
                        if config_int_rows is None else config_int_rows)
                start_sb = None
                if warm_start and p0.nstates == 2:
                    start_sb = np.zeros((len(todo), 2))
                    for n, k in enumerate(todo):
                        g = warm_start_guess(s_est, trans[k][0], trans[k][1]) if stage == 0 else None
                        start_sb[n, 0] = 3.0 if g is None else g
                t0 = time.time()
                out = run_transition_ensemble(
                    params, [stage_pos[k] for k in todo], R, device=device,
                    seed=(p0.seeds[0] + 7919 * attempt, 1 + layer + 64 * stage), iters_per_round=iters,
                    max_rounds=max_rounds, hist_bins=hist_bins, hist_rmax=rmax, wl=wl,
                    eq_iters=eq_iters, comm=comm, dist_rows=rows, config_int_rows=cfgs,
                    keep_rounds=keep_rounds, wl_replicas=wl_replicas, start_s_bias=start_sb,
                    strict=False, list_capacity=list_cap)
                # a transition stopped by a replica error flag: the reference's process dies with
                # an exception there (main.cc:48-56,219-225) and tools/rerun.py runs it again;
                # nothing of it is written, the loop below comes back to it with another seed
                failed = [k for n, k in enumerate(todo) if out["failed"][n]]
                if failed:
                    if rank == 0:
                        log("layer %d: replica error flags in %s (attempt %d)" % (
                            layer, [(jobs[k][stage][2], int(out["err"][todo.index(k)])) for k in failed],
                            attempt))
                    if any(int(out["err"][todo.index(k)]) & ERR_LIST_OVERFLOW for k in failed):
                        # a compact state holds more live predictions than the default list:
                        # 8 per bead next, then every nonlocal pair (cannot overflow)
                        list_cap = 8 * p0.nbeads if list_cap < 8 * p0.nbeads else p0.nbeads * p0.nbeads
                        list_cap_floor = max(list_cap_floor, 8 * p0.nbeads)
                dt = time.time() - t0
                events += int(out["events"][0])
                for n, k in enumerate(todo):
                    if out["failed"][n]:
                        continue
                    bits_in, bits_out, _ = trans[k]
                    p, dj, name = jobs[k][stage]
                    up = out["unique_pos"][n]
                    summary.append((fmt(bits_in), name.split("_", 3)[3], float(out["s_bias"][n][0])))
                    if stage == len(jobs[k]) - 1:
                        s_est[(bits_in, bits_out)] = float(out["s_bias"][n][0])
                    if up is not None:
                        stage_pos[k] = up
                        if stage == len(jobs[k]) - 1:
                            new_structs.setdefault(bits_out, up)
                    elif stage < len(jobs[k]) - 1:
                        raise RuntimeError("%s: no bonded structure to start the next stair stage from" % name)
                    if rank != 0:
                        continue
                    with open(os.path.join(out_dir, name + ".json"), "w") as f:
                        json.dump(dj, f)
                    uv = out["unique_vel"][n]
                    cr = out["counts_rounds"][n]
                    nl = max(p.n_nonlocal, 1)
                    ds = {"s_bias": out["s_bias"][n].astype(np.float32),  # f32 as main.cc:445-446
                          # ConfigCountWriter: a zero row, then one row per G-test round
                          # (main.cc:394,488)
                          "config_count": np.vstack([np.zeros((1, p.nstates), np.uint64),
                                                     cr if len(cr) else out["counts"][n][None]]),
                          # per-MD-step bond distances of the last rounds (main.cc:199-200); what
                          # tools/mfpt.py:76-77 reads
                          "dist": out["dist"][n].reshape(-1, nl) if "dist" in out else np.zeros((0, nl)),
                          "config/int": (out["config_int"][n] if "config_int" in out
                                         else np.zeros(0, np.uint64)),
                          "dist_hist": out["dist_hist"][n],
                          "unique_config/pos": (up[None] if up is not None
                                                else np.zeros((0, p.nbeads, 3))),
                          "unique_config/vel": (uv[None] if uv is not None
                                                else np.zeros((0, p.nbeads, 3))),
                          "unique_config/config": np.ones(1 if up is not None else 0, np.uint64)}
                    at = {"sigma": np.float64(p.sigma), "length": np.float64(p.length),
                          "nsteps": np.uint32(p.nsteps), "dist_hist_rmax": np.float64(rmax),
                          "replicas": np.uint32(R * world), "err_flags": np.uint32(out["err"][n]),
                          "nonlocal_bonds": np.array(p.bond_list("nonlocal"), np.uint32).reshape(-1, 2),
                          "transient_bonds": np.array(p.bond_list("transient"), np.uint32).reshape(-1, 2),
                          "permanent_bonds": np.array(p.bond_list("permanent"), np.uint32).reshape(-1, 2)}
                    write_output(os.path.join(out_dir, name + ".h5"), ds, at)
                    with open(os.path.join(out_dir, name + ".log"), "w") as f:
                        f.write("replicas %d rounds %d accepted %s s_bias %s counts %s err_flags %d\n" % (
                            R * world, out["rounds"][n], bool(out["accepted"][n]),
                            out["s_bias"][n].tolist(), out["counts"][n].tolist(), int(out["err"][n])))
                if rank == 0:
                    # tools/diff_s_bias.py:72-74 rows, appended batch by batch (an interrupted
                    # campaign keeps what it has)
                    with open(os.path.join(out_dir, "diff_s_bias.csv"), "a", newline="") as f:
                        csv.writer(f).writerows(summary[n_written:])
                    n_written = len(summary)
                    log("layer %d%s: %d transitions, %d replicas each%s, rounds %s, %.1f s (%.1f s device), "
                        "events %d, collectives %d" % (
                            layer, " stage %d" % stage if stage else "", len(todo), R * world,
                            " over %d ranks" % world if world > 1 else "",
                            _brief(out["rounds"]), dt, out["seconds_device"], int(out["events"][0]),
                            out["collectives"]))
                todo = failed
                if not todo:
                    break
            if todo:
                raise RuntimeError("layer %d: %d transitions still raise replica error flags after 3 "
                                   "attempts: %s" % (layer, len(todo), [jobs[k][stage][2] for k in todo]))
        for b_, v in new_structs.items():
            structures.setdefault(b_, v)
        missing = [t[1] for lt in layers[layer + 1:layer + 2] for t in lt[1] if t[0] not in structures]
        if missing:
            raise RuntimeError("layer %d produced no structure for %d states" % (layer, len(set(missing))))
        if rank == 0:  # what a later run needs to go on from here (see `restart`)
            keys = list(structures)
            est = list(s_est.items())
            tmp = os.path.join(out_dir, "restart.tmp.npz")
            np.savez_compressed(tmp, layers_done=layer, state_bits=np.array([fmt(k) for k in keys]),
                                state_pos=np.array([structures[k] for k in keys]),
                                est_in=np.array([fmt(k[0]) for k, _ in est]),
                                est_out=np.array([fmt(k[1]) for k, _ in est]),
                                est_val=np.array([v for _, v in est], dtype=np.float64))
            os.replace(tmp, os.path.join(out_dir, "restart.npz"))
    return summary, time.time() - t_start, events


def _brief(rounds):
    r = np.asarray(rounds)
    return r.tolist() if len(r) <= 12 else "min %d median %d max %d" % (r.min(), np.median(r), r.max())


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
    ap.add_argument("--replicas", type=int, default=512,
                    help="replicas per transition (in total over all ranks)")
    ap.add_argument("--fill", default="0",
                    help="top thin layers up to about this many replicas per GPU "
                         "('auto': what the GPU runs concurrently)")
    ap.add_argument("--total-iter", type=int, default=None,
                    help="override total_iter of the master json (iterations a G-test round "
                         "pools over all replicas of a transition)")
    ap.add_argument("--device", type=int, default=None)
    ap.add_argument("--iters-per-round", type=int, default=None)
    ap.add_argument("--max-rounds", type=int, default=40)
    ap.add_argument("--max-layers", type=int, default=None)
    ap.add_argument("--no-wl", action="store_true", help="skip the per-replica Wang-Landau pre-estimate")
    ap.add_argument("--eq-iters", type=int, default=None)
    ap.add_argument("--hist-bins", type=int, default=2048)
    ap.add_argument("--restart", default=None,
                    help="restart.npz of an interrupted run: skip its finished layers")
    ap.add_argument("--warm-start", action="store_true",
                    help="start s_bias from the same bond's converged value one layer earlier")
    ap.add_argument("--wl-replicas", type=int, default=0,
                    help="replicas per transition that run the Wang-Landau phase (0: all)")
    ap.add_argument("--dist-rows", type=int, default=None,
                    help="raw dist rows kept per transition and round (default min(total_iter*nsteps, 4096))")
    ap.add_argument("--config-int-rows", type=int, default=None)
    args = ap.parse_args(argv)
    if bool(args.rc) != bool(args.bp):
        ap.error("--rc and --bp go together (tools/init_json.py:180-183)")
    with open(args.json) as f:
        master = json.load(f)
    if args.total_iter:
        master["total_iter"] = args.total_iter
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    device = args.device if args.device is not None else int(os.environ.get("LOCAL_RANK", "0"))
    comm = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("gloo")  # passes the 128-byte NCCL id around, nothing else
        comm = Comm.from_torch(device)
    # invoked the way tools/run.py invokes init_json.py: outputs go to the current directory
    out_dir = args.out or ("." if args.exe is not None else
                           os.path.splitext(os.path.basename(args.json))[0])
    from .engine import resident_replicas
    fill = resident_replicas(master["nbeads"], device) if args.fill == "auto" else int(args.fill)
    summary, wall, events = run_campaign(
        master, out_dir, args.replicas, device, args.iters_per_round, args.max_rounds,
        args.max_layers, hist_bins=args.hist_bins, rank=rank, world=world, wl=not args.no_wl,
        eq_iters=args.eq_iters, seed_increment=args.seed_increment, stair_bp=args.bp,
        stair_rc=args.rc, comm=comm, fill=fill, dist_rows=args.dist_rows,
        config_int_rows=args.config_int_rows, wl_replicas=args.wl_replicas, warm_start=args.warm_start, restart=args.restart)
    if rank == 0:
        print("%d transitions on %d GPU(s) in %.1f s -> %.0f transitions/hour, %.3e collision "
              "events (%.3e events/s whole campaign)" % (
                  len(summary), world, wall, 3600.0 * len(summary) / max(wall, 1e-9), events,
                  events / max(wall, 1e-9)))
    if comm is not None:
        comm.close()


if __name__ == "__main__":
    main()
