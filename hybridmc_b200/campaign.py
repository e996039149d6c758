"""Layer scheduler over the bond-state lattice — the batched counterpart of the
reference's campaign driver (tools/init_json.py:14-130): the transitions of one
lattice layer (same number of bonds already formed) are independent and run as ONE
ensemble launch; layer L+1 starts from a structure harvested from layer L in which
the new bond is formed (the reference passes it through `unique_config/pos` of the
previous layer's HDF5 file, src/snapshot.cc:122-131).

The per-transition estimator itself (the phase loop of the reference's main,
src/main.cc:411-491: equilibrate -> Wang-Landau -> G-test loop, over R replicas with
pooled counts) is C++: `hmc_run_program_ensemble` in csrc/hmc_program.cu.
"""
import numpy as np

from .engine import (PHASE_EQ, PHASE_MAIN, Engine, compute_entropy, g_test, init_pos_chain,
                     run_program_ensemble)
from .params import params_from_json, transition_jsons


def layers_of(master_json, seed_increment=1):
    """[(layer, [(bits_in, bits_out, HmcParams)])] in BFS order."""
    out = {}
    for layer, bits_in, bits_out, d in transition_jsons(master_json, seed_increment):
        out.setdefault(layer, []).append((bits_in, bits_out, params_from_json(d)))
    return [(k, out[k]) for k in sorted(out)]


def harvest_start_structures(master_json, device=0, replicas=128, eq_iters=4, max_rounds=400,
                             seed=12345, verbose=False):
    """Walk the lattice layer by layer with a small ensemble per transition and return
    {bits: pos[nbeads][3]} — one valid structure per bond state (the empty state from
    init_pos, every other one the first recorded configuration with the transient bond
    formed, i.e. the reference's unique_config/pos row)."""
    layers = layers_of(master_json)
    nb_bonds = len(master_json["nonlocal_bonds"])
    empty = tuple([False] * nb_bonds)
    p0 = layers[0][1][0][2]
    structures = {empty: init_pos_chain(p0)}
    for layer, trans in layers:
        params = [t[2] for t in trans]
        with Engine(params, replicas, device) as e:
            pos = np.stack([np.broadcast_to(structures[t[0]], (replicas,) + structures[t[0]].shape)
                            for t in trans]).reshape(e.R, e.nb, 3)
            e.set_positions(pos)
            e.init_config_from_positions()
            e.init_s()
            e.seed(seed + layer, 77)
            e.reset_clocks()
            e.run(PHASE_EQ, 0, eq_iters)
            e.init_config_from_positions()
            e.reset_clocks()
            missing = set(range(len(trans)))
            it = 0
            while missing and it < max_rounds:
                e.run(PHASE_MAIN, it, 1)
                it += 1
                upos, _, found = e.get_unique_config()
                found = found.reshape(len(trans), replicas)
                for t in list(missing):
                    w = np.nonzero(found[t])[0]
                    if len(w):
                        structures.setdefault(trans[t][1], upos.reshape(
                            len(trans), replicas, e.nb, 3)[t, w[0]].copy())
                        missing.discard(t)
            _, _, _, err = e.get_counts()
            if verbose:
                print("layer", layer, "transitions", len(trans), "rounds", it, "missing",
                      len(missing), "err", int(np.bitwise_or.reduce(err)))
            if missing:
                raise RuntimeError("no bonded configuration found for %d transitions of layer %d"
                                   % (len(missing), layer))
    return structures


def run_transition_ensemble(params_list, start_pos, replicas, device=0, seed=(1, 2),
                            iters_per_round=None, eq_iters=None, max_rounds=None, wl=True,
                            hist_bins=0, hist_rmax=0.0, verbose=False, comm=None, **kw):
    """The reference's per-transition program (main.cc:411-491) for a batch of
    transitions of one layer, R replicas each (per rank of `comm`), pooled flat-histogram
    updates: a thin front of the library's C++ phase loop `hmc_run_program_ensemble`
    (engine.run_program_ensemble), which is the only implementation of it.

    Per G-test round every replica runs `iters_per_round` iterations of run_trajectory
    (default total_iter / (R * world) rounded up, so that a round pools about as many
    recorded configurations as one reference round).  Raises engine.EnsembleError when a
    replica raised an error flag (the reference throws at the same points)."""
    return run_program_ensemble(params_list, start_pos, replicas, device=device, seed=seed,
                                iters_per_round=iters_per_round or 0,
                                eq_iters=eq_iters or 0, max_rounds=max_rounds or 0, wl=wl,
                                hist_bins=hist_bins, hist_rmax=hist_rmax, verbose=verbose,
                                comm=comm, **kw)


def stair_entropy(stage_s_bias):
    """Entropy difference of a staircase transition from the s_bias[0] of its stages, as
    tools/diff_s_bias_stair.py:47-58 combines them: log of the sum of exp(sum of subset)
    over all non-empty subsets of the stage values = log(prod(1 + e^v) - 1)."""
    v = np.asarray(stage_s_bias, dtype=np.float64)
    return float(np.log(np.prod(1.0 + np.exp(v)) - 1.0))


def collect_stair(rows):
    """diff_s_bias.csv rows (bits_in, bits_out[_rc], s) -> {(bits_in, bits_out): value}: plain
    transitions keep their value, the stages of a staircase transition are combined with
    stair_entropy (tools/diff_s_bias_stair.py:20-58)."""
    stages = {}
    for bits_in, tail, s in rows:
        bits_out = tail.split("_")[0]
        stages.setdefault((bits_in, bits_out), []).append(float(s))
    return {k: (v[0] if len(v) == 1 else stair_entropy(v)) for k, v in stages.items()}


def state_entropies(nbonds, diffs):
    """Per-state entropies from per-transition differences, as tools/avg_s_bias.py:36-72 forms
    them: S(all bonds formed) = 0; walking the lattice one broken bond at a time, the entropy
    of a state is the MEAN over the bonds b it lacks of  S(state + b) + diff[(state, state + b)]
    (each parent's own mean is final before its children are visited, because a layer is
    complete before the next one starts).  diffs: {(bits_in, bits_out): s_bias[0]} with bit
    strings; returns {bits: S}."""
    full = "1" * nbonds
    S = {full: 0.0}
    layer = [full]
    while layer:
        acc = {}
        for parent in layer:
            for b in range(nbonds):
                if parent[b] == "1":
                    child = parent[:b] + "0" + parent[b + 1:]
                    acc.setdefault(child, []).append(diffs[(child, parent)] + S[parent])
        for child, vals in acc.items():
            S[child] = float(np.mean(vals))
        layer = sorted(acc)
    return S


# ---- multi-GPU: one process per GPU, shard + one all-reduce per G-test round ----------

def shard(n_items, rank, world):
    """Round-robin assignment of the independent units (transitions of a layer, or replica
    blocks of one transition) to ranks: rank r owns items r, r+world, ..."""
    return list(range(rank, n_items, world))


def allreduce_counts(counts):
    """Sum the per-state counts `counts` (integer array, any shape) over all ranks of the
    default torch.distributed group.  Integer sums are exact and order-independent, so
    every rank then computes the identical s_bias update redundantly (no broadcast).
    Works on CPU (gloo) and on a device accumulator view (nccl)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return counts
    if isinstance(counts, torch.Tensor):
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
        return counts
    t = torch.from_numpy(np.ascontiguousarray(counts).astype(np.int64))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.numpy().astype(np.uint64).reshape(np.shape(counts))


def pooled_round(local_counts, s_bias, p):
    """One G-test round of the phase loop on pooled counts (main.cc:471-491):
    all-reduce, compute_entropy, g_test.  Returns (s_bias, accepted, pooled_counts)."""
    pooled = allreduce_counts(np.asarray(local_counts, dtype=np.uint64))
    out = np.array(s_bias, dtype=np.float64).reshape(pooled.shape)
    acc = np.zeros(len(pooled), dtype=bool) if pooled.ndim == 2 else None
    if pooled.ndim == 1:
        return compute_entropy(pooled, out, p.pos_scale, p.neg_scale), g_test(pooled, p.sig_level)[0], pooled
    for t in range(len(pooled)):
        out[t] = compute_entropy(pooled[t], out[t], p.pos_scale, p.neg_scale)
        acc[t] = g_test(pooled[t], p.sig_level)[0]
    return out, acc, pooled
