"""Host-side mirror of the reference's configuration layer.

`HmcParams` is the ctypes image of `hmc_params` (include/hybridmc_b200.h), the
POD copy of the reference's `Param` (src/params.h:10-101).  `params_from_json`
follows `from_json(Param)` (src/main.cc:500-562): same keys, same optional keys
(`stair`, `stair_bonds`, `p_rc`), unknown keys (`config_in`, `config_out`)
ignored, bond pairs normalised to i<j (src/config.cc:6-13), bit index = list
position.  `transition_params` enumerates the per-transition parameter sets the
campaign driver writes (tools/init_json.py:43-66,133-139).
"""
import ctypes as C
import json

HMC_MAX_BONDS = 64
HMC_MAX_BEADS = 64
HMC_MAX_SEEDS = 8

_PAIRS = (C.c_uint32 * 2) * HMC_MAX_BONDS


class HmcParams(C.Structure):
    _fields_ = [
        ("m", C.c_double), ("sigma", C.c_double),
        ("near_min", C.c_double), ("near_max", C.c_double),
        ("nnear_min", C.c_double), ("nnear_max", C.c_double),
        ("rh", C.c_double), ("rc", C.c_double),
        ("stair", C.c_double), ("p_rc", C.c_double),
        ("length", C.c_double),
        ("del_t", C.c_double), ("del_t_wl", C.c_double),
        ("gamma", C.c_double), ("gamma_f", C.c_double),
        ("temp", C.c_double),
        ("pos_scale", C.c_double), ("neg_scale", C.c_double), ("sig_level", C.c_double),
        ("tries", C.c_uint64),
        ("has_stair", C.c_uint32), ("has_p_rc", C.c_uint32),
        ("nbeads", C.c_uint32), ("ncell", C.c_uint32),
        ("nsteps", C.c_uint32), ("nsteps_eq", C.c_uint32),
        ("nsteps_wl", C.c_uint32), ("write_step", C.c_uint32),
        ("mc_moves", C.c_uint32), ("mc_write", C.c_uint32),
        ("total_iter", C.c_uint32), ("total_iter_eq", C.c_uint32),
        ("max_g_test_count", C.c_uint32), ("max_nbonds", C.c_uint32),
        ("n_nonlocal", C.c_uint32), ("n_transient", C.c_uint32),
        ("n_permanent", C.c_uint32), ("n_seeds", C.c_uint32),
        ("nonlocal_bonds", _PAIRS), ("transient_bonds", _PAIRS),
        ("permanent_bonds", _PAIRS),
        ("seeds", C.c_uint32 * HMC_MAX_SEEDS),
    ]

    @property
    def nstates(self):
        return 1 << self.n_transient

    def bond_list(self, which):
        arr = getattr(self, which + "_bonds")
        n = getattr(self, "n_" + which)
        return [(arr[k][0], arr[k][1]) for k in range(n)]


def _set_bonds(p, which, pairs):
    if len(pairs) > HMC_MAX_BONDS:
        raise ValueError("more than %d %s bonds" % (HMC_MAX_BONDS, which))
    arr = getattr(p, which + "_bonds")
    for k, (i, j) in enumerate(pairs):
        i, j = int(i), int(j)
        if i > j:  # NonlocalBonds ctor, src/config.cc:6-13
            i, j = j, i
        arr[k][0], arr[k][1] = i, j
    setattr(p, "n_" + which, len(pairs))


def params_from_json(data):
    """dict (or path to a JSON file) -> HmcParams, as src/main.cc:500-562."""
    if isinstance(data, str):
        with open(data) as f:
            data = json.load(f)
    p = HmcParams()
    p.m = data["m"]
    p.sigma = data["sigma_bb"]
    for k in ("near_min", "near_max", "nnear_min", "nnear_max", "rh", "rc"):
        setattr(p, k, data[k])
    if "stair" in data:
        p.stair, p.has_stair = data["stair"], 1
    # the reference dereferences p_rc whenever stair is set (main.cc:535, UB when
    # absent); here an absent p_rc is simply "no permanent-bond radius".
    if "p_rc" in data:
        p.p_rc, p.has_p_rc = data["p_rc"], 1
    if p.has_stair and (p.stair < p.rc or (p.has_p_rc and p.stair < p.p_rc)):
        raise RuntimeError("stair boundary is smaller than rc")  # main.cc:535-537
    _set_bonds(p, "nonlocal", data["nonlocal_bonds"])
    _set_bonds(p, "transient", data["transient_bonds"])
    _set_bonds(p, "permanent", data["permanent_bonds"])
    p.tries = data["tries"]
    p.nbeads = data["nbeads"]
    p.length = data["length"]
    p.ncell = data["ncell"]
    p.nsteps = data["nsteps"]
    p.nsteps_eq = data["nsteps_eq"]
    p.del_t = data["del_t"]
    p.nsteps_wl = data["nsteps_wl"]
    p.del_t_wl = data["del_t_wl"]
    p.gamma = data["gamma"]
    p.gamma_f = data["gamma_f"]
    p.write_step = data["write_step"]
    seeds = list(data["seeds"])
    if len(seeds) > HMC_MAX_SEEDS:
        raise ValueError("too many seeds")
    for k, s in enumerate(seeds):
        p.seeds[k] = s
    p.n_seeds = len(seeds)
    p.temp = data["temp"]
    p.mc_moves = data["mc_moves"]
    p.mc_write = data["mc_write"]
    p.total_iter = data["total_iter"]
    p.total_iter_eq = data["total_iter_eq"]
    p.pos_scale = data["pos_scale"]
    p.neg_scale = data["neg_scale"]
    p.sig_level = data["sig_level"]
    p.max_nbonds = data["max_nbonds"]
    p.max_g_test_count = data["max_g_test_count"]
    if p.nbeads > HMC_MAX_BEADS:
        raise ValueError("nbeads > %d" % HMC_MAX_BEADS)
    return p


def transition_jsons(master, seed_increment=1):
    """Per-transition JSON dicts in the order tools/init_json.py enqueues them
    (BFS over the bond-state lattice from the empty state; `count` is the
    running seed counter incremented for every (state, bond) pair, also the
    bonded ones).  Returns [(layer, bits_in, bits_out, dict)]."""
    data = dict(master)
    bonds = sorted(sorted(b) for b in data["nonlocal_bonds"])
    data["nonlocal_bonds"] = bonds
    nb = len(bonds)
    out, seen, frontier, count = [], set(), [tuple([False] * nb)], 0
    while frontier:
        nxt = []
        for bits_in in frontier:
            if bits_in in seen:
                continue
            seen.add(bits_in)
            for i in range(nb):
                count += 1
                if bits_in[i]:
                    continue
                bits_out = list(bits_in)
                bits_out[i] = True
                bits_out = tuple(bits_out)
                d = dict(data)
                d["transient_bonds"] = [bonds[i]]
                d["permanent_bonds"] = [bonds[k] for k in range(nb) if bits_in[k]]
                fmt = lambda b: "".join("1" if x else "0" for x in b)
                d["config_in"] = int(fmt(bits_in), 2)
                d["config_out"] = int(fmt(bits_out), 2)
                d["seeds"] = [count, seed_increment]
                out.append((sum(bits_in), bits_in, bits_out, d))
                nxt.append(bits_out)
        frontier = nxt
    return out
