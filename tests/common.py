"""Shared helpers for the parity tests (oracle side + engine side)."""
import ctypes as C
import json
import os

import numpy as np

from hybridmc_b200.params import params_from_json, transition_jsons
from oracle_lib import OrcMt, Sim, load_oracle

HERE = os.path.dirname(os.path.abspath(__file__))
CONFIGS = os.path.join(HERE, "golden", "configs")


def load_config(name, **overrides):
    with open(os.path.join(CONFIGS, name + ".json")) as f:
        d = json.load(f)
    d.update(overrides)
    return d


def transition_params(name, index=0, **overrides):
    """params of the index-th per-transition run the campaign driver generates"""
    d = load_config(name)
    layer, bits_in, bits_out, dj = transition_jsons(d)[index]
    dj.update(overrides)
    return params_from_json(dj)


class OracleRun:
    """One oracle replica + its restated std::mt19937, seeded like main.cc:379-380."""

    def __init__(self, params, lib=None):
        self.lib = lib or load_oracle()
        self.p = params
        self.sim = Sim(self.lib, "orc", params)
        self.g = OrcMt()
        seeds = (C.c_uint32 * params.n_seeds)(*list(params.seeds)[:params.n_seeds])
        self.lib.orc_mt_seed_seq(C.byref(self.g), seeds, params.n_seeds)
        self.vsig = float(np.sqrt(params.temp / params.m))

    def init_pos(self):
        pos = np.zeros((self.p.nbeads, 3))
        rc = self.lib.orc_init_pos(C.byref(self.p), C.byref(self.g), pos.ctypes.data_as(C.c_void_p))
        assert rc == 0
        self.sim.set_pos(pos)
        return pos

    def normals(self, nsteps):
        out = np.zeros((nsteps, self.p.nbeads, 3))
        for s in range(nsteps):
            self.lib.orc_mt_normals(C.byref(self.g), C.c_double(self.vsig),
                                    out[s].ctypes.data_as(C.c_void_p), 3 * self.p.nbeads)
        return out

    def tape(self, n):
        words = np.zeros(n, dtype=np.uint32)
        st, ct = np.zeros(n), np.zeros(n)
        self.lib.orc_mt_tape(C.byref(self.g), words.ctypes.data_as(C.c_void_p),
                             st.ctypes.data_as(C.c_void_p), ct.ctypes.data_as(C.c_void_p), n)
        return words, st, ct

    def discard(self, n):
        self.lib.orc_mt_discard(C.byref(self.g), C.c_uint(int(n)))

    def crankshaft(self, move_offset, n_moves, words, st, ct):
        cur = C.c_uint(0)
        err = self.lib.orc_crankshaft(
            self.sim.h, C.c_uint(move_offset), C.c_uint(n_moves),
            words.ctypes.data_as(C.c_void_p), st.ctypes.data_as(C.c_void_p),
            ct.ctypes.data_as(C.c_void_p), C.c_uint(len(words)), C.byref(cur))
        return int(err), int(cur.value)


def first_trace_mismatch(a, b):
    n = min(len(a), len(b))
    for k in range(n):
        if a[k].tobytes() != b[k].tobytes():
            return k
    return None if len(a) == len(b) else n
