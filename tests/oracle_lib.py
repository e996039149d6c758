"""ctypes loaders for the TEST-ONLY oracle libraries (oracle/libhmc_oracle.so and,
when built, oracle/_ref/libhmc_ref.so).  Product code never imports this."""
import ctypes as C
import os
import subprocess

import numpy as np

from hybridmc_b200.params import HmcParams

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")


class HmcEvent(C.Structure):
    _fields_ = [("t", C.c_double), ("type", C.c_uint32), ("i", C.c_uint32),
                ("j", C.c_uint32), ("pad", C.c_uint32)]


EVENT_DTYPE = np.dtype([("t", "<f8"), ("type", "<u4"), ("i", "<u4"), ("j", "<u4"),
                        ("pad", "<u4")])


class OrcMainResult(C.Structure):
    _fields_ = [("s_bias", C.c_double * 16), ("config_count", C.c_uint64 * 16),
                ("collisions", C.c_uint64), ("cell_crossings", C.c_uint64),
                ("heap_pops", C.c_uint64), ("predictions", C.c_uint64),
                ("gtest_iters", C.c_uint32), ("accepted", C.c_int), ("err", C.c_int),
                ("seconds_md", C.c_double)]


class RefMainResult(C.Structure):
    _fields_ = [("s_bias", C.c_double * 16), ("config_count", C.c_uint64 * 16),
                ("gtest_iters", C.c_uint32), ("accepted", C.c_int), ("err", C.c_int),
                ("seconds_md", C.c_double), ("pos_final", C.c_double * (64 * 3)),
                ("config_final", C.c_uint64)]


class OrcMt(C.Structure):
    _fields_ = [("mt", C.c_uint32 * 624), ("idx", C.c_int)]


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def build_oracle():
    subprocess.run(["make", "-s", "-C", ORACLE_DIR, "liboracle"], check=True)


def load_oracle():
    path = os.path.join(ORACLE_DIR, "libhmc_oracle.so")
    if not os.path.exists(path):
        build_oracle()
    lib = C.CDLL(path)
    lib.orc_create.restype = C.c_void_p
    lib.orc_create.argtypes = [C.POINTER(HmcParams)]
    lib.orc_get_config.restype = C.c_uint64
    lib.orc_config_int.restype = C.c_uint64
    lib.orc_compute_g_val.restype = C.c_double
    lib.orc_mt_canonical.restype = C.c_double
    lib.orc_rc2_inner.restype = C.c_double
    lib.orc_rc2_outer.restype = C.c_double
    for name in ("orc_destroy", "orc_set_pos", "orc_get_pos", "orc_get_vel", "orc_set_config",
                 "orc_get_config", "orc_init_config_from_pos", "orc_set_s_bias",
                 "orc_get_s_bias", "orc_init_s", "orc_reset_clocks", "orc_reset_counts",
                 "orc_md_step", "orc_crankshaft", "orc_enable_trace", "orc_get_trace",
                 "orc_enable_samples", "orc_get_dist", "orc_get_config_int",
                 "orc_get_counts", "orc_get_event_counts", "orc_get_unique_config",
                 "orc_check_local_dist", "orc_check_nonlocal_dist", "orc_rodrigues",
                 "orc_config_int", "orc_last_moves_done"):
        fn = getattr(lib, name)
        fn.argtypes = None
    return lib


def load_ref():
    """The compiled UNMODIFIED reference (oracle/_ref/libhmc_ref.so) or None."""
    path = os.path.join(ORACLE_DIR, "_ref", "libhmc_ref.so")
    if not os.path.exists(path):
        return None
    lib = C.CDLL(path)
    lib.ref_create.restype = C.c_void_p
    lib.ref_create.argtypes = [C.POINTER(HmcParams)]
    lib.ref_get_config.restype = C.c_uint64
    lib.ref_compute_g_val.restype = C.c_double
    lib.ref_time_md.restype = C.c_double
    lib.ref_mt_next.restype = C.c_uint32
    return lib


class Sim:
    """Thin OO wrapper shared by the oracle ('orc') and the reference ('ref')."""

    def __init__(self, lib, prefix, params):
        self.lib, self.pfx, self.p = lib, prefix, params
        self.h = C.c_void_p(getattr(lib, prefix + "_create")(C.byref(params)))
        if not self.h:
            raise RuntimeError("create failed")
        self.nb = params.nbeads
        self.ns = params.nstates

    def _f(self, name):
        return getattr(self.lib, self.pfx + "_" + name)

    def close(self):
        if self.h:
            self._f("destroy")(self.h)
            self.h = None

    def set_pos(self, pos):
        pos = np.ascontiguousarray(pos, dtype=np.float64)
        self._f("set_pos")(self.h, _dp(pos))

    def get_pos(self):
        out = np.zeros((self.nb, 3))
        self._f("get_pos")(self.h, _dp(out))
        return out

    def get_vel(self):
        out = np.zeros((self.nb, 3))
        self._f("get_vel")(self.h, _dp(out))
        return out

    def set_config(self, c):
        self._f("set_config")(self.h, C.c_uint64(c))

    def get_config(self):
        return int(self._f("get_config")(self.h))

    def init_config_from_pos(self):
        self._f("init_config_from_pos")(self.h)

    def set_s_bias(self, sb):
        sb = np.ascontiguousarray(sb, dtype=np.float64)
        self._f("set_s_bias")(self.h, _dp(sb), C.c_int(len(sb)))

    def get_s_bias(self):
        out = np.zeros(self.ns)
        self._f("get_s_bias")(self.h, _dp(out), C.c_int(self.ns))
        return out

    def init_s(self):
        self._f("init_s")(self.h)

    def reset_clocks(self):
        self._f("reset_clocks")(self.h)

    def md_step(self, phase, step, normals):
        normals = np.ascontiguousarray(normals, dtype=np.float64)
        return int(self._f("md_step")(self.h, C.c_int(phase), C.c_uint(step), _dp(normals)))

    def enable_trace(self, cap):
        self._f("enable_trace")(self.h, C.c_uint(cap))
        self._cap = cap

    def get_trace(self):
        buf = np.zeros(self._cap, dtype=EVENT_DTYPE)
        n = self._f("get_trace")(self.h, buf.ctypes.data_as(C.c_void_p), C.c_uint(self._cap))
        return buf[:min(n, self._cap)], n

    def event_counts(self):
        ev = (C.c_uint64 * 6)()
        self._f("get_event_counts")(self.h, ev)
        return list(ev)
