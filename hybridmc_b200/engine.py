"""ctypes mirror of the C-ABI (include/hybridmc_b200.h).

`Engine` wraps one `hmc_handle`.  Method names follow the reference's functions
(run_trajectory_eq / wang_landau / run_trajectory -> `run(PHASE_*)`,
compute_entropy, g_test ...).  Nothing here computes trajectories: every call is
forwarded to libhybridmc_b200.so, and a missing library or GPU raises.
"""
import ctypes as C
import os

import numpy as np

from .params import HmcParams

PHASE_EQ, PHASE_WL, PHASE_MAIN = 0, 1, 2

ERR_LOCAL_OVERLAP, ERR_NONLOCAL_OVERLAP, ERR_ENERGY = 1, 2, 4
ERR_NAN, ERR_TAPE_UNDERRUN, ERR_SAMPLE_OVERFLOW = 8, 16, 32
ERR_LIST_OVERFLOW, ERR_CRANK_STUCK = 64, 128

EVENT_DTYPE = np.dtype([("t", "<f8"), ("type", "<u4"), ("i", "<u4"), ("j", "<u4"),
                        ("pad", "<u4")])

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhybridmc_b200.so")

# every symbol include/hybridmc_b200.h declares
EXPORTS = [
    "hmc_timer_start", "hmc_timer_stop", "hmc_resident_replicas", "hmc_debug_bounds_violations",
    "hmc_create", "hmc_destroy", "hmc_last_error", "hmc_abi_version",
    "hmc_set_positions", "hmc_get_positions", "hmc_get_velocities", "hmc_set_config",
    "hmc_get_config", "hmc_init_config_from_positions", "hmc_set_s_bias", "hmc_get_s_bias",
    "hmc_reset_clocks", "hmc_reset_counts", "hmc_seed", "hmc_run", "hmc_sync",
    "hmc_md_steps_injected", "hmc_crankshaft_injected", "hmc_get_counts",
    "hmc_counts_device_ptr", "hmc_get_event_counts", "hmc_enable_samples", "hmc_get_dist",
    "hmc_get_config_int", "hmc_enable_dist_hist", "hmc_get_dist_hist",
    "hmc_get_unique_config", "hmc_enable_trace", "hmc_get_trace", "hmc_compute_entropy",
    "hmc_compute_g_val", "hmc_g_test", "hmc_wl_outer_iterations", "hmc_get_launch_stats", "hmc_init_pos_chain", "hmc_fp64_peak",
    "hmc_get_moves_done", "hmc_init_pos_engine", "hmc_run_program_exact",
    "hmc_run_program_exact_sink", "hmc_get_schedule_state", "hmc_set_schedule_state",
    "hmc_set_s_bias_per_replica",
    "hmc_dist_hist_device_ptr",
    "hmc_comm_unique_id", "hmc_comm_create", "hmc_comm_create_from_file", "hmc_comm_destroy",
    "hmc_comm_wrap_nccl",
    "hmc_comm_rank", "hmc_comm_world", "hmc_comm_nccl_version", "hmc_comm_collectives",
    "hmc_allreduce_counts", "hmc_allreduce_totals", "hmc_allreduce_hist",
    "hmc_comm_allreduce_host_u64", "hmc_comm_bcast_host", "hmc_comm_allreduce_device_u64",
    "hmc_set_replica_ids", "hmc_enable_replica_counts", "hmc_get_replica_counts",
    "hmc_set_replica_skip", "hmc_run_program_ensemble", "hmc_set_list_capacity",
    "hmc_get_list_capacity",
    "hmc_h5_create", "hmc_h5_dataset", "hmc_h5_attribute", "hmc_h5_write", "hmc_h5_destroy",
]

_lib = None


class EngineError(RuntimeError):
    pass


class ScheduleState(C.Structure):
    _fields_ = [("t_now", C.c_double), ("md_draws", C.c_uint64), ("mc_draws", C.c_uint64),
                ("seed0", C.c_uint32), ("seed1", C.c_uint32)]


def load_library(path=None):
    """dlopen the CUDA library; fails loudly if it has not been built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or os.environ.get("HMC_LIB") or LIB_PATH  # HMC_LIB: experimental variant
    if not os.path.exists(path):
        raise EngineError(
            "%s is missing: build it with `python -m hybridmc_b200.build` "
            "(nvcc, sm_100a). There is no CPU fallback." % path)
    lib = C.CDLL(path)
    lib.hmc_last_error.restype = C.c_char_p
    lib.hmc_last_error.argtypes = [C.c_void_p]
    lib.hmc_create.argtypes = [C.POINTER(HmcParams), C.c_int, C.c_int, C.c_int,
                               C.POINTER(C.c_void_p)]
    lib.hmc_compute_g_val.restype = C.c_double
    lib.hmc_compute_g_val.argtypes = [C.c_void_p, C.c_int]
    lib.hmc_wl_outer_iterations.restype = C.c_uint
    lib.hmc_wl_outer_iterations.argtypes = [C.POINTER(HmcParams)]
    lib.hmc_compute_entropy.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_double]
    lib.hmc_g_test.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_void_p,
                               C.c_void_p]
    lib.hmc_enable_dist_hist.argtypes = [C.c_void_p, C.c_uint, C.c_double]
    lib.hmc_debug_bounds_violations.restype = C.c_longlong
    lib.hmc_comm_collectives.restype = C.c_uint64
    lib.hmc_comm_collectives.argtypes = [C.c_void_p]
    lib.hmc_comm_create.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    lib.hmc_comm_create_from_file.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int,
                                              C.POINTER(C.c_void_p)]
    for name in ("hmc_comm_destroy", "hmc_comm_rank", "hmc_comm_world"):
        getattr(lib, name).argtypes = [C.c_void_p]
    lib.hmc_allreduce_counts.argtypes = [C.c_void_p, C.c_void_p]
    lib.hmc_allreduce_totals.argtypes = [C.c_void_p, C.c_void_p]
    lib.hmc_allreduce_hist.argtypes = [C.c_void_p, C.c_void_p]
    lib.hmc_comm_allreduce_host_u64.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int]
    lib.hmc_comm_bcast_host.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int]
    lib.hmc_set_replica_ids.argtypes = [C.c_void_p, C.c_uint, C.c_uint]
    lib.hmc_enable_replica_counts.argtypes = [C.c_void_p, C.c_int]
    lib.hmc_get_replica_counts.argtypes = [C.c_void_p, C.c_void_p]
    lib.hmc_set_replica_skip.argtypes = [C.c_void_p, C.c_void_p]
    for name in ("hmc_destroy", "hmc_set_positions", "hmc_get_positions", "hmc_get_velocities",
                 "hmc_set_config", "hmc_get_config", "hmc_init_config_from_positions",
                 "hmc_reset_clocks", "hmc_reset_counts", "hmc_sync", "hmc_get_event_counts",
                 "hmc_get_dist", "hmc_get_config_int", "hmc_get_dist_hist",
                 "hmc_get_unique_config", "hmc_get_launch_stats", "hmc_counts_device_ptr"):
        getattr(lib, name).argtypes = None
    _lib = lib
    return lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class Engine:
    """One ensemble on one GPU: T transitions x R replicas each."""

    def __init__(self, transitions, replicas_per_transition=1, device=0):
        if isinstance(transitions, HmcParams):
            transitions = [transitions]
        self.lib = load_library()
        self.T = len(transitions)
        self.Rper = int(replicas_per_transition)
        self.R = self.T * self.Rper
        self.params = list(transitions)
        self.p = self.params[0]
        self.nb = self.p.nbeads
        self.nstates = self.p.nstates
        arr = (HmcParams * self.T)(*transitions)
        h = C.c_void_p()
        rc = self.lib.hmc_create(arr, self.T, self.Rper, int(device), C.byref(h))
        if rc != 0:
            raise EngineError(self.lib.hmc_last_error(None).decode())
        self.h = h
        self.n_nonlocal = max(p.n_nonlocal for p in transitions)
        self._trace_cap = 0
        self._dist_cap = 0
        self._cfg_cap = 0
        self._hist_bins = 0

    # -- multi-GPU / independent-program mode --
    def set_replica_ids(self, first, total_per_transition):
        """RNG identity of this handle's replicas inside a multi-rank ensemble."""
        self._ck(self.lib.hmc_set_replica_ids(self.h, int(first), int(total_per_transition)))

    def allreduce_counts(self, comm):
        """config_count summed over the ranks of `comm`, in place, on the engine's stream."""
        self._ck(self.lib.hmc_allreduce_counts(self.h, comm.c if comm is not None else None))

    def allreduce_totals(self, comm):
        self._ck(self.lib.hmc_allreduce_totals(self.h, comm.c if comm is not None else None))

    def allreduce_hist(self, comm):
        self._ck(self.lib.hmc_allreduce_hist(self.h, comm.c if comm is not None else None))

    def set_list_capacity(self, entries):
        """entries of the per-replica list of live nonlocal predictions (0: default)"""
        self._ck(self.lib.hmc_set_list_capacity(self.h, C.c_uint(int(entries))))

    def get_list_capacity(self):
        return int(self.lib.hmc_get_list_capacity(self.h))

    def enable_replica_counts(self, on=True):
        self._ck(self.lib.hmc_enable_replica_counts(self.h, 1 if on else 0))

    def get_replica_counts(self):
        out = np.zeros((self.R, self.nstates), dtype=np.uint64)
        self._ck(self.lib.hmc_get_replica_counts(self.h, _p(out)))
        return out

    def set_replica_skip(self, skip):
        if skip is None:
            self._ck(self.lib.hmc_set_replica_skip(self.h, None))
        else:
            s = np.ascontiguousarray(skip, dtype=np.uint8)
            assert s.shape == (self.R,)
            self._ck(self.lib.hmc_set_replica_skip(self.h, _p(s)))

    # -- plumbing --
    def _ck(self, rc):
        if rc != 0:
            raise EngineError(self.lib.hmc_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.hmc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- state --
    def set_positions(self, pos):
        pos = np.ascontiguousarray(pos, dtype=np.float64).reshape(self.R, self.nb, 3)
        self._ck(self.lib.hmc_set_positions(self.h, _p(pos)))

    def get_positions(self):
        out = np.empty((self.R, self.nb, 3))
        self._ck(self.lib.hmc_get_positions(self.h, _p(out)))
        return out

    def get_velocities(self):
        out = np.empty((self.R, self.nb, 3))
        self._ck(self.lib.hmc_get_velocities(self.h, _p(out)))
        return out

    def set_config(self, cfg):
        cfg = np.ascontiguousarray(np.broadcast_to(np.asarray(cfg, dtype=np.uint64), (self.R,)))
        self._ck(self.lib.hmc_set_config(self.h, _p(cfg)))

    def get_config(self):
        out = np.empty(self.R, dtype=np.uint64)
        self._ck(self.lib.hmc_get_config(self.h, _p(out)))
        return out

    def init_config_from_positions(self):
        self._ck(self.lib.hmc_init_config_from_positions(self.h))

    def set_s_bias(self, s_bias):
        sb = np.ascontiguousarray(
            np.broadcast_to(np.asarray(s_bias, dtype=np.float64), (self.T, self.nstates)))
        self._ck(self.lib.hmc_set_s_bias(self.h, _p(sb), self.nstates))

    def get_s_bias(self):
        out = np.empty((self.R, self.nstates))
        self._ck(self.lib.hmc_get_s_bias(self.h, _p(out), self.nstates))
        return out

    def init_s(self):
        """init_s, src/entropy.cc:14-20."""
        nb = self.p.n_transient
        sb = np.array([3.0 * (nb - bin(i).count("1")) for i in range(self.nstates)])
        self.set_s_bias(sb)
        return sb

    def reset_clocks(self):
        self._ck(self.lib.hmc_reset_clocks(self.h))

    def reset_counts(self):
        self._ck(self.lib.hmc_reset_counts(self.h))

    # -- production path --
    def seed(self, s0, s1=0):
        self._ck(self.lib.hmc_seed(self.h, C.c_uint32(s0), C.c_uint32(s1)))

    def run(self, phase, first_iter, n_iters):
        self._ck(self.lib.hmc_run(self.h, int(phase), C.c_uint(first_iter), C.c_uint(n_iters)))

    def sync(self):
        self._ck(self.lib.hmc_sync(self.h))

    def wl_outer_iterations(self):
        return int(self.lib.hmc_wl_outer_iterations(C.byref(self.p)))

    # -- reference-exact path --
    def md_steps_injected(self, phase, first_step, normals):
        normals = np.ascontiguousarray(normals, dtype=np.float64)
        nsteps = normals.size // (self.R * self.nb * 3)
        assert normals.size == self.R * nsteps * self.nb * 3
        self._ck(self.lib.hmc_md_steps_injected(self.h, int(phase), C.c_uint(first_step),
                                                C.c_uint(nsteps), _p(normals)))

    def crankshaft_injected(self, move_offset, n_moves, words, sin_tab, cos_tab):
        words = np.ascontiguousarray(words, dtype=np.uint32).reshape(self.R, -1)
        n = words.shape[1]
        sin_tab = np.ascontiguousarray(sin_tab, dtype=np.float64).reshape(self.R, n)
        cos_tab = np.ascontiguousarray(cos_tab, dtype=np.float64).reshape(self.R, n)
        consumed = np.zeros(self.R, dtype=np.uint32)
        self._ck(self.lib.hmc_crankshaft_injected(
            self.h, C.c_uint(move_offset), C.c_uint(n_moves), _p(words), _p(sin_tab),
            _p(cos_tab), C.c_uint(n), _p(consumed)))
        return consumed

    # -- results --
    def get_counts(self):
        cc = np.zeros((self.T, self.nstates), dtype=np.uint64)
        formed = np.zeros(self.T, dtype=np.uint64)
        broken = np.zeros(self.T, dtype=np.uint64)
        err = np.zeros(self.R, dtype=np.uint32)
        self._ck(self.lib.hmc_get_counts(self.h, _p(cc), self.nstates, _p(formed), _p(broken),
                                         _p(err)))
        return cc, formed, broken, err

    def counts_device_ptr(self):
        ptr, n = C.c_void_p(), C.c_uint64()
        self._ck(self.lib.hmc_counts_device_ptr(self.h, C.byref(ptr), C.byref(n)))
        return ptr.value, int(n.value)

    def get_event_counts(self):
        ev = np.zeros(4, dtype=np.uint64)
        self._ck(self.lib.hmc_get_event_counts(self.h, _p(ev)))
        return ev

    def enable_samples(self, dist_rows, config_int_cap=0):
        self._ck(self.lib.hmc_enable_samples(self.h, C.c_uint(dist_rows), C.c_uint(config_int_cap)))
        self._dist_cap, self._cfg_cap = dist_rows, config_int_cap

    def get_dist(self):
        stride = max(self.n_nonlocal, 1)
        out = np.zeros((self.R, self._dist_cap, stride))
        n = np.zeros(self.R, dtype=np.uint32)
        self._ck(self.lib.hmc_get_dist(self.h, _p(out), _p(n)))
        return out, n

    def get_config_int(self):
        out = np.zeros((self.R, self._cfg_cap), dtype=np.uint64)
        n = np.zeros(self.R, dtype=np.uint32)
        self._ck(self.lib.hmc_get_config_int(self.h, _p(out), _p(n)))
        return out, n

    def enable_dist_hist(self, nbins, rmax):
        self._ck(self.lib.hmc_enable_dist_hist(self.h, C.c_uint(nbins), C.c_double(rmax)))
        self._hist_bins = nbins

    def get_dist_hist(self):
        out = np.zeros((self.T, max(self.n_nonlocal, 1), 2, self._hist_bins), dtype=np.uint64)
        self._ck(self.lib.hmc_get_dist_hist(self.h, _p(out)))
        return out

    def dist_hist_device_ptr(self):
        """(device address, length in u64) of the histogram block, for an in-place all-reduce"""
        ptr, n = C.c_void_p(), C.c_uint64()
        self._ck(self.lib.hmc_dist_hist_device_ptr(self.h, C.byref(ptr), C.byref(n)))
        return ptr.value, int(n.value)

    def get_unique_config(self):
        pos = np.zeros((self.R, self.nb, 3))
        vel = np.zeros((self.R, self.nb, 3))
        found = np.zeros(self.R, dtype=np.uint8)
        self._ck(self.lib.hmc_get_unique_config(self.h, _p(pos), _p(vel), _p(found)))
        return pos, vel, found

    def enable_trace(self, replica, capacity):
        self._ck(self.lib.hmc_enable_trace(self.h, int(replica), C.c_uint(capacity)))
        self._trace_cap = capacity

    def get_trace(self):
        buf = np.zeros(max(self._trace_cap, 1), dtype=EVENT_DTYPE)
        n = C.c_uint()
        self._ck(self.lib.hmc_get_trace(self.h, _p(buf), C.c_uint(self._trace_cap), C.byref(n)))
        return buf[:min(n.value, self._trace_cap)], int(n.value)

    # -- snapshot / resume (ensemble form of src/snapshot.cc) --
    def snapshot(self):
        st = ScheduleState()
        self._ck(self.lib.hmc_get_schedule_state(self.h, C.byref(st)))
        return dict(pos=self.get_positions(), config=self.get_config(), s_bias=self.get_s_bias(),
                    schedule=(st.t_now, st.md_draws, st.mc_draws, st.seed0, st.seed1))

    def restore(self, snap):
        self.set_positions(snap["pos"])
        self.set_config(snap["config"])
        sb = np.ascontiguousarray(snap["s_bias"], dtype=np.float64).reshape(self.R, self.nstates)
        self._ck(self.lib.hmc_set_s_bias_per_replica(self.h, _p(sb), self.nstates))
        st = ScheduleState(*snap["schedule"])
        self._ck(self.lib.hmc_set_schedule_state(self.h, C.byref(st)))

    def launch_stats(self):
        n, ms = C.c_uint64(), C.c_float()
        self._ck(self.lib.hmc_get_launch_stats(self.h, C.byref(n), C.byref(ms)))
        return int(n.value), float(ms.value)

    def timer_start(self):
        self._ck(self.lib.hmc_timer_start(self.h))

    def timer_stop(self):
        """ms on the engine's stream since timer_start (kernels, all-reduce, copies, gaps)"""
        ms = C.c_float()
        self._ck(self.lib.hmc_timer_stop(self.h, C.byref(ms)))
        return float(ms.value)


# ---- entropy.cc mirror (host, O(nstates)) ---------------------------------------

def compute_entropy(config_count, s_bias, pos_scale, neg_scale):
    """compute_entropy, src/entropy.cc:36-69; returns the updated s_bias."""
    lib = load_library()
    cc = np.ascontiguousarray(config_count, dtype=np.uint64)
    sb = np.array(s_bias, dtype=np.float64)
    rc = lib.hmc_compute_entropy(_p(cc), _p(sb), len(cc), C.c_double(pos_scale),
                                 C.c_double(neg_scale))
    if rc != 0:
        raise EngineError("hmc_compute_entropy failed")
    return sb


def compute_g_val(config_count):
    lib = load_library()
    cc = np.ascontiguousarray(config_count, dtype=np.uint64)
    return float(lib.hmc_compute_g_val(_p(cc), len(cc)))


def g_test(config_count, sig_level):
    """g_test, src/entropy.cc:92-116 -> (accepted, g_val, g_crit, p_val)."""
    lib = load_library()
    cc = np.ascontiguousarray(config_count, dtype=np.uint64)
    g, crit, pv = C.c_double(), C.c_double(), C.c_double()
    ok = lib.hmc_g_test(_p(cc), len(cc), C.c_double(sig_level), C.byref(g), C.byref(crit),
                        C.byref(pv))
    return bool(ok), g.value, crit.value, pv.value


def init_pos_chain(params, seeds=None):
    """init_pos (src/hardspheres.cc:37-138) with std::mt19937(seed_seq(seeds))."""
    lib = load_library()
    if seeds is None:
        seeds = list(params.seeds)[:params.n_seeds]
    arr = (C.c_uint32 * len(seeds))(*[int(s) & 0xffffffff for s in seeds])
    pos = np.zeros((params.nbeads, 3))
    rc = lib.hmc_init_pos_chain(C.byref(params), arr, C.c_uint(len(seeds)), _p(pos))
    if rc != 0:
        raise EngineError("number of tries exceeded while placing bead in box")
    return pos


def resident_replicas(nbeads, device=0):
    """replicas of an nbeads-bead chain that run concurrently on one GPU"""
    n = load_library().hmc_resident_replicas(C.c_uint(int(nbeads)), int(device))
    if n < 0:
        raise EngineError("hmc_resident_replicas: no CUDA device %d" % device)
    return int(n)


def fp64_peak(device=0):
    """(DFMA TFLOP/s, DMUL+DADD TFLOP/s) measured on `device`."""
    lib = load_library()
    a, b = C.c_double(), C.c_double()
    rc = lib.hmc_fp64_peak(int(device), C.byref(a), C.byref(b))
    if rc != 0:
        raise EngineError("hmc_fp64_peak failed (%d)" % rc)
    return a.value, b.value


class ProgramResult(C.Structure):
    _fields_ = [("s_bias", C.c_double * 16), ("config_count", C.c_uint64 * 16),
                ("collisions", C.c_uint64), ("cell_crossings", C.c_uint64),
                ("config_final", C.c_uint64), ("gtest_iters", C.c_uint32),
                ("accepted", C.c_int), ("err_flags", C.c_uint32), ("pad", C.c_uint32),
                ("seconds", C.c_double), ("pos_final", C.c_double * (64 * 3)),
                ("message", C.c_char * 256)]


def run_program_exact(params, device=0, pos_in=None, max_gtest_iters=0, dist_rows_cap=0):
    """The reference program (src/main.cc:362-491) for one replica with the reference's
    RNG contract std::mt19937(seed_seq(seeds)), trajectories on the GPU.
    Returns (ProgramResult, dist[rows][n_nonlocal] or None)."""
    lib = load_library()
    res = ProgramResult()
    pin = None
    if pos_in is not None:
        pin = np.ascontiguousarray(pos_in, dtype=np.float64)
    dist = None
    nrows = C.c_uint(0)
    if dist_rows_cap:
        dist = np.zeros((dist_rows_cap, max(params.n_nonlocal, 1)))
    rc = lib.hmc_run_program_exact(C.byref(params), int(device), _p(pin) if pin is not None else None,
                                   C.c_uint(max_gtest_iters), C.c_uint(dist_rows_cap),
                                   _p(dist) if dist is not None else None, C.byref(nrows),
                                   C.byref(res))
    if rc != 0:
        raise EngineError(res.message.decode() or "hmc_run_program_exact failed (%d)" % rc)
    return res, (dist[:nrows.value] if dist is not None else None)


# ---- multi-GPU communicator (the library's own NCCL, include/hybridmc_b200.h) ----------

COMM_ID_BYTES = 128


class Comm:
    """hmc_comm: one rank of the all-reduce group.  `Comm.from_torch()` hands the id
    around with torch.distributed (any backend); `Comm.from_file()` uses a shared file."""

    def __init__(self, c, rank, world):
        self.lib, self.c, self.rank, self.world = load_library(), c, rank, world

    @staticmethod
    def unique_id():
        lib = load_library()
        buf = (C.c_uint8 * COMM_ID_BYTES)()
        if lib.hmc_comm_unique_id(buf) != 0:
            raise EngineError(lib.hmc_last_error(None).decode())
        return bytes(buf)

    @classmethod
    def create(cls, id_bytes, rank, world, device):
        lib = load_library()
        buf = (C.c_uint8 * COMM_ID_BYTES).from_buffer_copy(id_bytes)
        c = C.c_void_p()
        if lib.hmc_comm_create(buf, int(rank), int(world), int(device), C.byref(c)) != 0:
            raise EngineError(lib.hmc_last_error(None).decode())
        return cls(c, rank, world)

    @classmethod
    def from_file(cls, path, rank, world, device):
        lib = load_library()
        c = C.c_void_p()
        if lib.hmc_comm_create_from_file(path.encode(), int(rank), int(world), int(device),
                                         C.byref(c)) != 0:
            raise EngineError(lib.hmc_last_error(None).decode())
        return cls(c, rank, world)

    @classmethod
    def from_torch(cls, device):
        """Inside an initialised torch.distributed job (torchrun): rank 0 makes the id, the
        default group broadcasts its 128 bytes."""
        import torch.distributed as dist
        rank, world = dist.get_rank(), dist.get_world_size()
        box = [cls.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        return cls.create(box[0], rank, world, device)

    def allreduce_u64(self, arr, op="sum"):
        a = np.ascontiguousarray(arr, dtype=np.uint64)
        rc = self.lib.hmc_comm_allreduce_host_u64(self.c, _p(a), a.size, {"sum": 0, "max": 1, "min": 2}[op])
        if rc != 0:
            raise EngineError(self.lib.hmc_last_error(None).decode())
        return a

    def bcast(self, arr, root=0):
        a = np.ascontiguousarray(arr)
        if self.lib.hmc_comm_bcast_host(self.c, _p(a), a.nbytes, int(root)) != 0:
            raise EngineError(self.lib.hmc_last_error(None).decode())
        return a

    @property
    def collectives(self):
        return int(self.lib.hmc_comm_collectives(self.c))

    def close(self):
        if self.c:
            self.lib.hmc_comm_destroy(self.c)
            self.c = None


# ---- the ensemble program (production path, hmc_run_program_ensemble) -------------------

class EnsembleOptions(C.Structure):
    _fields_ = [("replicas", C.c_uint32), ("iters_per_round", C.c_uint32),
                ("eq_iters", C.c_uint32), ("max_rounds", C.c_uint32),
                ("wang_landau", C.c_uint32), ("independent", C.c_uint32),
                ("hist_bins", C.c_uint32), ("dist_rows", C.c_uint32),
                ("config_int_rows", C.c_uint32), ("seed0", C.c_uint32), ("seed1", C.c_uint32),
                ("wl_replicas", C.c_uint32), ("hist_rmax", C.c_double), ("comm", C.c_void_p),
                ("list_capacity", C.c_uint32), ("pad2", C.c_uint32), ("start_s_bias", C.c_void_p)]


class EnsembleResult(C.Structure):
    _fields_ = [("s_bias", C.c_double * 16), ("config_count", C.c_uint64 * 16),
                ("formed", C.c_uint64), ("broken", C.c_uint64), ("rounds", C.c_uint32),
                ("accepted", C.c_int), ("err_flags", C.c_uint32), ("have_unique", C.c_int),
                ("unique_pos", C.c_double * (64 * 3)), ("unique_vel", C.c_double * (64 * 3))]


class EnsembleTotals(C.Structure):
    _fields_ = [("collisions", C.c_uint64), ("cell_crossings", C.c_uint64),
                ("seconds", C.c_double), ("seconds_device", C.c_double),
                ("collectives", C.c_uint64), ("message", C.c_char * 256)]


_ON_ROUND = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_uint, C.POINTER(C.c_uint64),
                        C.POINTER(C.c_double), C.c_int, C.c_int, C.c_double, C.c_double, C.c_double)
_ON_DIST = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.POINTER(C.c_double), C.c_uint, C.c_uint)
_ON_CFG = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.POINTER(C.c_uint64), C.c_uint)
_ON_HIST = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.POINTER(C.c_uint64), C.c_uint, C.c_uint)
_ON_REPL = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_uint32),
                       C.POINTER(C.c_uint8), C.c_uint, C.c_int)


class EnsembleSink(C.Structure):
    _fields_ = [("user", C.c_void_p), ("on_round", _ON_ROUND), ("on_dist", _ON_DIST),
                ("on_config_int", _ON_CFG), ("on_hist", _ON_HIST), ("on_replicas", _ON_REPL)]


class EnsembleError(EngineError):
    """hmc_run_program_ensemble failed; `.results` holds what it produced (err_flags)."""

    def __init__(self, msg, results=None, code=0):
        super().__init__(msg)
        self.results, self.code = results, code


def run_program_ensemble(params_list, start_pos, replicas, device=0, seed=(1, 2),
                         iters_per_round=0, eq_iters=0, max_rounds=0, wl=True,
                         independent=False, hist_bins=0, hist_rmax=0.0, dist_rows=0,
                         config_int_rows=0, comm=None, verbose=False, keep_rounds=None, wl_replicas=0, start_s_bias=None, strict=True,
                         list_capacity=0):
    """hmc_run_program_ensemble: the reference's per-transition program (main.cc:411-491)
    for T transitions of one layer x `replicas` replicas each on THIS rank, pooled over the
    ranks of `comm`.  Returns a dict: s_bias[T][nstates], rounds[T], accepted[T], counts[T],
    counts_rounds (per transition: the config_count row of every round), err[T], formed,
    broken, unique_pos/unique_vel (list, None where no bonded structure was recorded),
    dist / config_int (lists of arrays, when requested; the rows of the last `keep_rounds`
    rounds, all by default), dist_hist[T][n_nonlocal][2][bins],
    replicas (independent mode: per transition (s_bias[R][ns], rounds[R], accepted[R])),
    events (collisions, cell crossings), seconds, seconds_device, collectives."""
    lib = load_library()
    T = len(params_list)
    p = params_list[0]
    ns, nb = p.nstates, p.nbeads
    arr = (HmcParams * T)(*params_list)
    pos = np.ascontiguousarray(np.stack([np.asarray(sp, dtype=np.float64) for sp in start_pos]))
    assert pos.shape == (T, nb, 3)
    opt = EnsembleOptions(replicas=int(replicas), iters_per_round=int(iters_per_round),
                          eq_iters=int(eq_iters), max_rounds=int(max_rounds),
                          wang_landau=1 if wl else 0, independent=1 if independent else 0,
                          hist_bins=int(hist_bins), dist_rows=int(dist_rows),
                          config_int_rows=int(config_int_rows), seed0=int(seed[0]) & 0xffffffff,
                          seed1=int(seed[1]) & 0xffffffff, hist_rmax=float(hist_rmax),
                          wl_replicas=int(wl_replicas), list_capacity=int(list_capacity),
                          comm=comm.c if comm is not None else None)
    if start_s_bias is not None:
        sb0 = np.zeros((T, 16))
        sb0[:, :ns] = np.asarray(start_s_bias, dtype=np.float64).reshape(T, ns)
        opt.start_s_bias = sb0.ctypes.data
    rounds_log = [[] for _ in range(T)]
    dist = [[] for _ in range(T)]
    cfg = [[] for _ in range(T)]
    hist = [None] * T
    repl = [None] * T

    def on_round(_u, t, rnd, cc, sb, n, ok, g, crit, pv):
        rounds_log[t].append(np.array([cc[i] for i in range(n)], dtype=np.uint64))
        if verbose:
            print("round", rnd, "transition", t, "s_bias", [sb[i] for i in range(n)], "counts",
                  [cc[i] for i in range(n)], "ACCEPTED" if ok else "REJECTED", g, crit)

    def on_dist(_u, t, rows, n, nbonds):
        dist[t].append(np.ctypeslib.as_array(rows, shape=(n, nbonds)).copy())
        if keep_rounds and len(dist[t]) > keep_rounds:
            del dist[t][0]

    def on_cfg(_u, t, vals, n):
        cfg[t].append(np.ctypeslib.as_array(vals, shape=(n,)).copy())
        if keep_rounds and len(cfg[t]) > keep_rounds:
            del cfg[t][0]

    def on_hist(_u, t, h, nl, nbins):
        hist[t] = np.ctypeslib.as_array(h, shape=(nl, 2, nbins)).copy()

    def on_repl(_u, t, sb, rnds, acc, R, n):
        repl[t] = (np.ctypeslib.as_array(sb, shape=(R, n)).copy(),
                   np.ctypeslib.as_array(rnds, shape=(R,)).copy(),
                   np.ctypeslib.as_array(acc, shape=(R,)).astype(bool))

    sink = EnsembleSink(None, _ON_ROUND(on_round),
                        _ON_DIST(on_dist) if dist_rows else _ON_DIST(),
                        _ON_CFG(on_cfg) if config_int_rows else _ON_CFG(),
                        _ON_HIST(on_hist) if hist_bins else _ON_HIST(),
                        _ON_REPL(on_repl) if independent else _ON_REPL())
    res = (EnsembleResult * T)()
    tot = EnsembleTotals()
    rc = lib.hmc_run_program_ensemble(arr, T, int(device), _p(pos), C.byref(opt), C.byref(sink),
                                      res, C.byref(tot))
    out = dict(
        s_bias=np.array([[r.s_bias[i] for i in range(ns)] for r in res]),
        counts=np.array([[r.config_count[i] for i in range(ns)] for r in res], dtype=np.uint64),
        rounds=np.array([r.rounds for r in res]), accepted=np.array([bool(r.accepted) for r in res]),
        err=np.array([r.err_flags for r in res], dtype=np.uint32),
        formed=np.array([r.formed for r in res], dtype=np.uint64),
        broken=np.array([r.broken for r in res], dtype=np.uint64),
        unique_pos=[np.array(r.unique_pos[:nb * 3]).reshape(nb, 3) if r.have_unique else None for r in res],
        unique_vel=[np.array(r.unique_vel[:nb * 3]).reshape(nb, 3) if r.have_unique else None for r in res],
        counts_rounds=[np.array(x, dtype=np.uint64).reshape(-1, ns) for x in rounds_log],
        events=np.array([tot.collisions, tot.cell_crossings], dtype=np.uint64),
        seconds=tot.seconds, seconds_device=tot.seconds_device, collectives=int(tot.collectives))
    if dist_rows:
        out["dist"] = [np.concatenate(d) if d else np.zeros((0, params_list[t].n_nonlocal))
                       for t, d in enumerate(dist)]
    if config_int_rows:
        out["config_int"] = [np.concatenate(c) if c else np.zeros(0, np.uint64) for c in cfg]
    if hist_bins:
        out["dist_hist"] = np.stack([h if h is not None else
                                     np.zeros((max(q.n_nonlocal for q in params_list), 2, hist_bins), np.uint64)
                                     for h in hist])
    if independent:
        out["replicas"] = repl
    # transitions stopped by a replica error flag (the reference's process would have died
    # with an exception); the other transitions of the batch are complete
    out["failed"] = (out["err"] & ~np.uint32(ERR_SAMPLE_OVERFLOW)) != 0
    if rc != 0 and (strict or rc != -8):
        raise EnsembleError(tot.message.decode() or "hmc_run_program_ensemble failed (%d)" % rc,
                            results=out, code=rc)
    return out

