#!/usr/bin/env python3
"""bench.py — collision events/s of the HybridMC hot path on B200.

Workload (BASELINE.json configs[1]): examples/test.json — 20 beads, 3 bonds, hence
12 per-transition runs (tools/init_json.py) — as an ensemble of 4096 replicas per
transition on ONE GPU (49 152 replicas).  A "step" is one iteration of the
reference's run_trajectory (src/main.cc:171-247: nsteps=8 hybrid-MC MD intervals of
del_t=15 with fresh Maxwell-Boltzmann velocities + mc_moves=20 crankshaft moves)
for EVERY replica, in one persistent kernel launch.  N GPUs: every rank runs the
same per-GPU workload with its own Philox key (weak scaling) and the per-state
counts are all-reduced (NCCL, int64 sum) once per step — the path's only exchange.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm
  python bench.py --impl reference [...]                          the reference's CPU
                                                                  path on all host cores
Prints ONE JSON line (rank 0).
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "collision events/sec (whole box)"
UNIT = "events/s"
# Algorithmic FP64 operations per collision event:
# F = 49 P + 24 P_outer + 24 P_cell + 54 with the reference's measured prediction counts.
KFLOP_PER_EVENT = {8: 0.85e3, 20: 1.89e3, 46: 3.39e3, 50: 3.47e3}


def load_master(name):
    with open(os.path.join(ROOT, "tests", "golden", "configs", name + ".json")) as f:
        return json.load(f)


# ----------------------------------------------------------------- CPU arms --

def _cpu_worker(args):
    """One host core: time MD steps of one transition with the reference's own code
    (oracle/_ref) or with the C restatement (oracle/).  TEST INFRASTRUCTURE used as the
    measured baseline only."""
    kind, master, index, seconds, seed = args
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from hybridmc_b200.params import params_from_json, transition_jsons
    import oracle_lib as ol
    d = transition_jsons(master)[index][3]
    d["seeds"] = [seed, 1]
    p = params_from_json(d)
    orc = ol.load_oracle()
    g = ol.OrcMt()
    seeds = (C.c_uint32 * 2)(seed, 1)
    orc.orc_mt_seed_seq(C.byref(g), seeds, 2)
    pos = np.zeros((p.nbeads, 3))
    assert orc.orc_init_pos(C.byref(p), C.byref(g), pos.ctypes.data_as(C.c_void_p)) == 0
    vs = float(np.sqrt(p.temp / p.m))
    events, spent, step = 0, 0.0, 0
    if kind == "reference":
        ref = ol.load_ref()
        s = ol.Sim(ref, "ref", p)
        s.set_pos(pos)
        s.init_s()
        s.reset_clocks()
        ref.ref_mt_seed_seq(s.h, seeds, 2)
        chunk = 8
        while spent < seconds:
            c, x = C.c_uint64(), C.c_uint64()
            dt = ref.ref_time_md(s.h, C.c_uint(step), C.c_uint(chunk), C.byref(c), C.byref(x))
            spent += dt
            events += c.value
            step += chunk
    else:
        s = ol.Sim(orc, "orc", p)
        s.set_pos(pos)
        s.init_s()
        s.reset_clocks()
        normals = np.zeros((p.nbeads, 3))
        while spent < seconds:
            orc.orc_mt_normals(C.byref(g), C.c_double(vs), normals.ctypes.data_as(C.c_void_p),
                               3 * p.nbeads)
            t0 = time.perf_counter()
            s.md_step(0, step, normals)
            spent += time.perf_counter() - t0
            step += 1
        events = s.event_counts()[0]
    return events, spent, step


def cpu_events_per_sec(config, seconds, nprocs):
    """(events/s summed over nprocs single-threaded worker processes, kind, sample text).
    Workers are plain subprocesses of this script (--cpu-worker), one per host core."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol
    kind = "reference" if ol.load_ref() is not None else "port"
    ol.load_oracle()
    ntrans = 3  # layer-0 transitions: valid from an init_pos chain (no permanent bonds)
    t0 = time.perf_counter()
    procs = [subprocess.Popen([sys.executable, os.path.abspath(__file__), "--cpu-worker",
                               "%s,%s,%d,%f,%d" % (kind, config, k % ntrans, seconds, 1000 + k)],
                              stdout=subprocess.PIPE, text=True) for k in range(nprocs)]
    res = []
    for p in procs:
        out, _ = p.communicate(timeout=120 + 10 * seconds)
        if p.returncode != 0:
            raise RuntimeError("cpu worker failed")
        res.append(json.loads(out.strip().splitlines()[-1]))
    wall = time.perf_counter() - t0
    rate = sum(r[0] / r[1] for r in res)
    steps = sum(r[2] for r in res)
    sample = ("%d process(es) x %.0f s of initialize_system+run_step (EQ-phase MD steps, del_t=15) "
              "on test.json layer-0 transitions, %d steps total, wall %.1f s" % (nprocs, seconds, steps, wall))
    return rate, kind, sample


def cpu_worker_main(spec):
    kind, config, index, seconds, seed = spec.split(",")
    res = _cpu_worker((kind, load_master(config), int(index), float(seconds), int(seed)))
    print(json.dumps(res))


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ncores = os.cpu_count() or 1
    per_step = 4.0
    vals = []
    for _ in range(args.warmup):
        cpu_events_per_sec(args.config, 1.0, ncores)
    t0 = time.perf_counter()
    kind = sample = None
    for _ in range(args.steps):
        v, kind, sample = cpu_events_per_sec(args.config, per_step, ncores)
        vals.append(v)
    wall = time.perf_counter() - t0
    value = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * wall / max(args.steps, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "examples/test.json, 12 per-transition runs, reference CPU event loop "
                               "(priority queue), one process per host core"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": ncores, "kind": kind,
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------ GPU arm --

class _CudaArray:
    """__cuda_array_interface__ view of the engine's u64 accumulator block, so that
    torch.distributed can all-reduce it in place."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (ptr, False),
                                         "version": 2}


def clocks_sampler_start(path, dev):
    q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    try:
        f = open(path, "w")
        return subprocess.Popen(["nvidia-smi", "-i", str(dev), "--query-gpu=" + q, "--format=csv",
                                 "-lms", "200"], stdout=f, stderr=subprocess.DEVNULL), f
    except Exception:
        return None, None


def clocks_summary(path):
    sm, mx, reasons = [], [], set()
    try:
        with open(path) as f:
            rows = f.read().strip().splitlines()[1:]
        for r in rows:
            c = [x.strip() for x in r.split(",")]
            sm.append(float(c[1].split()[0]))
            mx.append(float(c[2].split()[0]))
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                  "sw_power_cap"), c[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
    except Exception:
        pass
    if not sm:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
    busy = sorted(sm)[len(sm) // 4:]  # drop idle samples at the edges
    return {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(mx), "reasons": sorted(reasons)}


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    from hybridmc_b200.campaign import harvest_start_structures, layers_of
    from hybridmc_b200.engine import PHASE_EQ, PHASE_MAIN, Engine, fp64_peak

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # NCCL (and anything else in the process) may write to stdout — its version banner does
    # even with NCCL_DEBUG_FILE set: park the real stdout and point fd 1 at stderr until the
    # one JSON line of the contract is printed
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    master = load_master(args.config)
    R = args.replicas
    # untimed set-up: one valid structure per bond state (layer walk), then the full
    # ensemble equilibrated for a few iterations
    structures = harvest_start_structures(master, device=local, replicas=128, seed=4242 + rank)
    trans = [t for _, ts in layers_of(master) for t in ts]
    params = [t[2] for t in trans]
    e = Engine(params, R, local)
    pos = np.stack([np.broadcast_to(structures[t[0]], (R,) + structures[t[0]].shape)
                    for t in trans]).reshape(e.R, e.nb, 3)
    e.set_positions(pos)
    e.init_config_from_positions()
    e.init_s()
    e.seed(2024 + rank, 7)
    e.reset_clocks()
    e.run(PHASE_EQ, 0, 2)
    e.init_config_from_positions()
    e.reset_clocks()
    e.reset_counts()
    e.sync()

    acc_ptr, acc_n = e.counts_device_ptr()
    acc_t = torch.as_tensor(_CudaArray(acc_ptr, acc_n), device=torch.device("cuda", local))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n_counts = e.T * e.nstates

    def one_step(it):
        e.reset_counts()
        e.run(PHASE_MAIN, it, 1)
        e.sync()
        ms = e.launch_stats()[1]
        if world > 1:
            # the single exchange step of the path: pooled per-state counts
            # (config_count[T][nstates], int64 sum => bit-exact, order-independent)
            dist.all_reduce(acc_t[:n_counts], op=dist.ReduceOp.SUM)
            torch.cuda.synchronize()
        return ms

    it = 0
    def flush_l2():
        flush.zero_()  # 256 MiB write > 126 MB L2
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        one_step(it)
        it += 1
        flush_l2()
    # ---- device-resident timed region ----
    ev0 = e.get_event_counts()
    l0 = e.launch_stats()[0]
    clk_path = os.path.join(ROOT, "gpurun_out", "clocks_rank%d.csv" % rank)
    os.makedirs(os.path.dirname(clk_path), exist_ok=True)
    proc, fh = clocks_sampler_start(clk_path, local)
    barrier()
    t0 = time.perf_counter()
    kernel_ms = []
    for _ in range(args.steps):
        kernel_ms.append(one_step(it))
        it += 1
        flush_l2()  # between timed iterations; not part of kernel_ms
    barrier()
    wall = time.perf_counter() - t0
    if proc:
        proc.terminate()
        fh.close()
    ev1 = e.get_event_counts()
    launches = e.launch_stats()[0] - l0
    local_events = float(ev1[0] - ev0[0])
    cells = float(ev1[1] - ev0[1])
    dev_ms = float(sum(kernel_ms))

    # ---- e2e: host-resident state, H2D + run + D2H every step ----
    host_pos = torch.empty((e.R, e.nb, 3), dtype=torch.float64).pin_memory().numpy()
    host_pos[...] = e.get_positions()
    host_cfg = e.get_config().copy()
    e.sync()
    ee0 = e.get_event_counts()
    barrier()
    te0 = time.perf_counter()
    for _ in range(args.steps):
        e.set_positions(host_pos)
        e.set_config(host_cfg)
        e.run(PHASE_MAIN, it, 1)
        it += 1
        host_pos[...] = e.get_positions()
        host_cfg = e.get_config()
        cc, formed, broken, err = e.get_counts()
    barrier()
    e2e_wall = time.perf_counter() - te0
    ee1 = e.get_event_counts()
    h2d = e.R * e.nb * 24 + e.R * 8
    d2h = e.R * e.nb * 24 + e.R * 8 + acc_n * 8 + e.R * 4

    e2e_events = float(ee1[0] - ee0[0])
    total_events = local_events
    if world > 1:
        # whole job: max time over ranks, events summed over ranks
        t = torch.tensor([dev_ms, wall, e2e_wall], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, wall, e2e_wall = [float(x) for x in t.tolist()]
        s = torch.tensor([total_events, e2e_events, cells], dtype=torch.float64, device="cuda")
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
        total_events, e2e_events, cells = [float(x) for x in s.tolist()]

    err_any = int(np.bitwise_or.reduce(e.get_counts()[3]))
    if rank == 0:
        value = total_events / (dev_ms * 1e-3)
        e2e_value = e2e_events / e2e_wall
        fma_tf, muladd_tf = fp64_peak(local)
        kflop = KFLOP_PER_EVENT.get(e.nb, 1.89e3)
        per_gpu_rate = value / world
        achieved = per_gpu_rate * kflop * 1e-12
        traffic = None
        prof = os.path.join(ROOT, "profiles", "r01_top_kernel.json")
        if os.path.exists(prof):
            try:
                traffic = json.load(open(prof)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": "examples/test.json: 12 per-transition runs x %d replicas per GPU, "
                                   "step = 1 run_trajectory iteration (8 MD intervals del_t=15 + 20 "
                                   "crankshaft moves) of every replica" % R,
                       "replicas_per_gpu": e.R, "nbeads": e.nb, "l2": "flushed between timed steps "
                       "(256 MiB write); state lives in shared memory",
                       "rng": "Philox4x32-10 per replica", "wall_s_timed_region": wall,
                       "cell_crossings_per_s": cells / (dev_ms * 1e-3),
                       "replica_error_flags": err_any},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches),
            "roofline": {"bound": "fp64", "achieved": achieved, "peak": muladd_tf,
                         "unit": "TFLOP/s", "frac": achieved / muladd_tf if muladd_tf else None,
                         "traffic": traffic,
                         "note": "peak = DMUL+DADD issue rate measured live by hmc_fp64_peak (the "
                                 "ceiling of a -fmad=false kernel; DFMA rate %.1f TFLOP/s); achieved = "
                                 "events/s/GPU x %.0f algorithmic FP64 ops per event" % (fma_tf, kflop)},
            "clocks": clocks_summary(clk_path),
        }
        if world == 1 and not args.no_cpu_baseline:
            v, kind, sample = cpu_events_per_sec(args.config, args.cpu_seconds, 1)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": kind,
                                    "sample": sample}
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    e.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="test")
    ap.add_argument("--replicas", type=int, default=4096, help="replicas per transition per GPU")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-worker", default=None, help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.cpu_worker:
        cpu_worker_main(args.cpu_worker)
        return
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
