#!/usr/bin/env python3
"""bench.py — collision events/s of the HybridMC hot path on B200.

Workload (BASELINE.json configs[1]): examples/test.json — 20 beads, 3 bonds, hence
12 per-transition runs (tools/init_json.py) — as an ensemble of 4096 replicas per
transition on ONE GPU (49 152 replicas), every transition started from the committed
structure of its bond state (tests/golden/start_structures_test.npz, made on the CPU by
the oracle).  A "step" is one round of the production loop at one iteration per round:
one iteration of the reference's run_trajectory (src/main.cc:171-247: nsteps=8 hybrid-MC
MD intervals of del_t=15 with fresh Maxwell-Boltzmann velocities + mc_moves=20 crankshaft
moves + sampling) for EVERY replica in one persistent kernel launch, then the all-reduce
of the per-state counts over the ranks (the library's NCCL communicator, on the engine's
stream), compute_entropy on the pooled counts and the new s_bias back to the device
(main.cc:471-491).  N GPUs: the replicas of every transition span the GPUs — 4096 per
transition on each rank (weak scaling), 4096 x N pooled per transition.  `value` is timed
with CUDA events on the engine's stream around the WHOLE step (kernel + all-reduce +
count read-back + s_bias write), max over ranks; only the L2 flush between steps is outside.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm
  python bench.py --impl reference [...]                          the reference's own
                                                                  run_trajectory on all host cores
Prints ONE JSON line (rank 0).
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "collision events/sec (whole box)"
UNIT = "events/s"
# Algorithmic FP64 operations per collision event:
# F = 49 P + 24 P_outer + 24 P_cell + 54 with the reference's measured prediction counts.
KFLOP_PER_EVENT = {8: 0.85e3, 20: 1.89e3, 46: 3.39e3, 50: 3.47e3}
EQ_ITERS = 2  # untimed equilibration iterations before the timed region, both arms


def load_master(name):
    with open(os.path.join(ROOT, "tests", "golden", "configs", name + ".json")) as f:
        return json.load(f)


def fmt(bits):
    return "".join("1" if b else "0" for b in bits)


def workload(config):
    """[(json dict, start structure)] of the bench: every transition of test.json from the
    committed start structures; for the other configs the layer-0 transitions from the
    init_pos chain (no fixture needed: no permanent bonds)."""
    from hybridmc_b200.params import params_from_json, transition_jsons
    master = load_master(config)
    trans = transition_jsons(master)
    fx = os.path.join(ROOT, "tests", "golden", "start_structures_%s.npz" % config)
    if os.path.exists(fx):
        st = np.load(fx)
        return [(d, np.ascontiguousarray(st[fmt(bits_in)])) for _, bits_in, _, d in trans]
    from hybridmc_b200.engine import init_pos_chain
    layer0 = [t for t in trans if t[0] == 0]
    pos = init_pos_chain(params_from_json(layer0[0][3]))
    return [(d, pos) for _, _, _, d in layer0]


# ----------------------------------------------------------------- CPU arms --

def _cpu_worker(config, index, rounds, iters, seed):
    """One host core, one transition of the workload: `rounds` x `iters` calls of the
    reference's own run_trajectory (oracle/_ref, unmodified sources; wall time of those
    calls only), then the oracle's bit-identical replay of the same run for the event count
    the reference does not keep (tests/test_oracle_vs_ref.py pins the pair).  Without
    oracle/_ref the replay itself is what is timed (kind "port").  TEST INFRASTRUCTURE used
    as the measured baseline only."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from hybridmc_b200.params import params_from_json
    import oracle_lib as ol
    wl = workload(config)
    d, start = wl[index % len(wl)]
    p = params_from_json(dict(d, seeds=[seed, 1]))
    orc = ol.load_oracle()
    ref = ol.load_ref()
    pos = np.zeros((p.nbeads, 3))
    sec, cfg, coll, cells = C.c_double(), C.c_uint64(), C.c_uint64(), C.c_uint64()
    args = (C.byref(p), start.ctypes.data_as(C.c_void_p), C.c_uint(EQ_ITERS), C.c_uint(rounds),
            C.c_uint(iters))
    seconds = None
    if ref is not None:
        assert ref.ref_bench_main_phase(*args, C.byref(sec), pos.ctypes.data_as(C.c_void_p),
                                        C.byref(cfg)) == 0
        seconds, ref_pos = sec.value, pos.copy()
    assert orc.orc_bench_main_phase(*args, C.byref(sec), pos.ctypes.data_as(C.c_void_p),
                                    C.byref(cfg), C.byref(coll), C.byref(cells)) == 0
    if ref is not None:
        assert ref_pos.tobytes() == pos.tobytes(), "oracle replay left the reference trajectory"
    else:
        seconds = sec.value
    return coll.value, seconds, rounds * iters, "reference" if ref is not None else "port"


def cpu_events_per_sec(config, nprocs, rounds, iters):
    """(events/s summed over nprocs single-threaded worker processes, kind, sample text).
    Workers are plain subprocesses of this script (--cpu-worker), one per host core, dealt
    round-robin over the transitions of the workload."""
    t0 = time.perf_counter()
    procs = [subprocess.Popen([sys.executable, os.path.abspath(__file__), "--cpu-worker",
                               "%s,%d,%d,%d,%d" % (config, k, rounds, iters, 1000 + k)],
                              stdout=subprocess.PIPE, text=True) for k in range(nprocs)]
    res = []
    for p in procs:
        out, _ = p.communicate(timeout=1800)
        if p.returncode != 0:
            raise RuntimeError("cpu worker failed")
        res.append(json.loads(out.strip().splitlines()[-1]))
    wall = time.perf_counter() - t0
    rate = sum(r[0] / r[1] for r in res)
    kind = res[0][3]
    sample = ("%d process(es), one per host core, transitions of %s dealt round-robin, each %d rounds x %d "
              "iterations of the reference's run_trajectory (8 MD intervals del_t=15 + 20 crankshaft moves "
              "+ sampling, compute_entropy per round) after %d equilibration iterations; %.1f s of "
              "run_trajectory per process on average, %d events; wall %.1f s incl. the untimed "
              "event-count replay" % (nprocs, config, rounds, iters, EQ_ITERS,
                                      float(np.mean([r[1] for r in res])), sum(r[0] for r in res), wall))
    return rate, kind, sample


def cpu_worker_main(spec):
    config, index, rounds, iters, seed = spec.split(",")
    print(json.dumps(_cpu_worker(config, int(index), int(rounds), int(iters), int(seed))))


def workload_text(config, nb, per_gpu=None):
    n = len(workload(config))
    return ("examples/%s.json: %d per-transition runs%s, step = 1 run_trajectory iteration (8 MD "
            "intervals del_t=15 + 20 crankshaft moves + sampling) of every replica + pooled "
            "compute_entropy" % (config, n, " x %d replicas per GPU" % per_gpu if per_gpu else ""))


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ncores = os.cpu_count() or 1
    vals = []
    for _ in range(min(args.warmup, 1)):
        cpu_events_per_sec(args.config, ncores, 1, 2)
    t0 = time.perf_counter()
    kind = sample = None
    for _ in range(args.steps):
        v, kind, sample = cpu_events_per_sec(args.config, ncores, args.cpu_rounds, args.cpu_iters)
        vals.append(v)
    wall = time.perf_counter() - t0
    value = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * wall / max(args.steps, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_text(args.config, 0) + "; reference CPU event loop (priority "
                               "queue), one process per host core"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": ncores, "kind": kind,
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------ GPU arm --

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


def campaign_leg(config, comm, rank, world, local, replicas, log):
    """BASELINE metric 2: the whole lattice sweep (every transition's s_bias + first-passage
    histograms, layers in order with structure hand-off) through run_campaign on all ranks.
    Returns dict(transitions, seconds, transitions_per_hour, events)."""
    from hybridmc_b200.run_campaign import run_campaign
    master = load_master(config)
    box = [tempfile.mkdtemp(prefix="hmc_bench_campaign_") if rank == 0 else None]
    if world > 1:
        import torch.distributed as dist
        dist.broadcast_object_list(box, src=0)
    t0 = time.perf_counter()
    summary, wall, events = run_campaign(master, box[0], replicas=replicas * world, device=local,
                                         iters_per_round=4, max_rounds=12, hist_bins=2048,
                                         rank=rank, world=world, comm=comm, wl=True, log=log)
    wall = time.perf_counter() - t0
    return {"transitions": len(summary), "seconds": wall,
            "replicas_per_transition": replicas * world, "iters_per_round": 4,
            "transitions_per_hour": 3600.0 * len(summary) / wall,
            "replica_transitions_per_hour": 3600.0 * len(summary) * replicas * world / wall,
            "events": int(events), "events_per_s_whole_campaign": events / wall,
            "scaling": "weak: %d replicas per transition on every GPU, pooled over the ranks" % replicas,
            "outputs": "s_bias, config_count, dist, dist_hist (first-passage statistics), "
                       "unique_config per transition (HDF5)"}


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    from hybridmc_b200.engine import (PHASE_EQ, PHASE_MAIN, Comm, Engine, compute_entropy,
                                      fp64_peak)
    from hybridmc_b200.params import params_from_json

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
    comm = None
    if world > 1:
        dist.init_process_group("cpu:gloo,cuda:nccl", device_id=torch.device("cuda", local))
        comm = Comm.from_torch(local)  # the library's own NCCL communicator (the data path)

    wl = workload(args.config)
    params = [params_from_json(d) for d, _ in wl]
    R = args.replicas
    # replicas of every transition span the ranks: rank r owns ids [r*R, (r+1)*R) of R*world
    e = Engine(params, R, local)
    e.set_replica_ids(rank * R, R * world)
    pos = np.stack([np.broadcast_to(sp, (R,) + sp.shape) for _, sp in wl]).reshape(e.R, e.nb, 3)
    e.set_positions(pos)
    e.init_config_from_positions()
    s_bias = np.tile(e.init_s(), (e.T, 1))
    e.seed(2024, 7)
    e.reset_clocks()
    e.run(PHASE_EQ, 0, EQ_ITERS)
    e.init_config_from_positions()
    e.reset_clocks()
    e.reset_counts()
    e.sync()
    p0 = params[0]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step(it):
        """one round of the production loop; returns (ms of the whole step on the engine's
        stream, ms of the kernel alone)"""
        e.timer_start()
        e.reset_clocks()  # every round starts at step 0, as the G-test loop does (main.cc:453-460)
        e.reset_counts()
        e.run(PHASE_MAIN, 0, 1)
        e.allreduce_counts(comm)  # the path's one exchange, on the engine's stream
        cc, _, _, _ = e.get_counts()
        for t in range(e.T):
            s_bias[t] = compute_entropy(cc[t], s_bias[t], p0.pos_scale, p0.neg_scale)
        e.set_s_bias(s_bias)
        ms = e.timer_stop()
        return ms, e.launch_stats()[1]

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
    c0 = comm.collectives if comm else 0
    clk_path = os.path.join(ROOT, "gpurun_out", "clocks_rank%d.csv" % rank)
    os.makedirs(os.path.dirname(clk_path), exist_ok=True)
    proc, fh = clocks_sampler_start(clk_path, local)
    barrier()
    t0 = time.perf_counter()
    step_ms, kernel_ms = [], []
    for _ in range(args.steps):
        a, b = one_step(it)
        step_ms.append(a)
        kernel_ms.append(b)
        it += 1
        flush_l2()  # between timed iterations; not part of step_ms
    barrier()
    wall = time.perf_counter() - t0
    if proc:
        proc.terminate()
        fh.close()
    ev1 = e.get_event_counts()
    launches = e.launch_stats()[0] - l0
    collectives = (comm.collectives if comm else 0) - c0
    local_events = float(ev1[0] - ev0[0])
    cells = float(ev1[1] - ev0[1])
    dev_ms = float(sum(step_ms))
    ker_ms = float(sum(kernel_ms))

    # ---- e2e: host-resident state through the C-ABI, H2D + run + exchange + D2H every step ----
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
        e.set_s_bias(s_bias)
        e.reset_clocks()
        e.reset_counts()
        e.run(PHASE_MAIN, 0, 1)
        e.allreduce_counts(comm)
        it += 1
        host_pos[...] = e.get_positions()
        host_cfg = e.get_config()
        cc, formed, broken, err = e.get_counts()
        for t in range(e.T):
            s_bias[t] = compute_entropy(cc[t], s_bias[t], p0.pos_scale, p0.neg_scale)
    barrier()
    e2e_wall = time.perf_counter() - te0
    ee1 = e.get_event_counts()
    n_acc = e.T * e.nstates + 2 * e.T
    h2d = e.R * e.nb * 24 + e.R * 8 + e.T * e.nstates * 8
    d2h = e.R * e.nb * 24 + e.R * 8 + n_acc * 8 + e.R * 4

    e2e_events = float(ee1[0] - ee0[0])
    total_events = local_events
    if world > 1:
        # whole job: max time over ranks, events summed over ranks
        t = torch.tensor([dev_ms, wall, e2e_wall, ker_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, wall, e2e_wall, ker_ms = [float(x) for x in t.tolist()]
        s = torch.tensor([total_events, e2e_events, cells], dtype=torch.float64, device="cuda")
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
        total_events, e2e_events, cells = [float(x) for x in s.tolist()]

    err_any = int(np.bitwise_or.reduce(e.get_counts()[3]))
    e.close()
    campaign = None
    if not args.no_campaign:
        campaign = campaign_leg(args.config if args.config == "test" else "test", comm, rank, world,
                                local, args.campaign_replicas,
                                (lambda *a: print(*a, file=sys.stderr)) if rank == 0 else (lambda *a: None))
    if rank == 0:
        value = total_events / (dev_ms * 1e-3)
        e2e_value = e2e_events / e2e_wall
        fma_tf, muladd_tf = fp64_peak(local)
        kflop = KFLOP_PER_EVENT.get(params[0].nbeads, 1.89e3)
        # roofline of the dominant kernel: its own launch durations (CUDA events around the
        # launch on the engine's stream), one GPU
        kernel_rate = local_events / (float(sum(kernel_ms)) * 1e-3)
        achieved = kernel_rate * kflop * 1e-12
        traffic = None
        prof = os.path.join(ROOT, "profiles", "r02_top_kernel.json")
        if os.path.exists(prof):
            try:
                traffic = json.load(open(prof)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        # HBM side of a launch: every replica's snapshot (positions in and out, velocities out,
        # bond state, s_bias row, error word) crosses HBM once per launch, the per-state
        # counts / histograms are a handful of atomics per recorded sample.  Algorithmic bytes
        # per launch over the measured launch duration = the HBM bandwidth the path needs.
        nbeads = int(params[0].nbeads)
        state_bytes = R * len(params) * (3 * 24 * nbeads + 2 * 8 + 16 + 2 * 4)
        ker_s = float(sum(kernel_ms)) * 1e-3 / max(len(kernel_ms), 1)
        hbm_peak = None
        try:
            hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs")
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": workload_text(args.config, params[0].nbeads, R),
                       "replicas_per_gpu": R * len(params), "nbeads": int(params[0].nbeads),
                       "replicas_per_transition_pooled": R * world,
                       "l2": "flushed between timed steps (256 MiB write); state lives in shared memory",
                       "rng": "Philox4x32-10 per replica, replica ids global over the ranks",
                       "timing": "CUDA events on the engine's stream around kernel + NCCL all-reduce of "
                                 "config_count + count read-back + compute_entropy + s_bias write",
                       "exchange": "%d NCCL all-reduces (library communicator) in the timed region"
                                   % collectives,
                       "kernel_only_ms_per_step": ker_ms / args.steps,
                       "kernel_only_events_per_s_per_gpu": kernel_rate,
                       "wall_s_timed_region": wall,
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
                                 "kernel-only events/s of one GPU x %.0f algorithmic FP64 ops per event; "
                                 "MEASURED_PEAKS.json has no FP64 entry" % (fma_tf, kflop)},
            "hbm": {"snapshot_bytes_per_launch": state_bytes,
                    "snapshot_GBps_needed": state_bytes / ker_s * 1e-9,
                    "ncu_dram_bytes_per_launch": traffic,
                    "ncu_dram_GBps": (traffic / ker_s * 1e-9) if traffic else None,
                    "hbm_peak_GBps_measured": hbm_peak,
                    "note": "replica state lives in shared memory for the whole launch; HBM sees one "
                            "snapshot read + write per replica and launch plus count/histogram atomics "
                            "(measured peak copy bandwidth of this pool: MEASURED_PEAKS.json)"},
            "clocks": clocks_summary(clk_path),
        }
        if campaign is not None:
            line["campaign"] = campaign
            line["metric2"] = {"metric": "transitions (s_bias + MFPT statistics) per hour",
                               "value": campaign["transitions_per_hour"], "unit": "transitions/h",
                               "n_gpus": world, "config": "examples/test.json, all 12 transitions, "
                               "%d replicas per transition" % campaign["replicas_per_transition"]}
        if world == 1 and not args.no_cpu_baseline:
            ncores = os.cpu_count() or 1
            v, kind, sample = cpu_events_per_sec(args.config, ncores, args.cpu_rounds, args.cpu_iters)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": ncores, "kind": kind,
                                    "sample": sample}
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if comm is not None:
        comm.close()
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
    ap.add_argument("--cpu-rounds", type=int, default=4)
    ap.add_argument("--cpu-iters", type=int, default=10, help="run_trajectory iterations per round")
    ap.add_argument("--no-campaign", action="store_true", help="skip the metric-2 campaign leg")
    ap.add_argument("--campaign-replicas", type=int, default=1024,
                    help="replicas per transition and GPU in the campaign leg")
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
