"""CPU, world_size 2, gloo: the multi-GPU plumbing of the path — sharding of the
independent units and the single all-reduce of per-state counts per G-test round —
gives every rank the same pooled counts and the same s_bias as a single process."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from common import load_config
    from hybridmc_b200.build import build_library
    from hybridmc_b200.campaign import pooled_round, shard
    from hybridmc_b200.params import params_from_json
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    build_library()
    p = params_from_json(load_config("8bead_1bond"))
    rng = np.random.default_rng(1234)
    all_counts = rng.integers(0, 400, size=(8, 3, 2)).astype(np.uint64)  # 8 replica blocks, 3 transitions
    mine = shard(8, rank, world)
    local = all_counts[mine].sum(axis=0)
    s0 = np.tile(np.array([3.0, 0.0]), (3, 1))
    s1, acc, pooled = pooled_round(local, s0, p)
    q.put((rank, mine, s1.tolist(), acc.tolist(), pooled.tolist()))
    dist.destroy_process_group()


def test_two_ranks_pool_counts_and_agree():
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from common import load_config
    from hybridmc_b200.campaign import pooled_round
    from hybridmc_b200.params import params_from_json
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    assert res[0][1] == [0, 2, 4, 6] and res[1][1] == [1, 3, 5, 7]  # disjoint cover
    assert res[0][2:] == res[1][2:]  # both ranks: same s_bias, verdicts, pooled counts
    # single-process reference
    rng = np.random.default_rng(1234)
    all_counts = rng.integers(0, 400, size=(8, 3, 2)).astype(np.uint64)
    p = params_from_json(load_config("8bead_1bond"))
    s1, acc, pooled = pooled_round(all_counts.sum(axis=0), np.tile(np.array([3.0, 0.0]), (3, 1)), p)
    assert res[0][2] == s1.tolist() and res[0][4] == pooled.tolist()
