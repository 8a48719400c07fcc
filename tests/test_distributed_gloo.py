"""world_size-2 gloo test of the multi-GPU host logic: shard -> per-rank solve -> all_gather must equal the
single-process result bitwise per instance.  The per-rank 'solve' here is the CPU ORACLE standing in for the GPU
(the sharding / gather code under test is the product code in mpc_benchmark_b200/distributed.py)."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, batch, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist

    import oracle_lib
    from mpc_benchmark_b200 import distributed, problems

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    prob = problems.cent_standing_problem(batch=batch, T=12)
    rng = np.random.default_rng(0)
    prob["x0"] = prob["x0"] + rng.normal(size=prob["x0"].shape) * 0.02  # distinct instances, same on every rank
    prob["xs"] = np.repeat(prob["x0"][:, None, :], 13, axis=1)
    lo, hi = distributed.shard_range(batch, rank, world)
    r = oracle_lib.solve(problems.sub_problem(prob, lo, hi), max_iters=5)
    full = distributed.gather_arrays(dict(xs=r["xs"], us=r["us"], iters=np.array([i.num_iters for i in r["info"]])), batch)
    if rank == 0:
        q.put(full)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("batch", [6, 5])
def test_sharded_equals_single_process(oracle, batch):
    from mpc_benchmark_b200 import distributed, problems

    assert [distributed.shard_range(5, r, 2) for r in range(2)] == [(0, 3), (3, 5)]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + batch) % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, batch, q)) for r in range(2)]
    for p in procs:
        p.start()
    full = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    prob = problems.cent_standing_problem(batch=batch, T=12)
    rng = np.random.default_rng(0)
    prob["x0"] = prob["x0"] + rng.normal(size=prob["x0"].shape) * 0.02
    prob["xs"] = np.repeat(prob["x0"][:, None, :], 13, axis=1)
    ref = oracle.solve(prob, max_iters=5)
    assert np.array_equal(full["xs"], ref["xs"]) and np.array_equal(full["us"], ref["us"])
    assert list(full["iters"]) == [i.num_iters for i in ref["info"]]
