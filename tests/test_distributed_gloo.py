"""world_size-2 gloo test (CPU) of the multi-GPU host logic: trajectory-range
sharding + gather/reduce.  The compute on each rank is the CPU oracle (this is a
test: the product's compute path is CUDA-only), so what is checked is that the
sharded + gathered result equals the single-process result, ragged shards included."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import cases
    import diffurch_b200 as d
    from diffurch_b200 import distributed as dd
    from oracle import oracle as orc

    rho = 20.0 + 20.0 * np.arange(n) / (n - 1)
    prm = np.stack([np.full(n, 10.0), rho, np.full(n, 8.0 / 3.0)], axis=1)
    y0 = np.tile([1.0, 2.0, 20.0], (n, 1))

    def make(y0s, prms):
        return (d.Solver.new().initial(y0s).equation(d.systems.Lorenz, prms).interval(0.0, 1.0)
                .stepsize(1.0 / 250.0))

    got = dd.run_sharded(make, y0, prm, rank, world, runner=lambda s: orc.run_solver(s))
    lo, hi = dd.shard_bounds(n, rank, world)
    stat = dd.all_reduce_sum([float(hi - lo), float(got["p"][lo:hi, 2].sum())])
    if rank == 0:
        np.savez(out_path, p=got["p"], t=got["t"], steps=got["steps"], stat=stat)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_bounds_cover_and_are_contiguous():
    sys.path.insert(0, ROOT)
    from diffurch_b200.distributed import shard_bounds
    for n in (1, 7, 8, 1000, 1 << 20):
        for world in (1, 2, 3, 8):
            prev = 0
            for r in range(world):
                lo, hi = shard_bounds(n, r, world)
                assert lo == prev and hi >= lo
                prev = hi
            assert prev == n
            sizes = [shard_bounds(n, r, world)[1] - shard_bounds(n, r, world)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_sharded_run_equals_single_process(tmp_path, oracle):
    n = 37  # ragged: 19 + 18
    out = str(tmp_path / "gathered.npz")
    port = _free_port()
    mp.spawn(_worker, args=(2, port, n, out), nprocs=2, join=True)
    got = np.load(out)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import diffurch_b200 as d
    rho = 20.0 + 20.0 * np.arange(n) / (n - 1)
    prm = np.stack([np.full(n, 10.0), rho, np.full(n, 8.0 / 3.0)], axis=1)
    y0 = np.tile([1.0, 2.0, 20.0], (n, 1))
    ref = oracle.run_solver(d.Solver.new().initial(y0).equation(d.systems.Lorenz, prm)
                            .interval(0.0, 1.0).stepsize(1.0 / 250.0))
    assert np.array_equal(got["p"], ref.p) and np.array_equal(got["t"], ref.t)
    assert np.array_equal(got["steps"], ref.steps.astype(np.int64))
    assert got["stat"][0] == n and abs(got["stat"][1] - ref.p[:, 2].sum()) < 1e-9
