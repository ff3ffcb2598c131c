"""Two-GPU checks (skipped on a one-GPU box): the ensemble shards by trajectory range, each
rank integrates its shard on its own GPU, and the final states are gathered over NCCL —
once through torch.distributed and once through the C-ABI `dfb_gather_final`
(ncclAllGather on the library's own packed buffers)."""
import ctypes as C
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:
        return 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import diffurch_b200 as d
    from diffurch_b200 import distributed as dd

    rho = 20.0 + 20.0 * np.arange(n) / (n - 1)
    prm = np.stack([np.full(n, 10.0), rho, np.full(n, 8.0 / 3.0)], axis=1)
    y0 = np.tile([1.0, 2.0, 20.0], (n, 1))

    def make(y0s, prms):
        return (d.Solver.new(device=rank).initial(y0s).equation(d.systems.Lorenz, prms)
                .interval(0.0, 2.0).stepsize(1.0 / 250.0))

    # (1) torch.distributed gather of ragged shards
    got = dd.run_sharded(make, y0, prm, rank, world, device=f"cuda:{rank}")

    # (2) C-ABI ncclAllGather (equal shards): raw ncclComm_t from torch's bundled NCCL
    m = n // world
    lo = rank * m
    ens = make(y0[lo:lo + m], prm[lo:lo + m]).build()
    ens.run()
    comm = dd.NcclComm(rank, world)
    t_all = np.empty(m * world)
    y_all = np.empty((m * world, 3))
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    rc = d.lib.dfb_gather_final(ens._h, comm.handle, world, dp(t_all), dp(y_all), None)
    assert rc == 0, d.lib.dfb_last_error()
    if rank == 0:
        np.savez(out_path, p=got["p"], t=got["t"], y_all=y_all, t_all=t_all)
    dist.barrier()
    comm.destroy()
    dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.skipif(_n_gpus() < 2, reason="needs 2 GPUs")
def test_two_gpu_sharded_run_and_nccl_gather(tmp_path):
    import torch.multiprocessing as mp
    import diffurch_b200 as d
    n = 2050  # ragged for the torch path: 1025 + 1025; the C-ABI leg uses 2 x 1025
    out = str(tmp_path / "g.npz")
    mp.spawn(_worker, args=(2, _free_port(), n, out), nprocs=2, join=True)
    got = np.load(out)
    rho = 20.0 + 20.0 * np.arange(n) / (n - 1)
    prm = np.stack([np.full(n, 10.0), rho, np.full(n, 8.0 / 3.0)], axis=1)
    one = (d.Solver.new().initial(np.tile([1., 2., 20.], (n, 1))).equation(d.systems.Lorenz, prm)
           .interval(0.0, 2.0).stepsize(1.0 / 250.0).run())
    assert np.array_equal(got["p"], one.p) and np.array_equal(got["t"], one.t)
    assert np.array_equal(got["y_all"], one.p) and np.array_equal(got["t_all"], one.t)


@pytest.mark.gpu
def test_one_rank_nccl_gather_runs_the_products_allgather(tmp_path):
    # The same worker with a ONE-rank communicator: `dfb_gather_final` still packs (t, y), calls
    # ncclAllGather on the library's own buffers and unpacks — so the product's NCCL path is exercised on a
    # one-GPU box too (the two-rank tests below need `gpurun --gpus 2`).
    import torch.multiprocessing as mp
    import diffurch_b200 as d
    n = 777
    out = str(tmp_path / "g1.npz")
    mp.spawn(_worker, args=(1, _free_port(), n, out), nprocs=1, join=True)
    got = np.load(out)
    rho = 20.0 + 20.0 * np.arange(n) / (n - 1)
    prm = np.stack([np.full(n, 10.0), rho, np.full(n, 8.0 / 3.0)], axis=1)
    one = (d.Solver.new().initial(np.tile([1., 2., 20.], (n, 1))).equation(d.systems.Lorenz, prm)
           .interval(0.0, 2.0).stepsize(1.0 / 250.0).run())
    assert np.array_equal(got["y_all"], one.p) and np.array_equal(got["t_all"], one.t)
    assert np.array_equal(got["p"], one.p)


def _lyap_worker(rank, world, port, n, tmax, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import cases
    import diffurch_b200 as d
    from diffurch_b200 import distributed as dd

    rho = np.linspace(20.0, 40.0, n)
    lo, hi = dd.shard_bounds(n, rank, world)
    ens = cases.lyap_lorenz(tmax, rho=rho[lo:hi], arith="strict")
    ens._device = rank
    ens = ens.build()
    ens.run()
    comm = dd.NcclComm(rank, world)
    m = hi - lo
    t_all, y_all, aux_all = np.empty(m * world), np.empty((m * world, 12)), np.empty((m * world, 3))
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    rc = d.lib.dfb_gather_final(ens._h, comm.handle, world, dp(t_all), dp(y_all), dp(aux_all))
    assert rc == 0, d.lib.dfb_last_error()
    if rank == 0:
        np.savez(out_path, spectra=aux_all / tmax, y_all=y_all)
    dist.barrier()
    comm.destroy()
    dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.skipif(_n_gpus() < 2, reason="needs 2 GPUs")
def test_two_gpu_lyapunov_spectra_gather(tmp_path):
    # BASELINE configs[4]: Lyapunov batch over a parameter grid, spectra gathered over NCCL
    import torch.multiprocessing as mp
    import cases
    n, tmax = 256, 20.0
    out = str(tmp_path / "l.npz")
    mp.spawn(_lyap_worker, args=(2, _free_port(), n, tmax, out), nprocs=2, join=True)
    got = np.load(out)
    one = cases.lyap_lorenz(tmax, rho=np.linspace(20.0, 40.0, n), arith="strict").run()
    assert one.totals["kernel_kind"] == 4
    assert np.array_equal(got["spectra"], one.aux / tmax)       # same kernel, same members: bit-identical
    assert np.array_equal(got["y_all"], one.p)
    # the three exponents sum to the trace of the Jacobian, -(sigma + 1 + beta)
    assert np.max(np.abs(got["spectra"].sum(axis=1) + (10.0 + 1.0 + 8.0 / 3.0))) < 1e-6


def _ndde_worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import diffurch_b200 as d
    from diffurch_b200 import lyapunov
    a, b = [x.ravel() for x in np.meshgrid(np.linspace(-2.0, 0.5, 33), np.linspace(-0.6, 0.6, 31))]
    lam = lyapunov.sharded_leading_exponents(d.systems.NddeLinear, a, b, rank, world, device="cuda", t1=60.0)
    if rank == 0:
        np.save(out_path, lam)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.skipif(_n_gpus() < 2, reason="needs 2 GPUs")
def test_two_gpu_ndde_lyapunov_grid_gather(tmp_path):
    # BASELINE configs[4] on a delay-equation grid: exponents of a ragged 1023-member grid, sharded 512 + 511
    import torch.multiprocessing as mp
    import diffurch_b200 as d
    from diffurch_b200 import lyapunov
    out = str(tmp_path / "lam.npy")
    mp.spawn(_ndde_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    a, b = [x.ravel() for x in np.meshgrid(np.linspace(-2.0, 0.5, 33), np.linspace(-0.6, 0.6, 31))]
    solver, window = lyapunov.linear_delay_grid(d.systems.NddeLinear, a, b, t1=60.0)
    st = solver.run()
    one = lyapunov.leading_exponent(st.trace_t, st.trace_p, st.trace_n, window, t_skip=20.0)
    assert np.array_equal(np.load(out), one)
