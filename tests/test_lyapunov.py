"""Lyapunov batch over a delay-equation parameter grid (BASELINE configs[4]).  The reference has no
(N)DDE Lyapunov test ("parity unpinned" upstream): the CUDA path is checked
against the oracle on the same estimator (north_star: exponents within 1e-3), and the estimator
against the rightmost characteristic root."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import diffurch_b200 as d  # noqa: E402
from diffurch_b200 import lyapunov  # noqa: E402


def grid(na, nb, neutral):
    bs = np.linspace(-0.6, 0.6, nb) if neutral else np.linspace(-2.0, 1.0, nb)
    return [x.ravel() for x in np.meshgrid(np.linspace(-2.0, 0.5, na), bs)]


def test_estimator_recovers_the_growth_rate_of_an_oscillating_mode():
    t = np.tile(np.arange(0.0, 100.0001, 0.1), (3, 1))
    alpha, omega = np.array([[-1.3], [0.0], [0.7]]), np.array([[2.0], [0.0], [5.0]])
    x = np.exp(alpha * t) * np.cos(omega * t + 0.3)
    n = np.array([t.shape[1]] * 3)
    lam = lyapunov.leading_exponent(t, x[:, :, None], n, window=10, t_skip=20.0)
    assert np.max(np.abs(lam - alpha[:, 0])) < 2e-3
    # a member that stopped early has no estimate; a shorter record still has one
    n[1] = 1
    n[2] = 600
    lam = lyapunov.leading_exponent(t, x[:, :, None], n, window=10, t_skip=20.0)
    assert np.isnan(lam[1]) and abs(lam[2] - 0.7) < 5e-3 and abs(lam[0] + 1.3) < 2e-3


@pytest.mark.parametrize("neutral", [False, True])
def test_oracle_grid_matches_the_rightmost_characteristic_root(oracle, neutral):
    system = d.systems.NddeLinear if neutral else d.systems.DdeLinear
    a, b = grid(5, 5, neutral)
    solver, window = lyapunov.linear_delay_grid(system, a, b, t1=100.0, arith="strict")
    o = oracle.run_solver(solver, n_threads=8)
    assert np.all(o.status == 0)
    lam = lyapunov.leading_exponent(o.trace_t, o.trace_p, o.trace_n, window, t_skip=20.0)
    root = np.array([lyapunov.rightmost_root(system, x, y) for x, y in zip(a, b)])
    assert np.max(np.abs(lam - root)) < 1e-3, np.max(np.abs(lam - root))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gloo_worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as orc
    a, b = grid(5, 3, True)                       # 15 members: ragged 8 + 7
    lam = lyapunov.sharded_leading_exponents(d.systems.NddeLinear, a, b, rank, world,
                                             runner=lambda s: orc.run_solver(s), t1=40.0, arith="strict")
    if rank == 0:
        np.save(out_path, lam)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_exponents_equal_single_process(tmp_path, oracle):
    import torch.multiprocessing as mp
    out = str(tmp_path / "lam.npy")
    mp.spawn(_gloo_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    a, b = grid(5, 3, True)
    solver, window = lyapunov.linear_delay_grid(d.systems.NddeLinear, a, b, t1=40.0, arith="strict")
    o = oracle.run_solver(solver)
    one = lyapunov.leading_exponent(o.trace_t, o.trace_p, o.trace_n, window, t_skip=20.0)
    assert np.array_equal(np.load(out), one)


@pytest.mark.gpu
@pytest.mark.parametrize("neutral", [False, True])
def test_gpu_grid_exponents_match_the_oracle(oracle, neutral):
    system = d.systems.NddeLinear if neutral else d.systems.DdeLinear
    a, b = grid(13, 11, neutral)                  # 143 members: ragged last warp
    solver, window = lyapunov.linear_delay_grid(system, a, b, t1=100.0, arith="fast")
    g = solver.run()
    assert g.totals["kernel_kind"] == d.abi.KERNEL_DDE_FIXED and np.all(g.status == 0)
    lam = lyapunov.leading_exponent(g.trace_t, g.trace_p, g.trace_n, window, t_skip=20.0)
    so, _ = lyapunov.linear_delay_grid(system, a, b, t1=100.0, arith="strict")
    o = oracle.run_solver(so, n_threads=8)
    lam_o = lyapunov.leading_exponent(o.trace_t, o.trace_p, o.trace_n, window, t_skip=20.0)
    assert np.array_equal(g.trace_n, o.trace_n) and np.array_equal(g.trace_t, o.trace_t)
    assert np.max(np.abs(lam - lam_o)) < 1e-6, np.max(np.abs(lam - lam_o))   # north_star: within 1e-3
    root = np.array([lyapunov.rightmost_root(system, x, y) for x, y in zip(a[::7], b[::7])])
    assert np.max(np.abs(lam[::7] - root)) < 5e-3                            # estimator accuracy, not parity


@pytest.mark.gpu
def test_gpu_lorenz_qr_spectrum_helper(oracle):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import cases
    g = cases.lyap_lorenz(20.0, rho=np.linspace(26.0, 30.0, 40), arith="strict").run()
    o = oracle.run_solver(cases.lyap_lorenz(20.0, rho=np.linspace(26.0, 30.0, 40), arith="strict"), n_threads=8)
    assert np.array_equal(lyapunov.qr_spectrum(g, 20.0), lyapunov.qr_spectrum(o, 20.0))
