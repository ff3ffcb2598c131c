#!/usr/bin/env python3
"""BASELINE configs[4] on N GPUs: Lyapunov spectra of a Lorenz rho grid (tests/lyapunov_exponents.rs:243-311:
variational system, h = 0.005, QR every 0.5) and leading exponents of an NDDE (a, b) grid, sharded by member
range, gathered with NCCL — `dfb_gather_final` (ncclAllGather on the library's packed buffers) for the spectra,
torch.distributed all_gather for the delay-equation exponents.  Run under torchrun:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_lyapunov_gather.py

Rank 0 prints one JSON object (kernel time = max over ranks of the CUDA-event time; gather = wall clock)."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29533")
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    import cases
    import diffurch_b200 as d
    from diffurch_b200 import distributed as dd
    from diffurch_b200 import lyapunov

    n_total = int(os.environ.get("DFB_LYAP_MEMBERS", 1 << 18))
    tmax = float(os.environ.get("DFB_LYAP_TMAX", 100.0))
    rho = np.linspace(20.0, 40.0, n_total)
    rho[n_total // 2] = 28.0  # the reference's own test point
    lo, hi = dd.shard_bounds(n_total, rank, world)
    assert (hi - lo) * world == n_total, "equal shards (ncclAllGather)"
    s = cases.lyap_lorenz(tmax, rho=rho[lo:hi], arith="fast")
    s._device = local
    ens = s.build()
    ens.run(); ens.synchronize()          # warm-up
    dist.barrier(); torch.cuda.synchronize()
    ens.run(); ens.synchronize()
    kernel_ms = ens.totals()["kernel_ms"]
    comm = dd.NcclComm(rank, world)
    m = hi - lo
    t_all, y_all, aux_all = np.empty(m * world), np.empty((m * world, 12)), np.empty((m * world, 3))
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    rc = d.lib.dfb_gather_final(ens._h, comm.handle, world, dp(t_all), dp(y_all), dp(aux_all))  # first call: NCCL connects
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    rc = d.lib.dfb_gather_final(ens._h, comm.handle, world, dp(t_all), dp(y_all), dp(aux_all))
    gather_s = time.perf_counter() - t0
    assert rc == 0, d.lib.dfb_last_error()
    spectra = aux_all / tmax
    km = torch.tensor([kernel_ms, gather_s * 1e3], dtype=torch.float64, device="cuda")
    dist.all_reduce(km, op=dist.ReduceOp.MAX)
    # delay-equation grid: leading exponents, sharded + all-gathered (diffurch_b200/lyapunov.py)
    a, b = [x.ravel() for x in np.meshgrid(np.linspace(-2.0, 0.5, 256), np.linspace(-0.6, 0.6, 256))]
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    lam = lyapunov.sharded_leading_exponents(d.systems.NddeLinear, a, b, rank, world, device="cuda", t1=100.0)
    ndde_s = time.perf_counter() - t0
    if rank == 0:
        i28 = n_total // 2
        ref = np.array([0.90566, 0.0, -14.57233])          # tests/lyapunov_exponents.rs:303-309
        res = {"n_gpus": world, "lorenz_grid_members": n_total, "members_per_gpu": m, "tmax": tmax,
               "kernel": "ode_fixed_periodic_kernel", "kernel_ms_max_over_ranks": float(km[0]),
               "rk_steps_per_s_all_gpus": n_total * (tmax / 0.005 + tmax / 0.5) / (float(km[0]) * 1e-3),
               "gather": "dfb_gather_final: ncclAllGather of (t, y[12], aux[3]) = %d B per member" % (16 * 8),
               "gather_bytes_total": int(n_total * 16 * 8), "gather_ms_max_over_ranks": float(km[1]),
               "spectrum_rho28": spectra[i28].tolist(), "abs_err_vs_reference_rho28": np.abs(spectra[i28] - ref).tolist(),
               "reference_tolerance": 7e-5 + 50.0 / tmax,
               "trace_identity_max_err": float(np.max(np.abs(spectra.sum(axis=1) + (10.0 + 1.0 + 8.0 / 3.0)))),
               "ndde_grid_members": int(a.size), "ndde_sharded_run_plus_all_gather_s": ndde_s,
               "ndde_leading_exponent_range": [float(np.nanmin(lam)), float(np.nanmax(lam))]}
        assert np.all(np.abs(spectra[i28] - ref) < 7e-5 + 50.0 / tmax)
        out.write(json.dumps(res) + "\n")
        out.flush()
    dist.barrier()
    comm.destroy()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
