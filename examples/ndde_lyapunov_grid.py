"""BASELINE configs[4]: a Lyapunov-exponent batch over an NDDE parameter grid.

x'(t) = a x(t - 1) + b x'(t - 1)   (the neutral test equation of tests/equation_types.rs:139-158,
with free coefficients), history sin(t), rktp64, h = 0.02, t in [0, 100].  One member per (a, b);
the leading exponent is the growth rate of the state-segment norm, estimated from the `Trace` samples
(diffurch_b200/lyapunov.py).  The reference has no (N)DDE Lyapunov test, so the check printed here is
against the rightmost characteristic root, lambda = (a + b lambda) exp(-lambda).

Under torchrun (one process per GPU) the grid is sharded by member range and the exponents are
gathered over NCCL."""
import os

import numpy as np

import _common  # noqa: F401
import diffurch_b200 as d
from diffurch_b200 import lyapunov


def main(na=64, nb=64, t1=100.0, system=None, check=16):
    system = system or d.systems.NddeLinear
    a, b = [x.ravel() for x in np.meshgrid(np.linspace(-2.0, 0.5, na), np.linspace(-0.6, 0.6, nb))]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        import torch
        import torch.distributed as dist
        rank = int(os.environ["RANK"])
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
        if not dist.is_initialized():
            dist.init_process_group("nccl")
        lam = lyapunov.sharded_leading_exponents(system, a, b, rank, world, device="cuda", t1=t1)
    else:
        rank = 0
        solver, window = lyapunov.linear_delay_grid(system, a, b, t1=t1)
        st = solver.run()
        assert st.totals["kernel_kind"] == d.abi.KERNEL_DDE_FIXED
        lam = lyapunov.leading_exponent(st.trace_t, st.trace_p, st.trace_n, window, t_skip=20.0)
    worst = 0.0
    pick = np.linspace(0, a.shape[0] - 1, check).astype(int) if check else []
    for i in pick:
        worst = max(worst, abs(lam[i] - lyapunov.rightmost_root(system, a[i], b[i])))
    if rank == 0:
        print(f"{a.shape[0]} members: exponents in [{np.nanmin(lam):.4f}, {np.nanmax(lam):.4f}], "
              f"{int(np.sum(lam > 0))} unstable; max |estimate - rightmost root| over {len(pick)} members = {worst:.2e}")
    return lam, a, b, worst


if __name__ == "__main__":
    main()
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()
