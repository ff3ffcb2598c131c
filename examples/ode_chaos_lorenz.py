"""examples/ode-chaos-lorenz.rs of the reference, as an ensemble over rho.

The reference integrates ONE trajectory to t = 2000 and records (t, x, y, z) from t >= 50 in an
`on_step` closure for a plot.  Here every member of a rho sweep is integrated; the `on_step`
closure becomes the output policy `Trace(every, capacity)`."""
import _common  # noqa: F401
import numpy as np

import diffurch_b200 as d


def main(n=4096, t1=60.0, trace_every=25):
    sigma, beta = 10.0, 8.0 / 3.0
    rho = np.linspace(20.0, 40.0, n) if n > 1 else np.array([28.0])
    prm = np.stack([np.full(n, sigma), rho, np.full(n, beta)], axis=1)
    st = (d.Solver.new()
          .initial(np.tile([1.0, 2.0, 20.0], (n, 1)))
          .equation(d.systems.Lorenz, prm)           # |s| [sigma (y - x), x (rho - z) - y, x y - beta z]
          .interval(0.0, t1)
          .stepsize(1.0 / 250.0)
          .on_step(d.Trace(trace_every, int(t1 * 250 / trace_every) + 2))
          .arith("fast")
          .run())
    keep = st.trace_t[0, :st.trace_n[0]] >= 50.0
    print(f"{n} members, {int(st.steps[0])} steps each, kernel {st.kernel_ms:.2f} ms; "
          f"{int(keep.sum())} recorded points per member with t >= 50")
    return st


if __name__ == "__main__":
    main()
