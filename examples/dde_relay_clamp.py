"""examples/dde-relay-clamp.rs: x'' + sigma x' + x = -alpha clamp(x(t - T), -1, 1), with the
discontinuities of the clamp propagated through the delay (`with_const_delay`) and located where
|x(t - T)| = 1.  Ensemble over alpha."""
import _common  # noqa: F401
import numpy as np

import diffurch_b200 as d


def main(n=1024, t1=30.0):
    sigma, T = 0.4, 1.0
    alpha = np.linspace(3.0, 7.0, n)
    st = (d.Solver.new()
          .initial(np.stack([np.ones(n), -alpha], axis=1))
          .equation(d.systems.DdeRelayClamp, np.stack([alpha, np.full(n, sigma), np.full(n, T)], axis=1))
          .interval(0.0, t1 * T)
          .stepsize(0.012)
          .with_const_delay(T, 2)
          .on(d.Locator.zero("abs_delayed_x_minus_1"), None)
          .arith("fast")
          .run())
    print(f"{n} members: {int(st.steps.sum())} steps, {int(st.events.sum())} located switchings, "
          f"{int((st.status != 0).sum())} failed")
    return st


if __name__ == "__main__":
    main()
