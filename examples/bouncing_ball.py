"""examples/bouncing-ball.rs: x'' = -g with a restitution bounce at x = 0, located by bisection
on the step's dense output.  Ensemble over the drop height."""
import _common  # noqa: F401
import numpy as np

import diffurch_b200 as d


def main(n=4096, t1=8.0):
    k, g = 0.9, 9.8
    x0 = 1.0 + np.arange(n) / n
    st = (d.Solver.new()
          .initial(np.stack([x0, np.full(n, -0.01)], axis=1))
          .equation(d.systems.BouncingBall, np.tile([g, k], (n, 1)))      # |s| [dx, -g]
          .interval(0.0, t1)
          .stepsize(0.005)
          .on(d.Locator.zero("dx"), None)                                # .on(Locator::zero(|s| s.p.dx), |_| {})
          .on_mut(d.Locator.below_zero("x"), "bounce")                   # |s| { s.p.x = 0.; s.p.dx = k * |s.p.dx| }
          .log_events(d.EventLog(128))
          .arith("strict")
          .run())
    first = st.event_t[0, np.argmax(st.event_index[0] == 1)]
    print(f"{n} members: {int(st.events.sum())} located events, first impact of member 0 at t = {first:.15f} "
          f"(analytic {(-0.01 + np.sqrt(0.0001 + 2 * g * x0[0])) / g:.15f})")
    return st


if __name__ == "__main__":
    main()
