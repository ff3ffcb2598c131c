"""examples/egg-billiard.rs: free flight inside x^2 + y^2 - y^3/3 = 1, specular reflection at the
boundary, `stop_integration` after a number of reflections.  Ensemble over the launch angle."""
import math

import _common  # noqa: F401
import numpy as np

import diffurch_b200 as d


def main(n=4096, reflections=200):
    phi = 2.0 * np.pi * np.arange(n) / n
    y0 = np.stack([np.zeros(n), np.full(n, 0.1), 0.5 * np.cos(phi), 0.5 * np.sin(phi)], axis=1)
    st = (d.Solver.new()
          .initial(y0)
          .equation(d.systems.EggBilliard, np.full((n, 1), float(reflections)))
          .interval(0.0, math.inf)                                        # `0. ..`
          .stepsize(0.5)
          .on_mut(d.Locator.above_zero("boundary"), "reflect")
          .max_steps(100 * reflections + 1000)
          .arith("fast")
          .run())
    speed = np.hypot(st.p[:, 2], st.p[:, 3])
    print(f"{n} members stopped after {int(st.events[0])} reflections, {int(st.steps.sum())} steps; "
          f"max |speed - 0.5| = {np.max(np.abs(speed - 0.5)):.1e}")
    return st


if __name__ == "__main__":
    main()
