"""examples/ode-2-relay-msign.rs: x'' = -2 sign(x) with the sign frozen at the start of each step
(`s.p_prev`), switching located at x = 0."""
import _common  # noqa: F401
import numpy as np

import diffurch_b200 as d


def main(n=256):
    x0 = 0.25 + 0.5 * np.arange(n) / n
    st = (d.Solver.new()
          .initial(np.stack([x0, np.zeros(n)], axis=1))
          .equation(d.systems.RelayMsign)                                 # |s| [y, -2 sign(p_prev.x)]
          .interval(0.5, 10.5)
          .stepsize(0.09)
          .on(d.Locator.zero("x"), None)
          .log_events(d.EventLog(64))
          .run())
    print(f"{n} members: {int(st.events.sum())} switchings; member 0 switches at {st.event_t[0, :4]}")
    return st


if __name__ == "__main__":
    main()
