"""examples/ode-linear-3-fibonacci.rs: a 3-dimensional linear system whose first component is the
n-th Fibonacci number at t = n; `Periodic::new(1.)` fires at the integers."""
import _common  # noqa: F401
import numpy as np

import diffurch_b200 as d


def main():
    a = [-0.21520447048200203, 0.43040894096400406, 3.141592653589793,
         0.43040894096400406, 0.21520447048200203, -1.941611038725466,
         -2.2732777998989695, 1.4049629462081452, -0.4812118250596034]
    st = (d.Solver.new().initial([0.0, 1.0, 0.0]).equation(d.systems.Linear3, a).interval(0.0, 50.0)
          .stepsize(0.1).on(d.Periodic.new(1.0), None).log_events(d.EventLog(64)).on_step(d.Trace(1, 1024)).run())
    t, x = st.trace_t[0, :st.trace_n[0]], st.trace_p[0, :st.trace_n[0], 0]
    for te in st.event_t[0, :st.event_n[0]][:12]:
        print(f"f_{te:02.0f} = {x[np.argmax(t == te)]:14.2f}")
    return st


if __name__ == "__main__":
    main()
