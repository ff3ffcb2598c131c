"""examples/ode-linear-1-exp.rs: x' = k x, k = ln 2, unit steps, so x(n) = 2^n.  The reference runs
this one in f32 (for its plotting crate); the device path is FP64."""
import math

import _common  # noqa: F401
import numpy as np

import diffurch_b200 as d


def main():
    k = math.log(2.0)
    st = (d.Solver.new().initial(np.ones((1, 1))).equation(d.systems.Linear1, np.array([[k]]))
          .interval(0.0, 16.0).stepsize(1.0).on_step(d.Trace(1, 32)).run())
    for t, x in zip(st.trace_t[0, :st.trace_n[0]], st.trace_p[0, :st.trace_n[0], 0]):
        print(f"t={t:<3.0f} x={x:<10.2f} rel error={(x - 2.0 ** t) / x:<8.0e}")
    return st


if __name__ == "__main__":
    main()
