"""An equation the library does not ship: the closure of `Solver::equation` restated as a CUDA
device functor (tests/plugins/van_der_pol.cuh), compiled against the same kernel templates and
registered at run time."""
import os

import _common
import numpy as np

import diffurch_b200 as d
from diffurch_b200 import user


def main(n=4096):
    header = os.path.join(_common.ROOT, "tests", "plugins", "van_der_pol.cuh")
    so = user.build_user_system(header, "SysVanDerPol")           # nvcc, cached by content hash
    vdp = user.load_user_system(so, name="van_der_pol", params=("mu",), detectors={"x": 1}, callbacks={"count": 1})
    mu = np.linspace(0.2, 4.0, n).reshape(-1, 1)
    st = (d.Solver.new().initial(np.tile([2.0, 0.0], (n, 1))).equation(vdp, mu).interval(0.0, 50.0)
          .stepsize(0.01).on_mut(d.Locator.zero("x"), "count").arith("fast").run())
    print(f"{n} members, kernel kind {st.totals['kernel_kind']}, {st.kernel_ms:.2f} ms; zero crossings of x in "
          f"[0, 50]: mu={mu[0, 0]:.1f} -> {int(st.aux[0, 0])}, mu={mu[-1, 0]:.1f} -> {int(st.aux[-1, 0])}")
    return st


if __name__ == "__main__":
    main()
