#!/usr/bin/env python3
"""One-GPU timing of the Lyapunov batch (BASELINE configs[4]): Lorenz rho grid, variational system,
h = 0.005, QR every 0.5, t in [0, tmax] (tests/lyapunov_exponents.rs:243-311).  Prints one JSON line.
DFB_PERIODIC_MINB selects the register-capped build of ode_fixed_periodic_kernel (tuning)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import cases
    n = int(os.environ.get("DFB_LYAP_MEMBERS", 1 << 18))
    tmax = float(os.environ.get("DFB_LYAP_TMAX", 100.0))
    rho = np.linspace(20.0, 40.0, n)
    rho[n // 2] = 28.0
    ens = cases.lyap_lorenz(tmax, rho=rho, arith="fast").build()
    ms = []
    for _ in range(4):
        ens.run()
        ens.synchronize()
        ms.append(ens.totals()["kernel_ms"])
    st = ens.collect()
    spec = np.asarray(st.aux)[n // 2] / tmax
    steps = int(np.asarray(st.steps, dtype=np.float64).sum())
    print(json.dumps({"members": n, "tmax": tmax, "kernel_ms": ms, "kernel_ms_min": min(ms[1:]),
                      "steps_per_s": steps / (min(ms[1:]) * 1e-3), "rho28_spectrum": spec.tolist(),
                      "minb": os.environ.get("DFB_PERIODIC_MINB", "default")}))


if __name__ == "__main__":
    main()
