#!/usr/bin/env python3
"""Kernel time of the 2^18-member DDE config (BASELINE configs[2]) — tuning helper."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
import diffurch_b200 as d  # noqa: E402

n = int(sys.argv[sys.argv.index("--members") + 1]) if "--members" in sys.argv else 1 << 18
k = 0.5 + np.arange(n) / (n - 1.0)
prm = cases.dde_params(k)
ens = (d.Solver.new().max_delay(1.0).initial(d.InitFn(d.SIN)).interval(0.0, 100.0).stepsize(0.02)
       .equation(d.systems.DdeLinear, prm).arith("fast").build())
ms = []
for i in range(6):
    ens.run(); ens.synchronize()
    if i >= 2:
        ms.append(ens.totals()["kernel_ms"])
tot = ens.totals()
t, y, _ = ens.final()
print(json.dumps({"env": {k_: v for k_, v in os.environ.items() if k_.startswith("DFB_")}, "kernel_ms": float(np.mean(ms)),
                  "min_ms": float(np.min(ms)), "steps_per_s": tot["steps"] / (np.mean(ms) * 1e-3),
                  "gbs": 96.0 * tot["steps"] / (np.mean(ms) * 1e-3) / 1e9, "n_failed": tot["n_failed"],
                  "err": float(np.max(np.abs(y[:, 0] - np.sin(k * t))))}))
