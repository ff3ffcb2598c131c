#!/usr/bin/env python3
"""Freeze the adaptive-path results of the INDEPENDENT numpy restatement (tests/test_oracle_independent.py
`run_numpy_adaptive`, written from stepsize.rs:64-85 and solver.rs:292-297) as hex floats:

    python tests/golden/make_adaptive_golden.py > tests/golden/adaptive_golden.json

Two variants per case: the reference's formula with libm `pow` (Python's math.pow = glibc pow = Rust's
f64::powf on Linux) and the portable n-th root the STRICT CUDA build uses.  The oracle and the GPU are
tested against this file; regenerate only when the restatement itself is corrected."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import test_oracle_independent as ti  # noqa: E402

out = {"generator": "tests/golden/make_adaptive_golden.py", "cases": {}}
for name, sys_name, rk_name, y0, prm, t0, t1, ctl in ti.adaptive_cases():
    rhs = ti.SYSTEMS[sys_name][3](prm)
    tab, b2 = ti.tableau(rk_name), ti.tableau_b2(rk_name)
    entry = {"t0": float(t0).hex(), "t1": float(t1).hex(), "y0": [float(v).hex() for v in y0],
             "params": None if prm is None else [float(v).hex() for v in prm], "controller": ctl}
    for root_name, root in (("libm_pow", None), ("portable_root", ti.portable_root)):
        t, p, steps, rejected, evals, h = ti.run_numpy_adaptive(rhs, y0, t0, t1, ctl, tab, b2, root=root)
        entry[root_name] = {"steps_rejected_evals": [steps, rejected, evals], "t": float(t).hex(),
                            "p": [float(v).hex() for v in p], "final_stepsize": float(h).hex()}
    out["cases"][name] = entry
print(json.dumps(out, indent=1))
