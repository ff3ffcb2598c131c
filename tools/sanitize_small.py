#!/usr/bin/env python3
"""Small invocation of every kernel family, meant to run under compute-sanitizer:
    compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
import diffurch_b200 as d  # noqa: E402

auto = d.AutomaticStepsize(1e-2, (1e-6, 1e-1), [1e-9] * 3, [1e-9] * 3, 4, 0.9, (0.2, 5.0))
runs = {
    "ode_fixed fast (ragged 333)": lambda: cases.lorenz(n=333, t1=0.5, arith="fast"),
    "ode_fixed strict": lambda: cases.lorenz(n=333, t1=0.5, arith="strict"),
    "ode_adaptive": lambda: cases.lorenz(n=333, t1=0.5, arith="fast").stepsize(auto),
    "ode_fixed_periodic": lambda: cases.lyap_lorenz(2.0, rho=np.linspace(20, 40, 77), arith="fast"),
    "ode_event bouncing ball": lambda: cases.bouncing_ball(n=97, t1=2.0, arith="fast"),
    "ode_event egg billiard + trace": lambda: cases.egg_billiard(n=65, reflections=5, arith="strict"),
    "dde_fixed (coefficient ring)": lambda: cases.eq_dde_sin(k=0.5 + np.arange(70) / 69.0, trace=False).arith("fast"),
    "general: dde strict": lambda: cases.eq_dde_sin(k=0.5 + np.arange(70) / 69.0, trace=False),
    "general: dde relay clamp (propagation + events)": lambda: cases.dde_relay_clamp(t1=3.0),
    "general: adaptive + trace": lambda: cases.lorenz(n=50, t1=0.5).stepsize(auto).on_step(d.Trace(1, 256)),
}
for name, mk in runs.items():
    r = mk().run()
    print(f"{name:50s} kernel_kind={r.totals['kernel_kind']} steps={int(r.steps.sum())} failed={r.totals['n_failed']}")
print("done")
