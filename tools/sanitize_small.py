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
    "dde_event: relay clamp ensemble (ragged 333, events + propagation)":
        lambda: cases.dde_relay_clamp_ensemble(333, arith="fast", t1=6.0, events=64, trace=d.Trace(5, 128)),
    "dde_event: relay clamp, small ring (overflow path)":
        lambda: cases.dde_relay_clamp_ensemble(40, arith="fast", t1=4.0, events=16).ring_capacity(64),
    "dde_event: propagated delays, euler (tests/delay_propagation.rs)": lambda: cases.delay_constant_2().arith("fast"),
    "dde_event: pantograph, unbounded max_delay": lambda: cases.delay_pantograph().arith("fast"),
    "dde_event: plain DDE + max_steps": lambda: cases.eq_dde_sin(k=0.5 + np.arange(40) / 39.0, t1=3.0, arith="fast").max_steps(100000),
    "dde_fixed PRE build (small shard)": lambda: cases.eq_dde_sin(k=0.5 + np.arange(2000) / 1999.0, t1=3.0, trace=False).arith("fast"),
}
from diffurch_b200 import user  # noqa: E402
vdp = user.load_user_system_from_source(os.path.join(ROOT, "tests", "plugins", "van_der_pol.cuh"), "SysVanDerPol",
                                        params=("mu",), detectors={"x": 1}, callbacks={"count": 1})
mu = np.linspace(0.3, 3.0, 77).reshape(-1, 1)
runs["nvrtc: ode_fixed (source system)"] = lambda: (d.Solver.new().initial(np.tile([2.0, 0.0], (77, 1))).equation(vdp, mu)
                                                    .interval(0.0, 1.0).stepsize(0.01).arith("fast"))
runs["nvrtc: general (events)"] = lambda: (d.Solver.new().initial(np.tile([2.0, 0.0], (77, 1))).equation(vdp, mu)
                                           .interval(0.0, 4.0).stepsize(0.01).arith("strict")
                                           .on_mut(d.Locator.zero("x"), "count"))
for name, mk in runs.items():
    r = mk().run()
    print(f"{name:50s} kernel_kind={r.totals['kernel_kind']} steps={int(r.steps.sum())} failed={r.totals['n_failed']}")
print("done")
