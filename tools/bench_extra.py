#!/usr/bin/env python3
"""Throughput of the non-headline BASELINE configs (events, Lyapunov batch, adaptive,
strict arithmetic) on one GPU.  Prints one JSON object; used for profiles/, not by the driver."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cases  # noqa: E402
import diffurch_b200 as d  # noqa: E402
from diffurch_b200 import lyapunov  # noqa: E402


def timed(make, reps=2):
    ens = make().build()
    ens.run()
    ens.synchronize()
    ms = []
    for _ in range(reps):
        ens.run()
        ens.synchronize()
        ms.append(ens.totals()["kernel_ms"])
    tot = ens.totals()
    ens.close()
    ms = float(np.mean(ms))
    return {"kernel_ms": ms, "steps": tot["steps"], "events": tot["events"], "rejected": tot["rejected"],
            "rhs_evals": tot["rhs_evals"], "n_failed": tot["n_failed"], "kernel_kind": tot["kernel_kind"],
            "steps_per_s": tot["steps"] / (ms * 1e-3), "events_per_s": tot["events"] / (ms * 1e-3)}


def relay_clamp(n, arith):
    # examples/dde-relay-clamp.rs:19-36 as an ensemble over alpha (DDE + propagation + located events)
    alpha = np.linspace(3.0, 7.0, n)
    prm = np.stack([alpha, np.full(n, 0.4), np.full(n, 1.0)], axis=1)
    y0 = np.stack([np.ones(n), -alpha], axis=1)
    return (d.Solver.new().initial(y0).equation(d.systems.DdeRelayClamp, prm).interval(0.0, 30.0).stepsize(0.012)
            .with_const_delay(1.0, 2).on(d.Locator.zero("abs_delayed_x_minus_1"), None).arith(arith))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0, help="fraction of the BASELINE ensemble sizes")
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    out = {}
    n_ev = int((1 << 19) * args.scale)
    n_ly = int((1 << 16) * args.scale)
    n_lo = int((1 << 20) * args.scale)
    todo = {
        "bouncing_ball_2^19_t8_fast": lambda: cases.bouncing_ball(n=n_ev, t1=8.0, arith="fast", events=0),
        "bouncing_ball_2^19_t8_strict": lambda: cases.bouncing_ball(n=n_ev, t1=8.0, arith="strict", events=0),
        "egg_billiard_2^19_200refl_fast": lambda: cases.egg_billiard(n=n_ev, reflections=200, arith="fast", events=0),
        "lyapunov_lorenz_2^16_rho_grid_t100_fast": lambda: cases.lyap_lorenz(100.0, rho=np.linspace(20.0, 40.0, n_ly), arith="fast"),
        "lorenz_2^20_adaptive_fast": lambda: cases.lorenz(n=n_lo, t1=50.0, arith="fast").stepsize(
            d.AutomaticStepsize(1e-2, (1e-6, 1e-1), [1e-9] * 3, [1e-9] * 3, 4, 0.9, (0.2, 5.0))),
        "dde_relay_clamp_2^16_t30_fast": lambda: relay_clamp(n_ly, "fast"),
        "dde_relay_clamp_2^16_t30_strict": lambda: relay_clamp(n_ly, "strict"),
        "lorenz_2^16_adaptive_with_events_fast": lambda: cases.lorenz(n=n_ly, t1=20.0, arith="fast").stepsize(
            d.AutomaticStepsize(1e-2, (1e-6, 1e-1), [1e-9] * 3, [1e-9] * 3, 4, 0.9, (0.2, 5.0))).on(d.Periodic(0.5, 0.0), None),
        "ndde_lyapunov_grid_2^16_t100_traced_every5_fast": lambda: lyapunov.linear_delay_grid(
            d.systems.NddeLinear, *[x.ravel() for x in np.meshgrid(np.linspace(-2.0, 0.5, 256),
                                                                     np.linspace(-0.6, 0.6, max(1, n_ly // 256)))])[0],
        "dde_sin_2^16_t20_traced_every50_fast": lambda: cases.eq_dde_sin(
            k=0.5 + np.arange(n_ly) / n_ly, t1=20.0, trace=False, arith="fast").on_step(d.Trace(50, 24)),
        "lorenz_2^18_t10_traced_every25_fast": lambda: cases.lorenz(n=n_lo // 4, t1=10.0, arith="fast",
                                                                     trace=d.Trace(25, 104)),
        "lorenz_2^18_adaptive_traced_every25_fast": lambda: cases.lorenz(n=n_lo // 4, t1=10.0, arith="fast", trace=d.Trace(25, 104)).stepsize(
            d.AutomaticStepsize(1e-2, (1e-6, 1e-1), [1e-9] * 3, [1e-9] * 3, 4, 0.9, (0.2, 5.0))),
        "lorenz_2^20_rk4_fast": lambda: cases.lorenz(n=n_lo, t1=50.0, arith="fast", rk=d.RK.rk4()),
        "lorenz_2^20_rktp64_strict": lambda: cases.lorenz(n=n_lo, t1=50.0, arith="strict"),
    }
    for name, mk in todo.items():
        if args.only and args.only not in name:
            continue
        t0 = time.time()
        try:
            out[name] = timed(mk)
        except Exception as e:  # noqa: BLE001
            out[name] = {"error": repr(e)}
        out[name]["wall_s"] = time.time() - t0
        print(name, json.dumps(out[name]), file=sys.stderr, flush=True)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
