#!/usr/bin/env python3
"""Freeze a few outputs of the CPU oracle (hex floats) as regression fixtures:

  * BASELINE configs[0]: the single Lorenz trajectory of examples/ode-chaos-lorenz.rs
    (sigma = 10, rho = 28, beta = 8/3, y0 = (1, 2, 20), rktp64, h = 1/250, t in [0, 50]);
  * the located event times of examples/bouncing-ball.rs on [0, 8.58] and the first 40 reflections
    of examples/egg-billiard.rs;
  * the DDE / NDDE sine solutions of tests/equation_types.rs at t = 10.

The reference itself cannot run here (nightly Rust, no toolchain), so these are NOT reference
outputs: they pin the oracle — which is pinned to the reference's own assertions by
tests/test_oracle_reference_tests.py — against silent drift, and give the STRICT CUDA path a
bit-exact target that does not need the oracle at run time.  Re-run after a deliberate change:
    python tests/golden/make_oracle_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cases  # noqa: E402
from oracle import oracle as orc  # noqa: E402


def hexes(a):
    return [float(x).hex() for x in np.asarray(a, dtype=np.float64).ravel()]


def build():
    orc.build()
    out = {}
    r = orc.run_solver(cases.lorenz(n=1, t1=50.0))
    out["lorenz_single_t50"] = {"t": hexes(r.t), "p": hexes(r.p), "steps": int(r.steps[0]), "rhs_evals": int(r.rhs_evals[0])}
    r = orc.run_solver(cases.bouncing_ball())
    n = int(r.event_n[0])
    out["bouncing_ball_t8.58"] = {"event_t": hexes(r.event_t[0, :n]), "event_index": [int(i) for i in r.event_index[0, :n]],
                                  "p": hexes(r.p), "steps": int(r.steps[0])}
    r = orc.run_solver(cases.egg_billiard(reflections=40))
    n = int(r.event_n[0])
    out["egg_billiard_40"] = {"event_t": hexes(r.event_t[0, :n]), "p": hexes(r.p), "steps": int(r.steps[0])}
    r = orc.run_solver(cases.eq_dde_sin(trace=False))
    out["dde_sin_t10"] = {"p": hexes(r.p), "steps": int(r.steps[0])}
    r = orc.run_solver(cases.eq_ndde_sin())
    out["ndde_sin_t10"] = {"p": hexes(r.p), "steps": int(r.steps[0])}
    return out


if __name__ == "__main__":
    with open(os.path.join(HERE, "oracle_golden.json"), "w") as f:
        json.dump(build(), f, indent=1)
    print("wrote", os.path.join(HERE, "oracle_golden.json"))
