"""A second, independent restatement of the reference's fixed-step path — numpy, written from
SURVEY Appendix A.1/A.2 (solver.rs:292-310, state.rs:94-120) without looking at oracle/ — against
the C++ oracle, bit for bit, over every rk.rs tableau (coefficients from tests/golden/tableaux.json,
which is generated from the reference source).  Two restatements that agree to the last bit on
random inputs are unlikely to share a misreading; the oracle is the checker of every GPU test."""
import json
import os

import zlib

import numpy as np
import pytest

import cases  # noqa: F401  (puts the package on the path the same way the other tests do)
import diffurch_b200 as d

HERE = os.path.dirname(os.path.abspath(__file__))
TABLEAUX = json.load(open(os.path.join(HERE, "golden", "tableaux.json")))["tableaux"]


def _vec(v):
    return np.array([float.fromhex(x) if isinstance(x, str) else float(x) for x in v])


def tableau(name):
    t = TABLEAUX[name]
    S = len(t["b"])
    return np.array([_vec(r) for r in t["a"]]).reshape(S, S), _vec(t["b"]), _vec(t["c"])


def run_numpy(rhs, y0, t0, t1, h, tab):
    a, b, c = tab
    S = len(b)
    t, p, d_curr, steps, evals = t0, np.array(y0, dtype=np.float64), None, 0, 0
    while t < t1:                                   # solver.rs:292
        hs = min(h, t1 - t)                         # solver.rs:293
        if d_curr is None:                          # state.rs:97-98 (first step: t_prev == t_curr)
            k = [rhs(t, p)]
            evals += 1
        else:
            k = [d_curr]                            # state.rs:95-96
        for i in range(1, S):                       # state.rs:105-110
            acc = np.zeros_like(p)
            for j in range(i):
                acc = acc + k[j] * a[i, j]          # includes the structural zeros, as upstream
            k.append(rhs(t + c[i] * hs, p + acc * hs))
            evals += 1
        acc = np.zeros_like(p)
        for j in range(S):
            acc = acc + k[j] * b[j]
        p = p + acc * hs                            # state.rs:112-113
        t = t + hs
        d_curr = rhs(t, p)                          # state.rs:115
        evals += 1
        steps += 1
    return t, p, steps, evals


SYSTEMS = {
    "lorenz": (d.systems.Lorenz, 3, 3,
               lambda prm: (lambda t, p: np.array([prm[0] * (p[1] - p[0]), p[0] * (prm[1] - p[2]) - p[1],
                                                   p[0] * p[1] - prm[2] * p[2]]))),
    "harmonic": (d.systems.Harmonic, 2, 0, lambda prm: (lambda t, p: np.array([p[1], -p[0]]))),
    "linear1": (d.systems.Linear1, 1, 1, lambda prm: (lambda t, p: np.array([p[0] * prm[0]]))),
    "time3": (d.systems.Time3, 3, 0, lambda prm: (lambda t, p: np.array([0.0, 1.0, t]))),
}


@pytest.mark.parametrize("rk_name", sorted(TABLEAUX))
def test_oracle_equals_independent_numpy_restatement(oracle, rk_name):
    rng = np.random.default_rng(zlib.crc32(rk_name.encode()))  # str hashes are salted per process
    tab = tableau(rk_name)
    for sys_name, (system, N, NP, make_rhs) in SYSTEMS.items():
        n = 3
        y0 = rng.uniform(-2.0, 2.0, (n, N)) + (np.array([0, 0, 20.0]) if sys_name == "lorenz" else 0.0)
        prm = rng.uniform(0.5, 3.0, (n, NP)) if NP else None
        if sys_name == "lorenz":
            prm = np.stack([np.full(n, 10.0), rng.uniform(20, 35, n), np.full(n, 8.0 / 3.0)], 1)
        t0 = float(rng.uniform(-1.0, 1.0))
        h = float(rng.choice([0.01, 1.0 / 64, 0.037]))
        t1 = t0 + float(rng.uniform(3.0, 25.0)) * h
        s = d.Solver.new().initial(y0).interval(t0, t1).stepsize(h).rk(d.RK.by_name(rk_name)).arith("strict")
        s = s.equation(system, prm) if prm is not None else s.equation(system)
        o = oracle.run_solver(s, n_threads=1)
        for i in range(n):
            t, p, steps, evals = run_numpy(make_rhs(prm[i] if prm is not None else None), y0[i], t0, t1, h, tab)
            assert o.steps[i] == steps and o.rhs_evals[i] == evals, (rk_name, sys_name)
            assert o.t[i] == t, (rk_name, sys_name)
            assert np.array_equal(o.p[i], p, equal_nan=True), (rk_name, sys_name, o.p[i], p)
