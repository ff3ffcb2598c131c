"""A second, independent restatement of the reference's fixed-step path — numpy, written from
SURVEY Appendix A.1/A.2 (solver.rs:292-310, state.rs:94-120) without looking at oracle/ — against
the C++ oracle, bit for bit, over every rk.rs tableau (coefficients from tests/golden/tableaux.json,
which is generated from the reference source).  Two restatements that agree to the last bit on
random inputs are unlikely to share a misreading; the oracle is the checker of every GPU test.

The second half does the same for the ADAPTIVE path, which no reference test exercises
(`AutomaticStepsize::update`, stepsize.rs:64-85, and the reject loop of `Solver::run`, solver.rs:293-297):
`run_numpy_adaptive` is written from the reference source, uses Python's `math.pow` (the same glibc `pow`
the oracle links, the reference's `f64::powf` on Linux) and must agree with the oracle bit for bit; its
results are frozen in tests/golden/adaptive_golden.json (tests/golden/make_adaptive_golden.py)."""
import json
import math
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


# ------------------------------------------------------------------------------------------------
# adaptive path: AutomaticStepsize (stepsize.rs:35-85) + the reject loop (solver.rs:292-297)
# ------------------------------------------------------------------------------------------------
def tableau_b2(name):
    return _vec(TABLEAUX[name]["b2"])


def f64_clamp(x, lo, hi):  # f64::clamp: NaN stays NaN
    if x < lo:
        return lo
    if x > hi:
        return hi
    return x


def portable_root(x, n):
    """Python restatement of oracle/diffurch_oracle.hpp `portable_root` (same IEEE operations)."""
    if x != x:
        return x
    if x < 0.0:
        return math.nan
    if x == 0.0 or x == math.inf or n <= 1:
        return x
    m, e = math.frexp(x)
    q, r = int(e / n), 0          # C integer division truncates toward zero
    r = e - q * n
    if r < 0:
        r += n
        q -= 1
    a = math.ldexp(m, r)
    y = a if a > 1.0 else 1.0
    for _ in range(48):
        p = 1.0
        for _ in range(n - 1):
            p = p * y
        y = ((n - 1.0) * y + a / p) / float(n)
    return math.ldexp(y, q)


def run_numpy_adaptive(rhs, y0, t0, t1, ctl, tab, b2, root=None):
    """Solver::run with an AutomaticStepsize controller and no events.  `ctl` = dict(stepsize, range, atol,
    rtol, order, fac, fac_range).  `root(x, n)` replaces (1/err).powf(1/(order+1)) when given."""
    a, b, c = tab
    S = len(b)
    h = ctl["stepsize"]                                   # init() is never called upstream
    t, p = t0, np.array(y0, dtype=np.float64)
    t_prev, p_prev, d_curr, d_prev = t0, p.copy(), np.zeros_like(p), np.zeros_like(p)
    steps = rejected = evals = 0
    floor_hit = False
    while t < t1:                                         # solver.rs:292
        while True:
            hs = min(h, t1 - t)                           # solver.rs:293 / :296
            # ---- State::make_step (state.rs:94-120)
            if t_prev != t:
                k = [d_curr]
            else:
                k = [rhs(t, p)]
                evals += 1
            t_prev, p_prev, d_prev = t, p, d_curr
            for i in range(1, S):
                acc = np.zeros_like(p)
                for j in range(i):
                    acc = acc + k[j] * a[i, j]
                k.append(rhs(t_prev + c[i] * hs, p_prev + acc * hs))
                evals += 1
            acc = np.zeros_like(p)
            for j in range(S):
                acc = acc + k[j] * b[j]
            p = p_prev + acc * hs
            t = t_prev + hs
            d_curr = rhs(t, p)
            evals += 1
            acc = np.zeros_like(p)
            for j in range(S):
                acc = acc + k[j] * (b2[j] - b[j])
            e = acc * hs
            # ---- AutomaticStepsize::update (stepsize.rs:64-85)
            err = None
            for n in range(len(p)):
                ae = abs(float(e[n]))
                v = ae / (ctl["atol"][n] + ae * ctl["rtol"][n])
                err = v if err is None else (v if (err != err or v > err) else err)   # f64::max
            inv = math.inf if err == 0.0 else 1.0 / err
            if root is None:
                factor = ctl["fac"] * (math.pow(inv, 1.0 / float(ctl["order"] + 1)) if inv == inv else inv)
            else:
                factor = ctl["fac"] * root(inv, ctl["order"] + 1)
            factor = f64_clamp(factor, *ctl["fac_range"])
            h_before = h
            h = f64_clamp(h * factor, *ctl["range"])
            if not err >= 1.0:
                break                                      # Accepted
            rejected += 1
            t, p, d_curr = t_prev, p_prev, d_prev          # undo_step (state.rs:146-150)
            if h == h_before:                              # library limit (DFB_TRAJ_STEPSIZE_FLOOR): the reference retries forever
                floor_hit = True
                break
        if floor_hit:
            break
        steps += 1                                         # commit_step; on_step
    return t, p, steps, rejected, evals, h


ADAPTIVE_TABLEAUX = ["rktp64", "rk43", "rk4", "three_eights", "heun3", "midpoint"]
ADAPTIVE_ORDER = {"rktp64": 4, "rk43": 3, "rk4": 2, "three_eights": 2, "heun3": 1, "midpoint": 1}  # order_embedded


def adaptive_cases():
    """(name, system key, tableau, y0, prm, t0, t1, ctl) — fixed inputs, also what the golden file freezes."""
    out = []
    for rk_name in ADAPTIVE_TABLEAUX:
        rng = np.random.default_rng(zlib.crc32(("adaptive:" + rk_name).encode()))
        for sys_name, (system, N, NP, make_rhs) in SYSTEMS.items():
            y0 = rng.uniform(-2.0, 2.0, N) + (np.array([0, 0, 20.0]) if sys_name == "lorenz" else 0.0)
            prm = rng.uniform(0.5, 3.0, NP) if NP else None
            if sys_name == "lorenz":
                prm = np.array([10.0, float(rng.uniform(20, 35)), 8.0 / 3.0])
            if sys_name == "linear1":
                prm = -prm                                   # decaying: the controller grows the step to its cap
            # tolerances a method of that order reaches in a few hundred steps (the floor of the step-size
            # range is a library limit, DFB_TRAJ_STEPSIZE_FLOOR, not reference behaviour: stay off it)
            tol = float(rng.choice([1e-6, 1e-8, 1e-10] if ADAPTIVE_ORDER[rk_name] >= 3 else [1e-2, 3e-3]))
            ctl = {"stepsize": 1e-2, "range": (1e-9, 0.25), "atol": [tol] * N, "rtol": [tol * 10.0] * N,
                   "order": ADAPTIVE_ORDER[rk_name], "fac": 0.9, "fac_range": (0.2, 5.0)}
            t0 = float(rng.uniform(-1.0, 1.0))
            t1 = t0 + float(rng.uniform(0.5, 2.0))
            out.append((f"{rk_name}/{sys_name}", sys_name, rk_name, y0, prm, t0, t1, ctl))
    return out


def adaptive_solver(sys_name, rk_name, y0, prm, t0, t1, ctl, arith="strict"):
    system = SYSTEMS[sys_name][0]
    auto = d.AutomaticStepsize(ctl["stepsize"], ctl["range"], ctl["atol"], ctl["rtol"], ctl["order"], ctl["fac"],
                               ctl["fac_range"])
    s = (d.Solver.new().initial(np.asarray(y0).reshape(1, -1)).interval(t0, t1).stepsize(auto)
         .rk(d.RK.by_name(rk_name)).arith(arith))
    return s.equation(system, np.asarray(prm).reshape(1, -1)) if prm is not None else s.equation(system)


@pytest.mark.parametrize("case", adaptive_cases(), ids=lambda c: c[0])
def test_adaptive_oracle_equals_independent_numpy_restatement(oracle, case):
    name, sys_name, rk_name, y0, prm, t0, t1, ctl = case
    rhs = SYSTEMS[sys_name][3](prm)
    tab, b2 = tableau(rk_name), tableau_b2(rk_name)
    o = oracle.run_solver(adaptive_solver(sys_name, rk_name, y0, prm, t0, t1, ctl), n_threads=1)
    t, p, steps, rejected, evals, _ = run_numpy_adaptive(rhs, y0, t0, t1, ctl, tab, b2)
    assert (o.steps[0], o.rejected[0], o.rhs_evals[0]) == (steps, rejected, evals), name
    assert o.t[0] == t and np.array_equal(o.p[0], p), name
    assert steps > 3
    # the same with the portable n-th root in both (what the STRICT CUDA build computes)
    o2 = oracle.run_solver(adaptive_solver(sys_name, rk_name, y0, prm, t0, t1, ctl), n_threads=1, portable_root=True)
    t2, p2, steps2, rejected2, evals2, _ = run_numpy_adaptive(rhs, y0, t0, t1, ctl, tab, b2, root=portable_root)
    assert (o2.steps[0], o2.rejected[0], o2.rhs_evals[0]) == (steps2, rejected2, evals2), name
    assert o2.t[0] == t2 and np.array_equal(o2.p[0], p2), name
    # libm pow and the portable root differ in the last bits of the factor only: same accept/reject
    # sequence on these problems, states equal to 1e-12
    assert (steps2, rejected2) == (steps, rejected), name
    assert np.allclose(p2, p, rtol=1e-12, atol=1e-13), name


def test_adaptive_rejections_are_exercised_and_frozen(oracle):
    """The golden file pins the numbers themselves (hex floats), so a change of the oracle, of numpy's or of
    libm's behaviour shows up as a diff of a committed file rather than as two sides drifting together."""
    gold = json.load(open(os.path.join(HERE, "golden", "adaptive_golden.json")))
    total_rejected = 0
    for name, sys_name, rk_name, y0, prm, t0, t1, ctl in adaptive_cases():
        g = gold["cases"][name]
        for root_name, portable in (("libm_pow", False), ("portable_root", True)):
            o = oracle.run_solver(adaptive_solver(sys_name, rk_name, y0, prm, t0, t1, ctl), n_threads=1,
                                  portable_root=portable)
            e = g[root_name]
            assert [int(o.steps[0]), int(o.rejected[0]), int(o.rhs_evals[0])] == e["steps_rejected_evals"], (name, root_name)
            assert float(o.t[0]).hex() == e["t"] and [float(v).hex() for v in o.p[0]] == e["p"], (name, root_name)
        total_rejected += g["libm_pow"]["steps_rejected_evals"][1]
    assert total_rejected > 10  # the reject loop (undo_step + k[0] re-evaluation) is really on the tested path


def test_portable_root_matches_its_python_restatement_and_pow(oracle):
    L = oracle.lib()
    rng = np.random.default_rng(7)
    for _ in range(4000):
        x = float(np.exp(rng.uniform(-9.0, 9.0)))
        n = int(rng.integers(2, 8))
        r = L.orc_portable_root(x, n)
        assert r == portable_root(x, n)
        assert abs(r - math.pow(x, 1.0 / n)) <= 2.0 * math.ulp(r)
    for x in (0.0, math.inf, 5e-324, 1e-310, 1.0, 2.0 ** 1000):
        assert L.orc_portable_root(x, 5) == portable_root(x, 5)
    assert math.isnan(L.orc_portable_root(math.nan, 5)) and math.isnan(L.orc_portable_root(-1.0, 5))
