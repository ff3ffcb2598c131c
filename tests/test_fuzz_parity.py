"""Randomised parity: seeded random Solver configurations, CUDA path (whatever kernel the
dispatcher picks) against the CPU oracle on the same inputs.

STRICT arithmetic must be BIT-IDENTICAL for systems without transcendental calls: final state,
times, step / event / rhs-evaluation counters, status words, event logs (bisection results) and
traces.  Every case prints its seed and configuration on failure; kernel kinds seen are
collected so the run also shows the dispatcher reached every ODE kernel family.
"""
import math
import os

import numpy as np
import pytest

import diffurch_b200 as d
from diffurch_b200 import abi

pytestmark = pytest.mark.gpu

# DFB_FUZZ_SCALE=k multiplies the number of seeds of every test (a longer campaign on demand)
SCALE = max(1, int(os.environ.get("DFB_FUZZ_SCALE", "1")))

RK_NAMES = ["euler", "midpoint", "heun2", "ralston2", "kutta3", "heun3", "ralston3", "wray3", "ssp3", "rk4",
            "three_eights", "rk43", "rktp64"]
SEEN = set()


def pick_rk(rng):
    r = rng.integers(0, len(RK_NAMES) + 4)
    if r < len(RK_NAMES):
        return d.RK.by_name(RK_NAMES[r]), RK_NAMES[r]
    if r == len(RK_NAMES):
        c2 = float(rng.uniform(0.2, 1.0))
        return d.RK.generic_order_2(c2), f"generic_order_2({c2})"
    if r == len(RK_NAMES) + 1:
        c2, c3 = float(rng.uniform(0.2, 0.6)), float(rng.uniform(0.65, 1.0))
        return d.RK.generic_order_3(c2, c3), f"generic_order_3({c2},{c3})"
    return d.RK.rktp64(), "rktp64"      # the default tableau gets extra weight


PANIC_BITS = abi.TRAJ_HISTORY_DELETED | abi.TRAJ_HISTORY_NOT_READY | abi.TRAJ_NO_DERIVATIVE


def same(g, o, what=""):
    # a member the reference would have panicked on has no final State upstream: the oracle leaves
    # NaN there, the device path leaves the state at the moment of the panic; everything else
    # (status word, counters up to the panic, event log, trace) must still agree
    alive = (o.status & PANIC_BITS) == 0
    for name in ("t", "p", "aux", "rhs_evals"):   # (the oracle reads its rhs counter off the returned State)
        x, y = getattr(g, name), getattr(o, name)
        if x is None or y is None or not np.size(x):
            continue
        assert np.array_equal(x[alive], y[alive], equal_nan=True), f"{name} differs {what}"
    for name in ("steps", "rejected", "events", "status"):
        x, y = getattr(g, name), getattr(o, name)
        if x is None or y is None or not np.size(x):
            continue
        assert np.array_equal(x, y, equal_nan=True), f"{name} differs {what}"
    if g.event_n is not None:
        assert np.array_equal(g.event_n, o.event_n), f"event_n differs {what}"
        assert np.array_equal(g.event_t, o.event_t, equal_nan=True), f"event times differ {what}"
        assert np.array_equal(g.event_index, o.event_index), f"event indices differ {what}"
    if g.trace_n is not None:
        assert np.array_equal(g.trace_n, o.trace_n), f"trace_n differs {what}"
        for i in range(len(g.trace_n)):
            n = int(g.trace_n[i])
            assert np.array_equal(g.trace_t[i, :n], o.trace_t[i, :n]), f"trace t differs (member {i}) {what}"
            assert np.array_equal(g.trace_p[i, :n], o.trace_p[i, :n], equal_nan=True), f"trace p differs {what}"


def ode_system(rng, n):
    """(system, params, y0, t0) of a random transcendental-free ODE system."""
    k = rng.integers(0, 6)
    if k == 0:
        rho = rng.uniform(5.0, 40.0, n)
        return d.systems.Lorenz, np.stack([np.full(n, 10.0), rho, np.full(n, 8.0 / 3.0)], 1), \
            rng.uniform(-5, 5, (n, 3)) + [0, 0, 20], 0.0
    if k == 1:
        return d.systems.Harmonic, None, rng.uniform(-2, 2, (n, 2)), float(rng.uniform(-1, 1))
    if k == 2:
        return d.systems.Linear1, rng.uniform(-2, 1, (n, 1)), rng.uniform(0.5, 2, (n, 1)), 0.0
    if k == 3:
        return d.systems.Linear3, rng.uniform(-1, 1, (n, 9)), rng.uniform(-1, 1, (n, 3)), 0.0
    if k == 4:
        return d.systems.Time3, None, rng.uniform(-1, 1, (n, 3)), float(rng.uniform(0, 2))
    return d.systems.Linear3Matrix, rng.uniform(-1, 1, (n, 9)), np.tile(np.eye(3).reshape(1, 9), (n, 1)), 0.0


@pytest.mark.parametrize("seed", range(24 * SCALE))
def test_fuzz_fixed_step_ode_strict_bit_exact(oracle, seed):
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.integers(1, 130))
    system, prm, y0, t0 = ode_system(rng, n)
    rk, rk_name = pick_rk(rng)
    h = float(rng.choice([0.5, 0.1, 0.03, 1.0 / 64, 0.0123456789]))
    t1 = t0 + float(rng.uniform(0.0, 40.0)) * h * (1 if rng.random() < 0.9 else 0)   # sometimes an empty interval
    trace = d.Trace(int(rng.integers(1, 5)), 64) if rng.random() < 0.4 else None
    periodic = rng.random() < 0.35 and system in (d.systems.Linear1, d.systems.Linear3Matrix)

    def mk():
        s = d.Solver.new().initial(y0).interval(t0, t1).stepsize(h).rk(rk).arith("strict")
        s = s.equation(system, prm) if prm is not None else s.equation(system)
        if trace is not None:
            s = s.on_step(trace)
        if periodic:
            cb = "renormalize" if system is d.systems.Linear1 else "qr_ln_abs"
            s = s.on_mut(d.Periodic(float(rng_period), float(rng_offset)), cb)
            if two_periodics:   # coinciding times: earliest wins, lowest index on ties, DedupLocF per locator
                s = s.on_mut(d.Periodic(float(rng_period) * 1.5, 0.0), cb)
        return s
    rng_period, rng_offset = rng.uniform(2.0, 9.0) * h, rng.uniform(0.0, 1.0) * h
    two_periodics = rng.random() < 0.5
    what = f"[seed {seed}: {system.name} n={n} rk={rk_name} h={h} t=[{t0},{t1}] trace={trace is not None} periodic={periodic}]"
    g = mk().run()
    o = oracle.run_solver(mk(), n_threads=4)
    SEEN.add(g.totals["kernel_kind"])
    if system is d.systems.Linear3Matrix and periodic:
        # the QR callback takes log(): compare the log sums to 1-2 ulp, everything else exactly
        assert np.max(np.abs(g.aux - o.aux)) <= 1e-12 * (1.0 + np.max(np.abs(o.aux))), what
        g.aux = o.aux
    if system is d.systems.Linear1 and periodic:
        assert np.max(np.abs(g.aux - o.aux)) <= 1e-12 * (1.0 + np.max(np.abs(o.aux))), what
        g.aux = o.aux
    same(g, o, what)
    if periodic:
        # the same problem in FAST arithmetic (pre-scaled / folded stages, predicted Periodic events): the event
        # protocol is time-grid arithmetic, so every counter and the final time are the oracle's; the state and
        # the accumulated logarithms within the FAST tolerance (linear systems: no chaotic amplification)
        f = mk().arith("fast").run()
        assert f.totals["kernel_kind"] == g.totals["kernel_kind"], what
        for name in ("steps", "events", "rhs_evals", "status"):
            assert np.array_equal(getattr(f, name), getattr(o, name)), f"{name} differs {what} fast"
        assert np.array_equal(f.t, o.t), what
        scale = np.maximum(np.abs(o.p), 1.0)
        assert np.max(np.abs(f.p - o.p) / scale) < 1e-10, what + " fast"
        if o.aux is not None and np.size(o.aux):
            assert np.max(np.abs(f.aux - o.aux)) < 1e-9 * (1.0 + np.max(np.abs(o.aux))), what + " fast"


DETECTIONS = [abi.DET_ZERO, abi.DET_ABOVE_ZERO, abi.DET_BELOW_ZERO, abi.DET_POSITIVE, abi.DET_NEGATIVE,
              abi.DET_SWITCH, abi.DET_SWITCH_TRUE, abi.DET_SWITCH_FALSE, abi.DET_IS_TRUE, abi.DET_IS_FALSE,
              abi.DET_STEP]
LOCATIONS = [abi.LOCATE_STEP_BEGIN, abi.LOCATE_STEP_END, abi.LOCATE_STEP_MIDDLE, abi.LOCATE_LERP,
             abi.LOCATE_BISECTION, abi.LOCATE_BISECTION_BOOL, abi.LOCATE_REGULA_FALSI]


@pytest.mark.parametrize("seed", range(40 * SCALE))
def test_fuzz_located_events_strict_bit_exact(oracle, seed):
    rng = np.random.default_rng(5000 + seed)
    n = int(rng.integers(1, 100))
    which = rng.integers(0, 3)
    rk, rk_name = (d.RK.rktp64(), "rktp64") if rng.random() < 0.6 else \
        ((d.RK.rk43(), "rk43") if rng.random() < 0.5 else (d.RK.rk4(), "rk4"))
    n_extra = int(rng.integers(0, 3))

    def random_locator(det_name):
        det = int(rng.choice(DETECTIONS))
        # a boolean-valued detection goes with the boolean bisection, a sign detection with the real ones
        boolean = det in (abi.DET_SWITCH, abi.DET_SWITCH_TRUE, abi.DET_SWITCH_FALSE, abi.DET_IS_TRUE, abi.DET_IS_FALSE)
        if boolean:
            loc = int(rng.choice([abi.LOCATE_STEP_BEGIN, abi.LOCATE_STEP_END, abi.LOCATE_BISECTION_BOOL]))
        elif det in (abi.DET_POSITIVE, abi.DET_NEGATIVE, abi.DET_STEP):
            loc = int(rng.choice([abi.LOCATE_STEP_BEGIN, abi.LOCATE_STEP_END]))
        else:
            loc = int(rng.choice([abi.LOCATE_STEP_BEGIN, abi.LOCATE_STEP_END, abi.LOCATE_LERP,
                                  abi.LOCATE_BISECTION, abi.LOCATE_REGULA_FALSI]))
        L = d.Locator._mk(det_name, det, loc)
        f = rng.integers(0, 6)
        if f == 1:
            L = L.every(int(rng.integers(1, 4)))
        elif f == 2:
            L = L.separated_by(float(rng.uniform(0.1, 2.0)))
        elif f == 3:
            a = float(rng.uniform(0.0, 3.0))
            L = L.in_interval(a, a + float(rng.uniform(0.5, 4.0)))
        elif f == 4:
            L = L.on_times(sorted(set(int(x) for x in rng.integers(0, 12, 4))))
        return L, (det, loc, int(f))

    descr = []
    if which == 0:      # bouncing ball with its own two locators + random extras
        y0 = np.stack([rng.uniform(0.5, 2.0, n), rng.uniform(-1.0, 1.0, n)], 1)
        # restitution >= 0.85 and t <= 2 keep the run short of the Zeno time (t_1 + (2 v_1 / g) k / (1 - k)
        # >= 3.9 for x0 >= 0.5), where the bounces accumulate and the reference itself never finishes
        prm = np.tile([9.8, float(rng.uniform(0.85, 0.95))], (n, 1))
        s0 = lambda: (d.Solver.new().initial(y0).equation(d.systems.BouncingBall, prm).interval(0.0, 2.0)
                      .stepsize(float(h)).on(d.Locator.zero("dx"), None).on_mut(d.Locator.below_zero("x"), "bounce"))
        h = rng.choice([0.005, 0.02, 0.05])
        names = ["dx", "x"]
    elif which == 1:    # relay with frozen sign
        y0 = np.stack([rng.uniform(0.1, 1.0, n), rng.uniform(-0.5, 0.5, n)], 1)
        s0 = lambda: (d.Solver.new().initial(y0).equation(d.systems.RelayMsign).interval(0.5, 8.5)
                      .stepsize(float(h)).on(d.Locator.zero("x"), None))
        h = rng.choice([0.09, 0.03, 0.2])
        names = ["x"]
    else:               # egg billiard
        phi = rng.uniform(0, 2 * np.pi, n)
        y0 = np.stack([np.zeros(n), np.full(n, 0.1), 0.5 * np.cos(phi), 0.5 * np.sin(phi)], 1)
        prm = np.full((n, 1), float(rng.integers(3, 40)))
        s0 = lambda: (d.Solver.new().initial(y0).equation(d.systems.EggBilliard, prm).interval(0.0, 60.0)
                      .stepsize(float(h)).on_mut(d.Locator.above_zero("boundary"), "reflect"))
        h = rng.choice([0.5, 0.25, 0.7])
        names = ["boundary"]
    extras = []
    for _ in range(n_extra):
        if rng.random() < 0.3:
            extras.append((d.Periodic(float(rng.uniform(0.2, 1.5)), float(rng.uniform(0.0, 0.3))), ("periodic",)))
        else:
            extras.append(random_locator(str(rng.choice(names))))
    use_trace = rng.random() < 0.5

    def mk():
        s = s0().rk(rk).arith("strict").log_events(d.EventLog(96)).max_steps(6000)   # guard: never spin
        for L, _ in extras:
            s = s.on(L, None)
        if use_trace:
            s = s.on_step(d.Trace(1, 1024))
        return s
    what = f"[seed {seed}: case {which} n={n} rk={rk_name} h={h} extras={[e[1] for e in extras]} trace={use_trace}]"
    g = mk().run()
    o = oracle.run_solver(mk(), n_threads=4)
    SEEN.add(g.totals["kernel_kind"])
    same(g, o, what)


@pytest.mark.parametrize("seed", range(12 * SCALE))
def test_fuzz_dde_strict_vs_oracle(oracle, seed):
    # the sin history / exact solution put sin() on both sides: compare to 1e-13, counters exactly
    rng = np.random.default_rng(9000 + seed)
    n = int(rng.integers(1, 80))
    k = rng.uniform(0.4, 1.4, n)
    tau = float(rng.choice([1.0, 0.5, 0.73]))
    h = float(rng.choice([0.02, 0.05, 0.011]))
    per_member = rng.random() < 0.3      # a different delay per member: the lanes of a warp disagree on the window
    taus = rng.choice([1.0, 0.5, 0.73, 0.25], size=n) if per_member else np.full(n, tau)
    tau = float(np.max(taus))
    rk, rk_name = [(d.RK.rktp64(), "rktp64"), (d.RK.rk43(), "rk43"), (d.RK.rk4(), "rk4")][rng.integers(0, 3)]
    neutral = rng.random() < 0.4
    import cases
    if neutral:
        prm, system, init = cases.ndde_params(k, taus), d.systems.NddeLinear, d.InitFn(d.SIN, d.SIN_DERIV)
    else:
        prm, system, init = cases.dde_params(k, taus), d.systems.DdeLinear, d.InitFn(d.SIN)
    t1 = float(rng.uniform(1.0, 6.0))
    for arith, tol in (("strict", 1e-13), ("fast", 1e-11)):
        def mk():
            return (d.Solver.new().max_delay(tau).initial(init).interval(0.0, t1).stepsize(h).rk(rk)
                    .equation(system, prm).arith(arith))
        g = mk().run()
        o = oracle.run_solver(mk(), n_threads=4)
        SEEN.add(g.totals["kernel_kind"])
        what = f"[seed {seed}: {system.name} n={n} tau={tau} per_member={per_member} h={h} rk={rk_name} t1={t1} {arith}]"
        assert np.array_equal(g.steps, o.steps) and np.array_equal(g.status, o.status), what
        assert np.array_equal(g.t, o.t), what
        assert np.max(np.abs(g.p - o.p)) < tol * max(1.0, float(np.max(np.abs(o.p)))), what


@pytest.mark.parametrize("seed", range(10 * SCALE))
def test_fuzz_adaptive_vs_oracle(oracle, seed):
    # AutomaticStepsize: STRICT (general kernel, libm-like pow on both sides, <= 2 ulp apart) and FAST
    # (ode_adaptive_kernel, Newton root) against the oracle: same step sequence up to one step near
    # err = 1, states within the controller's tolerance
    rng = np.random.default_rng(7000 + seed)
    n = int(rng.integers(1, 150))
    system, prm, y0, t0 = ode_system(rng, n)
    if system is d.systems.Time3:
        system, prm, y0, t0 = d.systems.Harmonic, None, rng.uniform(-2, 2, (n, 2)), 0.0
    N = y0.shape[1]
    tol = float(rng.choice([1e-6, 1e-8, 1e-10]))
    rk, rk_name, order = [(d.RK.rktp64(), "rktp64", 4), (d.RK.rk43(), "rk43", 3), (d.RK.rk4(), "rk4", 2)][rng.integers(0, 3)]
    ctl = d.AutomaticStepsize(stepsize=float(rng.choice([1e-2, 1e-3])), stepsize_range=(1e-9, 0.25), atol=[tol] * N,
                              rtol=[tol] * N, order=order, fac=float(rng.uniform(0.8, 0.95)), fac_range=(0.2, 5.0))
    t1 = t0 + float(rng.uniform(0.5, 3.0))

    def mk(arith):
        s = d.Solver.new().initial(y0).interval(t0, t1).stepsize(ctl).rk(rk).arith(arith)
        return s.equation(system, prm) if prm is not None else s.equation(system)
    o = oracle.run_solver(mk("strict"), n_threads=4)
    what = f"[seed {seed}: {system.name} n={n} rk={rk_name} tol={tol} t=[{t0},{t1}]]"
    for arith in ("strict", "fast"):
        g = mk(arith).run()
        SEEN.add(g.totals["kernel_kind"])
        assert np.array_equal(g.status, o.status), what
        ok = o.status == 0
        assert np.max(np.abs(g.steps[ok].astype(np.int64) - o.steps[ok].astype(np.int64)), initial=0) <= 1, what
        assert np.array_equal(g.t[ok], o.t[ok]), what
        scale = np.maximum(np.abs(o.p[ok]), 1.0)
        assert np.max(np.abs(g.p[ok] - o.p[ok]) / scale, initial=0.0) < 100 * tol, what


@pytest.mark.parametrize("seed", range(16 * SCALE))
def test_fuzz_dde_events_and_propagation_strict_bit_exact(oracle, seed):
    # delay systems without transcendental calls, through the general kernel: the relay-clamp DDE
    # (examples/dde-relay-clamp.rs) and y' = 0 with propagated discontinuities
    # (tests/delay_propagation.rs), random delays / steps / tableaux / smoothing orders
    rng = np.random.default_rng(11000 + seed)
    n = int(rng.integers(1, 70))
    rk, rk_name = [(d.RK.rktp64(), "rktp64"), (d.RK.rk43(), "rk43"), (d.RK.rk4(), "rk4"), (d.RK.euler(), "euler"),
                   (d.RK.heun3(), "heun3")][rng.integers(0, 5)]
    if rng.random() < 0.5:
        T_ = float(rng.choice([1.0, 0.7, 1.3]))
        alpha = rng.uniform(2.0, 7.0, n)
        prm = np.stack([alpha, np.full(n, float(rng.uniform(0.1, 0.8))), np.full(n, T_)], 1)
        y0 = np.stack([rng.uniform(0.5, 1.5, n), -alpha * rng.uniform(0.5, 1.5, n)], 1)
        h = float(rng.choice([0.012, 0.05, 0.1])) * T_
        smoothing = int(rng.integers(0, 3))

        def mk():
            s = (d.Solver.new().initial(y0).equation(d.systems.DdeRelayClamp, prm).interval(0.0, 8.0 * T_)
                 .stepsize(h).rk(rk).with_const_delay(T_, smoothing).arith("strict").max_steps(5000)
                 .log_events(d.EventLog(128)))
            if rng_use_detector:
                s = s.on(d.Locator.zero("abs_delayed_x_minus_1"), None)
            return s
        rng_use_detector = rng.random() < 0.7
        what = f"[seed {seed}: relay clamp n={n} T={T_} h={h} rk={rk_name} smoothing={smoothing} det={rng_use_detector}]"
    else:
        delays = sorted(set(float(x) for x in rng.choice([0.5, 1.0, 1.25, 0.8], size=int(rng.integers(1, 3)))))
        h = float(rng.choice([0.75, 7.0 / 8.0, 0.3, 0.123456789]))
        discos = [(0.0, int(rng.integers(0, 2)))]
        smoothing = int(rng.integers(0, 2))
        pantograph = rng.random() < 0.3

        def mk():
            s = (d.Solver.new().rk(rk).initial(np.zeros((n, 1))).equation(d.systems.Zero1)
                 .stepsize(h).arith("strict").max_steps(5000).on_step(d.Trace(1, 256)))
            if pantograph:
                return (s.initial_disco([(1.0, 0)]).interval(1.0, 40.0).with_delayed_argument("half_time", smoothing)
                        .max_delay(math.inf))
            s = s.initial_disco(discos).interval(0.0, 6.0)
            for dl in delays:
                s = s.with_const_delay(dl, smoothing)
            return s
        what = f"[seed {seed}: zero1 n={n} delays={delays} h={h} rk={rk_name} smoothing={smoothing} pantograph={pantograph}]"
    g = mk().run()
    o = oracle.run_solver(mk(), n_threads=4)
    SEEN.add(g.totals["kernel_kind"])
    same(g, o, what)
    # the same problem in FAST arithmetic: dde_event_kernel where a (system, tableau shape) instantiation
    # exists, the general kernel otherwise.  y' = 0 makes the time grid pure time arithmetic (FAST = STRICT
    # bit for bit); the relay-clamp DDE must keep the event sequence and stay within the FAST tolerances
    f = mk().arith("fast").run()
    SEEN.add(f.totals["kernel_kind"])
    compiled = ("rk4", "euler") if "zero1" in what else ("rktp64", "rk4")   # kernels_dde_event.cu shape_compiled
    if rk_name in compiled:
        assert f.totals["kernel_kind"] == abi.KERNEL_DDE_EVENT, what
    if "zero1" in what:
        same(f, o, what + " fast")
    else:
        assert np.array_equal(f.status, o.status), what
        ok = o.status == 0
        assert np.array_equal(f.steps[ok], o.steps[ok]) and np.array_equal(f.events[ok], o.events[ok]), what
        assert np.array_equal(f.event_n[ok], o.event_n[ok]) and np.array_equal(f.event_index[ok], o.event_index[ok]), what
        assert np.allclose(f.event_t[ok], o.event_t[ok], rtol=0.0, atol=1e-9, equal_nan=True), what
        assert np.allclose(f.p[ok], o.p[ok], rtol=1e-8, atol=1e-8), what


@pytest.mark.parametrize("seed", range(10 * SCALE))
def test_fuzz_output_policies_and_ragged_sizes(oracle, seed):
    # traces with `every` > 1 and a capacity smaller than the number of samples, event logs that
    # overflow their capacity (the count keeps running), member counts around the warp / block edges
    rng = np.random.default_rng(13000 + seed)
    n = int(rng.choice([1, 2, 31, 32, 33, 63, 64, 65, 127, 128, 129, 255, 257]))
    every, cap, ecap = int(rng.integers(1, 7)), int(rng.integers(1, 40)), int(rng.integers(1, 5))
    y0 = np.stack([rng.uniform(0.5, 2.0, n), rng.uniform(-1.0, 1.0, n)], 1)
    prm = np.tile([9.8, 0.9], (n, 1))

    def mk():
        return (d.Solver.new().initial(y0).equation(d.systems.BouncingBall, prm).interval(0.0, 2.0)
                .stepsize(0.01).on(d.Locator.zero("dx"), None).on_mut(d.Locator.below_zero("x"), "bounce")
                .on_step(d.Trace(every, cap)).log_events(d.EventLog(ecap)).arith("strict").max_steps(3000))
    g = mk().run()
    o = oracle.run_solver(mk(), n_threads=4)
    same(g, o, f"[seed {seed}: n={n} every={every} cap={cap} ecap={ecap}]")
    assert np.all(g.trace_n <= cap)


@pytest.mark.parametrize("seed", range(16 * SCALE))
def test_fuzz_fast_arithmetic_within_tolerance(oracle, seed):
    # FAST (FMA, pre-scaled rows, pre-combined interpolants) against the STRICT oracle on non-chaotic
    # systems: north_star's 1e-12 relative for fixed-step trajectories, event times within 1e-10
    rng = np.random.default_rng(17000 + seed)
    n = int(rng.integers(1, 200))
    rk, rk_name = pick_rk(rng)
    if rng.random() < 0.6:
        k = rng.integers(0, 3)
        if k == 0:
            system, prm, y0 = d.systems.Harmonic, None, rng.uniform(-2, 2, (n, 2))
        elif k == 1:
            system, prm, y0 = d.systems.Linear1, rng.uniform(-2, 0.5, (n, 1)), rng.uniform(0.5, 2, (n, 1))
        else:
            system, prm, y0 = d.systems.Time3, None, rng.uniform(-1, 1, (n, 3))
        h = float(rng.choice([0.1, 0.03, 1.0 / 64]))
        t1 = float(rng.uniform(5.0, 60.0)) * h

        def mk(arith):
            s = d.Solver.new().initial(y0).interval(0.0, t1).stepsize(h).rk(rk).arith(arith)
            return s.equation(system, prm) if prm is not None else s.equation(system)
        g = mk("fast").run()
        o = oracle.run_solver(mk("strict"), n_threads=4)
        what = f"[seed {seed}: {system.name} n={n} rk={rk_name} h={h} t1={t1}]"
        assert np.array_equal(g.steps, o.steps) and np.array_equal(g.t, o.t) and np.all(g.status == 0), what
        assert np.max(np.abs(g.p - o.p) / np.maximum(np.abs(o.p), 1.0)) < 1e-12, what
    else:
        y0 = np.stack([rng.uniform(0.5, 2.0, n), rng.uniform(-1.0, 1.0, n)], 1)
        prm = np.tile([9.8, float(rng.uniform(0.85, 0.95))], (n, 1))
        h = float(rng.choice([0.005, 0.01, 0.02]))
        rk43 = rng.random() < 0.3

        def mk(arith):
            s = (d.Solver.new().initial(y0).equation(d.systems.BouncingBall, prm).interval(0.0, 2.0).stepsize(h)
                 .on(d.Locator.zero("dx"), None).on_mut(d.Locator.below_zero("x"), "bounce")
                 .log_events(d.EventLog(64)).arith(arith).max_steps(5000))
            return s.rk(d.RK.rk43()) if rk43 else s
        g = mk("fast").run()
        o = oracle.run_solver(mk("strict"), n_threads=4)
        what = f"[seed {seed}: bouncing ball n={n} h={h} rk43={rk43}]"
        assert g.totals["kernel_kind"] == abi.KERNEL_ODE_EVENT
        assert np.array_equal(g.event_n, o.event_n) and np.array_equal(g.event_index, o.event_index), what
        assert np.array_equal(g.steps, o.steps), what
        assert np.max(np.abs(g.event_t - o.event_t)) < 1e-10, what
        assert np.max(np.abs(g.p - o.p)) < 1e-9, what


@pytest.mark.parametrize("seed", range(12 * SCALE))
def test_fuzz_traced_hot_kernels_fast_vs_oracle(oracle, seed):
    # Trace(every, capacity) inside the fixed-step, DDE (one and two members per thread) and adaptive hot
    # kernels, FAST arithmetic, against the STRICT oracle: sample counts and the shared time grid bit for
    # bit, samples within the FAST tolerance
    rng = np.random.default_rng(21000 + seed)
    n = int(rng.choice([1, 31, 32, 33, 64, 65, 127, 129, 200, 257]))
    every, cap = int(rng.integers(1, 9)), int(rng.choice([1, 2, 7, 40, 400, 4000]))
    which = int(rng.integers(0, 3))
    import cases
    if which == 0:      # fixed-step ODE: traceable instantiations are rktp64 and rk4
        rk, rk_name = [(d.RK.rktp64(), "rktp64"), (d.RK.rk4(), "rk4")][rng.integers(0, 2)]
        system, prm, y0 = [(d.systems.Harmonic, None, rng.uniform(-2, 2, (n, 2))),
                           (d.systems.Linear1, rng.uniform(-2, 0.5, (n, 1)), rng.uniform(0.5, 2, (n, 1)))][rng.integers(0, 2)]
        h = float(rng.choice([0.1, 0.03, 1.0 / 64]))
        t1 = float(rng.uniform(5.0, 300.0)) * h

        def mk(arith):
            s = d.Solver.new().initial(y0).interval(0.0, t1).stepsize(h).rk(rk).arith(arith).on_step(d.Trace(every, cap))
            return s.equation(system, prm) if prm is not None else s.equation(system)
        kind, tol = abi.KERNEL_ODE_FIXED, 1e-12
        what = f"[seed {seed}: {system.name} n={n} rk={rk_name} h={h} t1={t1} every={every} cap={cap}]"
    elif which == 1:    # constant-delay DDE / NDDE
        k = rng.uniform(0.4, 1.4, n)
        tau = float(rng.choice([1.0, 0.5, 0.73]))
        h = float(rng.choice([0.02, 0.05, 0.011]))
        rk, rk_name = [(d.RK.rktp64(), "rktp64"), (d.RK.rk43(), "rk43"), (d.RK.rk4(), "rk4")][rng.integers(0, 3)]
        if rng.random() < 0.4:
            prm, system, init = cases.ndde_params(k, tau), d.systems.NddeLinear, d.InitFn(d.SIN, d.SIN_DERIV)
        else:
            prm, system, init = cases.dde_params(k, tau), d.systems.DdeLinear, d.InitFn(d.SIN)
        t1 = float(rng.uniform(0.5, 6.0))
        tpt = str(rng.integers(1, 3))

        def mk(arith):
            return (d.Solver.new().max_delay(tau).initial(init).interval(0.0, t1).stepsize(h).rk(rk)
                    .equation(system, prm).arith(arith).on_step(d.Trace(every, cap)))
        kind, tol = abi.KERNEL_DDE_FIXED, 1e-11
        what = f"[seed {seed}: {system.name} n={n} tau={tau} h={h} rk={rk_name} t1={t1} every={every} cap={cap} tpt={tpt}]"
        os.environ["DFB_DDE_TPT"] = tpt
    else:               # adaptive: every member has its own grid; consistency + exact solution
        amp = rng.uniform(0.5, 3.0, n)
        y0 = np.stack([np.zeros(n), amp], 1)
        ctl = d.AutomaticStepsize(1e-2, (1e-6, 1e-1), [1e-9] * 2, [1e-9] * 2, 4, 0.9, (0.2, 5.0))
        t1 = float(rng.uniform(0.5, 8.0))
        f = (d.Solver.new().initial(y0).equation(d.systems.Harmonic).interval(0.0, t1).stepsize(ctl)
             .arith("fast").on_step(d.Trace(every, cap)).run())
        what = f"[seed {seed}: adaptive harmonic n={n} t1={t1} every={every} cap={cap}]"
        assert f.totals["kernel_kind"] == abi.KERNEL_ODE_ADAPTIVE and np.all(f.status == 0), what
        assert np.array_equal(f.trace_n, np.minimum(f.steps // every + 1, cap)), what
        for i in range(0, n, 7):
            m = int(f.trace_n[i])
            t = f.trace_t[i, :m]
            assert t[0] == 0.0 and np.all(np.diff(t) > 0) and t[-1] <= t1, what
            assert np.max(np.abs(f.trace_p[i, :m, 0] - amp[i] * np.sin(t))) < 1e-6, what
        return
    try:
        g = mk("fast").run()
    finally:
        os.environ.pop("DFB_DDE_TPT", None)
    o = oracle.run_solver(mk("strict"), n_threads=4)
    assert g.totals["kernel_kind"] == kind and np.all(g.status == 0), what
    assert np.array_equal(g.steps, o.steps) and np.array_equal(g.t, o.t), what
    assert np.array_equal(g.trace_n, o.trace_n) and np.all(g.trace_n <= cap), what
    m = int(g.trace_n[0])
    assert np.all(g.trace_n == m), what
    assert np.array_equal(g.trace_t[:, :m], o.trace_t[:, :m]), what
    scale = np.maximum(np.abs(o.trace_p[:, :m]), 1.0)
    assert np.max(np.abs(g.trace_p[:, :m] - o.trace_p[:, :m]) / scale) < tol, what
    assert np.max(np.abs(g.p - o.p) / np.maximum(np.abs(o.p), 1.0)) < tol, what


def test_fuzz_reached_every_ode_kernel_family():
    # runs last in this file: the random configurations above must have exercised these kernels
    if not SEEN:
        pytest.skip("fuzz cases did not run in this session")
    assert {abi.KERNEL_ODE_FIXED, abi.KERNEL_ODE_EVENT, abi.KERNEL_GENERAL, abi.KERNEL_ODE_ADAPTIVE,
            abi.KERNEL_DDE_FIXED, abi.KERNEL_DDE_EVENT}.issubset(SEEN), SEEN
