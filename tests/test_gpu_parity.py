"""Parity of the CUDA path (through the C ABI) against the CPU oracle.

STRICT arithmetic: bit-exact for every system without transcendental calls
(integer/branch decisions, step counts, event times, final states, traces).
FAST arithmetic (FMA, pre-scaled rows): within 1e-12 relative on non-chaotic
horizons, as BASELINE.json's north_star states.  Systems that call sin/cos/log
on the device (CUDA libm is within 1-2 ulp of glibc, not bit-equal) are compared
to a tolerance written at each test.
"""
import math
import os

import numpy as np
import pytest

import cases
import diffurch_b200 as d

pytestmark = pytest.mark.gpu


def run_both(oracle, make, **kw):
    g = make(**kw).run()
    o = oracle.run_solver(make(**kw), n_threads=8)
    return g, o


KERNEL_GENERAL, KERNEL_ODE_FIXED, KERNEL_DDE_FIXED, KERNEL_ODE_ADAPTIVE = 0, 1, 2, 3
def run_general(make):
    os.environ["DFB_FORCE_GENERAL"] = "1"
    try:
        r = make().run()
    finally:
        del os.environ["DFB_FORCE_GENERAL"]
    assert r.totals["kernel_kind"] == KERNEL_GENERAL
    return r


def assert_same_trace(g, o, exact=True, rtol=0.0, atol=0.0):
    assert np.array_equal(g.trace_n, o.trace_n)
    for i in range(len(g.trace_n)):
        n = int(g.trace_n[i])
        if exact:
            assert np.array_equal(g.trace_t[i, :n], o.trace_t[i, :n])
            assert np.array_equal(g.trace_p[i, :n], o.trace_p[i, :n])
        else:
            assert np.allclose(g.trace_t[i, :n], o.trace_t[i, :n], rtol=rtol, atol=atol)
            assert np.allclose(g.trace_p[i, :n], o.trace_p[i, :n], rtol=rtol, atol=atol)


def assert_same_counters(g, o):
    assert np.array_equal(g.steps, o.steps)
    assert np.array_equal(g.rejected, o.rejected)
    assert np.array_equal(g.events, o.events)
    assert np.array_equal(g.rhs_evals, o.rhs_evals)
    assert np.array_equal(g.status, o.status)


# ---------------------------------------------------------------- golden vectors on the GPU
def test_gpu_delay_propagation_golden_vectors():
    # tests/delay_propagation.rs:21-24, 45-50, 68-71, 91-93 — exact f64 equality
    g = cases.delay_constant_1().run()
    assert list(g.trace_t[0, :g.trace_n[0]]) == [0.0, 0.75, 1.0, 1.75, 2.0, 2.75, 3.0, 3.75, 4.0, 4.75, 5.0]
    g = cases.delay_constant_2().run()
    s = 7.0 / 8.0
    assert list(g.trace_t[0, :g.trace_n[0]]) == [0.0, s, 1.0, 1.25, 2., 2.25, 2.5, 3.0, 3.25, 3.5,
                                                  3.75, 4., 4.25, 4.5, 4.75, 5.0]
    g = cases.delay_constant_1_smoothing().run()
    assert list(g.trace_t[0, :g.trace_n[0]]) == [0., 0.75, 1., 1.75, 2., 2.75, 3., 3.75, 4.5, 5.25, 6.]
    g = cases.delay_pantograph().run()
    ts = set(g.trace_t[0, :g.trace_n[0]])
    for i in range(1, 11):
        assert 2.0 ** i in ts


def test_gpu_equation_types_golden():
    # tests/equation_types.rs:15, 38-40: exact; :77, :94, :110: reference tolerances
    g = cases.eq_constant().run()
    for t, p in zip(g.trace_t[0, :g.trace_n[0]], g.trace_p[0]):
        assert list(p) == [t, -t, 0.0]
    g = cases.eq_time().run()
    for t, p in zip(g.trace_t[0, :g.trace_n[0]], g.trace_p[0]):
        assert list(p) == [0.0, t, t * t / 2.0]
    for arith in ("strict", "fast"):
        g = cases.eq_ode_exponent().arith(arith).run()
        n = g.trace_n[0]
        assert np.all(np.abs(g.trace_p[0, :n, 0] - np.exp(-g.trace_t[0, :n])) < 1e-14)
        g = cases.eq_ode_harmonic().arith(arith).run()
        n = g.trace_n[0]
        assert np.all(np.abs(g.trace_p[0, :n, 0] - np.sin(g.trace_t[0, :n])) < 1e-13)
        g = cases.eq_odet_lin().arith(arith).run()
        n = g.trace_n[0]
        assert np.all(np.abs(g.trace_p[0, :n, 0] - g.trace_t[0, :n] ** -2.0) < 1e-13)


# ---------------------------------------------------------------- strict = bit-exact vs the oracle
@pytest.mark.parametrize("name", ["delay_constant_1", "delay_constant_2", "delay_constant_1_smoothing",
                                  "delay_pantograph", "eq_constant", "eq_time", "eq_ode_exponent",
                                  "eq_ode_harmonic", "eq_odet_lin"])
def test_strict_bit_exact_traces(oracle, name):
    g, o = run_both(oracle, getattr(cases, name))
    assert_same_counters(g, o)
    assert_same_trace(g, o, exact=True)
    assert np.array_equal(g.t, o.t) and np.array_equal(g.p, o.p)


@pytest.mark.parametrize("rk", ["rktp64", "rk4", "rk43", "three_eights", "euler", "midpoint", "heun2",
                                "ralston2", "kutta3", "heun3", "ralston3", "wray3", "ssp3"])
def test_lorenz_strict_bit_exact_all_tableaux(oracle, rk):
    # both the hot-path kernel and the oracle, 257 trajectories (ragged last block), t in [0, 5]
    mk = lambda: cases.lorenz(n=257, t1=5.0, rk=d.RK.by_name(rk))
    g, o = run_both(oracle, mk)
    assert_same_counters(g, o)
    assert np.array_equal(g.t, o.t)
    assert np.array_equal(g.p, o.p)


def test_lorenz_hot_path_equals_general_path(oracle):
    mk = lambda: cases.lorenz(n=300, t1=3.0)
    hot = mk().run()
    os.environ["DFB_FORCE_GENERAL"] = "1"
    try:
        gen = mk().run()
    finally:
        del os.environ["DFB_FORCE_GENERAL"]
    assert np.array_equal(hot.p, gen.p) and np.array_equal(hot.t, gen.t)
    assert np.array_equal(hot.steps, gen.steps) and np.array_equal(hot.rhs_evals, gen.rhs_evals)
    # a traced run (rktp64: the hot kernel's TRACE instantiation); its last sample is the final state
    tr = cases.lorenz(n=300, t1=3.0, trace=d.Trace(1, 800)).run()
    assert np.array_equal(tr.p, hot.p)
    n = tr.trace_n[0]
    assert n == hot.steps[0] + 1
    assert np.array_equal(tr.trace_p[:, n - 1, :], hot.p)


def test_lorenz_fast_within_1e12_short_horizon(oracle):
    # north_star: fixed-step trajectories within 1e-12 relative; chaotic systems on short
    # horizons.  Rounding-level differences grow along the flow (rho up to 40 and a start far
    # from the attractor): 1e-12 holds to t = 0.25, and the growth to t = 2 stays below 1e-10.
    for rk in ("rktp64", "rk4"):
        for t1, tol in ((0.25, 1e-12), (2.0, 1e-10)):
            mk = lambda arith: cases.lorenz(n=512, t1=t1, rk=d.RK.by_name(rk), arith=arith)
            g = mk("fast").run()
            o = oracle.run_solver(mk("strict"), n_threads=8)
            assert np.array_equal(g.steps, o.steps)
            rel = np.max(np.abs(g.p - o.p) / np.maximum(np.abs(o.p), 1.0))
            assert rel < tol, (rk, t1, rel)


def test_linear_systems_fast_within_1e12(oracle):
    # non-chaotic: harmonic oscillator and exponential decay over the reference's horizons
    for mk in (cases.eq_ode_harmonic, cases.eq_ode_exponent):
        g = mk().arith("fast").run()
        o = oracle.run_solver(mk())
        n = g.trace_n[0]
        assert np.array_equal(g.trace_t[0, :n], o.trace_t[0, :n])
        rel = np.max(np.abs(g.trace_p[0, :n] - o.trace_p[0, :n]) / np.maximum(np.abs(o.trace_p[0, :n]), 1e-3))
        assert rel < 1e-12, rel


def test_fibonacci_periodic_events(oracle):
    # examples/ode-linear-3-fibonacci.rs: Periodic(1) fires at t = 1..50, x(n) = F_n
    g, o = run_both(oracle, cases.fibonacci)
    assert_same_counters(g, o)
    assert_same_trace(g, o, exact=True)
    assert np.array_equal(g.event_t, o.event_t)
    ne = g.event_n[0]
    assert list(g.event_t[0, :ne]) == [float(i) for i in range(1, 51)]
    # x(10) = 55
    n = g.trace_n[0]
    i10 = list(g.trace_t[0, :n]).index(10.0)
    assert abs(g.trace_p[0, i10, 0] - 55.0) < 1e-6


# ---------------------------------------------------------------- events
def test_bouncing_ball_strict_bit_exact(oracle):
    mk = lambda: cases.bouncing_ball(n=96, t1=8.0, trace=d.Trace(1, 2048))
    g, o = run_both(oracle, mk)
    assert_same_counters(g, o)
    assert np.array_equal(g.event_n, o.event_n)
    assert np.array_equal(g.event_t, o.event_t)          # bisection results, bit for bit
    assert np.array_equal(g.event_index, o.event_index)
    assert_same_trace(g, o, exact=True)
    assert np.array_equal(g.p, o.p)
    # the single-trajectory anchors of SURVEY §8c
    g1 = cases.bouncing_ball().run()
    impacts = g1.event_t[0, :g1.event_n[0]][g1.event_index[0, :g1.event_n[0]] == 1]
    assert len(impacts) == 78 and g1.steps[0] + 1 == 1815
    assert abs(impacts[0] - (-0.01 + math.sqrt(19.6001)) / 9.8) < 2e-15


def test_bouncing_ball_batching_threshold_does_not_change_results():
    # warp-level event batching is a scheduling decision only: per-trajectory results must
    # not depend on the batch threshold or on the bounded wait
    for mk in (lambda: cases.bouncing_ball(n=200, t1=6.0), lambda: cases.egg_billiard(n=100, reflections=40)):
        res = []
        for thr, wait in (("1", "0"), ("8", "4"), ("32", "16"), ("32", "100000")):
            os.environ["DFB_LOCATE_BATCH"], os.environ["DFB_LOCATE_WAIT"] = thr, wait
            try:
                res.append(mk().run())
            finally:
                del os.environ["DFB_LOCATE_BATCH"], os.environ["DFB_LOCATE_WAIT"]
        for r in res[1:]:
            assert np.array_equal(r.p, res[0].p) and np.array_equal(r.event_t, res[0].event_t)
            assert np.array_equal(r.steps, res[0].steps) and np.array_equal(r.events, res[0].events)


def test_bouncing_ball_fast_event_times_within_tolerance(oracle):
    g = cases.bouncing_ball(n=64, t1=6.0, arith="fast").run()
    o = oracle.run_solver(cases.bouncing_ball(n=64, t1=6.0), n_threads=8)
    assert np.array_equal(g.event_n, o.event_n)
    assert np.array_equal(g.event_index, o.event_index)
    assert np.max(np.abs(g.event_t - o.event_t)) < 1e-11
    assert np.max(np.abs(g.p - o.p)) < 1e-10


def test_egg_billiard_strict_bit_exact(oracle):
    mk = lambda: cases.egg_billiard(n=64, reflections=60)
    g, o = run_both(oracle, mk)
    assert_same_counters(g, o)
    assert np.all(g.status == d.abi.TRAJ_STOPPED)
    assert np.all(np.isinf(g.t))
    assert np.array_equal(g.event_t, o.event_t)
    assert np.array_equal(g.p, o.p)
    assert np.array_equal(g.aux, o.aux) and np.all(g.aux[:, 0] == 60.0)
    g1 = cases.egg_billiard(reflections=200).run()
    assert g1.steps[0] + 1 == 1497  # SURVEY §8c anchor


def test_relay_msign_strict_bit_exact(oracle):
    g, o = run_both(oracle, cases.relay_msign)
    assert_same_counters(g, o)
    assert_same_trace(g, o, exact=True)
    assert np.array_equal(g.event_t, o.event_t)
    # analytic solution of examples/ode-2-relay-msign.rs:8-13 at the recorded times
    n = g.trace_n[0]
    t = g.trace_t[0, :n]
    sgn = np.sign((t * 0.5) % 1.0 - 0.5)
    x = (t - np.floor(t)) * (t - np.ceil(t)) * sgn
    assert np.max(np.abs(g.trace_p[0, :n, 0] - x)) < 1e-9


@pytest.mark.parametrize("detection,location", [
    (d.abi.DET_ZERO, d.abi.LOCATE_LERP), (d.abi.DET_ZERO, d.abi.LOCATE_REGULA_FALSI),
    (d.abi.DET_ZERO, d.abi.LOCATE_STEP_BEGIN), (d.abi.DET_ZERO, d.abi.LOCATE_STEP_END),
    (d.abi.DET_ABOVE_ZERO, d.abi.LOCATE_BISECTION), (d.abi.DET_BELOW_ZERO, d.abi.LOCATE_BISECTION),
    (d.abi.DET_SWITCH, d.abi.LOCATE_BISECTION_BOOL), (d.abi.DET_SWITCH_TRUE, d.abi.LOCATE_BISECTION_BOOL),
    (d.abi.DET_SWITCH_FALSE, d.abi.LOCATE_BISECTION_BOOL)])
def test_all_detection_and_location_methods(oracle, detection, location):
    # harmonic-like relay system, every detection/location kind the reference defines
    def mk():
        s = (d.Solver.new().initial([0.25, 0.0]).equation(d.systems.RelayMsign).interval(0.5, 6.5)
             .stepsize(0.09).log_events(d.EventLog(64)).on_step(d.Trace(1, 256)))
        loc = d.Locator._mk("x", detection, location)
        return s.on(loc, None)
    g, o = run_both(oracle, mk)
    assert_same_counters(g, o)
    assert np.array_equal(g.event_t, o.event_t)
    assert_same_trace(g, o, exact=True)


def test_filters_every_separated_in_interval(oracle):
    for flt in (lambda L: L.every(3), lambda L: L.separated_by(2.5), lambda L: L.in_interval(2.0, 5.0),
                lambda L: L.on_times([0, 2, 3, 7])):
        def mk():
            return (d.Solver.new().initial([0.25, 0.0]).equation(d.systems.RelayMsign)
                    .interval(0.5, 10.5).stepsize(0.09).log_events(d.EventLog(64))
                    .on(flt(d.Locator.zero("x")), None).on_step(d.Trace(1, 256)))
        g, o = run_both(oracle, mk)
        assert_same_counters(g, o)
        assert np.array_equal(g.event_t, o.event_t)
        assert g.event_n[0] > 0



# ---------------------------------------------------------------- fixed step + located events (ode_event_kernel)
KERNEL_ODE_EVENT = 5


def _same_everything(a, b):
    for name in ("t", "p", "aux", "steps", "events", "rhs_evals", "status", "rejected"):
        x, y = getattr(a, name), getattr(b, name)
        if x is not None and np.size(x):
            assert np.array_equal(x, y, equal_nan=True), name
    if a.event_n is not None:
        assert np.array_equal(a.event_n, b.event_n)
        assert np.array_equal(a.event_t, b.event_t) and np.array_equal(a.event_index, b.event_index)
    if a.trace_n is not None:
        assert_same_trace(a, b, exact=True)


@pytest.mark.parametrize("case", ["bouncing_ball", "egg_billiard", "relay_msign", "fibonacci"])
def test_event_kernel_strict_equals_general_kernel_and_oracle(oracle, case):
    mk = {"bouncing_ball": lambda: cases.bouncing_ball(n=96, t1=8.0, trace=d.Trace(1, 2048)),
          "egg_billiard": lambda: cases.egg_billiard(n=64, reflections=60),
          "relay_msign": cases.relay_msign, "fibonacci": cases.fibonacci}[case]
    hot = mk().run()
    assert hot.totals["kernel_kind"] == KERNEL_ODE_EVENT
    gen = run_general(mk)
    _same_everything(hot, gen)
    o = oracle.run_solver(mk(), n_threads=8)
    assert_same_counters(hot, o)
    assert np.array_equal(hot.event_t, o.event_t) and np.array_equal(hot.p, o.p)


def test_event_kernel_compile_time_locators_match_run_time_descriptors(oracle):
    # FAST builds of the shipped event examples know their locators at compile time (SPEC: detector id,
    # detection kind, DedupLocF folded into the kernel); DFB_EVENT_RUNTIME_SPEC=1 runs the same problems with the
    # descriptors read per step.  Same arithmetic either way: identical bits, and the oracle's event sequence.
    def relay():   # examples/ode-2-relay-msign.rs:15-26 without its Trace (a traced problem is not a LEAN/SPEC build)
        return (d.Solver.new().initial(np.tile([0.25, 0.0], (40, 1)) + 0.01 * np.arange(40).reshape(-1, 1))
                .equation(d.systems.RelayMsign).interval(0.5, 10.5).stepsize(0.09)
                .on(d.Locator.zero("x"), None).log_events(d.EventLog(64)))
    for mk in (lambda: cases.bouncing_ball(n=70, t1=6.0), lambda: cases.egg_billiard(n=70, reflections=12), relay):
        spec = mk().arith("fast").run()
        os.environ["DFB_EVENT_RUNTIME_SPEC"] = "1"
        try:
            rt = mk().arith("fast").run()
        finally:
            del os.environ["DFB_EVENT_RUNTIME_SPEC"]
        assert spec.totals["kernel_kind"] == KERNEL_ODE_EVENT and rt.totals["kernel_kind"] == KERNEL_ODE_EVENT
        _same_everything(spec, rt)
        if mk is relay:   # a relay chatters around x = 0: FAST rounding may legitimately re-order its crossings
            continue
        o = oracle.run_solver(mk().arith("strict"), n_threads=4)
        assert np.array_equal(spec.event_n, o.event_n) and np.array_equal(spec.event_index, o.event_index)
        assert np.array_equal(spec.steps, o.steps) and np.array_equal(spec.events, o.events)
        assert np.nanmax(np.abs(spec.event_t - o.event_t)) < 1e-9


def test_event_kernel_dynamic_assignment_ragged_costs(oracle):
    # members stop after very different numbers of reflections (1 .. 97), far more members than one
    # block: lanes hand in finished members and pull new ones from the work counter; results are
    # indexed by member and must not depend on which lane integrated what
    n = 3001
    refl = (np.arange(n) * 7 % 97 + 1).astype(np.float64)

    def mk(arith="strict"):
        return _egg_with_reflections(n, refl, arith)
    hot = mk().run()
    assert hot.totals["kernel_kind"] == KERNEL_ODE_EVENT
    assert np.array_equal(hot.events, refl.astype(np.uint64)) and np.all(hot.status == d.abi.TRAJ_STOPPED)
    o = oracle.run_solver(mk(), n_threads=8)
    assert_same_counters(hot, o)
    assert np.array_equal(hot.p, o.p) and np.array_equal(hot.aux, o.aux)
    # FAST: pre-combined interpolant -> event times within tolerance, same event counts
    f = mk("fast").run()
    assert f.totals["kernel_kind"] == KERNEL_ODE_EVENT
    assert np.array_equal(f.events, hot.events)
    few = refl <= 10                      # the billiard map is chaotic: compare short itineraries
    assert np.max(np.abs(f.p[few] - hot.p[few])) < 1e-9


def _egg_with_reflections(n, refl, arith):
    phi = 2.0 * np.pi * np.arange(n) / n
    y0 = np.stack([np.zeros(n), np.full(n, 0.1), 0.5 * np.cos(phi), 0.5 * np.sin(phi)], axis=1)
    return (d.Solver.new().initial(y0).equation(d.systems.EggBilliard, refl.reshape(-1, 1))
            .interval(0.0, math.inf).stepsize(0.5).on_mut(d.Locator.above_zero("boundary"), "reflect")
            .arith(arith).max_steps(100000))


def test_event_kernel_fast_event_log_vs_oracle(oracle):
    g = cases.bouncing_ball(n=64, t1=6.0, arith="fast").run()
    assert g.totals["kernel_kind"] == KERNEL_ODE_EVENT
    o = oracle.run_solver(cases.bouncing_ball(n=64, t1=6.0), n_threads=8)
    assert np.array_equal(g.event_n, o.event_n) and np.array_equal(g.event_index, o.event_index)
    assert np.max(np.abs(g.event_t - o.event_t)) < 1e-11     # "event times within the solver tolerance"
    assert np.max(np.abs(g.p - o.p)) < 1e-10


def test_event_kernel_step_limit_status(oracle):
    # max_steps guard on an unbounded interval (the egg billiard never stops by itself with a huge quota)
    mk = lambda: (_egg_with_reflections(40, np.full(40, 1e9), "strict").max_steps(500))
    hot = mk().run()
    assert hot.totals["kernel_kind"] == KERNEL_ODE_EVENT
    o = oracle.run_solver(mk(), n_threads=4)
    assert np.all(hot.status == 16) and np.array_equal(hot.status, o.status)
    assert np.array_equal(hot.steps, o.steps) and np.all(hot.steps == 500)

# ---------------------------------------------------------------- DDE / NDDE
def test_dde_constant_history_strict_bit_exact(oracle):
    # x' = a x + b x(t - 1) with constant history: no transcendental call anywhere
    def mk():
        prm = np.stack([np.linspace(-1.0, 0.5, 40), np.linspace(-2.0, -0.5, 40), np.ones(40), np.zeros(40)], 1)
        return (d.Solver.new().max_delay(1.0).initial(np.ones((40, 1))).interval(0.0, 6.0)
                .stepsize(0.02).equation(d.systems.DdeLinear, prm).on_step(d.Trace(1, 320)))
    g, o = run_both(oracle, mk)
    assert np.all(g.status == 0)
    assert_same_counters(g, o)
    assert_same_trace(g, o, exact=True)


def test_dde_sin_reference_tolerance_and_oracle(oracle):
    # tests/equation_types.rs:133: |x - sin t| < 1e-11; vs oracle: sin() differs by <= 2 ulp
    for arith in ("strict", "fast"):
        g = cases.eq_dde_sin(arith=arith).run()
        assert g.status[0] == 0
        n = g.trace_n[0]
        assert n == 502  # 500 steps of 0.02 land one ulp short of 10, then one closing step
        assert np.all(np.abs(g.trace_p[0, :n, 0] - np.sin(g.trace_t[0, :n])) < 1e-11)
        o = oracle.run_solver(cases.eq_dde_sin())
        assert np.array_equal(g.trace_t[0, :n], o.trace_t[0, :n])
        assert np.max(np.abs(g.trace_p[0, :n] - o.trace_p[0, :n])) < 1e-13


def test_dde_sin_ensemble_k_sweep(oracle):
    k = 0.5 + np.arange(200) / 199.0
    mk = lambda: cases.eq_dde_sin(k=k, t1=10.0, trace=False)
    g, o = run_both(oracle, mk)
    assert np.all(g.status == 0)
    assert np.array_equal(g.steps, o.steps)
    assert np.max(np.abs(g.p - o.p)) < 1e-13
    assert np.max(np.abs(g.p[:, 0] - np.sin(k * 10.0))) < 1e-10


def test_ndde_sin(oracle):
    g = cases.eq_ndde_sin().run()
    n = g.trace_n[0]
    assert g.status[0] == 0 and n == 502
    assert np.all(np.abs(g.trace_p[0, :n, 0] - np.sin(g.trace_t[0, :n])) < 1e-11)  # :155
    o = oracle.run_solver(cases.eq_ndde_sin())
    assert np.max(np.abs(g.trace_p[0, :n] - o.trace_p[0, :n])) < 1e-13


def test_dde_hot_path_fast_vs_oracle(oracle):
    # FAST + fixed step + no trace -> the streaming-window kernel with pre-combined
    # coefficient records (dev/dde_fixed.cuh).  Non-chaotic: within 1e-12 of the oracle.
    k = 0.5 + np.arange(333) / 332.0          # ragged: 333 = 2 blocks + 77
    for t1 in (0.5, 1.5, 10.0, 30.0):         # inside the history, across t = tau, long
        mk = lambda arith: cases.eq_dde_sin(k=k, t1=t1, trace=False, arith=arith)
        g = mk("fast").run()
        o = oracle.run_solver(mk("strict"), n_threads=8)
        assert np.all(g.status == 0)
        assert np.array_equal(g.steps, o.steps) and np.array_equal(g.rhs_evals, o.rhs_evals)
        assert np.array_equal(g.t, o.t)
        assert np.max(np.abs(g.p - o.p)) < 1e-12, (t1, np.max(np.abs(g.p - o.p)))
        assert np.max(np.abs(g.p[:, 0] - np.sin(k * t1))) < 1e-10
    # the general kernel in FAST arithmetic agrees too (same inputs, forced)
    os.environ["DFB_FORCE_GENERAL"] = "1"
    try:
        gen = cases.eq_dde_sin(k=k, t1=10.0, trace=False, arith="fast").run()
    finally:
        del os.environ["DFB_FORCE_GENERAL"]
    hot = cases.eq_dde_sin(k=k, t1=10.0, trace=False, arith="fast").run()
    assert np.max(np.abs(gen.p - hot.p)) < 1e-12


def test_ndde_hot_path_fast_vs_oracle(oracle):
    k = 0.5 + np.arange(130) / 129.0
    def mk(arith):
        return (d.Solver.new().initial(d.InitFn(d.SIN, d.SIN_DERIV))
                .equation(d.systems.NddeLinear, cases.ndde_params(k)).interval(0.0, 10.0)
                .max_delay(1.0).stepsize(0.02).arith(arith))
    g = mk("fast").run()
    o = oracle.run_solver(mk("strict"), n_threads=8)
    assert np.all(g.status == 0) and np.array_equal(g.steps, o.steps)
    assert np.max(np.abs(g.p - o.p)) < 1e-11
    assert np.max(np.abs(g.p[:, 0] - np.sin(k * 10.0))) < 1e-9


def test_dde_hot_path_other_tableaux_and_delays(oracle):
    # rk43 (S=5, I=4 cubic interpolant), rk4 (linear interpolant), delays from 1.6 h to 40 h,
    # constant history (no transcendental call)
    for rk, h in (("rk43", 0.02), ("rk4", 0.01), ("rktp64", 0.05)):
        for tau in (1.6 * h, 2.5 * h, 3.0 * h, 40.0 * h):
            n = 70
            prm = np.stack([np.linspace(-1.0, 0.5, n), np.linspace(-2.0, -0.5, n), np.full(n, tau), np.zeros(n)], 1)
            def mk(arith):
                return (d.Solver.new().max_delay(tau).initial(np.ones((n, 1))).interval(0.0, 3.0)
                        .rk(d.RK.by_name(rk)).stepsize(h).equation(d.systems.DdeLinear, prm).arith(arith))
            g = mk("fast").run()
            o = oracle.run_solver(mk("strict"), n_threads=8)
            assert np.all(o.status == 0) and np.all(g.status == 0), (rk, tau / h)
            assert np.array_equal(g.steps, o.steps)
            assert np.max(np.abs(g.p - o.p) / np.maximum(np.abs(o.p), 1.0)) < 1e-12, (rk, tau / h)



def test_dde_hot_path_delay_is_an_exact_multiple_of_the_step(oracle):
    # tau / h an integer puts every delayed stage time one ulp away from a record boundary.  The
    # reference then reads the record that STARTS at the boundary; for y'(t - tau) of a linear
    # interpolant (rk4) the neighbouring record's value differs by O(h).  Found by the fuzz campaign
    # (tests/test_fuzz_parity.py, seed 54 at DFB_FUZZ_SCALE=6): the window used to slide on
    # t - tau instead of the earliest lookup of the step.
    k = 0.4 + np.arange(64) / 63.0
    for tau, h, rk in ((0.5, 0.05, d.RK.rk4()), (1.0, 0.125, d.RK.rk4()), (0.5, 0.025, d.RK.rk43()),
                       (1.0, 0.02, d.RK.rktp64())):
        for system, prm, init in ((d.systems.NddeLinear, cases.ndde_params(k, tau), d.InitFn(d.SIN, d.SIN_DERIV)),
                                  (d.systems.DdeLinear, cases.dde_params(k, tau), d.InitFn(d.SIN))):
            mk = lambda arith: (d.Solver.new().max_delay(tau).initial(init).interval(0.0, 4.0).stepsize(h).rk(rk)
                                .equation(system, prm).arith(arith))
            f = mk("fast").run()
            assert f.totals["kernel_kind"] == d.abi.KERNEL_DDE_FIXED and np.all(f.status == 0)
            o = oracle.run_solver(mk("strict"), n_threads=4)
            assert np.array_equal(f.steps, o.steps) and np.array_equal(f.t, o.t)
            assert np.max(np.abs(f.p - o.p)) < 1e-11, (system.name, tau, h)


def test_dde_hot_path_builds_agree_lean_loop(oracle):
    # The three builds of the constant-delay hot path — one member per thread without the register cap (PRE:
    # pre-evaluated delayed values, lean steady loop, record w+3 prefetched into 8 staging slots), one member
    # per thread capped at 72 registers, two members per thread — run the same arithmetic per member: identical
    # bits among themselves, the oracle's step grid, the FAST tolerance against the oracle's state.  Delays from
    # 6 to 200 steps put the window anywhere from right behind the write head to the far end of the ring.
    k = 0.5 + np.arange(133) / 132.0                       # ragged: 4 warps + 5 members
    for tau, h, t1 in ((1.0, 0.02, 25.0), (0.3, 0.05, 30.0), (2.0, 0.01, 12.0), (1.0, 0.037, 20.0),
                       (1.0, 0.0123456789, 120.0), (0.77, 0.1, 400.0)):   # ~10^4 and 4000 steps: many window slides
        mk = lambda arith: (d.Solver.new().max_delay(tau).initial(d.InitFn(d.SIN)).interval(0.0, t1).stepsize(h)
                            .equation(d.systems.DdeLinear, cases.dde_params(k, tau)).arith(arith))
        o = oracle.run_solver(mk("strict"), n_threads=8)
        runs = {}
        for name, env in (("pre", {"DFB_DDE_TPT": "1"}), ("capped", {"DFB_DDE_TPT": "1", "DFB_DDE_RELAXED": "0"}),
                          ("two", {"DFB_DDE_TPT": "2"})):
            os.environ.update(env)
            try:
                g = mk("fast").run()
            finally:
                for e in env:
                    del os.environ[e]
            assert g.totals["kernel_kind"] == d.abi.KERNEL_DDE_FIXED and np.all(g.status == 0), name
            assert np.array_equal(g.steps, o.steps) and np.array_equal(g.t, o.t), name
            assert np.max(np.abs(g.p - o.p)) < 1e-9, (name, tau, h)
            runs[name] = g
        assert np.array_equal(runs["pre"].p, runs["capped"].p), (tau, h)
        assert np.array_equal(runs["pre"].p, runs["two"].p), (tau, h)


def test_dde_hot_path_trace_policy(oracle):
    # Trace (every k-th on_step call, capacity-bounded) inside both DDE hot kernels: the time grid and
    # the sample counts are the oracle's bit for bit, the samples within the FAST tolerance, the last
    # sample (stride 1) is the final state of the untraced run.
    k = 0.5 + np.arange(165) / 164.0              # ragged: 5 warps + 5 members
    for tpt in ("1", "2"):
        os.environ["DFB_DDE_TPT"] = tpt
        try:
            for every, cap, t1 in ((1, 600, 10.0), (7, 600, 10.0), (3, 50, 10.0), (1, 600, 0.03)):
                mk = lambda arith: cases.eq_dde_sin(k=k, t1=t1, trace=False, arith=arith).on_step(d.Trace(every, cap))
                g = mk("fast").run()
                assert g.totals["kernel_kind"] == d.abi.KERNEL_DDE_FIXED and np.all(g.status == 0)
                o = oracle.run_solver(mk("strict"), n_threads=8)
                assert np.array_equal(g.steps, o.steps)
                assert np.array_equal(g.trace_n, o.trace_n)
                n = int(g.trace_n[0])
                assert np.all(g.trace_n == n)
                assert np.array_equal(g.trace_t[:, :n], o.trace_t[:, :n])
                assert np.max(np.abs(g.trace_p[:, :n] - o.trace_p[:, :n])) < 1e-12, (tpt, every, cap)
                if every == 1 and cap > n:
                    plain = cases.eq_dde_sin(k=k, t1=t1, trace=False, arith="fast").run()
                    assert np.array_equal(g.trace_p[:, n - 1, :], plain.p) and np.array_equal(g.p, plain.p)
        finally:
            del os.environ["DFB_DDE_TPT"]


def test_dde_hot_path_trace_stops_counting_at_a_panic(oracle):
    # the reference stops at a panic; a panicked member of the hot kernel keeps stepping (its warp stays
    # converged) but its sample count is frozen where the reference's trace ends
    k = 0.5 + np.arange(70) / 69.0
    for tpt in ("1", "2"):
        os.environ["DFB_DDE_TPT"] = tpt
        try:
            for must_panic, tweak in ((True, lambda s: s.stepsize(1.5)), (True, lambda s: s.max_delay(0.5)),
                                      (False, lambda s: s.max_delay(0.97)), (False, lambda s: s.max_delay(0.9))):
                mk = lambda arith: tweak(cases.eq_dde_sin(k=k, t1=4.0, trace=False, arith=arith)).on_step(d.Trace(1, 400))
                g, o = mk("fast").run(), oracle.run_solver(mk("strict"), n_threads=4)
                assert g.totals["kernel_kind"] == d.abi.KERNEL_DDE_FIXED
                assert np.array_equal(g.status, o.status) and (np.any(o.status != 0) or not must_panic)
                assert np.array_equal(g.trace_n, o.trace_n)
                for i in range(0, 70, 9):
                    n = int(g.trace_n[i])
                    assert np.array_equal(g.trace_t[i, :n], o.trace_t[i, :n])
                    assert np.max(np.abs(g.trace_p[i, :n] - o.trace_p[i, :n])) < 1e-12
        finally:
            del os.environ["DFB_DDE_TPT"]


def test_dde_hot_path_panics_become_status_bits(oracle):
    mk = lambda: cases.eq_dde_sin(trace=False, arith="fast").stepsize(1.5)   # delay <= h
    g, o = mk().run(), oracle.run_solver(mk())
    assert g.status[0] & d.abi.TRAJ_HISTORY_NOT_READY and o.status[0] & d.abi.TRAJ_HISTORY_NOT_READY
    mk = lambda: cases.eq_dde_sin(trace=False, arith="fast").max_delay(0.5)  # window < delay
    g, o = mk().run(), oracle.run_solver(mk())
    assert g.status[0] & d.abi.TRAJ_HISTORY_DELETED and o.status[0] & d.abi.TRAJ_HISTORY_DELETED
    # a window only slightly shorter than the delay: the reference's prune rule
    # (t_deque[1] < t_prev - max_delay) decides; the kernel replays it exactly
    for md in (0.97, 0.985, 0.995, 1.0):
        mk = lambda: cases.eq_dde_sin(trace=False, arith="fast", t1=4.0).max_delay(md)
        g, o = mk().run(), oracle.run_solver(mk())
        assert (g.status[0] & d.abi.TRAJ_HISTORY_DELETED) == (o.status[0] & d.abi.TRAJ_HISTORY_DELETED), md


def test_full_size_dde_properties(oracle):
    # BASELINE config 3 at full size: 2^18 members, k sweep, t in [0, 100]; the exact solution
    # sin(k t) is the size-independent check, plus duplicated inputs -> identical bits
    n = 1 << 18
    k = 0.5 + np.arange(n) / (n - 1.0)
    k[n // 2:] = k[: n // 2]
    g = cases.eq_dde_sin(k=k, t1=100.0, trace=False, arith="fast").run()
    assert np.all(g.status == 0)
    assert np.all(g.steps == g.steps[0]) and 5000 <= g.steps[0] <= 5001
    assert np.array_equal(g.p[: n // 2], g.p[n // 2:])
    assert np.max(np.abs(g.p[:, 0] - np.sin(k * g.t))) < 1e-9
    sub = slice(0, n, n // 64)
    o = oracle.run_solver(cases.eq_dde_sin(k=k[sub], t1=100.0, trace=False), n_threads=8)
    assert np.max(np.abs(g.p[sub] - o.p)) < 1e-11


def test_dde_panics_become_status_bits(oracle):
    s = lambda: cases.eq_dde_sin().stepsize(1.5)           # delay <= h  (state.rs:172-177)
    g, o = run_both(oracle, s)
    assert g.status[0] & d.abi.TRAJ_HISTORY_NOT_READY and o.status[0] & d.abi.TRAJ_HISTORY_NOT_READY
    s = lambda: cases.eq_ndde_sin().initial(d.InitFn(d.SIN))  # no derivative (initial_condition.rs:33)
    g, o = run_both(oracle, s)
    assert g.status[0] & d.abi.TRAJ_NO_DERIVATIVE and o.status[0] & d.abi.TRAJ_NO_DERIVATIVE
    s = lambda: cases.eq_dde_sin().max_delay(0.5)           # window shorter than the delay (state.rs:166-171)
    g, o = run_both(oracle, s)
    assert g.status[0] & d.abi.TRAJ_HISTORY_DELETED and o.status[0] & d.abi.TRAJ_HISTORY_DELETED


def test_dde_relay_clamp_strict_bit_exact(oracle):
    mk = lambda: cases.dde_relay_clamp(t1=20.0)
    g, o = run_both(oracle, mk)
    assert g.status[0] == 0
    assert_same_counters(g, o)
    assert np.array_equal(g.event_t, o.event_t) and np.array_equal(g.event_index, o.event_index)
    assert_same_trace(g, o, exact=True)
    assert g.events[0] > 10


def test_adaptive_strict_is_step_for_step_identical_to_the_golden_and_the_oracle(oracle):
    # STRICT adaptive (general kernel): error norm, clamps, reject loop, k[0] re-evaluation and the portable
    # n-th root are the oracle's operations in the oracle's order -> every accepted step, every rejection and
    # the final state are bit-identical to tests/golden/adaptive_golden.json ("portable_root" variant, frozen
    # from an independent numpy restatement) and to the oracle run with the same root
    import json
    import test_oracle_independent as ti
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "adaptive_golden.json")))["cases"]
    for name, sys_name, rk_name, y0, prm, t0, t1, ctl in ti.adaptive_cases():
        mk = lambda: ti.adaptive_solver(sys_name, rk_name, y0, prm, t0, t1, ctl, arith="strict").on_step(d.Trace(1, 12000))
        g = mk().run()
        e = gold[name]["portable_root"]
        assert g.totals["kernel_kind"] == KERNEL_GENERAL and g.status[0] == 0, name
        assert [int(g.steps[0]), int(g.rejected[0]), int(g.rhs_evals[0])] == e["steps_rejected_evals"], name
        assert float(g.t[0]).hex() == e["t"] and [float(v).hex() for v in g.p[0]] == e["p"], name
        o = oracle.run_solver(mk(), n_threads=1, portable_root=True)
        assert_same_counters(g, o)
        assert_same_trace(g, o, exact=True)   # every accepted step: same time, same state, bit for bit
        # FAST hot kernel (MUFU-seeded root, FMA): same problem within the solver tolerance
        f = ti.adaptive_solver(sys_name, rk_name, y0, prm, t0, t1, ctl, arith="fast").run()
        assert f.status[0] == 0 and abs(int(f.steps[0]) - int(g.steps[0])) <= max(2, int(g.steps[0]) // 50), name
        scale = np.maximum(np.abs(g.p[0]), 1.0)
        assert np.all(np.abs(f.p[0] - g.p[0]) / scale < 2e3 * max(ctl["atol"][0], 1e-9)), (name, f.p[0], g.p[0])


# ---------------------------------------------------------------- dde_event_kernel (DDE + events + propagation)
KERNEL_DDE_EVENT = 6


def test_dde_event_kernel_relay_clamp_fast_vs_oracle(oracle):
    # examples/dde-relay-clamp.rs:19-36 through the hot kernel: same event sequence, same step counts,
    # event times within 1e-11, trajectory within 1e-10 of the (strict) oracle over t in [0, 20]
    mk = lambda: cases.dde_relay_clamp(arith="fast", t1=20.0)
    g = mk().run()
    o = oracle.run_solver(cases.dde_relay_clamp(arith="strict", t1=20.0), n_threads=1)
    assert g.totals["kernel_kind"] == KERNEL_DDE_EVENT and g.status[0] == 0
    assert_same_counters(g, o)
    n = int(g.event_n[0])
    assert n == int(o.event_n[0]) and n > 10
    assert np.array_equal(g.event_index[0, :n], o.event_index[0, :n])
    assert np.allclose(g.event_t[0, :n], o.event_t[0, :n], rtol=0.0, atol=1e-11)
    assert_same_trace(g, o, exact=False, rtol=1e-10, atol=1e-10)
    assert g.totals["locate_iters"] > 40 * g.events[0]  # ~45 bisection iterations per located event


def test_dde_event_kernel_equals_general_fast_kernel():
    # the two FAST kernels implement the same arithmetic contract: same events, states within 1e-11
    mk = lambda: cases.dde_relay_clamp_ensemble(96, arith="fast", t1=12.0, events=256, trace=d.Trace(7, 200))
    hot = mk().run()
    gen = run_general(mk)
    assert hot.totals["kernel_kind"] == KERNEL_DDE_EVENT
    assert np.array_equal(hot.steps, gen.steps) and np.array_equal(hot.events, gen.events)
    assert np.array_equal(hot.status, gen.status) and np.all(hot.status == 0)
    assert np.array_equal(hot.event_n, gen.event_n) and np.array_equal(hot.event_index, gen.event_index)
    assert np.allclose(hot.event_t, gen.event_t, rtol=0.0, atol=1e-11)
    assert np.allclose(hot.p, gen.p, rtol=1e-11, atol=1e-11)
    assert_same_trace(hot, gen, exact=False, rtol=1e-11, atol=1e-11)


def test_dde_event_kernel_ensemble_vs_oracle_ragged(oracle):
    # 1000 members (a ragged last warp, members pulled from the work queue in any order): per-member
    # results do not depend on the assignment; checked member by member against the oracle
    mk = lambda a: cases.dde_relay_clamp_ensemble(1000, arith=a, t1=8.0, events=128)
    g = mk("fast").run()
    o = oracle.run_solver(mk("strict"), n_threads=8)
    assert g.totals["kernel_kind"] == KERNEL_DDE_EVENT
    assert_same_counters(g, o)
    assert np.array_equal(g.event_n, o.event_n) and np.array_equal(g.event_index, o.event_index)
    assert np.allclose(g.event_t, o.event_t, rtol=0.0, atol=1e-11)
    assert np.allclose(g.p, o.p, rtol=1e-10, atol=1e-10)
    g2 = mk("fast").run()
    assert np.array_equal(g.p, g2.p) and np.array_equal(g.event_t, g2.event_t)  # run-to-run identical


def test_dde_event_kernel_propagation_golden_vectors():
    # tests/delay_propagation.rs:21-24, 45-50, 68-71, 91-93 in FAST arithmetic: the time grid of these
    # problems is pure time arithmetic (rhs = 0), so the hot kernel must hit the exact f64 vectors too
    g = cases.delay_constant_1().arith("fast").run()
    assert g.totals["kernel_kind"] == KERNEL_DDE_EVENT
    assert list(g.trace_t[0, :g.trace_n[0]]) == [0.0, 0.75, 1.0, 1.75, 2.0, 2.75, 3.0, 3.75, 4.0, 4.75, 5.0]
    g = cases.delay_constant_2().arith("fast").run()
    s = 7.0 / 8.0
    assert g.totals["kernel_kind"] == KERNEL_DDE_EVENT
    assert list(g.trace_t[0, :g.trace_n[0]]) == [0.0, s, 1.0, 1.25, 2., 2.25, 2.5, 3.0, 3.25, 3.5,
                                                  3.75, 4., 4.25, 4.5, 4.75, 5.0]
    g = cases.delay_constant_1_smoothing().arith("fast").run()
    assert g.totals["kernel_kind"] == KERNEL_DDE_EVENT
    assert list(g.trace_t[0, :g.trace_n[0]]) == [0., 0.75, 1., 1.75, 2., 2.75, 3., 3.75, 4.5, 5.25, 6.]
    g = cases.delay_pantograph().arith("fast").run()
    assert g.totals["kernel_kind"] == KERNEL_DDE_EVENT
    ts = set(g.trace_t[0, :g.trace_n[0]])
    for i in range(1, 11):
        assert 2.0 ** i in ts


def test_dde_event_kernel_panics_and_plain_dde(oracle):
    # (a) a DDE without locators but with a trace + an on-step policy the streaming kernel does not take
    #     (max_steps) lands here; (b) the reference's panics become the same status bits
    mk = lambda a: cases.eq_dde_sin(k=np.linspace(0.6, 1.4, 40), t1=6.0, trace=True, arith=a).max_steps(100000)
    g = mk("fast").run()
    o = oracle.run_solver(mk("strict"), n_threads=4)
    assert g.totals["kernel_kind"] == KERNEL_DDE_EVENT
    assert_same_counters(g, o)
    assert_same_trace(g, o, exact=False, rtol=1e-11, atol=1e-12)
    s = lambda a: cases.eq_dde_sin(arith=a).max_delay(0.5).max_steps(100000)  # window shorter than the delay
    g, o = s("fast").run(), oracle.run_solver(s("strict"))
    assert g.totals["kernel_kind"] == KERNEL_DDE_EVENT
    assert g.status[0] & d.abi.TRAJ_HISTORY_DELETED and o.status[0] & d.abi.TRAJ_HISTORY_DELETED
    assert g.steps[0] == o.steps[0]


# ---------------------------------------------------------------- adaptive
def auto_stepsize(n_state, tol=1e-9):
    return d.AutomaticStepsize(stepsize=1e-2, stepsize_range=(1e-6, 1e-1), atol=[tol] * n_state,
                               rtol=[tol] * n_state, order=4, fac=0.9, fac_range=(0.2, 5.0))


def test_adaptive_harmonic_vs_oracle(oracle):
    # AutomaticStepsize is exercised by no reference test ("parity unpinned" upstream);
    # pow() of CUDA and glibc agree to <= 2 ulp, so step sequences agree to tolerance.
    def mk(arith="strict"):
        return (d.Solver.new().initial(np.tile([0.0, 1.0], (33, 1))).equation(d.systems.Harmonic)
                .interval(0.0, 10.0).stepsize(auto_stepsize(2)).on_step(d.Trace(1, 4096)).arith(arith))
    g, o = run_both(oracle, mk)
    assert np.all(g.status == 0)
    assert np.array_equal(g.steps, o.steps) and np.array_equal(g.rejected, o.rejected)
    n = g.trace_n[0]
    # "adaptive and event times within the solver tolerance" (atol = rtol = 1e-9)
    assert np.max(np.abs(g.trace_t[0, :n] - o.trace_t[0, :n])) < 1e-8
    assert np.max(np.abs(g.p - o.p)) < 1e-9
    assert np.max(np.abs(g.p[:, 0] - math.sin(10.0))) < 1e-6
    assert g.rejected[0] > 0 or g.steps[0] < 10.0 / 1e-2  # the controller actually adapts
    f = mk("fast").run()
    assert np.max(np.abs(f.p - o.p)) < 1e-9


def test_adaptive_lorenz_short(oracle):
    def mk():
        return cases.lorenz(n=65, t1=2.0).stepsize(auto_stepsize(3))
    g, o = run_both(oracle, mk)
    assert np.all(g.status == 0)
    assert np.max(np.abs(g.steps.astype(np.int64) - o.steps.astype(np.int64))) <= 1
    assert np.max(np.abs(g.p - o.p) / np.maximum(np.abs(o.p), 1.0)) < 1e-7



# ---------------------------------------------------------------- adaptive hot path (ode_adaptive_kernel)


def check_adaptive_counters(g, S):
    # rhs evaluations as the reference counts them: S per attempt, k[0] at the start and after every undo
    att = g.steps + g.rejected
    assert np.array_equal(g.rhs_evals, att * S + 1 + g.rejected)


def test_adaptive_hot_path_harmonic_vs_oracle(oracle):
    # 257 members (ragged block) with different initial amplitudes -> different step sequences
    amp = 1.0 + np.arange(257) / 64.0
    y0 = np.stack([np.zeros(257), amp], axis=1)

    def mk(arith="fast"):
        return (d.Solver.new().initial(y0).equation(d.systems.Harmonic).interval(0.0, 10.0)
                .stepsize(auto_stepsize(2)).arith(arith))
    f = mk().run()
    assert f.totals["kernel_kind"] == KERNEL_ODE_ADAPTIVE
    o = oracle.run_solver(mk("strict"), n_threads=8)
    assert np.all(f.status == 0)
    check_adaptive_counters(f, 7)
    # "adaptive ... within the solver tolerance" (atol = rtol = 1e-9): the controller's root is a
    # Newton iteration instead of libm pow, so a step count may differ by one near err = 1
    assert np.max(np.abs(f.steps.astype(np.int64) - o.steps.astype(np.int64))) <= 1
    assert np.max(np.abs(f.rejected.astype(np.int64) - o.rejected.astype(np.int64))) <= 1
    assert np.array_equal(f.t, o.t)
    assert np.max(np.abs(f.p - o.p)) < 1e-9
    assert np.max(np.abs(f.p[:, 0] - amp * math.sin(10.0))) < 1e-6
    # and against the general kernel in the same arithmetic
    os.environ["DFB_FORCE_GENERAL"] = "1"
    try:
        gen = mk().run()
    finally:
        del os.environ["DFB_FORCE_GENERAL"]
    assert gen.totals["kernel_kind"] == KERNEL_GENERAL
    assert np.max(np.abs(f.p - gen.p)) < 1e-11
    assert np.max(np.abs(f.steps.astype(np.int64) - gen.steps.astype(np.int64))) <= 1


def test_adaptive_hot_path_trace_policy(oracle):
    # Trace inside the adaptive hot kernel: every member has its own time grid.  Sample counts follow
    # from the step counts, the last sample (stride 1) is the final state, every sample lies within
    # the solver tolerance of the exact solution, and members whose step counts equal the oracle's
    # have the oracle's sample times within the controller's sensitivity.
    amp = 1.0 + np.arange(257) / 64.0
    y0 = np.stack([np.zeros(257), amp], axis=1)
    for every, cap in ((1, 4096), (5, 4096), (2, 20)):
        def mk(arith="fast"):
            return (d.Solver.new().initial(y0).equation(d.systems.Harmonic).interval(0.0, 10.0)
                    .stepsize(auto_stepsize(2)).arith(arith).on_step(d.Trace(every, cap)))
        f = mk().run()
        assert f.totals["kernel_kind"] == KERNEL_ODE_ADAPTIVE and np.all(f.status == 0)
        assert np.array_equal(f.trace_n, np.minimum(f.steps // every + 1, cap))
        o = oracle.run_solver(mk("strict"), n_threads=8)
        same = np.flatnonzero((f.steps == o.steps) & (f.rejected == o.rejected))
        assert same.size > 200
        for i in same[::16]:
            n = int(f.trace_n[i])
            assert n == int(o.trace_n[i])
            # the error estimate is a difference of O(1e-9) quantities: FMA contraction moves it by
            # ~1e-7 relative, hence the step sizes by ~1e-10 each, accumulating along the grid
            assert np.max(np.abs(f.trace_t[i, :n] - o.trace_t[i, :n])) < 1e-5
            assert np.max(np.abs(f.trace_p[i, :n] - o.trace_p[i, :n])) < 1e-4
        for i in range(0, 257, 8):
            n = int(f.trace_n[i])
            t = f.trace_t[i, :n]
            assert t[0] == 0.0 and np.all(np.diff(t) > 0)
            assert np.max(np.abs(f.trace_p[i, :n, 0] - amp[i] * np.sin(t))) < 1e-6
        if every == 1:
            n = f.trace_n
            assert np.array_equal(f.trace_t[np.arange(257), n - 1], f.t)
            assert np.array_equal(f.trace_p[np.arange(257), n - 1, :], f.p)
            plain = (d.Solver.new().initial(y0).equation(d.systems.Harmonic).interval(0.0, 10.0)
                     .stepsize(auto_stepsize(2)).arith("fast").run())
            assert np.array_equal(plain.p, f.p) and np.array_equal(plain.steps, f.steps)


@pytest.mark.parametrize("rk,order", [("rktp64", 4), ("rk43", 3), ("rk4", 2), ("heun2", 1), ("three_eights", 2)])
def test_adaptive_hot_path_tableaux(oracle, rk, order):
    lam = -0.5 - np.arange(100) / 50.0
    ctl = d.AutomaticStepsize(stepsize=1e-2, stepsize_range=(1e-7, 1e-1), atol=[1e-8], rtol=[1e-8],
                              order=order, fac=0.9, fac_range=(0.2, 5.0))

    def mk(arith="fast"):
        return (d.Solver.new().initial(np.ones((100, 1))).equation(d.systems.Linear1, lam.reshape(-1, 1))
                .interval(0.0, 3.0).stepsize(ctl).rk(d.RK.by_name(rk)).arith(arith))
    f = mk().run()
    assert f.totals["kernel_kind"] == KERNEL_ODE_ADAPTIVE
    o = oracle.run_solver(mk("strict"), n_threads=8)
    assert np.all(f.status == 0) and np.all(o.status == 0)
    check_adaptive_counters(f, d.RK.by_name(rk).S)
    assert np.max(np.abs(f.steps.astype(np.int64) - o.steps.astype(np.int64))) <= 1
    assert np.array_equal(f.t, o.t)
    assert np.max(np.abs(f.p - o.p) / np.abs(o.p)) < 1e-8


def test_adaptive_hot_path_lorenz_short(oracle):
    def mk(arith="fast"):
        return cases.lorenz(n=1000, t1=2.0, arith=arith).stepsize(auto_stepsize(3))
    f = mk().run()
    assert f.totals["kernel_kind"] == KERNEL_ODE_ADAPTIVE
    o = oracle.run_solver(mk("strict"), n_threads=8)
    assert np.all(f.status == 0)
    check_adaptive_counters(f, 7)
    assert np.max(np.abs(f.steps.astype(np.int64) - o.steps.astype(np.int64))) <= 1
    assert np.max(np.abs(f.p - o.p) / np.maximum(np.abs(o.p), 1.0)) < 1e-7


def test_adaptive_stepsize_floor_becomes_status_bit(oracle):
    # a tolerance the smallest allowed step cannot meet: the reference retries forever
    # (solver.rs:294-297); the library stops the trajectory with DFB_TRAJ_STEPSIZE_FLOOR
    ctl = d.AutomaticStepsize(stepsize=1e-1, stepsize_range=(5e-2, 1e-1), atol=[1e-14] * 3, rtol=[1e-14] * 3,
                              order=4, fac=0.9, fac_range=(0.2, 5.0))

    def mk(arith):
        return cases.lorenz(n=40, t1=2.0, arith=arith).stepsize(ctl)
    o = oracle.run_solver(mk("strict"), n_threads=4)
    g = mk("strict").run()   # general kernel
    f = mk("fast").run()     # adaptive hot path
    assert f.totals["kernel_kind"] == KERNEL_ODE_ADAPTIVE and g.totals["kernel_kind"] == KERNEL_GENERAL
    for r in (o, g, f):
        assert np.all(r.status & 256)
    assert np.array_equal(g.steps, o.steps) and np.array_equal(g.rejected, o.rejected)
    assert np.array_equal(f.steps, o.steps) and np.array_equal(f.rejected, o.rejected)
    assert np.all(np.isnan(f.t)) and np.all(np.isnan(g.t))
    # kutta3's "embedded" weights b2 = [0, 1, 0] (rk.rs) give an O(h) error estimate: with tol = 1e-8
    # the controller hits stepsize_range.start = 1e-7 at once (9 rejections, no accepted step)
    lam = -0.5 - np.arange(50) / 50.0
    ctl3 = d.AutomaticStepsize(stepsize=1e-2, stepsize_range=(1e-7, 1e-1), atol=[1e-8], rtol=[1e-8],
                               order=1, fac=0.9, fac_range=(0.2, 5.0))

    def mk3(arith):
        return (d.Solver.new().initial(np.ones((50, 1))).equation(d.systems.Linear1, lam.reshape(-1, 1))
                .interval(0.0, 3.0).stepsize(ctl3).rk(d.RK.kutta3()).arith(arith))
    o3 = oracle.run_solver(mk3("strict"), n_threads=4)
    f3 = mk3("fast").run()
    assert f3.totals["kernel_kind"] == KERNEL_ODE_ADAPTIVE
    assert np.all(o3.status == 256) and np.all(f3.status == 256)
    assert np.array_equal(f3.steps, o3.steps) and np.array_equal(f3.rejected, o3.rejected)
    assert np.all(f3.rejected == 9) and np.all(f3.steps == 0)


def test_adaptive_hot_path_empty_interval_and_degenerate_errors():
    # t1 == t0: no step, no rhs evaluation (solver.rs:292)
    s = cases.lorenz(n=33, t1=0.0, arith="fast").stepsize(auto_stepsize(3)).run()
    assert s.totals["kernel_kind"] == KERNEL_ODE_ADAPTIVE
    assert np.all(s.steps == 0) and np.all(s.rhs_evals == 0) and np.all(s.t == 0.0)
    assert np.array_equal(s.p, np.tile([1.0, 2.0, 20.0], (33, 1)))
    # y' = 0 * y: the error estimate is exactly 0 -> err = 0/atol = 0 -> factor clamps to fac_max
    z = (d.Solver.new().initial(np.ones((5, 1))).equation(d.systems.Linear1, np.zeros((5, 1)))
         .interval(0.0, 1.0).stepsize(auto_stepsize(1)).arith("fast").run())
    assert z.totals["kernel_kind"] == KERNEL_ODE_ADAPTIVE
    assert np.all(z.status == 0) and np.all(z.p == 1.0) and np.all(z.t == 1.0)
    # h: 1e-2 -> 5e-2 -> 1e-1 (stepsize_range end) ... : 1e-2, 5e-2, then 0.1 steps to t = 1
    assert np.all(z.rejected == 0) and np.all(z.steps == z.steps[0]) and 10 <= z.steps[0] <= 13


def test_full_size_adaptive_lorenz_properties():
    # BASELINE configs[1] "fixed and adaptive step": 2^20 members; size-independent properties
    n = 1 << 20
    s = cases.lorenz(n=n, t1=50.0, arith="fast").stepsize(auto_stepsize(3)).run()
    assert s.totals["kernel_kind"] == KERNEL_ODE_ADAPTIVE
    assert s.totals["n_failed"] == 0
    assert np.all(s.t == 50.0)
    check_adaptive_counters(s, 7)
    assert np.all(s.steps >= 500)  # h <= stepsize_range.end = 0.1
    # on the attractor (rho in [20, 40]): bounded, z > 0
    assert np.all(np.abs(s.p) < 100.0) and np.all(s.p[:, 2] > 0.0)
    # same member integrated alone = same member integrated inside the ensemble (dynamic assignment
    # of trajectories to lanes must not leak into results)
    pick = np.array([0, 1, 31, 32, 4097, n // 2, n - 1])
    rho = 20.0 + 20.0 * pick / (n - 1)
    alone = cases.lorenz(n=len(pick), t1=50.0, arith="fast", rho=rho).stepsize(auto_stepsize(3)).run()
    assert np.array_equal(alone.p, s.p[pick]) and np.array_equal(alone.steps, s.steps[pick])

# ---------------------------------------------------------------- Lyapunov
def test_lyapunov_scalar(oracle):
    lam = [-5., -4., -3., -2., -1., 1., 2., 3., 4., 5.]
    for tmax in (10., 100., 1000.):
        g = cases.lyap_linear_1_const(lam, tmax).run()
        assert np.all(np.abs(np.array(lam) - g.aux[:, 0] / tmax) < 2.0 / tmax)   # lyapunov_exponents.rs:42
        o = oracle.run_solver(cases.lyap_linear_1_const(lam, tmax))
        assert np.array_equal(g.steps, o.steps) and np.array_equal(g.p, o.p)
        assert np.max(np.abs(g.aux - o.aux)) < 1e-9 * tmax
        g = cases.lyap_linear_1_sin(lam, tmax).run()
        assert np.all(np.abs(np.array(lam) - g.aux[:, 0] / tmax) < 10.0 / tmax)  # :80


def test_lyapunov_matrix(oracle):
    real = [[2., 0., -2.], [3., 1., -2.], [5., 3., -2.], [3., 2., 1.], [4., -1., -1.]]
    mats = [cases.real_jordan_matrix(e) for e in real]
    for tmax in (10., 100.):
        g = cases.lyap_linear_3(mats, tmax).run()
        o = oracle.run_solver(cases.lyap_linear_3(mats, tmax))
        assert np.array_equal(g.steps, o.steps)
        assert np.array_equal(g.p, o.p)  # Householder QR: sqrt and division are IEEE on both sides
        for e, l in zip(real, g.aux / tmax):
            tol = 40.0 if (e[0] == e[1] or e[1] == e[2]) else 2.0
            assert np.max(np.abs(np.array(e) - l)) < tol / tmax                  # :161


def test_lyapunov_lorenz(oracle):
    # tests/lyapunov_exponents.rs:243-311 at tmax = 100 and 1000; oracle within 1e-3 (north_star)
    ref = np.array([0.90566, 0.0, -14.57233])
    for tmax in (100., 1000.):
        g = cases.lyap_lorenz(tmax).run()
        assert g.status[0] == 0
        lam = g.aux[0] / tmax
        assert np.max(np.abs(lam - ref)) < 0.00007 + 50.0 / tmax
    o = oracle.run_solver(cases.lyap_lorenz(100.))
    g = cases.lyap_lorenz(100.).run()
    assert np.array_equal(g.steps, o.steps) and g.events[0] == 200 and g.steps[0] + 1 == 20119
    assert np.max(np.abs(g.aux[0] / 100. - o.aux[0] / 100.)) < 1e-3
    assert np.max(np.abs(g.aux[0] - o.aux[0])) < 1e-8      # same trajectory; only log() differs
    # a rho grid, one trajectory per rho
    rho = np.linspace(25.0, 35.0, 40)
    gg = cases.lyap_lorenz(50., rho=rho).run()
    assert np.all(gg.status == 0) and np.all(gg.aux[:, 0] / 50. > 0.3)  # chaotic band


def test_lyapunov_lorenz_fast_vs_strict_ensemble_mean():
    # north_star: "chaotic systems compared on short horizons plus Lyapunov exponents within
    # 1e-3".  A FAST and a STRICT trajectory decorrelate after t ~ 40, so single finite-time
    # exponents differ by O(1/sqrt(T)); the comparison that is meaningful is between ensemble
    # means over the same perturbed initial conditions.
    n, tmax = 4096, 300.0
    rng = np.random.default_rng(11)
    def mk(arith):
        s = cases.lyap_lorenz(tmax, rho=np.full(n, 28.0), arith=arith)
        y0 = np.tile(np.array([10., 15., 20., 1., 0., 0., 0., 1., 0., 0., 0., 1.]), (n, 1))
        y0[:, :3] += 1e-3 * np.random.default_rng(11).uniform(-1, 1, size=(n, 3))
        return s.initial(y0)
    f, g = mk("fast").run(), mk("strict").run()
    assert np.all(f.status == 0) and np.all(g.status == 0)
    lf, lg = f.aux.mean(axis=0) / tmax, g.aux.mean(axis=0) / tmax
    assert np.max(np.abs(lf - lg)) < 1e-3, (lf, lg)
    # the three exponents sum to the trace of the Jacobian, -(sigma + 1 + beta), on both
    assert abs(lf.sum() + (10.0 + 1.0 + 8.0 / 3.0)) < 1e-6
    assert abs(lg.sum() + (10.0 + 1.0 + 8.0 / 3.0)) < 1e-6



# ---------------------------------------------------------------- fixed step + Periodic callbacks (ode_fixed_periodic_kernel)
KERNEL_ODE_FIXED_PERIODIC = 4


@pytest.mark.parametrize("arith", ["strict", "fast"])
def test_periodic_kernel_lyapunov_lorenz_equals_general_and_oracle(oracle, arith):
    rho = np.linspace(20.0, 40.0, 77)   # 77: ragged warp
    tmax = 20.0 if arith == "strict" else 5.0
    mk = lambda: cases.lyap_lorenz(tmax, rho=rho, arith=arith)
    hot = mk().run()
    assert hot.totals["kernel_kind"] == KERNEL_ODE_FIXED_PERIODIC
    gen = run_general(mk)
    for a in ("steps", "events", "rhs_evals", "status", "rejected"):
        assert np.array_equal(getattr(hot, a), getattr(gen, a)), a
    assert np.array_equal(hot.t, gen.t)
    if arith == "strict":
        # same operation order in both kernels: bit-identical state and log sums
        assert np.array_equal(hot.p, gen.p) and np.array_equal(hot.aux, gen.aux)
        o = oracle.run_solver(mk(), n_threads=8)
        assert_same_counters(hot, o)
        assert np.array_equal(hot.p, o.p)                      # QR: sqrt and division are IEEE on both sides
        assert np.max(np.abs(hot.aux - o.aux)) < 1e-9          # log() is within 1-2 ulp of glibc
    else:
        # FAST: pre-scaled rows + FMA vs the general kernel's contraction: rounding-level differences
        # amplified by the chaotic flow (leading exponent up to ~1.4 at rho = 40: e^7 ~ 1e3 over t = 5)
        assert np.max(np.abs(hot.p[:, :3] - gen.p[:, :3])) < 1e-9
        assert np.max(np.abs(hot.aux - gen.aux)) < 1e-9


def test_periodic_kernel_two_locators_and_event_log(oracle):
    # two Periodic locators (different periods and an offset) + the event log; times coincide at
    # t = 1.5, 3.0, ...: the strictly-earliest / lowest-index rule and DedupLocF decide who fires
    lam = np.array([-1.0, -0.5, 0.25, 0.5]).reshape(-1, 1)

    def mk():
        return (d.Solver.new().initial(np.ones((4, 1))).interval(0.0, 7.0).stepsize(0.04)
                .equation(d.systems.Linear1, lam)
                .on_mut(d.Periodic(0.3, 0.0), "renormalize").on_mut(d.Periodic(0.5, 0.25), "renormalize")
                .log_events(d.EventLog(64)))
    hot = mk().run()
    assert hot.totals["kernel_kind"] == KERNEL_ODE_FIXED_PERIODIC
    o = oracle.run_solver(mk(), n_threads=2)
    assert_same_counters(hot, o)
    assert np.array_equal(hot.event_n, o.event_n)
    ne = int(hot.event_n[0])
    assert ne > 30
    assert np.array_equal(hot.event_t[:, :ne], o.event_t[:, :ne])
    assert np.array_equal(hot.event_index[:, :ne], o.event_index[:, :ne])
    assert np.array_equal(hot.p, o.p) and np.array_equal(hot.t, o.t)
    assert np.max(np.abs(hot.aux - o.aux)) < 1e-12
    gen = run_general(mk)
    assert np.array_equal(hot.p, gen.p) and np.array_equal(hot.aux, gen.aux)
    assert np.array_equal(hot.event_t[:, :ne], gen.event_t[:, :ne])


@pytest.mark.parametrize("h,period,offset,t1", [
    (0.25, 0.5, 0.0, 60.0),      # boundaries exactly on grid points (all values dyadic)
    (0.01, 0.1, 0.0, 40.0),      # inexact period and step: floor((t - o) / T) sits within an ulp of an integer
    (0.037, 0.5, 0.123, 500.0),  # incommensurate, long horizon (the quiet-interval margin grows with t)
    (0.3, 0.1, 0.05, 12.0),      # several periods per step: every step has an event
])
def test_periodic_kernel_quiet_interval_same_events_as_oracle(oracle, h, period, offset, t1):
    # the kernel skips the floor((t - o) / T) test while t_next is provably below the next period boundary
    # (dev/ode_fixed.cuh, `t_quiet`) and predicts events instead of stepping and undoing: event times, indices
    # and every counter must still be the reference's (loc.rs:318-334, solver.rs:299-308), bit for bit
    lam = np.array([-0.01, 0.003, 0.0]).reshape(-1, 1)

    def mk():
        return (d.Solver.new().initial(np.ones((3, 1))).interval(0.0, t1).stepsize(h)
                .equation(d.systems.Linear1, lam).on_mut(d.Periodic(period, offset), "renormalize")
                .log_events(d.EventLog(2048)))
    hot = mk().run()
    assert hot.totals["kernel_kind"] == KERNEL_ODE_FIXED_PERIODIC
    o = oracle.run_solver(mk(), n_threads=2)
    assert_same_counters(hot, o)
    assert np.array_equal(hot.event_n, o.event_n) and int(hot.event_n[0]) > 10
    ne = min(int(hot.event_n[0]), 2048)
    assert np.array_equal(hot.event_t[:, :ne], o.event_t[:, :ne])
    assert np.array_equal(hot.p, o.p) and np.array_equal(hot.t, o.t)


def test_periodic_kernel_two_boundaries_inside_one_step(oracle):
    # two Periodic locators whose boundaries regularly fall inside the SAME step: the event of the earlier one
    # cuts the step short and the later one's boundary still lies ahead — the quiet-interval bound computed for
    # the whole step must not survive the cut (regression: fuzz seed 110 at DFB_FUZZ_SCALE=8 lost those events)
    lam = np.array([-0.3, 0.2, 0.05, -0.01]).reshape(-1, 1)
    for h, p1, o1, p2 in ((0.1, 0.25, 0.03, 0.375), (0.0123456789, 0.05, 0.004, 0.075), (0.5, 0.3, 0.1, 0.45)):
        def mk():
            return (d.Solver.new().initial(np.ones((4, 1))).interval(0.0, 40.0 * h).stepsize(h)
                    .equation(d.systems.Linear1, lam)
                    .on_mut(d.Periodic(p1, o1), "renormalize").on_mut(d.Periodic(p2, 0.0), "renormalize")
                    .log_events(d.EventLog(1024)))
        hot = mk().run()
        assert hot.totals["kernel_kind"] == KERNEL_ODE_FIXED_PERIODIC
        o = oracle.run_solver(mk(), n_threads=2)
        assert_same_counters(hot, o)
        assert np.array_equal(hot.event_n, o.event_n) and int(hot.event_n[0]) > 10
        ne = min(int(hot.event_n[0]), 1024)
        assert np.array_equal(hot.event_t[:, :ne], o.event_t[:, :ne])
        assert np.array_equal(hot.event_index[:, :ne], o.event_index[:, :ne])
        assert np.array_equal(hot.p, o.p) and np.array_equal(hot.t, o.t)
        assert np.max(np.abs(hot.aux - o.aux)) < 1e-12


def test_periodic_kernel_stop_integration_is_per_lane(oracle):
    # the egg-billiard reflection callback (examples/egg-billiard.rs:22-38) on a Periodic clock: member
    # i calls stop_integration at its (i % 7 + 1)-th event, so lanes of one warp stop at different
    # times while the rest of the warp keeps stepping
    n = 70
    max_refl = (np.arange(n) % 7 + 1).astype(np.float64).reshape(-1, 1)
    phi = 2.0 * np.pi * np.arange(n) / n
    y0 = np.stack([np.zeros(n), np.full(n, 0.1), 0.5 * np.cos(phi), 0.5 * np.sin(phi)], axis=1)

    def mk():
        return (d.Solver.new().initial(y0).interval(0.0, 5.0).stepsize(0.05)
                .equation(d.systems.EggBilliard, max_refl).on_mut(d.Periodic(0.4, 0.0), "reflect"))
    hot = mk().run()
    assert hot.totals["kernel_kind"] == KERNEL_ODE_FIXED_PERIODIC
    o = oracle.run_solver(mk(), n_threads=4)
    assert_same_counters(hot, o)
    assert np.all(hot.status == 64) and np.all(np.isinf(hot.t))       # DFB_TRAJ_STOPPED, t = +inf
    assert np.array_equal(hot.events, (np.arange(n) % 7 + 1).astype(np.uint64))
    assert np.array_equal(hot.p, o.p) and np.array_equal(hot.aux, o.aux)
    gen = run_general(mk)
    assert np.array_equal(hot.p, gen.p) and np.array_equal(hot.steps, gen.steps)


def test_periodic_kernel_scalar_and_matrix_lyapunov_kinds():
    # the reference's other Lyapunov tests also take the periodic kernel
    assert cases.lyap_linear_1_const([-1.0, 2.0], 10.0).run().totals["kernel_kind"] == KERNEL_ODE_FIXED_PERIODIC
    assert cases.lyap_linear_1_sin([-1.0, 2.0], 10.0).run().totals["kernel_kind"] == KERNEL_ODE_FIXED_PERIODIC
    m = [cases.real_jordan_matrix([2., 0., -2.])]
    assert cases.lyap_linear_3(m, 10.0).run().totals["kernel_kind"] == KERNEL_ODE_FIXED_PERIODIC

# ---------------------------------------------------------------- full-size properties
def test_full_size_lorenz_properties():
    # BASELINE config 2 at full size (2^20 members, t in [0,50], h = 1/250, rktp64, FAST)
    n = 1 << 20
    rho = 20.0 + 20.0 * np.arange(n) / (n - 1)
    rho[n // 2:] = rho[: n // 2]          # duplicate the first half: identical inputs -> identical bits
    g = cases.lorenz(n=n, t1=50.0, arith="fast", rho=rho).run()
    assert np.all(g.status == 0)
    # 12500 steps of 1/250 accumulate to one ulp short of 50, then one closing step (as upstream)
    assert np.all(g.steps == 12501) and np.all(g.rhs_evals == 12501 * 7 + 1)
    assert np.all(g.t == g.t[0]) and abs(g.t[0] - 50.0) < 1e-9
    assert np.array_equal(g.p[: n // 2], g.p[n // 2:])
    assert np.all(np.isfinite(g.p)) and np.max(np.abs(g.p)) < 200.0   # stays on the attractor
    again = cases.lorenz(n=n, t1=50.0, arith="fast", rho=rho).run()
    assert np.array_equal(again.p, g.p)                               # deterministic


def test_full_size_lorenz_strict_subsample_is_bit_exact_at_t50(oracle):
    # the same 2^20-member launch in STRICT arithmetic, and 64 members spread over the sweep re-integrated
    # by the CPU oracle to t = 50 (12 501 steps): STRICT is the reference's rounding, so chaos is no excuse —
    # every sampled member must agree to the last bit at full size, wherever it sat in the grid
    n = 1 << 20
    rho = 20.0 + 20.0 * np.arange(n) / (n - 1)
    g = cases.lorenz(n=n, t1=50.0, arith="strict", rho=rho).run()
    assert g.totals["kernel_kind"] == KERNEL_ODE_FIXED and np.all(g.status == 0) and np.all(g.steps == 12501)
    idx = np.unique(np.concatenate([np.linspace(0, n - 1, 60).astype(np.int64), [1, 31, 32, n - 2]]))
    o = oracle.run_solver(cases.lorenz(n=len(idx), t1=50.0, arith="strict", rho=rho[idx]), n_threads=8)
    assert np.array_equal(g.p[idx], o.p) and np.array_equal(g.t[idx], o.t)
    assert np.array_equal(g.steps[idx], o.steps) and np.array_equal(g.rhs_evals[idx], o.rhs_evals)


def test_full_size_linearity_property():
    # y' = A y is linear: scaling y0 by a power of two scales the result exactly, in both arithmetics
    n = 1 << 18
    rng = np.random.default_rng(7)
    a = rng.uniform(-1.0, 1.0, size=(n, 9))
    y0 = rng.uniform(-1.0, 1.0, size=(n, 3))
    for arith in ("strict", "fast"):
        mk = lambda s: (d.Solver.new().initial(y0 * s).equation(d.systems.Linear3, a)
                        .interval(0.0, 2.0).stepsize(0.01).arith(arith))
        g1, g4 = mk(1.0).run(), mk(4.0).run()
        assert np.array_equal(g1.p * 4.0, g4.p)


def test_full_size_bouncing_ball_properties():
    # BASELINE configs[3] at full size: 2^19 members, x0 = 1 + i/n, t in [0, 8], FAST arithmetic.
    # Between impacts the flight is a parabola that rktp64 integrates exactly up to rounding, so the
    # impact times are known in closed form: t_1 = (v0 + sqrt(v0^2 + 2 g x0)) / g, then the flight
    # times shrink by the restitution factor k = 0.9.
    n = 1 << 19
    g = cases.bouncing_ball(n=n, t1=8.0, arith="fast", events=8).run()
    assert g.totals["kernel_kind"] == d.abi.KERNEL_ODE_EVENT and g.totals["n_failed"] == 0
    x0 = 1.0 + np.arange(n) / n
    v0, grav, k = -0.01, 9.8, 0.9
    impacts = g.event_t[np.arange(n), np.argmax(g.event_index[:, :8] == 1, axis=1)]   # first "bounce" event
    t1 = (v0 + np.sqrt(v0 * v0 + 2.0 * grav * x0)) / grav
    assert np.max(np.abs(impacts - t1)) < 1e-12
    # number of impacts before t = 8: t_m = t_1 + (2 v_1 / g) k (1 - k^(m-1)) / (1 - k), v_1 = sqrt(v0^2 + 2 g x0)
    v1 = np.sqrt(v0 * v0 + 2.0 * grav * x0)
    m = np.arange(1, 200)[None, :]
    tm = t1[:, None] + (2.0 * v1[:, None] / grav) * k * (1.0 - k ** (m - 1)) / (1.0 - k)
    n_imp = (tm < 8.0).sum(axis=1)
    apex = g.events.astype(np.int64) - n_imp          # the no-op `zero(dx)` locator fires once per flight
    assert np.all(np.abs(apex - n_imp) <= 1)
    assert np.all(g.p[:, 0] > -1e-9) and np.all(np.abs(g.t - 8.0) < 1e-12)


def test_full_size_egg_billiard_properties():
    # 2^19 members, 200 reflections each: the flow between reflections is free flight and a
    # reflection is an isometry of the velocity, so speed is conserved; the ball stays in the egg.
    n = 1 << 19
    g = cases.egg_billiard(n=n, reflections=200, arith="fast", events=0).run()
    assert g.totals["kernel_kind"] == d.abi.KERNEL_ODE_EVENT and g.totals["n_failed"] == 0
    assert np.all(g.status == d.abi.TRAJ_STOPPED) and np.all(g.events == 200) and np.all(g.aux[:, 0] == 200.0)
    speed = np.hypot(g.p[:, 2], g.p[:, 3])
    assert np.max(np.abs(speed - 0.5)) < 1e-12
    x, y = g.p[:, 0], g.p[:, 1]
    assert np.max(np.abs(x * x + y * y - y ** 3 / 3.0 - 1.0)) < 1e-9   # stopped ON the boundary


def test_full_size_lyapunov_batch_properties():
    # BASELINE configs[4]: 2^16-member rho grid; the exponents sum to the trace of the Jacobian,
    # -(sigma + 1 + beta), for every member; the leading one is positive on the chaotic band
    n, tmax = 1 << 16, 50.0
    rho = np.linspace(24.0, 40.0, n)
    g = cases.lyap_lorenz(tmax, rho=rho, arith="fast").run()
    assert g.totals["kernel_kind"] == d.abi.KERNEL_ODE_FIXED_PERIODIC and g.totals["n_failed"] == 0
    lam = g.aux / tmax
    assert np.max(np.abs(lam.sum(axis=1) + (10.0 + 1.0 + 8.0 / 3.0))) < 1e-6
    assert np.all(g.events == 100) and np.all(g.steps == g.steps[0])
    assert np.mean(lam[:, 0] > 0.3) > 0.9


def test_on_start_mut_runs_once_before_the_first_on_step(oracle):
    # Solver::on_start_mut (solver.rs:190-197, run at solver.rs:289): the registered callback mutates
    # the initial State; the first trace sample (on_step at t_init, solver.rs:290) already sees it
    lam = np.linspace(-1.0, 1.0, 37).reshape(-1, 1)
    y0 = np.linspace(0.5, 4.0, 37).reshape(-1, 1)
    for extra in ("none", "periodic", "statefn"):
        def mk():
            s = (d.Solver.new().initial(y0).equation(d.systems.Linear1, lam).interval(0.0, 2.0).stepsize(0.05)
                 .on_start_mut("renormalize").on_step(d.Trace(1, 64)))
            if extra == "periodic":
                s = s.on_mut(d.Periodic(0.3, 0.0), "renormalize")
            if extra == "statefn":
                s = s.on(d.Locator.step().every(7), None)
            return s
        g = mk().run()
        o = oracle.run_solver(mk(), n_threads=2)
        assert np.all(g.trace_p[:, 0, 0] == 1.0)                       # |y0| / |y0|
        assert_same_counters(g, o)
        assert_same_trace(g, o, exact=True)
        assert np.array_equal(g.p, o.p) and np.max(np.abs(g.aux - o.aux)) < 1e-13
        assert np.all(np.abs(g.aux[:, 0] - np.log(y0[:, 0])) < 1e-12) or extra == "periodic"
    # the register-only hot paths do not take callbacks: an on_start_mut problem goes to a kernel that does
    plain = (d.Solver.new().initial(y0).equation(d.systems.Linear1, lam).interval(0.0, 2.0).stepsize(0.05)
             .on_start_mut("renormalize").arith("fast").run())
    assert plain.totals["kernel_kind"] == KERNEL_GENERAL
    ref = (d.Solver.new().initial(np.ones_like(y0)).equation(d.systems.Linear1, lam).interval(0.0, 2.0)
           .stepsize(0.05).arith("fast").run())
    assert np.max(np.abs(plain.p - ref.p)) < 1e-13 and np.all(np.abs(plain.aux[:, 0] - np.log(y0[:, 0])) < 1e-12)
