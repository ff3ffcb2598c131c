"""The CPU oracle against the reference's OWN test assertions.

Every test here re-creates one live `#[test]` of /root/reference/tests/*.rs (or an
in-crate unit test) through the oracle and applies the reference's assertion
verbatim: this is what pins the oracle, since the Rust crate cannot be built here.
"""
import math

import numpy as np
import pytest

import cases


def trace_of(r, i=0):
    n = int(r.trace_n[i])
    return r.trace_t[i, :n], r.trace_p[i, :n]


# ---------------------------------------------------------------- delay_propagation.rs
def test_delay_constant_1(oracle):  # tests/delay_propagation.rs:6-25
    r = oracle.run_solver(cases.delay_constant_1())
    ts, _ = trace_of(r)
    assert list(ts) == [0.0, 0.75, 1.0, 1.75, 2.0, 2.75, 3.0, 3.75, 4.0, 4.75, 5.0]


def test_delay_constant_2(oracle):  # :27-52
    r = oracle.run_solver(cases.delay_constant_2())
    ts, _ = trace_of(r)
    s = 7.0 / 8.0
    assert list(ts) == [0.0, s, 1.0, 1.25, 2., 2.25, 2.5, 3.0, 3.25, 3.5, 3.75, 4., 4.25, 4.5,
                        4.75, 5.0]


def test_delay_constant_1_smoothing(oracle):  # :54-72
    r = oracle.run_solver(cases.delay_constant_1_smoothing())
    ts, _ = trace_of(r)
    assert list(ts) == [0., 0.75, 1., 1.75, 2., 2.75, 3., 3.75, 4.5, 5.25, 6.]


def test_delay_pantograph(oracle):  # :74-94
    r = oracle.run_solver(cases.delay_pantograph())
    ts, _ = trace_of(r)
    assert r.trace_n[0] < 16384
    for i in range(1, 11):
        assert 2.0 ** i in set(ts)


# ---------------------------------------------------------------- equation_types.rs
def test_eq_constant(oracle):  # :6-17  exact equality with [t, -t, 0]
    r = oracle.run_solver(cases.eq_constant())
    ts, ps = trace_of(r)
    assert len(ts) == 41
    for t, p in zip(ts, ps):
        assert list(p) == [t, -t, 0.0]


def test_eq_time(oracle):  # :19-65  exact equality with [0, t, t*t/2]
    r = oracle.run_solver(cases.eq_time())
    ts, ps = trace_of(r)
    assert len(ts) == 21
    for t, p in zip(ts, ps):
        assert list(p) == [0.0, t, t * t / 2.0]


def test_eq_ode_exponent(oracle):  # :67-79
    r = oracle.run_solver(cases.eq_ode_exponent())
    ts, ps = trace_of(r)
    assert len(ts) > 200
    assert np.all(np.abs(ps[:, 0] - np.exp(-ts)) < 1e-14)


def test_eq_ode_harmonic(oracle):  # :81-96
    r = oracle.run_solver(cases.eq_ode_harmonic())
    ts, ps = trace_of(r)
    assert len(ts) > 250
    assert np.all(np.abs(ps[:, 0] - np.sin(ts)) < 1e-13)


def test_eq_odet_lin(oracle):  # :98-113
    r = oracle.run_solver(cases.eq_odet_lin())
    ts, ps = trace_of(r)
    assert len(ts) > 450
    assert np.all(np.abs(ps[:, 0] - ts ** -2.0) < 1e-13)


def test_eq_dde_sin(oracle):  # :115-136
    r = oracle.run_solver(cases.eq_dde_sin())
    assert r.status[0] == 0
    ts, ps = trace_of(r)
    assert len(ts) > 500
    assert np.all(np.abs(ps[:, 0] - np.sin(ts)) < 1e-11)


def test_eq_ndde_sin(oracle):  # :138-158
    r = oracle.run_solver(cases.eq_ndde_sin())
    assert r.status[0] == 0
    ts, ps = trace_of(r)
    assert len(ts) > 500
    assert np.all(np.abs(ps[:, 0] - np.sin(ts)) < 1e-11)


def test_dde_value_only_history_panics_on_derivative(oracle, dfb):
    # initial_condition.rs:33: InitFn(f, ()) has no derivative -> `unimplemented!`
    import diffurch_b200 as d
    s = cases.eq_ndde_sin().initial(d.InitFn(d.SIN))
    r = oracle.run_solver(s)
    assert r.status[0] & d.abi.TRAJ_NO_DERIVATIVE


def test_dde_delay_not_larger_than_step_panics(oracle, dfb):
    # state.rs:172-177: StateRef::p never sees the in-flight step, so delay <= h panics
    import diffurch_b200 as d
    s = cases.eq_dde_sin().stepsize(1.5)
    r = oracle.run_solver(s)
    assert r.status[0] & d.abi.TRAJ_HISTORY_NOT_READY


# ---------------------------------------------------------------- lyapunov_exponents.rs
TMAX_FAST = [10., 100., 1000., 10_000.]
TMAX_ALL = TMAX_FAST + [100_000., 200_000.]
LAMBDAS = [-5., -4., -3., -2., -1., 1., 2., 3., 4., 5.]


def test_lyap_linear_1_const(oracle):  # :9-45
    r = oracle.run_solver(cases.lyap_linear_1_const([0.0], 100.0))
    assert r.aux[0, 0] / 100.0 == 0.0
    for tmax in TMAX_ALL:
        r = oracle.run_solver(cases.lyap_linear_1_const(LAMBDAS, tmax), n_threads=4)
        assert np.all(r.status == 0)
        computed = r.aux[:, 0] / tmax
        assert np.all(np.abs(np.array(LAMBDAS) - computed) < 2.0 / tmax), (tmax, computed)


def test_lyap_linear_1_sin(oracle):  # :47-83
    r = oracle.run_solver(cases.lyap_linear_1_sin([0.0], 100.0))
    assert r.aux[0, 0] / 100.0 == 0.0
    for tmax in TMAX_ALL:
        r = oracle.run_solver(cases.lyap_linear_1_sin(LAMBDAS, tmax), n_threads=4)
        computed = r.aux[:, 0] / tmax
        assert np.all(np.abs(np.array(LAMBDAS) - computed) < 10.0 / tmax), (tmax, computed)


REAL_CASES = [[2., 0., -2.], [3., 1., -2.], [5., 3., -2.], [3., 2., 1.], [2., 2., 1.],
              [4., -1., -1.], [3., 2., 2.], [0., 0., 0.]]


def test_lyap_linear_const_3_real(oracle):  # :85-164
    mats = [cases.real_jordan_matrix(e) for e in REAL_CASES]
    for tmax in TMAX_FAST:
        r = oracle.run_solver(cases.lyap_linear_3(mats, tmax, "qr_ln"), n_threads=4)
        assert np.all(r.status == 0)
        for e, lam in zip(REAL_CASES, r.aux / tmax):
            norm = np.max(np.abs(np.array(e) - lam))
            tol = 40.0 if (e[0] == e[1] or e[1] == e[2]) else 2.0
            assert norm < tol / tmax, (tmax, e, lam)


COMPLEX_CASES = [(4., (-3., 0.)), (-1., (5., -5.)), (0., (-2., 3.)), (2., (1., -4.)),
                 (-5., (0., 2.)), (3., (4., -1.)), (-4., (-5., 5.)), (1., (3., -2.)),
                 (5., (2., 4.)), (-2., (-1., 1.))]


def test_lyap_linear_const_3_complex(oracle):  # :166-240
    mats = [cases.complex_matrix(l1, re, im) for l1, (re, im) in COMPLEX_CASES]
    for tmax in TMAX_FAST:
        r = oracle.run_solver(cases.lyap_linear_3(mats, tmax, "qr_ln_abs"), n_threads=4)
        for (l1, (re, im)), lam in zip(COMPLEX_CASES, r.aux / tmax):
            true = [l1, re, re] if l1 > re else [re, re, l1]
            norm = np.max(np.abs(np.array(true) - lam))
            assert norm < 2.0 / tmax, (tmax, true, lam)


REF_LORENZ = np.array([0.90566, 0.00000, -14.57233])


@pytest.mark.parametrize("tmax", [100., 1000., 10_000.])
def test_lyap_lorenz(oracle, tmax):  # :242-311
    r = oracle.run_solver(cases.lyap_lorenz(tmax))
    assert r.status[0] == 0
    lam = r.aux[0] / tmax
    err = np.max(np.abs(lam - REF_LORENZ))
    assert err < 0.00007 + 50.0 / tmax, (tmax, lam)


@pytest.mark.slow
def test_lyap_lorenz_long(oracle):  # the tmax = 1e5 leg of :254
    tmax = 100_000.
    r = oracle.run_solver(cases.lyap_lorenz(tmax))
    lam = r.aux[0] / tmax
    assert np.max(np.abs(lam - REF_LORENZ)) < 0.00007 + 50.0 / tmax, lam


# ---------------------------------------------------------------- util/partition_point.rs
def _ppl(oracle, data, start, thr):
    import ctypes as C
    arr = (C.c_int64 * len(data))(*data)
    calls = C.c_uint64()
    i = oracle.lib().orc_partition_point_linear(arr, len(data), start, thr, C.byref(calls))
    return i, calls.value


def test_partition_point_linear(oracle):  # src/util/partition_point.rs:56-129
    dq = [1, 2, 3, 3, 5, 6, 7]
    assert _ppl(oracle, dq, 0, 5)[0] == 4       # test1
    assert _ppl(oracle, dq, 4, 10)[0] == 7      # test2
    assert _ppl(oracle, dq, 4, -4)[0] == 0      # test3
    assert _ppl(oracle, dq, 0, 5) == (4, 5)     # count_1
    assert _ppl(oracle, dq, 2, 5) == (4, 3)     # count_2
    assert _ppl(oracle, dq, 6, 5) == (4, 4)     # count_3
    assert _ppl(oracle, dq, 666, 5) == (4, 4)   # count_4
    assert _ppl(oracle, [], 3, 5)[0] == 0


# ---------------------------------------------------------------- restatement-derived anchors
def test_bouncing_ball_anchors(oracle):
    # SURVEY §8c [scratch-verified]: first impact = (-0.01 + sqrt(19.6001)) / 9.8
    r = oracle.run_solver(cases.bouncing_ball())
    assert r.status[0] == 0
    idx = r.event_index[0, :r.event_n[0]]
    tt = r.event_t[0, :r.event_n[0]]
    impacts = tt[idx == 1]
    analytic = (-0.01 + math.sqrt(0.0001 + 2 * 9.8 * 1.0)) / 9.8
    assert abs(impacts[0] - analytic) < 2e-15
    assert len(impacts) == 78
    assert r.steps[0] + 1 == 1815  # on_step invocations incl. t_init


def test_egg_billiard_anchors(oracle):
    r = oracle.run_solver(cases.egg_billiard(reflections=200))
    assert r.status[0] == 64  # stopped by stop_integration
    assert math.isinf(r.t[0])
    assert r.aux[0, 0] == 200.0
    assert r.steps[0] + 1 == 1497
    v = r.p[0]
    assert abs(math.hypot(v[2], v[3]) - 0.5) < 1e-14  # speed conserved by the reflection
