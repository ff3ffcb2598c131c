"""Problem builders shared by the oracle tests (CPU) and the parity tests (GPU).

Each function returns a ``diffurch_b200.Solver`` configured exactly like one of the
reference's live tests / examples (file:line cited), optionally widened to an
ensemble.  The same builder object feeds the CUDA path (``.run()``) and the oracle
(``oracle.run_solver``), so both sides see one input.
"""
import math

import numpy as np

import diffurch_b200 as d
from diffurch_b200 import RK, InitFn, Locator, Periodic, Solver, Trace, systems
from diffurch_b200.solver import EventLog


# ---- tests/delay_propagation.rs ------------------------------------------------
def delay_constant_1():  # :6-25
    return (Solver.new().rk(RK.euler()).initial([0.0]).equation(systems.Zero1)
            .initial_disco([(0.0, 0)]).with_const_delay(1.0, 0).interval(0.0, 5.0)
            .stepsize(0.75).on_step(Trace(1, 64)))


def delay_constant_2():  # :27-52
    return (Solver.new().rk(RK.euler()).stepsize(7.0 / 8.0).initial([0.0]).equation(systems.Zero1)
            .initial_disco([(0.0, 0)]).interval(0.0, 5.0).with_const_delay(1.0, 0)
            .with_const_delay(1.25, 0).max_delay(1.25).on_step(Trace(1, 64)))


def delay_constant_1_smoothing():  # :54-72
    return (Solver.new().rk(RK.rk4()).initial([0.0]).equation(systems.Zero1)
            .initial_disco([(0.0, 1)]).with_const_delay(1.0, 1).interval(0.0, 6.0)
            .stepsize(0.75).on_step(Trace(1, 64)))


def delay_pantograph():  # :74-94
    return (Solver.new().rk(RK.rk4()).stepsize(0.123456789101112).initial([0.0])
            .equation(systems.Zero1).initial_disco([(1.0, 0)]).interval(1.0, 2.0 ** 10)
            .with_delayed_argument("half_time", 0).on_step(Trace(1, 16384))
            .max_delay(math.inf))


# ---- tests/equation_types.rs ---------------------------------------------------
def eq_constant():  # :6-17
    return (Solver.new().rk(RK.euler()).stepsize(0.25).initial([0.0, 0.0, 0.0])
            .interval(0.0, 10.0).equation(systems.Constant3).on_step(Trace(1, 64)))


def eq_time():  # :19-41 (and time2 :43-65, same system)
    return (Solver.new().interval(0.0, 10.0).rk(RK.midpoint()).stepsize(0.5)
            .initial([0.0, 0.0, 0.0]).equation(systems.Time3).on_step(Trace(1, 64)))


def eq_ode_exponent(n=1):  # :67-79   y' = -y  (Linear1 with lambda = -1: p * -1 == -p bit for bit)
    return (Solver.new().initial(np.ones((n, 1))).interval(0.0, 10.0).rk(RK.rktp64())
            .stepsize(0.05).equation(systems.Linear1, -np.ones((n, 1))).on_step(Trace(1, 256)))


def eq_ode_harmonic():  # :81-96
    return (Solver.new().rk(RK.rktp64()).stepsize(0.04).initial([0.0, 1.0])
            .equation(systems.Harmonic).interval(0.0, 10.0).on_step(Trace(1, 512)))


def eq_odet_lin():  # :98-113
    return (Solver.new().rk(RK.rktp64()).stepsize(0.02).interval(1.0, 10.0)
            .initial(InitFn(d.INV_SQUARE)).equation(systems.PowerDecay).on_step(Trace(1, 512)))


def dde_params(k, tau=1.0):
    k = np.atleast_1d(np.asarray(k, dtype=np.float64))
    a = k / np.tan(k * tau)
    b = -k / np.sin(k * tau)
    return np.stack([a, b, np.full_like(k, tau), k], axis=1)


def eq_dde_sin(k=1.0, t1=10.0, trace=True, arith="strict"):  # :115-136
    s = (Solver.new().max_delay(1.0).initial(InitFn(d.SIN)).interval(0.0, t1).stepsize(0.02)
         .equation(systems.DdeLinear, dde_params(k)).arith(arith))
    if trace:
        s = s.on_step(Trace(1, int(t1 / 0.02) + 8))
    return s


def ndde_params(k, tau=1.0):
    k = np.atleast_1d(np.asarray(k, dtype=np.float64))
    a = -k * np.sin(k * tau)
    b = np.cos(k * tau)
    return np.stack([a, b, np.full_like(k, tau), k], axis=1)


def eq_ndde_sin(k=1.0, t1=10.0):  # :138-158
    return (Solver.new().initial(InitFn(d.SIN, d.SIN_DERIV)).equation(systems.NddeLinear, ndde_params(k))
            .interval(0.0, t1).max_delay(1.0).stepsize(0.02).on_step(Trace(1, int(t1 / 0.02) + 8)))


# ---- tests/lyapunov_exponents.rs -------------------------------------------------
def lyap_linear_1_const(eigenvalues, tmax):  # :9-45
    lam = np.atleast_1d(np.asarray(eigenvalues, dtype=np.float64)).reshape(-1, 1)
    return (Solver.new().initial(np.ones((lam.shape[0], 1))).interval(0.0, tmax)
            .equation(systems.Linear1, lam).on_mut(Periodic(0.3, 0.0), "renormalize"))


def lyap_linear_1_sin(eigenvalues, tmax):  # :47-83
    lam = np.atleast_1d(np.asarray(eigenvalues, dtype=np.float64)).reshape(-1, 1)
    return (Solver.new().initial(np.ones((lam.shape[0], 1))).interval(0.0, tmax)
            .equation(systems.Linear1Sin, lam).on_mut(Periodic(1.3, 0.0), "renormalize"))


def real_jordan_matrix(eigenvalues):  # :89-109
    c = np.array([[1., 3., -2.], [4., 2., -1.], [-1., 1., -5.]])
    j = np.diag(np.asarray(eigenvalues, dtype=np.float64))
    if j[0, 0] == j[1, 1]:
        j[0, 1] = 1.0
    if j[1, 1] == j[2, 2]:
        j[1, 2] = 1.0
    return c @ j @ np.linalg.inv(c)


def complex_matrix(l1, re, im):  # :170-183
    c = np.array([[1., -3., -2.], [4., -2., -1.], [-1., 1., -5.]])
    j = np.array([[l1, 0., 0.], [0., re, -im], [0., im, re]])
    return c @ j @ np.linalg.inv(c)


def lyap_linear_3(a_matrices, tmax, callback="qr_ln"):  # :111-131 / :187-210
    a = np.asarray(a_matrices, dtype=np.float64).reshape(-1, 9)  # row-major
    eye = np.tile(np.eye(3).reshape(1, 9), (a.shape[0], 1))      # column-major identity == identity
    return (Solver.new().initial(eye).interval(0.0, tmax).equation(systems.Linear3Matrix, a)
            .on_mut(Periodic(1.0, 0.0), callback))


def lyap_lorenz(tmax, rho=28.0, arith="strict"):  # :242-311
    rho = np.atleast_1d(np.asarray(rho, dtype=np.float64))
    n = rho.shape[0]
    y0 = np.tile(np.array([10., 15., 20., 1., 0., 0., 0., 1., 0., 0., 0., 1.]), (n, 1))
    prm = np.stack([np.full(n, 10.0), rho, np.full(n, 8.0 / 3.0)], axis=1)
    return (Solver.new().stepsize(0.005).initial(y0).interval(0.0, tmax)
            .equation(systems.LorenzVariational, prm).on_mut(Periodic(0.5, 0.0), "qr").arith(arith))


# ---- examples -------------------------------------------------------------------
def lorenz(n=1, t1=50.0, h=1.0 / 250.0, rk=None, arith="strict", rho=None, trace=None):
    # examples/ode-chaos-lorenz.rs:17-33; rho sweep of SURVEY §8(d) config 2
    if rho is None:
        rho = np.array([28.0]) if n == 1 else 20.0 + 20.0 * np.arange(n) / (n - 1)
    prm = np.stack([np.full(n, 10.0), rho, np.full(n, 8.0 / 3.0)], axis=1)
    s = (Solver.new().initial(np.tile([1.0, 2.0, 20.0], (n, 1))).equation(systems.Lorenz, prm)
         .interval(0.0, t1).stepsize(h).arith(arith))
    if rk is not None:
        s = s.rk(rk)
    if trace is not None:
        s = s.on_step(trace)
    return s


def bouncing_ball(n=1, t1=8.58, arith="strict", events=256, trace=None):
    # examples/bouncing-ball.rs:21-33; x0 sweep of SURVEY §8(d) config 4
    x0 = np.ones(n) if n == 1 else 1.0 + np.arange(n) / n
    y0 = np.stack([x0, np.full(n, -0.01)], axis=1)
    prm = np.tile([9.8, 0.9], (n, 1))
    s = (Solver.new().initial(y0).equation(systems.BouncingBall, prm).interval(0.0, t1)
         .stepsize(0.005).on(Locator.zero("dx"), None).on_mut(Locator.below_zero("x"), "bounce")
         .arith(arith))
    if events:
        s = s.log_events(EventLog(events))
    if trace is not None:
        s = s.on_step(trace)
    return s


def egg_billiard(n=1, reflections=200, arith="strict", events=256, phi=None):
    # examples/egg-billiard.rs:12-40 (the example's own y0 when n == 1)
    if n == 1 and phi is None:
        y0 = np.array([[0.0, 0.1, 0.3, 0.4]])
    else:
        if phi is None:
            phi = 2.0 * np.pi * np.arange(n) / n
        y0 = np.stack([np.zeros(n), np.full(n, 0.1), 0.5 * np.cos(phi), 0.5 * np.sin(phi)], axis=1)
    prm = np.full((n, 1), float(reflections))
    s = (Solver.new().initial(y0).equation(systems.EggBilliard, prm).interval(0.0, math.inf)
         .stepsize(0.5).on_mut(Locator.above_zero("boundary"), "reflect").arith(arith)
         .max_steps(100 * reflections + 1000))
    if events:
        s = s.log_events(EventLog(events))
    return s


def relay_msign(arith="strict"):  # examples/ode-2-relay-msign.rs:15-26
    return (Solver.new().initial([0.25, 0.0]).equation(systems.RelayMsign).interval(0.5, 10.5)
            .stepsize(0.09).on(Locator.zero("x"), None).on_step(Trace(1, 512)).arith(arith)
            .log_events(EventLog(64)))


def dde_relay_clamp(arith="strict", t1=100.0):  # examples/dde-relay-clamp.rs:19-36
    alpha, sigma, T, v0 = 5.0, 0.4, 1.0, 1.0
    return (Solver.new().initial([1.0, -alpha * v0]).equation(systems.DdeRelayClamp, [alpha, sigma, T])
            .interval(0.0, t1 * T).stepsize(0.012).with_const_delay(T, 2)
            .on_step(Trace(1, 16384)).on(Locator.zero("abs_delayed_x_minus_1"), None)
            .arith(arith).log_events(EventLog(512)))


def dde_relay_clamp_ensemble(n, arith="fast", t1=30.0, events=0, trace=None):
    # examples/dde-relay-clamp.rs:19-36 as an ensemble over alpha (DDE + propagation + located events)
    alpha = np.linspace(3.0, 7.0, n)
    prm = np.stack([alpha, np.full(n, 0.4), np.full(n, 1.0)], axis=1)
    y0 = np.stack([np.ones(n), -alpha], axis=1)
    s = (Solver.new().initial(y0).equation(systems.DdeRelayClamp, prm).interval(0.0, t1).stepsize(0.012)
         .with_const_delay(1.0, 2).on(Locator.zero("abs_delayed_x_minus_1"), None).arith(arith))
    if events:
        s = s.log_events(EventLog(events))
    if trace is not None:
        s = s.on_step(trace)
    return s


def fibonacci(arith="strict"):  # examples/ode-linear-3-fibonacci.rs:11-25
    a = [-0.21520447048200203, 0.43040894096400406, 3.141592653589793,
         0.43040894096400406, 0.21520447048200203, -1.941611038725466,
         -2.2732777998989695, 1.4049629462081452, -0.4812118250596034]
    return (Solver.new().initial([0.0, 1.0, 0.0]).equation(systems.Linear3, a).interval(0.0, 50.0)
            .stepsize(0.1).on(Periodic.new(1.0), None).log_events(EventLog(64))
            .on_step(Trace(1, 1024)).arith(arith))
