"""Lyapunov exponents of an ensemble (BASELINE configs[4]: a Lyapunov batch over a parameter grid).

Two estimators, both host-side post-processing of what the device path returns:

* ``qr_spectrum`` — the reference's own scheme (tests/lyapunov_exponents.rs:243-311): the variational
  block is re-orthonormalised by a ``Periodic`` callback ("qr") that accumulates ``ln |R_ii|`` in the
  trajectory's aux slots; the spectrum is ``aux / tmax``.

* ``leading_exponent`` — for delay equations.  The reference has no DDE/NDDE Lyapunov test
  ("parity unpinned upstream"): a mutating callback sees ``t`` and ``p`` only
  (state_fn.rs:236-245), so the history segment cannot be renormalised.  For a *linear* (N)DDE the
  equation is its own variational equation and no renormalisation is needed on horizons where
  ``exp(lambda t)`` stays inside the FP64 range: the exponent is the growth rate of the norm of the
  state SEGMENT ``x_t = x(t + s), s in [-tau, 0]``.  The solver samples the trajectory with the
  ``Trace`` output policy (the on_step closure of examples/ode-chaos-lorenz.rs:25-32) — on the DDE
  hot kernel — and the estimator fits ``ln ||x_t||`` against ``t`` by least squares, which averages
  the oscillation of a complex dominant pair.
"""
import numpy as np

from . import systems
from .initial_condition import SIN, SIN_DERIV, InitFn
from .solver import Solver, Trace


def qr_spectrum(state, tmax):
    """Spectrum from the accumulated ``ln |R_ii|`` of the "qr" callback (lyapunov_exponents.rs:296-309)."""
    return np.asarray(state.aux, dtype=np.float64) / float(tmax)


def segment_log_norm(trace_t, trace_p, trace_n, window):
    """``ln`` of the RMS norm of the last ``window`` samples, per member and sample.

    trace_t: [n, cap]; trace_p: [n, cap, N]; trace_n: [n].  Returns (t [n, m], lognorm [n, m], valid
    [n, m]) with m = cap - window + 1; entry j covers samples j .. j + window - 1 and is stamped with
    the time of its last sample.
    """
    t = np.asarray(trace_t, dtype=np.float64)
    p = np.asarray(trace_p, dtype=np.float64)
    cnt = np.asarray(trace_n).astype(np.int64)
    n, cap = t.shape
    window = int(window)
    if window < 1 or window > cap:
        raise ValueError("window must be in 1..capacity")
    idx = np.arange(cap)[None, :]
    sq = np.where(idx < cnt[:, None], np.sum(p * p, axis=2), 0.0)
    m = cap - window + 1
    energy = np.zeros((n, m))
    for j in range(window):                 # shifted adds, not a prefix sum: a decaying solution spans
        energy += sq[:, j:j + m]            # hundreds of decades and a difference of prefix sums cancels
    energy /= window
    last = idx[:, window - 1:]
    valid = last < cnt[:, None]
    with np.errstate(divide="ignore", invalid="ignore"):
        lognorm = 0.5 * np.log(energy)
    valid &= np.isfinite(lognorm)
    return t[:, window - 1:], lognorm, valid


def leading_exponent(trace_t, trace_p, trace_n, window, t_skip=0.0):
    """Least-squares slope of ``ln ||x_t||`` over the samples with ``t >= t_skip``; NaN where fewer
    than two usable samples exist (a member that failed early)."""
    t, y, ok = segment_log_norm(trace_t, trace_p, trace_n, window)
    ok = ok & (t >= t_skip)
    w = ok.astype(np.float64)
    m = w.sum(axis=1)
    y = np.where(ok, y, 0.0)
    with np.errstate(divide="ignore", invalid="ignore"):
        tm = (w * t).sum(axis=1) / m
        ym = (w * y).sum(axis=1) / m
        dt = (t - tm[:, None]) * w
        slope = (dt * (y - ym[:, None])).sum(axis=1) / (dt * dt).sum(axis=1)
    slope[m < 2] = np.nan
    return slope


def linear_delay_grid(system, a, b, tau=1.0, t1=100.0, h=0.02, samples_per_delay=10, k=1.0, arith="fast"):
    """Solver for a grid of scalar linear delay equations (``systems.DdeLinear``:
    x' = a x + b x(t - tau); ``systems.NddeLinear``: x' = a x(t - tau) + b x'(t - tau);
    tests/equation_types.rs:131, 150), history ``sin(k t)``, sampled ``samples_per_delay`` times per
    delay.  Returns (solver, window) — ``window`` is what ``leading_exponent`` needs."""
    a = np.ravel(np.asarray(a, dtype=np.float64))
    b = np.ravel(np.asarray(b, dtype=np.float64))
    if a.shape != b.shape:
        raise ValueError("a and b must have the same number of members")
    n = a.shape[0]
    steps_per_delay = int(round(tau / h))
    every = max(1, steps_per_delay // int(samples_per_delay))
    n_steps = int(np.ceil((t1 - 0.0) / h)) + 2
    prm = np.stack([a, b, np.full(n, float(tau)), np.full(n, float(k))], axis=1)
    init = InitFn(SIN, SIN_DERIV) if system is systems.NddeLinear else InitFn(SIN)
    solver = (Solver.new().max_delay(float(tau)).initial(init).interval(0.0, float(t1)).stepsize(float(h))
              .equation(system, prm).arith(arith).on_step(Trace(every, n_steps // every + 2)))
    return solver, max(1, steps_per_delay // every)


def rightmost_root(system, a, b, tau=1.0, n_branches=48):
    """Real part of the rightmost characteristic root (host check for the estimator, not a reference
    quantity).  DDE: lambda = a + b exp(-lambda tau); NDDE: lambda = (a + b lambda) exp(-lambda tau).
    Newton from a fan of starting points (four per 2 pi of imaginary part); scalar inputs."""
    neutral = system is systems.NddeLinear
    best = -np.inf
    for m in range(n_branches):
        for re0 in (1.0, 0.0, -1.0, -3.0):
            z = complex(re0, 0.5 * np.pi * m / tau)
            for _ in range(200):
                if not (abs(z) < 1e3 and z.real * tau > -500.0):  # diverged
                    break
                e = np.exp(-z * tau)
                if neutral:
                    f = z - (a + b * z) * e
                    df = 1.0 - b * e + tau * (a + b * z) * e
                else:
                    f = z - a - b * e
                    df = 1.0 + tau * b * e
                if df == 0:
                    break
                dz = f / df
                z -= dz
                if abs(dz) < 1e-14 * max(1.0, abs(z)):
                    break
            if not (abs(z) < 1e3 and z.real * tau > -500.0):
                continue
            e = np.exp(-z * tau)
            res = abs(z - ((a + b * z) * e if neutral else a + b * e))
            if np.isfinite(z.real) and res < 1e-10 * max(1.0, abs(z)):
                best = max(best, z.real)
    return best


def sharded_leading_exponents(system, a, b, rank, world, runner=None, group=None, device=None, t_skip=20.0, **grid_kw):
    """The grid of ``linear_delay_grid`` sharded by contiguous member range over ``world`` ranks (one
    process per GPU); each rank integrates and post-processes its shard, then the exponents are
    gathered in member order (NCCL with ``device="cuda"``, gloo in the CPU tests).  ``runner(solver)``
    defaults to ``solver.run()`` (the CUDA path)."""
    from . import distributed as dd

    a = np.ravel(np.asarray(a, dtype=np.float64))
    b = np.ravel(np.asarray(b, dtype=np.float64))
    lo, hi = dd.shard_bounds(a.shape[0], rank, world)
    solver, window = linear_delay_grid(system, a[lo:hi], b[lo:hi], **grid_kw)
    if device is not None and runner is None and str(device).startswith("cuda"):
        import torch
        solver._device = torch.cuda.current_device()
    st = solver.run() if runner is None else runner(solver)
    lam = leading_exponent(st.trace_t, st.trace_p, st.trace_n, window, t_skip=t_skip)
    return dd.all_gather_rows(lam, group=group, device=device)
