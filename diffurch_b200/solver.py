"""Host-side mirror of the reference's ``Solver`` builder (src/solver.rs) over the C ABI.

Same setters, same argument meaning, plus the ensemble axis::

    Solver.new()                         # rktp64, h = 0.05, max_delay = 0   (solver.rs:80-96)
        .initial(y0)                     # [n_traj, N] constants, or InitFn(...)
        .equation(systems.Lorenz, params)  # registered device functor + captured values
        .interval(0.0, 50.0)
        .stepsize(1 / 250)               # float, or AutomaticStepsize(...)
        .rk(RK.rk4())
        .on(Locator.zero("dx"), None)    # non-mutating callback (None = `|_| {}`)
        .on_mut(Locator.below_zero("x"), "bounce")
        .with_const_delay(1.0, 0) / .with_delayed_argument("half_time", 0)
        .max_delay(1.25) / .initial_disco([(0.0, 0)])
        .on_step(Trace(every=1, capacity=512))   # output policy instead of a host closure
        .run()                           # -> EnsembleState

``run`` consumes the solver (``Solver::run(self)``, solver.rs:263): a second call
raises.  Host closures cannot run on the device, so ``equation``/``on``/``on_mut``
name registered functors and ``on_step`` takes an output policy.
"""
import ctypes as C
import math

import numpy as np

from . import _abi as abi
from . import systems as _systems
from ._lib import DfbError, check, lib
from .initial_condition import InitFn
from .loc import Locator, Periodic, _Loc
from .stepsize import AutomaticStepsize


class Trace:
    """Output policy: record (t, p) at every ``every``-th ``on_step`` invocation
    (the one at t_init is invocation 0, solver.rs:290), at most ``capacity`` samples."""

    def __init__(self, every=1, capacity=1024):
        self.every = int(every)
        self.capacity = int(capacity)


class EventLog:
    """Output policy: record (time, locator index) of every fired callback."""

    def __init__(self, capacity=256):
        self.capacity = int(capacity)


class EnsembleState:
    """What ``Solver::run`` returns (the final ``State``, solver.rs:313), per trajectory."""

    def __init__(self):
        self.t = None        # [n]
        self.p = None        # [n, N]
        self.aux = None      # [n, n_aux]  (callback accumulators)
        self.steps = None
        self.rejected = None
        self.events = None
        self.rhs_evals = None
        self.status = None
        self.trace_n = None  # [n]
        self.trace_t = None  # [n, capacity]
        self.trace_p = None  # [n, capacity, N]
        self.event_n = None
        self.event_t = None
        self.event_index = None
        self.kernel_ms = None
        self.totals = None


def _dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class Ensemble:
    """Thin RAII wrapper of a ``dfb_handle`` (one device, one stream)."""

    def __init__(self, problem, n_traj, device=0):
        self.problem = problem
        self.n = int(n_traj)
        self.system = _systems.BY_ID[problem.system_id]
        self._h = C.c_void_p()
        check(lib.dfb_create(C.byref(problem), self.n, int(device), C.byref(self._h)))

    def close(self):
        if self._h:
            lib.dfb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def set_initial(self, y0):
        y0 = np.ascontiguousarray(y0, dtype=np.float64).reshape(self.n, self.system.n_state)
        check(lib.dfb_set_initial(self._h, _dptr(y0)))

    def set_params(self, params):
        if self.system.n_params == 0:
            return
        params = np.ascontiguousarray(params, dtype=np.float64).reshape(self.n, self.system.n_params)
        check(lib.dfb_set_params(self._h, _dptr(params)))

    def set_initial_device(self, ptr):
        check(lib.dfb_set_initial_device(self._h, C.c_void_p(int(ptr))))

    def set_params_device(self, ptr):
        check(lib.dfb_set_params_device(self._h, C.c_void_p(int(ptr))))

    def run(self, stream=0):
        check(lib.dfb_run(self._h, C.c_void_p(int(stream))))

    def synchronize(self):
        check(lib.dfb_synchronize(self._h))

    def final(self, want_aux=True, out=None):
        """Final State of every member: (t [n], p [n, N], aux [n, A]).  `out=(t, p)` or `(t, p, aux)`
        reuses caller-owned arrays (e.g. pinned host memory, so the download is a true async DMA)."""
        n, N, A = self.n, self.system.n_state, self.system.n_aux
        if out is not None:
            t, y = out[0], out[1]
            assert t.shape == (n,) and y.shape == (n, N) and t.dtype == y.dtype == np.float64
            aux = out[2] if (len(out) > 2 and A and want_aux) else (np.zeros((n, A)) if (A and want_aux) else None)
        else:
            t = np.empty(n)
            y = np.empty((n, N))
            aux = np.zeros((n, A)) if (A and want_aux) else None
        check(lib.dfb_get_final(self._h, _dptr(t), _dptr(y), _dptr(aux) if aux is not None else None))
        return t, y, (aux if aux is not None else np.zeros((n, 0)))

    def final_device_ptrs(self):
        a, b, c = C.c_void_p(), C.c_void_p(), C.c_void_p()
        check(lib.dfb_final_device(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def stats(self):
        n = self.n
        out = {k: np.empty(n, dtype=np.uint64) for k in ("steps", "rejected", "events", "rhs_evals")}
        out["status"] = np.empty(n, dtype=np.uint32)
        s = abi.Stats()
        for k in ("steps", "rejected", "events", "rhs_evals"):
            setattr(s, k, out[k].ctypes.data_as(C.POINTER(C.c_uint64)))
        s.status = out["status"].ctypes.data_as(C.POINTER(C.c_uint32))
        check(lib.dfb_get_stats(self._h, C.byref(s)))
        return out

    def totals(self):
        t = abi.Totals()
        check(lib.dfb_get_totals(self._h, C.byref(t)))
        return {k: getattr(t, k) for k, _ in abi.Totals._fields_ if not k.startswith("_")}

    def trace(self):
        n, N, cap = self.n, self.system.n_state, self.problem.trace_capacity
        ns = np.empty(n, dtype=np.uint32)
        t = np.empty((n, cap))
        y = np.empty((n, cap, N))
        check(lib.dfb_get_trace(self._h, ns.ctypes.data_as(C.POINTER(C.c_uint32)), _dptr(t), _dptr(y)))
        return ns, t, y

    def events(self):
        n, cap = self.n, self.problem.event_capacity
        ne = np.empty(n, dtype=np.uint32)
        t = np.empty((n, cap))
        idx = np.empty((n, cap), dtype=np.int32)
        check(lib.dfb_get_events(self._h, ne.ctypes.data_as(C.POINTER(C.c_uint32)), _dptr(t),
                                 idx.ctypes.data_as(C.POINTER(C.c_int32))))
        return ne, t, idx

    def collect(self):
        st = EnsembleState()
        st.t, st.p, st.aux = self.final()
        s = self.stats()
        st.steps, st.rejected, st.events = s["steps"], s["rejected"], s["events"]
        st.rhs_evals, st.status = s["rhs_evals"], s["status"]
        if self.problem.trace_every > 0 and self.problem.trace_capacity > 0:
            st.trace_n, st.trace_t, st.trace_p = self.trace()
        if self.problem.event_capacity > 0:
            st.event_n, st.event_t, st.event_index = self.events()
        st.totals = self.totals()
        st.kernel_ms = st.totals["kernel_ms"]
        return st


def default_problem():
    p = abi.Problem()
    check(lib.dfb_problem_default(C.byref(p)))
    return p


class Solver:
    """Builder; field names follow ``pub struct Solver`` (solver.rs:51-77)."""

    def __init__(self):
        self._problem = default_problem()
        self._system = None
        self._params = None
        self._y0 = None
        self._n_traj = None
        self._consumed = False
        self._device = 0

    # Solver::new::<T, P>() (solver.rs:80-96)
    @staticmethod
    def new(n_traj=None, device=0):
        s = Solver()
        s._n_traj = n_traj
        s._device = device
        return s

    # ---- setters, in the reference's order (solver.rs:113-261) ----------------
    def initial_disco(self, initial_disco):
        lst = list(initial_disco)
        if len(lst) > abi.MAX_DISCO:
            raise ValueError("too many initial discontinuities")
        self._problem.n_disco = len(lst)
        for i, (t, order) in enumerate(lst):
            self._problem.disco_t[i] = float(t)
            self._problem.disco_order[i] = int(order)
        return self

    def stepsize(self, new_stepsize):
        p = self._problem
        if isinstance(new_stepsize, AutomaticStepsize):
            a = new_stepsize
            p.stepsize_kind = abi.STEP_AUTOMATIC
            p.automatic.stepsize = float(a.stepsize)
            p.automatic.stepsize_min, p.automatic.stepsize_max = map(float, a.stepsize_range)
            for i, v in enumerate(a.atol):
                p.automatic.atol[i] = float(v)
            for i, v in enumerate(a.rtol):
                p.automatic.rtol[i] = float(v)
            p.automatic.order = int(a.order)
            p.automatic.fac = float(a.fac)
            p.automatic.fac_min, p.automatic.fac_max = map(float, a.fac_range)
            p.automatic.initial_stepsize = float(a.initial_stepsize or 0.0)
        else:
            p.stepsize_kind = abi.STEP_FIXED
            p.h = float(new_stepsize)
        return self

    def max_delay(self, max_delay):
        self._problem.max_delay = float(max_delay)
        return self

    def equation(self, system, params=None):
        self._system = system
        self._problem.system_id = system.id
        self._params = params
        return self

    def initial(self, new_initial):
        if isinstance(new_initial, InitFn):
            self._problem.history_id = new_initial.history_id
            self._y0 = None
        else:
            self._problem.history_id = abi.HIST_CONSTANT
            self._y0 = np.asarray(new_initial, dtype=np.float64)
        return self

    def interval(self, start, end=math.inf):
        # RangeBounds (interval.rs:8-25): unbounded start = 0, unbounded end = +inf
        self._problem.t0 = 0.0 if start is None else float(start)
        self._problem.t1 = math.inf if end is None else float(end)
        return self

    def rk(self, new_rk):
        self._problem.rk = new_rk
        return self

    def on_step(self, policy):
        if not isinstance(policy, Trace):
            raise TypeError("on_step takes an output policy (Trace); host closures cannot run on the device")
        self._problem.trace_every = policy.every
        self._problem.trace_capacity = policy.capacity
        return self

    def on_start_mut(self, callback):
        """``Solver::on_start_mut`` (solver.rs:190-197): a registered mutating callback run once on
        the initial State, before the first ``on_step``."""
        if self._system is None:
            raise ValueError("call .equation(system, params) before registering callbacks")
        cb = self._system.callbacks[callback] if isinstance(callback, str) else int(callback or 0)
        self._problem.on_start_callback_id = cb
        return self

    def on_start(self, callback=None):
        """``Solver::on_start`` (solver.rs:181-188) takes a read-only closure: it observes the
        initial state, which the caller of a batched solve already holds.  Accepted for source
        compatibility; nothing is sent to the device."""
        return self

    def on_stop(self, callback=None):
        """``Solver::on_stop`` (solver.rs:172-179), read-only: it observes the final State, which
        ``run()`` returns.  Accepted for source compatibility."""
        return self

    def log_events(self, policy):
        self._problem.event_capacity = policy.capacity
        return self

    def _append_loc(self, loc, callback, dedup):
        if self._system is None:
            raise ValueError("call .equation(system, params) before registering events")
        p = self._problem
        if p.n_loc >= abi.MAX_LOCATORS:
            raise ValueError("too many locators")
        loc.fill(p.loc[p.n_loc], self._system, callback, dedup)
        p.n_loc += 1
        return self

    def on(self, loc, callback=None):
        """``Solver::on`` (solver.rs:199-216): locator + non-mutating callback, DedupLocF-wrapped."""
        return self._append_loc(loc, callback, dedup=True)

    def on_mut(self, loc, callback):
        """``Solver::on_mut`` (solver.rs:219-236)."""
        return self._append_loc(loc, callback, dedup=True)

    def with_delayed_argument(self, delayed, smoothing_order):
        """``Solver::with_delayed_argument`` (solver.rs:239-248): not deduplicated."""
        return self._append_loc(Locator.propagated_discontinuity(delayed, smoothing_order), None, False)

    def with_const_delay(self, delay, smoothing_order):
        """``Solver::with_const_delay`` (solver.rs:251-261): raises max_delay, then
        registers the propagated-discontinuity locator for ``t - delay``."""
        self._problem.max_delay = max(self._problem.max_delay, float(delay))
        loc = _Loc(abi.LOC_PROPAGATION, detector=0, delay=delay, smoothing_order=smoothing_order)
        return self._append_loc(loc, None, dedup=False)

    # ---- options that have no reference counterpart ----------------------------
    def arith(self, mode):
        self._problem.arith = {"strict": abi.ARITH_STRICT, "fast": abi.ARITH_FAST}[mode]
        return self

    def max_steps(self, n):
        self._problem.max_steps = int(n)
        return self

    def ring_capacity(self, n):
        self._problem.ring_capacity = int(n)
        return self

    # ---- run -------------------------------------------------------------------
    def _resolve(self):
        if self._system is None:
            raise ValueError("no equation set")
        sysm = self._system
        n = self._n_traj
        y0 = self._y0
        params = self._params
        if y0 is not None:
            y0 = np.atleast_1d(y0)
            if y0.ndim == 1:
                y0 = y0.reshape(1, -1) if y0.size == sysm.n_state else y0.reshape(-1, sysm.n_state)
            if y0.shape[1] != sysm.n_state:
                raise ValueError(f"initial must have {sysm.n_state} components")
        if params is not None:
            params = np.atleast_1d(np.asarray(params, dtype=np.float64))
            if params.ndim == 1:
                params = params.reshape(1, -1) if params.size == sysm.n_params else params.reshape(-1, sysm.n_params)
        cands = [a.shape[0] for a in (y0, params) if a is not None and a.shape[0] > 1]
        if n is None:
            n = max(cands) if cands else 1
        if y0 is None:
            y0 = np.zeros((n, sysm.n_state))
        if y0.shape[0] == 1 and n > 1:
            y0 = np.repeat(y0, n, axis=0)
        if sysm.n_params:
            if params is None:
                raise ValueError(f"system {sysm.name} needs params {sysm.params}")
            if params.shape[0] == 1 and n > 1:
                params = np.repeat(params, n, axis=0)
            if params.shape != (n, sysm.n_params):
                raise ValueError("params shape mismatch")
        if y0.shape[0] != n:
            raise ValueError("initial shape mismatch")
        return n, np.ascontiguousarray(y0), (np.ascontiguousarray(params) if params is not None else None)

    def problem(self):
        """The flattened C-ABI descriptor (a copy)."""
        p = abi.Problem()
        C.memmove(C.byref(p), C.byref(self._problem), C.sizeof(abi.Problem))
        return p

    def build(self):
        """Create the device handle with inputs uploaded, without running."""
        n, y0, params = self._resolve()
        ens = Ensemble(self.problem(), n, self._device)
        ens.set_initial(y0)
        if params is not None:
            ens.set_params(params)
        return ens

    def run(self):
        if self._consumed:
            raise RuntimeError("Solver::run(self) consumes the solver (solver.rs:263)")
        self._consumed = True
        with self.build() as ens:
            ens.run()
            return ens.collect()


__all__ = ["Solver", "Ensemble", "EnsembleState", "Trace", "EventLog", "Locator", "Periodic",
           "InitFn", "AutomaticStepsize", "DfbError"]
