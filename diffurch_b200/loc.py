"""Event specifications: mirror of the reference's ``loc.rs``.

``Locator.zero(f)``, ``Locator.below_zero(f)`` ... (loc.rs:913-924) pair a
detection method with a location method; ``Periodic`` (loc.rs:292-335);
``Locator.propagated_discontinuity`` (loc.rs:942-967); the ``Filter`` methods
``every / separated_by / in_interval`` (loc.rs:613-664).  ``f`` is the name (or
id) of a detector registered for the system.
"""
import copy

from . import _abi as abi


class _Loc:
    def __init__(self, kind, detector=0, detection=0, location=0, period=0.0, offset=0.0,
                 delay=0.0, smoothing_order=0):
        self.kind = kind
        self.detector = detector
        self.detection = detection
        self.location = location
        self.period = period
        self.offset = offset
        self.delay = delay
        self.smoothing_order = smoothing_order
        self.filter = abi.FILTER_NONE
        self.filter_n = 0
        self.filter_a = 0.0
        self.filter_b = 0.0
        self.filter_times = ()

    # ---- Filter trait (loc.rs:603-689) ----
    def _with_filter(self, kind, n=0, a=0.0, b=0.0):
        if self.filter != abi.FILTER_NONE:
            raise NotImplementedError("nested filters are not supported by the device path")
        r = copy.copy(self)
        r.filter, r.filter_n, r.filter_a, r.filter_b = kind, int(n), float(a), float(b)
        return r

    def every(self, n):
        return self._with_filter(abi.FILTER_EVERY, n=n)

    def separated_by(self, diff):
        return self._with_filter(abi.FILTER_SEPARATED_BY, a=diff)

    def in_interval(self, start, end):
        return self._with_filter(abi.FILTER_IN_INTERVAL, a=start, b=end)

    def on_times(self, indices):
        """``Filter::on_times(iter)`` (loc.rs:665-688): pass the detections whose 0-based index is
        in `indices`.  The reference takes any iterator and walks it monotonically; the device path
        takes a finite, strictly increasing list of at most ``MAX_FILTER_TIMES`` indices."""
        idx = [int(i) for i in indices]
        if len(idx) > abi.MAX_FILTER_TIMES or any(b <= a for a, b in zip(idx, idx[1:])) or any(i < 0 for i in idx):
            raise ValueError(f"on_times takes <= {abi.MAX_FILTER_TIMES} strictly increasing indices")
        r = self._with_filter(abi.FILTER_ON_TIMES, n=len(idx))
        r.filter_times = tuple(idx)
        return r

    # location-method overrides (the reference builds these by hand from LocatorStateFn)
    def with_location(self, location):
        r = copy.copy(self)
        r.location = int(location)
        return r

    def fill(self, L, system, callback, dedup):
        det = self.detector
        if isinstance(det, str):
            det = system.detectors[det]
        cb = callback
        if isinstance(cb, str):
            cb = system.callbacks[cb]
        L.kind = self.kind
        L.detector_id = int(det or 0)
        L.detection = self.detection
        L.location = self.location
        L.callback_id = int(cb or 0)
        L.dedup = 1 if dedup else 0
        L.smoothing_order = int(self.smoothing_order)
        L.filter = self.filter
        L.filter_n = self.filter_n
        L.period = float(self.period)
        L.offset = float(self.offset)
        L.delay = float(self.delay)
        L.filter_a = self.filter_a
        L.filter_b = self.filter_b
        for i in range(abi.MAX_FILTER_TIMES):
            L.filter_times[i] = self.filter_times[i] if i < len(self.filter_times) else 0


class Locator:
    """``Locator::<T, P>`` constructors (loc.rs:913-967)."""

    @staticmethod
    def _mk(f, detection, location):
        return _Loc(abi.LOC_STATEFN, detector=f, detection=detection, location=location)

    @staticmethod
    def zero(f):
        return Locator._mk(f, abi.DET_ZERO, abi.LOCATE_BISECTION)

    @staticmethod
    def above_zero(f):
        return Locator._mk(f, abi.DET_ABOVE_ZERO, abi.LOCATE_BISECTION)

    @staticmethod
    def below_zero(f):
        return Locator._mk(f, abi.DET_BELOW_ZERO, abi.LOCATE_BISECTION)

    @staticmethod
    def positive(f):
        return Locator._mk(f, abi.DET_POSITIVE, abi.LOCATE_STEP_BEGIN)

    @staticmethod
    def negative(f):
        return Locator._mk(f, abi.DET_NEGATIVE, abi.LOCATE_STEP_BEGIN)

    @staticmethod
    def switch(f):
        return Locator._mk(f, abi.DET_SWITCH, abi.LOCATE_BISECTION_BOOL)

    @staticmethod
    def switch_true(f):
        return Locator._mk(f, abi.DET_SWITCH_TRUE, abi.LOCATE_BISECTION_BOOL)

    @staticmethod
    def switch_false(f):
        return Locator._mk(f, abi.DET_SWITCH_FALSE, abi.LOCATE_BISECTION_BOOL)

    @staticmethod
    def is_true(f):
        return Locator._mk(f, abi.DET_IS_TRUE, abi.LOCATE_BISECTION_BOOL)

    @staticmethod
    def is_false(f):
        return Locator._mk(f, abi.DET_IS_FALSE, abi.LOCATE_BISECTION_BOOL)

    @staticmethod
    def step():
        return Locator._mk(0, abi.DET_STEP, abi.LOCATE_STEP_END)

    @staticmethod
    def propagated_discontinuity(delayed, smoothing_order):
        return _Loc(abi.LOC_PROPAGATION, detector=delayed, smoothing_order=smoothing_order)


class Periodic(_Loc):
    """``Periodic { period, offset }`` (loc.rs:292-308)."""

    def __init__(self, period, offset=0.0):
        super().__init__(abi.LOC_PERIODIC, period=period, offset=offset)

    @staticmethod
    def new(period):
        return Periodic(period)

    def with_offset(self, offset):
        return Periodic(self.period, offset)
