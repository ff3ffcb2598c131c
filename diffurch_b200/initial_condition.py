"""History functions: mirror of the reference's ``initial_condition.rs``.

A plain array passed to ``Solver.initial`` is the constant initial condition
(initial_condition.rs:15-22).  ``InitFn(f)`` / ``InitFn(f, df)`` (:24-50) take the
names of registered device functors instead of host closures.
"""
from . import _abi as abi

SIN = "sin(k t)"
SIN_DERIV = "k cos(k t)"
INV_SQUARE = "t^-2"


class InitFn:
    def __init__(self, f, df=None):
        self.f = f
        self.df = df
        if f == SIN and df == SIN_DERIV:
            self.history_id = abi.HIST_SIN
        elif f == SIN and df is None:
            self.history_id = abi.HIST_SIN_NODERIV
        elif f == INV_SQUARE and df is None:
            self.history_id = abi.HIST_INV_SQUARE
        else:
            raise ValueError(f"no registered history functor for InitFn({f!r}, {df!r})")
