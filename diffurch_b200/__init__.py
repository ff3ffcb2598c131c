"""diffurch_b200 — B200-native ensemble integrator for the hot path of
Danila-Bain/diffurch (the explicit Runge-Kutta step loop), behind a mirror of the
reference's ``Solver`` builder.  CUDA (sm_100a) only: importing this package
loads ``libdiffurch_b200.so`` and fails loudly if it has not been built.
"""
from . import _abi as abi  # noqa: F401
from . import systems  # noqa: F401
from ._lib import LIB_PATH, DfbError, lib  # noqa: F401
from .initial_condition import INV_SQUARE, SIN, SIN_DERIV, InitFn  # noqa: F401
from .loc import Locator, Periodic  # noqa: F401
from .rk import RK, ButcherTableau, ButcherTableu  # noqa: F401
from .solver import (Ensemble, EnsembleState, EventLog, Solver, Trace,  # noqa: F401
                     default_problem)
from .stepsize import AutomaticStepsize  # noqa: F401
from . import lyapunov  # noqa: F401,E402

__version__ = "0.1.0"
