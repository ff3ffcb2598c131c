"""Step-size controllers: mirror of the reference's ``stepsize.rs``.

A bare float is the fixed-step controller (stepsize.rs:18-32);
:class:`AutomaticStepsize` mirrors the struct of stepsize.rs:35-44 field for field.
"""
from dataclasses import dataclass
from typing import Optional, Sequence, Tuple


@dataclass
class AutomaticStepsize:
    stepsize: float
    stepsize_range: Tuple[float, float]
    atol: Sequence[float]
    rtol: Sequence[float]
    order: int
    fac: float
    fac_range: Tuple[float, float]
    initial_stepsize: Optional[float] = None  # dead in the reference: init() is never called
