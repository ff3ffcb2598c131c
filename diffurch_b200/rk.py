"""Butcher tableaux: mirror of the reference's ``rk.rs`` constructors.

``RK.rktp64()``, ``RK.rk4()``, ... return the C-ABI ``Tableau`` POD filled by the
library (reference ``ButcherTableu::<name>()``, src/rk.rs:55-430; the reference
spells the type ``ButcherTableu`` and aliases it ``RK``, src/lib.rs:20-21).
"""
import ctypes as C

from . import _abi as abi
from ._lib import check, lib

NAMES = ["euler", "midpoint", "heun2", "ralston2", "kutta3", "heun3", "ralston3", "wray3",
         "ssp3", "rk4", "three_eights", "rk43", "rktp64"]


def _by_name(name):
    t = abi.Tableau()
    check(lib.dfb_tableau_by_name(name.encode(), C.byref(t)))
    return t


class ButcherTableu:
    """Namespace of constructors, like ``impl ButcherTableu``."""

    @staticmethod
    def by_name(name):
        return _by_name(name)

    @staticmethod
    def generic_order_2(c2):
        t = abi.Tableau()
        check(lib.dfb_tableau_generic_order_2(float(c2), C.byref(t)))
        return t

    @staticmethod
    def generic_order_3(c2, c3, b3=None):
        t = abi.Tableau()
        check(lib.dfb_tableau_generic_order_3(float(c2), float(c3), 0 if b3 is None else 1,
                                              0.0 if b3 is None else float(b3), C.byref(t)))
        return t


for _n in NAMES:
    setattr(ButcherTableu, _n, staticmethod(lambda _n=_n: _by_name(_n)))

RK = ButcherTableu
ButcherTableau = ButcherTableu
