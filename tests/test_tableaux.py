"""Butcher tableaux: product and oracle, bit for bit against the golden file
generated from /root/reference/src/rk.rs, plus the consistency checks the
reference keeps (commented out) at rk.rs:432-640."""
import json
import os

import numpy as np
import pytest

import diffurch_b200 as d
from diffurch_b200 import _abi as abi

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "tableaux.json")))["tableaux"]


def unpack(t):
    S, I = t.S, t.I
    a = np.array(t.a[:]).reshape(abi.MAX_STAGES, abi.MAX_STAGES)[:S, :S]
    bi = np.array(t.bi[:]).reshape(abi.MAX_STAGES, abi.MAX_INTERP)[:S, :I]
    return a, np.array(t.b[:S]), np.array(t.b2[:S]), np.array(t.c[:S]), bi


def check_against_gold(name, t):
    g = GOLD[name]
    assert (t.S, t.I) == (g["S"], g["I"])
    assert (t.order, t.order_embedded, t.order_interpolant) == (g["order"], g["order_embedded"], g["order_interpolant"])
    a, b, b2, c, bi = unpack(t)
    fh = np.vectorize(float.fromhex)
    assert np.array_equal(a, fh(np.array(g["a"])))
    assert np.array_equal(b, fh(np.array(g["b"])))
    assert np.array_equal(b2, fh(np.array(g["b2"])))
    assert np.array_equal(c, fh(np.array(g["c"])))
    assert np.array_equal(bi, fh(np.array(g["bi"])))


@pytest.mark.parametrize("name", sorted(GOLD))
def test_product_tableau_bits(name):
    check_against_gold(name, d.RK.by_name(name))


@pytest.mark.parametrize("name", sorted(GOLD))
def test_oracle_tableau_bits(oracle, name):
    check_against_gold(name, oracle.tableau(name))


@pytest.mark.parametrize("name", sorted(GOLD))
def test_tableau_consistency(name):  # rk.rs:432-640 (commented-out upstream)
    a, b, b2, c, bi = unpack(d.RK.by_name(name))
    assert np.allclose(a.sum(axis=1), c, atol=3e-16)          # sum_j a_ij = c_i
    order3_family = name in ("kutta3", "heun3", "ralston3", "wray3", "ssp3")
    if order3_family:
        # Reference quirk, replicated bit for bit: rk.rs:128 computes
        # b_1 = 1 - (c2 + 3 c3 - 2) / (6 c2 c3), which is not 1 - b_2 - b_3, so the
        # weights of the generic_order_3 family do not sum to one upstream either.
        assert abs(b.sum() - 1.0) > 1e-3
    else:
        assert abs(b.sum() - 1.0) < 3e-16
    if name != "euler":
        assert abs(b2.sum() - 1.0) < 2e-15
    assert np.allclose(bi.sum(axis=1), b, atol=1e-15)         # b_i(1) = b_i
    assert np.all(np.triu(a) == 0.0)                          # explicit
    assert np.all(bi[:, 0] == 0.0)                            # b_i(0) = 0


def test_generic_families_match_named():
    for name, args in [("midpoint", (0.5,)), ("heun2", (1.0,)), ("ralston2", (2 / 3,))]:
        check_against_gold(name, d.RK.generic_order_2(*args))
    for name, args in [("kutta3", (0.5, 1.0)), ("heun3", (1 / 3, 2 / 3)), ("wray3", (8 / 15, 2 / 3))]:
        check_against_gold(name, d.RK.generic_order_3(*args))
    # sic: the reference reports order 2 for its third-order family (rk.rs:138)
    assert d.RK.kutta3().order == 2


def test_dense_output_endpoints(oracle):
    import ctypes as C
    t = oracle.tableau("rktp64")
    k = np.array([0.3, -0.2, 0.5, 0.7, -0.1, 0.4, 0.9])
    kp = k.ctypes.data_as(C.POINTER(C.c_double))
    y0, h = 1.25, 0.02
    assert oracle.lib().orc_dense_output(C.byref(t), 0, y0, h, 0.0, kp) == y0
    y1 = oracle.lib().orc_dense_output(C.byref(t), 0, y0, h, 1.0, kp)
    a, b, b2, c, bi = unpack(t)
    assert abs(y1 - (y0 + h * float(b @ k))) < 1e-15
    # derivative at theta=0 is k_0 (bi[0][1] = 1 is the only linear coefficient)
    assert oracle.lib().orc_dense_output(C.byref(t), 1, y0, h, 0.0, kp) == k[0]
    # powi multiply sequence
    th = 0.37
    assert oracle.lib().orc_powi(th, 3) == th * (th * th)
    assert oracle.lib().orc_powi(th, 4) == (th * th) * (th * th)
    assert oracle.lib().orc_powi(th, 0) == 1.0
