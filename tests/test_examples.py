"""The reference's live examples (examples/*.rs), re-entered in examples/*.py, run end to end."""
import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "examples"))

pytestmark = pytest.mark.gpu


def run(name, **kw):
    return importlib.import_module(name).main(**kw)


def test_examples_run_and_give_the_documented_results():
    st = run("ode_chaos_lorenz", n=64, t1=52.0)
    assert np.all(st.status == 0) and st.trace_n[0] > 500
    st = run("bouncing_ball", n=64)
    assert np.all(st.status == 0) and st.events[0] >= 40
    st = run("egg_billiard", n=64, reflections=50)
    assert np.all(st.events == 50)
    st = run("dde_relay_clamp", n=64, t1=10.0)
    assert np.all(st.status == 0) and st.events.sum() > 0
    st = run("ode_2_relay_msign", n=32)
    assert np.all(st.status == 0) and np.all(st.events >= 5)
    st = run("ode_linear_1_exp")
    assert abs(st.p[0, 0] - 2.0 ** 16) / 2.0 ** 16 < 2e-5            # rktp64 with h = 1, k = ln 2: ~7e-7 per step
    st = run("ode_linear_3_fibonacci")
    t, x = st.trace_t[0, :st.trace_n[0]], st.trace_p[0, :st.trace_n[0], 0]
    assert abs(x[np.argmax(t == 10.0)] - 55.0) < 1e-6 and abs(x[np.argmax(t == 20.0)] - 6765.0) < 1e-3
    lam, a, b, worst = run("ndde_lyapunov_grid", na=16, nb=8, t1=100.0, check=8)
    assert lam.shape == (128,) and np.all(np.isfinite(lam)) and worst < 2e-3
    st = run("user_system_van_der_pol", n=64)
    assert np.all(st.status == 0) and np.all(st.aux[:, 0] >= 8)
