"""Frozen oracle outputs (tests/golden/oracle_golden.json, made by tests/golden/make_oracle_golden.py):
the oracle must still produce them (CPU), and the STRICT CUDA path must hit them bit for bit (GPU)
— BASELINE configs[0] (the single Lorenz trajectory of examples/ode-chaos-lorenz.rs) included."""
import json
import os

import numpy as np
import pytest

import cases

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "oracle_golden.json")))


def unhex(v):
    return np.array([float.fromhex(x) for x in v])


def check(run):
    r = run(cases.lorenz(n=1, t1=50.0))
    g = GOLD["lorenz_single_t50"]
    assert int(r.steps[0]) == g["steps"] == 12501 and int(r.rhs_evals[0]) == g["rhs_evals"]
    assert np.array_equal(r.t, unhex(g["t"])) and np.array_equal(r.p.ravel(), unhex(g["p"]))
    r = run(cases.bouncing_ball())
    g = GOLD["bouncing_ball_t8.58"]
    n = int(r.event_n[0])
    assert n == len(g["event_t"]) and int(r.steps[0]) == g["steps"]
    assert np.array_equal(r.event_t[0, :n], unhex(g["event_t"]))
    assert [int(i) for i in r.event_index[0, :n]] == g["event_index"]
    assert np.array_equal(r.p.ravel(), unhex(g["p"]))
    r = run(cases.egg_billiard(reflections=40))
    g = GOLD["egg_billiard_40"]
    n = int(r.event_n[0])
    assert np.array_equal(r.event_t[0, :n], unhex(g["event_t"])) and np.array_equal(r.p.ravel(), unhex(g["p"]))
    return r


def test_oracle_reproduces_its_frozen_outputs(oracle):
    check(lambda s: oracle.run_solver(s))
    for case, key in ((cases.eq_dde_sin(trace=False), "dde_sin_t10"), (cases.eq_ndde_sin(), "ndde_sin_t10")):
        r = oracle.run_solver(case)
        assert np.array_equal(r.p.ravel(), unhex(GOLD[key]["p"])) and int(r.steps[0]) == GOLD[key]["steps"]


@pytest.mark.gpu
def test_strict_cuda_path_hits_the_frozen_outputs():
    check(lambda s: s.run())
    # the delay equations call sin() in their history function: CUDA and glibc agree to 1-2 ulp
    for case, key in ((cases.eq_dde_sin(trace=False), "dde_sin_t10"), (cases.eq_ndde_sin(), "ndde_sin_t10")):
        r = case.run()
        assert int(r.steps[0]) == GOLD[key]["steps"]
        assert np.max(np.abs(r.p.ravel() - unhex(GOLD[key]["p"]))) < 1e-13
