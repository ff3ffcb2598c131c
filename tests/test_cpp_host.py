"""The header-only C++ mirror of the Solver builder (csrc/host/solver.hpp): compiles
against the C ABI (CPU check), and on a GPU gives the same bits as the ctypes path."""
import os
import subprocess

import numpy as np
import pytest

import cases
import diffurch_b200 as d

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "host_solver_demo.cpp")
EXE = os.path.join(ROOT, "tests", "cpp", "_build", "host_solver_demo")


def build_demo():
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    libdir = os.path.dirname(d.LIB_PATH)
    if (not os.path.exists(EXE)) or os.path.getmtime(EXE) < max(os.path.getmtime(SRC), os.path.getmtime(d.LIB_PATH)):
        subprocess.run(["g++", "-std=c++17", "-O1", SRC, "-o", EXE, f"-L{libdir}", "-ldiffurch_b200",
                        f"-Wl,-rpath,{libdir}"], check=True)
    return EXE


def test_cpp_mirror_compiles_and_refuses_without_gpu():
    exe = build_demo()
    from conftest import has_cuda
    r = subprocess.run([exe], capture_output=True, text=True)
    if not has_cuda():
        assert r.returncode == 3 and "no CUDA device" in r.stderr


@pytest.mark.gpu
def test_cpp_mirror_matches_ctypes_path():
    r = subprocess.run([build_demo()], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = {ln.split()[0]: ln.split()[1:] for ln in r.stdout.strip().splitlines()}
    prm = np.array([[10.0, 20.0 + 10.0 * i, 8.0 / 3.0] for i in range(3)])
    g = (d.Solver.new().initial(np.tile([1., 2., 20.], (3, 1))).equation(d.systems.Lorenz, prm)
         .interval(0.0, 1.0).stepsize(1.0 / 250.0).run())
    vals = np.array([float(x) for x in out["lorenz"]]).reshape(3, 4)
    assert np.array_equal(vals[:, 0], g.steps.astype(float)) and np.array_equal(vals[:, 1:], g.p)
    b = cases.bouncing_ball(t1=4.0, events=64).run()
    ball = [float(x) for x in out["ball"]]
    assert ball[0] == b.steps[0] and ball[1] == b.events[0] and ball[2:4] == list(b.p[0])
    assert ball[4:] == list(b.event_t[0, :len(ball) - 4])
    assert [float(x) for x in out["delay"]] == [0.0, 0.75, 1.0, 1.75, 2.0, 2.75, 3.0, 3.75, 4.0, 4.75, 5.0]
