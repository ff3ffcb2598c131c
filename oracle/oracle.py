"""TEST INFRASTRUCTURE ONLY — ctypes wrapper of oracle/_build/liboracle.so.

Imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs, never by the product.  Takes the same C-ABI ``Problem``
descriptor as the product so both sides see identical inputs.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from diffurch_b200 import _abi as abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")


def build(verbose=False):
    """Compile the oracle (make is incremental)."""
    r = subprocess.run(["make", "-C", _HERE], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stdout)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        dp = C.POINTER(C.c_double)
        u64p = C.POINTER(C.c_uint64)
        u32p = C.POINTER(C.c_uint32)
        L.orc_run.restype = C.c_int
        L.orc_run.argtypes = [C.POINTER(abi.Problem), C.c_size_t, dp, dp, C.c_int, dp, dp, dp,
                              u64p, u64p, u64p, u64p, u32p, u32p, dp, dp, u32p, dp,
                              C.POINTER(C.c_int32)]
        L.orc_tableau_by_name.restype = C.c_int
        L.orc_tableau_by_name.argtypes = [C.c_char_p, C.POINTER(abi.Tableau)]
        L.orc_tableau_generic_order_2.restype = C.c_int
        L.orc_tableau_generic_order_2.argtypes = [C.c_double, C.POINTER(abi.Tableau)]
        L.orc_tableau_generic_order_3.restype = C.c_int
        L.orc_tableau_generic_order_3.argtypes = [C.c_double, C.c_double, C.c_int, C.c_double,
                                                  C.POINTER(abi.Tableau)]
        L.orc_dense_output.restype = C.c_double
        L.orc_dense_output.argtypes = [C.POINTER(abi.Tableau), C.c_int, C.c_double, C.c_double,
                                       C.c_double, dp]
        L.orc_partition_point_linear.restype = C.c_size_t
        L.orc_partition_point_linear.argtypes = [C.POINTER(C.c_int64), C.c_size_t, C.c_size_t,
                                                 C.c_int64, u64p]
        L.orc_householder_qr3.restype = None
        L.orc_householder_qr3.argtypes = [dp, dp, dp]
        L.orc_powi.restype = C.c_double
        L.orc_powi.argtypes = [C.c_double, C.c_int]
        L.orc_set_portable_root.restype = None
        L.orc_set_portable_root.argtypes = [C.c_int]
        L.orc_portable_root.restype = C.c_double
        L.orc_portable_root.argtypes = [C.c_double, C.c_uint]
        L.orc_system_info.restype = C.c_int
        L.orc_system_info.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                      C.POINTER(C.c_int)]
        _lib = L
    return _lib


def tableau(name):
    t = abi.Tableau()
    rc = lib().orc_tableau_by_name(name.encode(), C.byref(t))
    if rc != 0:
        raise ValueError(name)
    return t


def system_info(system_id):
    a, b, c = C.c_int(), C.c_int(), C.c_int()
    if lib().orc_system_info(system_id, C.byref(a), C.byref(b), C.byref(c)) != 0:
        raise ValueError(system_id)
    return a.value, b.value, c.value


class Result:
    pass


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def run(problem, y0, params=None, n_threads=1):
    """Run the CPU oracle on every trajectory.  y0 [n, N], params [n, P]."""
    N, P, A = system_info(problem.system_id)
    y0 = np.ascontiguousarray(y0, dtype=np.float64).reshape(-1, N)
    n = y0.shape[0]
    if P:
        params = np.ascontiguousarray(params, dtype=np.float64).reshape(n, P)
    else:
        params = None
    r = Result()
    r.t = np.empty(n)
    r.p = np.empty((n, N))
    r.aux = np.zeros((n, max(A, 1)))
    r.steps = np.zeros(n, dtype=np.uint64)
    r.rejected = np.zeros(n, dtype=np.uint64)
    r.events = np.zeros(n, dtype=np.uint64)
    r.rhs_evals = np.zeros(n, dtype=np.uint64)
    r.status = np.zeros(n, dtype=np.uint32)
    cap = problem.trace_capacity if problem.trace_every > 0 else 0
    r.trace_n = np.zeros(n, dtype=np.uint32)
    r.trace_t = np.zeros((n, max(cap, 1)))
    r.trace_p = np.zeros((n, max(cap, 1), N))
    ecap = problem.event_capacity
    r.event_n = np.zeros(n, dtype=np.uint32)
    r.event_t = np.zeros((n, max(ecap, 1)))
    r.event_index = np.full((n, max(ecap, 1)), -1, dtype=np.int32)
    u64 = lambda a: a.ctypes.data_as(C.POINTER(C.c_uint64))
    u32 = lambda a: a.ctypes.data_as(C.POINTER(C.c_uint32))
    rc = lib().orc_run(C.byref(problem), n, _dp(y0), _dp(params), int(n_threads),
                       _dp(r.t), _dp(r.p), _dp(r.aux) if A else None,
                       u64(r.steps), u64(r.rejected), u64(r.events), u64(r.rhs_evals),
                       u32(r.status),
                       u32(r.trace_n), _dp(r.trace_t) if cap else None,
                       _dp(r.trace_p) if cap else None,
                       u32(r.event_n), _dp(r.event_t) if ecap else None,
                       r.event_index.ctypes.data_as(C.POINTER(C.c_int32)) if ecap else None)
    if rc != 0:
        raise RuntimeError(f"orc_run failed: {rc}")
    r.aux = r.aux[:, :A]
    if not cap:
        r.trace_t = r.trace_t[:, :0]
        r.trace_p = r.trace_p[:, :0]
    return r


def run_solver(solver, n_threads=1, portable_root=False):
    """Run the oracle on what a diffurch_b200.Solver builder describes.  `portable_root`: compute the
    AutomaticStepsize factor with the portable n-th root the STRICT CUDA build uses instead of libm pow
    (see diffurch_oracle.hpp)."""
    n, y0, params = solver._resolve()
    lib().orc_set_portable_root(1 if portable_root else 0)
    try:
        return run(solver.problem(), y0, params, n_threads=n_threads)
    finally:
        lib().orc_set_portable_root(0)
