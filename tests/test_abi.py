"""CPU-only checks of the drop-in boundary: the C-ABI library loads, exports every
symbol include/diffurch_b200.h declares, validates descriptors, and refuses to
compute without a CUDA device (there is no CPU fallback)."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest

import diffurch_b200 as d
from diffurch_b200 import _abi as abi
from conftest import has_cuda

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "diffurch_b200.h")


def test_header_symbols_exported():
    text = open(HEADER).read()
    declared = re.findall(r"^DFB_API\s+[\w\s\*]+?\b(dfb_\w+)\s*\(", text, flags=re.M)
    assert len(declared) >= 20
    assert sorted(declared) == sorted(abi.EXPORTED_SYMBOLS)
    raw = C.CDLL(d.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), name


def test_struct_layout_matches_default_problem():
    p = d.default_problem()
    assert p.abi_version == abi.ABI_VERSION
    # Solver::new defaults (solver.rs:80-96)
    assert (p.rk.S, p.rk.I, p.rk.order) == (7, 5, 6)
    assert p.h == 0.05 and p.max_delay == 0.0 and p.n_loc == 0
    assert p.t0 == 0.0 and np.isinf(p.t1)
    # enum/define values agree with the header text
    text = open(HEADER).read()
    for name, val in re.findall(r"#define (DFB_MAX_\w+) (\d+)", text):
        assert getattr(abi, name[4:]) == int(val), name
    for name, val in re.findall(r"(DFB_SYS_\w+) = (\d+)", text):
        assert getattr(abi, name[4:]) == int(val), name


def test_system_info_matches_python_registry():
    for s in d.systems.ALL:
        info = abi.SystemInfo()
        assert d.lib.dfb_system_info_get(s.id, C.byref(info)) == 0
        assert info.n_state == s.n_state and info.n_params == s.n_params
        assert info.n_aux == s.n_aux and bool(info.uses_history) == s.uses_history
        assert info.n_detectors == len(s.detectors)
        assert info.n_callbacks == len(s.callbacks)


def test_invalid_descriptors_are_rejected_without_gpu():
    p = d.default_problem()
    p.system_id = 999
    h = C.c_void_p()
    assert d.lib.dfb_create(C.byref(p), 4, 0, C.byref(h)) == abi.ERR_INVALID
    p = d.default_problem()
    p.system_id = abi.SYS_LORENZ
    p.h = -1.0
    assert d.lib.dfb_create(C.byref(p), 4, 0, C.byref(h)) == abi.ERR_INVALID
    p = d.default_problem()
    p.system_id = abi.SYS_LORENZ
    p.n_loc = 1
    p.loc[0].kind = abi.LOC_STATEFN
    p.loc[0].detector_id = 3  # Lorenz registers no detector
    assert d.lib.dfb_create(C.byref(p), 4, 0, C.byref(h)) == abi.ERR_INVALID
    assert b"detector" in d.lib.dfb_last_error()
    # empty ensemble
    p = d.default_problem()
    p.system_id = abi.SYS_LORENZ
    assert d.lib.dfb_create(C.byref(p), 0, 0, C.byref(h)) == abi.ERR_INVALID
    # generic_order_3 with arguments the reference panics on (rk.rs:200)
    with pytest.raises(d.DfbError):
        d.RK.generic_order_3(2.0 / 3.0, 0.5)


@pytest.mark.skipif(has_cuda(), reason="CPU-only behaviour")
def test_no_cpu_fallback():
    s = d.Solver.new().initial([1., 2., 20.]).equation(d.systems.Lorenz, [10., 28., 8 / 3]).interval(0, 1)
    with pytest.raises(d.DfbError) as e:
        s.run()
    assert e.value.code == abi.ERR_NO_DEVICE


def test_solver_consumed_by_run_and_builder_semantics():
    s = d.Solver.new().initial([0.0]).equation(d.systems.Zero1)
    # with_const_delay raises max_delay (solver.rs:259); a later max_delay() overrides
    s.with_const_delay(1.0, 0)
    assert s.problem().max_delay == 1.0
    s.with_const_delay(1.25, 0).max_delay(0.5)
    p = s.problem()
    assert p.max_delay == 0.5 and p.n_loc == 2
    assert p.loc[0].kind == abi.LOC_PROPAGATION and p.loc[0].dedup == 0
    s2 = d.Solver.new().initial([1., 0.]).equation(d.systems.BouncingBall, [9.8, .9])
    s2.on(d.Locator.zero("dx")).on_mut(d.Locator.below_zero("x"), "bounce")
    p2 = s2.problem()
    assert [p2.loc[i].dedup for i in range(2)] == [1, 1]
    assert p2.loc[1].detection == abi.DET_BELOW_ZERO and p2.loc[1].location == abi.LOCATE_BISECTION
    assert p2.loc[1].callback_id == 1 and p2.loc[0].callback_id == abi.CB_NONE


def test_ctypes_layout_equals_the_c_header(tmp_path):
    """sizeof / offsetof of every struct of include/diffurch_b200.h, as gcc lays it out, against the
    ctypes mirror (and therefore, through tests/test_rust_sys_matches_abi.py, the Rust bindings)."""
    import ctypes as C
    import subprocess
    structs = {"dfb_tableau": abi.Tableau, "dfb_locator": abi.Locator, "dfb_auto_stepsize": abi.AutoStepsize,
               "dfb_problem": abi.Problem, "dfb_system_info": abi.SystemInfo, "dfb_stats": abi.Stats,
               "dfb_totals": abi.Totals}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "diffurch_b200.h"', "int main(void) {"]
    for name, ct in structs.items():
        lines.append(f'  printf("{name} %zu\\n", sizeof({name}));')
        for field, _ in ct._fields_:
            lines.append(f'  printf("{name}.{field} %zu\\n", offsetof({name}, {field}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = dict(ln.split() for ln in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for name, ct in structs.items():
        assert int(out[name]) == C.sizeof(ct), name
        for field, _ in ct._fields_:
            assert int(out[f"{name}.{field}"]) == getattr(ct, field).offset, f"{name}.{field}"
