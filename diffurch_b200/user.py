"""User-defined systems: `Solver.equation` for an equation the library does not ship.

The reference takes any Rust closure (``Solver::equation``, solver.rs:135; detectors and
callbacks through ``on`` / ``on_mut``).  A host closure cannot run on a GPU, so the user
restates it as a CUDA device functor — a struct with the signatures of
``csrc/dev/systems.cuh`` (``rhs`` / ``detector`` / ``callback`` over the same views the
reference's closures see: ``t, p, d, t_prev, p_prev, p(tau), d(tau)``; mutable ``t, p`` and
``stop_integration()`` for callbacks) — and this module compiles it, with nvcc for sm_100a,
against the SAME kernel templates the built-in systems use (register-resident fixed-step,
adaptive, periodic, event, DDE coefficient-ring and general kernels, both arithmetic
contracts) into a plugin shared object that `dfb_register_plugin` loads.

    so = build_user_system("van_der_pol.cuh", "SysVanDerPol")
    VanDerPol = load_user_system(so, params=("mu",))
    Solver.new().equation(VanDerPol, mu).initial(y0).interval(0, 20).stepsize(0.01).run()
"""
import concurrent.futures as cf
import ctypes as C
import hashlib
import os
import subprocess

from . import _abi as abi
from ._lib import check, lib
from .build import ARCH, COMMON, CSRC, NVCC
from .systems import System


def _jobs():
    jobs = []
    for arith, tag, extra in ((0, "strict", ["-fmad=false"]), (1, "fast", [])):
        for src in ("kernels_ode_fixed.cu", "kernels_ode_event.cu", "kernels_general.cu"):
            jobs.append((src, f"{src[:-3]}_{tag}.o", [f"-DDFB_ARITH={arith}", "-DDFB_GROUP=1"] + extra))
    jobs.append(("kernels_ode_adaptive.cu", "kernels_ode_adaptive_fast.o", ["-DDFB_ARITH=1"]))
    jobs.append(("kernels_dde_fixed.cu", "kernels_dde_fixed_fast.o", ["-DDFB_ARITH=1"]))
    jobs.append(("kernels_dde_event.cu", "kernels_dde_event_fast.o", ["-DDFB_ARITH=1"]))
    jobs.append(("plugin_entry.cu", "plugin_entry.o", []))
    return jobs


def build_user_system(header, struct_name, out_dir=None, force=False, verbose=False):
    """Compile the device functor `struct_name` defined in `header` into a plugin .so.

    Returns the path of the shared object (cached by the header's content hash)."""
    header = os.path.abspath(header)
    src = open(header, "rb").read()
    tag = hashlib.sha256(src + struct_name.encode()).hexdigest()[:12]
    out_dir = os.path.abspath(out_dir or os.path.join(os.path.dirname(header), "_build"))
    obj_dir = os.path.join(out_dir, f"{struct_name}_{tag}_obj")
    so = os.path.join(out_dir, f"libdfb_user_{struct_name}_{tag}.so")
    # kernel sources the plugin is compiled from (csrc/host is the C++ builder mirror: not a dependency)
    newest_dep = max(os.path.getmtime(os.path.join(r, f)) for r, _, fs in os.walk(CSRC) for f in fs
                     if os.path.basename(r) != "host" and f != "api.cu")
    newest_dep = max(newest_dep, os.path.getmtime(os.path.join(CSRC, "..", "..", "include", "diffurch_b200.h")))
    if not force and os.path.exists(so) and os.path.getmtime(so) >= max(newest_dep, os.path.getmtime(header)):
        return so
    os.makedirs(obj_dir, exist_ok=True)
    defs = [f'-DDFB_PLUGIN_HEADER="{header}"', f"-DDFB_PLUGIN_SYSTEM={struct_name}", "-I", CSRC]

    def compile_one(job):
        s, o, flags = job
        cmd = [NVCC] + ARCH + COMMON + defs + flags + ["-c", os.path.join(CSRC, s), "-o", os.path.join(obj_dir, o)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return o, r.returncode, r.stdout + r.stderr, cmd

    jobs = _jobs()
    with cf.ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
        for o, rc, log, cmd in ex.map(compile_one, jobs):
            if verbose:
                print(f"[diffurch_b200.user] {o}: rc={rc}")
            if rc != 0:
                raise RuntimeError("nvcc failed on " + o + ":\n" + " ".join(cmd) + "\n" + log)
    objs = [os.path.join(obj_dir, j[1]) for j in jobs]
    r = subprocess.run([NVCC] + ARCH + ["-shared", "-o", so] + objs + ["-lcudart"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return so


def build_user_system_from_source(source, struct_name, out_dir, **kw):
    """Same as :func:`build_user_system` for functor code held in a string (e.g. generated from a
    symbolic right-hand side): the text is written to ``<out_dir>/<struct_name>.cuh`` first.  It may
    omit the ``#include "user_system.cuh"`` / ``namespace DFB_NS`` wrapper."""
    os.makedirs(out_dir, exist_ok=True)
    if "user_system.cuh" not in source:
        source = '#include "user_system.cuh"\nnamespace DFB_NS {\n' + source + "\n}  // namespace DFB_NS\n"
    path = os.path.join(out_dir, struct_name + ".cuh")
    if not os.path.exists(path) or open(path).read() != source:
        with open(path, "w") as f:
            f.write(source)
    return build_user_system(path, struct_name, out_dir=os.path.join(out_dir, "_build"), **kw)


def load_user_system(so_path, name=None, params=(), detectors=None, callbacks=None, reference=""):
    """Register a plugin built by :func:`build_user_system`; returns the `System` descriptor that
    ``Solver.equation`` takes.  `params`, `detectors` and `callbacks` name the functor's parameter
    slots and its system-local detector / callback ids (1-based, as in the shipped systems)."""
    sid = C.c_int32()
    check(lib.dfb_register_plugin(os.path.abspath(so_path).encode(), C.byref(sid)))
    info = abi.SystemInfo()
    check(lib.dfb_system_info_get(sid.value, C.byref(info)))
    params = tuple(params)
    if len(params) != info.n_params:
        raise ValueError(f"the functor declares NP = {info.n_params} parameters, {len(params)} names given")
    from . import systems
    return systems.register(System(name or os.path.basename(so_path), sid.value, info.n_state, params,
                                   n_aux=info.n_aux, detectors=dict(detectors or {}),
                                   callbacks=dict(callbacks or {}), uses_history=bool(info.uses_history),
                                   reference=reference))


def load_user_system_from_source(source, struct_name, name=None, params=(), detectors=None, callbacks=None,
                                 reference=""):
    """Register a device functor given as SOURCE TEXT; no nvcc needed.  The functor is compiled at run time
    with NVRTC (`dfb_register_system_from_source`), one kernel instantiation at a time on first use, from
    the same kernel templates as the built-in systems.  Fixed-step problems without events take the
    register-resident hot kernel, everything else the general kernel."""
    if os.path.exists(source):
        source = open(source).read()
    sid = C.c_int32()
    check(lib.dfb_register_system_from_source(source.encode(), struct_name.encode(), C.byref(sid)))
    info = abi.SystemInfo()
    check(lib.dfb_system_info_get(sid.value, C.byref(info)))
    params = tuple(params)
    if len(params) != info.n_params:
        raise ValueError(f"the functor declares NP = {info.n_params} parameters, {len(params)} names given")
    from . import systems
    return systems.register(System(name or struct_name, sid.value, info.n_state, params, n_aux=info.n_aux,
                                   detectors=dict(detectors or {}), callbacks=dict(callbacks or {}),
                                   uses_history=bool(info.uses_history), reference=reference))


def nvrtc_stats():
    """Run-time compilation counters: kernels compiled, disk-cache hits, total / last compile time (ms)."""
    a, b = C.c_uint32(), C.c_uint32()
    t, l = C.c_double(), C.c_double()
    check(lib.dfb_nvrtc_stats(C.byref(a), C.byref(b), C.byref(t), C.byref(l)))
    return {"kernels_compiled": a.value, "disk_cache_hits": b.value, "compile_ms_total": t.value,
            "last_compile_ms": l.value}
