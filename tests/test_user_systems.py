"""User-defined systems (dfb_register_plugin): a device functor the library does not ship is
compiled against the same kernel templates as the built-in systems and must behave like them.

The plugins are built by `__graft_entry__.build()` (tests/plugins/__init__.py); the CPU tests
check loading and the registry, the GPU tests check numerics:
  * SysUserLorenz / SysUserDde restate built-in systems -> bit-identical results in every
    kernel family (fixed, adaptive, general, DDE coefficient ring);
  * SysVanDerPol is new -> checked against a numpy restatement of the reference's make_step
    (state.rs:94-120, operation order of SURVEY Appendix A.2) with the golden rktp64 tableau.
"""
import ctypes as C
import json
import os

import numpy as np
import pytest

import cases
import diffurch_b200 as d
import plugins
from diffurch_b200 import user

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def built():
    return plugins.build_all()   # cached: a no-op when __graft_entry__.build() has run


def test_plugin_loads_and_registers(built):
    vdp = user.load_user_system(built["van_der_pol"], name="van_der_pol", params=("mu",),
                                detectors={"x": 1}, callbacks={"count": 1})
    assert vdp.id >= d.abi.SYS_USER_BASE and vdp.n_state == 2 and vdp.n_params == 1 and vdp.n_aux == 1
    again = user.load_user_system(built["van_der_pol"], params=("mu",))
    assert again.id == vdp.id                       # registering the same path twice is idempotent
    lor = user.load_user_system(built["user_lorenz"], params=("sigma", "rho", "beta"))
    assert lor.id != vdp.id and lor.n_state == 3 and not lor.uses_history
    dde = user.load_user_system(built["user_dde"], params=("a", "b", "tau", "k"))
    assert dde.uses_history
    info = d.abi.SystemInfo()
    assert d.lib.dfb_system_info_get(vdp.id, C.byref(info)) == 0 and info.n_detectors == 1
    with pytest.raises(ValueError):
        user.load_user_system(built["user_lorenz"], params=("only_one",))
    sid = C.c_int32()
    assert d.lib.dfb_register_plugin(b"/nonexistent/libnothing.so", C.byref(sid)) == d.abi.ERR_INVALID
    assert d.lib.dfb_register_plugin(d.LIB_PATH.encode(), C.byref(sid)) == d.abi.ERR_INVALID  # not a plugin
    assert d.lib.dfb_system_info_get(d.abi.SYS_USER_BASE + 999, C.byref(info)) == d.abi.ERR_INVALID


# ---------------------------------------------------------------------------------------- GPU
def golden_tableau(name):
    t = json.load(open(os.path.join(HERE, "golden", "tableaux.json")))["tableaux"][name]
    f = lambda v: np.array([float.fromhex(x) if isinstance(x, str) else float(x) for x in v])
    S = len(t["b"])
    return np.array([f(r) for r in t["a"]]).reshape(S, S), f(t["b"]), f(t["c"])


def numpy_fixed_step(rhs, y0, t0, t1, h, tab):
    """Solver::run with a scalar stepsize (solver.rs:292-310) + make_step (state.rs:94-120):
    separately rounded multiplies and adds, left folds from zero, d_curr reused as k[0]."""
    a, b, c = tab
    S = len(b)
    t = t0
    p = np.array(y0, dtype=np.float64)
    d_curr = None
    n = 0
    while t < t1:
        hs = min(h, t1 - t)
        k = [rhs(t, p) if d_curr is None else d_curr]
        for i in range(1, S):
            acc = np.zeros_like(p)
            for j in range(i):
                acc = acc + k[j] * a[i, j]
            k.append(rhs(t + c[i] * hs, p + acc * hs))
        acc = np.zeros_like(p)
        for j in range(S):
            acc = acc + k[j] * b[j]
        p = p + acc * hs
        t = t + hs
        d_curr = rhs(t, p)
        n += 1
    return t, p, n


@pytest.mark.gpu
def test_user_lorenz_is_bit_identical_to_builtin(built):
    ulor = user.load_user_system(built["user_lorenz"], params=("sigma", "rho", "beta"))
    n = 300
    rho = 20.0 + 20.0 * np.arange(n) / (n - 1)
    prm = np.stack([np.full(n, 10.0), rho, np.full(n, 8.0 / 3.0)], axis=1)
    y0 = np.tile([1.0, 2.0, 20.0], (n, 1))
    auto = d.AutomaticStepsize(1e-2, (1e-6, 1e-1), [1e-9] * 3, [1e-9] * 3, 4, 0.9, (0.2, 5.0))
    for arith in ("strict", "fast"):
        for variant in ("fixed", "rk4", "adaptive", "trace"):
            def mk(system):
                s = d.Solver.new().initial(y0).equation(system, prm).interval(0.0, 3.0).stepsize(1.0 / 250.0).arith(arith)
                if variant == "rk4":
                    s = s.rk(d.RK.rk4())
                if variant == "adaptive":
                    s = s.stepsize(auto)
                if variant == "trace":
                    s = s.on_step(d.Trace(50, 32))
                return s
            u, b = mk(ulor).run(), mk(d.systems.Lorenz).run()
            assert u.totals["kernel_kind"] == b.totals["kernel_kind"], (arith, variant)
            assert np.array_equal(u.p, b.p) and np.array_equal(u.t, b.t), (arith, variant)
            assert np.array_equal(u.steps, b.steps) and np.array_equal(u.rejected, b.rejected)
            assert np.array_equal(u.rhs_evals, b.rhs_evals) and np.all(u.status == 0)
            if variant == "trace":
                assert np.array_equal(u.trace_p, b.trace_p)


@pytest.mark.gpu
def test_user_dde_is_bit_identical_to_builtin(built):
    udde = user.load_user_system(built["user_dde"], params=("a", "b", "tau", "k"))
    k = 0.5 + np.arange(200) / 199.0
    prm = cases.dde_params(k)
    for arith, kind in (("fast", 2), ("strict", 0)):
        def mk(system):
            return (d.Solver.new().max_delay(1.0).initial(d.InitFn(d.SIN)).interval(0.0, 10.0).stepsize(0.02)
                    .equation(system, prm).arith(arith))
        u, b = mk(udde).run(), mk(d.systems.DdeLinear).run()
        assert u.totals["kernel_kind"] == kind and b.totals["kernel_kind"] == kind
        assert np.array_equal(u.p, b.p) and np.array_equal(u.steps, b.steps) and np.all(u.status == 0)
        assert np.max(np.abs(u.p[:, 0] - np.sin(k * u.t))) < 1e-11      # tests/equation_types.rs:133


@pytest.mark.gpu
def test_van_der_pol_against_numpy_restatement(built):
    vdp = user.load_user_system(built["van_der_pol"], name="van_der_pol", params=("mu",),
                                detectors={"x": 1}, callbacks={"count": 1})
    mu = np.array([0.5, 1.0, 2.0, 3.5])
    y0 = np.tile([2.0, 0.0], (4, 1))
    h, t1 = 0.01, 3.0
    g = (d.Solver.new().initial(y0).equation(vdp, mu.reshape(-1, 1)).interval(0.0, t1).stepsize(h)
         .arith("strict").run())
    assert g.totals["kernel_kind"] == 1 and np.all(g.status == 0)
    tab = golden_tableau("rktp64")
    for i, m in enumerate(mu):
        rhs = lambda t, p, m=m: np.array([p[1], m * (1.0 - p[0] * p[0]) * p[1] - p[0]])
        t, p, n = numpy_fixed_step(rhs, y0[i], 0.0, t1, h, tab)
        assert g.steps[i] == n and g.t[i] == t
        assert np.array_equal(g.p[i], p), (g.p[i], p)         # STRICT: the reference's rounding, bit for bit
    f = (d.Solver.new().initial(y0).equation(vdp, mu.reshape(-1, 1)).interval(0.0, t1).stepsize(h)
         .arith("fast").run())
    assert np.max(np.abs(f.p - g.p) / np.maximum(np.abs(g.p), 1.0)) < 1e-12   # north_star tolerance


@pytest.mark.gpu
def test_van_der_pol_events_on_user_detector(built):
    vdp = user.load_user_system(built["van_der_pol"], name="van_der_pol", params=("mu",),
                                detectors={"x": 1}, callbacks={"count": 1})
    mu = np.linspace(0.2, 3.0, 50).reshape(-1, 1)
    y0 = np.tile([2.0, 0.0], (50, 1))

    def mk(arith):
        return (d.Solver.new().initial(y0).equation(vdp, mu).interval(0.0, 30.0).stepsize(0.01)
                .on_mut(d.Locator.zero("x"), "count").log_events(d.EventLog(64)).arith(arith))
    s = mk("strict").run()
    assert s.totals["kernel_kind"] == 5 and np.all(s.status == 0)
    assert np.array_equal(s.aux[:, 0], s.events.astype(np.float64)) and np.all(s.events >= 4)
    gen = None
    os.environ["DFB_FORCE_GENERAL"] = "1"
    try:
        gen = mk("strict").run()
    finally:
        del os.environ["DFB_FORCE_GENERAL"]
    assert gen.totals["kernel_kind"] == 0
    assert np.array_equal(gen.event_t, s.event_t) and np.array_equal(gen.p, s.p)
    # x(t*) = 0 at every located time: re-integrate to the first event time and look at x
    i = 7
    t_ev = s.event_t[i, 0]
    r = (d.Solver.new().initial(y0[i:i + 1]).equation(vdp, mu[i:i + 1]).interval(0.0, t_ev).stepsize(0.01)
         .arith("strict").run())
    assert abs(r.p[0, 0]) < 1e-12
    f = mk("fast").run()
    assert np.array_equal(f.events, s.events) and np.max(np.abs(f.event_t - s.event_t)) < 1e-10


@pytest.mark.gpu
def test_callback_may_switch_a_captured_parameter(built):
    # examples/loc-switch-parameter.rs: gravity flips at every crossing of x = 0 (a `Cell` captured by
    # the equation and the callback upstream; here the callback writes the parameter block)
    sw = user.load_user_system(built["switch_parameter"], params=("g", "k"), detectors={"dx": 1, "x": 2},
                               callbacks={"switch": 1})
    n = 40
    x0 = 1.0 + np.arange(n) / n
    prm = np.tile([9.8, 0.9], (n, 1))

    def mk(arith, t1=6.0):
        return (d.Solver.new().initial(np.stack([x0, np.full(n, -0.01)], 1)).equation(sw, prm).interval(0.0, t1)
                .stepsize(0.005).on(d.Locator.zero("dx"), None).on_mut(d.Locator.zero("x"), "switch")
                .log_events(d.EventLog(128)).arith(arith))
    s = mk("strict").run()
    assert s.totals["kernel_kind"] == d.abi.KERNEL_ODE_EVENT and np.all(s.status == 0)
    crossings = s.aux[:, 0]
    assert np.all(crossings >= 3)
    # first crossing = the free-fall time of the built-in bouncing ball; after it gravity points up,
    # so the particle is below zero until the next crossing: sign(x) = (-1)^crossings
    t1 = (-0.01 + np.sqrt(0.0001 + 2 * 9.8 * x0)) / 9.8
    first = s.event_t[np.arange(n), np.argmax(s.event_index == 1, axis=1)]
    assert np.max(np.abs(first - t1)) < 1e-13
    assert np.all(np.sign(s.p[:, 0]) == (-1.0) ** crossings)
    gen = None
    os.environ["DFB_FORCE_GENERAL"] = "1"
    try:
        gen = mk("strict").run()
    finally:
        del os.environ["DFB_FORCE_GENERAL"]
    assert gen.totals["kernel_kind"] == d.abi.KERNEL_GENERAL
    assert np.array_equal(gen.p, s.p) and np.array_equal(gen.event_t, s.event_t) and np.array_equal(gen.aux, s.aux)
    f = mk("fast").run()
    assert np.array_equal(f.aux, s.aux) and np.max(np.abs(f.event_t - s.event_t)) < 1e-10


# ---------------------------------------------------------------------------- from SOURCE (NVRTC)
VDP_SOURCE = open(plugins.header("van_der_pol")).read()
INLINE_SOURCE = """
struct SysDuffingInline : SysBase {          // x'' + delta x' + alpha x + beta x^3 = 0, no wrapper around it
  static constexpr int N = 2, NP = 3;
  template <class V>
  __device__ __forceinline__ static void rhs(const V& s, const double* prm, double* out) {
    out[0] = s.p[1];
    out[1] = -prm[0] * s.p[1] - prm[1] * s.p[0] - prm[2] * (s.p[0] * s.p[0]) * s.p[0];
  }
};
"""


def test_source_system_registers_and_builds_without_a_gpu_or_nvcc(tmp_path, monkeypatch):
    """dfb_register_system_from_source: the functor's static facts come out of NVRTC's output (no GPU, no
    nvcc on PATH), and the kernel instantiation a problem needs compiles to a CUBIN for sm_100a."""
    monkeypatch.setenv("PATH", "/nonexistent")            # nvcc is not reachable
    monkeypatch.setenv("DFB_NVRTC_CACHE", str(tmp_path))
    vdp = user.load_user_system_from_source(VDP_SOURCE, "SysVanDerPol", params=("mu",), detectors={"x": 1},
                                            callbacks={"count": 1})
    assert vdp.id >= d.abi.SYS_USER_BASE and (vdp.n_state, vdp.n_params, vdp.n_aux) == (2, 1, 1)
    again = user.load_user_system_from_source(VDP_SOURCE, "SysVanDerPol", params=("mu",))
    assert again.id == vdp.id                              # idempotent per (source, struct)
    duf = user.load_user_system_from_source(INLINE_SOURCE, "SysDuffingInline", params=("delta", "alpha", "beta"))
    assert duf.id != vdp.id and duf.n_state == 2 and not duf.uses_history
    before = user.nvrtc_stats()
    for arith in (d.abi.ARITH_FAST, d.abi.ARITH_STRICT):
        rk = d.RK.rktp64()
        assert d.lib.dfb_nvrtc_prebuild(vdp.id, arith, C.byref(rk), 2, 0) == 0, d.lib.dfb_last_error()
    rk4 = d.RK.rk4()
    assert d.lib.dfb_nvrtc_prebuild(duf.id, d.abi.ARITH_FAST, C.byref(rk4), 1, 1) == 0, d.lib.dfb_last_error()
    after = user.nvrtc_stats()
    assert after["kernels_compiled"] == before["kernels_compiled"] + 3 and after["last_compile_ms"] > 0.0
    assert len([f for f in os.listdir(tmp_path) if f.endswith(".cubin")]) == 3
    # the same request again is served from the disk cache
    assert d.lib.dfb_nvrtc_prebuild(duf.id, d.abi.ARITH_FAST, C.byref(rk4), 1, 1) == 0
    assert user.nvrtc_stats()["disk_cache_hits"] == after["disk_cache_hits"] + 1
    sid = C.c_int32()
    assert d.lib.dfb_register_system_from_source(b"struct Broken : SysBase { int", b"Broken", C.byref(sid)) == d.abi.ERR_INVALID
    assert b"NVRTC" in d.lib.dfb_last_error()
    assert d.lib.dfb_nvrtc_prebuild(d.abi.SYS_LORENZ, 1, C.byref(rk4), 1, 0) == d.abi.ERR_INVALID  # not a source system


@pytest.mark.gpu
def test_source_systems_are_bit_identical_to_the_nvcc_plugins(built, tmp_path, monkeypatch):
    """Van der Pol and the user Lorenz restatement, registered from source with nvcc absent from PATH: the
    hot fixed-step kernel (exact tableau sparsity) and the general kernel give the same bits as the plugin
    built by nvcc from the same functor; the first-call compile time is reported."""
    monkeypatch.setenv("PATH", "/nonexistent")
    monkeypatch.setenv("DFB_NVRTC_CACHE", str(tmp_path))
    vdp_p = user.load_user_system(built["van_der_pol"], name="van_der_pol", params=("mu",), detectors={"x": 1},
                                  callbacks={"count": 1})
    vdp_s = user.load_user_system_from_source(VDP_SOURCE, "SysVanDerPol", name="van_der_pol_src", params=("mu",),
                                              detectors={"x": 1}, callbacks={"count": 1})
    lor_p = user.load_user_system(built["user_lorenz"], params=("sigma", "rho", "beta"))
    lor_s = user.load_user_system_from_source(plugins.header("user_lorenz"), "SysUserLorenz", name="user_lorenz_src",
                                              params=("sigma", "rho", "beta"))
    mu = np.linspace(0.3, 3.5, 333).reshape(-1, 1)
    n = 300
    prm = np.stack([np.full(n, 10.0), 20.0 + 20.0 * np.arange(n) / (n - 1), np.full(n, 8.0 / 3.0)], axis=1)
    auto = d.AutomaticStepsize(1e-2, (1e-6, 1e-1), [1e-9] * 3, [1e-9] * 3, 4, 0.9, (0.2, 5.0))
    t_first = None
    for arith in ("strict", "fast"):
        for variant in ("fixed", "rk4", "trace", "adaptive", "event"):
            def mk_vdp(system):
                s = (d.Solver.new().initial(np.tile([2.0, 0.0], (333, 1))).equation(system, mu).interval(0.0, 3.0)
                     .stepsize(0.01).arith(arith))
                if variant == "rk4":
                    s = s.rk(d.RK.rk4())
                if variant == "trace":
                    s = s.on_step(d.Trace(25, 16))
                if variant == "adaptive":
                    s = s.stepsize(d.AutomaticStepsize(1e-2, (1e-6, 1e-1), [1e-9] * 2, [1e-9] * 2, 4, 0.9, (0.2, 5.0)))
                if variant == "event":
                    s = s.on_mut(d.Locator.zero("x"), "count").log_events(d.EventLog(16))
                return s

            def mk_lor(system):
                s = d.Solver.new().initial(np.tile([1.0, 2.0, 20.0], (n, 1))).equation(system, prm).interval(0.0, 2.0).stepsize(1.0 / 250.0).arith(arith)
                if variant == "rk4":
                    s = s.rk(d.RK.rk4())
                if variant == "trace":
                    s = s.on_step(d.Trace(50, 16))
                if variant == "adaptive":
                    s = s.stepsize(auto)
                return s
            before = user.nvrtc_stats()
            a, b = mk_vdp(vdp_s).run(), mk_vdp(vdp_p).run()
            if t_first is None:
                t_first = user.nvrtc_stats()["last_compile_ms"]
            hot = variant in ("fixed", "rk4", "trace")
            assert a.totals["kernel_kind"] == (d.abi.KERNEL_ODE_FIXED if hot else d.abi.KERNEL_GENERAL), (arith, variant)
            # the plugin takes the adaptive / event HOT kernels where the source system takes the general one:
            # the same bits in STRICT (one arithmetic contract), the FAST tolerance otherwise
            exact = hot or arith == "strict"
            if exact:
                assert np.array_equal(a.p, b.p) and np.array_equal(a.t, b.t) and np.array_equal(a.steps, b.steps), (arith, variant)
                assert np.array_equal(a.rhs_evals, b.rhs_evals), (arith, variant)
            else:
                assert np.max(np.abs(a.p - b.p)) < 1e-7 and np.max(np.abs(a.steps.astype(np.int64) - b.steps.astype(np.int64))) <= 2
            assert np.all(a.status == 0), (arith, variant)
            if variant == "trace":
                assert np.array_equal(a.trace_p, b.trace_p) and np.array_equal(a.trace_n, b.trace_n)
            if variant == "event":
                assert np.array_equal(a.events, b.events) and np.array_equal(a.aux, b.aux) and np.sum(a.events) >= 100
                assert np.array_equal(a.event_t, b.event_t) if exact else np.max(np.abs(a.event_t - b.event_t)) < 1e-10
            if variant == "adaptive" and exact:
                assert np.array_equal(a.rejected, b.rejected)
            if variant != "event":
                u, v = mk_lor(lor_s).run(), mk_lor(lor_p).run()
                if exact:
                    assert np.array_equal(u.p, v.p) and np.array_equal(u.steps, v.steps), (arith, variant)
                    assert np.array_equal(u.p, mk_lor(d.systems.Lorenz).run().p), (arith, variant)
                else:
                    assert np.max(np.abs(u.p - v.p) / np.maximum(np.abs(v.p), 1.0)) < 1e-5, (arith, variant)
                assert np.all(u.status == 0)
            assert user.nvrtc_stats()["kernels_compiled"] >= before["kernels_compiled"]
    st = user.nvrtc_stats()
    assert st["kernels_compiled"] >= 10 and t_first is not None and t_first > 0.0
    print(f"[nvrtc] {st['kernels_compiled']} kernels compiled in {st['compile_ms_total']:.0f} ms "
          f"(first call {t_first:.0f} ms, mean {st['compile_ms_total'] / st['kernels_compiled']:.0f} ms)")
