"""Loader of the in-tree CUDA library ``libdiffurch_b200.so`` (the C ABI).

The product has no CPU execution path: if the library is missing this module
raises at import, and without a CUDA device ``dfb_create`` fails with
``ERR_NO_DEVICE`` (surfaced as :class:`DfbError`).
"""
import ctypes as C
import os

from . import _abi as abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DFB_LIB_PATH") or os.path.join(_HERE, "libdiffurch_b200.so")  # DFB_LIB_PATH: A/B of two builds (tuning)


class DfbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"diffurch_b200 error {code}: {msg}")
        self.code = code


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m diffurch_b200.build` "
            "(nvcc, sm_100a).  diffurch_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    dp = C.POINTER(C.c_double)
    H = C.c_void_p
    sigs = {
        "dfb_tableau_by_name": (C.c_int, [C.c_char_p, C.POINTER(abi.Tableau)]),
        "dfb_tableau_generic_order_2": (C.c_int, [C.c_double, C.POINTER(abi.Tableau)]),
        "dfb_tableau_generic_order_3": (C.c_int, [C.c_double, C.c_double, C.c_int, C.c_double,
                                                  C.POINTER(abi.Tableau)]),
        "dfb_problem_default": (C.c_int, [C.POINTER(abi.Problem)]),
        "dfb_system_info_get": (C.c_int, [C.c_int32, C.POINTER(abi.SystemInfo)]),
        "dfb_register_plugin": (C.c_int, [C.c_char_p, C.POINTER(C.c_int32)]),
        "dfb_register_system_from_source": (C.c_int, [C.c_char_p, C.c_char_p, C.POINTER(C.c_int32)]),
        "dfb_nvrtc_stats": (C.c_int, [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), dp, dp]),
        "dfb_nvrtc_prebuild": (C.c_int, [C.c_int32, C.c_int32, C.POINTER(abi.Tableau), C.c_int32, C.c_int32]),
        "dfb_create": (C.c_int, [C.POINTER(abi.Problem), C.c_size_t, C.c_int, C.POINTER(H)]),
        "dfb_destroy": (None, [H]),
        "dfb_set_initial": (C.c_int, [H, dp]),
        "dfb_set_params": (C.c_int, [H, dp]),
        "dfb_set_initial_device": (C.c_int, [H, C.c_void_p]),
        "dfb_set_params_device": (C.c_int, [H, C.c_void_p]),
        "dfb_run": (C.c_int, [H, C.c_void_p]),
        "dfb_synchronize": (C.c_int, [H]),
        "dfb_get_final": (C.c_int, [H, dp, dp, dp]),
        "dfb_get_stats": (C.c_int, [H, C.POINTER(abi.Stats)]),
        "dfb_get_totals": (C.c_int, [H, C.POINTER(abi.Totals)]),
        "dfb_final_device": (C.c_int, [H, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                       C.POINTER(C.c_void_p)]),
        "dfb_get_trace": (C.c_int, [H, C.POINTER(C.c_uint32), dp, dp]),
        "dfb_get_events": (C.c_int, [H, C.POINTER(C.c_uint32), dp, C.POINTER(C.c_int32)]),
        "dfb_gather_final": (C.c_int, [H, C.c_void_p, C.c_int, dp, dp, dp]),
        "dfb_measure_fp64_peak": (C.c_int, [C.c_int, dp, dp]),
        "dfb_last_error": (C.c_char_p, []),
        "dfb_abi_version": (C.c_int, []),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.dfb_abi_version() != abi.ABI_VERSION:
        raise ImportError("libdiffurch_b200.so ABI version mismatch; rebuild")
    return lib


lib = _load()


def check(rc):
    if rc != 0:
        raise DfbError(rc, lib.dfb_last_error().decode())
    return rc
