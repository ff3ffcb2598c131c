"""ctypes mirror of ``include/diffurch_b200.h`` (the C ABI).

Pure declarations: struct layouts, enum values, no behaviour.  Kept in lock-step
with the header; ``tests/test_abi.py`` checks sizes/offsets against the values the
compiled library reports.
"""
import ctypes as C

ABI_VERSION = 3
MAX_STAGES = 7
MAX_INTERP = 5
MAX_STATE = 12
MAX_PARAMS = 12
MAX_AUX = 4
MAX_LOCATORS = 4
MAX_DISCO = 8

# dfb_status
OK, ERR_INVALID, ERR_UNSUPPORTED, ERR_NO_DEVICE, ERR_CUDA, ERR_STATE, ERR_NCCL = 0, -1, -2, -3, -4, -5, -6

# per-trajectory status bits
TRAJ_OK = 0
TRAJ_HISTORY_DELETED = 1
TRAJ_HISTORY_NOT_READY = 2
TRAJ_NO_DERIVATIVE = 4
TRAJ_NONFINITE = 8
TRAJ_STEP_LIMIT = 16
TRAJ_RING_OVERFLOW = 32
TRAJ_STOPPED = 64
TRAJ_DISCO_OVERFLOW = 128
TRAJ_STEPSIZE_FLOOR = 256
# dfb_kernel_kind (dfb_totals.kernel_kind)
KERNEL_GENERAL, KERNEL_ODE_FIXED, KERNEL_DDE_FIXED, KERNEL_ODE_ADAPTIVE = 0, 1, 2, 3
KERNEL_ODE_FIXED_PERIODIC, KERNEL_ODE_EVENT, KERNEL_DDE_EVENT = 4, 5, 6

# dfb_system_id
SYS_LORENZ = 1
SYS_LORENZ_VARIATIONAL = 2
SYS_LINEAR1 = 3
SYS_LINEAR1_SIN = 4
SYS_LINEAR3 = 5
SYS_LINEAR3_MATRIX = 6
SYS_HARMONIC = 7
SYS_POWER_DECAY = 8
SYS_CONSTANT3 = 9
SYS_TIME3 = 10
SYS_ZERO1 = 11
SYS_DDE_LINEAR = 12
SYS_NDDE_LINEAR = 13
SYS_BOUNCING_BALL = 14
SYS_EGG_BILLIARD = 15
SYS_RELAY_MSIGN = 16
SYS_DDE_RELAY_CLAMP = 17
SYS_USER_BASE = 1000  # ids of systems registered with dfb_register_plugin start here

# dfb_history_id
HIST_CONSTANT, HIST_SIN, HIST_SIN_NODERIV, HIST_INV_SQUARE = 0, 1, 2, 3

# dfb_locator_kind
LOC_STATEFN, LOC_PERIODIC, LOC_PROPAGATION = 0, 1, 2

# dfb_detection
(DET_ZERO, DET_ABOVE_ZERO, DET_BELOW_ZERO, DET_POSITIVE, DET_NEGATIVE, DET_SWITCH,
 DET_SWITCH_TRUE, DET_SWITCH_FALSE, DET_IS_TRUE, DET_IS_FALSE, DET_STEP) = range(11)

# dfb_location
(LOCATE_STEP_BEGIN, LOCATE_STEP_END, LOCATE_STEP_MIDDLE, LOCATE_LERP, LOCATE_BISECTION,
 LOCATE_BISECTION_BOOL, LOCATE_REGULA_FALSI) = range(7)

# dfb_filter
FILTER_NONE, FILTER_EVERY, FILTER_SEPARATED_BY, FILTER_IN_INTERVAL, FILTER_ON_TIMES = 0, 1, 2, 3, 4
MAX_FILTER_TIMES = 8

CB_NONE = 0

STEP_FIXED, STEP_AUTOMATIC = 0, 1
ARITH_STRICT, ARITH_FAST = 0, 1


class Tableau(C.Structure):
    _fields_ = [
        ("S", C.c_int32), ("I", C.c_int32),
        ("order", C.c_int32), ("order_embedded", C.c_int32), ("order_interpolant", C.c_int32),
        ("_pad", C.c_int32),
        ("a", C.c_double * (MAX_STAGES * MAX_STAGES)),
        ("b", C.c_double * MAX_STAGES),
        ("b2", C.c_double * MAX_STAGES),
        ("c", C.c_double * MAX_STAGES),
        ("bi", C.c_double * (MAX_STAGES * MAX_INTERP)),
    ]


class Locator(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("detector_id", C.c_int32), ("detection", C.c_int32),
        ("location", C.c_int32), ("callback_id", C.c_int32), ("dedup", C.c_int32),
        ("smoothing_order", C.c_int32), ("filter", C.c_int32), ("filter_n", C.c_int32),
        ("_pad", C.c_int32),
        ("period", C.c_double), ("offset", C.c_double), ("delay", C.c_double),
        ("filter_a", C.c_double), ("filter_b", C.c_double),
        ("filter_times", C.c_int32 * MAX_FILTER_TIMES),
    ]


class AutoStepsize(C.Structure):
    _fields_ = [
        ("stepsize", C.c_double),
        ("stepsize_min", C.c_double), ("stepsize_max", C.c_double),
        ("atol", C.c_double * MAX_STATE), ("rtol", C.c_double * MAX_STATE),
        ("order", C.c_uint32), ("_pad", C.c_uint32),
        ("fac", C.c_double), ("fac_min", C.c_double), ("fac_max", C.c_double),
        ("initial_stepsize", C.c_double),
    ]


class Problem(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("system_id", C.c_int32), ("history_id", C.c_int32),
        ("arith", C.c_int32),
        ("rk", Tableau),
        ("t0", C.c_double), ("t1", C.c_double),
        ("stepsize_kind", C.c_int32), ("n_loc", C.c_int32),
        ("h", C.c_double),
        ("automatic", AutoStepsize),
        ("max_delay", C.c_double),
        ("n_disco", C.c_int32), ("on_start_callback_id", C.c_int32),
        ("disco_t", C.c_double * MAX_DISCO),
        ("disco_order", C.c_int32 * MAX_DISCO),
        ("loc", Locator * MAX_LOCATORS),
        ("trace_every", C.c_int32), ("trace_capacity", C.c_int32),
        ("event_capacity", C.c_int32), ("ring_capacity", C.c_int32),
        ("max_steps", C.c_int64),
    ]


class SystemInfo(C.Structure):
    _fields_ = [
        ("n_state", C.c_int32), ("n_params", C.c_int32), ("n_aux", C.c_int32),
        ("n_detectors", C.c_int32), ("n_callbacks", C.c_int32), ("uses_history", C.c_int32),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("steps", C.POINTER(C.c_uint64)), ("rejected", C.POINTER(C.c_uint64)),
        ("events", C.POINTER(C.c_uint64)), ("rhs_evals", C.POINTER(C.c_uint64)),
        ("status", C.POINTER(C.c_uint32)),
    ]


class Totals(C.Structure):
    _fields_ = [
        ("n_traj", C.c_uint64),
        ("steps", C.c_uint64), ("rejected", C.c_uint64), ("events", C.c_uint64),
        ("rhs_evals", C.c_uint64), ("n_failed", C.c_uint64),
        ("kernel_ms", C.c_double),
        ("kernel_launches", C.c_uint32), ("kernel_kind", C.c_uint32),
        ("locate_iters", C.c_uint64),
    ]


# Every symbol include/diffurch_b200.h declares (checked by tests/test_abi.py).
EXPORTED_SYMBOLS = [
    "dfb_tableau_by_name", "dfb_tableau_generic_order_2", "dfb_tableau_generic_order_3",
    "dfb_problem_default", "dfb_system_info_get", "dfb_register_plugin",
    "dfb_register_system_from_source", "dfb_nvrtc_stats", "dfb_nvrtc_prebuild",
    "dfb_create", "dfb_destroy",
    "dfb_set_initial", "dfb_set_params", "dfb_set_initial_device", "dfb_set_params_device",
    "dfb_run", "dfb_synchronize",
    "dfb_get_final", "dfb_get_stats", "dfb_get_totals", "dfb_final_device",
    "dfb_get_trace", "dfb_get_events", "dfb_gather_final",
    "dfb_measure_fp64_peak", "dfb_last_error", "dfb_abi_version",
]
