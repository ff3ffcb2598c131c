/*
 * diffurch_b200.h — C ABI of the B200-native ensemble integrator.
 *
 * Drop-in boundary for the hot path of Danila-Bain/diffurch: the explicit
 * Runge–Kutta step loop (reference `src/solver.rs:263-314`) run over many
 * independent trajectories.  The reference has NO FFI surface of its own
 * (SURVEY.md §8b): its only seam is the Rust `Solver` builder.  This header is
 * the C-ABI a Rust `-sys` crate (or ctypes) binds; each item
 * cites the reference interface it replaces.
 *
 * Conventions
 *   - every function returns 0 (DFB_OK) or a negative dfb_status; nothing
 *     throws or aborts across the boundary; `dfb_last_error()` gives text.
 *   - the reference's `panic!`s (history lookup outside the stored range,
 *     `state.rs:166-178`; missing derivative `initial_condition.rs:33`) become
 *     per-trajectory status bits (DFB_TRAJ_*), never a process abort.
 *   - caller owns every pointer it passes; the library owns device memory
 *     behind the opaque handle.  One handle <-> one CUDA device <-> one stream;
 *     a handle is not thread-safe (the reference `Solver::run(self)` is
 *     single-threaded and consumes the solver).
 *   - host arrays are array-of-states `[n_traj][N]` (what a caller holding a
 *     `Vec<P>` has); device arrays are structure-of-arrays `[N][n_traj]`.
 *   - there is no CPU fallback: without a CUDA device `dfb_create` fails with
 *     DFB_ERR_NO_DEVICE.
 */
#ifndef DIFFURCH_B200_H
#define DIFFURCH_B200_H

#ifdef __CUDACC_RTC__
/* run-time compilation (NVRTC) has no system headers: the LP64 types this header uses */
typedef unsigned long size_t;
typedef unsigned long uint64_t;
typedef long int64_t;
typedef unsigned int uint32_t;
typedef int int32_t;
#else
#include <stddef.h>
#include <stdint.h>
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define DFB_ABI_VERSION 3

#if defined(__GNUC__)
#define DFB_API __attribute__((visibility("default")))
#else
#define DFB_API
#endif

#define DFB_MAX_STAGES 7   /* rktp64: S = 7  (rk.rs:313) */
#define DFB_MAX_INTERP 5   /* rktp64: I = 5 */
#define DFB_MAX_STATE 12   /* Lorenz + 3x3 variation (tests/lyapunov_exponents.rs:256) */
#define DFB_MAX_PARAMS 12
#define DFB_MAX_AUX 4
#define DFB_MAX_LOCATORS 4
#define DFB_MAX_DISCO 8

typedef enum dfb_status {
  DFB_OK = 0,
  DFB_ERR_INVALID = -1,      /* bad descriptor field / argument */
  DFB_ERR_UNSUPPORTED = -2,  /* combination has no compiled kernel */
  DFB_ERR_NO_DEVICE = -3,    /* no CUDA device: the product has no CPU path */
  DFB_ERR_CUDA = -4,         /* CUDA runtime error, see dfb_last_error() */
  DFB_ERR_STATE = -5,        /* call order (e.g. get_final before run) */
  DFB_ERR_NCCL = -6
} dfb_status;

/* Per-trajectory status bits (dfb_stats.status). */
#define DFB_TRAJ_OK 0u
#define DFB_TRAJ_HISTORY_DELETED 1u    /* state.rs:166-171 "deleted time range" */
#define DFB_TRAJ_HISTORY_NOT_READY 2u  /* state.rs:172-177 "not yet computed" (delay <= h) */
#define DFB_TRAJ_NO_DERIVATIVE 4u      /* initial_condition.rs:33 unimplemented D>=1 */
#define DFB_TRAJ_NONFINITE 8u          /* state became NaN/Inf (reference would keep going) */
#define DFB_TRAJ_STEP_LIMIT 16u        /* max_steps guard hit */
#define DFB_TRAJ_RING_OVERFLOW 32u     /* history ring capacity exceeded (library limit) */
#define DFB_TRAJ_STOPPED 64u           /* StateRefMut::stop_integration (state_fn.rs:92-97) */
#define DFB_TRAJ_DISCO_OVERFLOW 128u   /* discontinuity list capacity exceeded */
#define DFB_TRAJ_STEPSIZE_FLOOR 256u   /* AutomaticStepsize rejected a step and could not shrink its
                                          stepsize (clamped at stepsize_range.start): the reference's
                                          retry loop (solver.rs:294-297) never ends here; the trajectory
                                          is stopped with t = NaN */

/* ---------------------------------------------------------------- tableau */
/* Mirror of `ButcherTableu<T,S,I>` (rk.rs:5-14).  Row-major, padded to the
 * maxima; entries outside [S]x[S] / [S]x[I] are ignored. */
typedef struct dfb_tableau {
  int32_t S;
  int32_t I;
  int32_t order;
  int32_t order_embedded;
  int32_t order_interpolant;
  int32_t _pad;
  double a[DFB_MAX_STAGES * DFB_MAX_STAGES]; /* a[i*DFB_MAX_STAGES + j] */
  double b[DFB_MAX_STAGES];
  double b2[DFB_MAX_STAGES];
  double c[DFB_MAX_STAGES];
  double bi[DFB_MAX_STAGES * DFB_MAX_INTERP]; /* bi[i*DFB_MAX_INTERP + j] */
} dfb_tableau;

/* Named constructors of rk.rs:55-430: "euler", "midpoint", "heun2",
 * "ralston2", "kutta3", "heun3", "ralston3", "wray3", "ssp3", "rk4",
 * "three_eights", "rk43", "rktp64". */
DFB_API int dfb_tableau_by_name(const char* name, dfb_tableau* out);
/* rk.rs:79 generic_order_2(c2) */
DFB_API int dfb_tableau_generic_order_2(double c2, dfb_tableau* out);
/* rk.rs:116 generic_order_3(c2, c3, b3: Option<f64>); has_b3 = 0 -> None.
 * Arguments the reference panics on (rk.rs:200) return DFB_ERR_INVALID. */
DFB_API int dfb_tableau_generic_order_3(double c2, double c3, int has_b3, double b3,
                                dfb_tableau* out);

/* ---------------------------------------------------------------- systems */
/* Registered device functors for `Solver::equation` (solver.rs:135).  A host
 * closure cannot run on the GPU, so the shipped example/test systems of the
 * reference are compiled in; params are what the reference closure captures. */
typedef enum dfb_system_id {
  /* examples/ode-chaos-lorenz.rs:19-22. N=3, params {sigma, rho, beta} */
  DFB_SYS_LORENZ = 1,
  /* tests/lyapunov_exponents.rs:264-283. N=12 column-major 3x4
   * [x y z | dv1 | dv2 | dv3], params {sigma, rho, beta}; aux = 3 log sums.
   * callback 1 = QR re-orthonormalisation (lyapunov_exponents.rs:289-296). */
  DFB_SYS_LORENZ_VARIATIONAL = 2,
  /* tests/lyapunov_exponents.rs:17 / equation_types.rs:76. N=1, x' = x*lambda,
   * params {lambda}; aux = 1 log sum; callback 1 = renormalise (:26-30). */
  DFB_SYS_LINEAR1 = 3,
  /* tests/lyapunov_exponents.rs:55. x' = (1+sin t)*lambda*x; same aux/callback */
  DFB_SYS_LINEAR1_SIN = 4,
  /* examples/ode-linear-3-fibonacci.rs:11-20. N=3, y' = A y, params = A row-major (9) */
  DFB_SYS_LINEAR3 = 5,
  /* tests/lyapunov_exponents.rs:107-131. N=9 column-major 3x3 Y, Y' = A Y,
   * params = A row-major (9); aux = 3 log sums; callback 1 = QR with ln(r)
   * of the signed diagonal (:127), callback 2 = QR with ln|r| (:206). */
  DFB_SYS_LINEAR3_MATRIX = 6,
  /* tests/equation_types.rs:89-92. N=2, [dx, -x] */
  DFB_SYS_HARMONIC = 7,
  /* tests/equation_types.rs:107. N=1, y' = -2 y / t */
  DFB_SYS_POWER_DECAY = 8,
  /* tests/equation_types.rs:14. N=3, y' = [1,-1,0] */
  DFB_SYS_CONSTANT3 = 9,
  /* tests/equation_types.rs:31. N=3, y' = [0,1,t] */
  DFB_SYS_TIME3 = 10,
  /* tests/delay_propagation.rs:13. N=1, y' = 0.
   * detector 1 = delayed argument t/2 (pantograph, :86). */
  DFB_SYS_ZERO1 = 11,
  /* tests/equation_types.rs:116-136. N=1, x' = a x + b x(t - tau),
   * params {a, b, tau, k}; history DFB_HIST_SIN uses k. */
  DFB_SYS_DDE_LINEAR = 12,
  /* tests/equation_types.rs:139-158. x' = a x(t-tau) + b x'(t-tau), params {a,b,tau,k} */
  DFB_SYS_NDDE_LINEAR = 13,
  /* examples/bouncing-ball.rs:21-33. N=2 {x, dx}, params {g, k};
   * detector 1 = dx, detector 2 = x; callback 1 = bounce (:28-31) */
  DFB_SYS_BOUNCING_BALL = 14,
  /* examples/egg-billiard.rs:12-40. N=4, params {max_reflections};
   * detector 1 = boundary function (:18-21); callback 1 = reflect (:22-38);
   * aux[0] = reflection counter */
  DFB_SYS_EGG_BILLIARD = 15,
  /* examples/ode-2-relay-msign.rs:17. N=2, [y, -2 sign(p_prev.x)]; detector 1 = x */
  DFB_SYS_RELAY_MSIGN = 16,
  /* examples/dde-relay-clamp.rs:24-35. N=2, params {alpha, sigma, T};
   * detector 1 = |x(t-T)| - 1 */
  DFB_SYS_DDE_RELAY_CLAMP = 17,
  /* ids >= DFB_SYS_USER_BASE are user systems registered with dfb_register_plugin */
  DFB_SYS_USER_BASE = 1000
} dfb_system_id;

/* `Solver::initial` argument kinds (initial_condition.rs:15-50). */
typedef enum dfb_history_id {
  DFB_HIST_CONSTANT = 0, /* any `U: Copy` constant: y(t)=y0, y'(t)=0 (:15-22) */
  DFB_HIST_SIN = 1,      /* InitFn(|t| sin(k t), |t| k cos(k t)) (equation_types.rs:147) */
  DFB_HIST_SIN_NODERIV = 2, /* InitFn(sol, ()) (equation_types.rs:127): D>=1 unimplemented */
  DFB_HIST_INV_SQUARE = 3   /* InitFn(|t| t^-2, ()) (equation_types.rs:106) */
} dfb_history_id;

/* ------------------------------------------------------------------ events */
typedef enum dfb_locator_kind {
  DFB_LOC_STATEFN = 0,     /* LocatorStateFn over a registered detector (loc.rs:36-80) */
  DFB_LOC_PERIODIC = 1,    /* Periodic{period, offset} (loc.rs:292-335) */
  DFB_LOC_PROPAGATION = 2  /* propagated_discontinuity (loc.rs:942-967) */
} dfb_locator_kind;

/* detection_method (loc.rs:127-143) */
typedef enum dfb_detection {
  DFB_DET_ZERO = 0, DFB_DET_ABOVE_ZERO = 1, DFB_DET_BELOW_ZERO = 2,
  DFB_DET_POSITIVE = 3, DFB_DET_NEGATIVE = 4,
  DFB_DET_SWITCH = 5, DFB_DET_SWITCH_TRUE = 6, DFB_DET_SWITCH_FALSE = 7,
  DFB_DET_IS_TRUE = 8, DFB_DET_IS_FALSE = 9, DFB_DET_STEP = 10
} dfb_detection;

/* location_method (loc.rs:200-285) */
typedef enum dfb_location {
  DFB_LOCATE_STEP_BEGIN = 0, DFB_LOCATE_STEP_END = 1, DFB_LOCATE_STEP_MIDDLE = 2,
  DFB_LOCATE_LERP = 3, DFB_LOCATE_BISECTION = 4, DFB_LOCATE_BISECTION_BOOL = 5,
  DFB_LOCATE_REGULA_FALSI = 6
} dfb_location;

/* Filter::{every, separated_by, in_interval, on_times} (loc.rs:613-688).
 * on_times takes the first filter_n entries of a strictly increasing
 * index list stored in filter_times. */
typedef enum dfb_filter {
  DFB_FILTER_NONE = 0, DFB_FILTER_EVERY = 1, DFB_FILTER_SEPARATED_BY = 2,
  DFB_FILTER_IN_INTERVAL = 3, DFB_FILTER_ON_TIMES = 4
} dfb_filter;
#define DFB_MAX_FILTER_TIMES 8 /* on_times: a finite index list (the reference takes any iterator) */

#define DFB_CB_NONE 0 /* `|_| {}` */

typedef struct dfb_locator {
  int32_t kind;            /* dfb_locator_kind */
  int32_t detector_id;     /* system-local id (STATEFN); 0 = const delay `t - delay`
                              for PROPAGATION (solver.rs:260), else system-local
                              delayed-argument id (solver.rs:239) */
  int32_t detection;       /* dfb_detection (STATEFN only) */
  int32_t location;        /* dfb_location  (STATEFN only; PROPAGATION is Bisection) */
  int32_t callback_id;     /* system-local id, DFB_CB_NONE = no-op */
  int32_t dedup;           /* 1 = wrapped in DedupLocF as `on`/`on_mut` do (solver.rs:210,230) */
  int32_t smoothing_order; /* PROPAGATION (loc.rs:350) */
  int32_t filter;          /* dfb_filter */
  int32_t filter_n;        /* every(n) */
  int32_t _pad;
  double period;           /* PERIODIC */
  double offset;           /* PERIODIC */
  double delay;            /* PROPAGATION with detector_id == 0 */
  double filter_a;         /* separated_by(diff) / in_interval start (inclusive) */
  double filter_b;         /* in_interval end (exclusive) */
  int32_t filter_times[DFB_MAX_FILTER_TIMES]; /* on_times (loc.rs:665-688): pass the detections whose
                              0-based index is in this strictly increasing list of filter_n entries */
} dfb_locator;

/* ---------------------------------------------------------------- stepsize */
typedef enum dfb_stepsize_kind {
  DFB_STEP_FIXED = 0,    /* bare scalar controller (stepsize.rs:18-32) */
  DFB_STEP_AUTOMATIC = 1 /* AutomaticStepsize (stepsize.rs:35-85) */
} dfb_stepsize_kind;

/* Mirror of AutomaticStepsize (stepsize.rs:35-44).  `initial_stepsize` is
 * dead in the reference (init() is never called by run) and is carried only
 * for completeness. */
typedef struct dfb_auto_stepsize {
  double stepsize;
  double stepsize_min, stepsize_max; /* stepsize_range */
  double atol[DFB_MAX_STATE];
  double rtol[DFB_MAX_STATE];
  uint32_t order;
  uint32_t _pad;
  double fac;
  double fac_min, fac_max;           /* fac_range */
  double initial_stepsize;           /* unused, see above */
} dfb_auto_stepsize;

/* Arithmetic contract.  STRICT rounds every multiply and add separately in
 * the reference's operation order (Rust never contracts to FMA): bit-identical
 * to the CPU path for systems without transcendental calls.  FAST lets the
 * kernel use FMA, pre-scaled tableau rows and pre-combined dense-output
 * coefficients: within 1e-12 relative of STRICT on non-chaotic horizons. */
typedef enum dfb_arith { DFB_ARITH_STRICT = 0, DFB_ARITH_FAST = 1 } dfb_arith;

/* ----------------------------------------------------------------- problem */
/* POD mirror of the fields of `Solver` (solver.rs:65-76) + the ensemble axis. */
typedef struct dfb_problem {
  int32_t abi_version;  /* DFB_ABI_VERSION */
  int32_t system_id;    /* dfb_system_id          <- Solver::equation */
  int32_t history_id;   /* dfb_history_id         <- Solver::initial */
  int32_t arith;        /* dfb_arith */
  dfb_tableau rk;       /*                        <- Solver::rk (default rktp64, solver.rs:88) */
  double t0, t1;        /*                        <- Solver::interval; t1 may be +inf */
  int32_t stepsize_kind; /* dfb_stepsize_kind     <- Solver::stepsize */
  int32_t n_loc;        /*                        <- Solver::on / on_mut / with_const_delay ... in call order */
  double h;             /* fixed stepsize (default 0.05, solver.rs:89) */
  dfb_auto_stepsize automatic;
  double max_delay;     /*                        <- Solver::max_delay (solver.rs:130) */
  int32_t n_disco;      /*                        <- Solver::initial_disco (solver.rs:114) */
  int32_t on_start_callback_id; /*               <- Solver::on_start_mut (solver.rs:190-197): system-local
                                   callback id run once on the initial State, before the first
                                   on_step; DFB_CB_NONE = none.  (Read-only on_start / on_stop closures,
                                   solver.rs:172-188, observe what the caller already holds: the initial
                                   state and the final State that dfb_get_final returns.) */
  double disco_t[DFB_MAX_DISCO];
  int32_t disco_order[DFB_MAX_DISCO];
  dfb_locator loc[DFB_MAX_LOCATORS];
  /* Output policy (replaces host `on_step` closures, which cannot run on the
   * device).  trace_every = k > 0 records (t, p) at every k-th `on_step`
   * invocation (the one at t_init is invocation 0, solver.rs:290) into a
   * [trace_capacity][1+N][n_traj] device buffer. */
  int32_t trace_every;
  int32_t trace_capacity;
  int32_t event_capacity; /* per-trajectory log of (time, locator index) of fired callbacks */
  int32_t ring_capacity;  /* history ring slots per trajectory; 0 = derive from max_delay/h */
  int64_t max_steps;      /* guard for unbounded intervals; 0 = no limit */
} dfb_problem;

/* Fills the defaults of `Solver::new` (solver.rs:80-96): rktp64, h = 0.05,
 * max_delay = 0, no events, t0 = 0, t1 = +inf, FAST arithmetic off (STRICT). */
DFB_API int dfb_problem_default(dfb_problem* p);

/* Static facts about a registered system. */
typedef struct dfb_system_info {
  int32_t n_state, n_params, n_aux, n_detectors, n_callbacks, uses_history;
} dfb_system_info;
DFB_API int dfb_system_info_get(int32_t system_id, dfb_system_info* out);

/* User-defined systems.  `Solver::equation` (solver.rs:135) takes ANY closure; a closure cannot
 * cross this ABI, so its CUDA restatement — a struct with the functor signatures of
 * diffurch_b200/csrc/dev/systems.cuh (rhs / detector / callback over StateRef / StateRefMut,
 * reference state_fn.rs:13-98) — is compiled against the same kernel templates into a plugin
 * shared object (python: diffurch_b200.user.build_user_system; or nvcc by hand)
 * and registered here.  Returns the system id to put into dfb_problem.system_id.  Registering
 * the same path twice returns the same id.  Plugins stay loaded for the life of the process. */
DFB_API int dfb_register_plugin(const char* so_path, int32_t* system_id_out);

/* The same from SOURCE text: the functor is compiled at run time with NVRTC for sm_100a (no nvcc on the
 * machine), one kernel instantiation at a time, on first use — the exact tableau sparsity, arithmetic
 * contract and trace policy a problem asks for; CUBINs are cached in memory and on disk (DFB_NVRTC_CACHE).
 * `cuda_source` is what a user-system header holds; it may omit the
 * `#include "user_system.cuh"` / `namespace DFB_NS {}` wrapper.  Fixed-step ODE problems without events run in
 * ode_fixed_kernel, everything else in integrate_general_kernel — the same templates as the built-in
 * systems, so results are bit-identical to the nvcc-built plugin of the same functor.  Registration needs no
 * GPU (the functor's static facts are read from NVRTC's output); running does. */
DFB_API int dfb_register_system_from_source(const char* cuda_source, const char* struct_name,
                                            int32_t* system_id_out);
/* Run-time compilation counters of this process: kernels compiled, disk-cache hits, total and last compile
 * time in milliseconds.  NULL pointers are skipped. */
DFB_API int dfb_nvrtc_stats(uint32_t* kernels_compiled, uint32_t* disk_cache_hits, double* compile_ms_total,
                            double* last_compile_ms);
/* Compile (or fetch from the cache) the ode_fixed_kernel instantiation of a source-registered system for
 * `rk` without launching it: pays the compile time up front; works without a GPU. */
DFB_API int dfb_nvrtc_prebuild(int32_t system_id, int32_t arith, const dfb_tableau* rk,
                               int32_t members_per_thread, int32_t trace);

/* ------------------------------------------------------------------ handle */
typedef struct dfb_handle dfb_handle;

/* Per-trajectory counters, all [n_traj]. */
typedef struct dfb_stats {
  uint64_t* steps;     /* committed steps = on_step invocations after t_init */
  uint64_t* rejected;  /* AutomaticStepsize rejections (solver.rs:294-297) */
  uint64_t* events;    /* located events whose callback fired (solver.rs:299-307) */
  uint64_t* rhs_evals; /* equation evaluations */
  uint32_t* status;    /* DFB_TRAJ_* bits */
} dfb_stats;

/* Aggregate of the last run on this handle (this rank only). */
typedef struct dfb_totals {
  uint64_t n_traj;
  uint64_t steps, rejected, events, rhs_evals;
  uint64_t n_failed;       /* trajectories with any status bit except STOPPED */
  double kernel_ms;        /* CUDA-event duration of the integration kernel(s) */
  uint32_t kernel_launches;
  uint32_t kernel_kind;    /* dfb_kernel_kind: which kernel family integrated the ensemble */
  uint64_t locate_iters;   /* iterations of the location loops (Bisection / BisectionBool / RegulaFalsi,
                              loc.rs:210-285) summed over the ensemble; counted by the event kernels
                              (DFB_KERNEL_ODE_EVENT, DFB_KERNEL_DDE_EVENT), 0 elsewhere */
} dfb_totals;

/* Kernel families (all CUDA; there is no CPU path). */
typedef enum dfb_kernel_kind {
  DFB_KERNEL_GENERAL = 0,      /* integrate_general_kernel: every Solver feature */
  DFB_KERNEL_ODE_FIXED = 1,    /* ode_fixed_kernel: fixed step, registers only */
  DFB_KERNEL_DDE_FIXED = 2,    /* dde_fixed_kernel: constant delay, coefficient ring */
  DFB_KERNEL_ODE_ADAPTIVE = 3, /* ode_adaptive_kernel: AutomaticStepsize, dynamic trajectory queue */
  DFB_KERNEL_ODE_FIXED_PERIODIC = 4, /* ode_fixed_periodic_kernel: fixed step + Periodic callbacks
                                        (warp-uniform event handling; the Lyapunov-batch shape) */
  DFB_KERNEL_ODE_EVENT = 5,         /* ode_event_kernel: fixed step + located state-function events
                                       (register-resident locators, warp-batched bisection) */
  DFB_KERNEL_DDE_EVENT = 6          /* dde_event_kernel: fixed step DDE/NDDE + located events +
                                       discontinuity propagation (per-lane coefficient ring with
                                       stored start times, cached-record lookup) */
} dfb_kernel_kind;

/* Solver::new + setters -> one handle for `n_traj` trajectories on `device`. */
DFB_API int dfb_create(const dfb_problem* desc, size_t n_traj, int device, dfb_handle** out);
DFB_API void dfb_destroy(dfb_handle* h);

/* Solver::initial: y0 is host [n_traj][N] (copied). */
DFB_API int dfb_set_initial(dfb_handle* h, const double* y0);
/* Closure captures (e.g. rho, ode-chaos-lorenz.rs:7): host [n_traj][n_params]. */
DFB_API int dfb_set_params(dfb_handle* h, const double* params);
/* Device-resident variants: SoA [N][n_traj] / [n_params][n_traj] on the handle's device. */
DFB_API int dfb_set_initial_device(dfb_handle* h, const double* d_y0);
DFB_API int dfb_set_params_device(dfb_handle* h, const double* d_params);

/* Solver::run (solver.rs:263-314) for every trajectory.  `stream` is a
 * cudaStream_t (NULL = legacy default stream).  Asynchronous; the getters
 * below synchronise. */
DFB_API int dfb_run(dfb_handle* h, void* stream);
DFB_API int dfb_synchronize(dfb_handle* h);

/* Final `State` (solver.rs:313): t_curr [n_traj], p_curr [n_traj][N], aux
 * [n_traj][n_aux] (callback accumulators).  NULL pointers are skipped. */
DFB_API int dfb_get_final(dfb_handle* h, double* t, double* y, double* aux);
DFB_API int dfb_get_stats(dfb_handle* h, dfb_stats* out);
DFB_API int dfb_get_totals(dfb_handle* h, dfb_totals* out);
/* Device pointers (SoA) of the final state, valid until the next run/destroy. */
DFB_API int dfb_final_device(dfb_handle* h, const double** d_t, const double** d_y,
                     const double** d_aux);
/* Trace: n_samples [n_traj], t [n_traj][trace_capacity], y
 * [n_traj][trace_capacity][N] (host, array-of-trajectories). */
DFB_API int dfb_get_trace(dfb_handle* h, uint32_t* n_samples, double* t, double* y);
/* Event log: n_events [n_traj] (may exceed event_capacity: count of all
 * events), t [n_traj][event_capacity], index [n_traj][event_capacity]. */
DFB_API int dfb_get_events(dfb_handle* h, uint32_t* n_events, double* t, int32_t* index);

/* Ensemble statistics across ranks: gathers the final (t, y, aux) of every
 * rank's handle into host arrays ordered by rank (ncclAllGather over
 * NVLink).  `nccl_comm` is an ncclComm_t owned by the caller; counts must
 * be equal on all ranks. */
DFB_API int dfb_gather_final(dfb_handle* h, void* nccl_comm, int n_ranks, double* t_all,
                     double* y_all, double* aux_all);

/* Measured FP64 FMA rate of the device (DFMA chain microbenchmark), TFLOP/s. */
DFB_API int dfb_measure_fp64_peak(int device, double* tflops, double* sm_mhz_est);

DFB_API const char* dfb_last_error(void);
DFB_API int dfb_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* DIFFURCH_B200_H */
