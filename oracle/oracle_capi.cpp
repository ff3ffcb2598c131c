// =============================================================================
// oracle_capi.cpp — TEST INFRASTRUCTURE ONLY (see diffurch_oracle.hpp).
//
// C entry points (ctypes) over the oracle: takes the SAME `dfb_problem`
// descriptor as the product's C ABI so that parity tests feed both sides one
// input.  Each trajectory gets its own Solver, exactly like a rayon-style loop
// over the reference's `Solver::run(self)` would (a Solver is
// consumed by run and its closures are FnMut, so nothing is shared).
// =============================================================================
#include <atomic>
#include <cstring>
#include <thread>
#include <vector>

#include "../include/diffurch_b200.h"
#include "diffurch_oracle.hpp"
#include "oracle_systems.hpp"
#include "oracle_tableaux.hpp"

namespace {
using namespace orc;

struct RunArgs {
  const dfb_problem* desc;
  size_t n_traj;
  const double* y0;      // [n][N]
  const double* params;  // [n][P] or null
  int n_params;
  int n_aux;
  double* t_final;
  double* y_final;
  double* aux_final;
  uint64_t *steps, *rejected, *events, *rhs_evals;
  uint32_t* status;
  uint32_t* n_samples;
  double *trace_t, *trace_y;
  uint32_t* n_events;
  double* ev_t;
  int32_t* ev_idx;
};

template <int S, int I>
ButcherTableu<S, I> tableau_from(const dfb_tableau& d) {
  ButcherTableu<S, I> t;
  t.order = d.order;
  t.order_embedded = d.order_embedded;
  t.order_interpolant = d.order_interpolant;
  for (int i = 0; i < S; ++i) {
    for (int j = 0; j < S; ++j) t.a[i][j] = d.a[i * DFB_MAX_STAGES + j];
    t.b[i] = d.b[i];
    t.b2[i] = d.b2[i];
    t.c[i] = d.c[i];
    for (int j = 0; j < I; ++j) t.bi[i][j] = d.bi[i * DFB_MAX_INTERP + j];
  }
  return t;
}

template <int S, int I>
void tableau_to(const ButcherTableu<S, I>& t, dfb_tableau* d) {
  std::memset(d, 0, sizeof(*d));
  d->S = S;
  d->I = I;
  d->order = t.order;
  d->order_embedded = t.order_embedded;
  d->order_interpolant = t.order_interpolant;
  for (int i = 0; i < S; ++i) {
    for (int j = 0; j < S; ++j) d->a[i * DFB_MAX_STAGES + j] = t.a[i][j];
    d->b[i] = t.b[i];
    d->b2[i] = t.b2[i];
    d->c[i] = t.c[i];
    for (int j = 0; j < I; ++j) d->bi[i * DFB_MAX_INTERP + j] = t.bi[i][j];
  }
}

template <int N>
InitialCondition<N> make_initial(const dfb_problem& d, const double* y0, const double* prm) {
  Vec<N> v;
  for (int i = 0; i < N; ++i) v[i] = y0[i];
  switch (d.history_id) {
    case DFB_HIST_CONSTANT: return InitialCondition<N>::constant_value(v);
    case DFB_HIST_SIN: {
      const double k = prm[3];
      return InitialCondition<N>::init_fn(
          [k](double t) {
            Vec<N> r = Vec<N>::zero();
            r[0] = std::sin(k * t);
            return r;
          },
          [k](double t) {
            Vec<N> r = Vec<N>::zero();
            r[0] = k * std::cos(k * t);
            return r;
          });
    }
    case DFB_HIST_SIN_NODERIV: {
      const double k = prm[3];
      return InitialCondition<N>::init_fn([k](double t) {
        Vec<N> r = Vec<N>::zero();
        r[0] = std::sin(k * t);
        return r;
      });
    }
    case DFB_HIST_INV_SQUARE:
      return InitialCondition<N>::init_fn([](double t) {
        Vec<N> r = Vec<N>::zero();
        r[0] = powi(t, -2);
        return r;
      });
  }
  return InitialCondition<N>::constant_value(v);
}

template <class Sys, int N, int S, int I>
std::unique_ptr<LocBase<N, S, I>> make_locator(const dfb_locator& L, const double* prm,
                                               double* aux) {
  using Ref = StateRef<N, S, I>;
  std::unique_ptr<LocBase<N, S, I>> base;
  if (L.kind == DFB_LOC_PERIODIC) {
    auto p = std::make_unique<Periodic<N, S, I>>();
    p->period = L.period;
    p->offset = L.offset;
    p->callback = Sys::callback(L.callback_id, prm, aux);
    base = std::move(p);
  } else if (L.kind == DFB_LOC_STATEFN) {
    auto p = std::make_unique<LocatorStateFn<N, S, I>>();
    if (L.detection == DFB_DET_STEP) p->f.f = [](const Ref&) { return 1.0; };
    else p->f.f = Sys::detector(L.detector_id, prm);
    p->detection = Detection(L.detection);
    p->location = Location(L.location);
    p->callback = Sys::callback(L.callback_id, prm, aux);
    base = std::move(p);
  }
  if (!base) return base;
  if (L.filter == DFB_FILTER_EVERY) {
    auto f = std::make_unique<FilterAfterDetection<N, S, I>>();
    f->loc = std::move(base);
    auto counter = make_every_counter(size_t(L.filter_n));
    f->filter = [counter](const Ref&) { return counter(); };
    base = std::move(f);
  } else if (L.filter == DFB_FILTER_SEPARATED_BY) {
    auto f = std::make_unique<FilterLocated<N, S, I>>();
    f->loc = std::move(base);
    auto prev = std::make_shared<double>(-std::numeric_limits<double>::max());
    const double diff = L.filter_a;
    f->filter = [prev, diff](const Ref& s) {  // loc.rs:638-648
      if (s.t >= *prev + diff) {
        *prev = s.t;
        return true;
      }
      return false;
    };
    base = std::move(f);
  } else if (L.filter == DFB_FILTER_IN_INTERVAL) {
    auto f = std::make_unique<FilterLocated<N, S, I>>();
    f->loc = std::move(base);
    const double lo = L.filter_a, hi = L.filter_b;
    f->filter = [lo, hi](const Ref& s) { return s.t >= lo && s.t < hi; };  // `lo..hi`
    base = std::move(f);
  } else if (L.filter == DFB_FILTER_ON_TIMES) {  // loc.rs:665-688
    auto f = std::make_unique<FilterAfterDetection<N, S, I>>();
    f->loc = std::move(base);
    std::vector<size_t> times(L.filter_times, L.filter_times + L.filter_n);
    auto i = std::make_shared<size_t>(0);
    auto next = std::make_shared<size_t>(0);  // the Peekable iterator's position
    f->filter = [times, i, next](const Ref&) {
      if (*next < times.size() && *i == times[*next]) {
        *next += 1;
        *i += 1;
        return true;
      }
      *i += 1;
      return false;
    };
    base = std::move(f);
  }
  return base;
}

template <class Sys, int N, int S, int I>
void run_one(const RunArgs& A, size_t i) {
  const dfb_problem& d = *A.desc;
  const double* prm = A.params ? A.params + i * A.n_params : nullptr;
  double aux[DFB_MAX_AUX] = {0, 0, 0, 0};
  Solver<N, S, I, decltype(Sys::rhs(prm))> solver(Sys::rhs(prm), tableau_from<S, I>(d.rk));
  solver.initial = make_initial<N>(d, A.y0 + i * N, prm);
  for (int k = 0; k < d.n_disco; ++k)
    solver.initial_disco.push_back(Disco{d.disco_t[k], size_t(d.disco_order[k])});
  solver.t_init = d.t0;
  solver.t_end = d.t1;
  solver.max_delay = d.max_delay;
  solver.max_steps = uint64_t(d.max_steps);
  if (d.stepsize_kind == DFB_STEP_AUTOMATIC) {
    const dfb_auto_stepsize& a = d.automatic;
    solver.stepsize.automatic = true;
    solver.stepsize.stepsize = a.stepsize;
    solver.stepsize.range_lo = a.stepsize_min;
    solver.stepsize.range_hi = a.stepsize_max;
    for (int k = 0; k < N; ++k) {
      solver.stepsize.atol[k] = a.atol[k];
      solver.stepsize.rtol[k] = a.rtol[k];
    }
    solver.stepsize.order = a.order;
    solver.stepsize.fac = a.fac;
    solver.stepsize.fac_lo = a.fac_min;
    solver.stepsize.fac_hi = a.fac_max;
  } else {
    solver.stepsize.stepsize = d.h;
  }
  for (int k = 0; k < d.n_loc; ++k) {
    const dfb_locator& L = d.loc[k];
    if (L.kind == DFB_LOC_PROPAGATION) {
      const double max_delay_keep = solver.max_delay;
      if (L.detector_id == 0) solver.with_const_delay(L.delay, size_t(L.smoothing_order));
      else solver.with_delayed_argument(Sys::detector(L.detector_id, prm), size_t(L.smoothing_order));
      solver.max_delay = max_delay_keep;  // the descriptor already holds the builder's final max_delay
    } else {
      auto loc = make_locator<Sys, N, S, I>(L, prm, aux);
      if (L.dedup) solver.on(std::move(loc));
      else solver.events_on_loc.push_back(std::move(loc));
    }
  }
  // output policy = an on_step closure
  uint32_t n_samples = 0;
  uint64_t on_step_calls = 0;
  const int every = d.trace_every, cap = d.trace_capacity;
  if (every > 0 && cap > 0 && A.trace_t) {
    double* tt = A.trace_t + i * size_t(cap);
    double* ty = A.trace_y + i * size_t(cap) * N;
    solver.events_on_step.push_back([&, tt, ty](const StateRef<N, S, I>& s) {
      if (on_step_calls % uint64_t(every) == 0 && n_samples < uint32_t(cap)) {
        tt[n_samples] = s.t;
        for (int k = 0; k < N; ++k) ty[size_t(n_samples) * N + k] = s.p[k];
        ++n_samples;
      }
      ++on_step_calls;
    });
  }
  solver.log_events = A.ev_t != nullptr && d.event_capacity > 0;
  if (d.on_start_callback_id != DFB_CB_NONE) solver.on_start_mut = Sys::callback(d.on_start_callback_id, prm, aux);

  uint32_t status = TRAJ_OK;
  double t_fin = std::numeric_limits<double>::quiet_NaN();
  Vec<N> y_fin;
  for (int k = 0; k < N; ++k) y_fin[k] = t_fin;
  uint64_t rhs_evals = 0;
  try {
    auto st = solver.run();
    t_fin = st.t_curr;
    y_fin = st.p_curr;
    rhs_evals = st.rhs_evals;
    if (std::isinf(t_fin)) status |= TRAJ_STOPPED;  // stop_integration (state_fn.rs:92-97)
    // library status, not a reference behaviour: the reference returns the non-finite state silently
    for (int k = 0; k < N; ++k)
      if (!(std::fabs(y_fin[k]) < std::numeric_limits<double>::infinity())) status |= TRAJ_NONFINITE;
  } catch (const Panic& p) {
    status |= p.status;
  }
  if (A.t_final) A.t_final[i] = t_fin;
  if (A.y_final)
    for (int k = 0; k < N; ++k) A.y_final[i * N + k] = y_fin[k];
  if (A.aux_final)
    for (int k = 0; k < A.n_aux; ++k) A.aux_final[i * A.n_aux + k] = aux[k];
  if (A.steps) A.steps[i] = solver.n_steps;
  if (A.rejected) A.rejected[i] = solver.n_rejected;
  if (A.events) A.events[i] = solver.n_events;
  if (A.rhs_evals) A.rhs_evals[i] = rhs_evals;
  if (A.status) A.status[i] = status;
  if (A.n_samples) A.n_samples[i] = n_samples;
  if (solver.log_events) {
    const size_t ecap = size_t(d.event_capacity);
    if (A.n_events) A.n_events[i] = uint32_t(solver.event_log.size());
    for (size_t k = 0; k < solver.event_log.size() && k < ecap; ++k) {
      A.ev_t[i * ecap + k] = solver.event_log[k].first;
      if (A.ev_idx) A.ev_idx[i * ecap + k] = solver.event_log[k].second;
    }
  }
}

template <class Sys, int N, int S, int I>
int run_all(const RunArgs& A, int n_threads) {
  if (n_threads <= 1) {
    for (size_t i = 0; i < A.n_traj; ++i) run_one<Sys, N, S, I>(A, i);
    return 0;
  }
  std::atomic<size_t> next{0};
  const size_t chunk = 64;
  std::vector<std::thread> pool;
  for (int w = 0; w < n_threads; ++w)
    pool.emplace_back([&]() {
      for (;;) {
        size_t b = next.fetch_add(chunk);
        if (b >= A.n_traj) break;
        size_t e = b + chunk < A.n_traj ? b + chunk : A.n_traj;
        for (size_t i = b; i < e; ++i) run_one<Sys, N, S, I>(A, i);
      }
    });
  for (auto& t : pool) t.join();
  return 0;
}

template <template <int, int> class SysT, int N>
int dispatch_shape(RunArgs& A, int n_threads) {
  const int S = A.desc->rk.S, I = A.desc->rk.I;
#define ORC_SHAPE(S_, I_) \
  if (S == S_ && I == I_) return run_all<SysT<S_, I_>, N, S_, I_>(A, n_threads);
  ORC_SHAPE(1, 2)
  ORC_SHAPE(2, 2)
  ORC_SHAPE(3, 2)
  ORC_SHAPE(4, 2)
  ORC_SHAPE(5, 4)
  ORC_SHAPE(7, 5)
#undef ORC_SHAPE
  return DFB_ERR_UNSUPPORTED;
}

struct SysInfo {
  int n_state, n_params, n_aux;
};
bool sys_info(int id, SysInfo* s) {
  switch (id) {
    case DFB_SYS_LORENZ: *s = {3, 3, 0}; return true;
    case DFB_SYS_LORENZ_VARIATIONAL: *s = {12, 3, 3}; return true;
    case DFB_SYS_LINEAR1: *s = {1, 1, 1}; return true;
    case DFB_SYS_LINEAR1_SIN: *s = {1, 1, 1}; return true;
    case DFB_SYS_LINEAR3: *s = {3, 9, 0}; return true;
    case DFB_SYS_LINEAR3_MATRIX: *s = {9, 9, 3}; return true;
    case DFB_SYS_HARMONIC: *s = {2, 0, 0}; return true;
    case DFB_SYS_POWER_DECAY: *s = {1, 0, 0}; return true;
    case DFB_SYS_CONSTANT3: *s = {3, 0, 0}; return true;
    case DFB_SYS_TIME3: *s = {3, 0, 0}; return true;
    case DFB_SYS_ZERO1: *s = {1, 0, 0}; return true;
    case DFB_SYS_DDE_LINEAR: *s = {1, 4, 0}; return true;
    case DFB_SYS_NDDE_LINEAR: *s = {1, 4, 0}; return true;
    case DFB_SYS_BOUNCING_BALL: *s = {2, 2, 0}; return true;
    case DFB_SYS_EGG_BILLIARD: *s = {4, 1, 1}; return true;
    case DFB_SYS_RELAY_MSIGN: *s = {2, 0, 0}; return true;
    case DFB_SYS_DDE_RELAY_CLAMP: *s = {2, 3, 0}; return true;
  }
  return false;
}
}  // namespace

extern "C" {

int orc_system_info(int system_id, int* n_state, int* n_params, int* n_aux) {
  SysInfo s;
  if (!sys_info(system_id, &s)) return DFB_ERR_INVALID;
  *n_state = s.n_state;
  *n_params = s.n_params;
  *n_aux = s.n_aux;
  return 0;
}

int orc_run(const dfb_problem* desc, size_t n_traj, const double* y0, const double* params,
            int n_threads, double* t_final, double* y_final, double* aux_final, uint64_t* steps,
            uint64_t* rejected, uint64_t* events, uint64_t* rhs_evals, uint32_t* status,
            uint32_t* n_samples, double* trace_t, double* trace_y, uint32_t* n_events,
            double* ev_t, int32_t* ev_idx) {
  SysInfo s;
  if (!desc || !sys_info(desc->system_id, &s)) return DFB_ERR_INVALID;
  if (!y0 || (s.n_params > 0 && !params)) return DFB_ERR_INVALID;
  RunArgs A{desc,   n_traj,   y0,     params,    s.n_params, s.n_aux, t_final,
            y_final, aux_final, steps, rejected, events,     rhs_evals, status,
            n_samples, trace_t, trace_y, n_events, ev_t,     ev_idx};
  switch (desc->system_id) {
    case DFB_SYS_LORENZ: return dispatch_shape<Lorenz, 3>(A, n_threads);
    case DFB_SYS_LORENZ_VARIATIONAL: return dispatch_shape<LorenzVariational, 12>(A, n_threads);
    case DFB_SYS_LINEAR1: return dispatch_shape<Linear1, 1>(A, n_threads);
    case DFB_SYS_LINEAR1_SIN: return dispatch_shape<Linear1Sin, 1>(A, n_threads);
    case DFB_SYS_LINEAR3: return dispatch_shape<Linear3, 3>(A, n_threads);
    case DFB_SYS_LINEAR3_MATRIX: return dispatch_shape<Linear3Matrix, 9>(A, n_threads);
    case DFB_SYS_HARMONIC: return dispatch_shape<Harmonic, 2>(A, n_threads);
    case DFB_SYS_POWER_DECAY: return dispatch_shape<PowerDecay, 1>(A, n_threads);
    case DFB_SYS_CONSTANT3: return dispatch_shape<Constant3, 3>(A, n_threads);
    case DFB_SYS_TIME3: return dispatch_shape<Time3, 3>(A, n_threads);
    case DFB_SYS_ZERO1: return dispatch_shape<Zero1, 1>(A, n_threads);
    case DFB_SYS_DDE_LINEAR: return dispatch_shape<DdeLinear, 1>(A, n_threads);
    case DFB_SYS_NDDE_LINEAR: return dispatch_shape<NddeLinear, 1>(A, n_threads);
    case DFB_SYS_BOUNCING_BALL: return dispatch_shape<BouncingBall, 2>(A, n_threads);
    case DFB_SYS_EGG_BILLIARD: return dispatch_shape<EggBilliard, 4>(A, n_threads);
    case DFB_SYS_RELAY_MSIGN: return dispatch_shape<RelayMsign, 2>(A, n_threads);
    case DFB_SYS_DDE_RELAY_CLAMP: return dispatch_shape<DdeRelayClamp, 2>(A, n_threads);
  }
  return DFB_ERR_UNSUPPORTED;
}

int orc_tableau_by_name(const char* name, dfb_tableau* out) {
  if (!name || !out) return DFB_ERR_INVALID;
#define ORC_TAB(n)            \
  if (!std::strcmp(name, #n)) { \
    tableau_to(n(), out);     \
    return 0;                 \
  }
  ORC_TAB(euler)
  ORC_TAB(midpoint)
  ORC_TAB(heun2)
  ORC_TAB(ralston2)
  ORC_TAB(kutta3)
  ORC_TAB(heun3)
  ORC_TAB(ralston3)
  ORC_TAB(wray3)
  ORC_TAB(ssp3)
  ORC_TAB(rk4)
  ORC_TAB(three_eights)
  ORC_TAB(rk43)
  ORC_TAB(rktp64)
#undef ORC_TAB
  return DFB_ERR_INVALID;
}
int orc_tableau_generic_order_2(double c2, dfb_tableau* out) {
  tableau_to(generic_order_2(c2), out);
  return 0;
}
int orc_tableau_generic_order_3(double c2, double c3, int has_b3, double b3, dfb_tableau* out) {
  ButcherTableu<3, 2> t;
  if (!generic_order_3(c2, c3, has_b3 != 0, b3, &t)) return DFB_ERR_INVALID;
  tableau_to(t, out);
  return 0;
}

// dense_output::<D> of a tableau on scalar state (rk.rs:17-52), for unit tests
double orc_dense_output(const dfb_tableau* tab, int D, double y_prev, double t_step,
                        double theta, const double* k) {
  if (tab->S == 7 && tab->I == 5) {
    auto t = tableau_from<7, 5>(*tab);
    Vec<1> kk[7];
    for (int i = 0; i < 7; ++i) kk[i][0] = k[i];
    Vec<1> y{{y_prev}};
    return D == 0 ? t.dense_output<0, 1>(y, t_step, theta, kk)[0]
                  : t.dense_output<1, 1>(y, t_step, theta, kk)[0];
  }
  if (tab->S == 5 && tab->I == 4) {
    auto t = tableau_from<5, 4>(*tab);
    Vec<1> kk[5];
    for (int i = 0; i < 5; ++i) kk[i][0] = k[i];
    Vec<1> y{{y_prev}};
    return D == 0 ? t.dense_output<0, 1>(y, t_step, theta, kk)[0]
                  : t.dense_output<1, 1>(y, t_step, theta, kk)[0];
  }
  return std::numeric_limits<double>::quiet_NaN();
}

// src/util/partition_point.rs: predicate `x < threshold` over an int deque;
// returns the index and writes the number of predicate calls.
size_t orc_partition_point_linear(const int64_t* data, size_t n, size_t start,
                                  int64_t threshold, uint64_t* pred_calls) {
  RingDeque<int64_t> dq;
  for (size_t i = 0; i < n; ++i) dq.push_back(data[i]);
  uint64_t calls = 0;
  size_t r = partition_point_linear(dq, start, [&](int64_t x) {
    ++calls;
    return x < threshold;
  });
  if (pred_calls) *pred_calls = calls;
  return r;
}

// 3x3 Householder QR (column-major), for unit tests of the Lyapunov callbacks
void orc_householder_qr3(const double* a, double* q, double* rdiag) {
  householder_qr3(a, q, rdiag);
}

double orc_powi(double a, int b) { return powi(a, b); }

// the n-th root used instead of libm pow by the STRICT CUDA build; `on` != 0 makes orc_run use it too
void orc_set_portable_root(int on) { g_portable_root = on != 0; }
double orc_portable_root(double x, unsigned n) { return portable_root(x, n); }

}  // extern "C"
