// Header-only C++17 mirror of the reference's `Solver` builder (src/solver.rs) over the
// C ABI of include/diffurch_b200.h, for callers that are compiled code.  Same setter
// names and argument meaning as the reference; `run()` consumes the solver like
// `Solver::run(self)` (solver.rs:263): a second call throws.
//
//   auto state = dfb::Solver::make()
//                    .initial(y0)                              // [n][N], row-major
//                    .equation(DFB_SYS_LORENZ, params)         // registered functor + captures
//                    .interval(0.0, 50.0).stepsize(1.0 / 250.0)
//                    .run();                                   // -> dfb::EnsembleState
#pragma once
#include <cmath>
#include <cstdint>
#include <limits>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../../include/diffurch_b200.h"

namespace dfb {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error("diffurch_b200 error " + std::to_string(c) + ": " + m), code(c) {}
};
inline void check(int rc) {
  if (rc != DFB_OK) throw Error(rc, dfb_last_error());
}

// A user system built into a plugin shared object: returns the system id
// `equation(...)` takes.
inline int register_user_system(const std::string& so_path) {
  int32_t id = 0;
  check(dfb_register_plugin(so_path.c_str(), &id));
  return int(id);
}

// rk.rs constructors
struct RK {
  static dfb_tableau by_name(const char* name) {
    dfb_tableau t;
    check(dfb_tableau_by_name(name, &t));
    return t;
  }
  static dfb_tableau rktp64() { return by_name("rktp64"); }
  static dfb_tableau rk4() { return by_name("rk4"); }
  static dfb_tableau rk43() { return by_name("rk43"); }
  static dfb_tableau euler() { return by_name("euler"); }
  static dfb_tableau midpoint() { return by_name("midpoint"); }
  static dfb_tableau generic_order_2(double c2) {
    dfb_tableau t;
    check(dfb_tableau_generic_order_2(c2, &t));
    return t;
  }
};

// stepsize.rs:35-44
struct AutomaticStepsize {
  double stepsize;
  std::pair<double, double> stepsize_range;
  std::vector<double> atol, rtol;
  uint32_t order;
  double fac;
  std::pair<double, double> fac_range;
};

// initial_condition.rs:24 — registered history functors instead of closures
struct InitFn {
  int history_id;
};

// loc.rs:913-967
struct Loc {
  dfb_locator l{};
  Loc& every(int n) { l.filter = DFB_FILTER_EVERY, l.filter_n = n; return *this; }
  Loc& separated_by(double diff) { l.filter = DFB_FILTER_SEPARATED_BY, l.filter_a = diff; return *this; }
  Loc& in_interval(double a, double b) { l.filter = DFB_FILTER_IN_INTERVAL, l.filter_a = a, l.filter_b = b; return *this; }
  // Filter::on_times (loc.rs:665-688): strictly increasing 0-based detection indices, at most DFB_MAX_FILTER_TIMES
  Loc& on_times(std::initializer_list<int> idx) {
    l.filter = DFB_FILTER_ON_TIMES;
    l.filter_n = 0;
    for (int i : idx)
      if (l.filter_n < DFB_MAX_FILTER_TIMES) l.filter_times[l.filter_n++] = i;
    return *this;
  }
};
struct Locator {
  static Loc mk(int detector, int detection, int location) {
    Loc r;
    r.l.kind = DFB_LOC_STATEFN;
    r.l.detector_id = detector;
    r.l.detection = detection;
    r.l.location = location;
    return r;
  }
  static Loc zero(int f) { return mk(f, DFB_DET_ZERO, DFB_LOCATE_BISECTION); }
  static Loc above_zero(int f) { return mk(f, DFB_DET_ABOVE_ZERO, DFB_LOCATE_BISECTION); }
  static Loc below_zero(int f) { return mk(f, DFB_DET_BELOW_ZERO, DFB_LOCATE_BISECTION); }
  static Loc positive(int f) { return mk(f, DFB_DET_POSITIVE, DFB_LOCATE_STEP_BEGIN); }
  static Loc negative(int f) { return mk(f, DFB_DET_NEGATIVE, DFB_LOCATE_STEP_BEGIN); }
  static Loc switch_(int f) { return mk(f, DFB_DET_SWITCH, DFB_LOCATE_BISECTION_BOOL); }
  static Loc switch_true(int f) { return mk(f, DFB_DET_SWITCH_TRUE, DFB_LOCATE_BISECTION_BOOL); }
  static Loc switch_false(int f) { return mk(f, DFB_DET_SWITCH_FALSE, DFB_LOCATE_BISECTION_BOOL); }
  static Loc is_true(int f) { return mk(f, DFB_DET_IS_TRUE, DFB_LOCATE_BISECTION_BOOL); }
  static Loc is_false(int f) { return mk(f, DFB_DET_IS_FALSE, DFB_LOCATE_BISECTION_BOOL); }
  static Loc step() { return mk(0, DFB_DET_STEP, DFB_LOCATE_STEP_END); }
};
inline Loc Periodic(double period, double offset = 0.0) {  // loc.rs:292-308
  Loc r;
  r.l.kind = DFB_LOC_PERIODIC;
  r.l.period = period;
  r.l.offset = offset;
  return r;
}

struct Trace {
  int every = 1, capacity = 1024;
};

// the final State of every trajectory (solver.rs:313) + counters
struct EnsembleState {
  size_t n = 0;
  int n_state = 0, n_aux = 0;
  std::vector<double> t, p, aux;                 // [n], [n][N], [n][A]
  std::vector<uint64_t> steps, rejected, events, rhs_evals;
  std::vector<uint32_t> status;
  std::vector<uint32_t> trace_n;
  std::vector<double> trace_t, trace_p;          // [n][cap], [n][cap][N]
  std::vector<uint32_t> event_n;
  std::vector<double> event_t;
  std::vector<int32_t> event_index;
  dfb_totals totals{};
};

class Solver {
 public:
  // Solver::new (solver.rs:80-96)
  static Solver make(int device = 0) {
    Solver s;
    check(dfb_problem_default(&s.p_));
    s.device_ = device;
    return s;
  }
  Solver& initial_disco(const std::vector<std::pair<double, int>>& d) {
    if (d.size() > DFB_MAX_DISCO) throw Error(DFB_ERR_INVALID, "too many initial discontinuities");
    p_.n_disco = int(d.size());
    for (size_t i = 0; i < d.size(); ++i) p_.disco_t[i] = d[i].first, p_.disco_order[i] = d[i].second;
    return *this;
  }
  Solver& stepsize(double h) {
    p_.stepsize_kind = DFB_STEP_FIXED;
    p_.h = h;
    return *this;
  }
  Solver& stepsize(const AutomaticStepsize& a) {
    p_.stepsize_kind = DFB_STEP_AUTOMATIC;
    p_.automatic.stepsize = a.stepsize;
    p_.automatic.stepsize_min = a.stepsize_range.first;
    p_.automatic.stepsize_max = a.stepsize_range.second;
    for (size_t i = 0; i < a.atol.size() && i < DFB_MAX_STATE; ++i) p_.automatic.atol[i] = a.atol[i];
    for (size_t i = 0; i < a.rtol.size() && i < DFB_MAX_STATE; ++i) p_.automatic.rtol[i] = a.rtol[i];
    p_.automatic.order = a.order;
    p_.automatic.fac = a.fac;
    p_.automatic.fac_min = a.fac_range.first;
    p_.automatic.fac_max = a.fac_range.second;
    return *this;
  }
  Solver& max_delay(double m) {
    p_.max_delay = m;
    return *this;
  }
  Solver& equation(int system_id, std::vector<double> params = {}) {
    p_.system_id = system_id;
    check(dfb_system_info_get(system_id, &info_));
    params_ = std::move(params);
    return *this;
  }
  Solver& initial(std::vector<double> y0) {
    p_.history_id = DFB_HIST_CONSTANT;
    y0_ = std::move(y0);
    return *this;
  }
  Solver& initial(InitFn f, size_t n_traj) {
    p_.history_id = f.history_id;
    n_from_initfn_ = n_traj;
    y0_.clear();
    return *this;
  }
  Solver& interval(double t0, double t1 = std::numeric_limits<double>::infinity()) {
    p_.t0 = t0;
    p_.t1 = t1;
    return *this;
  }
  Solver& rk(const dfb_tableau& t) {
    p_.rk = t;
    return *this;
  }
  Solver& on_step(Trace tr) {
    p_.trace_every = tr.every;
    p_.trace_capacity = tr.capacity;
    return *this;
  }
  Solver& log_events(int capacity) {
    p_.event_capacity = capacity;
    return *this;
  }
  Solver& on(Loc loc, int callback_id = DFB_CB_NONE) { return append(loc, callback_id, 1); }      // solver.rs:199
  Solver& on_mut(Loc loc, int callback_id) { return append(loc, callback_id, 1); }                // solver.rs:219
  Solver& on_start_mut(int callback_id) { p_.on_start_callback_id = callback_id; return *this; }   // solver.rs:190
  Solver& on_start() { return *this; }  // solver.rs:181: read-only, observes the initial state the caller holds
  Solver& on_stop() { return *this; }   // solver.rs:172: read-only, observes the final State run() returns
  Solver& with_delayed_argument(int delayed_id, int smoothing_order) {                            // solver.rs:239
    Loc l;
    l.l.kind = DFB_LOC_PROPAGATION;
    l.l.detector_id = delayed_id;
    l.l.smoothing_order = smoothing_order;
    return append(l, DFB_CB_NONE, 0);
  }
  Solver& with_const_delay(double delay, int smoothing_order) {                                   // solver.rs:251
    p_.max_delay = std::fmax(p_.max_delay, delay);
    Loc l;
    l.l.kind = DFB_LOC_PROPAGATION;
    l.l.detector_id = 0;
    l.l.delay = delay;
    l.l.smoothing_order = smoothing_order;
    return append(l, DFB_CB_NONE, 0);
  }
  Solver& arith(dfb_arith a) {
    p_.arith = a;
    return *this;
  }
  Solver& max_steps(int64_t n) {
    p_.max_steps = n;
    return *this;
  }
  const dfb_problem& problem() const { return p_; }

  // Solver::run(self)
  EnsembleState run() {
    if (consumed_) throw Error(DFB_ERR_STATE, "Solver::run(self) consumes the solver (solver.rs:263)");
    consumed_ = true;
    if (info_.n_state == 0) throw Error(DFB_ERR_INVALID, "no equation set");
    const int N = info_.n_state, P = info_.n_params, A = info_.n_aux;
    size_t n = y0_.empty() ? n_from_initfn_ : y0_.size() / size_t(N);
    if (n == 0) throw Error(DFB_ERR_INVALID, "no initial condition set");
    if (y0_.empty()) y0_.assign(n * N, 0.0);
    if (P > 0 && params_.size() == size_t(P) && n > 1) {  // one capture set for every trajectory
      std::vector<double> rep(n * P);
      for (size_t i = 0; i < n; ++i)
        for (int c = 0; c < P; ++c) rep[i * P + c] = params_[c];
      params_.swap(rep);
    }
    if (size_t(P) * n != params_.size()) throw Error(DFB_ERR_INVALID, "params shape mismatch");
    dfb_handle* h = nullptr;
    check(dfb_create(&p_, n, device_, &h));
    struct Guard {
      dfb_handle* h;
      ~Guard() { dfb_destroy(h); }
    } guard{h};
    check(dfb_set_initial(h, y0_.data()));
    if (P > 0) check(dfb_set_params(h, params_.data()));
    check(dfb_run(h, nullptr));
    EnsembleState st;
    st.n = n;
    st.n_state = N;
    st.n_aux = A;
    st.t.resize(n);
    st.p.resize(n * N);
    st.aux.resize(n * size_t(A));
    check(dfb_get_final(h, st.t.data(), st.p.data(), A ? st.aux.data() : nullptr));
    st.steps.resize(n), st.rejected.resize(n), st.events.resize(n), st.rhs_evals.resize(n), st.status.resize(n);
    dfb_stats s{st.steps.data(), st.rejected.data(), st.events.data(), st.rhs_evals.data(), st.status.data()};
    check(dfb_get_stats(h, &s));
    if (p_.trace_every > 0 && p_.trace_capacity > 0) {
      st.trace_n.resize(n);
      st.trace_t.resize(n * size_t(p_.trace_capacity));
      st.trace_p.resize(n * size_t(p_.trace_capacity) * N);
      check(dfb_get_trace(h, st.trace_n.data(), st.trace_t.data(), st.trace_p.data()));
    }
    if (p_.event_capacity > 0) {
      st.event_n.resize(n);
      st.event_t.resize(n * size_t(p_.event_capacity));
      st.event_index.resize(n * size_t(p_.event_capacity));
      check(dfb_get_events(h, st.event_n.data(), st.event_t.data(), st.event_index.data()));
    }
    check(dfb_get_totals(h, &st.totals));
    return st;
  }

 private:
  Solver& append(const Loc& loc, int callback_id, int dedup) {
    if (p_.n_loc >= DFB_MAX_LOCATORS) throw Error(DFB_ERR_INVALID, "too many locators");
    dfb_locator l = loc.l;
    l.callback_id = callback_id;
    l.dedup = dedup;
    p_.loc[p_.n_loc++] = l;
    return *this;
  }
  dfb_problem p_{};
  dfb_system_info info_{};
  std::vector<double> y0_, params_;
  size_t n_from_initfn_ = 0;
  int device_ = 0;
  bool consumed_ = false;
};

}  // namespace dfb
