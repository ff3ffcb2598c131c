// =============================================================================
// diffurch_oracle.hpp — TEST INFRASTRUCTURE ONLY.  NOT PART OF THE PRODUCT.
//
// CPU restatement (C++17) of the hot path of Danila-Bain/diffurch, operation
// for operation, used as the parity oracle for the CUDA kernels and as the
// "port" CPU baseline of bench.py.  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may build, load or call
// anything under oracle/.  The product (diffurch_b200/, include/) never does.
//
// Why a restatement: the reference is a nightly-Rust crate with an un-vendored
// git dependency (hlist2_trait_macro@0c5ccef) and there is no cargo/rustc in
// this environment, so it cannot be compiled here.
//
// Parity pinning: oracle/ is checked by tests/test_oracle_*.py against every
// live assertion of the reference's own tests: the exact f64 step-time vectors
// of tests/delay_propagation.rs:21-24,45-50,68-71,91-93; the exact equalities
// of tests/equation_types.rs:15,38-40,62-64; the tolerances of
// tests/equation_types.rs:77,94,110,133,155; the Lyapunov bounds of
// tests/lyapunov_exponents.rs; and src/util/partition_point.rs:61-128.
// `AutomaticStepsize` (src/stepsize.rs:64-85) has NO test in the reference:
// for that one function the oracle is "parity unpinned" (restated from source
// only).
//
// Must be compiled with -ffp-contract=off: Rust never contracts a*b+c to FMA.
//
// File:line citations are into /root/reference/.
// =============================================================================
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <functional>
#include <limits>
#include <memory>
#include <stdexcept>
#include <utility>
#include <vector>

namespace orc {

// ---------------------------------------------------------------------------
// RealVectorSpace (src/traits.rs:5-31): element-wise, separately rounded ops.
// ---------------------------------------------------------------------------
template <int N>
struct Vec {
  double v[N];
  static Vec zero() {
    Vec r;
    for (int i = 0; i < N; ++i) r.v[i] = 0.0;
    return r;
  }
  double& operator[](int i) { return v[i]; }
  const double& operator[](int i) const { return v[i]; }
  Vec operator+(const Vec& o) const {
    Vec r;
    for (int i = 0; i < N; ++i) r.v[i] = v[i] + o.v[i];
    return r;
  }
  Vec operator-(const Vec& o) const {
    Vec r;
    for (int i = 0; i < N; ++i) r.v[i] = v[i] - o.v[i];
    return r;
  }
  Vec operator*(double s) const {
    Vec r;
    for (int i = 0; i < N; ++i) r.v[i] = v[i] * s;
    return r;
  }
  Vec operator/(double s) const {
    Vec r;
    for (int i = 0; i < N; ++i) r.v[i] = v[i] / s;
    return r;
  }
  Vec operator-() const {
    Vec r;
    for (int i = 0; i < N; ++i) r.v[i] = -v[i];
    return r;
  }
  Vec& operator+=(const Vec& o) {
    for (int i = 0; i < N; ++i) v[i] = v[i] + o.v[i];
    return *this;
  }
};

// Rust `f64::powi` lowers to llvm.powi = compiler-rt __powidf2: square-and-
// multiply from the low bit (th^2 = th*th, th^3 = th*th^2,
// th^4 = th^2*th^2).  Restated here so the rounding sequence is explicit.
inline double powi(double a, int b) {
  const bool recip = b < 0;
  double r = 1.0;
  while (true) {
    if (b & 1) r *= a;
    b /= 2;
    if (b == 0) break;
    a *= a;
  }
  return recip ? 1.0 / r : r;
}

// ---------------------------------------------------------------------------
// ButcherTableu (src/rk.rs:5-14) and dense_output (src/rk.rs:17-52).
// ---------------------------------------------------------------------------
template <int S, int I>
struct ButcherTableu {
  int order = 0, order_embedded = 0, order_interpolant = 0;
  double a[S][S] = {};
  double b[S] = {};
  double b2[S] = {};
  double c[S] = {};
  double bi[S][I] = {};

  // rk.rs:17-52
  template <int D, int N>
  Vec<N> dense_output(const Vec<N>& y_prev, double t_step, double theta,
                      const Vec<N>* k) const {
    static_assert(D == 0 || D == 1, "rk.rs:48 unimplemented!()");
    Vec<N> delta = Vec<N>::zero();
    if (D == 0) {
      for (int i = 0; i < S; ++i) {
        double b_i = 0.0;
        for (int j = 0; j < I; ++j) b_i += bi[i][j] * powi(theta, j);
        delta += k[i] * b_i;
      }
      return y_prev + delta * t_step;
    } else {
      for (int i = 0; i < S; ++i) {
        double b_i = 0.0;
        for (int j = 1; j < I; ++j) b_i += bi[i][j] * powi(theta, j - 1) * double(j);
        delta += k[i] * b_i;
      }
      return delta;
    }
  }
};

// ---------------------------------------------------------------------------
// Errors standing in for the reference's panics.
// ---------------------------------------------------------------------------
enum TrajStatus : uint32_t {
  TRAJ_OK = 0,
  TRAJ_HISTORY_DELETED = 1,    // state.rs:166-171
  TRAJ_HISTORY_NOT_READY = 2,  // state.rs:172-177
  TRAJ_NO_DERIVATIVE = 4,      // initial_condition.rs:33,47
  TRAJ_NONFINITE = 8,
  TRAJ_STEP_LIMIT = 16,
  TRAJ_STOPPED = 64,
  TRAJ_STEPSIZE_FLOOR = 256,
};
struct Panic {
  uint32_t status;
};

// ---------------------------------------------------------------------------
// VecDeque stand-in: growable power-of-two ring (what alloc::VecDeque is).
// ---------------------------------------------------------------------------
template <class T>
class RingDeque {
 public:
  RingDeque() : buf_(8), head_(0), len_(0) {}
  size_t size() const { return len_; }
  bool empty() const { return len_ == 0; }
  void push_back(const T& x) {
    if (len_ == buf_.size()) grow();
    buf_[(head_ + len_) & (buf_.size() - 1)] = x;
    ++len_;
  }
  void pop_front() {
    head_ = (head_ + 1) & (buf_.size() - 1);
    --len_;
  }
  const T& operator[](size_t i) const { return buf_[(head_ + i) & (buf_.size() - 1)]; }
  const T& front() const { return (*this)[0]; }
  const T& back() const { return (*this)[len_ - 1]; }
  // VecDeque::partition_point (binary search), used at state.rs:165
  template <class P>
  size_t partition_point(P pred) const {
    size_t lo = 0, hi = len_;
    while (lo < hi) {
      size_t mid = lo + (hi - lo) / 2;
      if (pred((*this)[mid])) lo = mid + 1;
      else hi = mid;
    }
    return lo;
  }

 private:
  void grow() {
    std::vector<T> nb(buf_.size() * 2);
    for (size_t i = 0; i < len_; ++i) nb[i] = (*this)[i];
    buf_.swap(nb);
    head_ = 0;
  }
  std::vector<T> buf_;
  size_t head_, len_;
};

// src/util/partition_point.rs:26-52
template <class T, class P>
size_t partition_point_linear(const RingDeque<T>& deque, size_t start, P pred) {
  if (deque.empty()) return 0;
  size_t i = start < deque.size() - 1 ? start : deque.size() - 1;
  if (pred(deque[i])) {
    i += 1;
    while (i < deque.size() && pred(deque[i])) i += 1;
  } else {
    while (i > 0 && !pred(deque[i - 1])) i -= 1;
  }
  return i;
}

// ---------------------------------------------------------------------------
// InitialCondition (src/initial_condition.rs:9-50).
// ---------------------------------------------------------------------------
template <int N>
struct InitialCondition {
  std::function<Vec<N>(double)> f;   // D = 0
  std::function<Vec<N>(double)> df;  // D = 1, empty => unimplemented (InitFn<F,()>)
  bool constant = false;             // constant: all derivatives are zero (:15-22)
  static InitialCondition constant_value(const Vec<N>& y0) {
    InitialCondition ic;
    ic.constant = true;
    ic.f = [y0](double) { return y0; };
    return ic;
  }
  static InitialCondition init_fn(std::function<Vec<N>(double)> f) {
    InitialCondition ic;
    ic.f = std::move(f);
    return ic;
  }
  static InitialCondition init_fn(std::function<Vec<N>(double)> f,
                                  std::function<Vec<N>(double)> df) {
    InitialCondition ic;
    ic.f = std::move(f);
    ic.df = std::move(df);
    return ic;
  }
  template <int D>
  Vec<N> eval(double t) const {
    if (D == 0) return f(t);
    if (constant) return Vec<N>::zero();
    if (D == 1 && df) return df(t);
    throw Panic{TRAJ_NO_DERIVATIVE};
  }
};

// ---------------------------------------------------------------------------
// StateHistory (src/state/state.rs:11-23, 161-187)
// ---------------------------------------------------------------------------
template <int N, int S>
struct KBlock {
  Vec<N> k[S];
};
struct Disco {
  double t;
  size_t order;
};

template <int N, int S, int I>
struct StateHistory {
  double t_span = 0.0;
  double t_init = 0.0;
  InitialCondition<N> p_init;
  RingDeque<double> t_deque;
  RingDeque<Vec<N>> p_deque;
  RingDeque<KBlock<N, S>> k_deque;
  RingDeque<Disco> disco_deque;
  ButcherTableu<S, I> rk;

  // state.rs:161-187
  template <int D>
  Vec<N> eval(double t) const {
    if (t <= t_init) return p_init.template eval<D>(t);
    size_t i = t_deque.partition_point([&](double t_i) { return t_i <= t; });
    if (i == 0) throw Panic{TRAJ_HISTORY_DELETED};
    if (i == t_deque.size()) throw Panic{TRAJ_HISTORY_NOT_READY};
    const Vec<N>& y_prev = p_deque[i - 1];
    const KBlock<N, S>& k = k_deque[i - 1];
    double t_prev = t_deque[i - 1];
    double t_next = t_deque[i];
    double t_step = t_next - t_prev;
    double theta = (t - t_prev) / t_step;
    return rk.template dense_output<D, N>(y_prev, t_step, theta, k.k);
  }
};

// ---------------------------------------------------------------------------
// StateRef / StateRefMut (src/state/state_fn.rs:13-98): the closure argument.
// ---------------------------------------------------------------------------
template <int N, int S, int I>
struct StateRef {
  double t;
  const Vec<N>& p;
  const Vec<N>& d;
  double t_prev;
  const Vec<N>& p_prev;
  const StateHistory<N, S, I>& history;
  Vec<N> p_at(double tt) const { return history.template eval<0>(tt); }  // state_fn.rs:46
  Vec<N> d_at(double tt) const { return history.template eval<1>(tt); }  // state_fn.rs:49
};
template <int N, int S, int I>
struct StateRefMut {
  double& t;
  Vec<N>& p;
  const Vec<N>& d;
  double t_prev;
  const Vec<N>& p_prev;
  StateHistory<N, S, I>& history;
  Vec<N> p_at(double tt) const { return history.template eval<0>(tt); }
  Vec<N> d_at(double tt) const { return history.template eval<1>(tt); }
  void stop_integration() { t = std::numeric_limits<double>::infinity(); }  // state_fn.rs:92-97
};

// ---------------------------------------------------------------------------
// State (src/state/state.rs:26-150)
// ---------------------------------------------------------------------------
template <int N, int S, int I>
struct State {
  StateHistory<N, S, I> history;
  double t_curr, t_prev;
  Vec<N> p_curr, p_prev, d_curr, d_prev, e_curr;
  ButcherTableu<S, I> rk;
  Vec<N> k_curr[S];
  uint64_t rhs_evals = 0;

  // state.rs:52-81
  State(double t_init, double t_span, InitialCondition<N> p_init,
        const std::vector<Disco>& disco_init, const ButcherTableu<S, I>& rk_)
      : rk(rk_) {
    Vec<N> p = p_init.template eval<0>(t_init);
    t_curr = t_prev = t_init;
    p_curr = p_prev = p;
    d_curr = d_prev = e_curr = Vec<N>::zero();
    for (int i = 0; i < S; ++i) k_curr[i] = Vec<N>::zero();
    history.rk = rk_;
    history.t_init = t_init;
    history.t_span = t_span;
    history.p_init = std::move(p_init);
    history.t_deque.push_back(t_init);
    history.p_deque.push_back(p);
    for (const Disco& d : disco_init) history.disco_deque.push_back(d);
  }

  // views (state_fn.rs:170-201)
  StateRef<N, S, I> view_curr() const {
    return StateRef<N, S, I>{t_curr, p_curr, d_curr, t_prev, p_prev, history};
  }
  StateRef<N, S, I> view_prev() const {
    return StateRef<N, S, I>{t_prev, p_prev, d_prev, t_prev, p_prev, history};
  }

  // state.rs:83-92
  template <int D>
  Vec<N> eval(double t) const {
    if (t >= t_prev && t <= t_curr) {
      double t_step = t_curr - t_prev;
      double theta = (t - t_prev) / t_step;
      return rk.template dense_output<D, N>(p_prev, t_step, theta, k_curr);
    }
    return history.template eval<D>(t);
  }

  // state.rs:94-120
  template <class Rhs>
  void make_step(Rhs& rhs, double t_step) {
    if (t_prev != t_curr) {
      k_curr[0] = d_curr;
    } else {
      k_curr[0] = rhs(view_curr());
      ++rhs_evals;
    }
    t_prev = t_curr;
    p_prev = p_curr;
    d_prev = d_curr;
    for (int i = 1; i < S; ++i) {
      t_curr = t_prev + rk.c[i] * t_step;
      Vec<N> acc = Vec<N>::zero();
      for (int j = 0; j < i; ++j) acc = acc + k_curr[j] * rk.a[i][j];
      p_curr = p_prev + acc * t_step;
      k_curr[i] = rhs(view_curr());
      ++rhs_evals;
    }
    {
      Vec<N> acc = Vec<N>::zero();
      for (int j = 0; j < S; ++j) acc = acc + k_curr[j] * rk.b[j];
      p_curr = p_prev + acc * t_step;
    }
    t_curr = t_prev + t_step;
    d_curr = rhs(view_curr());
    ++rhs_evals;
    {
      Vec<N> acc = Vec<N>::zero();
      for (int j = 0; j < S; ++j) acc = acc + k_curr[j] * (rk.b2[j] - rk.b[j]);
      e_curr = acc * t_step;
    }
  }

  // state.rs:122-139
  void commit_step() {
    history.t_deque.push_back(t_curr);
    history.p_deque.push_back(p_curr);
    KBlock<N, S> kb;
    for (int i = 0; i < S; ++i) kb.k[i] = k_curr[i];
    history.k_deque.push_back(kb);
    double t_tail = t_prev - history.t_span;
    while (history.t_deque.size() > 1 && history.t_deque[1] < t_tail) {
      history.t_deque.pop_front();
      history.p_deque.pop_front();
      history.k_deque.pop_front();
    }
    while (!history.disco_deque.empty() && history.disco_deque.front().t < t_tail)
      history.disco_deque.pop_front();
  }
  // state.rs:141-144
  void make_zero_step() {
    t_prev = t_curr;
    p_prev = p_curr;
  }
  // state.rs:146-150
  void undo_step() {
    t_curr = t_prev;
    p_curr = p_prev;
    d_curr = d_prev;
  }
};

// ---------------------------------------------------------------------------
// Step-size control (src/stepsize.rs)
// ---------------------------------------------------------------------------
enum class StepStatus { Rejected, Accepted };

// The reference computes the step-size factor with f64::powf, i.e. the platform's libm `pow`
// (stepsize.rs:74): its last bit differs between libm builds, and CUDA's pow is not glibc's.  To pin the
// CONTROL FLOW of the adaptive path (error norm, clamps, reject loop, k[0] re-evaluation) bit for bit on
// the GPU, the n-th root can be switched to `portable_root`: frexp / ldexp (exact) and a fixed number of
// Newton steps in plain IEEE +, *, / — the same bits on any IEEE machine compiled without contraction.
// The STRICT CUDA build always uses its own copy of this function (csrc/dev/integrator.cuh); the oracle
// uses it when `g_portable_root` is set (tests), libm pow otherwise (the reference's behaviour).  It agrees
// with pow(x, 1/n) to 2 ulp for 1e-4 < x < 1e4 (up to 10 ulp at |ln x| = 60: pow's exponent is the ROUNDED
// 1/n, so pow itself is that far from the n-th root; tests/test_oracle_independent.py).
inline bool g_portable_root = false;
inline double portable_root(double x, unsigned n) {
  if (x != x) return x;
  if (x < 0.0) return std::numeric_limits<double>::quiet_NaN();
  if (x == 0.0) return 0.0;                                            // pow(0, 1/n) = 0
  if (x == std::numeric_limits<double>::infinity()) return x;          // pow(inf, 1/n) = inf
  if (n <= 1u) return x;
  int e = 0;
  const double m = std::frexp(x, &e);                                  // x = m 2^e, m in [0.5, 1)
  int q = e / int(n), r = e - q * int(n);
  if (r < 0) {
    r += int(n);
    q -= 1;
  }
  const double a = std::ldexp(m, r);                                   // a in [0.5, 2^(n-1)), x = a 2^(q n)
  double y = a > 1.0 ? a : 1.0;                                        // Newton from above: monotone, then quadratic
  const double nm1 = double(n - 1u), dn = double(n);
  for (int it = 0; it < 48; ++it) {
    double p = 1.0;
    for (unsigned j = 0; j + 1u < n; ++j) p = p * y;                   // y^(n-1)
    y = (nm1 * y + a / p) / dn;
  }
  return std::ldexp(y, q);
}

template <int N>
struct Stepsize {
  bool automatic = false;
  // fixed: bare scalar (stepsize.rs:18-32); automatic: AutomaticStepsize (:35-44)
  double stepsize = 0.05;
  double range_lo = 0.0, range_hi = std::numeric_limits<double>::infinity();
  Vec<N> atol = Vec<N>::zero(), rtol = Vec<N>::zero();
  uint32_t order = 4;
  double fac = 0.9, fac_lo = 0.2, fac_hi = 5.0;

  double get() const { return stepsize; }
  // stepsize.rs:29-31 (fixed) and :64-85 (automatic)
  StepStatus update(const Vec<N>& error) {
    if (!automatic) return StepStatus::Accepted;
    // .map(..).reduce(T::max): nalgebra's RealField::max for f64 is f64::max
    double err = 0.0;
    bool first = true;
    for (int i = 0; i < N; ++i) {
      double e = std::fabs(error[i]) / (atol[i] + std::fabs(error[i]) * rtol[i]);
      err = first ? e : std::fmax(err, e);
      first = false;
    }
    double factor = fac * (g_portable_root ? portable_root(1.0 / err, order + 1u) : std::pow(1.0 / err, 1.0 / double(order + 1)));
    factor = clamp(factor, fac_lo, fac_hi);
    stepsize *= factor;
    stepsize = clamp(stepsize, range_lo, range_hi);
    return err >= 1.0 ? StepStatus::Rejected : StepStatus::Accepted;
  }
  // f64::clamp: NaN stays NaN
  static double clamp(double x, double lo, double hi) {
    if (x < lo) return lo;
    if (x > hi) return hi;
    return x;
  }
};

// ---------------------------------------------------------------------------
// Event location (src/loc.rs)
// ---------------------------------------------------------------------------
enum class Detection {  // loc.rs:127-143
  Zero, AboveZero, BelowZero, Positive, Negative,
  Switch, SwitchTrue, SwitchFalse, IsTrue, IsFalse, Step
};
enum class Location {  // loc.rs:200-285
  StepBegin, StepEnd, StepMiddle, Lerp, Bisection, BisectionBool, RegulaFalsi
};

// EvalState (state_fn.rs:131-143): a scalar function of the state, evaluated
// on the current view, the previous view, or the dense output at time t.
template <int N, int S, int I>
struct ScalarStateFn {
  std::function<double(const StateRef<N, S, I>&)> f;
  double eval_curr(const State<N, S, I>& st) const { return f(st.view_curr()); }
  double eval_prev(const State<N, S, I>& st) const { return f(st.view_prev()); }
  // state_fn.rs:190-201
  double eval_at(const State<N, S, I>& st, double t) const {
    Vec<N> y = st.template eval<0>(t);
    Vec<N> dy = st.template eval<1>(t);
    return f(StateRef<N, S, I>{t, y, dy, t, y, st.history});
  }
};

template <int N, int S, int I>
struct LocBase {
  using St = State<N, S, I>;
  virtual ~LocBase() = default;
  virtual bool detect(const St& st) = 0;
  virtual double locate(const St& st) = 0;
  // Locate::detect_and_locate default (loc.rs:31-33)
  virtual bool detect_and_locate(const St& st, double* time) {
    if (!detect(st)) return false;
    *time = locate(st);
    return true;
  }
  // EvalMutState::eval_mut of the attached callback
  virtual void eval_mut(St& st) = 0;
};

// LocatorStateFn<.., F, Detection, Location> (loc.rs:36-80) + LocCallback (:710-786)
template <int N, int S, int I>
struct LocatorStateFn : LocBase<N, S, I> {
  using St = State<N, S, I>;
  ScalarStateFn<N, S, I> f;  // bool-valued detectors return 0.0 / 1.0
  Detection detection;
  Location location;
  std::function<void(StateRefMut<N, S, I>&)> callback;  // empty = `|_| {}`

  bool detect(const St& st) override {
    switch (detection) {
      case Detection::Zero: {
        double curr = f.eval_curr(st), prev = f.eval_prev(st);
        return (curr > 0.0 && prev <= 0.0) || (curr <= 0.0 && prev > 0.0);
      }
      case Detection::AboveZero: {
        double curr = f.eval_curr(st), prev = f.eval_prev(st);
        return curr > 0.0 && prev <= 0.0;
      }
      case Detection::BelowZero: {
        double curr = f.eval_curr(st), prev = f.eval_prev(st);
        return curr < 0.0 && prev >= 0.0;
      }
      case Detection::Positive: return f.eval_curr(st) >= 0.0;
      case Detection::Negative: return f.eval_curr(st) <= 0.0;
      case Detection::Switch: {
        bool curr = f.eval_curr(st) != 0.0, prev = f.eval_prev(st) != 0.0;
        return curr != prev;
      }
      case Detection::SwitchTrue: {
        bool curr = f.eval_curr(st) != 0.0, prev = f.eval_prev(st) != 0.0;
        return curr && !prev;
      }
      case Detection::SwitchFalse: {
        bool curr = f.eval_curr(st) != 0.0, prev = f.eval_prev(st) != 0.0;
        return !curr && prev;
      }
      case Detection::IsTrue: return f.eval_curr(st) != 0.0;
      case Detection::IsFalse: return !(f.eval_curr(st) != 0.0);
      case Detection::Step: return true;
    }
    return false;
  }

  double locate(const St& st) override { return locate_with(f, location, st); }

  // loc.rs:200-285, shared with the Propagator
  template <class F>
  static double locate_with(const F& fn, Location location, const St& st) {
    switch (location) {
      case Location::StepBegin: return st.t_prev;
      case Location::StepEnd: return st.t_curr;
      case Location::StepMiddle: return 0.5 * (st.t_curr - st.t_prev);  // sic, loc.rs:202-204
      case Location::Lerp: {
        double curr = fn.eval_curr(st), prev = fn.eval_prev(st);
        return (curr * st.t_prev - prev * st.t_curr) / (curr - prev);
      }
      case Location::BisectionBool: {
        double l = st.t_prev, r = st.t_curr;
        double m = 0.5 * (l + r);
        if (fn.eval_prev(st) != 0.0) std::swap(l, r);
        double w = std::fabs(r - l);
        double w_prev = 2.0 * w;
        while (w < w_prev) {
          w_prev = w;
          if (fn.eval_at(st, m) != 0.0) r = m;
          else l = m;
          m = 0.5 * (l + r);
          w = std::fabs(r - l);
        }
        return rust_max(l, r);
      }
      case Location::Bisection: {
        double l = st.t_prev, r = st.t_curr;
        double m = 0.5 * (l + r);
        if (fn.eval_curr(st) < 0.0) std::swap(l, r);
        double w = std::fabs(r - l);
        double w_prev = 2.0 * w;
        while (w < w_prev) {
          w_prev = w;
          if (fn.eval_at(st, m) < 0.0) l = m;
          else r = m;
          m = 0.5 * (l + r);
          w = std::fabs(r - l);
        }
        return rust_max(l, r);
      }
      case Location::RegulaFalsi: {
        double l = st.t_prev, r = st.t_curr;
        if (fn.eval_curr(st) < 0.0) std::swap(l, r);
        double w = std::fabs(r - l);
        double w_prev = 2.0 * w;
        while (w < w_prev) {
          w_prev = w;
          double f_l = fn.eval_at(st, l);
          double f_r = fn.eval_at(st, r);
          double m = (f_r * l - f_l * r) / (f_r - f_l);
          double f_m = fn.eval_at(st, m);
          if (f_m < 0.0) r = m;
          else l = m;
          w = std::fabs(r - l);
        }
        return rust_max(l, r);
      }
    }
    return st.t_curr;
  }
  static double rust_max(double a, double b) { return std::fmax(a, b); }

  void eval_mut(St& st) override {
    if (callback) {
      StateRefMut<N, S, I> m{st.t_curr, st.p_curr, st.d_curr, st.t_prev, st.p_prev, st.history};
      callback(m);
    }
  }
};

// Periodic (loc.rs:292-335)
template <int N, int S, int I>
struct Periodic : LocBase<N, S, I> {
  using St = State<N, S, I>;
  double period, offset;
  std::function<void(StateRefMut<N, S, I>&)> callback;
  bool detect(const St& st) override {
    double prev = std::floor((st.t_prev - offset) / period);
    double curr = std::floor((st.t_curr - offset) / period);
    return prev < curr;
  }
  double locate(const St& st) override {
    return std::floor((st.t_curr - offset) / period) * period + offset;
  }
  void eval_mut(St& st) override {
    if (callback) {
      StateRefMut<N, S, I> m{st.t_curr, st.p_curr, st.d_curr, st.t_prev, st.p_prev, st.history};
      callback(m);
    }
  }
};

// Propagator + Propagation (loc.rs:344-454)
template <int N, int S, int I>
struct PropagatedDiscontinuity : LocBase<N, S, I> {
  using St = State<N, S, I>;
  ScalarStateFn<N, S, I> delayed_argument;
  size_t smoothing_order = 0;
  double tracked_disco_t = -std::numeric_limits<double>::max();  // T::min_value()
  size_t tracked_disco_order = std::numeric_limits<size_t>::max();
  size_t tracked_queue_index = 0;

  // loc.rs:378-410
  bool detect(const St& st) override {
    double prev = delayed_argument.eval_prev(st);
    double curr = delayed_argument.eval_curr(st);
    size_t partition_prev = partition_point_linear(
        st.history.disco_deque, tracked_queue_index, [&](const Disco& d) { return d.t <= prev; });
    size_t partition_curr = partition_point_linear(
        st.history.disco_deque, partition_prev, [&](const Disco& d) { return d.t <= curr; });
    tracked_queue_index = partition_prev < partition_curr ? partition_prev : partition_curr;
    if (partition_prev != partition_curr) {
      const Disco& d = st.history.disco_deque[tracked_queue_index];
      tracked_disco_t = d.t;
      tracked_disco_order = d.order + smoothing_order;
      return true;
    }
    return false;
  }
  // EvalState for Propagator (loc.rs:424-432): delayed argument minus tracked time
  struct Shifted {
    const PropagatedDiscontinuity* self;
    double eval_curr(const St& st) const {
      return self->delayed_argument.eval_curr(st) - self->tracked_disco_t;
    }
    double eval_prev(const St& st) const {
      return self->delayed_argument.eval_prev(st) - self->tracked_disco_t;
    }
    double eval_at(const St& st, double t) const {
      return self->delayed_argument.eval_at(st, t) - self->tracked_disco_t;
    }
  };
  double locate(const St& st) override {
    return LocatorStateFn<N, S, I>::locate_with(Shifted{this}, Location::Bisection, st);
  }
  // loc.rs:446-453
  void eval_mut(St& st) override {
    if (tracked_disco_order < size_t(st.rk.order))
      st.history.disco_deque.push_back(Disco{st.t_curr, tracked_disco_order});
  }
};

// Filters (loc.rs:465-700).  The filter closure is stateful (FnMut).
template <int N, int S, int I>
struct FilterAfterDetection : LocBase<N, S, I> {  // loc.rs:465-505: every(n), on_times
  using St = State<N, S, I>;
  std::unique_ptr<LocBase<N, S, I>> loc;
  std::function<bool(const StateRef<N, S, I>&)> filter;
  bool detect(const St& st) override { return loc->detect(st) && filter(st.view_curr()); }
  double locate(const St& st) override { return loc->locate(st); }
  void eval_mut(St& st) override { loc->eval_mut(st); }
};
template <int N, int S, int I>
struct FilterLocated : LocBase<N, S, I> {  // loc.rs:549-601: separated_by, in_interval
  using St = State<N, S, I>;
  std::unique_ptr<LocBase<N, S, I>> loc;
  std::function<bool(const StateRef<N, S, I>&)> filter;
  bool filter_at(const St& st, double t) {
    Vec<N> y = st.template eval<0>(t);
    Vec<N> dy = st.template eval<1>(t);
    return filter(StateRef<N, S, I>{t, y, dy, t, y, st.history});
  }
  bool detect(const St& st) override {
    double t;
    return loc->detect_and_locate(st, &t) && filter_at(st, t);
  }
  double locate(const St& st) override { return loc->locate(st); }
  // NB: `Solver::on` wraps the locator in LocCallback, whose Locate impl uses
  // the DEFAULT detect_and_locate (loc.rs:733-746), i.e. detect() then
  // locate(); FilterLocated's own override (loc.rs:592-600) is never reached
  // through the builder, so it is not restated.
  void eval_mut(St& st) override { loc->eval_mut(st); }
};
// Filter::every (loc.rs:613-632)
inline std::function<bool()> make_every_counter(size_t n) {
  auto i = std::make_shared<size_t>(n - 1);
  return [i, n]() {
    *i += 1;
    if (*i >= n) {
      *i = 0;
      return true;
    }
    return false;
  };
}

// DedupLocF (loc.rs:970-1052)
template <int N, int S, int I>
struct DedupLocF : LocBase<N, S, I> {
  using St = State<N, S, I>;
  bool has_last_call = false;
  double last_call = 0.0;
  std::unique_ptr<LocBase<N, S, I>> loc_f;
  bool detect(const St& st) override {  // never reached via run (loc.rs:984-988 self-recurses)
    return (!has_last_call || st.t_prev > last_call) && loc_f->detect(st);
  }
  double locate(const St& st) override { return loc_f->locate(st); }
  bool detect_and_locate(const St& st, double* time) override {  // loc.rs:1002-1011
    if (!has_last_call || st.t_prev > last_call) return loc_f->detect_and_locate(st, time);
    return false;
  }
  void eval_mut(St& st) override {  // loc.rs:1023-1026
    has_last_call = true;
    last_call = st.t_curr;
    loc_f->eval_mut(st);
  }
};

// ---------------------------------------------------------------------------
// Solver (src/solver.rs)
// ---------------------------------------------------------------------------
template <int N, int S, int I, class Rhs>
struct Solver {
  using St = State<N, S, I>;
  using Ref = StateRef<N, S, I>;
  using RefMut = StateRefMut<N, S, I>;

  Rhs equation;
  InitialCondition<N> initial;
  std::vector<Disco> initial_disco;
  double t_init = 0.0, t_end = std::numeric_limits<double>::infinity();  // interval.rs:8-25
  ButcherTableu<S, I> rk;
  Stepsize<N> stepsize;
  double max_delay = 0.0;
  std::vector<std::function<void(const Ref&)>> events_on_step, events_on_start, events_on_stop;
  std::function<void(StateRefMut<N, S, I>&)> on_start_mut;
  std::vector<std::unique_ptr<LocBase<N, S, I>>> events_on_loc;
  uint64_t max_steps = 0;  // oracle-only guard for unbounded intervals

  // per-run counters (not in the reference; for the stats the product reports)
  uint64_t n_steps = 0, n_rejected = 0, n_events = 0;
  std::vector<std::pair<double, int>> event_log;
  bool log_events = false;

  explicit Solver(Rhs rhs, const ButcherTableu<S, I>& tab) : equation(std::move(rhs)), rk(tab) {}

  // Solver::on / on_mut (solver.rs:199-236): DedupLocF{ LocCallback(loc, cb) }
  void on(std::unique_ptr<LocBase<N, S, I>> loc) {
    auto d = std::make_unique<DedupLocF<N, S, I>>();
    d->loc_f = std::move(loc);
    events_on_loc.push_back(std::move(d));
  }
  // Solver::with_delayed_argument (solver.rs:239-248): NOT deduplicated
  void with_delayed_argument(std::function<double(const Ref&)> delayed, size_t smoothing_order) {
    auto p = std::make_unique<PropagatedDiscontinuity<N, S, I>>();
    p->delayed_argument.f = std::move(delayed);
    p->smoothing_order = smoothing_order;
    events_on_loc.push_back(std::move(p));
  }
  // Solver::with_const_delay (solver.rs:251-261)
  void with_const_delay(double delay, size_t smoothing_order) {
    max_delay = std::fmax(max_delay, delay);
    with_delayed_argument([delay](const Ref& s) { return s.t - delay; }, smoothing_order);
  }

  // HListLocateEarliest::locate_earliest (loc.rs:821-835, 877-882)
  bool locate_earliest(const St& st, size_t* index, double* time) {
    bool found = false;
    size_t earliest_index = 0;
    double earliest_time = 0.0;
    for (size_t self_index = 0; self_index < events_on_loc.size(); ++self_index) {
      double self_time;
      if (events_on_loc[self_index]->detect_and_locate(st, &self_time) &&
          (!found || self_time < earliest_time)) {
        earliest_index = self_index;
        earliest_time = self_time;
        found = true;
      }
    }
    *index = earliest_index;
    *time = earliest_time;
    return found;
  }

  static void fire(std::vector<std::function<void(const Ref&)>>& list, St& st) {
    for (auto& cb : list) cb(st.view_curr());  // state_fn.rs:255-261: non-mut fn => eval_curr
  }

  // solver.rs:263-314
  St run() {
    St state(t_init, max_delay, initial, initial_disco, rk);
    auto& rhs = equation;
    fire(events_on_start, state);
    if (on_start_mut) {  // a `new_mut` element of events_on_start (solver.rs:190-197)
      StateRefMut<N, S, I> m{state.t_curr, state.p_curr, state.d_curr, state.t_prev, state.p_prev, state.history};
      on_start_mut(m);
    }
    fire(events_on_step, state);
    while (state.t_curr < t_end) {
      state.make_step(rhs, std::fmin(stepsize.get(), t_end - state.t_curr));
      double h_ctl = stepsize.get();
      while (stepsize.update(state.e_curr) == StepStatus::Rejected) {
        ++n_rejected;
        state.undo_step();
        // Not a reference behaviour: when a rejection leaves the controller's stepsize unchanged
        // (clamped at stepsize_range.start) the retry is the same step from the same state with the
        // same error, and the reference's loop never ends.  The CUDA path stops such a trajectory
        // with DFB_TRAJ_STEPSIZE_FLOOR; so does this checker.
        if (stepsize.get() == h_ctl) throw Panic{TRAJ_STEPSIZE_FLOOR};
        h_ctl = stepsize.get();
        state.make_step(rhs, std::fmin(stepsize.get(), t_end - state.t_curr));
      }
      size_t index;
      double time;
      if (locate_earliest(state, &index, &time) && time >= state.t_prev) {
        state.undo_step();
        state.make_step(rhs, time - state.t_curr);
        state.commit_step();
        state.make_zero_step();
        events_on_loc[index]->eval_mut(state);
        ++n_events;
        if (log_events) event_log.emplace_back(time, int(index));
      }
      state.commit_step();
      ++n_steps;
      fire(events_on_step, state);
      if (max_steps && n_steps >= max_steps && state.t_curr < t_end) throw Panic{TRAJ_STEP_LIMIT};
    }
    fire(events_on_stop, state);
    return state;
  }
};

}  // namespace orc
