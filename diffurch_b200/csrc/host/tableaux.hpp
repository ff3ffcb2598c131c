// Host-side Butcher tableaux of the product: the named constructors of the
// reference's `ButcherTableu` (src/rk.rs:55-430) flattened into the C-ABI
// `dfb_tableau` POD.  Entries are checked bit for bit against
// tests/golden/tableaux.json (generated from the reference source).
#pragma once
#include <cstring>
#include <string>

#include "../../../include/diffurch_b200.h"

namespace dfb_host {

struct TabBuilder {
  dfb_tableau t;
  TabBuilder(int S, int I, int order, int emb, int interp) {
    std::memset(&t, 0, sizeof(t));
    t.S = S;
    t.I = I;
    t.order = order;
    t.order_embedded = emb;
    t.order_interpolant = interp;
  }
  double& a(int i, int j) { return t.a[i * DFB_MAX_STAGES + j]; }
  double& bi(int i, int j) { return t.bi[i * DFB_MAX_INTERP + j]; }
  // all tableaux but rk43/rktp64 interpolate linearly: bi = [0, b_i] (rk.rs:84,134,237)
  void linear_interpolant() {
    for (int i = 0; i < t.S; ++i) bi(i, 1) = t.b[i];
  }
};

inline dfb_tableau tab_euler() {  // rk.rs:60
  TabBuilder B(1, 2, 1, 0, 1);
  B.t.b[0] = 1.0;
  B.linear_interpolant();
  return B.t;
}

inline dfb_tableau tab_generic_order_2(double c2) {  // rk.rs:79
  TabBuilder B(2, 2, 2, 1, 1);
  B.a(1, 0) = c2;
  B.t.b[0] = 1. - 1. / (2. * c2);
  B.t.b[1] = 1. / (2. * c2);
  B.t.b2[0] = 1.;
  B.t.c[1] = c2;
  B.linear_interpolant();
  return B.t;
}

// rk.rs:116-203; false where the reference panics ("Provided arguments are incorrect.")
inline bool tab_generic_order_3(double c2, double c3, bool has_b3, double b3, dfb_tableau* out) {
  TabBuilder B(3, 2, /*order (sic, rk.rs:138)*/ 2, 1, 1);
  const double two_thirds = 2. / 3.;
  if (c2 != 0. && c3 != 0. && c2 != c3 && c2 != two_thirds && !has_b3) {
    const double den = c2 * (2. - 3. * c2);
    B.a(2, 0) = (c3 * (3. * c2 * (1. - c2) - c3)) / den;
    B.a(2, 1) = (c3 * (c3 - c2)) / den;
    B.t.b[0] = 1. - (1. * c2 + 3. * c3 - 2.) / (6. * c2 * c3);
    B.t.b[1] = (3. * c3 - 2.) / (6. * c2 * (c3 - c2));
    B.t.b[2] = (2. - 3. * c2) / (6. * c3 * (c3 - c2));
  } else if (c2 == two_thirds && c3 == two_thirds && has_b3 && b3 != 0.) {
    B.a(2, 0) = two_thirds - 1. / (4. * b3);
    B.a(2, 1) = 1. / (4. * b3);
    B.t.b[0] = 1. / 4.;
    B.t.b[1] = 3. / 4. - b3;
    B.t.b[2] = b3;
  } else if (c2 == two_thirds && c3 == 0. && has_b3 && b3 != 0.) {
    B.a(2, 0) = -1. / (4. * b3);
    B.a(2, 1) = 1. / (4. * b3);
    B.t.b[0] = 1. / 4. - b3;
    B.t.b[1] = 3. / 4.;
    B.t.b[2] = b3;
  } else {
    return false;
  }
  B.a(1, 0) = c2;
  B.t.b2[0] = 1. - 1. / (2. * c2);
  B.t.b2[1] = 1. / (2. * c2);
  B.t.c[1] = c2;
  B.t.c[2] = c3;
  B.linear_interpolant();
  *out = B.t;
  return true;
}

inline dfb_tableau tab_rk4_family(bool three_eights) {  // rk.rs:226, 251
  TabBuilder B(4, 2, 4, 2, 1);
  if (!three_eights) {
    B.a(1, 0) = 0.5;
    B.a(2, 1) = 0.5;
    B.a(3, 2) = 1.;
    const double b[4] = {1. / 6., 1. / 3., 1. / 3., 1. / 6.};
    const double c[4] = {0., 0.5, 0.5, 1.};
    for (int i = 0; i < 4; ++i) B.t.b[i] = b[i], B.t.c[i] = c[i];
    B.t.b2[1] = 1.;
  } else {
    B.a(1, 0) = 1. / 3.;
    B.a(2, 0) = -1. / 3.;
    B.a(2, 1) = 1.;
    B.a(3, 0) = 1.;
    B.a(3, 1) = -1.;
    B.a(3, 2) = 1.;
    const double b[4] = {1. / 8., 3. / 8., 3. / 8., 1. / 8.};
    const double c[4] = {0., 1. / 3., 2. / 3., 1.};
    for (int i = 0; i < 4; ++i) B.t.b[i] = b[i], B.t.c[i] = c[i];
    B.t.b2[0] = -0.5;
    B.t.b2[1] = 1.5;
  }
  B.linear_interpolant();
  return B.t;
}

inline dfb_tableau tab_rk43() {  // rk.rs:280
  TabBuilder B(5, 4, 4, 3, 3);
  B.a(1, 0) = 0.5;
  B.a(2, 1) = 0.5;
  B.a(3, 2) = 1.;
  const double a4[4] = {5. / 32., 7. / 32., 13. / 32., -1. / 32.};
  for (int j = 0; j < 4; ++j) B.a(4, j) = a4[j];
  const double b[5] = {1. / 6., 1. / 3., 1. / 3., 1. / 6., 0.};
  const double b2[5] = {-1. / 2., 7. / 3., 7. / 3., 13. / 6., -16. / 3.};
  const double c[5] = {0., 0.5, 0.5, 1., 0.75};
  for (int i = 0; i < 5; ++i) B.t.b[i] = b[i], B.t.b2[i] = b2[i], B.t.c[i] = c[i];
  B.bi(0, 1) = 1.;
  B.bi(0, 2) = -1.5;
  B.bi(0, 3) = 2. / 3.;
  B.bi(1, 2) = 1.;
  B.bi(1, 3) = -2. / 3.;
  B.bi(2, 2) = 1.;
  B.bi(2, 3) = -2. / 3.;
  B.bi(3, 2) = -0.5;
  B.bi(3, 3) = 2. / 3.;
  return B.t;
}

// rk.rs:313-429 — the Solver::new default (solver.rs:88).  Strictly-lower
// triangle of `a` packed row by row; 25 significant digits are enough for the
// nearest double (checked against the golden file).
inline dfb_tableau tab_rktp64() {
  TabBuilder B(7, 5, 6, 4, 4);
  static const double a_packed[21] = {
      1.481481481481481481481481e-1,
      5.555555555555555555555556e-2,  1.666666666666666666666667e-1,
      1.924198250728862973760933e-1,  -5.313411078717201166180758e-1, 7.674927113702623906705539e-1,
      2.713826497395833333333333e-1,  -2.8179931640625e-1,            1.019193209134615384615385e-1,
      5.959973457532051282051282e-1,
      -1.214068134869227267952803e-1, 4.776141018744569040427029e-1,  1.219229696847908092027121e-1,
      8.207866862482693812853459e-3,  2.828926442959615505062426e-1,
      3.231094658910629296668429e-1,  -6.103913273400317292437864e-1, 4.584686754163931997661289e-1,
      5.750574080671156627813328e-1,  -5.737923452226768178174563e-1, 8.275481231881367548469380e-1};
  int q = 0;
  for (int i = 1; i < 7; ++i)
    for (int j = 0; j < i; ++j) B.a(i, j) = a_packed[q++];
  static const double b[7] = {7.277777777777777777777778e-2, 0.0,
                              2.875212707069050352632442e-1, 1.897484622039683218771094e-1,
                              1.058173634868255073516797e-1, 2.690954432848408180476492e-1,
                              7.503968253968253968253968e-2};
  static const double b2[7] = {1.032266604751811852403568e-1, 0.0,
                               1.561154205613407102547654e-1, 3.863491885106390780272092e-1,
                               -1.207309520868435132048711e-1, 4.0e-1,
                               7.503968253968253968253968e-2};
  static const double c[7] = {0.0,
                              1.481481481481481481481481e-1,
                              2.222222222222222222222222e-1,
                              4.285714285714285714285714e-1,
                              6.875e-1,
                              7.692307692307692307692308e-1,
                              1.0};
  for (int i = 0; i < 7; ++i) B.t.b[i] = b[i], B.t.b2[i] = b2[i], B.t.c[i] = c[i];
  // continuous extension b_i(theta) = sum_j bi[i][j] theta^j; row 1 is all zero
  static const double bi0[4] = {1.0, -3.188888888888888888888889, 3.629629629629629629629630,
                                -1.367962962962962962962963};
  static const double bi2[3] = {3.136640970969329178284402, -4.875676562243726422830900,
                                2.026556861981302279809742};
  static const double bi3[3] = {1.443066607721780135573239, -2.687321937321937321937322,
                                1.434003791804125508241192};
  static const double bi4[3] = {-1.841056785040315663063990, 5.434162520729684908789386,
                                -3.487288372202543738373716};
  static const double bi6[3] = {4.502380952380952380952381e-1, -1.500793650793650793650794,
                                1.125595238095238095238095};
  for (int j = 0; j < 4; ++j) B.bi(0, j + 1) = bi0[j];
  for (int j = 0; j < 3; ++j) {
    B.bi(2, j + 2) = bi2[j];
    B.bi(3, j + 2) = bi3[j];
    B.bi(4, j + 2) = bi4[j];
    B.bi(6, j + 2) = bi6[j];
  }
  B.bi(5, 4) = 2.690954432848408180476492e-1;
  return B.t;
}

inline bool tab_by_name(const std::string& name, dfb_tableau* out) {
  dfb_tableau t;
  bool ok = true;
  if (name == "euler") t = tab_euler();
  else if (name == "midpoint") t = tab_generic_order_2(0.5);
  else if (name == "heun2") t = tab_generic_order_2(1.);
  else if (name == "ralston2") t = tab_generic_order_2(2. / 3.);
  else if (name == "kutta3") ok = tab_generic_order_3(0.5, 1., false, 0., &t);
  else if (name == "heun3") ok = tab_generic_order_3(1. / 3., 2. / 3., false, 0., &t);
  else if (name == "ralston3") ok = tab_generic_order_3(1. / 2., 3. / 4., false, 0., &t);
  else if (name == "wray3") ok = tab_generic_order_3(8. / 15., 2. / 3., false, 0., &t);
  else if (name == "ssp3") ok = tab_generic_order_3(1.0, 0.5, false, 0., &t);
  else if (name == "rk4") t = tab_rk4_family(false);
  else if (name == "three_eights") t = tab_rk4_family(true);
  else if (name == "rk43") t = tab_rk43();
  else if (name == "rktp64") t = tab_rktp64();
  else return false;
  if (ok) *out = t;
  return ok;
}

}  // namespace dfb_host
