// =============================================================================
// oracle_tableaux.hpp — TEST INFRASTRUCTURE ONLY (see diffurch_oracle.hpp).
//
// The named Butcher tableaux of src/rk.rs:55-430, restated.  The parametric
// families are evaluated with the same f64 expressions as the reference
// (rk.rs:79-98, 116-203); rktp64 carries the reference's decimal digits
// (rk.rs:313-429).  tests/test_oracle_tableaux.py checks every entry bit for
// bit against tests/golden/tableaux.json (generated from the reference source
// by tests/golden/make_tableaux_golden.py) and re-creates the commented-out
// consistency checks of rk.rs:432-640.
// =============================================================================
#pragma once
#include <cstring>

#include "diffurch_oracle.hpp"

namespace orc {

inline ButcherTableu<1, 2> euler() {  // rk.rs:60-71
  ButcherTableu<1, 2> t;
  t.order = 1;
  t.order_embedded = 0;
  t.order_interpolant = 1;
  t.a[0][0] = 0.0;
  t.b[0] = 1.0;
  t.b2[0] = 0.0;
  t.c[0] = 0.0;
  t.bi[0][0] = 0.0;
  t.bi[0][1] = 1.0;
  return t;
}

inline ButcherTableu<2, 2> generic_order_2(double c2) {  // rk.rs:79-98
  ButcherTableu<2, 2> t;
  t.order = 2;
  t.order_embedded = 1;
  t.order_interpolant = 1;
  t.a[1][0] = c2;
  t.b[0] = 1. - 1. / (2. * c2);
  t.b[1] = 1. / (2. * c2);
  t.b2[0] = 1.;
  t.b2[1] = 0.;
  t.c[0] = 0.;
  t.c[1] = c2;
  t.bi[0][1] = t.b[0];
  t.bi[1][1] = t.b[1];
  return t;
}
inline ButcherTableu<2, 2> midpoint() { return generic_order_2(0.5); }
inline ButcherTableu<2, 2> heun2() { return generic_order_2(1.); }
inline ButcherTableu<2, 2> ralston2() { return generic_order_2(2. / 3.); }

// rk.rs:116-203.  Returns false where the reference panics (rk.rs:200).
inline bool generic_order_3(double c2, double c3, bool has_b3, double b3,
                            ButcherTableu<3, 2>* out) {
  ButcherTableu<3, 2> t;
  t.order = 2;  // sic: rk.rs:138,164,190 report order 2 for the 3rd-order family
  t.order_embedded = 1;
  t.order_interpolant = 1;
  if (c2 != 0. && c3 != 0. && c2 != c3 && c2 != 2. / 3. && !has_b3) {  // case I
    t.a[1][0] = c2;
    t.a[2][0] = (c3 * (3. * c2 * (1. - c2) - c3)) / (c2 * (2. - 3. * c2));
    t.a[2][1] = (c3 * (c3 - c2)) / (c2 * (2. - 3. * c2));
    t.b[0] = 1. - (1. * c2 + 3. * c3 - 2.) / (6. * c2 * c3);
    t.b[1] = (3. * c3 - 2.) / (6. * c2 * (c3 - c2));
    t.b[2] = (2. - 3. * c2) / (6. * c3 * (c3 - c2));
  } else if (c2 == 2. / 3. && c3 == 2. / 3. && has_b3 && b3 != 0.) {  // case II
    t.a[1][0] = c2;
    t.a[2][0] = 2. / 3. - 1. / (4. * b3);
    t.a[2][1] = 1. / (4. * b3);
    t.b[0] = 1. / 4.;
    t.b[1] = 3. / 4. - b3;
    t.b[2] = b3;
  } else if (c2 == 2. / 3. && c3 == 0. && has_b3 && b3 != 0.) {  // case III
    t.a[1][0] = c2;
    t.a[2][0] = -1. / (4. * b3);
    t.a[2][1] = 1. / (4. * b3);
    t.b[0] = 1. / 4. - b3;
    t.b[1] = 3. / 4.;
    t.b[2] = b3;
  } else {
    return false;
  }
  t.b2[0] = 1. - 1. / (2. * c2);
  t.b2[1] = 1. / (2. * c2);
  t.b2[2] = 0.;
  t.c[0] = 0.;
  t.c[1] = c2;
  t.c[2] = c3;
  for (int i = 0; i < 3; ++i) t.bi[i][1] = t.b[i];
  *out = t;
  return true;
}
inline ButcherTableu<3, 2> go3(double c2, double c3) {
  ButcherTableu<3, 2> t;
  generic_order_3(c2, c3, false, 0., &t);
  return t;
}
inline ButcherTableu<3, 2> kutta3() { return go3(0.5, 1.); }
inline ButcherTableu<3, 2> heun3() { return go3(1. / 3., 2. / 3.); }
inline ButcherTableu<3, 2> ralston3() { return go3(1. / 2., 3. / 4.); }
inline ButcherTableu<3, 2> wray3() { return go3(8. / 15., 2. / 3.); }
inline ButcherTableu<3, 2> ssp3() { return go3(1.0, 0.5); }

inline ButcherTableu<4, 2> rk4() {  // rk.rs:226-248
  ButcherTableu<4, 2> t;
  t.order = 4;
  t.order_embedded = 2;
  t.order_interpolant = 1;
  t.a[1][0] = 0.5;
  t.a[2][1] = 0.5;
  t.a[3][2] = 1.;
  const double b[4] = {1. / 6., 1. / 3., 1. / 3., 1. / 6.};
  const double b2[4] = {0., 1., 0., 0.};
  const double c[4] = {0., 0.5, 0.5, 1.};
  for (int i = 0; i < 4; ++i) {
    t.b[i] = b[i];
    t.b2[i] = b2[i];
    t.c[i] = c[i];
    t.bi[i][1] = b[i];
  }
  return t;
}

inline ButcherTableu<4, 2> three_eights() {  // rk.rs:251-273
  ButcherTableu<4, 2> t;
  t.order = 4;
  t.order_embedded = 2;
  t.order_interpolant = 1;
  t.a[1][0] = 1. / 3.;
  t.a[2][0] = -1. / 3.;
  t.a[2][1] = 1.;
  t.a[3][0] = 1.;
  t.a[3][1] = -1.;
  t.a[3][2] = 1.;
  const double b[4] = {1. / 8., 3. / 8., 3. / 8., 1. / 8.};
  const double b2[4] = {-0.5, 1.5, 0., 0.};
  const double c[4] = {0., 1. / 3., 2. / 3., 1.};
  for (int i = 0; i < 4; ++i) {
    t.b[i] = b[i];
    t.b2[i] = b2[i];
    t.c[i] = c[i];
    t.bi[i][1] = b[i];
  }
  return t;
}

inline ButcherTableu<5, 4> rk43() {  // rk.rs:280-307
  ButcherTableu<5, 4> t;
  t.order = 4;
  t.order_embedded = 3;
  t.order_interpolant = 3;
  t.a[1][0] = 0.5;
  t.a[2][1] = 0.5;
  t.a[3][2] = 1.;
  t.a[4][0] = 5. / 32.;
  t.a[4][1] = 7. / 32.;
  t.a[4][2] = 13. / 32.;
  t.a[4][3] = -1. / 32.;
  const double b[5] = {1. / 6., 1. / 3., 1. / 3., 1. / 6., 0.};
  const double b2[5] = {-1. / 2., 7. / 3., 7. / 3., 13. / 6., -16. / 3.};
  const double c[5] = {0., 0.5, 0.5, 1., 0.75};
  const double bi[5][4] = {{0., 1., -1.5, 2. / 3.},
                           {0., 0., 1., -2. / 3.},
                           {0., 0., 1., -2. / 3.},
                           {0., 0., -0.5, 2. / 3.},
                           {0., 0., 0., 0.}};
  for (int i = 0; i < 5; ++i) {
    t.b[i] = b[i];
    t.b2[i] = b2[i];
    t.c[i] = c[i];
    for (int j = 0; j < 4; ++j) t.bi[i][j] = bi[i][j];
  }
  return t;
}

inline ButcherTableu<7, 5> rktp64() {  // rk.rs:313-429 (the Solver::new default)
  ButcherTableu<7, 5> t;
  t.order = 6;
  t.order_embedded = 4;
  t.order_interpolant = 4;
  t.a[1][0] = 0.148148148148148148148148148148148;
  t.a[2][0] = 0.0555555555555555555555555555555556;
  t.a[2][1] = 0.166666666666666666666666666666667;
  t.a[3][0] = 0.192419825072886297376093294460641;
  t.a[3][1] = -0.531341107871720116618075801749271;
  t.a[3][2] = 0.767492711370262390670553935860058;
  t.a[4][0] = 0.271382649739583333333333333333333;
  t.a[4][1] = -0.28179931640625;
  t.a[4][2] = 0.101919320913461538461538461538462;
  t.a[4][3] = 0.595997345753205128205128205128205;
  t.a[5][0] = -0.121406813486922726795280277301215;
  t.a[5][1] = 0.477614101874456904042702859270907;
  t.a[5][2] = 0.121922969684790809202712084577398;
  t.a[5][3] = 0.00820786686248269381285345905535723;
  t.a[5][4] = 0.282892644295961550506242643628322;
  t.a[6][0] = 0.323109465891062929666842940243258;
  t.a[6][1] = -0.610391327340031729243786356425172;
  t.a[6][2] = 0.458468675416393199766128880472551;
  t.a[6][3] = 0.575057408067115662781332786609229;
  t.a[6][4] = -0.573792345222676817817456258459895;
  t.a[6][5] = 0.827548123188136754846938007560029;
  const double b[7] = {0.0727777777777777777777777777777778,
                       0.0,
                       0.287521270706905035263244218468099,
                       0.189748462203968321877109418822433,
                       0.105817363486825507351679735198248,
                       0.269095443284840818047649167193759,
                       0.0750396825396825396825396825396825};
  const double b2[7] = {0.103226660475181185240356837989974,
                        0.0,
                        0.156115420561340710254765377422461,
                        0.386349188510639078027209177837778,
                        -0.120730952086843513204871075789895,
                        0.4,
                        0.0750396825396825396825396825396825};
  const double c[7] = {0.0,
                       0.148148148148148148148148148148148,
                       0.222222222222222222222222222222222,
                       0.428571428571428571428571428571429,
                       0.6875,
                       0.769230769230769230769230769230769,
                       1.0};
  const double bi[7][5] = {
      {0.0, 1.0, -3.18888888888888888888888888888889, 3.6296296296296296296296296296296296,
       -1.36796296296296296296296296296296},
      {0.0, 0.0, 0.0, 0.0, 0.0},
      {0.0, 0.0, 3.1366409709693291782844021649991799, -4.8756765622437264228309004428407413,
       2.0265568619813022798097424963096605},
      {0.0, 0.0, 1.4430666077217801355732390215148836, -2.6873219373219373219373219373219373,
       1.4340037918041255082411923346294870},
      {0.0, 0.0, -1.8410567850403156630639903928632699, 5.4341625207296849087893864013266998,
       -3.4872883722025437383737162732651819},
      {0.0, 0.0, 0.0, 0.0, 0.269095443284840818047649167193759},
      {0.0, 0.0, 0.450238095238095238095238095238095, -1.50079365079365079365079365079365,
       1.12559523809523809523809523809524}};
  for (int i = 0; i < 7; ++i) {
    t.b[i] = b[i];
    t.b2[i] = b2[i];
    t.c[i] = c[i];
    for (int j = 0; j < 5; ++j) t.bi[i][j] = bi[i][j];
  }
  return t;
}

}  // namespace orc
