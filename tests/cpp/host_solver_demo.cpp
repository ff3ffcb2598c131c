// Exercises the C++ mirror of the Solver builder (diffurch_b200/csrc/host/solver.hpp)
// through the C ABI.  Prints one line per case: name followed by the values the Python
// test compares with the ctypes path.  Exit code 3 = no CUDA device (expected on CPU).
#include <cstdio>
#include <vector>

#include "../../diffurch_b200/csrc/host/solver.hpp"

int main() {
  try {
    // examples/ode-chaos-lorenz.rs:17-24, three members of a rho sweep, t in [0, 1]
    {
      std::vector<double> y0, prm;
      for (int i = 0; i < 3; ++i) {
        y0.insert(y0.end(), {1.0, 2.0, 20.0});
        prm.insert(prm.end(), {10.0, 20.0 + 10.0 * i, 8.0 / 3.0});
      }
      auto st = dfb::Solver::make().initial(y0).equation(DFB_SYS_LORENZ, prm).interval(0.0, 1.0)
                    .stepsize(1.0 / 250.0).run();
      std::printf("lorenz");
      for (size_t i = 0; i < st.n; ++i)
        std::printf(" %llu %.17g %.17g %.17g", (unsigned long long)st.steps[i], st.p[3 * i], st.p[3 * i + 1], st.p[3 * i + 2]);
      std::printf("\n");
    }
    // examples/bouncing-ball.rs:21-33
    {
      auto st = dfb::Solver::make().initial({1.0, -0.01}).equation(DFB_SYS_BOUNCING_BALL, {9.8, 0.9})
                    .interval(0.0, 4.0).stepsize(0.005).log_events(64)
                    .on(dfb::Locator::zero(1)).on_mut(dfb::Locator::below_zero(2), 1).run();
      std::printf("ball %llu %llu %.17g %.17g", (unsigned long long)st.steps[0], (unsigned long long)st.events[0],
                  st.p[0], st.p[1]);
      for (uint32_t e = 0; e < st.event_n[0] && e < 6; ++e) std::printf(" %.17g", st.event_t[e]);
      std::printf("\n");
    }
    // tests/delay_propagation.rs:6-25
    {
      auto st = dfb::Solver::make().rk(dfb::RK::euler()).initial({0.0}).equation(DFB_SYS_ZERO1)
                    .initial_disco({{0.0, 0}}).with_const_delay(1.0, 0).interval(0.0, 5.0).stepsize(0.75)
                    .on_step(dfb::Trace{1, 64}).run();
      std::printf("delay");
      for (uint32_t i = 0; i < st.trace_n[0]; ++i) std::printf(" %.17g", st.trace_t[i]);
      std::printf("\n");
    }
  } catch (const dfb::Error& e) {
    std::fprintf(stderr, "%s\n", e.what());
    return e.code == DFB_ERR_NO_DEVICE ? 3 : 1;
  }
  return 0;
}
