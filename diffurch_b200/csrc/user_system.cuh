// What a user-system header includes.  Example (tests/plugins/van_der_pol.cuh):
//
//   #include "user_system.cuh"
//   namespace DFB_NS {
//   struct SysVanDerPol : SysBase {          // x'' - mu (1 - x^2) x' + x = 0
//     static constexpr int N = 2, NP = 1;    // state size, parameters per trajectory
//     template <class V>                     // V = StateRef: {t, p, d, t_prev, p_prev, p_at(), d_at()}
//     __device__ __forceinline__ static void rhs(const V& s, const double* prm, double* out) {
//       out[0] = s.p[1];
//       out[1] = prm[0] * (1.0 - s.p[0] * s.p[0]) * s.p[1] - s.p[0];
//     }
//   };
//   }  // namespace DFB_NS
//
// A callback may also WRITE the parameter block (`callback(int id, M& s, double* prm, double* aux)`):
// that is how a parameter captured in a `Cell` by both the equation and a callback upstream is
// restated (tests/plugins/switch_parameter.cuh).  Discrete mode variables of `#[state(discrete)]`
// (macros/derive_state) are state components with a zero derivative, or aux slots.
//
// Optional members (defaults in SysBase): NA (aux accumulators), ND / NC (detector / callback
// counts), `detector(id, view, prm)`, `callback(id, mut_view, prm, aux)`, `uses_history` +
// `has_const_delay` + `const_delay(prm)` for delay equations, `pure_rhs = false` if rhs reads
// p_prev / t_prev / d / the history.
//   * `reads_d = true` if ANY functor (rhs, detector, callback) reads the view's `d` (state_fn.rs:17).
//     During the stages the reference's `d` is STALE (the derivative at the end of the previous step);
//     the register-resident kernels (ode_fixed / periodic / event / adaptive / dde_event) carry the
//     fresh k[0] there, so a system that reads `d` must say so and is then integrated by the general
//     kernel, which keeps the reference's stale value.  Leaving it false while reading `s.d` silently
//     diverges from the reference.
//   * `pure_detectors = true` if every detector depends on (t, p, d, history) only — not on t_prev /
//     p_prev: dde_event_kernel then reuses a detector's value at a step end as the next eval_prev.
// The header is compiled twice, into namespaces dfb_strict (-fmad=false) and dfb_fast; DFB_NS names
// the current one.  The same text can be registered at run time without nvcc
// (dfb_register_system_from_source, NVRTC).
#pragma once
#include "dev/systems.cuh"
