// oracle/refbuild/include/boost/numeric/odeint.hpp — TEST INFRASTRUCTURE (oracle/_ref build): stand-in for Boost.odeint
// (CRAN package BH, unpinned in the reference's DESCRIPTION:15-16), restated from its published algorithms for exactly the
// three entry points the reference calls (R/modelTemplate.R:321-342) — SURVEY §8c rules T1 and T2:
//   integrate_const(runge_kutta4<State>(), sys, x, t0, t1, dt)   fixed-step classical RK4; steps while (time + dt) - t1 <= eps,
//                                                               time = t0 + step*dt, the remainder interval is NOT integrated
//   integrate(sys, x, t0, t1, dt) = integrate_adaptive(controlled_runge_kutta<runge_kutta_dopri5<State>>(), ...) with
//                                   eps_abs = eps_rel = 1e-6, a_x = a_dxdt = 1, landing exactly on t1
// plus the customisation points the reference specialises for arma::mat (is_resizeable, same_size_impl, resize_impl;
// R/modelTemplate.R:20-48).  State algebra: element-wise over begin()/end(), like odeint's default range_algebra.
#pragma once
#include <cfloat>
#include <cmath>
#include <stdexcept>

namespace boost {
struct true_type { static const bool value = true; typedef true_type type; };
struct false_type { static const bool value = false; typedef false_type type; };
namespace numeric { namespace odeint {

template <class State> struct is_resizeable { typedef boost::false_type type; const static bool value = type::value; };
template <class S1, class S2> struct same_size_impl { static bool same_size(const S1&, const S2&) { return true; } };
template <class S1, class S2> struct resize_impl { static void resize(S1&, const S2&) {} };

namespace detail {
template <class State> void adjust(State& tmp, const State& x) {  // resizer: temporaries take the shape of the state
  if (is_resizeable<State>::value && !same_size_impl<State, State>::same_size(tmp, x)) resize_impl<State, State>::resize(tmp, x);
}
}  // namespace detail

template <class State>
class runge_kutta4 {
  State k1, k2, k3, k4, xt;
 public:
  template <class System>
  void do_step(System& sys, State& x, double t, double dt) {
    detail::adjust(k1, x); detail::adjust(k2, x); detail::adjust(k3, x); detail::adjust(k4, x); detail::adjust(xt, x);
    const double dh = dt / 2, dt6 = dt / 6, dt3 = dt / 3;
    sys(x, k1, t);
    { auto a = xt.begin(); auto b = x.begin(); auto c = k1.begin(); for (; a != xt.end(); ++a, ++b, ++c) *a = *b + dh * *c; }
    sys(xt, k2, t + dh);
    { auto a = xt.begin(); auto b = x.begin(); auto c = k2.begin(); for (; a != xt.end(); ++a, ++b, ++c) *a = *b + dh * *c; }
    sys(xt, k3, t + dh);
    { auto a = xt.begin(); auto b = x.begin(); auto c = k3.begin(); for (; a != xt.end(); ++a, ++b, ++c) *a = *b + dt * *c; }
    sys(xt, k4, t + dt);
    { auto b = x.begin(); auto c1 = k1.begin(); auto c2 = k2.begin(); auto c3 = k3.begin(); auto c4 = k4.begin();
      for (; b != x.end(); ++b, ++c1, ++c2, ++c3, ++c4) *b = *b + dt6 * *c1 + dt3 * *c2 + dt3 * *c3 + dt6 * *c4; }
  }
};

// integrate_const for a plain stepper (odeint/integrate/detail/integrate_const.hpp, stepper_tag overload)
template <class Stepper, class System, class State>
size_t integrate_const(Stepper stepper, System system, State& x, double t0, double t1, double dt) {
  double time = t0;
  int step = 0;
  while ((time + dt) - t1 <= DBL_EPSILON) {   // less_eq_with_sign(time + dt, t1, dt) for dt > 0
    stepper.do_step(system, x, time, dt);
    ++step;
    time = t0 + static_cast<double>(step) * dt;
  }
  return (size_t)step;
}

// integrate(): controlled Dormand-Prince 5(4) with FSAL, default error checker and step adjuster
template <class System, class State>
size_t integrate(System system, State& x, double t0, double t1, double dt) {
  const double a21 = 1.0 / 5, a31 = 3.0 / 40, a32 = 9.0 / 40, a41 = 44.0 / 45, a42 = -56.0 / 15, a43 = 32.0 / 9,
               a51 = 19372.0 / 6561, a52 = -25360.0 / 2187, a53 = 64448.0 / 6561, a54 = -212.0 / 729,
               a61 = 9017.0 / 3168, a62 = -355.0 / 33, a63 = 46732.0 / 5247, a64 = 49.0 / 176, a65 = -5103.0 / 18656,
               c1 = 35.0 / 384, c3 = 500.0 / 1113, c4 = 125.0 / 192, c5 = -2187.0 / 6784, c6 = 11.0 / 84;
  const double dc1 = c1 - 5179.0 / 57600, dc3 = c3 - 7571.0 / 16695, dc4 = c4 - 393.0 / 640, dc5 = c5 - (-92097.0 / 339200),
               dc6 = c6 - 187.0 / 2100, dc7 = -1.0 / 40;
  const double eps_abs = 1e-6, eps_rel = 1e-6;
  State k1, k2, k3, k4, k5, k6, k7, xt, xn, xerr;
  for (State* s : {&k1, &k2, &k3, &k4, &k5, &k6, &k7, &xt, &xn, &xerr}) detail::adjust(*s, x);
  const size_t n = (size_t)(x.end() - x.begin());
  auto at = [](State& s, size_t i) -> double& { return *(s.begin() + i); };
  double t = t0;
  size_t steps = 0;
  system(x, k1, t);   // a fresh stepper: the FSAL derivative is computed at the start of the interval
  while (t1 - t > DBL_EPSILON) {               // less_with_sign(t, t1, dt)
    if ((t + dt) - t1 > DBL_EPSILON) dt = t1 - t;
    int trials = 0;
    for (;;) {
      if (++trials > 500) throw std::overflow_error("Max number of iterations exceeded (500). A new step size was not found.");
      for (size_t i = 0; i < n; i++) at(xt, i) = at(x, i) + dt * a21 * at(k1, i);
      system(xt, k2, t + dt * (1.0 / 5));
      for (size_t i = 0; i < n; i++) at(xt, i) = at(x, i) + dt * a31 * at(k1, i) + dt * a32 * at(k2, i);
      system(xt, k3, t + dt * (3.0 / 10));
      for (size_t i = 0; i < n; i++) at(xt, i) = at(x, i) + dt * a41 * at(k1, i) + dt * a42 * at(k2, i) + dt * a43 * at(k3, i);
      system(xt, k4, t + dt * (4.0 / 5));
      for (size_t i = 0; i < n; i++) at(xt, i) = at(x, i) + dt * a51 * at(k1, i) + dt * a52 * at(k2, i) + dt * a53 * at(k3, i) + dt * a54 * at(k4, i);
      system(xt, k5, t + dt * (8.0 / 9));
      for (size_t i = 0; i < n; i++) at(xt, i) = at(x, i) + dt * a61 * at(k1, i) + dt * a62 * at(k2, i) + dt * a63 * at(k3, i) + dt * a64 * at(k4, i) + dt * a65 * at(k5, i);
      system(xt, k6, t + dt);
      for (size_t i = 0; i < n; i++) at(xn, i) = at(x, i) + dt * c1 * at(k1, i) + dt * c3 * at(k3, i) + dt * c4 * at(k4, i) + dt * c5 * at(k5, i) + dt * c6 * at(k6, i);
      system(xn, k7, t + dt);
      double err = 0.0;
      for (size_t i = 0; i < n; i++) {
        at(xerr, i) = dt * dc1 * at(k1, i) + dt * dc3 * at(k3, i) + dt * dc4 * at(k4, i) + dt * dc5 * at(k5, i) + dt * dc6 * at(k6, i) + dt * dc7 * at(k7, i);
        const double e = std::fabs(at(xerr, i)) / (eps_abs + eps_rel * (std::fabs(at(x, i)) + std::fabs(dt) * std::fabs(at(k1, i))));
        err = (e > err || e != e) ? e : err;   // norm_inf; a NaN propagates
      }
      if (err != err) {   // odeint accepts such a step (NaN > 1 is false) and carries the NaN state to t1: same outcome, decided here
        for (size_t i = 0; i < n; i++) at(x, i) = err;
        return steps;
      }
      if (err > 1.0) {
        dt *= std::max(0.9 * std::pow(err, -1.0 / 3.0), 0.2);   // -1/(error_order - 1), error_order = 4
        continue;
      }
      t += dt;
      if (err < 0.5) {
        err = std::max(std::pow(5.0, -5.0), err);               // stepper_order = 5
        dt *= 0.9 * std::pow(err, -1.0 / 5.0);
      }
      for (size_t i = 0; i < n; i++) { at(x, i) = at(xn, i); at(k1, i) = at(k7, i); }
      ++steps;
      break;
    }
  }
  return steps;
}

}}}  // namespace boost::numeric::odeint
