// op_counter.h — TEST INFRASTRUCTURE (tests/hostsim with -DPSY_COUNT_OPS): replaces `double` in the kernel source by a wrapper
// that counts the floating-point operations the SOURCE executes (a*b+c written out counts as mul + add = 2, fma() as 2, one per
// add / mul / div / sqrt / exp / log / pow / comparison-free ops), so psydiff_b200/flops.py's closed-form count can be checked
// against a dynamic count of the real code path.  Included by cuda_host_shim.h AFTER every system header.
#pragma once
#include <type_traits>
// arithmetic inside constant expressions (the constexpr tableau of dopri5) is not counted — it is not executed
#define PSY_TICK(c) do { if (!std::is_constant_evaluated()) psy_ops::c++; } while (0)

struct psy_counted;
namespace psy_ops { extern thread_local unsigned long long add, mul, fma_, div_, special; }
namespace psy_ops { thread_local unsigned long long add = 0, mul = 0, fma_ = 0, div_ = 0, special = 0; }

struct psy_counted {
  double v;
  psy_counted() = default;
  constexpr psy_counted(double x) : v(x) {}
  constexpr psy_counted(int x) : v(x) {}
  constexpr psy_counted(long long x) : v((double)x) {}
  explicit operator double() const { return v; }
  explicit operator int() const { return (int)v; }
  psy_counted& operator+=(psy_counted o) { psy_ops::add++; v += o.v; return *this; }
  psy_counted& operator-=(psy_counted o) { psy_ops::add++; v -= o.v; return *this; }
  psy_counted& operator*=(psy_counted o) { psy_ops::mul++; v *= o.v; return *this; }
  psy_counted& operator/=(psy_counted o) { psy_ops::div_++; v /= o.v; return *this; }
};
constexpr psy_counted operator+(psy_counted a, psy_counted b) { PSY_TICK(add); return psy_counted(a.v + b.v); }
constexpr psy_counted operator-(psy_counted a, psy_counted b) { PSY_TICK(add); return psy_counted(a.v - b.v); }
constexpr psy_counted operator*(psy_counted a, psy_counted b) { PSY_TICK(mul); return psy_counted(a.v * b.v); }
constexpr psy_counted operator/(psy_counted a, psy_counted b) { PSY_TICK(div_); return psy_counted(a.v / b.v); }
constexpr psy_counted operator-(psy_counted a) { return psy_counted(-a.v); }
#define PSY_MIXED(OP)                                                                             \
  constexpr psy_counted operator OP(psy_counted a, double b) { return a OP psy_counted(b); }      \
  constexpr psy_counted operator OP(double a, psy_counted b) { return psy_counted(a) OP b; }      \
  constexpr psy_counted operator OP(psy_counted a, int b) { return a OP psy_counted(b); }         \
  constexpr psy_counted operator OP(int a, psy_counted b) { return psy_counted(a) OP b; }
PSY_MIXED(+) PSY_MIXED(-) PSY_MIXED(*) PSY_MIXED(/)
#undef PSY_MIXED
#define PSY_CMP(OP)                                                                \
  inline bool operator OP(psy_counted a, psy_counted b) { return a.v OP b.v; }     \
  inline bool operator OP(psy_counted a, double b) { return a.v OP b; }            \
  inline bool operator OP(double a, psy_counted b) { return a OP b.v; }
PSY_CMP(<) PSY_CMP(>) PSY_CMP(<=) PSY_CMP(>=) PSY_CMP(==) PSY_CMP(!=)
#undef PSY_CMP
inline psy_counted fma(psy_counted a, psy_counted b, psy_counted c) { psy_ops::fma_++; return psy_counted(__builtin_fma(a.v, b.v, c.v)); }
inline psy_counted fma(psy_counted a, psy_counted b, double c) { return fma(a, b, psy_counted(c)); }
inline psy_counted fma(psy_counted a, double b, psy_counted c) { return fma(a, psy_counted(b), c); }
inline psy_counted fma(double a, psy_counted b, psy_counted c) { return fma(psy_counted(a), b, c); }
inline psy_counted fma(psy_counted a, double b, double c) { return fma(a, psy_counted(b), psy_counted(c)); }
inline psy_counted fma(double a, psy_counted b, double c) { return fma(psy_counted(a), b, psy_counted(c)); }
inline psy_counted sqrt(psy_counted a) { psy_ops::special++; return psy_counted(__builtin_sqrt(a.v)); }
inline psy_counted exp(psy_counted a) { psy_ops::special++; return psy_counted(__builtin_exp(a.v)); }
inline psy_counted log(psy_counted a) { psy_ops::special++; return psy_counted(__builtin_log(a.v)); }
inline psy_counted pow(psy_counted a, psy_counted b) { psy_ops::special++; return psy_counted(__builtin_pow(a.v, b.v)); }
inline psy_counted pow(psy_counted a, double b) { psy_ops::special++; return psy_counted(__builtin_pow(a.v, b)); }
inline psy_counted sin(psy_counted a) { psy_ops::special++; return psy_counted(__builtin_sin(a.v)); }
inline psy_counted cos(psy_counted a) { psy_ops::special++; return psy_counted(__builtin_cos(a.v)); }
inline psy_counted tanh(psy_counted a) { psy_ops::special++; return psy_counted(__builtin_tanh(a.v)); }
inline psy_counted fabs(psy_counted a) { return psy_counted(__builtin_fabs(a.v)); }
inline psy_counted fmax(psy_counted a, psy_counted b) { return psy_counted(__builtin_fmax(a.v, b.v)); }
inline bool isfinite(psy_counted a) { return __builtin_isfinite(a.v); }
static inline psy_counted psy_sim_rcp_seed(psy_counted x) { return psy_counted(psy_sim_rcp_seed(x.v)); }
static inline psy_counted psy_sim_rsqrt_seed(psy_counted x) { return psy_counted(psy_sim_rsqrt_seed(x.v)); }
static inline psy_counted psy_longlong_as_counted(long long v) { double d; memcpy(&d, &v, sizeof d); return psy_counted(d); }
#define __longlong_as_double psy_longlong_as_counted
#define double psy_counted
