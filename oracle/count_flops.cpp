// ORACLE tooling (test infrastructure, not product code) — freezes F_eval, the ALGORITHMIC per-knot evaluation FLOPs of
// SURVEY 8(d) / BASELINE.md section 4, by running the oracle's own knot evaluation with an INSTRUMENTED SCALAR: every `double`
// of the oracle headers is replaced by a counting wrapper, so each +, -, *, / (and sqrt / sin / cos / atan2, tallied apart) the
// restated algorithm performs is counted exactly.  Built and run by tools/count_eval_flops.py -> profiles/eval_flops.json.
#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <vector>
#include <stdint.h>

static thread_local unsigned long long g_add = 0, g_mul = 0, g_div = 0, g_sqrt = 0, g_trig = 0;

struct Cnt {
  double v;
  Cnt() = default;
  Cnt(double x) : v(x) {}
  Cnt(int x) : v(x) {}
  Cnt(unsigned long x) : v((double)x) {}
  Cnt(bool x) : v(x) {}
  operator double() const { return v; }
};
inline Cnt operator+(Cnt a, Cnt b) { g_add++; return Cnt(a.v + b.v); }
inline Cnt operator-(Cnt a, Cnt b) { g_add++; return Cnt(a.v - b.v); }
inline Cnt operator*(Cnt a, Cnt b) { g_mul++; return Cnt(a.v * b.v); }
inline Cnt operator/(Cnt a, Cnt b) { g_div++; return Cnt(a.v / b.v); }
inline Cnt operator-(Cnt a) { return Cnt(-a.v); }
#define MIXED(op)                                                                                                         \
  inline Cnt operator op(Cnt a, double b) { return a op Cnt(b); }                                                         \
  inline Cnt operator op(double a, Cnt b) { return Cnt(a) op b; }                                                         \
  inline Cnt operator op(Cnt a, int b) { return a op Cnt(b); }                                                            \
  inline Cnt operator op(int a, Cnt b) { return Cnt(a) op b; }
MIXED(+) MIXED(-) MIXED(*) MIXED(/)
inline Cnt &operator+=(Cnt &a, Cnt b) { a = a + b; return a; }
inline Cnt &operator-=(Cnt &a, Cnt b) { a = a - b; return a; }
inline Cnt &operator*=(Cnt &a, Cnt b) { a = a * b; return a; }
inline Cnt &operator/=(Cnt &a, Cnt b) { a = a / b; return a; }
inline bool operator<(Cnt a, Cnt b) { return a.v < b.v; }
inline bool operator>(Cnt a, Cnt b) { return a.v > b.v; }
inline bool operator<=(Cnt a, Cnt b) { return a.v <= b.v; }
inline bool operator>=(Cnt a, Cnt b) { return a.v >= b.v; }
inline bool operator==(Cnt a, Cnt b) { return a.v == b.v; }
inline bool operator!=(Cnt a, Cnt b) { return a.v != b.v; }
#define CMPMIX(op)                                                                                                        \
  inline bool operator op(Cnt a, double b) { return a.v op b; }                                                           \
  inline bool operator op(double a, Cnt b) { return a op b.v; }                                                           \
  inline bool operator op(Cnt a, int b) { return a.v op b; }                                                              \
  inline bool operator op(int a, Cnt b) { return a op b.v; }
CMPMIX(<) CMPMIX(>) CMPMIX(<=) CMPMIX(>=) CMPMIX(==) CMPMIX(!=)
namespace std {
inline Cnt sqrt(Cnt a) { g_sqrt++; return Cnt(::sqrt(a.v)); }
inline Cnt sin(Cnt a) { g_trig++; return Cnt(::sin(a.v)); }
inline Cnt cos(Cnt a) { g_trig++; return Cnt(::cos(a.v)); }
inline Cnt atan2(Cnt a, Cnt b) { g_trig++; return Cnt(::atan2(a.v, b.v)); }
inline Cnt fabs(Cnt a) { return Cnt(::fabs(a.v)); }
inline Cnt pow(Cnt a, Cnt b) { g_trig++; return Cnt(::pow(a.v, b.v)); }
inline bool isfinite(Cnt a) { return std::isfinite(a.v); }
inline Cnt max(Cnt a, Cnt b) { return a.v < b.v ? b : a; }
inline Cnt min(Cnt a, Cnt b) { return b.v < a.v ? b : a; }
} // namespace std

#define double Cnt
#include "knot.hpp"
#undef double

using namespace orc;

extern "C" {
// counts[5] = add/sub, mul, div, sqrt, trig for ONE evaluation of the knot (kn != NULL) or of the terminal knot (kn == NULL, tm)
void orc_count_eval(const mpc_robot_t *rb, const mpc_config_t *cfg, const mpc_knot_t *kn, const mpc_term_t *tm, const double *x, const double *u,
                    const double *xn, int derivs, unsigned long long *counts) {
  Problem P(rb, *cfg);
  KnotEval e; e.resize(P.d);
  std::vector<Cnt> X(x, x + P.d.nx), U(u, u + P.d.m), XN(xn, xn + P.d.nx);
  g_add = g_mul = g_div = g_sqrt = g_trig = 0;
  if (kn) eval_knot(P, *kn, X.data(), U.data(), XN.data(), derivs != 0, e); else eval_term(P, *tm, X.data(), e);
  counts[0] = g_add; counts[1] = g_mul; counts[2] = g_div; counts[3] = g_sqrt; counts[4] = g_trig;
}
}
