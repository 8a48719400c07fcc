// ORACLE (test infrastructure, not product code) — spatial algebra and SE(3) Lie-group maps.
//
// CPU restatement of the Pinocchio conventions the reference relies on:
// spatial vectors are [linear; angular], placements act as y = R x + p, integrate is M*exp6(d),
// difference(a,b) = log6(a^-1 b) (reference use: fulldynamic_talos.py:522 `space.difference`).
// Pinocchio itself is NOT in /root/reference (pip dependency, README.md:16), so parity with it is
// UNPINNED; every Jacobian here is checked against forward-mode AD (Dual) in tests/.
//
// Everything is templated on the scalar so the same value code runs with `double` and `Dual`.
#pragma once
#include <array>
#include <cmath>

namespace orc {

// ---------------------------------------------------------------- forward-mode dual number
struct Dual {
  double v = 0, d = 0;
  Dual() = default;
  Dual(double v_) : v(v_), d(0) {}
  Dual(double v_, double d_) : v(v_), d(d_) {}
};
inline Dual operator+(Dual a, Dual b) { return {a.v + b.v, a.d + b.d}; }
inline Dual operator-(Dual a, Dual b) { return {a.v - b.v, a.d - b.d}; }
inline Dual operator-(Dual a) { return {-a.v, -a.d}; }
inline Dual operator*(Dual a, Dual b) { return {a.v * b.v, a.d * b.v + a.v * b.d}; }
inline Dual operator/(Dual a, Dual b) { return {a.v / b.v, (a.d * b.v - a.v * b.d) / (b.v * b.v)}; }
inline Dual &operator+=(Dual &a, Dual b) { a = a + b; return a; }
inline Dual &operator-=(Dual &a, Dual b) { a = a - b; return a; }
inline Dual &operator*=(Dual &a, Dual b) { a = a * b; return a; }
inline bool operator<(Dual a, Dual b) { return a.v < b.v; }
inline bool operator>(Dual a, Dual b) { return a.v > b.v; }
inline Dual sqrt(Dual a) { double s = std::sqrt(a.v); return {s, a.d / (2 * s)}; }
inline Dual sin(Dual a) { return {std::sin(a.v), std::cos(a.v) * a.d}; }
inline Dual cos(Dual a) { return {std::cos(a.v), -std::sin(a.v) * a.d}; }
inline Dual atan2(Dual y, Dual x) {
  double r2 = x.v * x.v + y.v * y.v;
  return {std::atan2(y.v, x.v), (x.v * y.d - y.v * x.d) / r2};
}
inline double val(double a) { return a; }
inline double val(Dual a) { return a.v; }
using std::atan2;
using std::cos;
using std::sin;
using std::sqrt;

template <class T> using V3 = std::array<T, 3>;
template <class T> using M3 = std::array<T, 9>;  // row-major
template <class T> using V6 = std::array<T, 6>;  // [lin; ang]
template <class T> using M6 = std::array<T, 36>; // row-major

template <class T> V3<T> add(const V3<T> &a, const V3<T> &b) { return {a[0] + b[0], a[1] + b[1], a[2] + b[2]}; }
template <class T> V3<T> sub(const V3<T> &a, const V3<T> &b) { return {a[0] - b[0], a[1] - b[1], a[2] - b[2]}; }
template <class T> V3<T> scale(const V3<T> &a, T s) { return {a[0] * s, a[1] * s, a[2] * s}; }
template <class T> T dot(const V3<T> &a, const V3<T> &b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
template <class T> V3<T> cross(const V3<T> &a, const V3<T> &b) {
  return {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
}
template <class T> V3<T> mul(const M3<T> &A, const V3<T> &x) {
  return {A[0] * x[0] + A[1] * x[1] + A[2] * x[2], A[3] * x[0] + A[4] * x[1] + A[5] * x[2],
          A[6] * x[0] + A[7] * x[1] + A[8] * x[2]};
}
template <class T> V3<T> mulT(const M3<T> &A, const V3<T> &x) {
  return {A[0] * x[0] + A[3] * x[1] + A[6] * x[2], A[1] * x[0] + A[4] * x[1] + A[7] * x[2],
          A[2] * x[0] + A[5] * x[1] + A[8] * x[2]};
}
template <class T> M3<T> mul(const M3<T> &A, const M3<T> &B) {
  M3<T> C;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
  return C;
}
template <class T> M3<T> transpose(const M3<T> &A) { return {A[0], A[3], A[6], A[1], A[4], A[7], A[2], A[5], A[8]}; }
template <class T> M3<T> skew(const V3<T> &a) { return {T(0), -a[2], a[1], a[2], T(0), -a[0], -a[1], a[0], T(0)}; }
template <class T> M3<T> eye3() { return {T(1), T(0), T(0), T(0), T(1), T(0), T(0), T(0), T(1)}; }
template <class T> M3<T> add(const M3<T> &A, const M3<T> &B) { M3<T> C; for (int i = 0; i < 9; i++) C[i] = A[i] + B[i]; return C; }
template <class T> M3<T> scale(const M3<T> &A, T s) { M3<T> C; for (int i = 0; i < 9; i++) C[i] = A[i] * s; return C; }

template <class T> struct SE3 {
  M3<T> R;
  V3<T> p;
  static SE3 identity() { return {eye3<T>(), {T(0), T(0), T(0)}}; }
};
template <class T> SE3<T> mul(const SE3<T> &a, const SE3<T> &b) { return {mul(a.R, b.R), add(mul(a.R, b.p), a.p)}; }
template <class T> SE3<T> inverse(const SE3<T> &a) { M3<T> Rt = transpose(a.R); return {Rt, scale(mul(Rt, a.p), T(-1))}; }
template <class T> SE3<T> se3_from12(const T *d) {
  SE3<T> M; for (int i = 0; i < 9; i++) M.R[i] = d[i]; for (int i = 0; i < 3; i++) M.p[i] = d[9 + i]; return M;
}
template <class T, class S> SE3<T> se3_cast(const S *d) {
  SE3<T> M; for (int i = 0; i < 9; i++) M.R[i] = T(d[i]); for (int i = 0; i < 3; i++) M.p[i] = T(d[9 + i]); return M;
}

template <class T> V3<T> lin(const V6<T> &m) { return {m[0], m[1], m[2]}; }
template <class T> V3<T> ang(const V6<T> &m) { return {m[3], m[4], m[5]}; }
template <class T> V6<T> mk6(const V3<T> &l, const V3<T> &a) { return {l[0], l[1], l[2], a[0], a[1], a[2]}; }
template <class T> V6<T> add(const V6<T> &a, const V6<T> &b) { V6<T> c; for (int i = 0; i < 6; i++) c[i] = a[i] + b[i]; return c; }
template <class T> V6<T> sub(const V6<T> &a, const V6<T> &b) { V6<T> c; for (int i = 0; i < 6; i++) c[i] = a[i] - b[i]; return c; }
template <class T> V6<T> scale(const V6<T> &a, T s) { V6<T> c; for (int i = 0; i < 6; i++) c[i] = a[i] * s; return c; }
template <class T> T dot(const V6<T> &a, const V6<T> &b) { T s = a[0] * b[0]; for (int i = 1; i < 6; i++) s += a[i] * b[i]; return s; }
template <class T> V6<T> zero6() { return {T(0), T(0), T(0), T(0), T(0), T(0)}; }

// motion: local -> world
template <class T> V6<T> act_motion(const SE3<T> &M, const V6<T> &m) {
  V3<T> w = mul(M.R, ang(m));
  return mk6(add(mul(M.R, lin(m)), cross(M.p, w)), w);
}
template <class T> V6<T> actinv_motion(const SE3<T> &M, const V6<T> &m) {
  return mk6(mulT(M.R, sub(lin(m), cross(M.p, ang(m)))), mulT(M.R, ang(m)));
}
// force: local -> world
template <class T> V6<T> act_force(const SE3<T> &M, const V6<T> &f) {
  V3<T> fl = mul(M.R, lin(f));
  return mk6(fl, add(mul(M.R, ang(f)), cross(M.p, fl)));
}
template <class T> V6<T> actinv_force(const SE3<T> &M, const V6<T> &f) {
  return mk6(mulT(M.R, lin(f)), mulT(M.R, sub(ang(f), cross(M.p, lin(f)))));
}
// m1 x m2 (motion) and m x* f (force)
template <class T> V6<T> cross_mm(const V6<T> &a, const V6<T> &b) {
  return mk6(add(cross(ang(a), lin(b)), cross(lin(a), ang(b))), cross(ang(a), ang(b)));
}
template <class T> V6<T> cross_mf(const V6<T> &a, const V6<T> &f) {
  return mk6(cross(ang(a), lin(f)), add(cross(ang(a), ang(f)), cross(lin(a), lin(f))));
}
template <class T> V6<T> mul(const M6<T> &A, const V6<T> &x) {
  V6<T> y;
  for (int i = 0; i < 6; i++) { T s = A[6 * i] * x[0]; for (int j = 1; j < 6; j++) s += A[6 * i + j] * x[j]; y[i] = s; }
  return y;
}
// 6x6 matrix of the motion action of M (Ad_M): m_world = Ad * m_local
template <class T> M6<T> action_matrix(const SE3<T> &M) {
  M6<T> A; for (auto &a : A) a = T(0);
  M3<T> pR = mul(skew(M.p), M.R);
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) { A[6 * i + j] = M.R[3 * i + j]; A[6 * i + 3 + j] = pR[3 * i + j]; A[6 * (i + 3) + 3 + j] = M.R[3 * i + j]; }
  return A;
}
// spatial inertia about the frame origin from (mass, com, I_c): [[m I, -m c^],[m c^, I_c - m c^ c^]]
template <class T> M6<T> inertia6(T m, const V3<T> &c, const M3<T> &Ic) {
  M6<T> I; for (auto &a : I) a = T(0);
  M3<T> cx = skew(c), cc = mul(cx, cx);
  for (int i = 0; i < 3; i++) {
    I[6 * i + i] = m;
    for (int j = 0; j < 3; j++) {
      I[6 * i + 3 + j] = -m * cx[3 * i + j];
      I[6 * (i + 3) + j] = m * cx[3 * i + j];
      I[6 * (i + 3) + 3 + j] = Ic[3 * i + j] - m * cc[3 * i + j];
    }
  }
  return I;
}

// ---------------------------------------------------------------- SO(3)/SE(3) exp/log
// coefficient helpers as functions of t2 = theta^2 (series below 1e-6 so Dual never divides by 0)
template <class T> void so3_coeffs(T t2, T &a, T &b, T &c) { // a = sin/θ, b=(1-cos)/θ², c=(θ-sin)/θ³
  if (val(t2) < 1e-6) {
    a = T(1) - t2 / T(6) + t2 * t2 / T(120);
    b = T(0.5) - t2 / T(24) + t2 * t2 / T(720);
    c = T(1) / T(6) - t2 / T(120) + t2 * t2 / T(5040);
  } else {
    T t = sqrt(t2);
    a = sin(t) / t; b = (T(1) - cos(t)) / t2; c = (t - sin(t)) / (t2 * t);
  }
}
template <class T> M3<T> exp3(const V3<T> &w) {
  T t2 = dot(w, w), a, b, c; so3_coeffs(t2, a, b, c);
  M3<T> W = skew(w);
  return add(add(eye3<T>(), scale(W, a)), scale(mul(W, W), b));
}
template <class T> V3<T> log3(const M3<T> &R) {
  V3<T> s = {(R[7] - R[5]) * T(0.5), (R[2] - R[6]) * T(0.5), (R[3] - R[1]) * T(0.5)}; // sinθ * axis
  T ct = (R[0] + R[4] + R[8] - T(1)) * T(0.5);
  T s2 = dot(s, s), f;
  if (val(s2) < 1e-6 && val(ct) > 0) {
    f = T(1) + s2 / T(6) + s2 * s2 * T(3.0 / 40.0) + s2 * s2 * s2 * T(15.0 / 336.0);
  } else {
    T sn = sqrt(s2); f = atan2(sn, ct) / sn;
  }
  return scale(s, f);
}
// c' of V^-1 = I - 1/2 w^ + c' w^ w^
template <class T> T vinv_coeff(T t2) {
  if (val(t2) < 1e-6) return T(1.0 / 12.0) + t2 / T(720) + t2 * t2 / T(30240);
  T t = sqrt(t2);
  return (T(1) - t * sin(t) / (T(2) * (T(1) - cos(t)))) / t2;
}
template <class T> SE3<T> exp6(const V6<T> &xi) {
  V3<T> v = lin(xi), w = ang(xi);
  T t2 = dot(w, w), a, b, c; so3_coeffs(t2, a, b, c);
  M3<T> W = skew(w), W2 = mul(W, W);
  SE3<T> M;
  M.R = add(add(eye3<T>(), scale(W, a)), scale(W2, b));
  M3<T> V = add(add(eye3<T>(), scale(W, b)), scale(W2, c));
  M.p = mul(V, v);
  return M;
}
template <class T> V6<T> log6(const SE3<T> &M) {
  V3<T> w = log3(M.R);
  T t2 = dot(w, w), cp = vinv_coeff(t2);
  M3<T> W = skew(w);
  M3<T> Vi = add(add(eye3<T>(), scale(W, T(-0.5))), scale(mul(W, W), cp));
  return mk6(mul(Vi, M.p), w);
}

// Right Jacobian of exp6 at xi (d exp6(xi+d) = exp6(xi) exp6(Jr d)) and its inverse.
// Barfoot, State Estimation for Robotics, eq. 7.85-7.86 with J_r(xi) = J_l(-xi).
inline void q_block(const V3<double> &rho, const V3<double> &phi, M3<double> &Q) {
  double t2 = dot(phi, phi), c1, c2, c3;
  if (t2 < 1e-6) {
    c1 = 1.0 / 6 - t2 / 120 + t2 * t2 / 5040;
    c2 = 1.0 / 24 - t2 / 720 + t2 * t2 / 40320;
    c3 = 1.0 / 120 - t2 / 2520 + t2 * t2 / 120960;
  } else {
    double t = std::sqrt(t2), s = std::sin(t), c = std::cos(t);
    c1 = (t - s) / (t2 * t);
    c2 = (t2 + 2 * c - 2) / (2 * t2 * t2);
    c3 = (2 * t - 3 * s + t * c) / (2 * t2 * t2 * t);
  }
  M3<double> P = skew(phi), Rr = skew(rho);
  M3<double> PR = mul(P, Rr), RP = mul(Rr, P), PRP = mul(PR, P);
  M3<double> PPR = mul(P, PR), RPP = mul(RP, P), PRPP = mul(PRP, P), PPRP = mul(P, PRP);
  for (int i = 0; i < 9; i++)
    Q[i] = 0.5 * Rr[i] + c1 * (PR[i] + RP[i] + PRP[i]) + c2 * (PPR[i] + RPP[i] - 3 * PRP[i]) + c3 * (PRPP[i] + PPRP[i]);
}
inline M6<double> Jexp6(const V6<double> &xi) { // right Jacobian
  V3<double> rho = lin(xi), phi = ang(xi);
  double t2 = dot(phi, phi), a, b, c; so3_coeffs(t2, a, b, c);
  M3<double> W = skew(phi), W2 = mul(W, W), J3, Q;
  for (int i = 0; i < 9; i++) J3[i] = (i % 4 == 0 ? 1.0 : 0.0) - b * W[i] + c * W2[i];
  q_block(scale(rho, double(-1.0)), scale(phi, double(-1.0)), Q);
  M6<double> J{};
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) { J[6 * i + j] = J3[3 * i + j]; J[6 * i + 3 + j] = Q[3 * i + j]; J[6 * (i + 3) + 3 + j] = J3[3 * i + j]; }
  return J;
}
inline M6<double> Jlog6(const SE3<double> &M) { // log6(M exp6(d)) ~ log6(M) + Jlog6 d
  V6<double> xi = log6(M);
  V3<double> rho = lin(xi), phi = ang(xi);
  double t2 = dot(phi, phi), cp = vinv_coeff(t2);
  M3<double> W = skew(phi), W2 = mul(W, W), J3i, Q;
  for (int i = 0; i < 9; i++) J3i[i] = (i % 4 == 0 ? 1.0 : 0.0) + 0.5 * W[i] + cp * W2[i];
  q_block(scale(rho, double(-1.0)), scale(phi, double(-1.0)), Q);
  M3<double> B = mul(mul(J3i, Q), J3i);
  M6<double> J{};
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) { J[6 * i + j] = J3i[3 * i + j]; J[6 * i + 3 + j] = -B[3 * i + j]; J[6 * (i + 3) + 3 + j] = J3i[3 * i + j]; }
  return J;
}

// ---------------------------------------------------------------- quaternion (x,y,z,w) <-> R
template <class T> M3<T> quat_to_R(const T *q) {
  T x = q[0], y = q[1], z = q[2], w = q[3];
  T n = x * x + y * y + z * z + w * w, s = T(2) / n;
  return {T(1) - s * (y * y + z * z), s * (x * y - z * w), s * (x * z + y * w),
          s * (x * y + z * w), T(1) - s * (x * x + z * z), s * (y * z - x * w),
          s * (x * z - y * w), s * (y * z + x * w), T(1) - s * (x * x + y * y)};
}
// q <- q (x) dq(w), dq = exp of rotation vector w; renormalised (pinocchio integrate)
template <class T> void quat_integrate(const T *q, const V3<T> &w, T *out) {
  T t2 = dot(w, w), sh, ch; // sh = sin(θ/2)/θ, ch = cos(θ/2)
  if (val(t2) < 1e-6) { sh = T(0.5) - t2 / T(48) + t2 * t2 / T(3840); ch = T(1) - t2 / T(8) + t2 * t2 / T(384); }
  else { T t = sqrt(t2); sh = sin(t * T(0.5)) / t; ch = cos(t * T(0.5)); }
  T dx = w[0] * sh, dy = w[1] * sh, dz = w[2] * sh, dw = ch;
  T x = q[0], y = q[1], z = q[2], ww = q[3];
  T ox = ww * dx + x * dw + y * dz - z * dy;
  T oy = ww * dy - x * dz + y * dw + z * dx;
  T oz = ww * dz + x * dy - y * dx + z * dw;
  T ow = ww * dw - x * dx - y * dy - z * dz;
  T n = sqrt(ox * ox + oy * oy + oz * oz + ow * ow);
  out[0] = ox / n; out[1] = oy / n; out[2] = oz / n; out[3] = ow / n;
}

} // namespace orc
