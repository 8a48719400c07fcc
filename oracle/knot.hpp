// ORACLE (test infrastructure, not product code) — per-knot evaluation of the three Talos OCPs.
//
// Restates what Aligator's StageModel::evaluate/computeFirstDerivatives/computeSecondDerivatives produce
// for exactly the stages the reference builds (no other stage structure is supported):
//   full dynamics   fulldynamic_talos.py:100-111,153-232 (createStage), 234-245 (terminal cost), 499-507
//   kinodynamics    kinodynamic_talos.py:107-173, 175-180
//   centroidal      centroidal_talos.py:202-247
// Aligator is a pip dependency absent from /root/reference (README.md:10) => PARITY UNPINNED.
// Conventions: SURVEY 8a (C1 sign, K4 Gauss-Newton Hessians), App. A2 (integrators), App. A7 (cone).
#pragma once
#include "rbd.hpp"
#include <cstring>
#include <vector>

namespace orc {

struct Dims { int nx, n, m, nc; };
inline Dims dims_of(int kind) {
  if (kind == MPC_KIND_CENT) return {9, 9, 12, 34};
  if (kind == MPC_KIND_KINO) return {57, 56, 34, 68};
  return {57, 56, 22, 78};
}

enum { SET_NONE = -1, SET_EQ = 0, SET_NEG = 1, SET_BOX = 2 };

struct KnotEval {
  int n = 0, m = 0, nc = 0, nx = 0;
  std::vector<double> xnext, gap, A, B, E6; // E6: 6x6 block of d gap / d x_{k+1} (rest is -I); identity-neg for vector spaces
  double cost = 0;
  std::vector<double> lx, lu, H; // H: (n+m)^2 row-major, [[xx, xu],[ux, uu]]
  std::vector<double> h, Cx, Cu, lo, hi;
  std::vector<int> ctype;
  double xdot[56], lam[12];
  void resize(const Dims &d) {
    nx = d.nx; n = d.n; m = d.m; nc = d.nc;
    xnext.assign(nx, 0); gap.assign(n, 0); A.assign(n * n, 0); B.assign(n * m, 0); E6.assign(36, 0);
    lx.assign(n, 0); lu.assign(m, 0); H.assign((n + m) * (n + m), 0);
    h.assign(nc, 0); Cx.assign(nc * n, 0); Cu.assign(nc * m, 0); lo.assign(nc, 0); hi.assign(nc, 0); ctype.assign(nc, SET_NONE);
  }
  void zero() {
    std::fill(gap.begin(), gap.end(), 0.0); std::fill(A.begin(), A.end(), 0.0); std::fill(B.begin(), B.end(), 0.0);
    std::fill(lx.begin(), lx.end(), 0.0); std::fill(lu.begin(), lu.end(), 0.0); std::fill(H.begin(), H.end(), 0.0);
    std::fill(h.begin(), h.end(), 0.0); std::fill(Cx.begin(), Cx.end(), 0.0); std::fill(Cu.begin(), Cu.end(), 0.0);
    std::fill(ctype.begin(), ctype.end(), (int)SET_NONE); cost = 0;
    std::memset(xdot, 0, sizeof xdot); std::memset(lam, 0, sizeof lam);
  }
};

// 17 x 6 wrench-cone matrix, r = A w <= 0, w = (f, tau) in the sole frame (App. A7; rows 0-8 agree with
// QP_utils.py:337-347 up to the sign convention C w >= l used there).
inline void cone_matrix(double mu, double L, double W, double *A) {
  std::memset(A, 0, sizeof(double) * 17 * 6);
  auto row = [&](int r, double fx, double fy, double fz, double tx, double ty, double tz) {
    double *a = A + 6 * r; a[0] = fx; a[1] = fy; a[2] = fz; a[3] = tx; a[4] = ty; a[5] = tz;
  };
  row(0, 0, 0, -1, 0, 0, 0);
  row(1, 1, 0, -mu, 0, 0, 0); row(2, -1, 0, -mu, 0, 0, 0);
  row(3, 0, 1, -mu, 0, 0, 0); row(4, 0, -1, -mu, 0, 0, 0);
  row(5, 0, 0, -W, 1, 0, 0);  row(6, 0, 0, -W, -1, 0, 0);
  row(7, 0, 0, -L, 0, 1, 0);  row(8, 0, 0, -L, 0, -1, 0);
  int r = 9;
  for (int s1 = 1; s1 >= -1; s1 -= 2)
    for (int s2 = 1; s2 >= -1; s2 -= 2) { // tau_z >= tau_z_min
      row(r++, s1 * W, s2 * L, -mu * (L + W), -s1 * mu, -s2 * mu, -1);
    }
  for (int s1 = 1; s1 >= -1; s1 -= 2)
    for (int s2 = 1; s2 >= -1; s2 -= 2) { // tau_z <= tau_z_max
      row(r++, s1 * W, s2 * L, -mu * (L + W), s1 * mu, s2 * mu, 1);
    }
}

// ---- manifold helpers (MultibodyPhaseSpace, App. A1)
template <class T> void mb_integrate(const T *x, const T *dx, T *out) {
  V3<T> dv = {dx[0], dx[1], dx[2]}, dw = {dx[3], dx[4], dx[5]};
  M3<T> R = quat_to_R(x + 3);
  SE3<T> e = exp6(mk6(dv, dw));
  V3<T> p = add(mul(R, e.p), V3<T>{x[0], x[1], x[2]});
  out[0] = p[0]; out[1] = p[1]; out[2] = p[2];
  quat_integrate(x + 3, dw, out + 3);
  for (int i = 0; i < NJ; i++) out[7 + i] = x[7 + i] + dx[6 + i];
  for (int i = 0; i < NV; i++) out[NQ + i] = x[NQ + i] + dx[NV + i];
}
template <class T> SE3<T> base_placement(const T *x) { return {quat_to_R(x + 3), {x[0], x[1], x[2]}}; }
// difference(x0, x1) = x1 (-) x0
template <class T> void mb_difference(const T *x0, const T *x1, T *out) {
  V6<T> l = log6(mul(inverse(base_placement(x0)), base_placement(x1)));
  for (int i = 0; i < 6; i++) out[i] = l[i];
  for (int i = 0; i < NJ; i++) out[6 + i] = x1[7 + i] - x0[7 + i];
  for (int i = 0; i < NV; i++) out[NV + i] = x1[NQ + i] - x0[NQ + i];
}

// accumulate a weighted residual cost 1/2 r^T W r with Jacobian J (nr x nz, only columns listed) into
// (cost, grad, H) over z = (x,u) of size nz = n+m
inline void add_residual_cost(const double *r, const double *w, int nr, const double *J, int nz, double &cost, double *grad, double *H) {
  for (int a = 0; a < nr; a++) {
    if (w[a] == 0.0) continue;
    cost += 0.5 * w[a] * r[a] * r[a];
    const double *Ja = J + a * nz;
    for (int i = 0; i < nz; i++) {
      if (Ja[i] == 0.0) continue;
      grad[i] += w[a] * Ja[i] * r[a];
      double wi = w[a] * Ja[i];
      for (int j = 0; j < nz; j++) H[i * nz + j] += wi * Ja[j];
    }
  }
}

struct Problem {
  const mpc_robot_t *rb;
  mpc_config_t cfg;
  Tree tree;
  Dims d;
  double Acone[17 * 6];
  Problem(const mpc_robot_t *r, const mpc_config_t &c) : rb(r), cfg(c), tree(r), d(dims_of(c.kind)) {
    cone_matrix(cfg.mu_fric, cfg.foot_L, cfg.foot_W, Acone);
  }
};

// ============================================================ centroidal (cent:202-247)
inline void eval_knot_cent(const Problem &P, const mpc_knot_t &kn, const double *x, const double *u, const double *xn, bool derivs, KnotEval &o) {
  const mpc_config_t &c = P.cfg;
  const int n = 9, m = 12, nz = 21;
  double mass = c.mass, dt = c.dt;
  const double *g = P.rb->gravity;
  double xd[9];
  for (int i = 0; i < 3; i++) { xd[i] = x[3 + i] / mass; xd[3 + i] = mass * g[i]; xd[6 + i] = 0; }
  V3<double> com = {x[0], x[1], x[2]};
  V3<double> ftot = {0, 0, 0};
  for (int k = 0; k < 2; k++) {
    if (kn.cs[k] == 0.0) continue;
    V3<double> f = {u[6 * k], u[6 * k + 1], u[6 * k + 2]}, t = {u[6 * k + 3], u[6 * k + 4], u[6 * k + 5]};
    V3<double> p = {kn.cpos[3 * k], kn.cpos[3 * k + 1], kn.cpos[3 * k + 2]};
    V3<double> mo = add(cross(sub(p, com), f), t);
    for (int i = 0; i < 3; i++) { xd[3 + i] += f[i]; xd[6 + i] += mo[i]; ftot[i] += f[i]; }
  }
  for (int i = 0; i < 9; i++) { o.xdot[i] = xd[i]; o.xnext[i] = x[i] + dt * xd[i]; o.gap[i] = o.xnext[i] - xn[i]; }
  // Jacobians of xdot
  double Fx[81] = {0}, Fu[9 * 12] = {0};
  for (int i = 0; i < 3; i++) Fx[i * 9 + 3 + i] = 1.0 / mass;
  M3<double> fx = skew(ftot); // d/dc sum (p-c) x f = +[f]x
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Fx[(6 + i) * 9 + j] = fx[3 * i + j];
  for (int k = 0; k < 2; k++) {
    if (kn.cs[k] == 0.0) continue;
    V3<double> p = {kn.cpos[3 * k], kn.cpos[3 * k + 1], kn.cpos[3 * k + 2]};
    M3<double> px = skew(sub(p, com));
    for (int i = 0; i < 3; i++) {
      Fu[(3 + i) * 12 + 6 * k + i] = 1.0;
      Fu[(6 + i) * 12 + 6 * k + 3 + i] = 1.0;
      for (int j = 0; j < 3; j++) Fu[(6 + i) * 12 + 6 * k + j] = px[3 * i + j];
    }
  }
  if (derivs) {
    for (int i = 0; i < 9; i++) {
      for (int j = 0; j < 9; j++) o.A[i * 9 + j] = (i == j ? 1.0 : 0.0) + dt * Fx[i * 9 + j];
      for (int j = 0; j < 12; j++) o.B[i * 12 + j] = dt * Fu[i * 12 + j];
    }
    for (int i = 0; i < 36; i++) o.E6[i] = (i % 7 == 0) ? -1.0 : 0.0;
  }
  // costs
  std::vector<double> grad(nz, 0.0);
  double J[3 * 21];
  double r3[3];
  // control cost ("state_cost", cent:224-225)
  for (int i = 0; i < m; i++) {
    double e = u[i] - kn.u_ref[i];
    o.cost += 0.5 * c.wu[i] * e * e; grad[n + i] += c.wu[i] * e; o.H[(n + i) * nz + n + i] += c.wu[i];
  }
  auto state_block = [&](int off, const double *ref, const double *w) {
    std::memset(J, 0, sizeof J);
    for (int i = 0; i < 3; i++) { r3[i] = x[off + i] - ref[i]; J[i * nz + off + i] = 1.0; }
    add_residual_cost(r3, w, 3, J, nz, o.cost, grad.data(), o.H.data());
  };
  double zero3[3] = {0, 0, 0};
  state_block(0, c.com_ref, c.w_com);
  state_block(3, zero3, c.w_linmom);
  state_block(6, zero3, c.w_angmom);
  // angular acceleration residual = xdot[6:9] (cent:217-219)
  std::memset(J, 0, sizeof J);
  for (int i = 0; i < 3; i++) { r3[i] = xd[6 + i]; for (int j = 0; j < 9; j++) J[i * nz + j] = Fx[(6 + i) * 9 + j]; for (int j = 0; j < 12; j++) J[i * nz + 9 + j] = Fu[(6 + i) * 12 + j]; }
  add_residual_cost(r3, c.w_angacc, 3, J, nz, o.cost, grad.data(), o.H.data());
  // linear acceleration residual = g + sum f / m (cent:214-216)
  std::memset(J, 0, sizeof J);
  for (int i = 0; i < 3; i++) { r3[i] = xd[3 + i] / mass; for (int j = 0; j < 12; j++) J[i * nz + 9 + j] = Fu[(3 + i) * 12 + j] / mass; }
  add_residual_cost(r3, c.w_linacc, 3, J, nz, o.cost, grad.data(), o.H.data());
  for (int i = 0; i < n; i++) o.lx[i] = grad[i];
  for (int i = 0; i < m; i++) o.lu[i] = grad[n + i];
  // constraints: wrench cones on active contacts (cent:242-245)
  for (int k = 0; k < 2; k++) {
    for (int r = 0; r < 17; r++) {
      int row = 17 * k + r;
      if (kn.cs[k] == 0.0) { o.ctype[row] = SET_NONE; continue; }
      o.ctype[row] = SET_NEG;
      double s = 0;
      for (int j = 0; j < 6; j++) { s += P.Acone[6 * r + j] * u[6 * k + j]; o.Cu[row * m + 6 * k + j] = P.Acone[6 * r + j]; }
      o.h[row] = s;
    }
  }
}

// ============================================================ full dynamics (full:100-111,153-232)
struct FootKin {
  SE3<double> oMf;
  double J[6 * NV]; // LOCAL frame Jacobian
};
inline void foot_kin(const Tree &tr, const Kin<double> &k, int foot, FootKin &f) {
  const mpc_robot_t &rb = *tr.rb;
  int b = rb.foot_body[foot];
  f.oMf = mul(k.oM[b], se3_cast<double>(rb.foot_place[foot]));
  for (int j = 0; j < NV; j++) {
    V6<double> col = tr.anc[body_of_dof(j)][b] ? actinv_motion(f.oMf, k.S[j]) : zero6<double>();
    for (int r = 0; r < 6; r++) f.J[r * NV + j] = col[r];
  }
}

// cost terms shared by the running and terminal full/kino stages: state, centroidal momentum, foot poses
inline void multibody_costs(const Problem &P, const Kin<double> &kin, const double *x, const double *wx, const double *wcent,
                            const double *wlf, const double *wrf, const double *lf_ref, const double *rf_ref, int nz, double &cost,
                            double *grad, double *H) {
  const mpc_config_t &c = P.cfg;
  // state cost: e = x (-) x_ref, J = blockdiag(Jlog6, I)
  {
    double e[56];
    mb_difference<double>(c.x_ref, x, e);
    M6<double> Jl = Jlog6(mul(inverse(base_placement<double>(c.x_ref)), base_placement<double>(x)));
    std::vector<double> J(56 * nz, 0.0);
    for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) J[i * nz + j] = Jl[6 * i + j];
    for (int i = 6; i < 56; i++) J[i * nz + i] = 1.0;
    add_residual_cost(e, wx, 56, J.data(), nz, cost, grad, H);
  }
  // centroidal momentum (full:160-162): r = h_g(q,v) - 0
  {
    V6<double> h = centroidal_momentum(kin);
    double dhq[6 * NV], Ag[6 * NV];
    centroidal_derivatives(P.tree, kin, dhq, Ag);
    std::vector<double> J(6 * nz, 0.0);
    for (int i = 0; i < 6; i++) for (int j = 0; j < NV; j++) { J[i * nz + j] = dhq[i * NV + j]; J[i * nz + NV + j] = Ag[i * NV + j]; }
    add_residual_cost(h.data(), wcent, 6, J.data(), nz, cost, grad, H);
  }
  // foot placement costs (full:164-167,183-185)
  for (int f = 0; f < 2; f++) {
    const double *w = f == 0 ? wlf : wrf;
    bool any = false; for (int i = 0; i < 6; i++) any |= (w[i] != 0.0);
    if (!any) continue;
    FootKin fk; foot_kin(P.tree, kin, f, fk);
    SE3<double> D = mul(inverse(se3_from12<double>(f == 0 ? lf_ref : rf_ref)), fk.oMf);
    V6<double> r = log6(D);
    M6<double> Jl = Jlog6(D);
    std::vector<double> J(6 * nz, 0.0);
    for (int i = 0; i < 6; i++) for (int j = 0; j < NV; j++) { double s = 0; for (int k = 0; k < 6; k++) s += Jl[6 * i + k] * fk.J[k * NV + j]; J[i * nz + j] = s; }
    add_residual_cost(r.data(), w, 6, J.data(), nz, cost, grad, H);
  }
}

// semi-implicit Euler on the multibody phase space + gap and its Jacobians (App. A2)
inline void semi_implicit_euler(double dt, const double *x, const double *acc, const double *ax /*NV x 56*/, const double *au /*NV x m*/,
                                int m, const double *xn, bool derivs, KnotEval &o) {
  const int n = 56;
  double dx[56];
  for (int i = 0; i < NV; i++) { dx[NV + i] = dt * acc[i]; dx[i] = dt * (x[NQ + i] + dx[NV + i]); }
  mb_integrate<double>(x, dx, o.xnext.data());
  mb_difference<double>(xn, o.xnext.data(), o.gap.data());
  if (!derivs) return;
  V6<double> dqb = {dx[0], dx[1], dx[2], dx[3], dx[4], dx[5]};
  M6<double> Jx6 = action_matrix(inverse(exp6(dqb))); // Jintegrate wrt x (base block)
  M6<double> Jd6 = Jexp6(dqb);                        // Jintegrate wrt dx (base block)
  SE3<double> D = mul(inverse(base_placement<double>(xn)), base_placement<double>(o.xnext.data()));
  M6<double> Jl = Jlog6(D), AdDi = action_matrix(inverse(D));
  for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) { double s = 0; for (int k = 0; k < 6; k++) s += Jl[6 * i + k] * AdDi[6 * k + j]; o.E6[6 * i + j] = -s; }
  // ddx/dx, ddx/du
  std::vector<double> Dx(n * n, 0.0), Du(n * m, 0.0);
  for (int i = 0; i < NV; i++) {
    for (int j = 0; j < n; j++) { Dx[(NV + i) * n + j] = dt * ax[i * n + j]; Dx[i * n + j] = dt * dt * ax[i * n + j]; }
    Dx[i * n + NV + i] += dt;
    for (int j = 0; j < m; j++) { Du[(NV + i) * m + j] = dt * au[i * m + j]; Du[i * m + j] = dt * dt * au[i * m + j]; }
  }
  // xnext tangent jacobians: Jint_x + Jint_dx * Dx ; Jint_dx * Du  (only the first 6 rows are non-trivial)
  std::vector<double> Fx(n * n, 0.0), Fu(n * m, 0.0);
  for (int i = 0; i < n; i++) {
    for (int j = 0; j < n; j++) {
      double s;
      if (i < 6) { s = 0; for (int k = 0; k < 6; k++) s += Jd6[6 * i + k] * Dx[k * n + j]; if (j < 6) s += Jx6[6 * i + j]; }
      else s = Dx[i * n + j] + (i == j ? 1.0 : 0.0);
      Fx[i * n + j] = s;
    }
    for (int j = 0; j < m; j++) {
      double s;
      if (i < 6) { s = 0; for (int k = 0; k < 6; k++) s += Jd6[6 * i + k] * Du[k * m + j]; }
      else s = Du[i * m + j];
      Fu[i * m + j] = s;
    }
  }
  // A = Jd2 Fx, B = Jd2 Fu with Jd2 = blockdiag(Jlog6(D), I)
  for (int i = 0; i < n; i++) {
    for (int j = 0; j < n; j++) {
      double s;
      if (i < 6) { s = 0; for (int k = 0; k < 6; k++) s += Jl[6 * i + k] * Fx[k * n + j]; } else s = Fx[i * n + j];
      o.A[i * n + j] = s;
    }
    for (int j = 0; j < m; j++) {
      double s;
      if (i < 6) { s = 0; for (int k = 0; k < 6; k++) s += Jl[6 * i + k] * Fu[k * m + j]; } else s = Fu[i * m + j];
      o.B[i * m + j] = s;
    }
  }
}

inline void eval_knot_full(const Problem &P, const mpc_knot_t &kn, const double *x, const double *u, const double *xn, bool derivs, KnotEval &o) {
  const mpc_config_t &c = P.cfg;
  const int n = 56, m = 22, nz = 78;
  const double *q = x, *v = x + NQ;
  double tau[NV];
  for (int i = 0; i < 6; i++) tau[i] = 0;
  for (int i = 0; i < NJ; i++) tau[6 + i] = u[i];
  bool active[2] = {kn.cs[0] != 0.0, kn.cs[1] != 0.0};
  if (!active[0] && !active[1]) active[0] = active[1] = true; // full:108-110 falls through to both contacts
  static thread_local CDyn<double> d;
  constrained_dynamics<double>(P.tree, c, q, v, tau, active, d);
  for (int i = 0; i < NV; i++) { o.xdot[i] = v[i]; o.xdot[NV + i] = d.a[i]; }
  for (int i = 0; i < 12; i++) o.lam[i] = d.lam[i];
  static thread_local CDynDerivs dd;
  std::vector<double> ax, au, lx_, lu_;
  if (derivs) {
    constrained_dynamics_derivatives(P.tree, c, v, d, dd);
    ax.assign(NV * n, 0.0); au.assign(NV * m, 0.0); lx_.assign(12 * n, 0.0); lu_.assign(12 * m, 0.0);
    for (int i = 0; i < NV; i++) {
      for (int j = 0; j < NV; j++) { ax[i * n + j] = dd.da_dq[i * NV + j]; ax[i * n + NV + j] = dd.da_dv[i * NV + j]; }
      for (int j = 0; j < m; j++) au[i * m + j] = dd.da_dtau[i * NV + 6 + j];
    }
    for (int i = 0; i < 12; i++) {
      for (int j = 0; j < NV; j++) { lx_[i * n + j] = dd.dl_dq[i * NV + j]; lx_[i * n + NV + j] = dd.dl_dv[i * NV + j]; }
      for (int j = 0; j < m; j++) lu_[i * m + j] = dd.dl_dtau[i * NV + 6 + j];
    }
  }
  semi_implicit_euler(c.dt, x, d.a, ax.data(), au.data(), m, xn, derivs, o);
  // ---- costs
  std::vector<double> grad(nz, 0.0);
  multibody_costs(P, d.kin, x, c.wx, c.w_cent, kn.w_lf, kn.w_rf, kn.lf_ref, kn.rf_ref, nz, o.cost, grad.data(), o.H.data());
  for (int i = 0; i < m; i++) { // control cost (full:176)
    double e = u[i] - kn.u_ref[i];
    o.cost += 0.5 * c.wu[i] * e * e; grad[n + i] += c.wu[i] * e; o.H[(n + i) * nz + n + i] += c.wu[i];
  }
  for (int f = 0; f < 2; f++) { // contact-force costs (full:187-201)
    if (kn.fcost[f] == 0.0) continue;
    double r[6];
    for (int i = 0; i < 6; i++) r[i] = d.lam[6 * f + i] - kn.f_ref[6 * f + i];
    std::vector<double> J(6 * nz, 0.0);
    if (derivs)
      for (int i = 0; i < 6; i++) { for (int j = 0; j < n; j++) J[i * nz + j] = lx_[(6 * f + i) * n + j]; for (int j = 0; j < m; j++) J[i * nz + n + j] = lu_[(6 * f + i) * m + j]; }
    add_residual_cost(r, c.w_force, 6, J.data(), nz, o.cost, grad.data(), o.H.data());
  }
  for (int i = 0; i < n; i++) o.lx[i] = grad[i];
  for (int i = 0; i < m; i++) o.lu[i] = grad[n + i];
  // ---- constraints
  for (int i = 0; i < m; i++) { // torque box (full:206-207)
    o.ctype[i] = SET_BOX; o.h[i] = u[i]; o.lo[i] = -P.rb->tau_max[i]; o.hi[i] = P.rb->tau_max[i]; o.Cu[i * m + i] = 1.0;
  }
  for (int i = 0; i < NJ; i++) { // joint box on r = neutral (-) x (full:208-209)
    int row = 22 + i;
    o.ctype[row] = SET_BOX; o.h[row] = -x[7 + i]; o.lo[row] = -P.rb->q_hi[i]; o.hi[row] = -P.rb->q_lo[i]; o.Cx[row * n + 6 + i] = -1.0;
  }
  for (int f = 0; f < 2; f++) // wrench cones on the contact forces (full:211-225)
    for (int r = 0; r < 17; r++) {
      int row = 44 + 17 * f + r;
      if (!active[f]) { o.ctype[row] = SET_NONE; continue; }
      o.ctype[row] = SET_NEG;
      double s = 0;
      for (int k = 0; k < 6; k++) s += P.Acone[6 * r + k] * d.lam[6 * f + k];
      o.h[row] = s;
      if (derivs) {
        for (int j = 0; j < n; j++) { double t = 0; for (int k = 0; k < 6; k++) t += P.Acone[6 * r + k] * lx_[(6 * f + k) * n + j]; o.Cx[row * n + j] = t; }
        for (int j = 0; j < m; j++) { double t = 0; for (int k = 0; k < 6; k++) t += P.Acone[6 * r + k] * lu_[(6 * f + k) * m + j]; o.Cu[row * m + j] = t; }
      }
    }
}

// ============================================================ kinodynamics (kino:107-173)
// u = [w_L(6), w_R(6), a_joint(22)], wrenches (f, tau) in world axes applied at the sole-frame origins.
// hdot' = [sum f_i ; sum (p_i - c) x f_i + tau_i] over active contacts (gravity is folded into F_g, see rbd.hpp).
template <class T> struct KinoVals {
  V6<T> hd;        // hdot' (without the m g term)
  T a[NV];         // generalized acceleration [a_base; a_joint]
  V3<T> p[2];      // sole origins
};
template <class T> void solve6(const T *A, const T *b, T *x) { // Gaussian elimination with partial pivoting, 6 x 6
  T M[6][7];
  for (int i = 0; i < 6; i++) { for (int j = 0; j < 6; j++) M[i][j] = A[6 * i + j]; M[i][6] = b[i]; }
  for (int c = 0; c < 6; c++) {
    int pv = c;
    for (int r = c + 1; r < 6; r++) if (std::fabs(val(M[r][c])) > std::fabs(val(M[pv][c]))) pv = r;
    if (pv != c) for (int j = 0; j < 7; j++) std::swap(M[pv][j], M[c][j]);
    for (int r = c + 1; r < 6; r++) { T f = M[r][c] / M[c][c]; for (int j = c; j < 7; j++) M[r][j] -= f * M[c][j]; }
  }
  for (int i = 5; i >= 0; i--) { T sacc = M[i][6]; for (int j = i + 1; j < 6; j++) sacc -= M[i][j] * x[j]; x[i] = sacc / M[i][i]; }
}
// centroidal momentum matrix A_g (6 x NV) from the composite inertias: column j = shift_to_com(Ic_J s_j)
template <class T> void centroidal_map(const Kin<T> &k, T *Ag) {
  for (int j = 0; j < NV; j++) {
    V6<T> Is = mul(k.Ic[body_of_dof(j)], k.S[j]);
    V3<T> l = lin(Is), aa = sub(ang(Is), cross(k.com, l));
    for (int r = 0; r < 3; r++) { Ag[r * NV + j] = l[r]; Ag[(3 + r) * NV + j] = aa[r]; }
  }
}
template <class T> void kino_dynamics(const Tree &tr, const mpc_robot_t &rb, const bool active[2], const Kin<T> &k, const T *v, const T *u, KinoVals<T> &o) {
  V3<T> fl = {T(0), T(0), T(0)}, fa = {T(0), T(0), T(0)};
  for (int i = 0; i < 2; i++) {
    SE3<T> oMf = mul(k.oM[rb.foot_body[i]], se3_cast<T>(rb.foot_place[i]));
    o.p[i] = oMf.p;
    if (!active[i]) continue;
    V3<T> f = {u[6 * i], u[6 * i + 1], u[6 * i + 2]}, t = {u[6 * i + 3], u[6 * i + 4], u[6 * i + 5]};
    fl = add(fl, f);
    fa = add(fa, add(cross(sub(oMf.p, k.com), f), t));
  }
  o.hd = mk6(fl, fa);
  // F_g(q, v, [0; a_j]) with a_world = -g, then A_b a_b = hd - F_g0
  T qdd[NV];
  for (int i = 0; i < 6; i++) qdd[i] = T(0);
  for (int i = 0; i < NJ; i++) qdd[6 + i] = u[12 + i];
  V6<T> F0 = centroidal_rate<T>(tr, k, v, qdd);
  T Ag[6 * NV], Ab[36], rhs[6], ab[6];
  centroidal_map<T>(k, Ag);
  for (int i = 0; i < 6; i++) { rhs[i] = o.hd[i] - F0[i]; for (int j = 0; j < 6; j++) Ab[6 * i + j] = Ag[i * NV + j]; }
  solve6<T>(Ab, rhs, ab);
  for (int i = 0; i < 6; i++) o.a[i] = ab[i];
  for (int i = 0; i < NJ; i++) o.a[6 + i] = u[12 + i];
}

inline void eval_knot_kino(const Problem &P, const mpc_knot_t &kn, const double *x, const double *u, const double *xn, bool derivs, KnotEval &o) {
  const mpc_config_t &c = P.cfg;
  const mpc_robot_t &rb = *P.rb;
  const int n = 56, m = 34, nz = 90;
  const double *v = x + NQ;
  bool active[2] = {kn.cs[0] != 0.0, kn.cs[1] != 0.0};
  static thread_local Kin<double> kin;
  forward_kin<double>(P.tree, x, v, kin);
  KinoVals<double> kv;
  kino_dynamics<double>(P.tree, rb, active, kin, v, u, kv);
  for (int i = 0; i < NV; i++) { o.xdot[i] = v[i]; o.xdot[NV + i] = kv.a[i]; }
  for (int i = 0; i < 12; i++) o.lam[i] = (active[i / 6]) ? u[i] : double(0.0);
  std::vector<double> ax(NV * n, 0.0), au(NV * m, 0.0), hdq(6 * NV, 0.0), hdu(6 * m, 0.0);
  FootKin fk[2];
  for (int f = 0; f < 2; f++) foot_kin(P.tree, kin, f, fk[f]);
  double Jc[3 * NV];
  com_jacobian(kin, Jc);
  if (derivs) {
    // d hdot'/dq (angular rows) and d hdot'/du
    for (int i = 0; i < 2; i++) {
      if (!active[i]) continue;
      V3<double> f = {u[6 * i], u[6 * i + 1], u[6 * i + 2]};
      M3<double> px = skew(sub(kv.p[i], kin.com));
      for (int r = 0; r < 3; r++) {
        hdu[r * m + 6 * i + r] = 1.0; hdu[(3 + r) * m + 6 * i + 3 + r] = 1.0;
        for (int cc = 0; cc < 3; cc++) hdu[(3 + r) * m + 6 * i + cc] = px[3 * r + cc];
      }
      int fb = rb.foot_body[i];
      for (int j = 0; j < NV; j++) {
        V3<double> dp = {0, 0, 0};
        if (P.tree.anc[body_of_dof(j)][fb]) dp = add(lin(kin.S[j]), cross(ang(kin.S[j]), kv.p[i]));
        V3<double> dc = {Jc[j], Jc[NV + j], Jc[2 * NV + j]};
        V3<double> d = cross(sub(dp, dc), f);
        for (int r = 0; r < 3; r++) hdq[(3 + r) * NV + j] += d[r];
      }
    }
    // d F_g / d(q, v) at fixed a
    V6<double> a[NB], Fs[NB];
    double tau[NV];
    rnea<double>(P.tree, kin, v, kv.a, nullptr, tau, a, Fs);
    double dFq[6 * NV], dFv[6 * NV], Ag[6 * NV], Ab[36];
    total_force_derivatives(P.tree, kin, a, Fs, dFq, dFv);
    centroidal_map<double>(kin, Ag);
    for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) Ab[6 * i + j] = Ag[i * NV + j];
    V3<double> Fl = lin(Fs[0]);
    for (int j = 0; j < n + m; j++) {
      double rhs[6] = {0, 0, 0, 0, 0, 0}, sol[6];
      if (j < NV) {
        V3<double> dl = {dFq[j], dFq[NV + j], dFq[2 * NV + j]}, da = {dFq[3 * NV + j], dFq[4 * NV + j], dFq[5 * NV + j]};
        V3<double> dc = {Jc[j], Jc[NV + j], Jc[2 * NV + j]};
        V3<double> dag = sub(sub(da, cross(dc, Fl)), cross(kin.com, dl));
        for (int r = 0; r < 3; r++) { rhs[r] = hdq[r * NV + j] - dl[r]; rhs[3 + r] = hdq[(3 + r) * NV + j] - dag[r]; }
      } else if (j < n) {
        int jj = j - NV;
        V3<double> dl = {dFv[jj], dFv[NV + jj], dFv[2 * NV + jj]}, da = {dFv[3 * NV + jj], dFv[4 * NV + jj], dFv[5 * NV + jj]};
        V3<double> dag = sub(da, cross(kin.com, dl));
        for (int r = 0; r < 3; r++) { rhs[r] = -dl[r]; rhs[3 + r] = -dag[r]; }
      } else {
        int jj = j - n;
        if (jj < 12) for (int r = 0; r < 6; r++) rhs[r] = hdu[r * m + jj];
        else for (int r = 0; r < 6; r++) rhs[r] = -Ag[r * NV + 6 + (jj - 12)];
      }
      solve6<double>(Ab, rhs, sol);
      for (int r = 0; r < 6; r++) { if (j < n) ax[r * n + j] = sol[r]; else au[r * m + (j - n)] = sol[r]; }
    }
    for (int i = 0; i < NJ; i++) au[(6 + i) * m + 12 + i] = 1.0;
  }
  semi_implicit_euler(c.dt, x, kv.a, ax.data(), au.data(), m, xn, derivs, o);
  // ---- costs: state, control, centroidal momentum, its derivative, foot poses (kino:137-157)
  std::vector<double> grad(nz, 0.0);
  multibody_costs(P, kin, x, c.wx, c.w_cent, kn.w_lf, kn.w_rf, kn.lf_ref, kn.rf_ref, nz, o.cost, grad.data(), o.H.data());
  for (int i = 0; i < m; i++) {
    double e = u[i] - kn.u_ref[i];
    o.cost += 0.5 * c.wu[i] * e * e; grad[n + i] += c.wu[i] * e; o.H[(n + i) * nz + n + i] += c.wu[i];
  }
  {
    double r[6];
    for (int i = 0; i < 3; i++) { r[i] = kv.hd[i] + kin.mass * rb.gravity[i]; r[3 + i] = kv.hd[3 + i]; }
    std::vector<double> J(6 * nz, 0.0);
    for (int i = 0; i < 6; i++) { for (int j = 0; j < NV; j++) J[i * nz + j] = hdq[i * NV + j]; for (int j = 0; j < m; j++) J[i * nz + n + j] = hdu[i * m + j]; }
    add_residual_cost(r, c.w_centder, 6, J.data(), nz, o.cost, grad.data(), o.H.data());
  }
  for (int i = 0; i < n; i++) o.lx[i] = grad[i];
  for (int i = 0; i < m; i++) o.lu[i] = grad[n + i];
  // ---- constraints: joint box, then per contact cone(17) + zero LOCAL frame velocity(6) (kino:161-171)
  for (int i = 0; i < NJ; i++) {
    o.ctype[i] = SET_BOX; o.h[i] = -x[7 + i]; o.lo[i] = -rb.q_hi[i]; o.hi[i] = -rb.q_lo[i]; o.Cx[i * n + 6 + i] = -1.0;
  }
  for (int f = 0; f < 2; f++) {
    int base = 22 + 23 * f, fb = rb.foot_body[f];
    for (int r = 0; r < 23; r++) o.ctype[base + r] = active[f] ? (r < 17 ? SET_NEG : SET_EQ) : SET_NONE;
    if (!active[f]) continue;
    for (int r = 0; r < 17; r++) {
      double sacc = 0;
      for (int k2 = 0; k2 < 6; k2++) { sacc += P.Acone[6 * r + k2] * u[6 * f + k2]; o.Cu[(base + r) * m + 6 * f + k2] = P.Acone[6 * r + k2]; }
      o.h[base + r] = sacc;
    }
    V6<double> vl = actinv_motion(fk[f].oMf, kin.v[fb]);
    for (int r = 0; r < 6; r++) o.h[base + 17 + r] = vl[r];
    if (derivs)
      for (int j = 0; j < NV; j++) {
        if (!P.tree.anc[body_of_dof(j)][fb]) continue;
        int J = body_of_dof(j), pJ = rb.parent[J];
        V6<double> vp = pJ >= 0 ? kin.v[pJ] : zero6<double>();
        V6<double> wl = actinv_motion(fk[f].oMf, cross_mm(kin.S[j], vp));
        for (int r = 0; r < 6; r++) { o.Cx[(base + 17 + r) * n + j] = -wl[r]; o.Cx[(base + 17 + r) * n + NV + j] = fk[f].J[r * NV + j]; }
      }
  }
}

// terminal cost + constraint. Outputs into o: cost, lx, H (n x n block used, stored with stride n+m), h[0:3], Cx rows 0..2
inline void eval_term(const Problem &P, const mpc_term_t &tm, const double *x, KnotEval &o) {
  const mpc_config_t &c = P.cfg;
  o.zero();
  if (c.kind == MPC_KIND_CENT) return; // empty CostStack, no constraint (cent:249,261)
  const int n = 56, nz = o.n + o.m;
  static thread_local Kin<double> kin;
  forward_kin<double>(P.tree, x, x + NQ, kin);
  std::vector<double> grad(nz, 0.0);
  bool anyc = false;
  for (int i = 0; i < 56; i++) anyc |= c.wx_term[i] != 0.0;
  for (int i = 0; i < 6; i++) anyc |= (c.w_cent_term[i] != 0.0) || (c.w_foot_term[i] != 0.0);
  if (anyc)
    multibody_costs(P, kin, x, c.wx_term, c.w_cent_term, c.w_foot_term, c.w_foot_term, tm.lf_ref, tm.rf_ref, nz, o.cost, grad.data(), o.H.data());
  for (int i = 0; i < n; i++) o.lx[i] = grad[i];
  if (tm.has_com_cstr != 0.0) { // CoM equality (full:499-507, kino:176-180)
    double Jc[3 * NV];
    com_jacobian(kin, Jc);
    for (int r = 0; r < 3; r++) {
      o.ctype[r] = SET_EQ; o.h[r] = kin.com[r] - tm.com_ref[r];
      for (int j = 0; j < NV; j++) o.Cx[r * n + j] = Jc[r * NV + j];
    }
  }
}

inline void eval_knot(const Problem &P, const mpc_knot_t &kn, const double *x, const double *u, const double *xn, bool derivs, KnotEval &o) {
  o.zero();
  if (P.cfg.kind == MPC_KIND_CENT) eval_knot_cent(P, kn, x, u, xn, derivs, o);
  else if (P.cfg.kind == MPC_KIND_KINO) eval_knot_kino(P, kn, x, u, xn, derivs, o);
  else eval_knot_full(P, kn, x, u, xn, derivs, o);
}

} // namespace orc
