// CPU ORACLE (test infrastructure, never shipped, never imported by the product package) — dense QP solver.
//
// Restates the boundary `proxsuite.proxqp.dense.QP.solve()` as the reference drives it (QP_utils.py:500-513,557-573 and the
// other solver classes of that file).  proxsuite is an absent pip dependency: PARITY UNPINNED against it.  What is
// restated is the published ProxQP algorithm (Bambade, El-Kazdadi, Taylor, Carpentier, RSS 2022):
//   outer loop  : bound-constrained-Lagrangian (BCL) update of the multipliers (y, z) or of the penalties (mu_eq, mu_in);
//   inner loop  : semismooth Newton on the proximal augmented Lagrangian in x, exact linesearch on the piecewise quadratic;
//   stop        : primal / dual residual (and duality gap) <= eps_abs + eps_rel * scale.
// Pinned instead by (tests/test_qp.py): KKT conditions of the returned point, agreement with an independent active-set
// enumeration on small problems and with scipy on the reference-shaped whole-body QPs.
// The Newton system is condensed to the primal block  K = H + rho I + A'A / mu_eq + C_act' C_act / mu_in  (SPD for rho > 0), which is
// the Schur complement of the primal-dual KKT matrix ProxQP factors with `DenseBackend.PrimalDualLDLT`: same step in exact arithmetic.
// Consequence of the condensed form: multiplier estimates are residual / mu, so the dual residual cannot go below ~ eps_machine |A||x| / mu;
// the penalty floors default to 1e-4 (proxsuite: 1e-9 / 1e-8), which leaves a floor of ~1e-7 on the whole-body QPs — the reference asks 1e-3.
// Not restated: Ruiz equilibration (the reference calls update(..., update_preconditioner=False); with it the iterates differ, the
// fixed point does not) and `primal_infeasibility_solving` (closest-feasible QP; only matters for infeasible problems).
#pragma once
#include "../include/mpcqp_b200.h"
#include <cmath>
#include <vector>

namespace orc {

inline void qp_default_settings(mpc_qp_settings_t &s) {
  s.eps_abs = 1e-5; s.eps_rel = 0.0; s.rho = 1e-6; s.mu_eq = 1e-3; s.mu_in = 1e-1; s.alpha_bcl = 0.1; s.beta_bcl = 0.9;
  s.mu_update_factor = 0.1; s.mu_min_eq = 1e-4; s.mu_min_in = 1e-4; s.max_iter = 10000; s.max_iter_in = 1500; s.check_duality_gap = 0;
  s.warm_start = 0;
}

struct QP {
  int n, ne, ni; // ni = general inequality rows; box rows (identity) follow when box
  bool box;
  const double *H, *g, *A, *b, *C, *l, *u, *lb, *ub;
  int nz() const { return ni + (box ? n : 0); }
  static bool inf(double v) { return std::fabs(v) >= 1e20; }
  // row i of the stacked inequality operator [C; I] applied to v
  double row(int i, const double *v) const {
    if (i >= ni) return v[i - ni];
    double s = 0; for (int j = 0; j < n; j++) s += C[i * n + j] * v[j];
    return s;
  }
  double lo(int i) const { return i < ni ? l[i] : lb[i - ni]; }
  double up(int i) const { return i < ni ? u[i] : ub[i - ni]; }
};

struct QPResiduals { double pri, dua, gap, obj, pri_scale, dua_scale, gap_scale; };

inline QPResiduals qp_residuals(const QP &q, const double *x, const double *y, const double *z) {
  const int n = q.n, nz = q.nz();
  QPResiduals r{};
  std::vector<double> Hx(n, 0.0), d(n, 0.0);
  double xHx = 0, gx = 0, nAx = 0, nCx = 0, nHx = 0, nAty = 0, nCtz = 0, ng = 0;
  for (int i = 0; i < n; i++) { for (int j = 0; j < n; j++) Hx[i] += q.H[i * n + j] * x[j]; xHx += x[i] * Hx[i]; gx += q.g[i] * x[i]; }
  for (int i = 0; i < n; i++) { d[i] = Hx[i] + q.g[i]; nHx = std::fmax(nHx, std::fabs(Hx[i])); ng = std::fmax(ng, std::fabs(q.g[i])); }
  double by = 0;
  std::vector<double> Aty(n, 0.0), Ctz(n, 0.0);
  for (int r_ = 0; r_ < q.ne; r_++) {
    double s = -q.b[r_];
    for (int j = 0; j < n; j++) { s += q.A[r_ * n + j] * x[j]; Aty[j] += q.A[r_ * n + j] * y[r_]; }
    r.pri = std::fmax(r.pri, std::fabs(s)); nAx = std::fmax(nAx, std::fabs(s + q.b[r_])); by += q.b[r_] * y[r_];
  }
  double bz = 0;
  for (int i = 0; i < nz; i++) {
    const double s = q.row(i, x);
    nCx = std::fmax(nCx, std::fabs(s));
    double viol = 0;
    if (!QP::inf(q.up(i))) viol += std::fmax(s - q.up(i), 0.0);
    if (!QP::inf(q.lo(i))) viol += std::fmin(s - q.lo(i), 0.0);
    r.pri = std::fmax(r.pri, std::fabs(viol));
    if (i < q.ni) for (int j = 0; j < n; j++) Ctz[j] += q.C[i * n + j] * z[i];
    else Ctz[i - q.ni] += z[i];
    if (z[i] > 0 && !QP::inf(q.up(i))) bz += q.up(i) * z[i];
    if (z[i] < 0 && !QP::inf(q.lo(i))) bz += q.lo(i) * z[i];
  }
  for (int i = 0; i < n; i++) {
    d[i] += Aty[i] + Ctz[i];
    r.dua = std::fmax(r.dua, std::fabs(d[i])); nAty = std::fmax(nAty, std::fabs(Aty[i])); nCtz = std::fmax(nCtz, std::fabs(Ctz[i]));
  }
  r.gap = xHx + gx + by + bz;
  r.obj = 0.5 * xHx + gx;
  r.pri_scale = std::fmax(nAx, nCx);
  r.dua_scale = std::fmax(std::fmax(nHx, ng), std::fmax(nAty, nCtz));
  r.gap_scale = std::fmax(std::fmax(std::fabs(xHx), std::fabs(gx)), std::fmax(std::fabs(by), std::fabs(bz)));
  return r;
}

// plain (unblocked) Cholesky + solve: an independent code path from the kernels' blocked routine
inline bool qp_chol_solve(std::vector<double> &K, int n, std::vector<double> &rhs) {
  for (int j = 0; j < n; j++) {
    double d = K[j * n + j];
    for (int k = 0; k < j; k++) d -= K[j * n + k] * K[j * n + k];
    if (!(d > 0) || !std::isfinite(d)) return false;
    d = std::sqrt(d); K[j * n + j] = d;
    for (int i = j + 1; i < n; i++) {
      double s = K[i * n + j];
      for (int k = 0; k < j; k++) s -= K[i * n + k] * K[j * n + k];
      K[i * n + j] = s / d;
    }
  }
  for (int i = 0; i < n; i++) { double s = rhs[i]; for (int k = 0; k < i; k++) s -= K[i * n + k] * rhs[k]; rhs[i] = s / K[i * n + i]; }
  for (int i = n - 1; i >= 0; i--) { double s = rhs[i]; for (int k = i + 1; k < n; k++) s -= K[k * n + i] * rhs[k]; rhs[i] = s / K[i * n + i]; }
  return true;
}

inline int qp_solve(const QP &q, const mpc_qp_settings_t &st, double *x, double *y, double *z, mpc_qp_info_t *info) {
  const int n = q.n, ne = q.ne, nz = q.nz();
  if (!st.warm_start) { for (int i = 0; i < n; i++) x[i] = 0; for (int i = 0; i < ne; i++) y[i] = 0; for (int i = 0; i < nz; i++) z[i] = 0; }
  double mue = st.mu_eq, mui = st.mu_in;
  const double eta_ext_init = std::pow(0.1, st.alpha_bcl), eps_in_min = std::fmin(st.eps_abs, 1e-9);
  double eta_ext = eta_ext_init, eta_in = 1.0;
  std::vector<double> xe(n), ye(ne), ze(nz), re(ne), su(nz), sl(nz), grad(n), dx(n), K(n * n), Adx(ne), cd(nz), AtA(n * n, 0.0);
  // A x, [C; I] x and H x are computed once and then carried along every accepted step (W (x + alpha dx) = W x + alpha W dx, the products with
  // dx being needed by the linesearch anyway): one O(n^2) sweep less per Newton step, and less cancellation noise in the residuals than
  // re-evaluating them at a nearly converged x
  std::vector<double> Ax(ne, 0.0), Sx(nz, 0.0), Hx(n, 0.0), Hdx(n, 0.0);
  for (int r = 0; r < ne; r++) for (int j = 0; j < n; j++) Ax[r] += q.A[r * n + j] * x[j];
  for (int i = 0; i < nz; i++) Sx[i] = q.row(i, x);
  for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) Hx[i] += q.H[i * n + j] * x[j];
  for (int r = 0; r < ne; r++) for (int i = 0; i < n; i++) { const double a = q.A[r * n + i]; if (a != 0) for (int j = 0; j < n; j++) AtA[i * n + j] += a * q.A[r * n + j]; }
  int status = 1, it = 0, it_in = 0, mu_updates = 0;
  QPResiduals R{};
  for (;; it++) {
    R = qp_residuals(q, x, y, z);
    if (!std::isfinite(R.pri) || !std::isfinite(R.dua)) { status = 2; break; }
    const bool ok = R.pri <= st.eps_abs + st.eps_rel * R.pri_scale && R.dua <= st.eps_abs + st.eps_rel * R.dua_scale &&
                    (!st.check_duality_gap || std::fabs(R.gap) <= st.eps_abs + st.eps_rel * R.gap_scale);
    if (ok) { status = 0; break; }
    if (it >= st.max_iter) break;
    xe.assign(x, x + n); ye.assign(y, y + ne); ze.assign(z, z + nz);
    // ---- inner loop: semismooth Newton on phi(x) = 1/2 x'Hx + g'x + rho/2 |x - xe|^2 + |A x - b + mu_e ye|^2 / (2 mu_e)
    //                                               + (|[s - u + mu_i ze]+|^2 + |[s - l + mu_i ze]-|^2) / (2 mu_i),  s = [C; I] x
    bool failed = false;
    for (int in = 0;; in++) {
      for (int r = 0; r < ne; r++) re[r] = Ax[r] - q.b[r] + mue * ye[r];
      for (int i = 0; i < nz; i++) {
        const double s = Sx[i];
        su[i] = QP::inf(q.up(i)) ? -1e300 : s - q.up(i) + mui * ze[i];
        sl[i] = QP::inf(q.lo(i)) ? 1e300 : s - q.lo(i) + mui * ze[i];
      }
      for (int i = 0; i < n; i++) {
        double s = q.g[i] + st.rho * (x[i] - xe[i]) + Hx[i];
        for (int r = 0; r < ne; r++) s += q.A[r * n + i] * re[r] / mue;
        grad[i] = s;
      }
      for (int i = 0; i < nz; i++) {
        const double t = (std::fmax(su[i], 0.0) + std::fmin(sl[i], 0.0)) / mui;
        if (t == 0) continue;
        if (i < q.ni) for (int j = 0; j < n; j++) grad[j] += q.C[i * n + j] * t; else grad[i - q.ni] += t;
      }
      double gn = 0; for (int i = 0; i < n; i++) gn = std::fmax(gn, std::fabs(grad[i]));
      if (!std::isfinite(gn)) { failed = true; break; }
      if (gn <= eta_in || in >= st.max_iter_in) break;
      // Newton matrix on the current active set
      for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) K[i * n + j] = q.H[i * n + j] + AtA[i * n + j] / mue + (i == j ? st.rho : 0.0);
      for (int i = 0; i < nz; i++) {
        if (!(su[i] > 0 || sl[i] < 0)) continue;
        if (i < q.ni) { for (int a = 0; a < n; a++) { const double ca = q.C[i * n + a]; if (ca != 0) for (int c = 0; c < n; c++) K[a * n + c] += ca * q.C[i * n + c] / mui; } }
        else K[(i - q.ni) * n + (i - q.ni)] += 1.0 / mui;
      }
      for (int i = 0; i < n; i++) dx[i] = -grad[i];
      if (!qp_chol_solve(K, n, dx)) { failed = true; break; }
      it_in++;
      // exact linesearch: phi'(alpha) = b0 + a0 alpha + sum_i c_i ([su_i + alpha c_i]+ + [sl_i + alpha c_i]-) / mu_i, nondecreasing piecewise linear
      double a0 = 0, b0 = 0;
      for (int r = 0; r < ne; r++) { double s = 0; for (int j = 0; j < n; j++) s += q.A[r * n + j] * dx[j]; Adx[r] = s; a0 += s * s / mue; b0 += s * re[r] / mue; }
      for (int i = 0; i < n; i++) {
        double hd = 0;
        for (int j = 0; j < n; j++) hd += q.H[i * n + j] * dx[j];
        Hdx[i] = hd;
        a0 += dx[i] * (hd + st.rho * dx[i]);
        b0 += dx[i] * (Hx[i] + q.g[i] + st.rho * (x[i] - xe[i]));
      }
      for (int i = 0; i < nz; i++) cd[i] = q.row(i, dx.data());
      auto dphi = [&](double t) {
        double s = b0 + a0 * t;
        for (int i = 0; i < nz; i++) s += cd[i] * (std::fmax(su[i] + t * cd[i], 0.0) + std::fmin(sl[i] + t * cd[i], 0.0)) / mui;
        return s;
      };
      double lo = 0.0, hi = INFINITY;
      for (int i = 0; i < nz; i++) {
        if (cd[i] == 0) continue;
        for (int w = 0; w < 2; w++) {
          const double sv = w ? sl[i] : su[i];
          if (std::fabs(sv) >= 1e299) continue;
          const double t = -sv / cd[i];
          if (!(t > 0)) continue;
          if (dphi(t) < 0) lo = std::fmax(lo, t); else hi = std::fmin(hi, t);
        }
      }
      const double tm = std::isfinite(hi) ? 0.5 * (lo + hi) : lo + 1.0;
      double slope = a0, icpt = b0;
      for (int i = 0; i < nz; i++) {
        if (su[i] + tm * cd[i] > 0) { slope += cd[i] * cd[i] / mui; icpt += cd[i] * su[i] / mui; }
        if (sl[i] + tm * cd[i] < 0) { slope += cd[i] * cd[i] / mui; icpt += cd[i] * sl[i] / mui; }
      }
      double alpha = -icpt / slope;
      alpha = std::fmin(std::fmax(alpha, lo), hi);
      if (!std::isfinite(alpha)) { failed = true; break; }
      double step = 0, xn = 1.0;
      for (int i = 0; i < n; i++) { step = std::fmax(step, std::fabs(alpha * dx[i])); xn = std::fmax(xn, std::fabs(x[i])); x[i] += alpha * dx[i]; Hx[i] += alpha * Hdx[i]; }
      for (int r = 0; r < ne; r++) Ax[r] += alpha * Adx[r];
      for (int i = 0; i < nz; i++) Sx[i] += alpha * cd[i];
      if (step <= 1e-14 * xn) break; // the Newton step is below the rounding level of x: the inner tolerance is not reachable in fp64
    }
    if (failed) { status = 2; break; }
    // ---- multiplier estimates at the inner solution and the BCL test
    double pri_new = 0;
    for (int r = 0; r < ne; r++) { const double s = Ax[r] - q.b[r]; y[r] = ye[r] + s / mue; pri_new = std::fmax(pri_new, std::fabs(s)); }
    for (int i = 0; i < nz; i++) {
      const double s = Sx[i];
      double viol = 0;
      if (!QP::inf(q.up(i))) viol += std::fmax(s - q.up(i), 0.0);
      if (!QP::inf(q.lo(i))) viol += std::fmin(s - q.lo(i), 0.0);
      pri_new = std::fmax(pri_new, std::fabs(viol));
      const double zu = QP::inf(q.up(i)) ? 0.0 : std::fmax(ze[i] + (s - q.up(i)) / mui, 0.0);
      const double zl = QP::inf(q.lo(i)) ? 0.0 : std::fmin(ze[i] + (s - q.lo(i)) / mui, 0.0);
      z[i] = zu + zl;
    }
    if (pri_new <= eta_ext) {
      eta_ext *= std::pow(mui, st.beta_bcl);
      eta_in = std::fmax(eta_in * mui, eps_in_min);
    } else {
      for (int r = 0; r < ne; r++) y[r] = ye[r];
      for (int i = 0; i < nz; i++) z[i] = ze[i];
      const double nmui = std::fmax(mui * st.mu_update_factor, st.mu_min_in), nmue = std::fmax(mue * st.mu_update_factor, st.mu_min_eq);
      if (nmui != mui || nmue != mue) mu_updates++;
      mui = nmui; mue = nmue;
      eta_ext = eta_ext_init * std::pow(mui, st.alpha_bcl);
      eta_in = std::fmax(mui, eps_in_min);
    }
  }
  if (info) {
    info->status = status; info->iter = it; info->iter_in = it_in; info->mu_updates = mu_updates;
    info->pri_res = R.pri; info->dua_res = R.dua; info->duality_gap = R.gap; info->objective = R.obj;
  }
  return status;
}

// IDSolver_ulim.computeMatrice (QP_utils.py:514-552) for nv = 28, nk = 2, force_size = 6: A [40][62], b [40], C [18][62], l [18].
inline void qp_assemble_id(const double *M, const double *nle, const double *Jc, const double *gamma, const double *a, const double *forces,
                           const int32_t *cs, double mu, double L, double W, double *A, double *b, double *C, double *l) {
  const int nv = 28, nk = 2, fs = 6, nf = nk * fs, n = 2 * nv - 6 + nf, ne = nv + nf;
  for (int i = 0; i < ne * n; i++) A[i] = 0;
  for (int i = 0; i < 9 * nk * n; i++) C[i] = 0;
  for (int i = 0; i < 9 * nk; i++) l[i] = 0;
  for (int i = 0; i < nv; i++) {
    double s = -nle[i];
    for (int j = 0; j < nv; j++) { A[i * n + j] = M[i * nv + j]; s -= M[i * nv + j] * a[j]; }
    for (int k = 0; k < nf; k++) if (cs[k / fs]) { A[i * n + nv + k] = -Jc[k * nv + i]; s += Jc[k * nv + i] * forces[k]; }
    if (i >= 6) A[i * n + nv + nf + i - 6] = -1.0; // -S (QP_utils.py:460-461,531)
    b[i] = s;
  }
  for (int k = 0; k < nf; k++) {
    double s = 0;
    if (cs[k / fs]) { s = -gamma[k]; for (int j = 0; j < nv; j++) { A[(nv + k) * n + j] = Jc[k * nv + j]; s -= Jc[k * nv + j] * a[j]; } }
    b[nv + k] = s;
  }
  const double Cmin[9][6] = {{-1, 0, mu, 0, 0, 0}, {1, 0, mu, 0, 0, 0}, {-1, 0, mu, 0, 0, 0}, {1, 0, mu, 0, 0, 0}, {0, 0, 1, 0, 0, 0},
                             {0, 0, W, -1, 0, 0}, {0, 0, W, 1, 0, 0},   {0, 0, L, 0, -1, 0},  {0, 0, L, 0, 1, 0}}; // QP_utils.py:474-484 (as written there)
  for (int i = 0; i < nk; i++) {
    if (!cs[i]) continue;
    const double *f = forces + i * fs;
    const double lv[9] = {f[0] - f[2] * mu, -f[0] - f[2] * mu, f[1] - f[2] * mu, -f[1] - f[2] * mu, -f[2],
                          f[3] - f[2] * W,  -f[3] - f[2] * W,  f[4] - f[2] * L,  -f[4] - f[2] * L}; // QP_utils.py:538-548
    for (int r = 0; r < 9; r++) { l[i * 9 + r] = lv[r]; for (int c = 0; c < fs; c++) C[(i * 9 + r) * n + nv + i * fs + c] = Cmin[r][c]; }
  }
}

} // namespace orc
