// ORACLE (test infrastructure, not product code) — ProxDDP (primal-dual augmented-Lagrangian DDP).
//
// Restates aligator::SolverProxDDPTpl<double>::run as the reference drives it
// (fulldynamic_talos.py:374-397 cold solve; :407,538-540 one-iteration MPC tick): ROLLOUT_LINEAR,
// force_initial_condition, tol 1e-5, mu_init 1e-8.  Aligator is a pip dependency absent from
// /root/reference (README.md:10) => PARITY UNPINNED.  The algorithm below follows the ProxDDP paper
// (Jallet et al.) and SURVEY App. A6; every constant that could not be checked against upstream is
// listed in DESIGN.md ("solver constants").
#pragma once
#include "knot.hpp"
#include "riccati.hpp"
#include <algorithm>
#include <cmath>
#include <limits>
#include <cstdio>
#include <cstdlib>
#include <chrono>

namespace orc {

struct SolverParams {
  double tol = 1e-5, mu_init = 1e-8;
  int max_iters = 100, max_al_iters = 100;
  double prim_alpha = 0.1, prim_beta = 0.9, dual_alpha = 1.0, dual_beta = 1.0;
  double mu_update_factor = 0.01, mu_lower_bound = 1e-8;
  double reg_init = 1e-9, reg_min = 1e-10, reg_max = 1e9, reg_inc = 10.0, reg_dec = 1.0 / 3.0;
  double ls_c1 = 1e-4, ls_alpha_min = 1e-7, ls_contr_min = 0.5, ls_contr_max = 0.8;
  int ls_max_steps = 20;
  double ls_dphi_rel = 1e-11; // predicted decrease below the fp64 resolution of the merit: Armijo is meaningless, take the step
  bool par_knots = false; // OpenMP over knots (reference: setNumThreads(8), fulldynamic_talos.py:385)
  // ---- ablation switches (tools/closed_loop_oracle.py; defaults = the normative algorithm the CUDA path mirrors).  They bound the
  // solver details that could not be checked against upstream Aligator (DESIGN "oracle-vs-Aligator ablation"):
  double mu_dyn_scale = 1.0;  // dynamics penalty mu_d = mu_dyn_scale * mu (upstream had a `dyn_al_scale` in some versions)
  int ls_mode = 0;            // 0 Armijo backtracking, 1 non-monotone Armijo (reference value = max of the last ls_window merits), 2 always alpha = 1
  int ls_window = 5;
  double dual_weight = 1.0;   // weight of the multiplier-distance terms of the PDAL merit
  int rollout = 0;            // 0 ROLLOUT_LINEAR (fulldynamic_talos.py:381 and every other reference script), 1 ROLLOUT_NONLINEAR
};

struct Instance {
  int T;
  const mpc_knot_t *knots; // [T]
  mpc_term_t term;
  std::vector<double> x0;
};

// normal-cone projection pieces for one row: returns v_plus; act = 1 if the row is active
inline double vplus_row(int type, double h, double ve, double mu, double lo, double hi, int &act) {
  double w = h + mu * ve;
  switch (type) {
  case SET_EQ: act = 1; return w / mu;
  case SET_NEG: if (w > 0) { act = 1; return w / mu; } act = 0; return 0.0;
  case SET_BOX:
    if (w > hi) { act = 1; return (w - hi) / mu; }
    if (w < lo) { act = 1; return (w - lo) / mu; }
    act = 0; return 0.0;
  default: act = 0; return 0.0;
  }
}
// h - Pi_C(h + mu ve): primal infeasibility of one row
inline double prim_row(int type, double h, double ve, double mu, double lo, double hi) {
  double w = h + mu * ve, pc;
  switch (type) {
  case SET_EQ: pc = 0; break;
  case SET_NEG: pc = std::min(w, 0.0); break;
  case SET_BOX: pc = std::min(std::max(w, lo), hi); break;
  default: return 0.0;
  }
  return h - pc;
}

struct Solver {
  const Problem &P;
  SolverParams prm;
  Dims d;
  int T = 0;
  // iterate
  std::vector<double> xs, us, vs, lams, vs_prev, lams_prev;
  std::vector<double> Kfb; // T x m x n feedback gains (controlFeedbacks)
  // workspace
  std::vector<KnotEval> ev; // T knots + terminal
  std::vector<KnotEval> tr; // trial evaluations
  std::vector<double> txs, tus, tvs, tlams;
  LQSolution sol;
  std::vector<double> gq, fbar, dbar, Cact, Dact, vplus, lplus; // per-knot LQ vectors
  std::vector<int> act;
  // results
  double mu = 0, inner_tol = 1, prim_tol = 1, preg = 0;
  double prim_infeas = 0, dual_infeas = 0, inner_crit = 0, traj_cost = 0, merit = 0;
  int num_iters = 0, al_iters = 0, conv = 0, status = 1, ls_evals = 0;
  double alpha_last = 0;
  std::vector<double> alphas;

  std::vector<double> merit_hist; // non-monotone linesearch memory (ls_mode 1); survives across run() calls of one Solver
  Solver(const Problem &p, const SolverParams &s) : P(p), prm(s), d(p.d) {}
  double mud() const { return mu * prm.mu_dyn_scale; }

  void setup(int T_) {
    T = T_;
    xs.assign((size_t)(T + 1) * d.nx, 0); us.assign((size_t)T * d.m, 0);
    vs.assign((size_t)(T + 1) * d.nc, 0); lams.assign((size_t)(T + 1) * d.n, 0);
    vs_prev = vs; lams_prev = lams;
    Kfb.assign((size_t)T * d.m * d.n, 0);
    ev.resize(T + 1); tr.resize(T + 1);
    for (auto &e : ev) e.resize(d);
    for (auto &e : tr) e.resize(d);
    txs = xs; tus = us; tvs = vs; tlams = lams;
    int nz = d.n + d.m;
    gq.assign((size_t)(T + 1) * nz, 0); fbar.assign((size_t)T * d.n, 0); dbar.assign((size_t)(T + 1) * d.nc, 0);
    Cact.assign((size_t)(T + 1) * d.nc * d.n, 0); Dact.assign((size_t)T * d.nc * d.m, 0);
    vplus.assign((size_t)(T + 1) * d.nc, 0); lplus.assign((size_t)(T + 1) * d.n, 0); act.assign((size_t)(T + 1) * d.nc, 0);
  }

  void integrate_state(const double *x, const double *dx, double alpha, double *out) const {
    if (P.cfg.kind == MPC_KIND_CENT) { for (int i = 0; i < d.n; i++) out[i] = x[i] + alpha * dx[i]; return; }
    double s[56]; for (int i = 0; i < 56; i++) s[i] = alpha * dx[i];
    mb_integrate<double>(x, s, out);
  }

  void evaluate(const Instance &in, const double *X, const double *U, bool derivs, std::vector<KnotEval> &E) {
#pragma omp parallel for schedule(dynamic) if (prm.par_knots)
    for (int k = 0; k <= T; k++) {
      if (k < T) eval_knot(P, in.knots[k], X + (size_t)k * d.nx, U + (size_t)k * d.m, X + (size_t)(k + 1) * d.nx, derivs, E[k]);
      else eval_term(P, in.term, X + (size_t)T * d.nx, E[T]);
    }
  }

  // PDAL merit at (E, V, L) with fixed estimates (vs_prev, lams_prev)
  double merit_value(const std::vector<KnotEval> &E, const double *V, const double *L, double *cost_out) const {
    double cost = 0, pen = 0;
    for (int k = 0; k <= T; k++) {
      cost += E[k].cost;
      for (int r = 0; r < d.nc; r++) {
        int a;
        if (E[k].ctype[r] == SET_NONE) continue;
        double vp = vplus_row(E[k].ctype[r], E[k].h[r], vs_prev[(size_t)k * d.nc + r], mu, E[k].lo[r], E[k].hi[r], a);
        double dv = vp - V[(size_t)k * d.nc + r];
        pen += 0.5 * mu * (vp * vp + prm.dual_weight * dv * dv);
      }
      if (k < T)
        for (int i = 0; i < d.n; i++) {
          double lp = lams_prev[(size_t)(k + 1) * d.n + i] + E[k].gap[i] / mud();
          double dl = lp - L[(size_t)(k + 1) * d.n + i];
          pen += 0.5 * mud() * (lp * lp + prm.dual_weight * dl * dl);
        }
    }
    if (cost_out) *cost_out = cost;
    return cost + pen;
  }

  // multipliers, Lagrangian gradients, infeasibilities, LQ right-hand sides at the current iterate
  void assemble() {
    const int n = d.n, m = d.m, nc = d.nc, nz = n + m;
    prim_infeas = 0; dual_infeas = 0; inner_crit = 0;
    for (int k = 0; k <= T; k++) {
      KnotEval &e = ev[k];
      double *g = &gq[(size_t)k * nz];
      for (int i = 0; i < n; i++) g[i] = e.lx[i];
      for (int i = 0; i < m; i++) g[n + i] = (k < T) ? e.lu[i] : 0.0;
      // constraints
      for (int r = 0; r < nc; r++) {
        size_t id = (size_t)k * nc + r;
        int a = 0;
        double vp = 0;
        if (e.ctype[r] != SET_NONE) {
          vp = vplus_row(e.ctype[r], e.h[r], vs_prev[id], mu, e.lo[r], e.hi[r], a);
          prim_infeas = std::max(prim_infeas, std::fabs(prim_row(e.ctype[r], e.h[r], vs_prev[id], mu, e.lo[r], e.hi[r])));
        }
        vplus[id] = vp; act[id] = a;
        dbar[id] = mu * (vp - vs[id]);
        inner_crit = std::max(inner_crit, std::fabs(dbar[id]));
        double v = vs[id];
        for (int j = 0; j < n; j++) { double c = e.Cx[r * n + j]; Cact[id * n + j] = a ? c : 0.0; if (v != 0.0) g[j] += c * v; }
        if (k < T) for (int j = 0; j < m; j++) { double c = e.Cu[r * m + j]; Dact[id * m + j] = a ? c : 0.0; if (v != 0.0) g[n + j] += c * v; }
      }
      if (k < T) {
        const double *l1 = &lams[(size_t)(k + 1) * n];
        for (int i = 0; i < n; i++) {
          size_t id = (size_t)(k + 1) * n + i;
          prim_infeas = std::max(prim_infeas, std::fabs(e.gap[i]));
          lplus[id] = lams_prev[id] + e.gap[i] / mud();
          fbar[(size_t)k * n + i] = mud() * (lplus[id] - lams[id]);
          inner_crit = std::max(inner_crit, std::fabs(fbar[(size_t)k * n + i]));
          double l = l1[i];
          if (l != 0.0) { for (int j = 0; j < n; j++) g[j] += e.A[i * n + j] * l; for (int j = 0; j < m; j++) g[n + j] += e.B[i * m + j] * l; }
        }
      }
      // E_{k-1}^T lam_k  (k >= 1: dynamics of knot k-1; k = 0: initial condition, E = +I)
      const double *lk = &lams[(size_t)k * n];
      if (k == 0) for (int i = 0; i < n; i++) g[i] += lk[i];
      else {
        const double *E6 = ev[k - 1].E6.data();
        for (int j = 0; j < n; j++) {
          if (j < 6 && n >= 6) { double t = 0; for (int i = 0; i < 6; i++) t += E6[6 * i + j] * lk[i]; g[j] += t; } else g[j] -= lk[j];
        }
      }
      for (int j = 0; j < nz; j++) {
        if (k == 0 && j < n) continue; // x0 is fixed (force_initial_condition): not a decision variable
        if (k == T && j >= n) continue;
        dual_infeas = std::max(dual_infeas, std::fabs(g[j]));
      }
    }
    inner_crit = std::max(inner_crit, dual_infeas);
    if (getenv("ORC_DIAG")) { // where the primal infeasibility sits: worst constraint row and worst dynamics gap
      double bc = 0, bg = 0; int kc = -1, rc = -1, kg = -1, ig = -1;
      for (int k = 0; k <= T; k++) {
        for (int r = 0; r < nc; r++) if (ev[k].ctype[r] != SET_NONE) { double v = std::fabs(prim_row(ev[k].ctype[r], ev[k].h[r], vs_prev[(size_t)k * nc + r], mu, ev[k].lo[r], ev[k].hi[r])); if (v > bc) { bc = v; kc = k; rc = r; } }
        if (k < T) for (int i = 0; i < n; i++) if (std::fabs(ev[k].gap[i]) > bg) { bg = std::fabs(ev[k].gap[i]); kg = k; ig = i; }
      }
      int nact = 0; for (size_t i = 0; i < act.size(); i++) nact += act[i];
      fprintf(stderr, "   diag: worst cstr %.3e at knot %d row %d | worst gap %.3e at knot %d coord %d | active rows %d\n", bc, kc, rc, bg, kg, ig, nact);
    }
  }

  void solve_lq() {
    const int n = d.n, m = d.m, nc = d.nc, nz = n + m;
    std::vector<LQKnot> kn(T);
    std::vector<std::vector<double>> Hreg(T);
    for (int k = 0; k < T; k++) {
      Hreg[k] = ev[k].H;
      for (int i = 0; i < nz; i++) Hreg[k][i * nz + i] += preg;
      kn[k] = {Hreg[k].data(), &gq[(size_t)k * nz], ev[k].A.data(), ev[k].B.data(),
               P.cfg.kind == MPC_KIND_CENT ? nullptr : ev[k].E6.data(), &fbar[(size_t)k * n],
               &Cact[(size_t)k * nc * n], &Dact[(size_t)k * nc * m], &dbar[(size_t)k * nc]};
    }
    std::vector<double> HT = ev[T].H;
    for (int i = 0; i < n; i++) HT[i * nz + i] += preg;
    int nct = (P.cfg.kind == MPC_KIND_CENT) ? 0 : 3;
    riccati_solve(n, m, nc, T, kn.data(), HT.data(), nz, &gq[(size_t)T * nz], &Cact[(size_t)T * nc * n], &dbar[(size_t)T * nc], nct, mud(), mu, sol);
    // rows of the terminal constraint beyond nct (and inactive rows everywhere) follow dv = dbar/mu
    for (int k = 0; k <= T; k++)
      for (int r = 0; r < nc; r++) {
        size_t id = (size_t)k * nc + r;
        if (!act[id]) sol.dvs[id] = dbar[id] / mu;
      }
    for (int k = 0; k < T; k++)
      for (int i = 0; i < m; i++)
        for (int j = 0; j < n; j++) Kfb[((size_t)k * m + i) * n + j] = sol.K[((size_t)k * (m + nc) + i) * (1 + n) + 1 + j];
  }

  double directional_derivative() const {
    const int n = d.n, m = d.m, nc = d.nc;
    double dphi = 0;
    for (int k = 0; k <= T; k++) {
      const KnotEval &e = ev[k];
      const double *dx = &sol.dxs[(size_t)k * n];
      const double *du = k < T ? &sol.dus[(size_t)k * m] : nullptr;
      for (int j = 0; j < n; j++) dphi += e.lx[j] * dx[j];
      if (k < T) for (int j = 0; j < m; j++) dphi += e.lu[j] * du[j];
      for (int r = 0; r < nc; r++) {
        size_t id = (size_t)k * nc + r;
        if (e.ctype[r] == SET_NONE) continue;
        if (act[id]) {
          double jd = 0;
          for (int j = 0; j < n; j++) jd += e.Cx[r * n + j] * dx[j];
          if (k < T) for (int j = 0; j < m; j++) jd += e.Cu[r * m + j] * du[j];
          dphi += ((1 + prm.dual_weight) * vplus[id] - prm.dual_weight * vs[id]) * jd;
        }
        dphi -= prm.dual_weight * mu * (vplus[id] - vs[id]) * sol.dvs[id];
      }
      if (k < T) {
        const double *dxn = &sol.dxs[(size_t)(k + 1) * n];
        for (int i = 0; i < n; i++) {
          size_t id = (size_t)(k + 1) * n + i;
          double jd = 0;
          for (int j = 0; j < n; j++) jd += e.A[i * n + j] * dx[j];
          for (int j = 0; j < m; j++) jd += e.B[i * m + j] * du[j];
          if (i < 6 && n >= 6) { for (int j = 0; j < 6; j++) jd += e.E6[6 * i + j] * dxn[j]; } else jd -= dxn[i];
          dphi += ((1 + prm.dual_weight) * lplus[id] - prm.dual_weight * lams[id]) * jd - prm.dual_weight * mud() * (lplus[id] - lams[id]) * sol.dlams[id];
        }
      }
    }
    return dphi;
  }

  void state_difference(const double *x0, const double *x1, double *out) const { // x1 (-) x0
    if (P.cfg.kind == MPC_KIND_CENT) { for (int i = 0; i < d.n; i++) out[i] = x1[i] - x0[i]; return; }
    mb_difference<double>(x0, x1, out);
  }

  // ROLLOUT_NONLINEAR (aligator RolloutType::NONLINEAR; BASELINE north_star "nonlinear forward rollout and line search"; no
  // reference script selects it, so the semantics below are the oracle's — PARITY UNPINNED like the rest of the solver).
  // The trial point is rolled out through the NONLINEAR dynamics under the affine policy of the LQ solve:
  //   dx_k = x+_k (-) x_k,  du = alpha ku + Ku dx_k,  dv = alpha kv + Kv dx_k,  dlam' = alpha pt + W [dx_k; du],
  //   x+_{k+1} = f(x+_k, u+_k) (+) mu_d (lam_e - lam'+),
  // i.e. the new shooting gap satisfies the dual-regularised dynamics gap = mu_d (lam'+ - lam_e) exactly (the linear rollout
  // satisfies its linearisation); alpha = 1 on an LQ problem reproduces the linear rollout.
  double try_step_nonlinear(const Instance &in, double alpha, double *cost_out) {
    const int n = d.n, m = d.m, nc = d.nc, nz = n + m, nr = 1 + n, s = m + nc;
    std::vector<double> dx(n), z(nz), slack(n);
    std::copy(xs.begin(), xs.begin() + d.nx, txs.begin());
    for (int i = 0; i < n; i++) tlams[i] = lams[i] + alpha * sol.dlams[i];
    for (int k = 0; k < T; k++) {
      const double *xt = &txs[(size_t)k * d.nx];
      state_difference(&xs[(size_t)k * d.nx], xt, dx.data());
      const double *Kk = &sol.K[(size_t)k * s * nr];
      for (int i = 0; i < s; i++) {
        double t = alpha * Kk[i * nr];
        for (int j = 0; j < n; j++) t += Kk[i * nr + 1 + j] * dx[j];
        if (i < m) { z[n + i] = t; tus[(size_t)k * m + i] = us[(size_t)k * m + i] + t; }
        else tvs[(size_t)k * nc + i - m] = vs[(size_t)k * nc + i - m] + t;
      }
      for (int j = 0; j < n; j++) z[j] = dx[j];
      const double *Wk = &sol.W[(size_t)k * n * nz];
      for (int i = 0; i < n; i++) {
        double t = alpha * sol.pt[(size_t)k * n + i];
        for (int j = 0; j < nz; j++) t += Wk[i * nz + j] * z[j];
        size_t id = (size_t)(k + 1) * n + i;
        tlams[id] = lams[id] + t;
        slack[i] = mud() * (lams_prev[id] - tlams[id]);
      }
      eval_knot(P, in.knots[k], xt, &tus[(size_t)k * m], xt, false, tr[k]); // (x_{k+1} argument unused: the gap is set below)
      integrate_state(tr[k].xnext.data(), slack.data(), 1.0, &txs[(size_t)(k + 1) * d.nx]);
      for (int i = 0; i < n; i++) tr[k].gap[i] = -slack[i];
    }
    state_difference(&xs[(size_t)T * d.nx], &txs[(size_t)T * d.nx], dx.data());
    for (int r = 0; r < nc; r++) {
      size_t id = (size_t)T * nc + r;
      double t = alpha * dbar[id];
      for (int j = 0; j < n; j++) t += Cact[id * n + j] * dx[j];
      tvs[id] = vs[id] + t / mu;
    }
    eval_term(P, in.term, &txs[(size_t)T * d.nx], tr[T]);
    return merit_value(tr, tvs.data(), tlams.data(), cost_out);
  }

  double try_step(const Instance &in, double alpha, double *cost_out) {
    if (prm.rollout == 1) return try_step_nonlinear(in, alpha, cost_out);
    const int n = d.n, m = d.m;
    for (int k = 0; k <= T; k++) integrate_state(&xs[(size_t)k * d.nx], &sol.dxs[(size_t)k * n], alpha, &txs[(size_t)k * d.nx]);
    for (size_t i = 0; i < us.size(); i++) tus[i] = us[i] + alpha * sol.dus[i];
    for (size_t i = 0; i < vs.size(); i++) tvs[i] = vs[i] + alpha * sol.dvs[i];
    for (size_t i = 0; i < lams.size(); i++) tlams[i] = lams[i] + alpha * sol.dlams[i];
    (void)m;
    evaluate(in, txs.data(), tus.data(), false, tr);
    return merit_value(tr, tvs.data(), tlams.data(), cost_out);
  }

  // Armijo backtracking with quadratic/cubic interpolation (proxsuite-nlp ArmijoLinesearch)
  double linesearch(const Instance &in, double phi0, double dphi0, double &phi_out, double &cost_out) {
    double alpha = 1.0, a_prev = 0, phi_prev = 0;
    double phi_ref = phi0; // Armijo reference value; non-monotone: the largest of the last ls_window accepted merits
    if (prm.ls_mode == 1) for (double v : merit_hist) phi_ref = std::max(phi_ref, v);
    for (int it = 0;; it++) {
      double c;
      double phi = try_step(in, alpha, &c);
      ls_evals++;
      phi_out = phi; cost_out = c;
      if (getenv("ORC_DIAG")) fprintf(stderr, "   ls: alpha %.4e phi %.10e (phi0 %.10e, armijo bound %.10e, cost %.6e)\n", alpha, phi, phi0, phi_ref + prm.ls_c1 * alpha * dphi0, c);
      if (prm.ls_mode == 2) return alpha;
      if (phi <= phi_ref + prm.ls_c1 * alpha * dphi0) return alpha;
      if (std::fabs(dphi0) <= prm.ls_dphi_rel * std::max(1.0, std::fabs(phi0)) && std::isfinite(phi)) return alpha;
      if (alpha <= prm.ls_alpha_min || it + 1 >= prm.ls_max_steps) return alpha;
      double a_new;
      if (it == 0) a_new = -dphi0 * alpha * alpha / (2.0 * (phi - phi0 - dphi0 * alpha));
      else {
        double r1 = phi - phi0 - dphi0 * alpha, r2 = phi_prev - phi0 - dphi0 * a_prev;
        double den = alpha * alpha * a_prev * a_prev * (alpha - a_prev);
        double a = (a_prev * a_prev * r1 - alpha * alpha * r2) / den;
        double b = (-a_prev * a_prev * a_prev * r1 + alpha * alpha * alpha * r2) / den;
        if (std::fabs(a) < 1e-300) a_new = -dphi0 / (2.0 * b);
        else { double disc = b * b - 3.0 * a * dphi0; a_new = (-b + std::sqrt(disc)) / (3.0 * a); }
      }
      if (!(a_new >= prm.ls_contr_min * alpha)) a_new = prm.ls_contr_min * alpha; // also catches NaN
      if (a_new > prm.ls_contr_max * alpha) a_new = prm.ls_contr_max * alpha;
      a_prev = alpha; phi_prev = phi;
      alpha = std::max(a_new, prm.ls_alpha_min);
    }
  }

  void tols_on_failure() { prim_tol = std::pow(mu, prm.prim_alpha); inner_tol = std::pow(mu, prm.dual_alpha); }
  void tols_on_success() { prim_tol *= std::pow(mu, prm.prim_beta); inner_tol *= std::pow(mu, prm.dual_beta); }

  // solver.run(problem, xs_init, us_init)
  int run(const Instance &in, const double *xs_init, const double *us_init, int max_iters) {
    std::copy(xs_init, xs_init + xs.size(), xs.begin());
    std::copy(us_init, us_init + us.size(), us.begin());
    if (P.cfg.force_initial_condition) std::copy(in.x0.begin(), in.x0.end(), xs.begin());
    vs_prev = vs; lams_prev = lams;
    mu = prm.mu_init; preg = prm.reg_init;
    tols_on_failure();
    inner_tol = std::max(inner_tol, prm.tol); prim_tol = std::max(prim_tol, prm.tol);
    num_iters = 0; al_iters = 0; conv = 0; status = 1; ls_evals = 0; alpha_last = 0;
    alphas.clear();
    while (num_iters < max_iters && al_iters < prm.max_al_iters) {
      auto t0_ = std::chrono::steady_clock::now();
      evaluate(in, xs.data(), us.data(), true, ev);
      auto t1_ = std::chrono::steady_clock::now();
      assemble();
      auto t2_ = std::chrono::steady_clock::now();
      merit = merit_value(ev, vs.data(), lams.data(), &traj_cost);
      if (!std::isfinite(merit)) { status = 2; break; }
      if (inner_crit <= inner_tol) { // inner problem solved: BCL outer update
        if (prim_infeas <= prim_tol) {
          tols_on_success();
          vs_prev = vs; lams_prev = lams;
          if (std::max(prim_infeas, dual_infeas) <= prm.tol) { conv = 1; status = 0; break; }
        } else {
          mu = std::max(mu * prm.mu_update_factor, prm.mu_lower_bound);
          tols_on_failure();
        }
        inner_tol = std::max(inner_tol, 0.01 * prm.tol); prim_tol = std::max(prim_tol, prm.tol);
        al_iters++;
        continue;
      }
      solve_lq();
      auto t3_ = std::chrono::steady_clock::now();
      double dphi0 = directional_derivative();
      if (getenv("ORC_CHECK_DPHI")) { double c_; for (double eps : {1e-4, 1e-6, 1e-8}) { double pe = try_step(in, eps, &c_); fprintf(stderr, "   dphi0 %.6e  fd(eps=%.0e) %.6e\n", dphi0, eps, (pe - merit) / eps); } }
      double phi_new, cost_new;
      double alpha = linesearch(in, merit, dphi0, phi_new, cost_new);
      alphas.push_back(alpha); alpha_last = alpha;
      if (prm.ls_mode == 1) { merit_hist.push_back(phi_new); if ((int)merit_hist.size() > prm.ls_window) merit_hist.erase(merit_hist.begin()); }
      if (getenv("ORC_TIMING")) { auto t4_ = std::chrono::steady_clock::now(); auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
        fprintf(stderr, "eval %.1f ms  assemble %.1f ms  lq %.1f ms  dphi+linesearch %.1f ms\n", ms(t0_, t1_), ms(t1_, t2_), ms(t2_, t3_), ms(t3_, t4_)); }
      if (getenv("ORC_VERBOSE")) fprintf(stderr, "it %3d prim %.3e dual %.3e inner %.3e merit %.10e dphi0 %.3e alpha %.3e ls %d preg %.1e mu %.1e al %d\n", num_iters, prim_infeas, dual_infeas, inner_crit, merit, dphi0, alpha, ls_evals, preg, mu, al_iters);
      if (!std::isfinite(phi_new)) { status = 2; break; }
      xs.swap(txs); us.swap(tus); vs.swap(tvs); lams.swap(tlams);
      ev.swap(tr); // values at the accepted point (xdot, contact forces: workspace read-back, full:467-480)
      merit = phi_new; traj_cost = cost_new;
      if (alpha <= prm.ls_alpha_min) { if (preg >= prm.reg_max) { status = 3; break; } preg = std::min(preg * prm.reg_inc, prm.reg_max); }
      else preg = std::max(preg * prm.reg_dec, prm.reg_min);
      num_iters++;
    }
    return conv;
  }
};

} // namespace orc
