// ORACLE (test infrastructure, not product code) — rigid-body algorithms for the Talos-shaped tree.
//
// Restates, from the published algorithms (Featherstone RBDA; Carpentier & Mansard 2018), what the
// reference obtains from Pinocchio through Aligator (App. A3-A4):
//   forwardKinematics / frame placements        fulldynamic_talos.py:50-51
//   constraintDynamics (CRBA + RNEA + contact KKT, ProximalSettings(1e-9,1e-10,1))   fulldynamic_talos.py:77-109
//   computeConstraintDynamicsDerivatives         (inside MultibodyConstraintFwdDynamics.dForward)
//   centroidal momentum + derivatives            fulldynamic_talos.py:160-162
// Pinocchio is not vendored in /root/reference => PARITY UNPINNED against upstream; value functions are
// templated so the hand-derived analytic Jacobians below are verified against forward-mode AD in tests.
//
// World-frame formulation: every spatial quantity is expressed at the world origin in world axes, so
// subtree accumulations are plain sums.
#pragma once
#include "../include/mpcb200.h"
#include "spatial.hpp"
#include <vector>

namespace orc {

constexpr int NB = MPC_NB, NV = MPC_NV, NQ = MPC_NQ, NJ = MPC_NJ;

inline int body_of_dof(int j) { return j < 6 ? 0 : j - 5; }

struct Tree {
  const mpc_robot_t *rb;
  bool anc[NB][NB]; // anc[a][b]: a is ancestor-or-self of b
  explicit Tree(const mpc_robot_t *r) : rb(r) {
    for (int a = 0; a < NB; a++)
      for (int b = 0; b < NB; b++) {
        bool f = false;
        for (int k = b; k >= 0; k = rb->parent[k]) if (k == a) { f = true; break; }
        anc[a][b] = f;
      }
  }
};

template <class T> struct Kin {
  SE3<T> oM[NB];
  V6<T> S[NV];   // world-frame motion subspace columns
  V6<T> v[NB];   // world-frame body spatial velocities
  M6<T> I[NB];   // world-frame body spatial inertias
  M6<T> Ic[NB];  // composite (subtree) inertias
  T mass;
  V3<T> com;
};

template <class T> M3<T> rot_axis(const double *ax, T q) {
  V3<T> w = {T(ax[0]) * q, T(ax[1]) * q, T(ax[2]) * q};
  return exp3(w);
}

// q: [p(3), quat xyzw (4), theta(22)], v: nv
template <class T> void forward_kin(const Tree &tr, const T *q, const T *v, Kin<T> &k) {
  const mpc_robot_t &rb = *tr.rb;
  k.oM[0].R = quat_to_R(q + 3);
  k.oM[0].p = {q[0], q[1], q[2]};
  for (int j = 0; j < 6; j++) { V6<T> e = zero6<T>(); e[j] = T(1); k.S[j] = act_motion(k.oM[0], e); }
  for (int b = 1; b < NB; b++) {
    SE3<T> pl = se3_cast<T>(rb.jplace[b]);
    SE3<T> jr; jr.R = rot_axis<T>(rb.axis[b], q[6 + b]); jr.p = {T(0), T(0), T(0)};
    k.oM[b] = mul(k.oM[rb.parent[b]], mul(pl, jr));
    V6<T> e = {T(0), T(0), T(0), T(rb.axis[b][0]), T(rb.axis[b][1]), T(rb.axis[b][2])};
    k.S[5 + b] = act_motion(k.oM[b], e);
  }
  // velocities
  k.v[0] = zero6<T>();
  for (int j = 0; j < 6; j++) k.v[0] = add(k.v[0], scale(k.S[j], v[j]));
  for (int b = 1; b < NB; b++) k.v[b] = add(k.v[rb.parent[b]], scale(k.S[5 + b], v[5 + b]));
  // inertias in the world frame
  k.mass = T(0);
  V3<T> mc = {T(0), T(0), T(0)};
  for (int b = 0; b < NB; b++) {
    V3<T> c = {T(rb.com[b][0]), T(rb.com[b][1]), T(rb.com[b][2])};
    V3<T> cw = add(mul(k.oM[b].R, c), k.oM[b].p);
    M3<T> Ib; for (int i = 0; i < 9; i++) Ib[i] = T(rb.inertia[b][i]);
    M3<T> Iw = mul(mul(k.oM[b].R, Ib), transpose(k.oM[b].R));
    k.I[b] = inertia6(T(rb.mass[b]), cw, Iw);
    k.mass += T(rb.mass[b]);
    mc = add(mc, scale(cw, T(rb.mass[b])));
  }
  k.com = scale(mc, T(1) / k.mass);
  for (int b = NB - 1; b >= 0; b--) {
    k.Ic[b] = k.I[b];
    for (int c = b + 1; c < NB; c++)
      if (rb.parent[c] == b) for (int i = 0; i < 36; i++) k.Ic[b][i] += k.Ic[c][i];
  }
}

// World-frame body accelerations for joint accelerations qdd (nullptr = 0). with_gravity uses a_world=-g.
template <class T> void body_accels(const Tree &tr, const Kin<T> &k, const T *v, const T *qdd, bool with_gravity, V6<T> *a) {
  const mpc_robot_t &rb = *tr.rb;
  V6<T> a0 = zero6<T>();
  if (with_gravity) { a0[0] = T(-rb.gravity[0]); a0[1] = T(-rb.gravity[1]); a0[2] = T(-rb.gravity[2]); }
  a[0] = a0; // joint velocity of the base is v[0] itself: v x (S qd) = v x v = 0
  if (qdd) for (int j = 0; j < 6; j++) a[0] = add(a[0], scale(k.S[j], qdd[j]));
  for (int b = 1; b < NB; b++) {
    V6<T> sj = scale(k.S[5 + b], v[5 + b]);
    a[b] = add(a[rb.parent[b]], cross_mm(k.v[b], sj));
    if (qdd) a[b] = add(a[b], scale(k.S[5 + b], qdd[5 + b]));
  }
}

// tau = ID(q,v,qdd) - J^T fext, fext[b] = world-frame external wrench on body b (may be null)
template <class T> void rnea(const Tree &tr, const Kin<T> &k, const T *v, const T *qdd, const V6<T> *fext, T *tau,
                            V6<T> *a_out = nullptr, V6<T> *F_out = nullptr) {
  const mpc_robot_t &rb = *tr.rb;
  V6<T> a[NB], F[NB];
  body_accels(tr, k, v, qdd, true, a);
  for (int b = 0; b < NB; b++) {
    F[b] = add(mul(k.I[b], a[b]), cross_mf(k.v[b], mul(k.I[b], k.v[b])));
    if (fext) F[b] = sub(F[b], fext[b]);
  }
  for (int b = NB - 1; b > 0; b--) F[rb.parent[b]] = add(F[rb.parent[b]], F[b]);
  for (int j = 0; j < NV; j++) tau[j] = dot(k.S[j], F[body_of_dof(j)]);
  if (a_out) for (int b = 0; b < NB; b++) a_out[b] = a[b];
  if (F_out) for (int b = 0; b < NB; b++) F_out[b] = F[b];
}

template <class T> void crba(const Tree &tr, const Kin<T> &k, T *M /*NV x NV*/) {
  for (int i = 0; i < NV; i++)
    for (int j = 0; j < NV; j++) {
      int bi = body_of_dof(i), bj = body_of_dof(j);
      T m = T(0);
      if (tr.anc[bi][bj]) m = dot(k.S[i], mul(k.Ic[bj], k.S[j]));
      else if (tr.anc[bj][bi]) m = dot(k.S[i], mul(k.Ic[bi], k.S[j]));
      M[i * NV + j] = m;
    }
}

// dense Cholesky (lower, in place) and solves, templated
template <class T> void chol(T *A, int n) {
  for (int j = 0; j < n; j++) {
    T d = A[j * n + j];
    for (int k = 0; k < j; k++) d -= A[j * n + k] * A[j * n + k];
    d = sqrt(d);
    A[j * n + j] = d;
    for (int i = j + 1; i < n; i++) {
      T s = A[i * n + j];
      for (int k = 0; k < j; k++) s -= A[i * n + k] * A[j * n + k];
      A[i * n + j] = s / d;
    }
  }
}
template <class T> void chol_solve(const T *L, int n, T *b, int nrhs, int ldb) { // b: n x nrhs row-major
  for (int c = 0; c < nrhs; c++) {
    for (int i = 0; i < n; i++) {
      T s = b[i * ldb + c];
      for (int k = 0; k < i; k++) s -= L[i * n + k] * b[k * ldb + c];
      b[i * ldb + c] = s / L[i * n + i];
    }
    for (int i = n - 1; i >= 0; i--) {
      T s = b[i * ldb + c];
      for (int k = i + 1; k < n; k++) s -= L[k * n + i] * b[k * ldb + c];
      b[i * ldb + c] = s / L[i * n + i];
    }
  }
}

// ------------------------------------------------------------------ contact kinematics
template <class T> struct Contact {
  int body;
  SE3<T> oMc;       // world placement of the contact frame c1
  T J[6 * NV];      // LOCAL frame Jacobian
  V6<T> vc;         // LOCAL spatial velocity
  V6<T> gamma;      // LOCAL spatial acceleration drift (qdd = 0, no gravity)
  V6<T> astar;      // Baumgarte desired acceleration
  SE3<T> c1Mc2;
};

template <class T> void contact_kin(const Tree &tr, const mpc_config_t &cfg, const Kin<T> &k, const T *v, int foot, Contact<T> &c) {
  const mpc_robot_t &rb = *tr.rb;
  c.body = rb.foot_body[foot];
  c.oMc = mul(k.oM[c.body], se3_cast<T>(rb.foot_place[foot]));
  for (int j = 0; j < NV; j++) {
    V6<T> col = tr.anc[body_of_dof(j)][c.body] ? actinv_motion(c.oMc, k.S[j]) : zero6<T>();
    for (int r = 0; r < 6; r++) c.J[r * NV + j] = col[r];
  }
  c.vc = actinv_motion(c.oMc, k.v[c.body]);
  V6<T> a[NB];
  body_accels<T>(tr, k, v, nullptr, false, a);
  c.gamma = actinv_motion(c.oMc, a[c.body]);
  c.c1Mc2 = mul(inverse(c.oMc), se3_cast<T>(cfg.contact_place[foot]));
  V6<T> lg = log6(c.c1Mc2);
  for (int r = 0; r < 6; r++) c.astar[r] = T(cfg.kp[r]) * lg[r] - T(cfg.kd[r]) * c.vc[r];
}

// Constrained forward dynamics (App. A3): [M J^T; J -mu I][a; -lam] = [tau - b; astar - gamma].
// active[2]: contacts present. Outputs a (NV), lam (12: left 0..5, right 6..11; zeros if inactive).
template <class T> struct CDyn {
  Kin<T> kin;
  Contact<T> con[2];
  int nact, act[2];
  T M[NV * NV], L[NV * NV];  // mass matrix and its Cholesky factor
  T b[NV];                   // nonlinear effects
  T Jm[12 * NV];             // stacked active Jacobians (nk x NV)
  T Y[NV * 12];              // M^-1 J^T (NV x nk)
  T G[144], LG[144];         // J M^-1 J^T + mu I and its factor (nk x nk)
  T a[NV], lam[12];
};

template <class T> void constrained_dynamics(const Tree &tr, const mpc_config_t &cfg, const T *q, const T *v, const T *tau,
                                             const bool active[2], CDyn<T> &d) {
  forward_kin(tr, q, v, d.kin);
  crba(tr, d.kin, d.M);
  rnea<T>(tr, d.kin, v, nullptr, nullptr, d.b);
  d.nact = 0;
  for (int f = 0; f < 2; f++) if (active[f]) { contact_kin(tr, cfg, d.kin, v, f, d.con[f]); d.act[d.nact++] = f; }
  int nk = 6 * d.nact;
  for (int i = 0; i < NV * NV; i++) d.L[i] = d.M[i];
  chol(d.L, NV);
  T afree[NV];
  for (int i = 0; i < NV; i++) afree[i] = tau[i] - d.b[i];
  chol_solve(d.L, NV, afree, 1, 1);
  for (int c = 0; c < d.nact; c++)
    for (int r = 0; r < 6; r++)
      for (int j = 0; j < NV; j++) { d.Jm[(6 * c + r) * NV + j] = d.con[d.act[c]].J[r * NV + j]; d.Y[j * 12 + 6 * c + r] = d.con[d.act[c]].J[r * NV + j]; }
  if (nk) chol_solve(d.L, NV, d.Y, nk, 12);
  T rhs[12];
  for (int r = 0; r < nk; r++) {
    const Contact<T> &c = d.con[d.act[r / 6]];
    T s = c.astar[r % 6] - c.gamma[r % 6];
    for (int j = 0; j < NV; j++) s -= d.Jm[r * NV + j] * afree[j];
    rhs[r] = s;
    for (int c2 = 0; c2 < nk; c2++) {
      T g = (r == c2) ? T(cfg.mu_contact) : T(0);
      for (int j = 0; j < NV; j++) g += d.Jm[r * NV + j] * d.Y[j * 12 + c2];
      d.G[r * nk + c2] = g;
    }
  }
  for (int i = 0; i < nk * nk; i++) d.LG[i] = d.G[i];
  if (nk) { chol(d.LG, nk); chol_solve(d.LG, nk, rhs, 1, 1); }
  for (int i = 0; i < 12; i++) d.lam[i] = T(0);
  for (int j = 0; j < NV; j++) {
    T s = afree[j];
    for (int r = 0; r < nk; r++) s += d.Y[j * 12 + r] * rhs[r];
    d.a[j] = s;
  }
  for (int c = 0; c < d.nact; c++) for (int r = 0; r < 6; r++) d.lam[6 * d.act[c] + r] = rhs[6 * c + r];
}

// ------------------------------------------------------------------ analytic derivatives (double)
// Tangent of inverse dynamics in the world frame (derivation in DESIGN.md "RBD derivatives").
// For dof j on body J with world twist s = S_j, parent body pJ:
//   w_j = s x v_pJ,  c_j = s x a_pJ - w_j x v_pJ,  e_J = v_J + v_pJ
//   g_k^q = I_k (c_j + w_j x v_k) + w_j x* (I_k v_k) + v_k x* (I_k w_j)        k in subtree(J)
//   g_k^v = I_k (s x (v_k - e_J)) + s x* (I_k v_k) + v_k x* (I_k s)
//   dtau_i/dq_j = -S_i^T sum_{k>=i} g_k^q  (i in subtree(J));  S_i^T (s x* Fnet_J - sum_{k>=J} g_k^q) (i ancestor dof)
//   dtau_i/dv_j =  S_i^T sum_{k>=max(i,J)} g_k^v
// Fnet = subtree inertial force minus external forces (held constant in the LOCAL frames).
struct IDDerivs {
  double dq[NV * NV], dv[NV * NV];
};

inline void id_derivatives(const Tree &tr, const Kin<double> &k, const V6<double> *a /*with gravity*/, const V6<double> *Fnet,
                           IDDerivs &out) {
  const mpc_robot_t &rb = *tr.rb;
  V6<double> a_world = zero6<double>();
  a_world[0] = -rb.gravity[0]; a_world[1] = -rb.gravity[1]; a_world[2] = -rb.gravity[2];
  for (int j = 0; j < NV; j++) {
    int J = body_of_dof(j), pJ = rb.parent[J];
    V6<double> s = k.S[j];
    V6<double> vp = pJ >= 0 ? k.v[pJ] : zero6<double>();
    V6<double> ap = pJ >= 0 ? a[pJ] : a_world;
    V6<double> w = cross_mm(s, vp);
    V6<double> cj = sub(cross_mm(s, ap), cross_mm(w, vp));
    V6<double> eJ = add(k.v[J], vp);
    V6<double> Gq[NB], Gv[NB];
    for (int b = 0; b < NB; b++) { Gq[b] = zero6<double>(); Gv[b] = zero6<double>(); }
    for (int b = NB - 1; b >= 0; b--) {
      if (!tr.anc[J][b]) continue;
      V6<double> Iv = mul(k.I[b], k.v[b]);
      V6<double> gq = add(add(mul(k.I[b], add(cj, cross_mm(w, k.v[b]))), cross_mf(w, Iv)), cross_mf(k.v[b], mul(k.I[b], w)));
      V6<double> gv = add(add(mul(k.I[b], cross_mm(s, sub(k.v[b], eJ))), cross_mf(s, Iv)), cross_mf(k.v[b], mul(k.I[b], s)));
      Gq[b] = add(Gq[b], gq); Gv[b] = add(Gv[b], gv);
      if (b != J) { Gq[rb.parent[b]] = add(Gq[rb.parent[b]], Gq[b]); Gv[rb.parent[b]] = add(Gv[rb.parent[b]], Gv[b]); }
    }
    V6<double> top = sub(cross_mf(s, Fnet[J]), Gq[J]);
    for (int i = 0; i < NV; i++) {
      int bi = body_of_dof(i);
      double dq = 0, dv = 0;
      if (tr.anc[J][bi]) { dq = -dot(k.S[i], Gq[bi]); dv = dot(k.S[i], Gv[bi]); }
      else if (tr.anc[bi][J]) { dq = dot(k.S[i], top); dv = dot(k.S[i], Gv[J]); }
      out.dq[i * NV + j] = dq; out.dv[i * NV + j] = dv;
    }
  }
}

// Derivatives of the constrained dynamics (App. A4). Outputs: da_dq, da_dv (NV x NV), da_dtau (NV x NV),
// dlam_dq, dlam_dv, dlam_dtau (12 x NV; rows of inactive contacts are zero).
struct CDynDerivs {
  double da_dq[NV * NV], da_dv[NV * NV], da_dtau[NV * NV];
  double dl_dq[12 * NV], dl_dv[12 * NV], dl_dtau[12 * NV];
};

inline void constrained_dynamics_derivatives(const Tree &tr, const mpc_config_t &cfg, const double *v, const CDyn<double> &d,
                                             CDynDerivs &o) {
  const mpc_robot_t &rb = *tr.rb;
  const Kin<double> &k = d.kin;
  int nk = 6 * d.nact;
  // inverse dynamics at (q, v, a) with the contact wrenches as external forces
  V6<double> fext[NB], a[NB], Fnet[NB];
  for (int b = 0; b < NB; b++) fext[b] = zero6<double>();
  for (int c = 0; c < d.nact; c++) {
    const Contact<double> &cc = d.con[d.act[c]];
    V6<double> l; for (int r = 0; r < 6; r++) l[r] = d.lam[6 * d.act[c] + r];
    fext[cc.body] = add(fext[cc.body], act_force(cc.oMc, l));
  }
  double tau_chk[NV];
  rnea<double>(tr, k, v, d.a, fext, tau_chk, a, Fnet);
  IDDerivs idd;
  id_derivatives(tr, k, a, Fnet, idd);
  // contact-acceleration derivatives R2 = d(alpha - astar)/d(q,v) at fixed qdd = a
  std::vector<double> R2q(12 * NV, 0.0), R2v(12 * NV, 0.0);
  V6<double> ang_[NB];
  body_accels<double>(tr, k, v, d.a, false, ang_); // spatial accelerations without gravity
  for (int c = 0; c < d.nact; c++) {
    const Contact<double> &cc = d.con[d.act[c]];
    M6<double> Jl = Jlog6(cc.c1Mc2);
    M6<double> Adi = action_matrix(inverse(cc.c1Mc2));
    for (int j = 0; j < NV; j++) {
      int J = body_of_dof(j);
      if (!tr.anc[J][cc.body]) continue;
      int pJ = rb.parent[J];
      V6<double> s = k.S[j];
      V6<double> vp = pJ >= 0 ? k.v[pJ] : zero6<double>();
      V6<double> ap = pJ >= 0 ? ang_[pJ] : zero6<double>();
      V6<double> w = cross_mm(s, vp);
      V6<double> cj = sub(cross_mm(s, ap), cross_mm(w, vp));
      V6<double> dalpha_q = scale(actinv_motion(cc.oMc, add(cj, cross_mm(w, k.v[cc.body]))), double(-1.0));
      V6<double> eJ = add(k.v[J], vp);
      V6<double> dalpha_v = actinv_motion(cc.oMc, cross_mm(s, sub(k.v[cc.body], eJ)));
      V6<double> Jc; for (int r = 0; r < 6; r++) Jc[r] = cc.J[r * NV + j];
      V6<double> dlog = scale(mul(Jl, mul(Adi, Jc)), double(-1.0));
      V6<double> wl = actinv_motion(cc.oMc, w); // d vc/dq = -wl
      for (int r = 0; r < 6; r++) {
        double dastar_q = cfg.kp[r] * dlog[r] + cfg.kd[r] * wl[r];
        double dastar_v = -cfg.kd[r] * Jc[r];
        R2q[(6 * c + r) * NV + j] = dalpha_q[r] - dastar_q;
        R2v[(6 * c + r) * NV + j] = dalpha_v[r] - dastar_v;
      }
    }
  }
  // solve:  dlam = G^-1 (J M^-1 R1 - R2),  da = Y dlam - M^-1 R1, for R1 in {dID/dq, dID/dv, -I}
  auto solve = [&](const double *R1, const double *R2, double *da, double *dl) {
    std::vector<double> X(R1, R1 + NV * NV);
    chol_solve(d.L, NV, X.data(), NV, NV); // M^-1 R1
    std::vector<double> rhs(12 * NV, 0.0);
    for (int r = 0; r < nk; r++)
      for (int j = 0; j < NV; j++) {
        double s = R2 ? -R2[r * NV + j] : double(0.0);
        for (int i = 0; i < NV; i++) s += d.Jm[r * NV + i] * X[i * NV + j];
        rhs[r * NV + j] = s;
      }
    if (nk) chol_solve(d.LG, nk, rhs.data(), NV, NV);
    for (int i = 0; i < NV; i++)
      for (int j = 0; j < NV; j++) {
        double s = -X[i * NV + j];
        for (int r = 0; r < nk; r++) s += d.Y[i * 12 + r] * rhs[r * NV + j];
        da[i * NV + j] = s;
      }
    for (int i = 0; i < 12 * NV; i++) dl[i] = 0;
    for (int c = 0; c < d.nact; c++)
      for (int r = 0; r < 6; r++)
        for (int j = 0; j < NV; j++) dl[(6 * d.act[c] + r) * NV + j] = rhs[(6 * c + r) * NV + j];
  };
  solve(idd.dq, R2q.data(), o.da_dq, o.dl_dq);
  solve(idd.dv, R2v.data(), o.da_dv, o.dl_dv);
  std::vector<double> mI(NV * NV, 0.0);
  for (int i = 0; i < NV; i++) mI[i * NV + i] = -1.0;
  solve(mI.data(), nullptr, o.da_dtau, o.dl_dtau);
}

// ------------------------------------------------------------------ centroidal momentum, CoM, frames
// h_g = [linear; angular about the CoM] in world axes (pinocchio computeCentroidalMomentum).
template <class T> V6<T> centroidal_momentum(const Kin<T> &k) {
  V6<T> h = zero6<T>();
  for (int b = 0; b < NB; b++) h = add(h, mul(k.I[b], k.v[b]));
  V3<T> l = lin(h);
  return mk6(l, sub(ang(h), cross(k.com, l)));
}
// Jcom (3 x NV): column j = linear part of Ic_J s_j over total mass, shifted to the subtree com.
inline void com_jacobian(const Kin<double> &k, double *Jc) {
  for (int j = 0; j < NV; j++) {
    V6<double> m = mul(k.Ic[body_of_dof(j)], k.S[j]);
    for (int r = 0; r < 3; r++) Jc[r * NV + j] = m[r] / k.mass;
  }
}
// dh_g/dq (6 x NV) and dh_g/dv = A_g (6 x NV)
inline void centroidal_derivatives(const Tree &tr, const Kin<double> &k, double *dh_dq, double *Ag) {
  const mpc_robot_t &rb = *tr.rb;
  V6<double> hsub[NB];
  for (int b = 0; b < NB; b++) hsub[b] = mul(k.I[b], k.v[b]);
  for (int b = NB - 1; b > 0; b--) hsub[rb.parent[b]] = add(hsub[rb.parent[b]], hsub[b]);
  V3<double> ptot = lin(hsub[0]);
  for (int j = 0; j < NV; j++) {
    int J = body_of_dof(j), pJ = rb.parent[J];
    V6<double> s = k.S[j];
    V6<double> vp = pJ >= 0 ? k.v[pJ] : zero6<double>();
    V6<double> w = cross_mm(s, vp);
    V6<double> dho = sub(cross_mf(s, hsub[J]), mul(k.Ic[J], w)); // d h_o / d q_j
    V6<double> Is = mul(k.Ic[J], s);                             // d h_o / d v_j
    V3<double> dc = scale(lin(Is), 1.0 / k.mass);                // d com / d q_j
    V3<double> dLq = sub(sub(ang(dho), cross(dc, ptot)), cross(k.com, lin(dho)));
    V3<double> dLv = sub(ang(Is), cross(k.com, lin(Is)));
    for (int r = 0; r < 3; r++) {
      dh_dq[r * NV + j] = dho[r]; dh_dq[(3 + r) * NV + j] = dLq[r];
      Ag[r * NV + j] = Is[r]; Ag[(3 + r) * NV + j] = dLv[r];
    }
  }
}

// ------------------------------------------------------------------ centroidal momentum RATE (kinodynamics, App. A5)
// F_g(q, v, a) = A_g a + dA_g v + (gravity folded in as a_world = -g): total inertial force of the tree expressed at the
// CoM, [linear; angular].  With a_world = -g this equals hdot_g - m g, so the kinodynamic base-acceleration equation reads
// F_g(q,v,a) = [sum f_i ; sum (p_i - c) x f_i + tau_i].
template <class T> V6<T> centroidal_rate(const Tree &tr, const Kin<T> &k, const T *v, const T *qdd) {
  T tau[NV];
  V6<T> F[NB];
  rnea<T>(tr, k, v, qdd, nullptr, tau, nullptr, F);
  V3<T> l = lin(F[0]);
  return mk6(l, sub(ang(F[0]), cross(k.com, l)));
}
// d F_o / d q_j and d F_o / d v_j of the TOTAL force about the world origin (6 x NV each): the "top" vectors of id_derivatives
inline void total_force_derivatives(const Tree &tr, const Kin<double> &k, const V6<double> *a /*with gravity*/, const V6<double> *Fsub,
                                    double *dFq, double *dFv) {
  const mpc_robot_t &rb = *tr.rb;
  V6<double> a_world = zero6<double>();
  a_world[0] = -rb.gravity[0]; a_world[1] = -rb.gravity[1]; a_world[2] = -rb.gravity[2];
  for (int j = 0; j < NV; j++) {
    int J = body_of_dof(j), pJ = rb.parent[J];
    V6<double> s = k.S[j];
    V6<double> vp = pJ >= 0 ? k.v[pJ] : zero6<double>();
    V6<double> ap = pJ >= 0 ? a[pJ] : a_world;
    V6<double> w = cross_mm(s, vp);
    V6<double> cj = sub(cross_mm(s, ap), cross_mm(w, vp));
    V6<double> eJ = add(k.v[J], vp);
    V6<double> Gq = zero6<double>(), Gv = zero6<double>();
    for (int b = 0; b < NB; b++) {
      if (!tr.anc[J][b]) continue;
      V6<double> Iv = mul(k.I[b], k.v[b]);
      Gq = add(Gq, add(add(mul(k.I[b], add(cj, cross_mm(w, k.v[b]))), cross_mf(w, Iv)), cross_mf(k.v[b], mul(k.I[b], w))));
      Gv = add(Gv, add(add(mul(k.I[b], cross_mm(s, sub(k.v[b], eJ))), cross_mf(s, Iv)), cross_mf(k.v[b], mul(k.I[b], s))));
    }
    V6<double> top = sub(cross_mf(s, Fsub[J]), Gq);
    for (int r = 0; r < 6; r++) { dFq[r * NV + j] = top[r]; dFv[r * NV + j] = Gv[r]; }
  }
}

} // namespace orc
