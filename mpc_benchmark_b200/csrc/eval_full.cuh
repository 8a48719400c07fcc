// Per-knot evaluation of the FULL-DYNAMICS Talos stage: one CTA per (instance, knot).
//
// Replaces, for one knot, what the reference obtains from five independent Aligator objects that each re-run
// pinocchio::constraintDynamics + computeConstraintDynamicsDerivatives (SURVEY finding 7):
//   MultibodyConstraintFwdDynamics + IntegratorSemiImplEuler      fulldynamic_talos.py:100-111
//   ContactForceResidual x2, MultibodyWrenchConeResidual x2        fulldynamic_talos.py:188-201, 211-225
//   QuadraticState/Control cost, CentroidalMomentumResidual, FramePlacementResidual x2   :160-185
//   torque / joint-limit boxes                                     :206-209
// Here (a, lambda, da, dlambda) are computed ONCE in shared memory and every cost/constraint reads them.
//
// World-frame formulation (all spatial quantities at the world origin, world axes): subtree accumulations are
// plain sums.  Derivative columns use composite quantities (DESIGN.md "RBD derivatives"):
//   dtau_i/dq_j = -S_i.(Ic_m c_j + Bc_m w_j),  m = body(i) in subtree(body(j));  ancestors: S_i.(s_j x* F_J - g_J)
//   dtau_i/dv_j = +S_i.(Ic_m c'_j + Bc_m s_j)
#pragma once
#include <cstddef>
#include "dmma.cuh"
#include "model.cuh"

namespace mpcdev {

// optional per-phase cycle counters of the evaluation kernel (-DMPC_PHASE_TIMING builds; knot 1 of instance 0)
#if defined(MPC_PHASE_TIMING) && !defined(MPC_HOST_EMU)
#define EPH_DECL long long eph_last = clock64(); long long eph_acc[16] = {0}
#define EPH(i) do { SYNC(); if (threadIdx.x == 0) { long long t_ = clock64(); eph_acc[i] += t_ - eph_last; eph_last = t_; } } while (0)
#define EPH_DUMP(ptr) do { if (threadIdx.x == 0 && (ptr)) for (int i_ = 0; i_ < 16; i_++) (ptr)[i_] = (double)eph_acc[i_]; } while (0)
#else
#define EPH_DECL
#define EPH(i)
#define EPH_DUMP(ptr)
#endif


constexpr int FN = 56, FM = 22, FNZ = 78, FNC = 78;

// merit / infeasibility partials written per knot
enum { SC_COST = 0, SC_PEN, SC_PRIM, SC_DUAL, SC_INNER, SC_COUNT = 8 };

struct KnotIO {
  // inputs
  const double *x, *u, *xn;        // x_k, u_k, x_{k+1}
  const mpc_knot_t *kn;
  const mpc_term_t *tm;
  const double *v, *v_prev;        // constraint multipliers of this knot (current / BCL estimate)
  const double *lam_k;             // co-state of x_k (dynamics k-1, or initial condition for k = 0)
  const double *lam_n, *lam_n_prev;// co-state of x_{k+1}
  double mu, preg;
  int k, T;
  // outputs (derivative pass)
  double *AB, *H, *lxu, *g, *T6, *E6, *gE_next, *fbar, *dbar, *vplus, *lplus, *CDact;
  int32_t *nca, *act_idx;
  // outputs (both passes)
  double *gap, *h, *scal, *xdot, *lamc;
  // ROLLOUT_NONLINEAR (fused rollout kernel, values pass only): x_{k+1} is not an input but DEFINED here as f(x, u) (+) slack and written out
  const double *slack; double *xn_out;
  double *scratch;   // derivative pass: this knot's slot of the Riccati W buffer (n x nz doubles, dead until the Riccati kernel writes it):
                     // holds the tangent X = da/dz (NV x nz) and dlam/dz (12 x nz), which stay L2-resident between the phases
  double *phase_out; // profiling builds only
};

template <bool WITH_DERIV> struct FullWsT {
  static constexpr int TOPS = WITH_DERIV ? 2 * NV * 6 : 8, BCS = WITH_DERIV ? NB * 36 : 8, HROWS = WITH_DERIV ? 40 : 1;
  static constexpr int CJ1 = WITH_DERIV ? 6 * FN : 8, CJ2 = WITH_DERIV ? 2 * 6 * NV : 8, M6 = WITH_DERIV ? 36 : 1, ZS = WITH_DERIV ? FNZ : 8;
  double x[NQ + NV], u[FM], xn[NQ + NV];
  double kn[sizeof(mpc_knot_t) / 8];
  double oM[NB * 12], S[NV * 6], v[NB * 6], a[NB * 6], I[NB * 10], Ic[NB * 10];
  double hb[NB * 6], hsub[NB * 6], f[NB * 6], Fsub[NB * 6];
  union { // Bc is dead once the derivative columns are built; the cost Jacobians are built afterwards
    double Bc[BCS];
    struct { double Jcent[CJ1], Jpose[CJ2]; } cj;
  };
  double top[TOPS]; // directly behind Bc: [Bc | top] doubles as the scratch of the explicit mass-matrix inverse (both are free until then)
  double U[NV * 6];
  union { // the mass-matrix factor is dead after X = M^-1 R1; the output-phase vectors live afterwards
    double M[NV * NV];
    struct {
      double P1[M6], P2[M6], E6[M6], T6[M6], Jlg[M6], Jrg[M6], AdD[M6], AdDi[M6], Jd[M6], Ade[M6];
      double lxu[ZS], g[ZS], hval[FNC], vpl[FNC], dbr[FNC], rowtmp[FNC];
    } late;
  };
  double acc[NV];
  double ofoot[24], Jf[2 * 6 * NV];
  double vc[12], gam[12], astar[12], c1Mc2[24], JlAd[2 * M6], lam[12], lgc[12];
  union { // the contact-solve scratch is dead once da/dlam are formed; the multipliers are staged into it afterwards
    struct { double Y[NV * 13], G[144], rhs[12], dinvM[4 * 64], dinvG[2 * 64]; };
    struct { double mv[FNC], mvp[FNC], mln[FN], mlnp[FN], mlk[FN]; }; // v, v_prev, lam_{k+1}, its estimate, lam_k
  };
  double rcent[6], rpose[12], Jlp[2 * M6];
  double estate[FN], Jls[M6];
  double dx[FN], xnext[NQ + NV], lgap[6], eexp[12], Dgap[12], Dl[36];
  double lpl[FN], fbr[FN];
  double com[3], scal[SC_COUNT], part[32];
  // Gauss-Newton Hessian as ONE tensor-core product H = sum_k w_k J_k' J_k: the weighted residual-Jacobian rows stay where they are
  // (dlam/dz rows in the global scratch, Jcent, Jpose, Jls in shared memory) and are addressed through this table
  double hrow_w[HROWS];
  int32_t hrow_len[HROWS];
  int32_t act[2], nact, sidx[2], ctype[FNC], isact[FNC], act_idx[FNC], nca;
};
using FullWs = FullWsT<true>;
static_assert(offsetof(FullWs, top) == offsetof(FullWs, Bc) + sizeof(double) * FullWs::BCS, "[Bc | top] must be contiguous");
static_assert(FullWs::BCS + FullWs::TOPS >= NV * NV + 3 * 64, "mass-matrix inverse scratch does not fit in [Bc | top]");

// ------------------------------------------------------------------ kinematics shared by running / terminal knots
// fills oM, S, I, v, hb, Ic, hsub, com, ofoot, Jf
template <class WS> HD void mb_kinematics(const DevModel &m, WS &w) {
  const mpc_robot_t &rb = m.rb;
  const double *q = w.x, *qd = w.x + NQ;
  // every body's placement relative to its parent (joint placement x joint rotation) is independent of the tree: all bodies
  // in parallel, off the level chain; the chain itself is then one SE(3) product per level
  PAR_FOR(b, NB) {
    double *o = w.oM + 12 * b;
    if (b == 0) { quat_to_R(q + 3, o); o[9] = q[0]; o[10] = q[1]; o[11] = q[2]; }
    else {
      // Rodrigues with the unit joint axis: R = I + sin(q) [a]x + (1 - cos q) [a]x^2
      double sn, cs, jr[12], K[9], K2[9];
      sincos(q[6 + b], &sn, &cs);
      skew3(rb.axis[b], K); mat3_mul(K, K, K2);
      for (int i = 0; i < 9; i++) jr[i] = ((i % 4 == 0) ? 1.0 : 0.0) + sn * K[i] + (1.0 - cs) * K2[i];
      jr[9] = jr[10] = jr[11] = 0;
      se3_mul(rb.jplace[b], jr, o);
    }
  }
  SYNC();
  for (int l = 1; l < m.nlevels; l++) {
    int n = m.level_start[l + 1] - m.level_start[l];
    PAR_FOR(t, n) {
      int b = m.level_body[m.level_start[l] + t];
      double loc[12];
      for (int i = 0; i < 12; i++) loc[i] = w.oM[12 * b + i];
      se3_mul(w.oM + 12 * rb.parent[b], loc, w.oM + 12 * b);
    }
    SYNC();
  }
  PAR_FOR(j, NV) {
    int b = body_of_dof(j);
    double e[6] = {0, 0, 0, 0, 0, 0};
    if (j < 6) e[j] = 1.0; else { e[3] = rb.axis[b][0]; e[4] = rb.axis[b][1]; e[5] = rb.axis[b][2]; }
    se3_act_motion(w.oM + 12 * b, e, w.S + 6 * j);
  }
  PAR_FOR(b, NB) {
    const double *o = w.oM + 12 * b;
    double cw[3], RI[9], Iw[9];
    mat3_vec(o, rb.com[b], cw);
    for (int i = 0; i < 3; i++) cw[i] += o[9 + i];
    mat3_mul(o, rb.inertia[b], RI);
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) Iw[3 * i + j] = RI[3 * i] * o[3 * j] + RI[3 * i + 1] * o[3 * j + 1] + RI[3 * i + 2] * o[3 * j + 2];
    double ms = rb.mass[b], c2 = dot3(cw, cw);
    double *I = w.I + 10 * b;
    I[0] = ms; I[1] = ms * cw[0]; I[2] = ms * cw[1]; I[3] = ms * cw[2];
    I[4] = Iw[0] + ms * (c2 - cw[0] * cw[0]); I[5] = Iw[1] - ms * cw[0] * cw[1]; I[6] = Iw[2] - ms * cw[0] * cw[2];
    I[7] = Iw[4] + ms * (c2 - cw[1] * cw[1]); I[8] = Iw[5] - ms * cw[1] * cw[2]; I[9] = Iw[8] + ms * (c2 - cw[2] * cw[2]);
  }
  PAR_FOR(f, 2) se3_mul(w.oM + 12 * rb.foot_body[f], rb.foot_place[f], w.ofoot + 12 * f);
  SYNC();
  PAR_FOR(e, NB * 6) {
    int b = e / 6, c = e % 6;
    uint32_t mask = m.ancdof_mask[b];
    double s = 0;
    for (int j = 0; j < NV; j++) if (mask >> j & 1) s += w.S[6 * j + c] * qd[j];
    w.v[e] = s;
  }
  PAR_FOR(e, 2 * NV) {
    int f = e / NV, j = e % NV;
    double col[6] = {0, 0, 0, 0, 0, 0};
    if (m.ancdof_mask[rb.foot_body[f]] >> j & 1) se3_actinv_motion(w.ofoot + 12 * f, w.S + 6 * j, col);
    for (int r = 0; r < 6; r++) w.Jf[(6 * f + r) * NV + j] = col[r];
  }
  PAR_FOR(e, NB * 10) {
    int b = e / 10, c = e % 10;
    uint32_t mask = m.sub_mask[b];
    double s = 0;
    for (int d = b; d < NB; d++) if (mask >> d & 1) s += w.I[10 * d + c];
    w.Ic[e] = s;
  }
  SYNC();
  PAR_FOR(b, NB) inertia_mul(w.I + 10 * b, w.v + 6 * b, w.hb + 6 * b);
  PAR_FOR(j, NV) inertia_mul(w.Ic + 10 * body_of_dof(j), w.S + 6 * j, w.U + 6 * j);
  SYNC();
  PAR_FOR(e, NB * 6) {
    int b = e / 6, c = e % 6;
    uint32_t mask = m.sub_mask[b];
    double s = 0;
    for (int d = b; d < NB; d++) if (mask >> d & 1) s += w.hb[6 * d + c];
    w.hsub[e] = s;
  }
  ONE_THREAD { for (int i = 0; i < 3; i++) w.com[i] = w.Ic[1 + i] / w.Ic[0]; }
  SYNC();
}

// Composite B of every subtree, Bc_b e_c = sum_{k in sub(b)} I_k (e_c x v_k) + e_c x* (I_k v_k) + v_k x* (I_k e_c):
// first each body's own 6 x 6 block (one column per work item, all bodies in parallel), then the subtree sums with one
// thread per matrix element walking the bodies leaves-to-root (bodies are ordered parents-first, so a child is complete
// before it is added to its parent): no barrier inside the accumulation and no O(subtree) work per item.
template <class WS> HD void mb_composite_B(const DevModel &m, WS &w) {
  PAR_FOR(e, NB * 6) {
    const int b = e / 6, c = e % 6;
    double ec[6] = {0, 0, 0, 0, 0, 0}, t0[6], t1[6], t2[6], t3[6], t4[6];
    ec[c] = 1.0;
    cross_mm(ec, w.v + 6 * b, t0);
    inertia_mul(w.I + 10 * b, t0, t1);
    cross_mf(ec, w.hb + 6 * b, t2);
    inertia_mul(w.I + 10 * b, ec, t3);
    cross_mf(w.v + 6 * b, t3, t4);
    for (int i = 0; i < 6; i++) w.Bc[36 * b + 6 * i + c] = t1[i] + t2[i] + t4[i];
  }
  SYNC();
  PAR_FOR(e, 36) {
    for (int b = NB - 1; b > 0; b--) w.Bc[36 * m.rb.parent[b] + e] += w.Bc[36 * b + e];
  }
  SYNC();
}

// centroidal momentum residual + Jacobian [dh/dq | A_g] (6 x 56), pose residuals + Jacobians, state error and — for running
// knots (gap_out != nullptr) — the base part of the semi-implicit Euler step, the shooting gap and the 6x6 Lie-group blocks
// P1 = Jlog6(D) Jexp6(dq), P2 = Jlog6(D) Ad(exp6(dq))^-1, E6 = -Jlog6(D) Ad(D^-1), T6 = -E6^-1 = Ad(D) Jexp6(log6 D).
// The single-thread Lie-group tasks are spread over the 4 warps in two dependent stages, the 6x6 products over all threads.
template <class WS, bool ROLL = false> HD void mb_cost_terms(const DevModel &m, WS &w, const double *lf_ref, const double *rf_ref, bool derivs, double *gap_out, const double *slack = nullptr) {
  const mpc_robot_t &rb = m.rb;
  // stage 1a: the relative placements whose logarithms are needed: both foot poses, the state error and the shooting gap
  PAR_FOR(task, 3 * 32) {
    if (task == 0) { // centroidal momentum value; integrator base block and gap placement
      const double *h = w.hsub;
      double c[3];
      cross3(w.com, h, c);
      for (int i = 0; i < 3; i++) { w.rcent[i] = h[i]; w.rcent[3 + i] = h[3 + i] - c[i]; }
      if (gap_out) {
        double pn[3], Mn[12], Mp[12];
        exp6(w.dx, w.eexp);
        mat3_vec(w.oM, w.eexp + 9, pn);
        for (int i = 0; i < 3; i++) w.xnext[i] = w.x[i] + pn[i];
        quat_integrate(w.x + 3, w.dx + 3, w.xnext + 3);
        if (ROLL) { // nonlinear rollout: x_{k+1} := f(x, u) (+) slack (base part; joints and velocities below)
          double es[12], Rn[9], ps[3];
          exp6(slack, es); quat_to_R(w.xnext + 3, Rn); mat3_vec(Rn, es + 9, ps);
          for (int i = 0; i < 3; i++) w.xn[i] = w.xnext[i] + ps[i];
          quat_integrate(w.xnext + 3, slack + 3, w.xn + 3);
        }
        quat_to_R(w.xn + 3, Mn); Mn[9] = w.xn[0]; Mn[10] = w.xn[1]; Mn[11] = w.xn[2];
        quat_to_R(w.xnext + 3, Mp); Mp[9] = w.xnext[0]; Mp[10] = w.xnext[1]; Mp[11] = w.xnext[2];
        se3_inv_mul(Mn, Mp, w.Dgap);
      }
    } else if (task == 32 || task == 64) { // foot pose placements
      int f = task == 32 ? 0 : 1;
      se3_inv_mul(f == 0 ? lf_ref : rf_ref, w.ofoot + 12 * f, w.Dl + 12 * f);
    } else if (task == 65) { // state error e = x (-) x_ref (second lane of the third warp: groups may have only three warps)
      double Mr[12], Mx[12];
      quat_to_R(m.cfg.x_ref + 3, Mr); Mr[9] = m.cfg.x_ref[0]; Mr[10] = m.cfg.x_ref[1]; Mr[11] = m.cfg.x_ref[2];
      for (int i = 0; i < 12; i++) Mx[i] = w.oM[i];
      se3_inv_mul(Mr, Mx, w.Dl + 24);
    }
  }
  SYNC();
  // stage 1b: ONE copy of log6 / Jlog6 run by up to four lanes of a warp side by side (the evaluation kernels are
  // instruction-fetch bound: every inlined copy of these long routines costs cold instruction-cache misses)
  PAR_FOR(item, gap_out ? 4 : 3) {
    const double *D = (item == 3) ? w.Dgap : w.Dl + 12 * item;
    double *out = (item < 2) ? w.rpose + 6 * item : (item == 2) ? w.estate : w.lgap;
    log6(D, out);
    if (derivs) Jlog6_from_log(out, (item < 2) ? w.Jlp + 36 * item : (item == 2) ? w.Jls : w.late.Jlg);
  }
  SYNC();
  if (gap_out) PAR_FOR(i, 6) gap_out[i] = w.fbr[i] = w.lgap[i]; // fbr temporarily holds the gap
  PAR_FOR(i, FN - 6) {
    int a = 6 + i;
    w.estate[a] = (a < NV) ? (w.x[7 + a - 6] - m.cfg.x_ref[7 + a - 6]) : (w.x[NQ + a - NV] - m.cfg.x_ref[NQ + a - NV]);
  }
  if (gap_out) {
    PAR_FOR(i, NJ) { w.xnext[7 + i] = w.x[7 + i] + w.dx[6 + i]; if (ROLL) w.xn[7 + i] = w.xnext[7 + i] + slack[6 + i]; gap_out[6 + i] = w.fbr[6 + i] = w.xnext[7 + i] - w.xn[7 + i]; }
    PAR_FOR(i, NV) { w.xnext[NQ + i] = w.x[NQ + i] + w.dx[NV + i]; if (ROLL) w.xn[NQ + i] = w.xnext[NQ + i] + slack[NV + i]; gap_out[NV + i] = w.fbr[NV + i] = w.xnext[NQ + i] - w.xn[NQ + i]; }
  }
  SYNC();
  if (!derivs) return;
  if (gap_out) {
    PAR_FOR(task, 3 * 32) {
      const double id[12] = {1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0};
      if (task == 0) { double inv[12]; se3_inv_mul(w.Dgap, id, inv); se3_action_matrix(inv, w.late.AdDi); }
      else if (task == 32 || task == 33) Jexp6(task == 32 ? w.dx : w.lgap, task == 32 ? w.late.Jd : w.late.Jrg); // two lanes, one copy
      else if (task == 64) { double einv[12]; se3_inv_mul(w.eexp, id, einv); se3_action_matrix(einv, w.late.Ade); se3_action_matrix(w.Dgap, w.late.AdD); }
    }
  }
  PAR_FOR(e, 2 * 6 * NV) { // Jpose = Jlog6 * Jf
    int f = e / (6 * NV), r = (e / NV) % 6, j = e % NV;
    const double *Jl = w.Jlp + 36 * f;
    double s = 0;
    for (int k = 0; k < 6; k++) s += Jl[6 * r + k] * w.Jf[(6 * f + k) * NV + j];
    w.cj.Jpose[e] = s;
  }
  PAR_FOR(j, NV) { // centroidal derivative columns
    int J = body_of_dof(j), pJ = rb.parent[J];
    const double *s = w.S + 6 * j;
    double vp[6] = {0, 0, 0, 0, 0, 0};
    if (pJ >= 0) for (int i = 0; i < 6; i++) vp[i] = w.v[6 * pJ + i];
    double wj[6], t1[6], t2[6], dho[6], Is[6];
    cross_mm(s, vp, wj);
    cross_mf(s, w.hsub + 6 * J, t1);
    inertia_mul(w.Ic + 10 * J, wj, t2);
    for (int i = 0; i < 6; i++) dho[i] = t1[i] - t2[i];
    inertia_mul(w.Ic + 10 * J, s, Is);
    const double im_ = rcp_(w.Ic[0]);
    double dc[3] = {Is[0] * im_, Is[1] * im_, Is[2] * im_};
    double c1[3], c2[3], c3[3];
    cross3(dc, w.hsub, c1); cross3(w.com, dho, c2); cross3(w.com, Is, c3);
    for (int r = 0; r < 3; r++) {
      w.cj.Jcent[r * FN + j] = dho[r]; w.cj.Jcent[(3 + r) * FN + j] = dho[3 + r] - c1[r] - c2[r];
      w.cj.Jcent[r * FN + NV + j] = Is[r]; w.cj.Jcent[(3 + r) * FN + NV + j] = Is[3 + r] - c3[r];
    }
  }
  SYNC();
  if (gap_out) {
    PAR_FOR(e, 4 * 36) {
      int which = e / 36, i = (e % 36) / 6, j = e % 6;
      const double *A = (which == 3) ? w.late.AdD : w.late.Jlg;
      const double *B = (which == 0) ? w.late.AdDi : (which == 1) ? w.late.Jd : (which == 2) ? w.late.Ade : w.late.Jrg;
      double sacc = 0;
      for (int k = 0; k < 6; k++) sacc += A[6 * i + k] * B[6 * k + j];
      if (which == 0) w.late.E6[6 * i + j] = -sacc; else if (which == 1) w.late.P1[6 * i + j] = sacc; else if (which == 2) w.late.P2[6 * i + j] = sacc; else w.late.T6[6 * i + j] = sacc;
    }
    SYNC();
  }
}

// value + gradient + Gauss-Newton Hessian contributions of the multibody cost stack at (a, b) / z
template <class WS> HD double mb_cost_value(const WS &w, const double *wx, const double *wcent, const double *wlf, const double *wrf) {
  double c = 0;
  for (int i = 0; i < FN; i++) c += 0.5 * wx[i] * w.estate[i] * w.estate[i];
  for (int i = 0; i < 6; i++) c += 0.5 * (wcent[i] * w.rcent[i] * w.rcent[i] + wlf[i] * w.rpose[i] * w.rpose[i] + wrf[i] * w.rpose[6 + i] * w.rpose[6 + i]);
  return c;
}
template <class WS> HD double mb_cost_grad(const WS &w, const double *wx, const double *wcent, const double *wlf, const double *wrf, int z) {
  double g = 0;
  if (z < 6) { for (int r = 0; r < 6; r++) g += wx[r] * w.Jls[6 * r + z] * w.estate[r]; }
  else if (z < FN) g += wx[z] * w.estate[z];
  if (z < FN) for (int r = 0; r < 6; r++) g += wcent[r] * w.cj.Jcent[r * FN + z] * w.rcent[r];
  if (z < NV) for (int r = 0; r < 6; r++) g += wlf[r] * w.cj.Jpose[r * NV + z] * w.rpose[r] + wrf[r] * w.cj.Jpose[(6 + r) * NV + z] * w.rpose[6 + r];
  return g;
}
template <class WS> HD double mb_cost_hess(const WS &w, const double *wx, const double *wcent, const double *wlf, const double *wrf, int a, int b) {
  double h = 0; // a <= b
  if (b < 6) { for (int r = 0; r < 6; r++) h += wx[r] * w.Jls[6 * r + a] * w.Jls[6 * r + b]; }
  else if (a == b && a < FN) h += wx[a];
  if (b < FN) for (int r = 0; r < 6; r++) if (wcent[r] != 0.0) h += wcent[r] * w.cj.Jcent[r * FN + a] * w.cj.Jcent[r * FN + b];
  if (b < NV) for (int r = 0; r < 6; r++) h += wlf[r] * w.cj.Jpose[r * NV + a] * w.cj.Jpose[r * NV + b] + wrf[r] * w.cj.Jpose[(6 + r) * NV + a] * w.cj.Jpose[(6 + r) * NV + b];
  return h;
}

// normal-cone projection pieces
HD double vplus_row(int type, double h, double ve, double mu, double lo, double hi, int &act, double &prim) {
  double wv = h + mu * ve;
  if (type == 0) { act = 1; prim = h; return wv * rcp_(mu); }
  if (type == 1) { if (wv > 0) { act = 1; prim = h; return wv * rcp_(mu); } act = 0; prim = h - wv; return 0.0; }
  if (type == 2) {
    if (wv > hi) { act = 1; prim = h - hi; return (wv - hi) * rcp_(mu); }
    if (wv < lo) { act = 1; prim = h - lo; return (wv - lo) * rcp_(mu); }
    act = 0; prim = h - wv; return 0.0;
  }
  act = 0; prim = 0; return 0.0;
}

// ------------------------------------------------------------------ running knot
template <bool DERIV, bool ROLL = false> HD void eval_full_knot(const DevModel &m, const KnotIO &io, FullWsT<DERIV> &w) {
  EPH_DECL;
  const mpc_robot_t &rb = m.rb;
  const mpc_config_t &cfg = m.cfg;
  const double dt = cfg.dt;
  PAR_FOR(i, NQ + NV) { w.x[i] = io.x[i]; w.xn[i] = io.xn[i]; }
  PAR_FOR(i, FM) w.u[i] = io.u[i];
  PAR_FOR(i, (int)(sizeof(mpc_knot_t) / 8)) w.kn[i] = reinterpret_cast<const double *>(io.kn)[i];
  SYNC();
  const mpc_knot_t &kn = *reinterpret_cast<const mpc_knot_t *>(w.kn);
  ONE_THREAD {
    bool l = kn.cs[0] != 0.0, r = kn.cs[1] != 0.0;
    if (!l && !r) l = r = true; // fulldynamic_talos.py:108-110: no-contact falls through to both contacts
    w.nact = 0; w.sidx[0] = w.sidx[1] = -1;
    if (l) { w.sidx[0] = w.nact; w.act[w.nact++] = 0; }
    if (r) { w.sidx[1] = w.nact; w.act[w.nact++] = 1; }
  }
  EPH(0);
  mb_kinematics(m, w);
  EPH(1);
  const int nact = w.nact, nk = 6 * nact;
  // derivative kernel: explicit mass-matrix inverse, raw right-hand sides [tau - b | J'] (NV x 13) in Yr; values kernel: solve in place in Y
  // [L^-1 (NV x NV) | 64 doubles of scratch per warp] in the free [Bc | top] block; the raw right-hand sides go to the global scratch
  double *Ls = reinterpret_cast<double *>(&w) + offsetof(FullWsT<DERIV>, Bc) / sizeof(double); // spans [Bc | top]
  double *X = io.scratch, *DL = io.scratch + NV * FNZ; // derivative pass only (null otherwise)
  double *Yr = DERIV ? DL : w.Y;
  const double a0[6] = {-rb.gravity[0], -rb.gravity[1], -rb.gravity[2], 0, 0, 0};
  // bias accelerations (qdd = 0, gravity folded in) and bias forces
  PAR_FOR(b, NB) {
    double acc[6] = {a0[0], a0[1], a0[2], 0, 0, 0};
    uint32_t mask = m.anc_mask[b];
    for (int k = 1; k <= b; k++)
      if (mask >> k & 1) {
        double sj[6], c[6];
        for (int i = 0; i < 6; i++) sj[i] = w.S[6 * (5 + k) + i] * w.x[NQ + 5 + k];
        cross_mm(w.v + 6 * k, sj, c);
        for (int i = 0; i < 6; i++) acc[i] += c[i];
      }
    for (int i = 0; i < 6; i++) w.a[6 * b + i] = acc[i];
    double t1[6], t2[6];
    inertia_mul(w.I + 10 * b, acc, t1);
    cross_mf(w.v + 6 * b, w.hb + 6 * b, t2);
    for (int i = 0; i < 6; i++) w.f[6 * b + i] = t1[i] + t2[i];
  }
  SYNC();
  PAR_FOR(e, NB * 6) {
    int b = e / 6, c = e % 6;
    uint32_t mask = m.sub_mask[b];
    double s = 0;
    for (int d = b; d < NB; d++) if (mask >> d & 1) s += w.f[6 * d + c];
    w.Fsub[e] = s;
  }
  PAR_FOR(e, NV * NV) { // CRBA
    int i = e / NV, j = e % NV, bi = body_of_dof(i), bj = body_of_dof(j);
    double v = 0;
    if (m.anc_mask[bj] >> bi & 1) v = dot6(w.S + 6 * i, w.U + 6 * j);
    else if (m.anc_mask[bi] >> bj & 1) v = dot6(w.S + 6 * j, w.U + 6 * i);
    w.M[e] = v;
  }
  PAR_FOR(c, nact) { // contact kinematics + Baumgarte terms
    int f = w.act[c], fb = rb.foot_body[f];
    se3_actinv_motion(w.ofoot + 12 * f, w.v + 6 * fb, w.vc + 6 * c);
    double an[6];
    for (int i = 0; i < 6; i++) an[i] = w.a[6 * fb + i] - a0[i];
    se3_actinv_motion(w.ofoot + 12 * f, an, w.gam + 6 * c);
    se3_inv_mul(w.ofoot + 12 * f, cfg.contact_place[f], w.c1Mc2 + 12 * c);
    double lg[6];
    log6(w.c1Mc2 + 12 * c, lg);
    for (int r = 0; r < 6; r++) w.astar[6 * c + r] = cfg.kp[r] * lg[r] - cfg.kd[r] * w.vc[6 * c + r];
    for (int r = 0; r < 6; r++) w.lgc[6 * c + r] = lg[r];
  }
  SYNC();
  PAR_FOR(j, NV) { // b = S^T Fsub ; rhs column 0 = tau - b ; columns 1.. = J^T
    double bj = dot6(w.S + 6 * j, w.Fsub + 6 * body_of_dof(j));
    Yr[j * 13] = (j >= 6 ? w.u[j - 6] : 0.0) - bj;
    for (int r = 0; r < nk; r++) Yr[j * 13 + 1 + r] = w.Jf[(6 * w.act[r / 6] + r % 6) * NV + j];
  }
  SYNC();
  EPH(2);
  if (DERIV) {
    // M <- M^-1 explicitly (one Cholesky + inverse factor + product): every solve with the mass matrix is then a plain matrix
    // product without barriers (Y here, the 78 right-hand sides of the tangent later).  The values-only kernel has just the
    // 13 columns of Y and keeps the factor + blocked triangular solves.
    spd_inverse_blocked(w.M, NV, NV, w.dinvM, Ls);
    EPH(3);
    PAR_FOR(e, NV * (1 + nk)) {
      const int i = e / (1 + nk), c = e % (1 + nk);
      double s = 0;
      for (int k = 0; k < NV; k++) s += w.M[i * NV + k] * Yr[k * 13 + c];
      w.Y[i * 13 + c] = s;
    }
    SYNC();
  } else {
    chol_blocked(w.M, NV, NV, w.dinvM);
    EPH(3);
    trsm_blocked(w.M, NV, NV, w.dinvM, w.Y, 1 + nk, 13);
  }
  PAR_FOR(e, nk * (nk + 1)) {
    int r = e / (nk + 1), c = e % (nk + 1);
    const double *Jr = w.Jf + (6 * w.act[r / 6] + r % 6) * NV;
    double s = 0;
    for (int i = 0; i < NV; i++) s += Jr[i] * w.Y[i * 13 + c];
    if (c == 0) w.rhs[r] = w.astar[r] - w.gam[r] - s;
    else w.G[r * 12 + c - 1] = s + ((r == c - 1) ? cfg.mu_contact : 0.0);
  }
  SYNC();
  if (DERIV) { // G <- G^-1 as well (6 x 6 or 12 x 12): lambda now, the 78 columns of dlambda later, all without triangular solves
    spd_inverse_blocked(w.G, nk, 12, w.dinvG, w.top); // `top` is free until the tangent columns are built
    PAR_FOR(r, nk) { double s = 0; for (int c = 0; c < nk; c++) s += w.G[r * 12 + c] * w.rhs[c]; w.astar[r] = s; }
    SYNC();
    PAR_FOR(r, nk) w.rhs[r] = w.astar[r];
    SYNC();
  } else {
    chol_blocked(w.G, nk, 12, w.dinvG);
    trsm_blocked(w.G, nk, 12, w.dinvG, w.rhs, 1, 1);
  }
  PAR_FOR(i, NV) {
    double s = w.Y[i * 13];
    for (int r = 0; r < nk; r++) s += w.Y[i * 13 + 1 + r] * w.rhs[r];
    w.acc[i] = s;
  }
  PAR_FOR(i, 12) { int f = i / 6; w.lam[i] = (w.sidx[f] >= 0) ? w.rhs[6 * w.sidx[f] + i % 6] : 0.0; }
  SYNC();
  PAR_FOR(i, NV) { io.xdot[i] = w.x[NQ + i]; io.xdot[NV + i] = w.acc[i]; }
  PAR_FOR(i, 12) io.lamc[i] = w.lam[i];

  EPH(4);
  if (DERIV) {
    // ---- inverse-dynamics tangent at (q, v, a) with the contact wrenches as external forces
    PAR_FOR(c, nact) { // JlAd = Jlog6(c1Mc2) * Ad(c2Mc1)
      double Jl[36], Ad[36], inv[12], id[12] = {1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0};
      Jlog6_from_log(w.lgc + 6 * c, Jl);
      se3_inv_mul(w.c1Mc2 + 12 * c, id, inv);
      se3_action_matrix(inv, Ad);
      mat6_mul(Jl, Ad, w.JlAd + 36 * c);
    }
    PAR_FOR(e, NB * 6) { // full accelerations
      int b = e / 6, c = e % 6;
      uint32_t mask = m.ancdof_mask[b];
      double s = w.a[e];
      for (int j = 0; j < NV; j++) if (mask >> j & 1) s += w.S[6 * j + c] * w.acc[j];
      w.a[e] = s;
    }
    SYNC();
    PAR_FOR(b, NB) {
      double t1[6], t2[6];
      inertia_mul(w.I + 10 * b, w.a + 6 * b, t1);
      cross_mf(w.v + 6 * b, w.hb + 6 * b, t2);
      for (int i = 0; i < 6; i++) t1[i] += t2[i];
      for (int c = 0; c < nact; c++)
        if (rb.foot_body[w.act[c]] == b) {
          double fw[6];
          se3_act_force(w.ofoot + 12 * w.act[c], w.rhs + 6 * c, fw);
          for (int i = 0; i < 6; i++) t1[i] -= fw[i];
        }
      for (int i = 0; i < 6; i++) w.f[6 * b + i] = t1[i];
    }
    PAR_FOR(e, NV * FNZ) X[e] = 0.0;
    PAR_FOR(e, 12 * FNZ) DL[e] = 0.0;
    SYNC();
    PAR_FOR(e, NB * 6) {
      int b = e / 6, c = e % 6;
      uint32_t mask = m.sub_mask[b];
      double s = 0;
      for (int d = b; d < NB; d++) if (mask >> d & 1) s += w.f[6 * d + c];
      w.Fsub[e] = s;
    }
    EPH(5);
    mb_composite_B(m, w);
    EPH(6);
    PAR_FOR(e, 2 * m.npairs) {
      int kind = e / m.npairs, p = e % m.npairs;
      int j = m.pair_j[p], mb = m.pair_m[p], J = body_of_dof(j), pJ = rb.parent[J];
      const double *s = w.S + 6 * j;
      double vp[6] = {0, 0, 0, 0, 0, 0}, ap[6] = {a0[0], a0[1], a0[2], 0, 0, 0};
      if (pJ >= 0) for (int i = 0; i < 6; i++) { vp[i] = w.v[6 * pJ + i]; ap[i] = w.a[6 * pJ + i]; }
      double cj[6], wj[6], t1[6], g[6], g2[6];
      if (kind == 0) {
        cross_mm(s, vp, wj);
        cross_mm(s, ap, cj);
        cross_mm(wj, vp, t1);
        for (int i = 0; i < 6; i++) cj[i] -= t1[i];
      } else {
        for (int i = 0; i < 6; i++) { wj[i] = s[i]; t1[i] = w.v[6 * J + i] + vp[i]; }
        cross_mm(s, t1, cj);
        for (int i = 0; i < 6; i++) cj[i] = -cj[i];
      }
      inertia_mul(w.Ic + 10 * mb, cj, g);
      mat6_vec(w.Bc + 36 * mb, wj, g2);
      for (int i = 0; i < 6; i++) g[i] += g2[i];
      int i0 = first_dof(mb), nd = ndof_of(mb);
      for (int d = 0; d < nd; d++) {
        double val = dot6(w.S + 6 * (i0 + d), g);
        X[(i0 + d) * FNZ + kind * NV + j] = kind == 0 ? -val : val;
      }
      if (mb == J) {
        if (kind == 0) { cross_mf(s, w.Fsub + 6 * J, t1); for (int i = 0; i < 6; i++) w.top[6 * j + i] = t1[i] - g[i]; }
        else for (int i = 0; i < 6; i++) w.top[6 * (NV + j) + i] = g[i];
      }
    }
    PAR_FOR(e, NJ) X[(6 + e) * FNZ + 2 * NV + e] = -1.0; // R1 for u: -B_act
    // R2 = d(alpha - astar)/d(q,v) for the active contacts
    PAR_FOR(e, nact * NV) {
      int c = e / NV, j = e % NV, f = w.act[c], fb = rb.foot_body[f];
      if (!(m.ancdof_mask[fb] >> j & 1)) continue;
      int J = body_of_dof(j), pJ = rb.parent[J];
      const double *s = w.S + 6 * j, *of = w.ofoot + 12 * f;
      double vp[6] = {0, 0, 0, 0, 0, 0}, ap[6] = {0, 0, 0, 0, 0, 0};
      if (pJ >= 0) for (int i = 0; i < 6; i++) { vp[i] = w.v[6 * pJ + i]; ap[i] = w.a[6 * pJ + i] - a0[i]; }
      double wj[6], cj[6], t1[6], t2[6], dq_[6], dv_[6], wl[6], Jc[6], dlog[6];
      cross_mm(s, vp, wj);
      cross_mm(s, ap, cj);
      cross_mm(wj, vp, t1);
      cross_mm(wj, w.v + 6 * fb, t2);
      for (int i = 0; i < 6; i++) t1[i] = cj[i] - t1[i] + t2[i];
      se3_actinv_motion(of, t1, dq_);
      for (int i = 0; i < 6; i++) t2[i] = w.v[6 * fb + i] - w.v[6 * J + i] - vp[i];
      cross_mm(s, t2, t1);
      se3_actinv_motion(of, t1, dv_);
      se3_actinv_motion(of, wj, wl);
      for (int r = 0; r < 6; r++) Jc[r] = w.Jf[(6 * f + r) * NV + j];
      mat6_vec(w.JlAd + 36 * c, Jc, dlog);
      for (int r = 0; r < 6; r++) {
        DL[(6 * c + r) * FNZ + j] = -dq_[r] - (-cfg.kp[r] * dlog[r] + cfg.kd[r] * wl[r]);
        DL[(6 * c + r) * FNZ + NV + j] = dv_[r] + cfg.kd[r] * Jc[r];
      }
    }
    SYNC();
    PAR_FOR(e, m.nanc) {
      int i = m.anc_i[e], j = m.anc_j[e];
      X[i * FNZ + j] = dot6(w.S + 6 * i, w.top + 6 * j);
      X[i * FNZ + NV + j] = dot6(w.S + 6 * i, w.top + 6 * (NV + j));
    }
    SYNC();
    EPH(7);
#ifdef MPC_HOST_EMU
    PAR_FOR(z, FNZ) { // X = M^-1 R1, one column per thread (column in registers, written back in place)
      double col[NV];
#pragma unroll
      for (int k = 0; k < NV; k++) col[k] = X[k * FNZ + z];
      for (int i = 0; i < NV; i++) {
        double s = 0;
#pragma unroll
        for (int k = 0; k < NV; k++) s += w.M[i * NV + k] * col[k];
        X[i * FNZ + z] = s;
      }
    }
    SYNC();
    EPH(8);
    PAR_FOR(z, FNZ) { // per column of z: dlam = G^-1 (J X - DL), then da/dz = -X + Y dlam (X column and dlam in registers)
      double xc[NV], t[12], dl[12];
#pragma unroll
      for (int i = 0; i < NV; i++) xc[i] = X[i * FNZ + z];
#pragma unroll
      for (int r = 0; r < 12; r++) {
        t[r] = 0.0;
        if (r < nk) {
          const double *Jr = w.Jf + (6 * w.act[r / 6] + r % 6) * NV;
          double s = -DL[r * FNZ + z];
#pragma unroll
          for (int i = 0; i < NV; i++) s += Jr[i] * xc[i];
          t[r] = s;
        }
      }
#pragma unroll
      for (int r = 0; r < 12; r++) {
        dl[r] = 0.0;
        if (r < nk) {
          double s = 0;
#pragma unroll
          for (int c = 0; c < 12; c++) if (c < nk) s += w.G[r * 12 + c] * t[c];
          dl[r] = s;
          DL[r * FNZ + z] = s;
        }
      }
#pragma unroll
      for (int i = 0; i < NV; i++) {
        double s = -xc[i];
#pragma unroll
        for (int r = 0; r < 12; r++) if (r < nk) s += w.Y[i * 13 + 1 + r] * dl[r];
        X[i * FNZ + z] = s; // da/dz
      }
    }
    SYNC();
#else
    { // Tangent of the constrained dynamics on the FP64 tensor pipe, one 8-column tile of z per warp, no block barrier inside:
      //   X <- M^-1 R1;  T = J X - DL;  dlam = G^-1 T (-> DL);  da/dz = -X + Y dlam (-> X)
      // Accumulator tiles that feed the next product as its K operand are transposed through a per-warp slice of the dead
      // [Bc | top] block (32 x 8 for X, 16 x 8 for T / dlam); the right-hand sides of the next tile are fetched from the
      // L2-resident scratch while the current one is being processed.
      const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
      double *Xs = Ls + 384 * WARP_ID, *Ts = Xs + 256;
      static_assert(FullWsT<DERIV>::BCS + FullWsT<DERIV>::TOPS >= 3 * 384 || !DERIV, "per-warp transpose tiles do not fit in [Bc | top]");
      const double *Jb = w.Jf + 6 * w.act[0] * NV; // rows of the active contacts are contiguous in Jf
      const int ks2 = (nk + 3) / 4;
      constexpr int NCT = (FNZ + 7) / 8;
      double bx[7], bnext[7], xn[4][2];
      { const int cb = 8 * WARP_ID + g;
#pragma unroll
        for (int s_ = 0; s_ < 7; s_++) bnext[s_] = (WARP_ID < NCT && cb < FNZ) ? X[(4 * s_ + t) * FNZ + cb] : 0.0; }
      for (int ct = WARP_ID; ct < NCT; ct += NWARPS) {
        const int cc = 8 * ct + 2 * t;
        const bool okc = cc < FNZ;
#pragma unroll
        for (int s_ = 0; s_ < 7; s_++) bx[s_] = bnext[s_];
        { const int cbn = 8 * (ct + NWARPS) + g; // right-hand sides of this warp's next tile
#pragma unroll
          for (int s_ = 0; s_ < 7; s_++) bnext[s_] = (ct + NWARPS < NCT && cbn < FNZ) ? X[(4 * s_ + t) * FNZ + cbn] : 0.0; }
#pragma unroll
        for (int ti = 0; ti < 4; ti++) {
          const int row = 8 * ti + g;
          double c0 = 0.0, c1 = 0.0;
#pragma unroll
          for (int s_ = 0; s_ < 7; s_++) dmma_8x8x4(c0, c1, (row < NV) ? w.M[(4 * s_ + t) * NV + row] : 0.0, bx[s_]);
          xn[ti][0] = c0; xn[ti][1] = c1;
          *reinterpret_cast<double2 *>(Xs + row * 8 + 2 * t) = make_double2(c0, c1);
        }
        __syncwarp();
#pragma unroll
        for (int s_ = 0; s_ < 7; s_++) bx[s_] = Xs[(4 * s_ + t) * 8 + g];
#pragma unroll
        for (int ti = 0; ti < 2; ti++) {
          const int row = 8 * ti + g;
          double c0 = 0.0, c1 = 0.0;
          if (row < nk && okc) { const double2 d = *reinterpret_cast<const double2 *>(DL + row * FNZ + cc); c0 = -d.x; c1 = -d.y; }
#pragma unroll
          for (int s_ = 0; s_ < 7; s_++) dmma_8x8x4(c0, c1, (row < nk) ? Jb[row * NV + 4 * s_ + t] : 0.0, bx[s_]);
          *reinterpret_cast<double2 *>(Ts + row * 8 + 2 * t) = make_double2(c0, c1);
        }
        __syncwarp();
        double dl[2][2];
#pragma unroll
        for (int ti = 0; ti < 2; ti++) {
          const int row = 8 * ti + g;
          double c0 = 0.0, c1 = 0.0;
          for (int s_ = 0; s_ < ks2; s_++) {
            const int k = 4 * s_ + t;
            dmma_8x8x4(c0, c1, (row < nk && k < nk) ? w.G[row * 12 + k] : 0.0, (k < nk) ? Ts[k * 8 + g] : 0.0);
          }
          dl[ti][0] = c0; dl[ti][1] = c1;
        }
        __syncwarp();
#pragma unroll
        for (int ti = 0; ti < 2; ti++) {
          const int row = 8 * ti + g;
          *reinterpret_cast<double2 *>(Ts + row * 8 + 2 * t) = make_double2(dl[ti][0], dl[ti][1]);
          if (row < nk && okc) *reinterpret_cast<double2 *>(DL + row * FNZ + cc) = make_double2(dl[ti][0], dl[ti][1]);
        }
        __syncwarp();
#pragma unroll
        for (int ti = 0; ti < 4; ti++) {
          const int row = 8 * ti + g;
          double c0 = -xn[ti][0], c1 = -xn[ti][1];
          for (int s_ = 0; s_ < ks2; s_++) {
            const int k = 4 * s_ + t;
            dmma_8x8x4(c0, c1, (row < NV && k < nk) ? w.Y[row * 13 + 1 + k] : 0.0, (k < nk) ? Ts[k * 8 + g] : 0.0);
          }
          if (row < NV && okc) *reinterpret_cast<double2 *>(X + row * FNZ + cc) = make_double2(c0, c1);
        }
        __syncwarp();
      }
    }
    SYNC();
#endif
  }

  EPH(9);
  // ---- semi-implicit Euler + gap (App. A2)
  PAR_FOR(i, NV) { double dv = dt * w.acc[i]; w.dx[NV + i] = dv; w.dx[i] = dt * (w.x[NQ + i] + dv); }
  // stage the multipliers (their global-load latency hides behind the cost-term phase)
  PAR_FOR(i, FNC) { w.mv[i] = io.v[i]; w.mvp[i] = io.v_prev[i]; }
  PAR_FOR(i, FN) { w.mln[i] = io.lam_n[i]; w.mlnp[i] = io.lam_n_prev[i]; w.mlk[i] = io.lam_k[i]; }
  SYNC();
  mb_cost_terms<FullWsT<DERIV>, ROLL>(m, w, kn.lf_ref, kn.rf_ref, DERIV, io.gap, io.slack);
  if (ROLL) PAR_FOR(i, NQ + NV) io.xn_out[i] = w.xn[i];

  EPH(10);
  if (DERIV) {
    // the kinematics block [oM | S | v | a | I | Ic] is dead from here on: stage the dlam/dz rows of the active contacts into it, so the
    // gradient / Hessian / constraint-row phases below read them from shared memory instead of the L2-resident scratch
    static_assert(NB * 12 + NV * 6 + 2 * NB * 6 + 2 * NB * 10 >= 12 * FNZ, "dlam/dz does not fit in the dead kinematics block");
    double *DLs = reinterpret_cast<double *>(&w) + offsetof(FullWsT<DERIV>, oM) / sizeof(double); // the whole block, not just oM
    PAR_FOR(e, nk * FNZ) DLs[e] = DL[e];
    DL = DLs; // (the barrier before the first read is the one inside the merit reduction below)
  }
  // ---- constraint values, multiplier estimates, activity
  // every thread accumulates the merit pieces of the rows / coordinates it owns; one group-wide reduction at the end
  double acc_cost = 0.0, acc_pen = 0.0, acc_prim = 0.0, acc_inner = 0.0;
  PAR_FOR(r, FNC) {
    int type = -1; double hv = 0, lo = 0, hi = 0;
    if (r < 22) { type = 2; hv = w.u[r]; lo = -rb.tau_max[r]; hi = rb.tau_max[r]; }
    else if (r < 44) { int i = r - 22; type = 2; hv = -w.x[7 + i]; lo = -rb.q_hi[i]; hi = -rb.q_lo[i]; }
    else {
      int f = (r - 44) / 17, rr = (r - 44) % 17;
      if (w.sidx[f] >= 0) { type = 1; for (int k = 0; k < 6; k++) hv += m.Acone[6 * rr + k] * w.lam[6 * f + k]; }
    }
    int act = 0; double prim = 0;
    const double vp = vplus_row(type, hv, w.mvp[r], io.mu, lo, hi, act, prim), dv = vp - w.mv[r];
    w.ctype[r] = type; w.late.hval[r] = hv; w.late.vpl[r] = vp; w.isact[r] = act;
    w.late.dbr[r] = io.mu * dv;
    io.h[r] = hv;
    if (type >= 0) {
      acc_pen += 0.5 * io.mu * (vp * vp + dv * dv);
      acc_prim = fmax(acc_prim, fabs(prim));
      acc_inner = fmax(acc_inner, fabs(io.mu * dv));
    }
  }
  PAR_FOR(i, FN) {
    const double gap = w.fbr[i], lp = w.mlnp[i] + gap * rcp_(io.mu), dl = lp - w.mln[i]; // fbr = gap here
    w.lpl[i] = lp;
    w.fbr[i] = io.mu * dl;
    acc_pen += 0.5 * io.mu * (lp * lp + dl * dl);
    acc_prim = fmax(acc_prim, fabs(gap));
    acc_inner = fmax(acc_inner, fabs(io.mu * dl));
    acc_cost += 0.5 * cfg.wx[i] * w.estate[i] * w.estate[i];
  }
  PAR_FOR(i, FM) { const double e = w.u[i] - kn.u_ref[i]; acc_cost += 0.5 * cfg.wu[i] * e * e; }
  PAR_FOR(i, 6) acc_cost += 0.5 * (cfg.w_cent[i] * w.rcent[i] * w.rcent[i] + kn.w_lf[i] * w.rpose[i] * w.rpose[i] + kn.w_rf[i] * w.rpose[6 + i] * w.rpose[6 + i]);
  PAR_FOR(e, 12) {
    const int f = e / 6, i = e % 6;
    if (kn.fcost[f] != 0.0) { const double d = w.lam[e] - kn.f_ref[e]; acc_cost += 0.5 * cfg.w_force[i] * d * d; }
  }
  reduce_sum2_max2(acc_cost, acc_pen, acc_prim, acc_inner, w.part, w.part + 16);
  ONE_THREAD { w.scal[SC_COST] = w.part[16]; w.scal[SC_PEN] = w.part[17]; w.scal[SC_PRIM] = w.part[18]; w.scal[SC_INNER] = w.part[19]; w.scal[SC_DUAL] = 0; }
  compact_flags(w.isact, FNC, w.act_idx, &w.nca, reinterpret_cast<int32_t *>(w.part));
  if (!DERIV) {
    PAR_FOR(i, SC_COUNT) io.scal[i] = w.scal[i];
    EPH(11);
    EPH_DUMP(io.phase_out);
    return;
  }

  EPH(11);
  // ---- LQ blocks to HBM: AB (56 x 78), Lagrangian gradient, cost gradient / Hessian, active constraint rows
  PAR_FOR(i, FNC) { io.dbar[i] = w.late.dbr[i]; io.vplus[i] = w.late.vpl[i]; io.act_idx[i] = (i < w.nca) ? w.act_idx[i] : -1; }
  PAR_FOR(i, FN) { io.fbar[i] = w.fbr[i]; io.lplus[i] = w.lpl[i]; }
  PAR_FOR(i, 36) { io.T6[i] = w.late.T6[i]; io.E6[i] = w.late.E6[i]; }
  ONE_THREAD io.nca[0] = w.nca;
  const double dt2 = dt * dt;
  PAR_FOR(z, FNZ) {
    double d6[6], acc = 0, xc[NV];
#pragma unroll
    for (int i = 0; i < NV; i++) xc[i] = X[i * FNZ + z]; // one pass over the column of the tangent (L2-resident scratch)
#pragma unroll
    for (int k = 0; k < 6; k++) d6[k] = dt2 * xc[k] + ((z == NV + k) ? dt : 0.0);
#pragma unroll
    for (int i = 0; i < 6; i++) {
      double s = (z < 6) ? w.late.P2[6 * i + z] : 0.0;
#pragma unroll
      for (int k = 0; k < 6; k++) s += w.late.P1[6 * i + k] * d6[k];
      io.AB[i * FNZ + z] = s; acc += s * w.mln[i];
    }
#pragma unroll
    for (int i = 6; i < NV; i++) {
      double s = dt2 * xc[i] + ((z == NV + i) ? dt : 0.0) + ((z == i) ? 1.0 : 0.0);
      io.AB[i * FNZ + z] = s; acc += s * w.mln[i];
    }
#pragma unroll
    for (int i = 0; i < NV; i++) {
      double s = dt * xc[i] + ((z == NV + i) ? 1.0 : 0.0);
      io.AB[(NV + i) * FNZ + z] = s; acc += s * w.mln[NV + i];
    }
    // cost gradient
    double lz = mb_cost_grad(w, cfg.wx, cfg.w_cent, kn.w_lf, kn.w_rf, z);
    if (z >= FN) lz += cfg.wu[z - FN] * (w.u[z - FN] - kn.u_ref[z - FN]);
    for (int f = 0; f < 2; f++)
      if (kn.fcost[f] != 0.0) {
        const double *Dl = DL + 6 * w.sidx[f] * FNZ;
        for (int r = 0; r < 6; r++) lz += cfg.w_force[r] * Dl[r * FNZ + z] * (w.lam[6 * f + r] - kn.f_ref[6 * f + r]);
      }
    w.late.lxu[z] = lz; io.lxu[z] = lz;
    // Lagrangian gradient: + C^T v  (+ E_{k-1}^T lam_k: vector part here, base block added by the reduction kernel)
    double gz = lz + acc;
    for (int r = 0; r < FNC; r++) {
      double vr = w.mv[r];
      if (vr == 0.0 || w.ctype[r] < 0) continue;
      double c;
      if (r < 22) c = (z == FN + r) ? 1.0 : 0.0;
      else if (r < 44) c = (z == 6 + r - 22) ? -1.0 : 0.0;
      else { int f = (r - 44) / 17, rr = (r - 44) % 17; const double *Dl = DL + 6 * w.sidx[f] * FNZ; c = 0; for (int k = 0; k < 6; k++) c += m.Acone[6 * rr + k] * Dl[k * FNZ + z]; }
      gz += c * vr;
    }
    if (z < FN) { if (io.k == 0) gz += w.mlk[z]; else if (z >= 6) gz -= w.mlk[z]; }
    w.late.g[z] = gz; io.g[z] = gz;
  }
  PAR_FOR(j, 6) { // E_k^T lam_{k+1}, base block, for the next knot's gradient
    double s = 0;
    for (int i = 0; i < 6; i++) s += w.late.E6[6 * i + j] * w.mln[i];
    io.gE_next[j] = s;
  }
  EPH(12);
#ifdef MPC_HOST_EMU
  { // Gauss-Newton Hessian H = sum_r w_r J_r' J_r (+ state/control diagonals) on 4 x 4 register tiles of the upper triangle
    constexpr int HT = (FNZ + 3) / 4;
    PAR_FOR(t, HT * HT) {
      const int ta = t / HT, tb = t % HT;
      if (tb < ta) continue;
      const int a0 = 4 * ta, b0 = 4 * tb;
      double acc[4][4];
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = 0.0;
      auto rank1 = [&](const double *J, int ncols, double wgt) {
        double va[4], vb[4];
#pragma unroll
        for (int i = 0; i < 4; i++) { va[i] = (a0 + i < ncols) ? wgt * J[a0 + i] : 0.0; vb[i] = (b0 + i < ncols) ? J[b0 + i] : 0.0; }
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
          for (int j = 0; j < 4; j++) acc[i][j] += va[i] * vb[j];
      };
      for (int f = 0; f < 2; f++)
        if (kn.fcost[f] != 0.0) for (int r = 0; r < 6; r++) rank1(DL + (6 * w.sidx[f] + r) * FNZ, FNZ, cfg.w_force[r]);
      if (a0 < FN) {
        for (int r = 0; r < 6; r++) if (cfg.w_cent[r] != 0.0) rank1(w.cj.Jcent + r * FN, FN, cfg.w_cent[r]);
        if (a0 < NV && b0 < NV)
          for (int r = 0; r < 6; r++) {
            if (kn.w_lf[r] != 0.0) rank1(w.cj.Jpose + r * NV, NV, kn.w_lf[r]);
            if (kn.w_rf[r] != 0.0) rank1(w.cj.Jpose + (6 + r) * NV, NV, kn.w_rf[r]);
          }
        if (b0 < 8 && a0 < 8) for (int r = 0; r < 6; r++) rank1(w.Jls + 6 * r, 6, cfg.wx[r]);
      }
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const int a = a0 + i, b = b0 + j;
          if (a >= FNZ || b >= FNZ || b < a) continue;
          double hv = acc[i][j];
          if (a == b) { hv += io.preg; if (a >= FN) hv += cfg.wu[a - FN]; else if (a >= 6) hv += cfg.wx[a]; }
          io.H[a * FNZ + b] = hv; io.H[b * FNZ + a] = hv;
        }
    }
  }
#else
  { // Gauss-Newton Hessian H = sum_k w_k J_k' J_k (+ state/control diagonals) as ONE tensor-core product over the 36 weighted
    // residual-Jacobian rows (12 contact-force rows dlam/dz, 6 centroidal momentum, 12 foot pose, 6 of the base block of the state
    // error; absent costs keep weight 0).  The rows stay where they are in shared memory and are addressed through a table of
    // offsets; every warp takes 2 x 2 blocks of 8 x 8 tiles of the FULL matrix, so the result leaves as row-major 16-byte stores.
    const double *base = reinterpret_cast<const double *>(&w);
    PAR_FOR(k, 36) {
      const double *p = w.Jls; int len = 0; double wgt = 0.0;
      if (k < 12) { const int f = k / 6, r = k % 6; if (kn.fcost[f] != 0.0 && w.sidx[f] >= 0) { p = DL + (6 * w.sidx[f] + r) * FNZ; len = FNZ; wgt = cfg.w_force[r]; } }
      else if (k < 18) { p = w.cj.Jcent + (k - 12) * FN; len = FN; wgt = cfg.w_cent[k - 12]; }
      else if (k < 30) { const int j = k - 18, r = j % 6; p = w.cj.Jpose + j * NV; len = NV; wgt = (j < 6) ? kn.w_lf[r] : kn.w_rf[r]; }
      else { p = w.Jls + 6 * (k - 30); len = 6; wgt = cfg.wx[k - 30]; }
      w.hrow_len[k] = (int32_t)(p - base) | (len << 20); // offset (doubles) and row length packed in one word
      w.hrow_w[k] = wgt;
    }
    SYNC();
    constexpr int NBK = ((FNZ + 7) / 8 + 1) / 2; // 5 blocks of 16 per side
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    for (int p = WARP_ID; p < NBK * NBK; p += NWARPS) {
      const int ti = 2 * (p / NBK), tj = 2 * (p % NBK);
      const int ia = ti * 8 + g, ib = tj * 8 + g;
      double c[2][2][2] = {{{0, 0}, {0, 0}}, {{0, 0}, {0, 0}}};
#pragma unroll 3
      for (int k0 = 0; k0 < 36; k0 += 4) {
        const int packed = w.hrow_len[k0 + t], off = packed & 0xFFFFF, len = packed >> 20;
        const double wk = w.hrow_w[k0 + t];
        const double *row = base + off;
        const double a0 = (ia < len) ? wk * row[ia] : 0.0, a1 = (ia + 8 < len) ? wk * row[ia + 8] : 0.0;
        const double b0 = (ib < len) ? row[ib] : 0.0, b1 = (ib + 8 < len) ? row[ib + 8] : 0.0;
        dmma_8x8x4(c[0][0][0], c[0][0][1], a0, b0);
        dmma_8x8x4(c[0][1][0], c[0][1][1], a0, b1);
        dmma_8x8x4(c[1][0][0], c[1][0][1], a1, b0);
        dmma_8x8x4(c[1][1][0], c[1][1][1], a1, b1);
      }
#pragma unroll
      for (int a = 0; a < 2; a++)
#pragma unroll
        for (int b = 0; b < 2; b++) {
          const int r = (ti + a) * 8 + g, cc = (tj + b) * 8 + 2 * t;
          if (r >= FNZ || cc >= FNZ) continue;
          double h0 = c[a][b][0], h1 = c[a][b][1];
          const double dg = io.preg + ((r >= FN) ? cfg.wu[r - FN] : ((r >= 6) ? cfg.wx[r] : 0.0));
          if (r == cc) h0 += dg; else if (r == cc + 1) h1 += dg;
          *reinterpret_cast<double2 *>(io.H + r * FNZ + cc) = make_double2(h0, h1);
        }
    }
  }
#endif
  EPH(13);
  PAR_FOR(e, w.nca * FNZ) { // active constraint rows, compacted
    int ai = e / FNZ, z = e % FNZ, r = w.act_idx[ai];
    double c;
    if (r < 22) c = (z == FN + r) ? 1.0 : 0.0;
    else if (r < 44) c = (z == 6 + r - 22) ? -1.0 : 0.0;
    else { int f = (r - 44) / 17, rr = (r - 44) % 17; const double *Dl = DL + 6 * w.sidx[f] * FNZ; c = 0; for (int k = 0; k < 6; k++) c += m.Acone[6 * rr + k] * Dl[k * FNZ + z]; }
    io.CDact[e] = c;
  }
  SYNC();
  PAR_FOR(c, 8) { // dual residual of this knot without the base block of the x-gradient (finalised in the reduction kernel)
    double dual = 0;
    for (int z = 6 + c; z < FNZ; z += 8) {
      if (io.k == 0 && z < FN) continue;
      dual = fmax(dual, fabs(w.late.g[z]));
    }
    w.part[c] = dual;
  }
  SYNC();
  PAR_FOR(i, SC_COUNT) {
    double v = w.scal[i];
    if (i == SC_DUAL) for (int c = 0; c < 8; c++) v = fmax(v, w.part[c]);
    io.scal[i] = v;
  }
  EPH(14);
  EPH_DUMP(io.phase_out);
}

// ------------------------------------------------------------------ terminal knot (fulldynamic_talos.py:234-245, 499-507)
template <bool DERIV> HD void eval_full_term(const DevModel &m, const KnotIO &io, FullWsT<DERIV> &w) {
  const mpc_config_t &cfg = m.cfg;
  PAR_FOR(i, NQ + NV) w.x[i] = io.x[i];
  SYNC();
  mb_kinematics(m, w);
  mb_cost_terms(m, w, io.tm->lf_ref, io.tm->rf_ref, DERIV, nullptr);
  const bool has_c = io.tm->has_com_cstr != 0.0;
  PAR_FOR(r, FNC) {
    int type = -1; double hv = 0;
    if (r < 3 && has_c) { type = 0; hv = w.com[r] - io.tm->com_ref[r]; }
    int act = 0; double prim = 0;
    double vp = vplus_row(type, hv, io.v_prev[r], io.mu, 0, 0, act, prim);
    w.ctype[r] = type; w.late.hval[r] = hv; w.late.vpl[r] = vp; w.isact[r] = act; w.late.dbr[r] = io.mu * (vp - io.v[r]); w.late.rowtmp[r] = fabs(prim);
    io.h[r] = hv;
  }
  SYNC();
  ONE_THREAD {
    double cost = mb_cost_value(w, cfg.wx_term, cfg.w_cent_term, cfg.w_foot_term, cfg.w_foot_term), pen = 0, prim = 0, inner = 0;
    int nca = 0;
    for (int r = 0; r < 3; r++) {
      if (w.ctype[r] < 0) continue;
      double dv = w.late.vpl[r] - io.v[r];
      pen += 0.5 * io.mu * (w.late.vpl[r] * w.late.vpl[r] + dv * dv);
      prim = fmax(prim, w.late.rowtmp[r]); inner = fmax(inner, fabs(w.late.dbr[r]));
      if (w.isact[r]) w.act_idx[nca++] = r;
    }
    w.nca = nca;
    w.scal[SC_COST] = cost; w.scal[SC_PEN] = pen; w.scal[SC_PRIM] = prim; w.scal[SC_INNER] = inner; w.scal[SC_DUAL] = 0;
  }
  SYNC();
  if (!DERIV) { PAR_FOR(i, SC_COUNT) io.scal[i] = w.scal[i]; return; }
  PAR_FOR(i, FNC) { io.dbar[i] = w.late.dbr[i]; io.vplus[i] = w.late.vpl[i]; io.act_idx[i] = (i < w.nca) ? w.act_idx[i] : -1; }
  ONE_THREAD io.nca[0] = w.nca;
  PAR_FOR(z, FNZ) {
    double lz = (z < FN) ? mb_cost_grad(w, cfg.wx_term, cfg.w_cent_term, cfg.w_foot_term, cfg.w_foot_term, z) : 0.0;
    io.lxu[z] = lz;
    double gz = lz;
    if (z < NV && has_c) for (int r = 0; r < 3; r++) gz += io.v[r] * w.U[6 * z + r] * rcp_(w.Ic[0]);
    if (z >= 6 && z < FN) gz -= io.lam_k[z];
    w.late.g[z] = gz; io.g[z] = gz;
  }
  PAR_FOR(e, FNZ * FNZ) {
    int a = e / FNZ, b = e % FNZ;
    double hv = 0;
    if (a < FN && b < FN) { hv = (a <= b) ? mb_cost_hess(w, cfg.wx_term, cfg.w_cent_term, cfg.w_foot_term, cfg.w_foot_term, a, b)
                                          : mb_cost_hess(w, cfg.wx_term, cfg.w_cent_term, cfg.w_foot_term, cfg.w_foot_term, b, a);
      if (a == b) hv += io.preg; }
    io.H[e] = hv;
  }
  PAR_FOR(e, w.nca * FNZ) {
    int ai = e / FNZ, z = e % FNZ, r = w.act_idx[ai];
    io.CDact[e] = (z < NV) ? w.U[6 * z + r] * rcp_(w.Ic[0]) : 0.0; // Jcom column = lin(Ic_J s_j) / mass
  }
  SYNC();
  ONE_THREAD {
    double dual = 0;
    for (int z = 6; z < FN; z++) dual = fmax(dual, fabs(w.late.g[z]));
    w.scal[SC_DUAL] = dual;
  }
  SYNC();
  PAR_FOR(i, SC_COUNT) io.scal[i] = w.scal[i];
}

} // namespace mpcdev
