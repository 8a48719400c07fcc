// Device-side gait / swing-foot reference generation: what the reference does in Python every MPC tick before
// solver.run — update_timings (talos_utils.py:350-373), footTrajectory.updateTrajectory (talos_utils.py:187-327: yaw-aligned footstep
// placement, degree-8 Bezier swing curve of ndcurves, geodesic rotation interpolation), the 2 x 100 setReference / contact_poses writes
// and the stage entering the horizon (fulldynamic_talos.py:444-510, kinodynamic_talos.py:362-409, centroidal_talos.py:354-384,459) —
// for every robot of the batch in ONE kernel, straight into the solver's device-resident per-knot parameter blocks.
// One group of threads per robot: thread 0 advances the start / final pose bookkeeping from the MEASURED foot placements, then one
// knot per thread samples the curves and writes its mpc_knot_t.  Everything is a pure function of the tick counter, the schedule and
// the four persistent poses, mirroring mpc_benchmark_b200/gait.py (GaitPlan.tick), which the tests compare it with.
#pragma once
#include "../../include/mpcb200.h"
#include "dev_common.cuh"

namespace mpcdev {

struct GaitCfg { // uniform over the batch (device copy owned by the solver handle)
  int32_t kind, T, T_ds, T_ss, nph, keep_forward, n_uref, pad_;
  int32_t n_ev[2][4];     // [mirror][to_rf, to_lf, la_rf, la_lf]: number of seeded countdowns
  int32_t ev[2][4][8];    // their seeds: phase index of the event + nsteps (fulldynamic_talos.py:268-280)
  double x_forward, y_forward, foot_yaw, y_gap, z_height, apex;
  double lf0[12], rf0[12], com0[3];
  double f_half, w_lfrf;
  const int8_t *phases;   // [2 mirror][nph][2]: contact flags (left, right)
  const double *urefs;    // [n_uref][MPC_MAXU] control references of the schedule (kino / cent), or null
};

struct GaitRobot { // persistent per robot: the poses footTrajectory keeps between ticks
  double start_l[12], final_l[12], start_r[12], final_r[12];
  double next_l[12], next_r[12]; // the references of the NEXT tick (LF[1], RF[1]): the soles of a robot that tracks them exactly
  int32_t mirror, pad_;
};

HD int gait_head(const int32_t *ev, int n, int t) { // head of a countdown list after t + 1 calls of update_timings
  for (int i = 0; i < n; i++) { const int v = ev[i] - (t + 1); if (v >= 0) return v; }
  return -1;
}
HD void gait_copy12(const double *a, double *b) { for (int i = 0; i < 12; i++) b[i] = a[i]; }
HD void gait_yaw_rot(double yaw, double *R) {
  const double c = cos(yaw), s = sin(yaw);
  R[0] = c; R[1] = -s; R[2] = 0; R[3] = s; R[4] = c; R[5] = 0; R[6] = 0; R[7] = 0; R[8] = 1;
}
// pose = (Rrot * R(ref), t(ref) + Rz(yaw(ref)) * tr)   (talos_utils.py:216-243); Rrot may be null (identity)
HD void gait_step_pose(const double *ref, const double *tr, const double *Rrot, double *out) {
  double Ry[9], d[3];
  gait_yaw_rot(atan2(ref[3], ref[0]), Ry);
  mat3_vec(Ry, tr, d);
  if (Rrot) mat3_mul(Rrot, ref, out); else for (int i = 0; i < 9; i++) out[i] = ref[i];
  for (int i = 0; i < 3; i++) out[9 + i] = ref[9 + i] + d[i];
}
HD void gait_log3(const double *R, double *w) {
  const double s[3] = {0.5 * (R[7] - R[5]), 0.5 * (R[2] - R[6]), 0.5 * (R[3] - R[1])};
  const double ct = 0.5 * (R[0] + R[4] + R[8] - 1.0), sn = sqrt(s[0] * s[0] + s[1] * s[1] + s[2] * s[2]);
  const double f = (sn < 1e-12) ? 1.0 : atan2(sn, ct) / sn;
  for (int i = 0; i < 3; i++) w[i] = s[i] * f;
}
HD void gait_exp3(const double *w, double *R) {
  const double th = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  const double W[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
  double W2[9];
  mat3_mul(W, W, W2);
  const double a = (th < 1e-12) ? 1.0 : sin(th) / th, b = (th < 1e-12) ? 0.0 : (1.0 - cos(th)) / (th * th);
  for (int i = 0; i < 9; i++) R[i] = ((i % 4 == 0) ? 1.0 : 0.0) + a * W[i] + b * W2[i];
}
// sample i of footTrajectory.foot_trajectory (talos_utils.py:298-317): countdown tt = land - i
HD void gait_foot_ref(int land, int i, int T_ss, double apex, const double *start, const double *fin, const double *lg, double *out) {
  if (land < 0) { gait_copy12(start, out); return; }
  const int tt = land - i;
  if (tt <= 0) { gait_copy12(fin, out); return; }
  if (tt > T_ss) { gait_copy12(start, out); return; }
  const double s = (double)(T_ss - tt) / (double)T_ss, u = 1.0 - s;
  // degree-8 Bernstein weights; control points: 4 x start, the lifted point 0.75 start + 0.25 final, 4 x final (talos_utils.py:281-296)
  double b[9], pw_s[9], pw_u[9];
  pw_s[0] = pw_u[0] = 1.0;
  for (int k = 1; k < 9; k++) { pw_s[k] = pw_s[k - 1] * s; pw_u[k] = pw_u[k - 1] * u; }
  const double binom[9] = {1, 8, 28, 56, 70, 56, 28, 8, 1};
  for (int k = 0; k < 9; k++) b[k] = binom[k] * pw_u[8 - k] * pw_s[k];
  const double b0 = b[0] + b[1] + b[2] + b[3], b1 = b[5] + b[6] + b[7] + b[8];
  for (int c = 0; c < 3; c++) {
    const double mid = 0.75 * start[9 + c] + 0.25 * fin[9 + c] + (c == 2 ? apex : 0.0);
    out[9 + c] = b0 * start[9 + c] + b[4] * mid + b1 * fin[9 + c];
  }
  const double ws[3] = {s * lg[0], s * lg[1], s * lg[2]};
  double E[9];
  gait_exp3(ws, E);
  mat3_mul(start, E, out); // R0 exp(s log(R0' R1))
}

// One MPC tick of the reference bookkeeping for robot b: tick counter t (0-based), measured sole placements lf / rf (12 doubles each).
// Writes the T knots of the horizon the solver sees at this tick and the terminal block.  sm: 2 x 12 + 2 x 3 + 8 doubles of shared scratch.
HD void gait_tick_group(const GaitCfg &g, GaitRobot &rs, int t, const double *lf, const double *rf, mpc_knot_t *knots, mpc_term_t *term, double *sm) {
  const int T = g.T, mir = rs.mirror ? 1 : 0;
  if (!lf) { lf = rs.next_l; rf = rs.next_r; } // perfect tracking: the soles are where last tick's plan wanted them now
  double *lgl = sm, *lgr = sm + 3;
  int32_t *heads = reinterpret_cast<int32_t *>(sm + 6); // to_rf, to_lf, la_rf, la_lf
  ONE_THREAD {
    const int to_rf = gait_head(g.ev[mir][0], g.n_ev[mir][0], t), to_lf = gait_head(g.ev[mir][1], g.n_ev[mir][1], t);
    const int la_rf = gait_head(g.ev[mir][2], g.n_ev[mir][2], t), la_lf = gait_head(g.ev[mir][3], g.n_ev[mir][3], t);
    heads[0] = to_rf; heads[1] = to_lf; heads[2] = la_rf; heads[3] = la_lf;
    // the scripts zero the forward step once no further landing is pending (full:448-449, kino, cent:365-366)
    double trR[3] = {g.x_forward, -g.y_gap - g.y_forward, g.z_height}, trL[3] = {g.x_forward, g.y_gap, g.z_height};
    bool zero = false;
    if (!g.keep_forward) {
      if (g.kind == MPC_KIND_FULL) zero = la_lf == -1;
      else if (g.kind == MPC_KIND_KINO) zero = la_rf == -1 && to_rf == -1;
      else zero = la_rf == -1;
    }
    if (zero) {
      trR[0] = 0; trR[2] = 0; trL[0] = 0;
      trL[2] = (g.kind == MPC_KIND_KINO) ? 0.0 : -0.01;
    }
    double Rd[9];
    gait_yaw_rot(g.foot_yaw, Rd);
    // footTrajectory.updateTrajectory, bookkeeping part (talos_utils.py:211-243)
    if (la_lf < 0) { gait_copy12(lf, rs.start_l); gait_copy12(lf, rs.final_l); }
    if (la_rf < 0) { gait_copy12(rf, rs.start_r); gait_copy12(rf, rs.final_r); }
    if (to_rf >= 0 && to_rf < g.T_ds) {
      gait_copy12(rf, rs.start_r);
      gait_step_pose(lf, trR, Rd, rs.final_r);
      gait_copy12(lf, rs.start_l);
      gait_step_pose(rs.final_r, trL, nullptr, rs.final_l);
    }
    if (to_lf >= 0 && to_lf < g.T_ds) {
      gait_copy12(lf, rs.start_l);
      gait_step_pose(rf, trL, nullptr, rs.final_l);
      gait_copy12(rf, rs.start_r);
      gait_step_pose(rs.final_l, trR, Rd, rs.final_r);
    }
    double D[9];
    mat3_mulT(rs.start_l, rs.final_l, D); gait_log3(D, lgl);
    mat3_mulT(rs.start_r, rs.final_r, D); gait_log3(D, lgr);
  }
  SYNC();
  const int la_rf = heads[2], la_lf = heads[3];
  const int8_t *ph = g.phases + (size_t)mir * g.nph * 2;
  PAR_FOR(j, T) {
    mpc_knot_t &k = knots[j];
    // write-then-rotate: slot j carries reference j + 1, the stage that has just entered (slot T - 1) its construction default
    double lref[12], rref[12];
    if (j < T - 1) {
      gait_foot_ref(la_lf, j + 1, g.T_ss, g.apex, rs.start_l, rs.final_l, lgl, lref);
      gait_foot_ref(la_rf, j + 1, g.T_ss, g.apex, rs.start_r, rs.final_r, lgr, rref);
    } else { gait_copy12(g.lf0, lref); gait_copy12(g.rf0, rref); }
    int pidx = t - (T - 1) + j;
    if (pidx < 0) pidx = 0;
    if (pidx > g.nph - 1) pidx = g.nph - 1;
    const bool cl = ph[2 * pidx] != 0, cr = ph[2 * pidx + 1] != 0;
    k.cs[0] = cl ? 1.0 : 0.0; k.cs[1] = cr ? 1.0 : 0.0;
    k.fcost[0] = k.fcost[1] = 0.0;
    for (int i = 0; i < 6; i++) { k.w_lf[i] = k.w_rf[i] = 0.0; k.cpos[i] = 0.0; }
    for (int i = 0; i < 12; i++) { k.lf_ref[i] = k.rf_ref[i] = 0.0; k.f_ref[i] = 0.0; }
    for (int i = 0; i < MPC_MAXU; i++) k.u_ref[i] = 0.0;
    if (g.kind != MPC_KIND_CENT) {
      for (int i = 0; i < 6; i++) { k.w_rf[i] = cl ? g.w_lfrf : 0.0; k.w_lf[i] = cr ? g.w_lfrf : 0.0; } // full:179-182, kino:145-148
      gait_copy12(lref, k.lf_ref); gait_copy12(rref, k.rf_ref);
    }
    if (g.kind == MPC_KIND_FULL) {
      k.fcost[0] = cl ? 1.0 : 0.0; k.fcost[1] = cr ? 1.0 : 0.0; // full:187-201
      k.f_ref[2] = g.f_half; k.f_ref[8] = g.f_half;               // every stage is built with the force references of index 0 (full:364-367)
    } else {
      const int ui = pidx < g.n_uref ? pidx : g.n_uref - 1;
      const double *u = g.urefs + (size_t)ui * MPC_MAXU;
      for (int i = 0; i < MPC_MAXU; i++) k.u_ref[i] = u[i];
      if (mir) for (int i = 0; i < 6; i++) { k.u_ref[i] = u[6 + i]; k.u_ref[6 + i] = u[i]; } // mirrored gait: the feet swap their force references
      if (g.kind == MPC_KIND_CENT) // contact positions follow the references for ACTIVE contacts only (cent:374-384)
        for (int i = 0; i < 3; i++) { k.cpos[i] = cl ? lref[9 + i] : g.lf0[9 + i]; k.cpos[3 + i] = cr ? rref[9 + i] : g.rf0[9 + i]; }
    }
  }
  ONE_THREAD { // terminal block (full:499-510, kino:402-409; the centroidal problem has no terminal constraint)
    double lT[12], rT[12];
    gait_foot_ref(la_lf, T - 1, g.T_ss, g.apex, rs.start_l, rs.final_l, lgl, lT);
    gait_foot_ref(la_rf, T - 1, g.T_ss, g.apex, rs.start_r, rs.final_r, lgr, rT);
    const double ident[12] = {1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0};
    const double *tl = g.kind == MPC_KIND_FULL ? lT : (g.kind == MPC_KIND_KINO ? ident : g.lf0);
    const double *tr = g.kind == MPC_KIND_FULL ? rT : (g.kind == MPC_KIND_KINO ? ident : g.rf0);
    gait_copy12(tl, term->lf_ref); gait_copy12(tr, term->rf_ref);
    if (g.kind == MPC_KIND_CENT) { term->com_ref[0] = term->com_ref[1] = term->com_ref[2] = 0.0; term->has_com_cstr = 0.0; }
    else {
      term->com_ref[0] = 0.5 * (lT[9] + rT[9]); term->com_ref[1] = 0.5 * (lT[10] + rT[10]); term->com_ref[2] = g.com0[2];
      term->has_com_cstr = 1.0;
    }
    double nl[12], nr[12];
    gait_foot_ref(la_lf, 1, g.T_ss, g.apex, rs.start_l, rs.final_l, lgl, nl);
    gait_foot_ref(la_rf, 1, g.T_ss, g.apex, rs.start_r, rs.final_r, lgr, nr);
    gait_copy12(nl, rs.next_l); gait_copy12(nr, rs.next_r);
  }
  SYNC();
}

} // namespace mpcdev
