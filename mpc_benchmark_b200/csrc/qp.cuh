// Batched dense QP solver: ONE group of threads per QP, every matrix of the QP resident in shared memory.
// Boundary replaced: proxsuite.proxqp.dense.QP.solve() as driven by QP_utils.py:500-513,557-573 (IDSolver_ulim) and the other
// solver classes of that file.  Normative algorithm: oracle/qp.hpp (ProxQP restated: proximal augmented Lagrangian, semismooth
// Newton with an exact piecewise-quadratic linesearch, BCL outer loop); this file mirrors it step by step.
// Written in the group idiom of dev_common.cuh (PAR_FOR / SYNC / ONE_THREAD), so the same source runs serially under the
// test-only host emulation (tests/emu).
//
// Shared memory per QP (n = 62, n_eq = 40, n_in = 18: 107 KB, two QPs per SM): H, the Newton matrix K (factored in place),
// A, C — all with an ODD leading dimension, so that one-row-per-thread and one-column-per-thread sweeps are both free of bank
// conflicts — and the vectors.  HBM traffic = the QP data once in, (x, y, z, info) once out.
#pragma once
#include "../../include/mpcqp_b200.h"
#include "dev_common.cuh"

namespace mpcdev {

struct QPArgs {
  int n, ne, ni, box, batch;
  const double *H, *g, *A, *b, *C, *l, *u, *lb, *ub;
  long long sH, sg, sA, sb, sC, sl, su, slb, sub; // batch strides in doubles (0 = shared)
  double *x, *y, *z;
  mpc_qp_info_t *info;
  mpc_qp_settings_t st;
};

HD int qp_ld(int n) { return n | 1; }
// doubles of shared memory one QP needs (host + device)
#ifdef MPC_HOST_EMU
#define QP_HOST_DEV inline
#else
#define QP_HOST_DEV inline __host__ __device__
#endif
QP_HOST_DEV int qp_smem_doubles(int n, int ne, int ni, int box) {
  const int ld = n | 1, nz = ni + (box ? n : 0), np = (n + CB - 1) / CB;
  return 2 * n * ld + ne * ld + ni * ld + np * CB * CB + 7 * n + 5 * ne + 9 * nz + 64 + 32;
}

// group-wide reduction: two sums, one maximum, one minimum -> out[0..3] (visible to every thread after the call)
HD void qp_reduce(double s0, double s1, double mx, double mn, double *scratch, double *out) {
#ifdef MPC_HOST_EMU
  out[0] = s0; out[1] = s1; out[2] = mx; out[3] = mn;
#else
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s0 += __shfl_xor_sync(0xffffffffu, s0, o);
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
  }
  if (LANE0) { scratch[4 * WARP_ID] = s0; scratch[4 * WARP_ID + 1] = s1; scratch[4 * WARP_ID + 2] = mx; scratch[4 * WARP_ID + 3] = mn; }
  SYNC();
  ONE_THREAD {
    double a = 0, b = 0, c = scratch[2], d = scratch[3];
    for (int w = 0; w < NWARPS; w++) { a += scratch[4 * w]; b += scratch[4 * w + 1]; c = fmax(c, scratch[4 * w + 2]); d = fmin(d, scratch[4 * w + 3]); }
    out[0] = a; out[1] = b; out[2] = c; out[3] = d;
  }
#endif
  SYNC();
}

HD bool qp_inf(double v) { return fabs(v) >= 1e20; }

struct QPView { // carved shared memory of one QP
  int n, ne, ni, nz, ld, box;
  double *Hs, *Ks, *As, *Cs, *Dinv;
  double *x, *xe, *g, *grad, *dx, *hxg, *hd;  // n
  double *y, *ye, *re, *b, *Adx;              // ne
  double *z, *ze, *su, *sl, *cd, *lo, *up, *s, *t; // nz
  double *red, *sc;                            // 64 reduction scratch, 32 scalars
};
HD QPView qp_carve(double *m, int n, int ne, int ni, int box) {
  QPView v;
  v.n = n; v.ne = ne; v.ni = ni; v.box = box; v.nz = ni + (box ? n : 0); v.ld = qp_ld(n);
  const int np = (n + CB - 1) / CB;
  v.Hs = m; m += n * v.ld; v.Ks = m; m += n * v.ld; v.As = m; m += ne * v.ld; v.Cs = m; m += ni * v.ld; v.Dinv = m; m += np * CB * CB;
  v.x = m; m += n; v.xe = m; m += n; v.g = m; m += n; v.grad = m; m += n; v.dx = m; m += n; v.hxg = m; m += n; v.hd = m; m += n;
  v.y = m; m += ne; v.ye = m; m += ne; v.re = m; m += ne; v.b = m; m += ne; v.Adx = m; m += ne;
  v.z = m; m += v.nz; v.ze = m; m += v.nz; v.su = m; m += v.nz; v.sl = m; m += v.nz; v.cd = m; m += v.nz; v.lo = m; m += v.nz; v.up = m; m += v.nz;
  v.s = m; m += v.nz; v.t = m; m += v.nz;
  v.red = m; m += 64; v.sc = m;
  return v;
}
// scalar slots in v.sc
enum { QS_PRI = 0, QS_DUA, QS_GAP, QS_OBJ, QS_PRIS, QS_DUAS, QS_GAPS, QS_A0, QS_B0, QS_LO, QS_HI, QS_ALPHA, QS_R0, QS_R1, QS_R2, QS_R3 };

HD double qp_rowdot(const double *row, const double *v, int n) {
  double s = 0;
  for (int j = 0; j < n; j++) s += row[j] * v[j];
  return s;
}

// primal / dual residuals, duality gap and their scales at (x, y, z) -> v.sc[QS_PRI..QS_GAPS]  (oracle/qp.hpp qp_residuals)
HD void qp_residuals(const QPView &v) {
  const int n = v.n, ne = v.ne, nz = v.nz, ld = v.ld;
  double pri = 0, nAx = 0, nCx = 0, by = 0, bz = 0;
  PAR_FOR(w, ne + nz) {
    if (w < ne) {
      const double ax = qp_rowdot(v.As + w * ld, v.x, n), s = ax - v.b[w];
      pri = fmax(pri, fabs(s)); nAx = fmax(nAx, fabs(ax)); by += v.b[w] * v.y[w];
    } else {
      const int i = w - ne;
      const double s = (i < v.ni) ? qp_rowdot(v.Cs + i * ld, v.x, n) : v.x[i - v.ni];
      nCx = fmax(nCx, fabs(s));
      double viol = 0;
      if (!qp_inf(v.up[i])) viol += fmax(s - v.up[i], 0.0);
      if (!qp_inf(v.lo[i])) viol += fmin(s - v.lo[i], 0.0);
      pri = fmax(pri, fabs(viol));
      const double zi = v.z[i];
      if (zi > 0 && !qp_inf(v.up[i])) bz += v.up[i] * zi;
      if (zi < 0 && !qp_inf(v.lo[i])) bz += v.lo[i] * zi;
    }
  }
  qp_reduce(by, bz, pri, 0.0, v.red, v.sc + QS_R0);
  const double by_t = v.sc[QS_R0], bz_t = v.sc[QS_R1], pri_t = v.sc[QS_R2];
  SYNC();
  qp_reduce(0.0, 0.0, fmax(nAx, nCx), 0.0, v.red, v.sc + QS_R0);
  const double pris_t = v.sc[QS_R2];
  SYNC();
  double xHx = 0, gx = 0, dua = 0, m1 = 0, m2 = 0;
  PAR_FOR(i, n) {
    const double hx = qp_rowdot(v.Hs + i * ld, v.x, n);
    double aty = 0, ctz = 0;
    for (int r = 0; r < ne; r++) aty += v.As[r * ld + i] * v.y[r];
    for (int k = 0; k < v.ni; k++) ctz += v.Cs[k * ld + i] * v.z[k];
    if (v.box) ctz += v.z[v.ni + i];
    xHx += v.x[i] * hx; gx += v.g[i] * v.x[i];
    dua = fmax(dua, fabs(hx + v.g[i] + aty + ctz));
    m1 = fmax(m1, fmax(fabs(hx), fabs(v.g[i])));
    m2 = fmax(m2, fmax(fabs(aty), fabs(ctz)));
  }
  qp_reduce(xHx, gx, dua, 0.0, v.red, v.sc + QS_R0);
  const double xHx_t = v.sc[QS_R0], gx_t = v.sc[QS_R1], dua_t = v.sc[QS_R2];
  SYNC();
  qp_reduce(0.0, 0.0, fmax(m1, m2), 0.0, v.red, v.sc + QS_R0);
  const double duas_t = v.sc[QS_R2];
  SYNC();
  ONE_THREAD {
    v.sc[QS_PRI] = pri_t; v.sc[QS_DUA] = dua_t; v.sc[QS_GAP] = xHx_t + gx_t + by_t + bz_t; v.sc[QS_OBJ] = 0.5 * xHx_t + gx_t;
    v.sc[QS_PRIS] = pris_t; v.sc[QS_DUAS] = duas_t;
    v.sc[QS_GAPS] = fmax(fmax(fabs(xHx_t), fabs(gx_t)), fmax(fabs(by_t), fabs(bz_t)));
  }
  SYNC();
}

// primal residual only (the BCL test)
HD double qp_primal_residual(const QPView &v) {
  const int n = v.n, ne = v.ne, nz = v.nz, ld = v.ld;
  double pri = 0;
  PAR_FOR(w, ne + nz) {
    if (w < ne) pri = fmax(pri, fabs(qp_rowdot(v.As + w * ld, v.x, n) - v.b[w]));
    else {
      const int i = w - ne;
      const double s = (i < v.ni) ? qp_rowdot(v.Cs + i * ld, v.x, n) : v.x[i - v.ni];
      double viol = 0;
      if (!qp_inf(v.up[i])) viol += fmax(s - v.up[i], 0.0);
      if (!qp_inf(v.lo[i])) viol += fmin(s - v.lo[i], 0.0);
      pri = fmax(pri, fabs(viol));
    }
  }
  qp_reduce(0.0, 0.0, pri, 0.0, v.red, v.sc + QS_R0);
  const double r = v.sc[QS_R2];
  SYNC();
  return r;
}

// K = H + rho I + A'A / mu_e + C_act' C_act / mu_i (lower triangle incl. diagonal), 2 x 2 register tiles
HD void qp_newton_matrix(const QPView &v, double rho, double mue, double mui) {
  const int n = v.n, ne = v.ne, ld = v.ld, th = (n + 1) / 2;
  const double ime = 1.0 / mue, imi = 1.0 / mui;
  PAR_FOR(tile, th * th) {
    const int ti = tile / th, tj = tile % th;
    if (tj > ti) continue;
    const int i0 = 2 * ti, j0 = 2 * tj;
    const bool i1 = i0 + 1 < n, j1 = j0 + 1 < n;
    const int i1o = i1 ? 1 : 0, j1o = j1 ? 1 : 0;
    double a00 = 0, a01 = 0, a10 = 0, a11 = 0;
    for (int r = 0; r < ne; r++) {
      const double *ar = v.As + r * ld;
      const double x0 = ar[i0], x1 = ar[i0 + i1o], y0 = ar[j0], y1 = ar[j0 + j1o];
      a00 += x0 * y0; a01 += x0 * y1; a10 += x1 * y0; a11 += x1 * y1;
    }
    a00 *= ime; a01 *= ime; a10 *= ime; a11 *= ime;
    double c00 = 0, c01 = 0, c10 = 0, c11 = 0;
    for (int k = 0; k < v.ni; k++) {
      if (!(v.su[k] > 0 || v.sl[k] < 0)) continue;
      const double *cr = v.Cs + k * ld;
      const double x0 = cr[i0], x1 = cr[i0 + i1o], y0 = cr[j0], y1 = cr[j0 + j1o];
      c00 += x0 * y0; c01 += x0 * y1; c10 += x1 * y0; c11 += x1 * y1;
    }
    a00 += c00 * imi; a01 += c01 * imi; a10 += c10 * imi; a11 += c11 * imi;
    if (i0 == j0) {
      double d0 = rho, d1 = rho;
      if (v.box) {
        if (v.su[v.ni + i0] > 0 || v.sl[v.ni + i0] < 0) d0 += imi;
        if (i1 && (v.su[v.ni + i0 + 1] > 0 || v.sl[v.ni + i0 + 1] < 0)) d1 += imi;
      }
      a00 += d0; a11 += d1;
    }
    v.Ks[i0 * ld + j0] = v.Hs[i0 * ld + j0] + a00;
    if (j1 && j0 + 1 <= i0) v.Ks[i0 * ld + j0 + 1] = v.Hs[i0 * ld + j0 + 1] + a01;
    if (i1) {
      v.Ks[(i0 + 1) * ld + j0] = v.Hs[(i0 + 1) * ld + j0] + a10;
      if (j1) v.Ks[(i0 + 1) * ld + j0 + 1] = v.Hs[(i0 + 1) * ld + j0 + 1] + a11;
    }
  }
  SYNC();
}

// The whole solve of QP `inst` by the calling group.  smem: qp_smem_doubles(...) doubles of the group's shared memory.
HD void qp_solve_group(const QPArgs &P, int inst, double *smem) {
  const int n = P.n, ne = P.ne, ni = P.ni;
  const QPView v = qp_carve(smem, n, ne, ni, P.box);
  const int nz = v.nz, ld = v.ld;
  const mpc_qp_settings_t &st = P.st;
  // ---- load the QP
  {
    const double *H = P.H + inst * P.sH, *A = P.A + inst * P.sA, *Cm = P.C + inst * P.sC;
    PAR_FOR(e, n * n) v.Hs[(e / n) * ld + e % n] = H[e];
    PAR_FOR(e, ne * n) v.As[(e / n) * ld + e % n] = A[e];
    PAR_FOR(e, ni * n) v.Cs[(e / n) * ld + e % n] = Cm[e];
    const double *g = P.g + inst * P.sg, *b = P.b + inst * P.sb, *l = P.l + inst * P.sl, *u = P.u + inst * P.su;
    PAR_FOR(i, n) { v.g[i] = g[i]; v.x[i] = st.warm_start ? P.x[(long long)inst * n + i] : 0.0; }
    PAR_FOR(r, ne) { v.b[r] = b[r]; v.y[r] = st.warm_start ? P.y[(long long)inst * ne + r] : 0.0; }
    PAR_FOR(i, nz) {
      v.lo[i] = (i < ni) ? l[i] : P.lb[inst * P.slb + i - ni];
      v.up[i] = (i < ni) ? u[i] : P.ub[inst * P.sub + i - ni];
      v.z[i] = st.warm_start ? P.z[(long long)inst * nz + i] : 0.0;
    }
  }
  SYNC();
  double mue = st.mu_eq, mui = st.mu_in;
  const double eta_ext_init = pow(0.1, st.alpha_bcl), eps_in_min = fmin(st.eps_abs, 1e-9);
  double eta_ext = eta_ext_init, eta_in = 1.0;
  int status = 1, it = 0, it_in = 0, mu_updates = 0;
  for (;; it++) {
    qp_residuals(v);
    const double pri = v.sc[QS_PRI], dua = v.sc[QS_DUA], gap = v.sc[QS_GAP];
    if (!isfinite(pri) || !isfinite(dua)) { status = 2; break; }
    const bool ok = pri <= st.eps_abs + st.eps_rel * v.sc[QS_PRIS] && dua <= st.eps_abs + st.eps_rel * v.sc[QS_DUAS] &&
                    (!st.check_duality_gap || fabs(gap) <= st.eps_abs + st.eps_rel * v.sc[QS_GAPS]);
    if (ok) { status = 0; break; }
    if (it >= st.max_iter) break;
    PAR_FOR(i, n) v.xe[i] = v.x[i];
    PAR_FOR(r, ne) v.ye[r] = v.y[r];
    PAR_FOR(i, nz) v.ze[i] = v.z[i];
    SYNC();
    bool failed = false;
    for (int in = 0;; in++) {
      // constraint residuals of the augmented Lagrangian
      PAR_FOR(w, ne + nz) {
        if (w < ne) v.re[w] = qp_rowdot(v.As + w * ld, v.x, n) - v.b[w] + mue * v.ye[w];
        else {
          const int i = w - ne;
          const double s = (i < ni) ? qp_rowdot(v.Cs + i * ld, v.x, n) : v.x[i - ni];
          const double su = qp_inf(v.up[i]) ? -1e300 : s - v.up[i] + mui * v.ze[i];
          const double sl = qp_inf(v.lo[i]) ? 1e300 : s - v.lo[i] + mui * v.ze[i];
          v.su[i] = su; v.sl[i] = sl;
          v.t[i] = (fmax(su, 0.0) + fmin(sl, 0.0)) / mui;
        }
      }
      SYNC();
      PAR_FOR(i, n) {
        const double hxg = v.g[i] + st.rho * (v.x[i] - v.xe[i]) + qp_rowdot(v.Hs + i * ld, v.x, n);
        double s = 0, c = 0;
        for (int r = 0; r < ne; r++) s += v.As[r * ld + i] * v.re[r];
        for (int k = 0; k < ni; k++) c += v.Cs[k * ld + i] * v.t[k];
        if (P.box) c += v.t[ni + i];
        v.hxg[i] = hxg;
        v.grad[i] = hxg + s / mue + c;
      }
      SYNC();
      double gn = 0;
      for (int i = 0; i < n; i++) gn = fmax(gn, fabs(v.grad[i])); // every thread: uniform control flow without another barrier
      if (!isfinite(gn)) { failed = true; break; }
      if (gn <= eta_in || in >= st.max_iter_in) break;
      // Newton step on the current active set
      qp_newton_matrix(v, st.rho, mue, mui);
      PAR_FOR(i, n) v.dx[i] = -v.grad[i];
      chol_blocked(v.Ks, n, ld, v.Dinv);
      trsm_blocked(v.Ks, n, ld, v.Dinv, v.dx, 1, 1);
      it_in++;
      // exact linesearch on the piecewise quadratic
      PAR_FOR(w, ne + nz + n) {
        if (w < ne) v.Adx[w] = qp_rowdot(v.As + w * ld, v.dx, n);
        else if (w < ne + nz) { const int i = w - ne; v.cd[i] = (i < ni) ? qp_rowdot(v.Cs + i * ld, v.dx, n) : v.dx[i - ni]; }
        else { const int i = w - ne - nz; v.hd[i] = qp_rowdot(v.Hs + i * ld, v.dx, n); }
      }
      SYNC();
      double a0 = 0, b0 = 0;
      PAR_FOR(w, ne + n) {
        if (w < ne) { a0 += v.Adx[w] * v.Adx[w] / mue; b0 += v.Adx[w] * v.re[w] / mue; }
        else { const int i = w - ne; a0 += v.dx[i] * (v.hd[i] + st.rho * v.dx[i]); b0 += v.dx[i] * v.hxg[i]; }
      }
      qp_reduce(a0, b0, 0.0, 0.0, v.red, v.sc + QS_A0); // -> QS_A0, QS_B0 (QS_LO / QS_HI overwritten below)
      a0 = v.sc[QS_A0]; b0 = v.sc[QS_B0];
      SYNC();
      double lo = 0.0, hi = INFINITY;
      PAR_FOR(c, 2 * nz) {
        const int i = c >> 1;
        const double sv = (c & 1) ? v.sl[i] : v.su[i], cdi = v.cd[i];
        if (cdi == 0 || fabs(sv) >= 1e299) continue;
        const double t = -sv / cdi;
        if (!(t > 0)) continue;
        double d = b0 + a0 * t;
        for (int k = 0; k < nz; k++) d += v.cd[k] * (fmax(v.su[k] + t * v.cd[k], 0.0) + fmin(v.sl[k] + t * v.cd[k], 0.0)) / mui;
        if (d < 0) lo = fmax(lo, t); else hi = fmin(hi, t);
      }
      qp_reduce(0.0, 0.0, lo, hi, v.red, v.sc + QS_R0);
      lo = v.sc[QS_R2]; hi = v.sc[QS_R3];
      SYNC();
      const double tm = isfinite(hi) ? 0.5 * (lo + hi) : lo + 1.0;
      double slope = 0, icpt = 0;
      PAR_FOR(i, nz) {
        const double cdi = v.cd[i];
        if (v.su[i] + tm * cdi > 0) { slope += cdi * cdi / mui; icpt += cdi * v.su[i] / mui; }
        if (v.sl[i] + tm * cdi < 0) { slope += cdi * cdi / mui; icpt += cdi * v.sl[i] / mui; }
      }
      qp_reduce(slope, icpt, 0.0, 0.0, v.red, v.sc + QS_R0);
      double alpha = -(b0 + v.sc[QS_R1]) / (a0 + v.sc[QS_R0]);
      SYNC();
      alpha = fmin(fmax(alpha, lo), hi);
      if (!isfinite(alpha)) { failed = true; break; }
      double step = 0, xn = 1.0;
      for (int i = 0; i < n; i++) { step = fmax(step, fabs(alpha * v.dx[i])); xn = fmax(xn, fabs(v.x[i])); } // every thread (uniform)
      SYNC();
      PAR_FOR(i, n) v.x[i] += alpha * v.dx[i];
      SYNC();
      if (step <= 1e-14 * xn) break; // the Newton step is below the rounding level of x
    }
    if (failed) { status = 2; break; }
    // multiplier estimates at the inner solution, BCL test
    PAR_FOR(w, ne + nz) {
      if (w < ne) v.y[w] = v.ye[w] + (qp_rowdot(v.As + w * ld, v.x, n) - v.b[w]) / mue;
      else {
        const int i = w - ne;
        const double s = (i < ni) ? qp_rowdot(v.Cs + i * ld, v.x, n) : v.x[i - ni];
        const double zu = qp_inf(v.up[i]) ? 0.0 : fmax(v.ze[i] + (s - v.up[i]) / mui, 0.0);
        const double zl = qp_inf(v.lo[i]) ? 0.0 : fmin(v.ze[i] + (s - v.lo[i]) / mui, 0.0);
        v.z[i] = zu + zl;
      }
    }
    SYNC();
    const double pri_new = qp_primal_residual(v);
    if (pri_new <= eta_ext) {
      eta_ext *= pow(mui, st.beta_bcl);
      eta_in = fmax(eta_in * mui, eps_in_min);
    } else {
      PAR_FOR(r, ne) v.y[r] = v.ye[r];
      PAR_FOR(i, nz) v.z[i] = v.ze[i];
      SYNC();
      const double nmui = fmax(mui * st.mu_update_factor, st.mu_min_in), nmue = fmax(mue * st.mu_update_factor, st.mu_min_eq);
      if (nmui != mui || nmue != mue) mu_updates++;
      mui = nmui; mue = nmue;
      eta_ext = eta_ext_init * pow(mui, st.alpha_bcl);
      eta_in = fmax(mui, eps_in_min);
    }
  }
  PAR_FOR(i, n) P.x[(long long)inst * n + i] = v.x[i];
  PAR_FOR(r, ne) P.y[(long long)inst * ne + r] = v.y[r];
  PAR_FOR(i, nz) P.z[(long long)inst * nz + i] = v.z[i];
  if (P.info) {
    ONE_THREAD {
      mpc_qp_info_t &o = P.info[inst];
      o.status = status; o.iter = it; o.iter_in = it_in; o.mu_updates = mu_updates;
      o.pri_res = v.sc[QS_PRI]; o.dua_res = v.sc[QS_DUA]; o.duality_gap = v.sc[QS_GAP]; o.objective = v.sc[QS_OBJ];
    }
  }
}

// IDSolver_ulim.computeMatrice (QP_utils.py:514-552) for one instance, nv = 28, nk = 2, 6-D contacts: fills A [40][62], b [40],
// C [18][62], l [18] of the handle from M, nle, Jc (LOCAL contact Jacobians, [12][28]), gamma [12], a [28], forces [12], cs [2].
HD void qp_assemble_id_group(const double *M, const double *nle, const double *Jc, const double *gamma, const double *a, const double *forces,
                             const int32_t *cs, double mu, double L, double W, double *A, double *b, double *C, double *l) {
  constexpr int nv = 28, nk = 2, fs = 6, nf = nk * fs, n = 2 * nv - 6 + nf, ne = nv + nf;
  PAR_FOR(e, ne * n) {
    const int r = e / n, c = e % n;
    double val = 0;
    if (r < nv) {
      if (c < nv) val = M[r * nv + c];
      else if (c < nv + nf) val = cs[(c - nv) / fs] ? -Jc[(c - nv) * nv + r] : 0.0;
      else val = (r >= 6 && c - nv - nf == r - 6) ? -1.0 : 0.0;
    } else if (c < nv) val = cs[(r - nv) / fs] ? Jc[(r - nv) * nv + c] : 0.0;
    A[e] = val;
  }
  PAR_FOR(r, ne) {
    double s;
    if (r < nv) {
      s = -nle[r];
      for (int j = 0; j < nv; j++) s -= M[r * nv + j] * a[j];
      for (int k = 0; k < nf; k++) if (cs[k / fs]) s += Jc[k * nv + r] * forces[k];
    } else {
      const int k = r - nv;
      s = 0;
      if (cs[k / fs]) { s = -gamma[k]; for (int j = 0; j < nv; j++) s -= Jc[k * nv + j] * a[j]; }
    }
    b[r] = s;
  }
  PAR_FOR(e, 9 * nk * n) {
    const int r = e / n, c = e % n, i = r / 9, rr = r % 9, cc = c - nv - i * fs;
    double val = 0;
    if (cs[i] && cc >= 0 && cc < fs) {
      // Cmin as written at QP_utils.py:474-484 (rows 2 and 3 repeat the x rows there)
      if (rr < 4) val = (cc == 0) ? ((rr & 1) ? 1.0 : -1.0) : (cc == 2 ? mu : 0.0);
      else if (rr == 4) val = (cc == 2) ? 1.0 : 0.0;
      else if (rr < 7) val = (cc == 2) ? W : (cc == 3 ? ((rr == 5) ? -1.0 : 1.0) : 0.0);
      else val = (cc == 2) ? L : (cc == 4 ? ((rr == 7) ? -1.0 : 1.0) : 0.0);
    }
    C[e] = val;
  }
  PAR_FOR(r, 9 * nk) {
    const int i = r / 9, rr = r % 9;
    const double *f = forces + i * fs;
    double val = 0;
    if (cs[i]) {
      switch (rr) { // QP_utils.py:538-548
      case 0: val = f[0] - f[2] * mu; break;
      case 1: val = -f[0] - f[2] * mu; break;
      case 2: val = f[1] - f[2] * mu; break;
      case 3: val = -f[1] - f[2] * mu; break;
      case 4: val = -f[2]; break;
      case 5: val = f[3] - f[2] * W; break;
      case 6: val = -f[3] - f[2] * W; break;
      case 7: val = f[4] - f[2] * L; break;
      default: val = -f[4] - f[2] * L; break;
      }
    }
    l[r] = val;
  }
}

} // namespace mpcdev
