// Proximal Riccati, tensor-core variant for state dimensions that are multiples of 8 (full / kinodynamic: n = 56).
// Same recursion and outputs as riccati_instance (riccati.cuh); what changes is HOW the three dense
// contractions per knot are executed:
//   Lambda^-1 = Linv' Linv,   Pt = Lambda^-1 P,   W = Pt [A B],   H = H_k + [A B]' W
// all run on the FP64 DMMA pipe out of shared memory (dmma.cuh), and the triangular solve with Lambda is replaced by an
// explicit inverse of its Cholesky factor built by independent per-warp column chains (no block barriers).
// Leading dimensions are padded to 8 (mod 16) doubles so the DMMA fragment loads are bank-conflict free.
#pragma once
#include "dmma.cuh"
#include "riccati.cuh"

namespace mpcdev {

// optional per-phase cycle counters (thread 0 of the CTA handling instance 0), enabled with -DMPC_PHASE_TIMING
#if defined(MPC_PHASE_TIMING) && !defined(MPC_HOST_EMU)
#define PHASE_DECL long long ph_last = clock64(); long long ph_acc[16] = {0}
#define PHASE(i) do { if (threadIdx.x == 0) { long long t_ = clock64(); ph_acc[i] += t_ - ph_last; ph_last = t_; } } while (0)
#define PHASE_DUMP(ptr) do { if (threadIdx.x == 0 && (ptr)) for (int i_ = 0; i_ < 16; i_++) (ptr)[i_] = (double)ph_acc[i_]; } while (0)
#else
#define PHASE_DECL
#define PHASE(i)
#define PHASE_DUMP(ptr)
#endif

template <int N, int M, int NC, int NCAP = NC> struct RicFastLayout {
  static constexpr int NZ = N + M, NR = 1 + N, S = M + NC;
  static constexpr int NBLK = N / 8;
  static constexpr int ZP = (NZ + 7) / 8 * 8;                       // padded n+m
  static constexpr int LDN = (N % 16 == 8) ? N : N + 8;             // ld of N x N buffers
  static constexpr int LDZ = (ZP % 16 == 8) ? ZP : ZP + 8;          // ld of N x ZP buffers (AB, W)
  static constexpr int LDH = ZP;                                    // ld of the Hessian buffer
  static_assert(N * LDZ <= 2 * N * LDN, "W must fit over the dead [P | G] buffers");
  static constexpr int phase1 = 3 * N * LDN + N * LDZ;  // P, G (later reused as W), Li, AB
  static constexpr int phase2 = NCAP * NZ + NCAP * NCAP + M * (NR + NCAP) + NCAP * NR + M * M; // sized for NCAP active rows
  static constexpr int un = phase1 > phase2 ? phase1 : phase2;
  static constexpr int vecs = 8 * ZP + 36 + 2 * NC + 256 + 64 * ((NC + 7) / 8) + 8 * 64 + 16;
  static constexpr int total = ZP * LDH + un + vecs;
};

template <int N, int M, int NC, int NCAP = NC> HD void riccati_instance_fast(const RiccatiIO &io, double *ws) {
  using Lay = RicFastLayout<N, M, NC, NCAP>;
  constexpr int NZ = Lay::NZ, NR = Lay::NR, S = Lay::S, ZP = Lay::ZP, LDN = Lay::LDN, LDZ = Lay::LDZ, LDH = Lay::LDH, NBLK = Lay::NBLK;
  static_assert(N % 8 == 0, "fast Riccati needs n % 8 == 0");
  const int T = io.T;
  const double mu = io.mu, mu_d = io.mu_d;
  PHASE_DECL;
  // ---- carve shared memory
  double *H = ws;                              // ZP x LDH, zero padded; [0:N,0:N] carries the value-function Hessian between knots
  double *U0 = H + ZP * LDH;
  double *P = U0, *G = P + N * LDN, *Li = G + N * LDN, *AB = Li + N * LDN, *W = P;  // phase 1 (W overwrites the dead P, G)
  double *CD = U0, *Sg = CD + NCAP * NZ, *Z = Sg + NCAP * NCAP, *Kv = Z + M * (NR + NCAP), *Rh = Kv + NCAP * NR;  // phase 2
  double *vec = U0 + Lay::un;
  double *p = vec, *pt = p + ZP, *gh = pt + ZP, *fb = gh + ZP, *tmp = fb + ZP, *dx = tmp + ZP, *z = dx + ZP, *pv = z + ZP;
  double *T6 = pv + ZP, *dbr = T6 + 36, *dva = dbr + NC, *red = dva + NC, *dinv = red + 256, *wtmp = dinv + 64 * ((NC + 7) / 8);
  // ---- zero the padding of H once; terminal value function: P = H_T[xx] + C'C/mu, p = g_T[x] + C' d/mu
  PAR_FOR(e, ZP * LDH) H[e] = 0.0;
  SYNC();
  {
    const double *HT = io.H + (size_t)T * NZ * NZ, *gT = io.g + (size_t)T * NZ;
    const int nca = io.nca[T];
    const double *CT = io.CDact + (size_t)T * NC * NZ, *dT = io.dbar + (size_t)T * NC;
    const int32_t *ai = io.act_idx + (size_t)T * NC;
    PAR_FOR(e, N * N) {
      int i = e / N, j = e % N;
      double s = HT[i * NZ + j];
      for (int r = 0; r < nca; r++) s += CT[r * NZ + i] * CT[r * NZ + j] / mu;
      H[i * LDH + j] = s;
    }
    PAR_FOR(i, N) {
      double s = gT[i];
      for (int r = 0; r < nca; r++) s += CT[r * NZ + i] * dT[ai[r]] / mu;
      p[i] = s;
    }
    SYNC();
  }
  for (int k = T - 1; k >= 0; k--) {
    const double *gAB = io.AB + (size_t)k * N * NZ, *gH = io.H + (size_t)k * NZ * NZ;
    int nca = io.nca[k];
    if (nca > NCAP) { nca = NCAP; ONE_THREAD { if (io.overflow) *io.overflow = 1; } } // more active rows than the shared-memory KKT holds
    // 1. stage [A B] (zero-padded columns), P <- value Hessian, E normalisation P <- T' P T, p <- T' p
    PAR_FOR(e, N * (ZP / 2)) {
      int i = e / (ZP / 2), j = (e % (ZP / 2)) * 2;
      double2 v = make_double2(0.0, 0.0);
      if (j < NZ) v = *reinterpret_cast<const double2 *>(gAB + i * NZ + j); // NZ even, rows 16-byte aligned
      *reinterpret_cast<double2 *>(AB + i * LDZ + j) = v;
    }
    PAR_FOR(e, N * N) { int i = e / N, j = e % N; P[i * LDN + j] = H[i * LDH + j]; }
    PAR_FOR(e, 36) T6[e] = io.T6[(size_t)k * 36 + e];
    PAR_FOR(i, N) fb[i] = io.fbar[(size_t)k * N + i];
    SYNC();
    PAR_FOR(e, N * 6) { int i = e / 6, j = e % 6; double s = 0; for (int l = 0; l < 6; l++) s += P[i * LDN + l] * T6[6 * l + j]; Li[e] = s; }
    SYNC();
    PAR_FOR(e, N * 6) { int i = e / 6, j = e % 6; P[i * LDN + j] = Li[e]; }
    SYNC();
    PAR_FOR(e, 6 * N) { int i = e / N, j = e % N; double s = 0; for (int l = 0; l < 6; l++) s += T6[6 * l + i] * P[l * LDN + j]; Li[e] = s; }
    PAR_FOR(i, 6) { double s = 0; for (int l = 0; l < 6; l++) s += T6[6 * l + i] * p[l]; tmp[i] = s; }
    SYNC();
    PAR_FOR(e, 6 * N) { int i = e / N, j = e % N; P[i * LDN + j] = Li[e]; }
    PAR_FOR(i, 6) p[i] = tmp[i];
    SYNC();
    PAR_FOR(e, N * LDN) Li[e] = 0.0;
    PHASE(0);
    // 2. G = chol(I + mu_d P);  pv = p + P f
    PAR_FOR(e, N * N) { int i = e / N, j = e % N; G[i * LDN + j] = mu_d * P[i * LDN + j] + ((i == j) ? 1.0 : 0.0); }
    PAR_FOR(i, N) { double s = p[i]; for (int j = 0; j < N; j++) s += P[i * LDN + j] * fb[j]; pv[i] = s; }
    SYNC();
    PHASE(1);
    chol_blocked(G, N, LDN, dinv);
    PHASE(2);
    // 3. Linv = G^-1 (lower triangular): independent forward-substitution chains, one per 8-column block and warp
    {
#ifdef MPC_HOST_EMU
      for (int jb = 0; jb < NBLK; jb++) {
        double *tt = wtmp;
#else
      for (int jb = (threadIdx.x >> 5); jb < NBLK; jb += (blockDim.x >> 5)) {
        double *tt = wtmp + 64 * (threadIdx.x >> 5);
#endif
        for (int ib = jb; ib < NBLK; ib++) {
          WARP_FOR(e, 64) { // tt = E_ij - sum_k L_ik X_kj
            int r = e >> 3, c = e & 7;
            double s = (ib == jb && r == c) ? 1.0 : 0.0;
            for (int kb = jb; kb < ib; kb++)
              for (int q = 0; q < 8; q++) s -= G[(8 * ib + r) * LDN + 8 * kb + q] * Li[(8 * kb + q) * LDN + 8 * jb + c];
            tt[e] = s;
          }
          WARP_SYNC();
          WARP_FOR(e, 64) { // X_ij = Dinv_i tt
            int r = e >> 3, c = e & 7;
            const double *Di = dinv + 64 * ib;
            double s = 0;
            for (int q = 0; q <= r; q++) s += Di[r * 8 + q] * tt[q * 8 + c];
            Li[(8 * ib + r) * LDN + 8 * jb + c] = s;
          }
          WARP_SYNC();
        }
      }
      SYNC();
    }
    PHASE(3);
    // 4. Lambda^-1 = Linv' Linv (into G);  5. Pt = Lambda^-1 P (into Li), pt = Lambda^-1 pv
    mma_tn(NBLK, NBLK, N, Li, LDN, Li, LDN, G, LDN, nullptr, 0, 0, 0, false);
    PAR_FOR(i, N) { double s = 0; for (int j = 0; j < N; j++) s += G[i * LDN + j] * pv[j]; pt[i] = s; }
    mma_tn(NBLK, NBLK, N, G, LDN, P, LDN, Li, LDN, nullptr, 0, 0, 0, false);
    PHASE(4);
    // 6. W = Pt [A B];  7. H = H_k + [A B]' W;  gh = g + [A B]' pt
    mma_tn(NBLK, ZP / 8, N, Li, LDN, AB, LDZ, W, LDZ, nullptr, 0, 0, 0, false);
    mma_tn(ZP / 8, ZP / 8, N, AB, LDZ, W, LDZ, H, LDH, gH, NZ, NZ, NZ, false);
    PAR_FOR(i, NZ) { double s = io.g[(size_t)k * NZ + i]; for (int l = 0; l < N; l++) s += AB[l * LDZ + i] * pt[l]; gh[i] = s; }
    PAR_FOR(e, N * (NZ / 2)) {
      int i = e / (NZ / 2), j = (e % (NZ / 2)) * 2;
      *reinterpret_cast<double2 *>(io.W + (size_t)k * N * NZ + i * NZ + j) = *reinterpret_cast<const double2 *>(W + i * LDZ + j);
    }
    PAR_FOR(i, N) io.pt[(size_t)k * N + i] = pt[i];
    SYNC();
    PHASE(5);
    // 8. KKT by block elimination (phase-2 buffers alias P/G/Li/AB/W, all dead now)
    const int ncol = NR + nca; // Z columns: [rh | Sh' | D']
    const double *gCD = io.CDact + (size_t)k * NC * NZ;
    const int32_t *ai = io.act_idx + (size_t)k * NC;
    PAR_FOR(e, nca * NZ) CD[e] = gCD[e];
    PAR_FOR(r, nca) dbr[r] = io.dbar[(size_t)k * NC + ai[r]];
    PAR_FOR(e, M * M) { int i = e / M, j = e % M; Rh[e] = 0.5 * (H[(N + i) * LDH + N + j] + H[(N + j) * LDH + N + i]); }
    SYNC();
    PAR_FOR(e, M * ncol) {
      int i = e / ncol, c = e % ncol;
      Z[i * ncol + c] = (c == 0) ? gh[N + i] : (c < NR ? H[(c - 1) * LDH + N + i] : CD[(c - NR) * NZ + N + i]);
    }
    SYNC();
    PHASE(6);
    chol_blocked(Rh, M, M, dinv);
    PHASE(7);
    trsm_blocked(Rh, M, M, dinv, Z, ncol, ncol);
    PHASE(8);
    PAR_FOR(e, nca * nca) {
      int r = e / nca, c = e % nca;
      double s = (r == c) ? mu : 0.0;
      for (int l = 0; l < M; l++) s += CD[r * NZ + N + l] * Z[l * ncol + NR + c];
      Sg[e] = s;
    }
    PAR_FOR(e, nca * NR) {
      int r = e / NR, c = e % NR;
      double s = (c == 0) ? dbr[r] : CD[r * NZ + c - 1];
      for (int l = 0; l < M; l++) s -= CD[r * NZ + N + l] * Z[l * ncol + c];
      Kv[e] = s;
    }
    SYNC();
    PHASE(9);
    if (nca > 0) { chol_blocked(Sg, nca, nca, dinv); trsm_blocked(Sg, nca, nca, dinv, Kv, NR, NR); }
    PHASE(10);
    PAR_FOR(e, M * NR) {
      int i = e / NR, c = e % NR;
      double s = -Z[i * ncol + c];
      for (int r = 0; r < nca; r++) s -= Z[i * ncol + NR + r] * Kv[r * NR + c];
      Z[i * ncol + c] = s;
    }
    SYNC();
    double *gK = io.K + (size_t)k * S * NR;
    PAR_FOR(e, M * NR) { int i = e / NR, c = e % NR; double v = Z[i * ncol + c]; gK[e] = v; if (c > 0) io.Kfb[((size_t)k * M + i) * N + c - 1] = v; }
    PAR_FOR(e, nca * NR) gK[M * NR + e] = Kv[e];
    PHASE(11);
    // 9. P = Qh + Sh Ku + C' Kv, p = qh + Sh ku + C' kv : in place in H[0:N,0:N] / p, then symmetrise
    { // 2 x 4 register tiles over (row i, column c of [p | P]); column 0 is the vector p
      constexpr int TC = (NR + 3) / 4;
      PAR_FOR(t, (N / 2) * TC) {
        const int i0 = (t / TC) * 2, c0 = (t % TC) * 4;
        double a[2][4];
#pragma unroll
        for (int r = 0; r < 2; r++)
#pragma unroll
          for (int q = 0; q < 4; q++) { int c = c0 + q; a[r][q] = (c >= NR) ? 0.0 : ((c == 0) ? gh[i0 + r] : H[(i0 + r) * LDH + c - 1]); }
        for (int l = 0; l < M; l++) {
          const double h0 = H[i0 * LDH + N + l], h1 = H[(i0 + 1) * LDH + N + l];
          const double *zr = Z + l * ncol + c0;
#pragma unroll
          for (int q = 0; q < 4; q++) { double zv = (c0 + q < NR) ? zr[q] : 0.0; a[0][q] += h0 * zv; a[1][q] += h1 * zv; }
        }
        for (int r = 0; r < nca; r++) {
          const double h0 = CD[r * NZ + i0], h1 = CD[r * NZ + i0 + 1];
          const double *kr = Kv + r * NR + c0;
#pragma unroll
          for (int q = 0; q < 4; q++) { double kv = (c0 + q < NR) ? kr[q] : 0.0; a[0][q] += h0 * kv; a[1][q] += h1 * kv; }
        }
#pragma unroll
        for (int r = 0; r < 2; r++)
#pragma unroll
          for (int q = 0; q < 4; q++) { int c = c0 + q; if (c == 0) p[i0 + r] = a[r][q]; else if (c < NR) H[(i0 + r) * LDH + c - 1] = a[r][q]; }
      }
    }
    SYNC();
    PAR_FOR(e, N * N) {
      int i = e / N, j = e % N;
      if (i < j) { double s = 0.5 * (H[i * LDH + j] + H[j * LDH + i]); H[i * LDH + j] = s; H[j * LDH + i] = s; }
    }
    SYNC();
    PHASE(12);
  }
  // ---- forward sweep, dx0 = 0 (force_initial_condition, fulldynamic_talos.py:384)
  PHASE(13);
  double acc = 0.0;
  PAR_FOR(i, N) { dx[i] = 0.0; io.dxs[i] = 0.0; io.dlams[i] = -p[i]; }
  SYNC();
  for (int k = 0; k <= T; k++) {
    const int nca = io.nca[k];
    const int32_t *ai = io.act_idx + (size_t)k * NC;
    const double *gdb = io.dbar + (size_t)k * NC, *gvp = io.vplus + (size_t)k * NC, *gv = io.v + (size_t)k * NC;
    double *gdv = io.dvs + (size_t)k * NC;
    if (k < T) {
      const double *gK = io.K + (size_t)k * S * NR;
      PAR_FOR(i, M + nca) {
        double s = gK[i * NR];
        for (int j = 0; j < N; j++) s += gK[i * NR + 1 + j] * dx[j];
        if (i < M) { z[N + i] = s; io.dus[(size_t)k * M + i] = s; } else dva[i - M] = s;
      }
      PAR_FOR(i, N) z[i] = dx[i];
      PAR_FOR(e, N * (NZ / 2)) {
        int i = e / (NZ / 2), j = (e % (NZ / 2)) * 2;
        *reinterpret_cast<double2 *>(AB + i * LDZ + j) = *reinterpret_cast<const double2 *>(io.AB + (size_t)k * N * NZ + i * NZ + j);
        *reinterpret_cast<double2 *>(W + i * LDZ + j) = *reinterpret_cast<const double2 *>(io.W + (size_t)k * N * NZ + i * NZ + j);
      }
      PAR_FOR(e, 36) T6[e] = io.T6[(size_t)k * 36 + e];
    } else {
      const double *CT = io.CDact + (size_t)T * NC * NZ;
      PAR_FOR(r, nca) { double s = gdb[ai[r]]; for (int j = 0; j < N; j++) s += CT[r * NZ + j] * dx[j]; dva[r] = s / mu; }
      PAR_FOR(i, N) z[i] = dx[i];
    }
    PAR_FOR(r, NC) gdv[r] = gdb[r] / mu;
    SYNC();
    PAR_FOR(r, nca) gdv[ai[r]] = dva[r];
    PAR_FOR(i, (k < T ? NZ : N)) acc += io.lxu[(size_t)k * NZ + i] * z[i];
    PAR_FOR(r, nca) { int row = ai[r]; acc += (2.0 * gvp[row] - gv[row]) * (mu * dva[r] - gdb[row]) - gdb[row] * dva[r]; }
    PAR_FOR(r, NC) {
      bool active = false;
      for (int q = 0; q < nca; q++) active |= (ai[q] == r);
      if (!active) acc -= gdb[r] * gdb[r] / mu;
    }
    if (k == T) break;
    PAR_FOR(i, N) {
      double s = io.pt[(size_t)k * N + i], a = io.fbar[(size_t)k * N + i];
      for (int j = 0; j < NZ; j++) { s += W[i * LDZ + j] * z[j]; a += AB[i * LDZ + j] * z[j]; }
      tmp[i] = a - mu_d * s;
      io.dlams[(size_t)(k + 1) * N + i] = s;
      double fbi = io.fbar[(size_t)k * N + i], lp = io.lplus[(size_t)(k + 1) * N + i], lm = io.lam[(size_t)(k + 1) * N + i];
      acc += (2.0 * lp - lm) * (mu_d * s - fbi) - fbi * s;
    }
    SYNC();
    PAR_FOR(i, N) {
      double s;
      if (i < 6) { s = 0; for (int l = 0; l < 6; l++) s += T6[6 * i + l] * tmp[l]; } else s = tmp[i];
      dx[i] = s; io.dxs[(size_t)(k + 1) * N + i] = s;
    }
    SYNC();
  }
  PHASE(14);
  PHASE_DUMP(io.phase_out);
  red[TID] = acc;
  SYNC();
  ONE_THREAD { double s = 0; for (int t = 0; t < NTHREADS; t++) s += red[t]; io.dphi[0] = s; }
  SYNC();
}

} // namespace mpcdev
