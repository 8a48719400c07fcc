// Proximal Riccati, tensor-core variant for state dimensions that are multiples of 8 (full / kinodynamic: n = 56).
// Same recursion and outputs as riccati_instance (riccati.cuh); what changes is HOW the three dense
// contractions per knot are executed:
//   Lambda^-1 = Linv' Linv,   Pt = Lambda^-1 P,   W = Pt [A B],   H = H_k + [A B]' W
// all run on the FP64 DMMA pipe out of shared memory (dmma.cuh), and the triangular solve with Lambda is replaced by an
// explicit inverse of its Cholesky factor built by independent per-warp column chains (no block barriers).
// Leading dimensions are padded to 8 (mod 16) doubles so the DMMA fragment loads are bank-conflict free.
// Around them: both Cholesky factorisations (Lambda, padded control block) are tensor-core blocked with a one-panel
// look-ahead (chol_mma), the control-block solve uses an in-place inverse of its factor and tile-wise triangular products,
// the value update is a DMMA product, and all operands arrive by cp.async.bulk copies tracked by mbarriers one knot ahead
// (backward pass: [A B], T6, fbar, H_k; forward sweep: W, [A B], gain rows, small vectors, double-buffered).
#pragma once
#include "dmma.cuh"
#include "riccati.cuh"

namespace mpcdev {

// optional per-phase cycle counters (thread 0 of the CTA handling instance 0), enabled with -DMPC_PHASE_TIMING
#if defined(MPC_PHASE_TIMING) && !defined(MPC_HOST_EMU)
#define PHASE_DECL long long ph_last = clock64(); long long ph_acc[32] = {0}
#define PHASE(i) do { SYNC(); if (threadIdx.x == 0) { long long t_ = clock64(); ph_acc[i] += t_ - ph_last; ph_last = t_; } } while (0)
#define PHASE_DUMP(ptr) do { if (threadIdx.x == 0 && (ptr)) for (int i_ = 0; i_ < 16; i_++) { (ptr)[i_] = (double)ph_acc[i_]; (ptr)[48 + i_] = (double)ph_acc[16 + i_]; } } while (0)
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
  static_assert(N * LDZ <= 2 * N * LDN - 8 * N, "W must fit behind P, next to the [pv | 0] tile");
  static_assert(ZP > NZ, "the Hessian update needs a spare padding column for pt");
  static constexpr int phase1 = 3 * N * LDN + N * LDZ;  // P, G + scratch (later reused as W), AB
  static constexpr int MR = (M + 7) / 8 * 8;                        // control block padded to whole 8 x 8 tiles (identity / zero padding)
  static constexpr int LDR = (MR % 16 == 8) ? MR : MR + 8;          // ld of the padded R^ buffer
  static constexpr int LDZMAX = (NR + NCAP + 7) / 8 * 8;            // ld of Z for NCAP active rows (runtime ld: round8(NR + nca))
  static constexpr int phase2 = NCAP * NZ + NCAP * NCAP + MR * LDZMAX + NCAP * NR + MR * LDR; // sized for NCAP active rows
  static constexpr int un = phase1 > phase2 ? phase1 : phase2;
  // phase-2 order [Z | Kv | Rh | CD | Sg]; [A B] sits at the END of the union so that knots with few active rows never
  // touch it in phase 2 and the next knot's [A B] can be prefetched early
  static constexpr int offKv = MR * LDZMAX, offRh = offKv + NCAP * NR, offCD = offRh + MR * LDR, offSg = offCD + NCAP * NZ;
  static constexpr int offAB = un - N * LDZ;
  static_assert(offAB >= 3 * N * LDN, "[A B] overlaps P / G / Li");
  // knots with MORE than NCAP active rows (possible only when NCAP < NC: a kinodynamic iterate with nearly every cone / box row
  // violated) use a second carving sized for NC rows in which the compacted rows [C D] stay in global memory: [Z | Kv | Rh | Sg]
  static constexpr int LDZBIG = (NR + NC + 7) / 8 * 8;
  static constexpr int offKvB = MR * LDZBIG, offRhB = offKvB + NC * NR, offSgB = offRhB + MR * LDR;
  static_assert(offSgB + NC * NC <= un, "overflow carving of the KKT buffers does not fit");
  static_assert(un % 2 == 0 && offAB % 2 == 0 && (ZP * LDH) % 2 == 0, "16-byte alignment of the bulk-copy destinations");
  static constexpr int vecs = 8 * ZP + 36 + 2 * NC + 64 * ((NC + 7) / 8) + 8 * 64 + 16; // the final reduction reuses wtmp
  static constexpr int total = ZP * LDH + un + vecs;
};

template <int N, int M, int NC, int NCAP = NC> HD void riccati_instance_fast(const RiccatiIO &io, double *ws) {
  using Lay = RicFastLayout<N, M, NC, NCAP>;
  constexpr int NZ = Lay::NZ, NR = Lay::NR, S = Lay::S, ZP = Lay::ZP, LDN = Lay::LDN, LDZ = Lay::LDZ, LDH = Lay::LDH, NBLK = Lay::NBLK, MR = Lay::MR, LDR = Lay::LDR;
  static_assert(N % 8 == 0, "fast Riccati needs n % 8 == 0");
  const int T = io.T;
  const double mu = io.mu, mu_d = io.mu_d;
  PHASE_DECL;
  // ---- carve shared memory
  double *H = ws;                              // ZP x LDH, zero padded; [0:N,0:N] carries the value-function Hessian between knots
  double *U0 = H + ZP * LDH;
  double *P = U0, *G = P + N * LDN, *AB = U0 + Lay::offAB, *W = G, *PV = G + 2 * N * LDN - 8 * N;  // phase 1 (W overwrites the dead factor G; PV: [pv | 0] tile)
  double *Z = U0, *CDs = U0 + Lay::offCD;  // phase 2 (Kv, Rh, Sg: carved per knot, see `big`)
  double *vec = U0 + Lay::un;
  double *p = vec, *pt = p + ZP, *gh = pt + ZP, *fb = gh + ZP, *tmp = fb + ZP, *dx = tmp + ZP, *z = dx + ZP, *pv = z + ZP;
  double *T6 = pv + ZP, *dbr = T6 + 36, *dva = dbr + NC, *dinv = dva + NC, *wtmp = dinv + 64 * ((NC + 7) / 8), *red = wtmp;  // (red: one slot per thread, <= 512 threads)
  unsigned long long *mbar = reinterpret_cast<unsigned long long *>(vec + Lay::vecs - 8); // four mbarriers in the spare tail of vecs
  ONE_THREAD { MBAR_INIT(mbar, 1); MBAR_INIT(mbar + 1, 1); MBAR_INIT(mbar + 2, 1); MBAR_INIT(mbar + 3, 1); }
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
  // [A B] of a knot goes global -> shared by bulk copies (cp.async.bulk, one 624-byte row per issuing thread into the padded
  // rows; completion counted on an mbarrier), issued one knot ahead so that the copy overlaps the tail of the previous knot;
  // the padding columns NZ..ZP-1 are zeroed with plain stores.  mbar[0..1]: forward-sweep stages, [2]: [A B] + T6 + fbar, [3]: H_k
  static_assert((NZ * 8) % 16 == 0 && (LDZ * 8) % 16 == 0 && (LDH * 8) % 16 == 0 && (N * 8) % 16 == 0, "bulk copies need 16-byte rows");
  auto stage_AB_async = [&](int kk) {
    const double *src = io.AB + (size_t)kk * N * NZ;
    ONE_THREAD {
      FENCE_PROXY_ASYNC();
      MBAR_EXPECT_TX(mbar + 2, (N * NZ + 36 + N) * 8);
      BULK_G2S(T6, io.T6 + (size_t)kk * 36, 36 * 8, mbar + 2);
      BULK_G2S(fb, io.fbar + (size_t)kk * N, N * 8, mbar + 2);
    }
    PAR_FOR(i, N) { FENCE_PROXY_ASYNC(); BULK_G2S(AB + i * LDZ, src + i * NZ, NZ * 8, mbar + 2); }
    PAR_FOR(e, N * (ZP - NZ)) { int i = e / (ZP - NZ), j = NZ + e % (ZP - NZ); AB[i * LDZ + j] = 0.0; }
  };
  if (T > 0) stage_AB_async(T - 1);
  for (int k = T - 1; k >= 0; k--) {
    const double *gH = io.H + (size_t)k * NZ * NZ;
    const int nca = io.nca[k];
    const bool big = (NCAP < NC) && nca > NCAP; // more active rows than the fast carving holds: [C D] is read from global memory
    // 1. ONE pass over the value Hessian left in H by the previous knot: P <- T' sym(H) T (E normalisation: T = blockdiag(T6, I)
    //    touches the 6 base rows / columns only), G <- I + mu_d P, and the normalised gradient tmp <- T' p
    MBAR_WAIT(mbar + 2, (T - 1 - k) & 1); // [A B], T6 and fbar of this knot (issued one knot ahead)
    PHASE(21);
    PAR_FOR(e, N * N) { // lanes: 8 consecutive j x 4 consecutive i, which keeps the transposed read at 8-way bank conflicts
      const int blk = e >> 5, l = e & 31, i = (blk / (N / 8)) * 4 + (l >> 3), j = (blk % (N / 8)) * 8 + (l & 7);
      double v;
      if (i >= 6 && j >= 6) v = 0.5 * (H[i * LDH + j] + H[j * LDH + i]);
      else if (i >= 6) { v = 0; for (int q = 0; q < 6; q++) v += 0.5 * (H[i * LDH + q] + H[q * LDH + i]) * T6[6 * q + j]; }
      else if (j >= 6) { v = 0; for (int q = 0; q < 6; q++) v += T6[6 * q + i] * (0.5 * (H[q * LDH + j] + H[j * LDH + q])); }
      else {
        v = 0;
        for (int q = 0; q < 6; q++) {
          double r = 0;
          for (int m = 0; m < 6; m++) r += 0.5 * (H[q * LDH + m] + H[m * LDH + q]) * T6[6 * m + j];
          v += T6[6 * q + i] * r;
        }
      }
      P[i * LDN + j] = v;
      G[i * LDN + j] = mu_d * v + ((i == j) ? 1.0 : 0.0);
    }
    PAR_FOR(i, N) { double v = p[i]; if (i < 6) { v = 0; for (int q = 0; q < 6; q++) v += T6[6 * q + i] * p[q]; } tmp[i] = v; }
    PAR_FOR(e, N * 8) PV[e] = 0.0;
    SYNC();
    PHASE(16);
    // H is dead now: H <- H_k asynchronously (lands before the Hessian update needs it);  pv = p + P f
    ONE_THREAD MBAR_EXPECT_TX(mbar + 3, NZ * NZ * 8);
    PAR_FOR(i, NZ) { FENCE_PROXY_ASYNC(); BULK_G2S(H + i * LDH, gH + i * NZ, NZ * 8, mbar + 3); }
    matvec_rows(P, LDN, N, N, fb, tmp, pv);
    SYNC();
    PAR_FOR(i, N) PV[8 * i] = pv[i];
    PHASE(1);
    // 2. G = chol(I + mu_d P)
    chol_mma<NBLK>(G, LDN, dinv);
    PHASE(2);
    // 3. [P | pv] <- Lambda^-1 [P | pv] in place: blocked forward / backward substitution, one warp per 8-column tile
    trsm_mma<NBLK>(G, LDN, dinv, P, LDN, NBLK, PV, 8, NBLK + 1);
    PHASE(4);
    // 4. W = Pt [A B] (into shared memory over the dead factor AND to HBM for the forward sweep, straight from the accumulators)
    mma_tn(NBLK, ZP / 8, N, P, LDN, AB, LDZ, W, LDZ, nullptr, 0, 0, 0, false, false, io.W + (size_t)k * N * NZ, NZ, NZ);
    PHASE(17);
    // pt rides in the first padding column of W, so [A B]' pt falls out of the Hessian update (column NZ of H)
    PAR_FOR(i, N) { const double v = PV[8 * i]; pt[i] = v; W[i * LDZ + NZ] = v; io.pt[(size_t)k * N + i] = v; }
    MBAR_WAIT(mbar + 3, (T - 1 - k) & 1); // H_k
    SYNC();
    // 5. H = H_k + [A B]' W in place (rows / columns >= NZ of H_k read as zero);  gh = g + [A B]' pt
    mma_tn(ZP / 8, ZP / 8, N, AB, LDZ, W, LDZ, H, LDH, H, LDH, ZP, NZ, true); // symmetric: upper blocks computed, lower mirrored
    PHASE(18);
    PAR_FOR(i, NZ) gh[i] = io.g[(size_t)k * NZ + i] + H[i * LDH + NZ];
    SYNC();
    // [A B]_k is dead: when the active rows of this knot keep phase 2 clear of the buffer, fetch the next knot's now
    const bool early = !big && ((nca == 0) ? Lay::offCD : ((Lay::offCD + nca * NZ > Lay::offSg + nca * nca) ? Lay::offCD + nca * NZ : Lay::offSg + nca * nca)) <= Lay::offAB;
    if (k > 0 && early) stage_AB_async(k - 1);
    PHASE(5);
    // 8. KKT by block elimination (phase-2 buffers alias P/G/Li/AB/W, all dead now)
    const int ncol = NR + nca, ldz = (ncol + 7) & ~7; // Z columns: [rh | Sh' | D'], padded with zero columns to whole tiles
    const double *gCD = io.CDact + (size_t)k * NC * NZ;
    const int32_t *ai = io.act_idx + (size_t)k * NC;
    double *Kv = big ? U0 + Lay::offKvB : U0 + Lay::offKv, *Rh = big ? U0 + Lay::offRhB : U0 + Lay::offRh, *Sg = big ? U0 + Lay::offSgB : U0 + Lay::offSg;
    const double *CD = big ? gCD : CDs;
    if (!big) PAR_FOR(e, nca * NZ) CDs[e] = gCD[e];
    PAR_FOR(r, nca) dbr[r] = io.dbar[(size_t)k * NC + ai[r]];
    PAR_FOR(e, MR * MR) { // R^ = sym(H_uu), padded with the identity to MR x MR
      int i = e / MR, j = e % MR;
      Rh[i * LDR + j] = (i < M && j < M) ? 0.5 * (H[(N + i) * LDH + N + j] + H[(N + j) * LDH + N + i]) : ((i == j) ? 1.0 : 0.0);
    }
    SYNC();
    PAR_FOR(e, MR * ldz) {
      int i = e / ldz, c = e % ldz;
      Z[e] = (i >= M || c >= ncol) ? 0.0 : ((c == 0) ? gh[N + i] : (c < NR ? H[(N + i) * LDH + c - 1] : CD[(c - NR) * NZ + N + i])); // H_ux row i (H is symmetric to rounding): conflict-free
    }
    SYNC();
    PHASE(6);
#ifdef MPC_HOST_EMU
    chol_blocked(Rh, MR, LDR, dinv);
    trsm_blocked(Rh, MR, LDR, dinv, Z, ncol, ldz);
#else
    // Z <- R^-1 Z on the tensor pipe: blocked Cholesky, then forward / backward substitution with one warp per 8-column tile of Z
    chol_mma<MR / 8>(Rh, LDR, dinv);
    PHASE(7);
    trsm_mma<MR / 8>(Rh, LDR, dinv, Z, ldz, ldz / 8, nullptr, 0, ldz / 8);
#endif
    PHASE(8);
    PAR_FOR(e, nca * nca) {
      int r = e / nca, c = e % nca;
      double s = (r == c) ? mu : 0.0;
      for (int l = 0; l < M; l++) s += CD[r * NZ + N + l] * Z[l * ldz + NR + c];
      Sg[e] = s;
    }
    PAR_FOR(e, nca * NR) {
      int r = e / NR, c = e % NR;
      double s = (c == 0) ? dbr[r] : CD[r * NZ + c - 1];
      for (int l = 0; l < M; l++) s -= CD[r * NZ + N + l] * Z[l * ldz + c];
      Kv[e] = s;
    }
    SYNC();
    PHASE(9);
    if (nca > 0) { chol_blocked(Sg, nca, nca, dinv); trsm_blocked(Sg, nca, nca, dinv, Kv, NR, NR); }
    PHASE(10);
    PAR_FOR(e, M * NR) {
      int i = e / NR, c = e % NR;
      double s = -Z[i * ldz + c];
      for (int r = 0; r < nca; r++) s -= Z[i * ldz + NR + r] * Kv[r * NR + c];
      Z[i * ldz + c] = s;
    }
    SYNC();
    double *gK = io.K + (size_t)k * S * NR;
    PAR_FOR(e, M * NR) { int i = e / NR, c = e % NR; double v = Z[i * ldz + c]; gK[e] = v; if (c > 0) io.Kfb[((size_t)k * M + i) * N + c - 1] = v; }
    PAR_FOR(e, nca * NR) gK[M * NR + e] = Kv[e];
    PHASE(11);
    // 9. P = Qh + Sh Ku + C' Kv, p = qh + Sh ku + C' kv : in place in H[0:N,0:N] / p (symmetrised when the next knot loads it).
    //    Sh Ku = H[N:, 0:N]' Z[:, 1:] runs on the DMMA pipe (K = MP, zero rows beyond M); the active-row part is usually empty.
    PAR_FOR(i, N) {
      double s = gh[i];
      for (int l = 0; l < M; l++) s += H[(N + l) * LDH + i] * Z[l * ldz]; // H_ux rows (as the DMMA update below): conflict-free
      for (int r = 0; r < nca; r++) s += CD[r * NZ + i] * Kv[r * NR];
      p[i] = s;
    }
    mma_tn(NBLK, NBLK, MR, H + N * LDH, LDH, Z + 1, ldz, H, LDH, H, LDH, N, N, false);
    if (nca > 0) { // 2 x 4 register tiles over (row i, column c)
      constexpr int TC = N / 4;
      PAR_FOR(t, (N / 2) * TC) {
        const int i0 = (t / TC) * 2, c0 = (t % TC) * 4;
        double a[2][4];
#pragma unroll
        for (int r = 0; r < 2; r++)
#pragma unroll
          for (int q = 0; q < 4; q++) a[r][q] = H[(i0 + r) * LDH + c0 + q];
        for (int r = 0; r < nca; r++) {
          const double h0 = CD[r * NZ + i0], h1 = CD[r * NZ + i0 + 1];
          const double *kr = Kv + r * NR + 1 + c0;
#pragma unroll
          for (int q = 0; q < 4; q++) { a[0][q] += h0 * kr[q]; a[1][q] += h1 * kr[q]; }
        }
#pragma unroll
        for (int r = 0; r < 2; r++)
#pragma unroll
          for (int q = 0; q < 4; q++) H[(i0 + r) * LDH + c0 + q] = a[r][q];
      }
      SYNC();
    }
    PHASE(19);
    if (k > 0 && !early) stage_AB_async(k - 1); // the phase-2 buffers aliasing [A B] are dead now
    PHASE(12);
  }
  // ---- forward sweep, dx0 = 0 (force_initial_condition, fulldynamic_talos.py:384)
  PHASE(13);
  double acc = 0.0;
  PAR_FOR(i, N) { dx[i] = 0.0; io.dxs[i] = 0.0; io.dlams[i] = -p[i]; }
  // everything that does not depend on the sequential chain is done for all knots at once: default multiplier steps
  // dv = dbar / mu of the inactive rows and their part of the directional derivative
  PAR_FOR(e, (T + 1) * NC) {
    const int kk = e / NC, r = e % NC, na = io.nca[kk];
    const int32_t *ak = io.act_idx + (size_t)kk * NC;
    const double db = io.dbar[e];
    bool active = false;
    for (int q = 0; q < na; q++) active |= (ak[q] == r);
    io.dvs[e] = db / mu;
    if (!active) acc -= db * db / mu;
  }
  // W_k, [A B]_k, the small per-knot vectors and the first KROWS gain rows of each knot are staged global -> shared one knot
  // ahead (bulk copies, two stages in the now dead H / phase buffers): the dependent chain dx_k -> du_k -> dx_{k+1} never waits on HBM
  constexpr int FW = N * NZ;
  constexpr int FX = 4 * N + NZ + 36; // pt_k, fbar_k, lplus_{k+1}, lam_{k+1}, lxu_k, T6_k
  constexpr int KROWS = (((ZP * LDH + Lay::un - 4 * FW - 2 * FX) / 2) / NR) & ~1;
  constexpr int FSTAGE = 2 * FW + FX + KROWS * NR;
  static_assert(KROWS >= M && 2 * FSTAGE <= ZP * LDH + Lay::un && FSTAGE % 2 == 0, "forward staging does not fit");
  // one thread arms the stage's mbarrier and issues nine bulk copies (cp.async.bulk: W, [A B], gain rows, six small vectors)
  auto stage_fwd = [&](int kk) {
    ONE_THREAD {
      double *bw = ws + (kk & 1) * FSTAGE, *ba = bw + FW, *bx = ba + FW, *bk = bx + FX;
      unsigned long long *bar = mbar + (kk & 1);
      int rows = M + io.nca[kk];
      if (rows > KROWS) rows = KROWS;
      rows = (rows + 1) & ~1; // whole 16-byte units (an extra row stays inside this knot's gain block)
      FENCE_PROXY_ASYNC();    // the buffer was last touched through the generic proxy
      MBAR_EXPECT_TX(bar, (2 * FW + rows * NR + 4 * N + NZ + 36) * 8);
      BULK_G2S(bw, io.W + (size_t)kk * FW, FW * 8, bar);
      BULK_G2S(ba, io.AB + (size_t)kk * FW, FW * 8, bar);
      BULK_G2S(bk, io.K + (size_t)kk * S * NR, rows * NR * 8, bar);
      BULK_G2S(bx, io.pt + (size_t)kk * N, N * 8, bar);
      BULK_G2S(bx + N, io.fbar + (size_t)kk * N, N * 8, bar);
      BULK_G2S(bx + 2 * N, io.lplus + (size_t)(kk + 1) * N, N * 8, bar);
      BULK_G2S(bx + 3 * N, io.lam + (size_t)(kk + 1) * N, N * 8, bar);
      BULK_G2S(bx + 4 * N, io.lxu + (size_t)kk * NZ, NZ * 8, bar);
      BULK_G2S(bx + 4 * N + NZ, io.T6 + (size_t)kk * 36, 36 * 8, bar);
    }
  };
  static_assert((FW * 8) % 16 == 0 && (N * 8) % 16 == 0 && (NZ * 8) % 16 == 0 && (FX * 8) % 16 == 0 && (2 * NR * 8) % 16 == 0 && (S * NR * 8) % 16 == 0,
                "bulk copies need 16-byte sizes and offsets");
  SYNC();
  if (T > 0) stage_fwd(0);
  for (int k = 0; k < T; k++) {
    const int nca = io.nca[k];
    if (k + 1 < T) stage_fwd(k + 1);
    MBAR_WAIT(mbar + (k & 1), (k >> 1) & 1); // stage k has landed (each barrier is used every other knot)
    const double *sW = ws + (k & 1) * FSTAGE, *sAB = sW + FW, *sX = sAB + FW, *sK = sX + FX;
    const double *gK = io.K + (size_t)k * S * NR;
    // du, dv of the active rows: RA rows per warp at a time (independent reduction chains), columns over lanes
    constexpr int RA = 3;
    for (int base = WARP_ID * RA; base < M + nca; base += NWARPS * RA) {
      double sa[RA];
#pragma unroll
      for (int r = 0; r < RA; r++) sa[r] = 0.0;
      LANE_FOR(j, N) {
        const double dj = dx[j];
#pragma unroll
        for (int r = 0; r < RA; r++) {
          const int i = base + r;
          if (i < M + nca) sa[r] += ((i < KROWS) ? sK[i * NR + 1 + j] : gK[i * NR + 1 + j]) * dj;
        }
      }
#pragma unroll
      for (int r = 0; r < RA; r++) sa[r] = WARP_SUM(sa[r]);
      if (LANE0) {
#pragma unroll
        for (int r = 0; r < RA; r++) {
          const int i = base + r;
          if (i >= M + nca) continue;
          const double v = sa[r] + ((i < KROWS) ? sK[i * NR] : gK[i * NR]);
          if (i < M) { z[N + i] = v; io.dus[(size_t)k * M + i] = v; } else dva[i - M] = v;
        }
      }
    }
    PAR_FOR(i, N) z[i] = dx[i];
    SYNC();
    if (nca > 0) {
      const int32_t *ai = io.act_idx + (size_t)k * NC;
      const double *gdb = io.dbar + (size_t)k * NC, *gvp = io.vplus + (size_t)k * NC, *gv = io.v + (size_t)k * NC;
      PAR_FOR(r, nca) {
        const int row = ai[r];
        io.dvs[(size_t)k * NC + row] = dva[r];
        acc += (2.0 * gvp[row] - gv[row]) * (mu * dva[r] - gdb[row]) - gdb[row] * dva[r];
      }
    }
    PAR_FOR(i, NZ) acc += sX[4 * N + i] * z[i];
    // dlam_{k+1} = pt + W z ; tmp = A dx + B du + fbar - mu_d dlam : RB rows per warp at a time
    constexpr int RB = 7;
    for (int base = WARP_ID * RB; base < N; base += NWARPS * RB) {
      double sl[RB], sa[RB];
#pragma unroll
      for (int r = 0; r < RB; r++) { sl[r] = 0.0; sa[r] = 0.0; }
      LANE_FOR(j, NZ) {
        const double zj = z[j];
#pragma unroll
        for (int r = 0; r < RB; r++)
          if (base + r < N) { sl[r] += sW[(base + r) * NZ + j] * zj; sa[r] += sAB[(base + r) * NZ + j] * zj; }
      }
#pragma unroll
      for (int r = 0; r < RB; r++) { sl[r] = WARP_SUM(sl[r]); sa[r] = WARP_SUM(sa[r]); }
      if (LANE0) {
#pragma unroll
        for (int r = 0; r < RB; r++) {
          const int i = base + r;
          if (i >= N) continue;
          const double fbi = sX[N + i], lp = sX[2 * N + i], lm = sX[3 * N + i], dl = sl[r] + sX[i];
          tmp[i] = sa[r] + fbi - mu_d * dl;
          io.dlams[(size_t)(k + 1) * N + i] = dl;
          acc += (2.0 * lp - lm) * (mu_d * dl - fbi) - fbi * dl;
        }
      }
    }
    SYNC();
    PAR_FOR(i, N) {
      double v;
      if (i < 6) { v = 0; for (int l = 0; l < 6; l++) v += sX[4 * N + NZ + 6 * i + l] * tmp[l]; } else v = tmp[i];
      dx[i] = v; io.dxs[(size_t)(k + 1) * N + i] = v;
    }
    PHASE(20);
    SYNC();
  }
  { // terminal knot: dv of the active terminal rows and the terminal cost gradient
    const int nca = io.nca[T];
    const int32_t *ai = io.act_idx + (size_t)T * NC;
    const double *gdb = io.dbar + (size_t)T * NC, *gvp = io.vplus + (size_t)T * NC, *gv = io.v + (size_t)T * NC;
    const double *CT = io.CDact + (size_t)T * NC * NZ;
    PAR_FOR(r, nca) {
      const int row = ai[r];
      double sv = gdb[row];
      for (int j = 0; j < N; j++) sv += CT[r * NZ + j] * dx[j];
      sv /= mu;
      io.dvs[(size_t)T * NC + row] = sv;
      acc += (2.0 * gvp[row] - gv[row]) * (mu * sv - gdb[row]) - gdb[row] * sv;
    }
    PAR_FOR(i, N) acc += io.lxu[(size_t)T * NZ + i] * dx[i];
  }
  PHASE(14);
  PHASE_DUMP(io.phase_out);
  red[TID] = acc;
  SYNC();
  ONE_THREAD { double s = 0; for (int t = 0; t < NTHREADS; t++) s += red[t]; io.dphi[0] = s; }
  SYNC();
}

} // namespace mpcdev
