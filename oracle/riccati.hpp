// ORACLE (test infrastructure, not product code) — dual-regularised constrained LQ solve (proximal Riccati).
//
// Restates aligator::gar::ProximalRiccatiSolver (serial form), which the reference selects with
// `solver.linear_solver_choice = aligator.LQ_SOLVER_PARALLEL` (fulldynamic_talos.py:383; the parallel
// variant returns the same solution up to round-off).  Aligator is absent from
// /root/reference => PARITY UNPINNED; the recursion below is derived in SURVEY App. A6 and verified in
// tests against a dense solve of the full KKT system.
//
// Knot LQ (deltas): min 1/2 z'Hz + g'z,  z=(dx,du)
//   s.t.  A dx + B du + E dx' + fbar = mu_d dlam'      (E = blockdiag(E6, -I))
//         C dx + D du + dbar        = mu  dv           (inactive rows of C,D are zero)
#pragma once
#include <cmath>
#include <cstring>
#include <vector>

namespace orc {

// no-pivot LDL^T of a symmetric quasi-definite matrix (in place: L strictly lower, D on the diagonal)
inline void ldlt(double *K, int s) {
  for (int j = 0; j < s; j++) {
    double d = K[j * s + j];
    for (int k = 0; k < j; k++) d -= K[j * s + k] * K[j * s + k] * K[k * s + k];
    K[j * s + j] = d;
    for (int i = j + 1; i < s; i++) {
      double t = K[i * s + j];
      for (int k = 0; k < j; k++) t -= K[i * s + k] * K[j * s + k] * K[k * s + k];
      K[i * s + j] = t / d;
    }
  }
}
inline void ldlt_solve(const double *K, int s, double *b, int nrhs) { // b: s x nrhs row-major
  for (int i = 0; i < s; i++)
    for (int k = 0; k < i; k++) { double l = K[i * s + k]; if (l != 0.0) { double *bi = b + i * nrhs; const double *bk = b + k * nrhs;
_Pragma("GCC ivdep") for (int c = 0; c < nrhs; c++) bi[c] -= l * bk[c]; } }
  for (int i = 0; i < s; i++) { double d = 1.0 / K[i * s + i]; for (int c = 0; c < nrhs; c++) b[i * nrhs + c] *= d; }
  for (int i = s - 1; i >= 0; i--)
    for (int k = i + 1; k < s; k++) { double l = K[k * s + i]; if (l != 0.0) { double *bi = b + i * nrhs; const double *bk = b + k * nrhs;
_Pragma("GCC ivdep") for (int c = 0; c < nrhs; c++) bi[c] -= l * bk[c]; } }
}
inline void chol_d(double *A, int n) {
  for (int j = 0; j < n; j++) {
    double d = A[j * n + j];
    for (int k = 0; k < j; k++) d -= A[j * n + k] * A[j * n + k];
    d = std::sqrt(d);
    A[j * n + j] = d;
    for (int i = j + 1; i < n; i++) {
      double s = A[i * n + j];
      for (int k = 0; k < j; k++) s -= A[i * n + k] * A[j * n + k];
      A[i * n + j] = s / d;
    }
  }
}
inline void chol_solve_d(const double *L, int n, double *b, int nrhs) { // b: n x nrhs row-major
  for (int i = 0; i < n; i++) {
    for (int k = 0; k < i; k++) { double l = L[i * n + k]; double *bi = b + i * nrhs; const double *bk = b + k * nrhs;
_Pragma("GCC ivdep") for (int c = 0; c < nrhs; c++) bi[c] -= l * bk[c]; }
    double d = 1.0 / L[i * n + i]; for (int c = 0; c < nrhs; c++) b[i * nrhs + c] *= d;
  }
  for (int i = n - 1; i >= 0; i--) {
    for (int k = i + 1; k < n; k++) { double l = L[k * n + i]; double *bi = b + i * nrhs; const double *bk = b + k * nrhs;
_Pragma("GCC ivdep") for (int c = 0; c < nrhs; c++) bi[c] -= l * bk[c]; }
    double d = 1.0 / L[i * n + i]; for (int c = 0; c < nrhs; c++) b[i * nrhs + c] *= d;
  }
}
// Dense products of the recursion, register-blocked (4 rows x 32 columns of C held in vector registers over the whole k loop; GCC
// vector extensions: AVX-512 / AVX2 code under -march=native, generic SIMD otherwise) — the CPU arm of the benchmark should not
// lose to the GPU because of a naive GEMM ("blocked or at least vectorised dense kernels").
typedef double v8d __attribute__((vector_size(64), aligned(8)));
template <bool TA> inline void gemm_blocked(int m, int n, int k, const double *__restrict A, const double *__restrict B, double *__restrict C) {
  // TA: A is k x m (C += A^T B); else A is m x k (C += A B).  B: k x n, C: m x n, all row-major.
  int jb = 0;
  for (; jb + 32 <= n; jb += 32)
    for (int ib = 0; ib < m; ib += 4) {
      const int rows = (m - ib < 4) ? m - ib : 4;
      v8d acc[4][4];
      for (int r = 0; r < 4; r++) for (int v = 0; v < 4; v++) { if (r < rows) std::memcpy(&acc[r][v], C + (size_t)(ib + r) * n + jb + 8 * v, 64); else acc[r][v] = v8d{0, 0, 0, 0, 0, 0, 0, 0}; }
      for (int p = 0; p < k; p++) {
        v8d b[4];
        for (int v = 0; v < 4; v++) std::memcpy(&b[v], B + (size_t)p * n + jb + 8 * v, 64);
        for (int r = 0; r < 4; r++) {
          const double a = (r < rows) ? (TA ? A[(size_t)p * m + ib + r] : A[(size_t)(ib + r) * k + p]) : 0.0;
          for (int v = 0; v < 4; v++) acc[r][v] += a * b[v];
        }
      }
      for (int r = 0; r < rows; r++) for (int v = 0; v < 4; v++) std::memcpy(C + (size_t)(ib + r) * n + jb + 8 * v, &acc[r][v], 64);
    }
  if (jb < n) // remaining columns: rank-1 updates, vectorised over j
    for (int p = 0; p < k; p++)
      for (int i = 0; i < m; i++) {
        const double a = TA ? A[(size_t)p * m + i] : A[(size_t)i * k + p];
        if (a == 0.0) continue;
        double *ci = C + (size_t)i * n; const double *bp = B + (size_t)p * n;
_Pragma("GCC ivdep") for (int j = jb; j < n; j++) ci[j] += a * bp[j];
      }
}
// C (m x n) += A^T (k x m)^T * B (k x n)
inline void gemm_tn(int m, int n, int k, const double *__restrict A, const double *__restrict B, double *__restrict C) { gemm_blocked<true>(m, n, k, A, B, C); }
// C (m x n) += A (m x k) * B (k x n)
inline void gemm_nn(int m, int n, int k, const double *__restrict A, const double *__restrict B, double *__restrict C) { gemm_blocked<false>(m, n, k, A, B, C); }
inline void inv6(const double *A, double *Ai) { // Gauss-Jordan with partial pivoting
  double M[6][12];
  for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) { M[i][j] = A[6 * i + j]; M[i][6 + j] = i == j; }
  for (int c = 0; c < 6; c++) {
    int p = c; for (int r = c + 1; r < 6; r++) if (std::fabs(M[r][c]) > std::fabs(M[p][c])) p = r;
    if (p != c) for (int j = 0; j < 12; j++) std::swap(M[p][j], M[c][j]);
    double d = 1.0 / M[c][c]; for (int j = 0; j < 12; j++) M[c][j] *= d;
    for (int r = 0; r < 6; r++) if (r != c) { double f = M[r][c]; if (f != 0.0) for (int j = 0; j < 12; j++) M[r][j] -= f * M[c][j]; }
  }
  for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) Ai[6 * i + j] = M[i][6 + j];
}

struct LQKnot { // views into caller storage
  const double *H, *g;     // (n+m)^2, n+m
  const double *A, *B;     // n x n, n x m
  const double *E6;        // 6x6 (nullptr => -I)
  const double *f;         // n
  const double *C, *D, *d; // nc x n, nc x m, nc
};

struct LQSolution {
  int n, m, nc, T, nct;
  std::vector<double> K;   // T x (m+nc) x (1+n): [k | K] rows: du then dv
  std::vector<double> W, pt, T6; // T x n x (n+m), T x n, T x 36
  std::vector<double> P0, p0;
  std::vector<double> PT, pT;
  std::vector<double> dxs, dus, dvs, dlams; // (T+1)n, T m, (T+1) nc, (T+1) n
};

struct RiccatiWork {
  std::vector<double> P, p, Pt, G, AB, Wk, Hh, gh, KK, rhs, tmp;
};

// terminal: HT (n x n with stride ldh), gT, CT (nct x n), dT
inline void riccati_solve(int n, int m, int nc, int T, const LQKnot *kn, const double *HT, int ldh, const double *gT, const double *CT,
                          const double *dT, int nct, double mu_d, double mu, LQSolution &sol) {
  const int nz = n + m, s = m + nc, nr = 1 + n;
  sol.n = n; sol.m = m; sol.nc = nc; sol.T = T; sol.nct = nct;
  sol.K.assign((size_t)T * s * nr, 0.0); sol.W.assign((size_t)T * n * nz, 0.0); sol.pt.assign((size_t)T * n, 0.0); sol.T6.assign((size_t)T * 36, 0.0);
  std::vector<double> P(n * n), p(n), Pt(n * n), G(n * n), W(n * nz), AB(n * nz), Hh(nz * nz), gh(nz), KK(s * s), rhs(s * nr), tmp(n), ptil(n), Wk_scratch;
  // terminal value function
  for (int i = 0; i < n; i++) { p[i] = gT[i]; for (int j = 0; j < n; j++) P[i * n + j] = HT[i * ldh + j]; }
  for (int r = 0; r < nct; r++)
    for (int i = 0; i < n; i++) { double c = CT[r * n + i] / mu; if (c == 0.0) continue; p[i] += c * dT[r]; for (int j = 0; j < n; j++) P[i * n + j] += c * CT[r * n + j]; }
  sol.PT = P; sol.pT = p;
  for (int k = T - 1; k >= 0; k--) {
    const LQKnot &q = kn[k];
    // 1. E normalisation: T6 = -E6^-1 ; P <- T'PT, p <- T'p on the first 6 coordinates
    double T6[36];
    if (q.E6) { double Ei[36]; inv6(q.E6, Ei); for (int i = 0; i < 36; i++) T6[i] = -Ei[i]; }
    else for (int i = 0; i < 36; i++) T6[i] = (i % 7 == 0) ? 1.0 : 0.0;
    std::memcpy(&sol.T6[(size_t)k * 36], T6, sizeof T6);
    if (n >= 6 && q.E6) {
      // columns
      for (int i = 0; i < n; i++) { double r6[6]; for (int j = 0; j < 6; j++) { double t = 0; for (int l = 0; l < 6; l++) t += P[i * n + l] * T6[6 * l + j]; r6[j] = t; } for (int j = 0; j < 6; j++) P[i * n + j] = r6[j]; }
      // rows
      for (int j = 0; j < n; j++) { double c6[6]; for (int i = 0; i < 6; i++) { double t = 0; for (int l = 0; l < 6; l++) t += T6[6 * l + i] * P[l * n + j]; c6[i] = t; } for (int i = 0; i < 6; i++) P[i * n + j] = c6[i]; }
      double c6[6]; for (int i = 0; i < 6; i++) { double t = 0; for (int l = 0; l < 6; l++) t += T6[6 * l + i] * p[l]; c6[i] = t; } for (int i = 0; i < 6; i++) p[i] = c6[i];
    }
    // 2. Lambda = I + mu_d P ; Pt = Lambda^-1 P ; ptil = Lambda^-1 (p + P f)
    for (int i = 0; i < n * n; i++) G[i] = mu_d * P[i];
    for (int i = 0; i < n; i++) G[i * n + i] += 1.0;
    chol_d(G.data(), n);
    Pt = P; chol_solve_d(G.data(), n, Pt.data(), n);
    for (int i = 0; i < n; i++) { double t = p[i]; for (int j = 0; j < n; j++) t += P[i * n + j] * q.f[j]; ptil[i] = t; }
    chol_solve_d(G.data(), n, ptil.data(), 1);
    // 3. W = Pt [A B];  Hh = H + [A B]' W ; gh = g + [A B]' ptil
    for (int i = 0; i < n; i++) { std::memcpy(&AB[i * nz], &q.A[i * n], sizeof(double) * n); std::memcpy(&AB[i * nz + n], &q.B[i * m], sizeof(double) * m); }
    std::fill(W.begin(), W.end(), 0.0);
    gemm_nn(n, nz, n, Pt.data(), AB.data(), W.data());
    std::memcpy(Hh.data(), q.H, sizeof(double) * nz * nz);
    gemm_tn(nz, nz, n, AB.data(), W.data(), Hh.data());
    for (int i = 0; i < nz; i++) { double t = q.g[i]; for (int l = 0; l < n; l++) t += AB[l * nz + i] * ptil[l]; gh[i] = t; }
    std::memcpy(&sol.W[(size_t)k * n * nz], W.data(), sizeof(double) * n * nz);
    std::memcpy(&sol.pt[(size_t)k * n], ptil.data(), sizeof(double) * n);
    // 4. KKT [Rh D'; D -mu I] X = -[rh Sh'; d C].  Rows of (C, D) that are entirely zero (inactive constraints) decouple
    //    exactly to kv = d/mu, Kv = 0 and are eliminated before the factorisation (same solution, smaller LDL').
    int nact = 0;
    std::vector<int> arow(nc);
    for (int r = 0; r < nc; r++) {
      bool nzr = false;
      for (int j = 0; j < n && !nzr; j++) nzr = q.C[r * n + j] != 0.0;
      for (int j = 0; j < m && !nzr; j++) nzr = q.D[r * m + j] != 0.0;
      if (nzr) arow[nact++] = r;
    }
    const int sa = m + nact;
    std::fill(KK.begin(), KK.begin() + sa * sa, 0.0);
    for (int i = 0; i < m; i++) for (int j = 0; j < m; j++) KK[i * sa + j] = 0.5 * (Hh[(n + i) * nz + n + j] + Hh[(n + j) * nz + n + i]);
    for (int a = 0; a < nact; a++) { int r = arow[a]; for (int j = 0; j < m; j++) { KK[(m + a) * sa + j] = q.D[r * m + j]; KK[j * sa + m + a] = q.D[r * m + j]; } KK[(m + a) * sa + m + a] = -mu; }
    std::vector<double> &ra = Wk_scratch;
    ra.assign((size_t)sa * nr, 0.0);
    for (int i = 0; i < m; i++) { ra[i * nr] = -gh[n + i]; for (int j = 0; j < n; j++) ra[i * nr + 1 + j] = -Hh[j * nz + n + i]; }
    for (int a = 0; a < nact; a++) { int r = arow[a]; ra[(m + a) * nr] = -q.d[r]; for (int j = 0; j < n; j++) ra[(m + a) * nr + 1 + j] = -q.C[r * n + j]; }
    ldlt(KK.data(), sa);
    ldlt_solve(KK.data(), sa, ra.data(), nr);
    std::fill(rhs.begin(), rhs.end(), 0.0);
    std::memcpy(rhs.data(), ra.data(), sizeof(double) * m * nr);
    for (int r = 0; r < nc; r++) rhs[(m + r) * nr] = q.d[r] / mu; // inactive rows; active ones overwritten next
    for (int a = 0; a < nact; a++) std::memcpy(&rhs[(m + arow[a]) * nr], &ra[(m + a) * nr], sizeof(double) * nr);
    std::memcpy(&sol.K[(size_t)k * s * nr], rhs.data(), sizeof(double) * s * nr);
    // 5. P = Qh + Sh Ku + C' Kv (symmetrised); p = qh + Sh ku + C' kv
    for (int i = 0; i < n; i++) {
      double t = gh[i];
      for (int l = 0; l < m; l++) t += Hh[i * nz + n + l] * rhs[l * nr];
      for (int r = 0; r < nc; r++) t += q.C[r * n + i] * rhs[(m + r) * nr];
      p[i] = t;
      double *row = &Pt[i * n];
      for (int j = 0; j < n; j++) row[j] = Hh[i * nz + j];
      for (int l = 0; l < m; l++) { double h = Hh[i * nz + n + l]; const double *kr = &rhs[l * nr + 1];
_Pragma("GCC ivdep") for (int j = 0; j < n; j++) row[j] += h * kr[j]; }
      for (int r = 0; r < nc; r++) { double c = q.C[r * n + i]; if (c == 0.0) continue; const double *kr = &rhs[(m + r) * nr + 1];
_Pragma("GCC ivdep") for (int j = 0; j < n; j++) row[j] += c * kr[j]; }
    }
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) P[i * n + j] = 0.5 * (Pt[i * n + j] + Pt[j * n + i]);
  }
  sol.P0 = P; sol.p0 = p;
  // forward, dx0 = 0
  sol.dxs.assign((size_t)(T + 1) * n, 0.0); sol.dus.assign((size_t)T * m, 0.0); sol.dvs.assign((size_t)(T + 1) * nc, 0.0); sol.dlams.assign((size_t)(T + 1) * n, 0.0);
  for (int i = 0; i < n; i++) sol.dlams[i] = -p[i];
  std::vector<double> z(nz);
  for (int k = 0; k < T; k++) {
    const double *dx = &sol.dxs[(size_t)k * n];
    const double *Kk = &sol.K[(size_t)k * s * nr];
    for (int i = 0; i < s; i++) {
      double t = Kk[i * nr]; for (int j = 0; j < n; j++) t += Kk[i * nr + 1 + j] * dx[j];
      if (i < m) sol.dus[(size_t)k * m + i] = t; else sol.dvs[(size_t)k * nc + i - m] = t;
    }
    for (int j = 0; j < n; j++) z[j] = dx[j];
    for (int j = 0; j < m; j++) z[n + j] = sol.dus[(size_t)k * m + j];
    const double *Wk = &sol.W[(size_t)k * n * nz];
    double *dl = &sol.dlams[(size_t)(k + 1) * n], *dxn = &sol.dxs[(size_t)(k + 1) * n];
    for (int i = 0; i < n; i++) {
      double t = sol.pt[(size_t)k * n + i], a = kn[k].f[i];
      for (int j = 0; j < nz; j++) t += Wk[i * nz + j] * z[j];
      for (int j = 0; j < n; j++) a += kn[k].A[i * n + j] * z[j];
      for (int j = 0; j < m; j++) a += kn[k].B[i * m + j] * z[n + j];
      dl[i] = t; tmp[i] = a - mu_d * t;
    }
    const double *T6 = &sol.T6[(size_t)k * 36];
    for (int i = 0; i < n; i++) {
      if (i < 6 && n >= 6 && kn[k].E6) { double t = 0; for (int l = 0; l < 6; l++) t += T6[6 * i + l] * tmp[l]; dxn[i] = t; } else dxn[i] = tmp[i];
    }
  }
  for (int r = 0; r < nct; r++) { double t = dT[r]; for (int j = 0; j < n; j++) t += CT[r * n + j] * sol.dxs[(size_t)T * n + j]; sol.dvs[(size_t)T * nc + r] = t / mu; }
}

} // namespace orc
