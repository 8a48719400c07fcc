// FP64 tensor-core GEMM on shared-memory operands (mma.sync.aligned.m8n8k4.f64 -> SASS DMMA.8x8x4).
//
// tcgen05/TMEM have no fp64 path on Blackwell, so dense fp64 contractions use the legacy DMMA pipe (same nominal
// peak as DFMA on B200, but 256 FMAs per warp instruction instead of 32: far lower issue and shared-memory
// pressure).  Used only where the stage matrices really are dense contractions (BASELINE north_star): Lambda^-1 P,
// Pt [A B] and [A B]' W in the proximal Riccati recursion.
//
// All products are in "TN" form C = A^T B with A stored [K][lda], B stored [K][ldb] (row-major), so both fragments
// read 8 consecutive doubles per k-row: conflict-free for leading dimensions = 8 (mod 16) doubles.
// Fragment layout (PTX ISA, m8n8k4 .f64): lane = 4*g + t;  a: A^T[row g][k t];  b: B[k t][col g];  c0,c1: C[row g][cols 2t, 2t+1].
#pragma once
#include "dev_common.cuh"

namespace mpcdev {

#ifndef MPC_HOST_EMU
__device__ __forceinline__ void dmma_8x8x4(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
#endif

// C (8*mt x 8*nt, ldc) = init + A^T B,  K multiple of 4.
//   init: 0 (Cinit == nullptr) or Cinit (ldci) restricted to rows < vr and cols < vc (zero outside) — lets the
//   Hessian update read H_k straight from global memory while keeping the shared copy zero-padded.
// upper_only: skip tiles strictly below the block diagonal (caller mirrors).
HD void mma_tn(int mt, int nt, int K, const double *A, int lda, const double *B, int ldb, double *C, int ldc, const double *Cinit, int ldci,
               int vr, int vc, bool upper_only) {
#ifdef MPC_HOST_EMU
  for (int i = 0; i < 8 * mt; i++)
    for (int j = 0; j < 8 * nt; j++) {
      if (upper_only && (j / 8) < (i / 8)) continue;
      double s = (Cinit && i < vr && j < vc) ? Cinit[i * ldci + j] : 0.0;
      for (int k = 0; k < K; k++) s += A[k * lda + i] * B[k * ldb + j];
      C[i * ldc + j] = s;
    }
#else
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  // each warp takes pairs of horizontally adjacent tiles (shares the A fragment)
  const int ntp = (nt + 1) / 2;
  for (int p = warp; p < mt * ntp; p += nwarps) {
    const int ti = p / ntp, tj = (p % ntp) * 2;
    const bool two = (tj + 1 < nt);
    if (upper_only && tj + (two ? 1 : 0) < ti) continue;
    const int row = ti * 8 + g, col0 = tj * 8 + 2 * t, col1 = col0 + 8;
    double c00 = 0, c01 = 0, c10 = 0, c11 = 0;
    if (Cinit && row < vr) {
      if (col0 < vc) c00 = Cinit[row * ldci + col0];
      if (col0 + 1 < vc) c01 = Cinit[row * ldci + col0 + 1];
      if (two && col1 < vc) c10 = Cinit[row * ldci + col1];
      if (two && col1 + 1 < vc) c11 = Cinit[row * ldci + col1 + 1];
    }
    const double *ap = A + t * lda + ti * 8 + g;
    const double *bp = B + t * ldb + tj * 8 + g;
    if (two) {
#pragma unroll 2
      for (int k0 = 0; k0 < K; k0 += 4) {
        double a = ap[k0 * lda], b0 = bp[k0 * ldb], b1 = bp[k0 * ldb + 8];
        dmma_8x8x4(c00, c01, a, b0);
        dmma_8x8x4(c10, c11, a, b1);
      }
    } else {
#pragma unroll 2
      for (int k0 = 0; k0 < K; k0 += 4) dmma_8x8x4(c00, c01, ap[k0 * lda], bp[k0 * ldb]);
    }
    *reinterpret_cast<double2 *>(C + row * ldc + col0) = make_double2(c00, c01);
    if (two) *reinterpret_cast<double2 *>(C + row * ldc + col1) = make_double2(c10, c11);
  }
#endif
  SYNC();
}

} // namespace mpcdev
