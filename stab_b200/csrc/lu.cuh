// lu.cuh -- CTA-cooperative LU with partial pivoting of C0 (n x n) applied to a 2n-column right
// hand side, then the upper-triangular back substitution: X = C0^-1 [R1 | R2].
// Reference: ZGETRF + two ZGETRS calls in spatial.f90:978-1004, producing the top half
// [-C0^-1 C1, -C0^-1 C2] of the companion matrix (spatial.f90:1007-1008); the right hand side
// arrives already negated from the assembly kernel and lives in the companion buffer (ldb = 2n).
//
// v1: right-looking elimination, one warp per trailing/RHS column (coalesced column walks), pivot
// search with IZAMAX semantics (|re|+|im|, first maximum).  Algorithmic work (8/3 + 16) n^3 flops.
#pragma once
#include "common.cuh"

namespace stab {

// sl: shared vector of n complex (multipliers of the current column)
SD_DEV int cta_lu_solve(const Cta& c, cplx* C, int n, int ldc, cplx* B, int nrhs, int ldb, cplx* sl, int* ipiv = nullptr) {
  int info = 0;
  for (int k = 0; k < n; ++k) {
    cplx* ck = C + (size_t)k * ldc;
    double best = -1.0; int bi = k;
    for (int r = k + c.tid; r < n; r += c.nt) {
      double m = cabs1(ck[r]);
      if (m > best) { best = m; bi = r; }
    }
    cta_argmax(c, best, bi);
    const int p = bi;
    if (ipiv && c.tid == 0) ipiv[k] = p;
    if (best == 0.0) { if (info == 0) info = k + 1; cta_sync(); continue; }
    if (p != k) {
      for (int j = c.tid; j < n; j += c.nt) { cplx t = C[k + (size_t)j * ldc]; C[k + (size_t)j * ldc] = C[p + (size_t)j * ldc]; C[p + (size_t)j * ldc] = t; }
      for (int j = c.tid; j < nrhs; j += c.nt) { cplx t = B[k + (size_t)j * ldb]; B[k + (size_t)j * ldb] = B[p + (size_t)j * ldb]; B[p + (size_t)j * ldb] = t; }
    }
    cta_sync();
    const cplx rp = cdiv(mk(1.0, 0.0), ck[k]);
    cta_sync();
    for (int r = k + 1 + c.tid; r < n; r += c.nt) { cplx l = ck[r] * rp; ck[r] = l; sl[r] = l; }
    cta_sync();
    // trailing update, one warp per column (matrix columns k+1..n-1, then all RHS columns)
    const int ntrail = (n - k - 1) + nrhs;
    for (int q = c.wid; q < ntrail; q += c.nw) {
      cplx* col = (q < n - k - 1) ? (C + (size_t)(k + 1 + q) * ldc) : (B + (size_t)(q - (n - k - 1)) * ldb);
      const cplx u = col[k];
      if (is_zero(u)) continue;
      for (int r = k + 1 + c.lane; r < n; r += c.ws) col[r] = col[r] - sl[r] * u;
    }
    cta_sync();
  }
  // back substitution U X = Y, one warp per right-hand-side column
  for (int q = c.wid; q < nrhs; q += c.nw) {
    cplx* col = B + (size_t)q * ldb;
    for (int k = n - 1; k >= 0; --k) {
      const cplx* uk = C + (size_t)k * ldc;
      const cplx xk = cdiv(col[k], uk[k]);
      warp_sync();
      if (c.lane == 0) col[k] = xk;
      if (!is_zero(xk))
        for (int r = c.lane; r < k; r += c.ws) col[r] = col[r] - uk[r] * xk;
      warp_sync();
    }
  }
  cta_sync();
  return info;
}

}  // namespace stab
