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

// Solve A x = b for one right-hand side with the factors and pivots left by cta_lu_solve
// (ZGETRS 'N' with nrhs = 1); b is overwritten by x.  b may live in shared or global memory.
SD_DEV void cta_lu_resolve(const Cta& c, const cplx* LU, int n, int ldc, const int* ipiv, cplx* b) {
  cta_sync();
  if (c.tid == 0)                                     // P: all row interchanges first (ZLASWP), the stored L has them applied
    for (int k = 0; k < n; ++k) {
      const int p = ipiv[k];
      if (p != k) { cplx t = b[k]; b[k] = b[p]; b[p] = t; }
    }
  for (int k = 0; k < n; ++k) {                       // L (unit lower)
    cta_sync();
    const cplx bk = b[k];
    if (!is_zero(bk))
      for (int r = k + 1 + c.tid; r < n; r += c.nt) fms_acc(b[r], LU[r + (size_t)k * ldc], bk);
  }
  for (int k = n - 1; k >= 0; --k) {                  // U
    cta_sync();
    if (c.tid == 0) b[k] = cdiv(b[k], LU[k + (size_t)k * ldc]);
    cta_sync();
    const cplx xk = b[k];
    if (!is_zero(xk))
      for (int r = c.tid; r < k; r += c.nt) fms_acc(b[r], LU[r + (size_t)k * ldc], xk);
  }
  cta_sync();
}

// Shift-invert inverse iteration on the pencil (A0, B0) near sigma (stage 4 of the north star;
// new functionality: the reference polishes modes with the external `shoot` program).
//   K = A0 - sigma B0 = P L U once;  repeat: z = K^-1 B0 x,  x = z/||z||,
//   lambda = (x^H A0 x)/(x^H B0 x),  resid = ||A0 x - lambda B0 x|| / (||A0 x|| + |lambda| ||B0 x||).
// One CTA.  K (n x n) is overwritten by the factors; x, u, v, z: n-vectors in global memory.
// out4: lambda (re, im), resid, iterations.
SD_DEV void cta_polish(const Cta& c, const cplx* A0, const cplx* B0, cplx* K, int n, cplx sigma, cplx* x, cplx* u, cplx* v,
                       int* ipiv, cplx* sl, int max_iters, double tol, double* out4) {
  for (size_t q = c.tid; q < (size_t)n * n; q += c.nt) K[q] = A0[q] - sigma * B0[q];
  cta_sync();
  const int sing = cta_lu_solve(c, K, n, n, nullptr, 0, n, sl, ipiv);
  cta_sync();
  double nx = 0.0;
  for (int r = c.tid; r < n; r += c.nt) nx += abs2(x[r]);
  nx = sqrt(cta_sum(c, nx));
  for (int r = c.tid; r < n; r += c.nt) x[r] = x[r] * (1.0 / nx);
  cta_sync();
  cplx lam = sigma;
  double resid = 1.0;
  int it = 0;
  for (; it < max_iters; ++it) {
    for (int r = c.tid; r < n; r += c.nt) {               // v = B0 x
      cplx s = mk(0.0, 0.0);
      for (int j = 0; j < n; ++j) fma_acc(s, B0[r + (size_t)j * n], x[j]);
      v[r] = s;
    }
    cta_sync();
    cta_lu_resolve(c, K, n, n, ipiv, v);                  // z = K^-1 B0 x
    double nz = 0.0;
    for (int r = c.tid; r < n; r += c.nt) nz += abs2(v[r]);
    nz = sqrt(cta_sum(c, nz));
    for (int r = c.tid; r < n; r += c.nt) x[r] = v[r] * (1.0 / nz);
    cta_sync();
    cplx xu = mk(0.0, 0.0), xv = mk(0.0, 0.0);
    for (int r = c.tid; r < n; r += c.nt) {               // u = A0 x, v = B0 x
      cplx s = mk(0.0, 0.0), t = mk(0.0, 0.0);
      for (int j = 0; j < n; ++j) { const cplx xj = x[j]; fma_acc(s, A0[r + (size_t)j * n], xj); fma_acc(t, B0[r + (size_t)j * n], xj); }
      u[r] = s; v[r] = t;
      fma_acc_conj(xu, x[r], s); fma_acc_conj(xv, x[r], t);
    }
    xu = cta_sum(c, xu); xv = cta_sum(c, xv);
    lam = cdiv(xu, xv);
    double rr = 0.0, nu = 0.0, nv = 0.0, dm = 0.0;
    for (int r = c.tid; r < n; r += c.nt) { rr += abs2(u[r] - lam * v[r]); nu += abs2(u[r]); nv += abs2(v[r]); }
    cta_sum4(c, rr, nu, nv, dm);
    resid = sqrt(rr) / (sqrt(nu) + cabs(lam) * sqrt(nv));   // relative to the terms of the pencil at this mode
    cta_sync();
    if (resid < tol) { ++it; break; }
  }
  // scale like temporal.f90:867-879: first entry of maximum modulus becomes 1
  double best = -1.0; int bi = 0;
  for (int r = c.tid; r < n; r += c.nt) { const double m = cabs(x[r]); if (m > best) { best = m; bi = r; } }
  cta_argmax(c, best, bi);
  const cplx sc = x[bi];
  cta_sync();
  for (int r = c.tid; r < n; r += c.nt) x[r] = cdiv(x[r], sc);
  if (c.tid == 0) { out4[0] = lam.re; out4[1] = lam.im; out4[2] = resid; out4[3] = (double)(sing ? -sing : it); }
  cta_sync();
}

}  // namespace stab
