// hessenberg.cuh -- CTA-cooperative unitary reduction of a balanced complex matrix to upper
// Hessenberg form, H = Q^H A Q with Q = H(ilo) H(ilo+1) ... H(ihi-1), H(j) = I - tau_j v_j v_j^H.
// Output layout is LAPACK's (ZGEHRD): H in the upper triangle + first subdiagonal, the
// Householder vectors below it, tau[] separately.  This is the stage ZGEEV spends its ZGEHRD
// time in (reference call site temporal.f90:803 / spatial.f90:1043).
//
// v1 ("streamed rank-2"): one CTA per matrix; per column j
//   pass 1  y = A(0:ihi, j+1:ihi) v                      (thread per row, coalesced column walk)
//   pass 2  per column c (one warp each): right update with y, dot with v, left update  -> the
//           column is read once from HBM/L2 and written once.
// Algorithmic work (40/3) n^3 real flops; traffic ~ 3 * 16 * n^3 / 3 bytes from L2/HBM.
#pragma once
#include "common.cuh"

namespace stab {

// Householder generator on (alpha, x[0:m-1)) following ZLARFG: returns tau, overwrites alpha by
// beta (real) and x by v(2:m).  All threads call it; x lives in global memory column `xp`.
SD_DEV cplx cta_zlarfg(const Cta& c, int m, cplx& alpha, cplx* xp) {
  if (m <= 0) return mk(0.0, 0.0);
  double ss = 0.0;
  for (int r = c.tid; r < m - 1; r += c.nt) ss += abs2(xp[r]);
  ss = cta_sum(c, ss);
  double xnorm = sqrt(ss);
  if (xnorm == 0.0 && alpha.im == 0.0) return mk(0.0, 0.0);
  double beta = -copysign(sqrt(alpha.re * alpha.re + alpha.im * alpha.im + ss), alpha.re);
  // (the safmin rescaling loop of ZLARFG is only reachable for |beta| < ~1e-292; operators here
  //  are O(1)...O(1e9), so it is omitted)
  cplx tau = mk((beta - alpha.re) / beta, -alpha.im / beta);
  cplx sc = cdiv(mk(1.0, 0.0), mk(alpha.re - beta, alpha.im));
  for (int r = c.tid; r < m - 1; r += c.nt) xp[r] = xp[r] * sc;
  alpha = mk(beta, 0.0);
  return tau;
}

// sv, sy: shared vectors of n complex each.
SD_DEV void cta_hessenberg(const Cta& c, cplx* A, int n, int lda, int ilo, int ihi, cplx* tau, cplx* sv, cplx* sy) {
  for (int j = c.tid; j < n; j += c.nt) tau[j] = mk(0.0, 0.0);
  cta_sync();
  for (int j = ilo; j < ihi; ++j) {
    const int m = ihi - j;                    // reflector length, acts on rows/cols j+1..ihi
    cplx* col = A + (size_t)j * lda;
    cplx alpha = col[j + 1];
    cta_sync();
    cplx tj = cta_zlarfg(c, m, alpha, col + j + 2);
    cta_sync();
    if (c.tid == 0) { tau[j] = tj; col[j + 1] = alpha; }
    if (is_zero(tj)) { cta_sync(); continue; }
    // v into shared (v[0] = 1 at row j+1)
    for (int r = c.tid; r < m; r += c.nt) sv[r] = (r == 0) ? mk(1.0, 0.0) : col[j + 1 + r];
    cta_sync();
    // pass 1: y(r) = sum_c A(r, j+1+c) v(c), r in [0, ihi]
    for (int r = c.tid; r <= ihi; r += c.nt) {
      cplx acc0 = mk(0.0, 0.0), acc1 = mk(0.0, 0.0);
      const cplx* ap = A + r + (size_t)(j + 1) * lda;
      int cc = 0;
      for (; cc + 1 < m; cc += 2) {
        fma_acc(acc0, ap[(size_t)cc * lda], sv[cc]);
        fma_acc(acc1, ap[(size_t)(cc + 1) * lda], sv[cc + 1]);
      }
      if (cc < m) fma_acc(acc0, ap[(size_t)cc * lda], sv[cc]);
      sy[r] = (acc0 + acc1) * tj;               // pre-scaled: tau * y
    }
    cta_sync();
    // pass 2: one warp per column
    const cplx ctj = conj(tj);
    for (int cidx = j + 1 + c.wid; cidx < n; cidx += c.nw) {
      cplx* ac = A + (size_t)cidx * lda;
      cplx w = mk(0.0, 0.0);
      if (cidx <= ihi) {
        const cplx vc = conj(sv[cidx - j - 1]);
        for (int r = c.lane; r <= ihi; r += c.ws) {
          cplx a = ac[r] - sy[r] * vc;
          ac[r] = a;
          if (r > j) fma_acc_conj(w, sv[r - j - 1], a);
        }
      } else {
        for (int r = j + 1 + c.lane; r <= ihi; r += c.ws) fma_acc_conj(w, sv[r - j - 1], ac[r]);
      }
      w = warp_sum(w);
      w = ctj * w;
      warp_sync();
      for (int r = j + 1 + c.lane; r <= ihi; r += c.ws) ac[r] = ac[r] - sv[r - j - 1] * w;
    }
    cta_sync();
  }
}

}  // namespace stab
