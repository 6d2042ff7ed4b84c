// evec.cuh -- right eigenvectors by shift-invert inverse iteration on the Hessenberg matrix
// (one warp per eigenvalue), back-transformation with the stored Householder reflectors,
// undoing of the balancing, and the normalisations of ZGEEV and temporal.f90.
//
// Reference behaviour: ZGEEV('N','V') returns unit-2-norm right eigenvectors whose largest
// component is real (temporal.f90:803, spatial.f90:1043); temporal.f90:867-879 then divides each
// column by its first max-|.| entry.  LAPACK reaches the vectors through the Schur form
// (ZHSEQR 'S' + ZTREVC); we keep the QR sweep eigenvalue-only and get each vector from ONE
// triangular solve with (H - lambda I) -- the ZHSEIN/ZLAEIN idea, and the "shift-invert inverse
// iteration" stage the north star names -- which is embarrassingly parallel over eigenvalues.
//
// Elimination order: H - lambda I is reduced to upper triangular U by COLUMN operations from the
// bottom row up (with column interchanges).  Columns of U become final in the order n-1, n-2, ...
// which is exactly the order a column-oriented back substitution consumes them, so U is never
// stored: each Hessenberg column is read once, combined with one carried column, used, dropped.
#pragma once
#include "common.cuh"

namespace stab {

// One warp.  On return y holds x with (H - lam I) x ~ b, b = bscale * start vector `variant`.
// c, y: warp-private shared vectors of n complex; flag: n bytes.
SD_DEV void warp_hess_solve(const Cta& w, const cplx* H, int n, int ldh, cplx lam, double eps3, double bscale,
                            int variant, cplx* c, cplx* y, unsigned char* flag) {
  const double rootn = sqrt((double)n);
  for (int r = w.lane; r < n; r += w.ws) {
    cplx a = H[r + (size_t)(n - 1) * ldh];
    if (r == n - 1) a -= lam;
    c[r] = a;
    // ZLAEIN's start vectors: all ones, or its "orthogonal" retry vectors
    double b = bscale;
    if (variant > 0) {
      b = bscale / (rootn + 1.0);
      if (r == 0) b = bscale;
      if (r == n - variant) b -= bscale * rootn;
    }
    y[r] = mk(b, 0.0);
  }
  warp_sync();
  for (int k = n - 1; k >= 1; --k) {
    const cplx* acol = H + (size_t)(k - 1) * ldh;
    const cplx ck = c[k];
    const cplx ak = acol[k];
    const bool sw = cabs1(ak) > cabs1(ck);
    cplx piv = sw ? ak : ck;
    if (is_zero(piv)) piv = mk(eps3, 0.0);
    const cplx yk = cdiv(y[k], piv);
    const cplx m = cdiv(sw ? ck : ak, piv);
    warp_sync();
    for (int r = w.lane; r < k; r += w.ws) {
      cplx a = acol[r];
      if (r == k - 1) a -= lam;
      const cplx cr = c[r];
      const cplx u = sw ? a : cr;
      c[r] = sw ? (cr - m * a) : (a - m * cr);
      y[r] -= yk * u;
    }
    if (w.lane == 0) { c[k] = m; y[k] = yk; flag[k] = sw ? 1 : 0; }
    warp_sync();
  }
  if (w.lane == 0) {
    cplx piv = c[0];
    if (is_zero(piv)) piv = mk(eps3, 0.0);
    y[0] = cdiv(y[0], piv);
    // x = T_{n-1} ... T_1 y
    for (int k = 1; k < n; ++k) {
      cplx yk = y[k] - c[k] * y[k - 1];
      if (flag[k]) { y[k] = y[k - 1]; y[k - 1] = yk; } else { y[k] = yk; }
    }
  }
  warp_sync();
}

// x := Q x with Q = H(ilo) ... H(ihi-1); reflector j stored in A(j+2:ihi, j), tau[j]
SD_DEV void warp_apply_q(const Cta& w, const cplx* A, int lda, int ilo, int ihi, const cplx* tau, cplx* x) {
  for (int j = ihi - 1; j >= ilo; --j) {
    const cplx tj = tau[j];
    if (is_zero(tj)) continue;
    const cplx* v = A + (size_t)j * lda;       // v(j+1) = 1 implicit, v(r) = A(r,j) for r >= j+2
    cplx s = mk(0.0, 0.0);
    for (int r = j + 1 + w.lane; r <= ihi; r += w.ws) {
      cplx vr = (r == j + 1) ? mk(1.0, 0.0) : v[r];
      fma_acc_conj(s, vr, x[r]);
    }
    s = warp_sum(s);
    s = tj * s;
    for (int r = j + 1 + w.lane; r <= ihi; r += w.ws) {
      cplx vr = (r == j + 1) ? mk(1.0, 0.0) : v[r];
      x[r] -= vr * s;
    }
    warp_sync();
  }
}

// ZGEBAK('B','R'): undo scaling on [ilo,ihi], then the permutations
SD_DEV void warp_gebak(const Cta& w, int n, int ilo, int ihi, const double* scale, cplx* x) {
  if (ilo != ihi)                             // ZGEBAK skips the scaling when ILO == IHI (scale(ilo) may hold a permutation)
    for (int r = ilo + w.lane; r <= ihi; r += w.ws) x[r] = x[r] * scale[r];
  warp_sync();
  if (w.lane == 0) {
    for (int i = ilo - 1; i >= 0; --i) {
      int k = (int)scale[i];
      if (k != i) { cplx t = x[i]; x[i] = x[k]; x[k] = t; }
    }
    for (int i = ihi + 1; i < n; ++i) {
      int k = (int)scale[i];
      if (k != i) { cplx t = x[i]; x[i] = x[k]; x[k] = t; }
    }
  }
  warp_sync();
}

// ZGEEV's normalisation: unit 2-norm, component of largest modulus made real (first maximum)
SD_DEV void warp_normalize_zgeev(const Cta& w, int n, cplx* x) {
  // pre-scale by the largest |re|+|im| so the squares can neither overflow nor underflow
  double mx = 0.0;
  for (int r = w.lane; r < n; r += w.ws) mx = fmax(mx, cabs1(x[r]));
  mx = warp_max(mx);
  if (mx == 0.0) return;
  const double pre = 1.0 / mx;
  double ss = 0.0;
  for (int r = w.lane; r < n; r += w.ws) { cplx z = x[r] * pre; ss += abs2(z); }
  ss = warp_sum(ss);
  double scl = pre / sqrt(ss);
  double best = -1.0; int bi = 0;
  for (int r = w.lane; r < n; r += w.ws) {
    cplx z = x[r] * scl; x[r] = z;
    double m2 = z.re * z.re + z.im * z.im;
    if (m2 > best) { best = m2; bi = r; }
  }
#ifndef STAB_EMU
  for (int o = 16; o > 0; o >>= 1) {
    double ov = __shfl_xor_sync(0xffffffffu, best, o);
    int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
#endif
  warp_sync();
  cplx xk = x[bi];
  double rt = sqrt(best);
  cplx tmp = mk(xk.re / rt, -xk.im / rt);
  warp_sync();
  for (int r = w.lane; r < n; r += w.ws) {
    cplx z = x[r] * tmp;
    if (r == bi) z.im = 0.0;
    x[r] = z;
  }
  warp_sync();
}

// temporal.f90:867-879: divide by the first entry of maximum |.| (strict >), rows [0, nrows)
SD_DEV void warp_scale_maxabs(const Cta& w, int nrows, cplx* x) {
  double best = 0.0; int bi = -1;
  for (int r = w.lane; r < nrows; r += w.ws) {
    double m = cabs(x[r]);
    if (m > best) { best = m; bi = r; }
  }
#ifndef STAB_EMU
  for (int o = 16; o > 0; o >>= 1) {
    double ov = __shfl_xor_sync(0xffffffffu, best, o);
    int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (oi >= 0 && (ov > best || (ov == best && (bi < 0 || oi < bi)))) { best = ov; bi = oi; }
  }
#endif
  warp_sync();
  if (bi < 0) return;
  const cplx sc = x[bi];
  warp_sync();
  for (int r = w.lane; r < nrows; r += w.ws) x[r] = cdiv(x[r], sc);
  warp_sync();
}

// Full per-eigenvalue pipeline executed by one warp.  Returns 0 or 1 (no growth / non-finite).
//   Hh  : Hessenberg matrix with the reflectors still stored below the subdiagonal (ZGEHRD layout);
//         entries below the first subdiagonal are ignored by construction of the solver.
//   kr  : last index of the diagonal block of H (between exactly-zero subdiagonals) the eigenvalue
//         belongs to.  As ZHSEIN does, the vector is computed from the leading (kr+1) x (kr+1)
//         submatrix and is exactly zero below: a solve with the whole matrix would pick up
//         cond(H22 - lam I) * eps garbage in the decoupled trailing part.
SD_DEV int warp_eigvec(const Cta& w, const cplx* Hh, int n, int ldh, int ilo, int ihi, const cplx* tau,
                       const double* scale, cplx lam, int kr, double hnorm, int scale_rows, cplx* c, cplx* y,
                       unsigned char* flag, cplx* out, bool raw = false) {
  const int m = kr + 1;
  const double smlnum = SD_SAFMIN * ((double)n / SD_ULP);
  const double eps3 = fmax(SD_ULP * hnorm, smlnum);
  const double rootn = sqrt((double)n);
  const double growto = 0.1 / rootn;
  const double bscale = eps3;                 // |b_i| ~ eps3 as in ZLAEIN
  int bad = 1;
  for (int its = 0; its < 4; ++its) {
    warp_hess_solve(w, Hh, m, ldh, lam, eps3, bscale, its, c, y, flag);
    double vn = 0.0;
    for (int r = w.lane; r < m; r += w.ws) vn += cabs1(y[r]);
    vn = warp_sum(vn);
    if (!(vn == vn) || vn > 1.0e300) break;   // NaN / overflow
    if (vn >= growto) { bad = 0; break; }                 // ZLAEIN's growth test
  }
  if (bad) {
    // no acceptable vector: return the unit vector of the last attempt's largest entry
    for (int r = w.lane; r < m; r += w.ws) y[r] = mk(r == 0 ? 1.0 : 0.0, 0.0);
    warp_sync();
  }
  for (int r = m + w.lane; r < n; r += w.ws) y[r] = mk(0.0, 0.0);
  warp_sync();
  if (!raw) {                                 // raw: leave the vector in the Hessenberg basis (GEMM back-transformation follows)
    warp_apply_q(w, Hh, ldh, ilo, ihi, tau, y);
    warp_gebak(w, n, ilo, ihi, scale, y);
    warp_normalize_zgeev(w, n, y);
    if (scale_rows > 0) warp_scale_maxabs(w, scale_rows, y);
  }
  for (int r = w.lane; r < n; r += w.ws) out[r] = y[r];
  warp_sync();
  return bad;
}

}  // namespace stab
