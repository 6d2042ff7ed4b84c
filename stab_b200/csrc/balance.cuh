// balance.cuh -- CTA-cooperative balancing of a general complex matrix, the preprocessing ZGEEV
// applies before the Hessenberg reduction (ZGEBAL job 'B': permute to isolate eigenvalues, then
// diagonal power-of-two scaling of the remaining block).  The reference reaches it through
// ZGEEV (temporal.f90:803, spatial.f90:1043); reproducing it keeps our rounding path inside
// LAPACK's noise floor on these badly scaled, highly non-normal operators (hard part 4).
//
// Algorithm: LAPACK 3.11+ ZGEBAL semantics (search order of the permutation loops, 2-norm
// based scaling test, radix 2), parallelised inside one CTA:
//   * isolation search keeps per-row / per-column non-zero counts that are updated incrementally
//     (O(n) per isolated eigenvalue instead of an O(n^2) rescan),
//   * the Gauss-Seidel scaling loop is sequential over i by definition; each step is two fused
//     block reductions (column: nrm2 + max, row: nrm2 + max) and a row/column rescale.
// Output convention (0-based): active block is [ilo, ihi] inclusive; scale[j] holds the
// 0-based permutation partner for j outside the block and the scaling factor inside.
#pragma once
#include "common.cuh"

namespace stab {

SD_DEV void cta_swap_cols(const Cta& c, cplx* A, int lda, int a, int b, int nrows) {
  if (a == b) return;
  for (int r = c.tid; r < nrows; r += c.nt) {
    cplx t = A[r + (size_t)a * lda]; A[r + (size_t)a * lda] = A[r + (size_t)b * lda]; A[r + (size_t)b * lda] = t;
  }
}
SD_DEV void cta_swap_rows(const Cta& c, cplx* A, int lda, int a, int b, int c0, int c1) {
  if (a == b) return;
  for (int j = c0 + c.tid; j < c1; j += c.nt) {
    cplx t = A[a + (size_t)j * lda]; A[a + (size_t)j * lda] = A[b + (size_t)j * lda]; A[b + (size_t)j * lda] = t;
  }
}

// Fused block reduction for the scaling loop: sums of cs, rs; (max, first index) of (cam, cai) and (ram, rai).
SD_DEV void cta_bal_reduce(const Cta& c, double& cs, double& rs, double& cam, int& cai, double& ram, int& rai) {
#ifdef STAB_EMU
  (void)c; (void)cs; (void)rs; (void)cam; (void)cai; (void)ram; (void)rai;
#else
  for (int o = 16; o > 0; o >>= 1) {
    cs += __shfl_xor_sync(0xffffffffu, cs, o);
    rs += __shfl_xor_sync(0xffffffffu, rs, o);
    double ov = __shfl_xor_sync(0xffffffffu, cam, o); int oi = __shfl_xor_sync(0xffffffffu, cai, o);
    if (ov > cam || (ov == cam && oi < cai)) { cam = ov; cai = oi; }
    ov = __shfl_xor_sync(0xffffffffu, ram, o); oi = __shfl_xor_sync(0xffffffffu, rai, o);
    if (ov > ram || (ov == ram && oi < rai)) { ram = ov; rai = oi; }
  }
  double* rd = c.red;                       // [nw][4] doubles + [nw][2] ints
  int* ri = reinterpret_cast<int*>(c.red + 4 * c.nw);
  cta_sync();
  if (c.lane == 0) {
    rd[4 * c.wid] = cs; rd[4 * c.wid + 1] = rs; rd[4 * c.wid + 2] = cam; rd[4 * c.wid + 3] = ram;
    ri[2 * c.wid] = cai; ri[2 * c.wid + 1] = rai;
  }
  cta_sync();
  double s0 = 0.0, s1 = 0.0, m0 = rd[2], m1 = rd[3]; int i0 = ri[0], i1 = ri[1];
  for (int w = 0; w < c.nw; ++w) {
    s0 += rd[4 * w]; s1 += rd[4 * w + 1];
    if (w > 0) {
      double ov = rd[4 * w + 2]; int oi = ri[2 * w];
      if (ov > m0 || (ov == m0 && oi < i0)) { m0 = ov; i0 = oi; }
      ov = rd[4 * w + 3]; oi = ri[2 * w + 1];
      if (ov > m1 || (ov == m1 && oi < i1)) { m1 = ov; i1 = oi; }
    }
  }
  cs = s0; rs = s1; cam = m0; cai = i0; ram = m1; rai = i1;
#endif
}

// cnt: int workspace of n entries (global or shared).  Returns ilo/ihi through pointers
// (every thread gets the same values).
SD_DEV void cta_balance(const Cta& c, cplx* A, int n, int lda, double* scale, int* cnt, int& ilo_out, int& ihi_out) {
  int k = 0;      // first active index
  int l = n;      // one past last active index
  // ---- row isolation: push rows with zero off-diagonal part (within columns [0,l)) down ----
  for (int r = c.tid; r < n; r += c.nt) {
    int m = 0;
    for (int j = 0; j < n; ++j)
      if (j != r && !is_zero(A[r + (size_t)j * lda])) ++m;
    cnt[r] = m;
  }
  cta_sync();
  bool done = false;
  bool noconv = true;
  while (noconv && !done) {
    noconv = false;
    int ip = l - 1;                       // DO I = L, 1, -1 (bound fixed at loop entry)
    while (ip >= 0) {
      int best = -1;
      for (int r = c.tid; r <= ip; r += c.nt)
        if (cnt[r] == 0) best = r;        // ascending scan: last hit is the largest
      best = cta_max_i(c, best);
      if (best < 0) break;
      const int i = best;
      if (c.tid == 0) scale[l - 1] = (double)i;
      if (i != l - 1) {
        cta_swap_cols(c, A, lda, i, l - 1, l);
        cta_sync();
        cta_swap_rows(c, A, lda, i, l - 1, k, n);
        if (c.tid == 0) { int t = cnt[i]; cnt[i] = cnt[l - 1]; cnt[l - 1] = t; }
      }
      cta_sync();
      noconv = true;
      if (l == 1) { done = true; break; }
      l = l - 1;
      // column l is now excluded from the search range
      for (int r = c.tid; r < l; r += c.nt)
        if (!is_zero(A[r + (size_t)l * lda])) cnt[r] -= 1;
      cta_sync();
      ip = i - 1;
    }
  }
  if (done) { ilo_out = 0; ihi_out = 0; return; }
  // ---- column isolation: push columns with zero off-diagonal part (rows [k,l)) left ----
  for (int j = c.tid; j < l; j += c.nt) {
    int m = 0;
    for (int r = k; r < l; ++r)
      if (r != j && !is_zero(A[r + (size_t)j * lda])) ++m;
    cnt[j] = m;
  }
  cta_sync();
  noconv = true;
  while (noconv) {
    noconv = false;
    int jp = k;                           // DO J = K, L
    const int lfix = l;
    while (jp < lfix) {
      int best = n;
      for (int j = jp + c.tid; j < lfix; j += c.nt)
        if (cnt[j] == 0 && j < best) best = j;
      best = cta_min_i(c, best);
      if (best >= n) break;
      const int j = best;
      if (c.tid == 0) scale[k] = (double)j;
      if (j != k) {
        cta_swap_cols(c, A, lda, j, k, l);
        cta_sync();
        cta_swap_rows(c, A, lda, j, k, k, n);
        if (c.tid == 0) { int t = cnt[j]; cnt[j] = cnt[k]; cnt[k] = t; }
      }
      cta_sync();
      noconv = true;
      // row k leaves the search range
      for (int jj = k + 1 + c.tid; jj < l; jj += c.nt)
        if (!is_zero(A[k + (size_t)jj * lda])) cnt[jj] -= 1;
      k = k + 1;
      cta_sync();
      jp = j + 1;
    }
  }
  // ---- scaling of the active block [k, l) ----
  for (int i = k + c.tid; i < l; i += c.nt) scale[i] = 1.0;
  cta_sync();
  const double radix = 2.0, factor = 0.95;
  const double sfmin1 = SD_SAFMIN / SD_ULP, sfmax1 = 1.0 / sfmin1;
  const double sfmin2 = sfmin1 * 2.0, sfmax2 = 1.0 / sfmin2;
  noconv = true;
  int guard = 0;
  while (noconv && guard++ < 200) {
    noconv = false;
    for (int i = k; i < l; ++i) {
      // column i: c = ||A(k:l, i)||_2, ca = max |A(0:l, i)| ; row i: r = ||A(i, k:l)||_2, ra = max |A(i, k:n)|
      double cs = 0.0, rs = 0.0, ca = 0.0, ra = 0.0;
      // IZAMAX picks by |re|+|im| but the value used is the true modulus of that entry; the two
      // orderings can differ, so track (cabs1, index) then evaluate |.| of the winner.
      double cam = -1.0, ram = -1.0; int cai = 0, rai = 0;
      for (int r = c.tid; r < l; r += c.nt) {
        cplx a = A[r + (size_t)i * lda];
        if (r >= k) cs += abs2(a);
        double m1 = cabs1(a);
        if (m1 > cam) { cam = m1; cai = r; }
      }
      for (int j = k + c.tid; j < n; j += c.nt) {
        cplx a = A[i + (size_t)j * lda];
        if (j < l) rs += abs2(a);
        double m1 = cabs1(a);
        if (m1 > ram) { ram = m1; rai = j; }
      }
      cta_bal_reduce(c, cs, rs, cam, cai, ram, rai);      // one fused reduction (two barriers) instead of three
      ca = cabs(A[cai + (size_t)i * lda]);
      ra = cabs(A[i + (size_t)rai * lda]);
      double cn = sqrt(cs), rn = sqrt(rs);
      if (cn == 0.0 || rn == 0.0) continue;
      double g = rn / radix, f = 1.0, s = cn + rn;
      while (cn < g && fmax(f, fmax(cn, ca)) < sfmax2 && fmin(rn, fmin(g, ra)) > sfmin2) {
        f *= radix; cn *= radix; ca *= radix; rn /= radix; g /= radix; ra /= radix;
      }
      g = cn / radix;
      while (g >= rn && fmax(rn, ra) < sfmax2 && fmin(fmin(f, cn), fmin(g, ca)) > sfmin2) {
        f /= radix; cn /= radix; g /= radix; ca /= radix; rn *= radix; ra *= radix;
      }
      if ((cn + rn) >= factor * s) continue;
      double sc = scale[i];
      if (f < 1.0 && sc < 1.0 && f * sc <= sfmin1) continue;
      if (f > 1.0 && sc > 1.0 && sc >= sfmax1 / f) continue;
      g = 1.0 / f;
      noconv = true;
      cta_sync();                       // everyone has read scale[i] / pivots before we modify
      if (c.tid == 0) scale[i] = sc * f;
      for (int j = k + c.tid; j < n; j += c.nt) A[i + (size_t)j * lda] = A[i + (size_t)j * lda] * g;
      cta_sync();
      for (int r = c.tid; r < l; r += c.nt) A[r + (size_t)i * lda] = A[r + (size_t)i * lda] * f;
      cta_sync();
    }
  }
  ilo_out = k;
  ihi_out = l - 1;
}

}  // namespace stab
