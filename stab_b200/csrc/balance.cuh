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
// wsp: double workspace of balance_wsp_doubles(n, bal_b) entries (shared memory on the device).
// row stride of the staged row block: n rounded up to 2 (mod 16) doubles, so that the bal_b rows a half warp stores at once
// (index e fastest) fall into different shared-memory banks (stride n = 640 put all eight into ONE bank: 8-way replays)
SD_HD int balance_rstride(int n) { return n + ((2 - (n & 15)) & 15); }
SD_HD size_t balance_wsp_doubles(int n, int bal_b) { return (size_t)2 * n + (size_t)bal_b * n + (size_t)bal_b * balance_rstride(n) + 2 + 4 * (size_t)bal_b; }
SD_DEV void cta_balance(const Cta& c, cplx* A, int n, int lda, double* scale, int* cnt, double* wsp, int bal_b, int& ilo_out, int& ihi_out) {
  int k = 0;      // first active index
  int l = n;      // one past last active index
  // ---- row isolation: push rows with zero off-diagonal part (within columns [0,l)) down ----
  for (int r = c.tid; r < n; r += c.nt) {       // eight loads in flight per thread: the scan is latency bound
    int m = 0;
    for (int j0 = 0; j0 < n; j0 += 8) {
      cplx v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = (j0 + u < n) ? A[r + (size_t)(j0 + u) * lda] : mk(0.0, 0.0);
#pragma unroll
      for (int u = 0; u < 8; ++u) if (j0 + u != r && !is_zero(v[u])) ++m;
    }
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
    for (int r0 = k; r0 < l; r0 += 8) {
      cplx v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = (r0 + u < l) ? A[(r0 + u) + (size_t)j * lda] : mk(0.0, 0.0);
#pragma unroll
      for (int u = 0; u < 8; ++u) if (r0 + u != j && !is_zero(v[u])) ++m;
    }
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
  // ZGEBAL's loop is Gauss-Seidel over i (each decision sees every scaling made before it).
  // Scaling by powers of two is exact, so the loop runs on the UNSCALED matrix: |a|^2 of bal_b
  // columns / rows at a time is staged in shared memory with coalesced, batched loads, one warp
  // takes the bal_b decisions in order, weighting every entry with the cumulative factors chosen
  // so far (fs = scale, fi = 1/scale), and the matrix itself is rescaled once at the end by a
  // single coalesced pass.  Same decisions as the entry-by-entry loop, without two block
  // reductions, three barriers and a strided row pass per index.
  double* fs = wsp;                       // n: cumulative column factor (1 outside [k,l))
  double* fi = wsp + n;                   // n: its exact inverse (row factor)
  double* cbuf = wsp + 2 * n;             // bal_b x n : |A(r, i_e)|^2
  double* rbuf = cbuf + (size_t)bal_b * n;   // bal_b x rst : |A(i_e, j)|^2
  const int rst = balance_rstride(n);
  int* flag = reinterpret_cast<int*>(rbuf + (size_t)bal_b * rst);
  double* part = rbuf + (size_t)bal_b * rst + 2;   // 4 x bal_b: partial sums / maxima of the block's indices
  int lb = 0; while ((1 << (lb + 1)) <= bal_b) ++lb;   // bal_b is a power of two
  const int B = 1 << lb;
  for (int i = c.tid; i < n; i += c.nt) { fs[i] = 1.0; fi[i] = 1.0; }
  cta_sync();
  const double radix = 2.0, factor = 0.95;
  const double sfmin1 = SD_SAFMIN / SD_ULP, sfmax1 = 1.0 / sfmin1;
  const double sfmin2 = sfmin1 * 2.0, sfmax2 = 1.0 / sfmin2;
  constexpr int U = 8;                    // loads in flight per thread
  noconv = true;
  int guard = 0;
  while (noconv && guard++ < 200) {
    if (c.tid == 0) *flag = 0;
    for (int i0 = k; i0 < l; i0 += B) {
      const int nb = (l - i0 < B) ? (l - i0) : B;
      // stage |a|^2: columns i0..i0+nb-1 (rows 0..l-1), rows i0..i0+nb-1 (columns k..n-1)
      for (int r = c.tid; r < l; r += c.nt) {
        double v[U];
#pragma unroll
        for (int e = 0; e < U; ++e) v[e] = (e < nb) ? abs2(A[r + (size_t)(i0 + e) * lda]) : 0.0;
#pragma unroll
        for (int e = 0; e < U; ++e) if (e < nb) cbuf[(size_t)e * n + r] = v[e];
      }
      const int totr = (n - k) << lb;
      for (int q0 = c.tid; q0 < totr; q0 += c.nt * U) {
        double v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int q = q0 + u * c.nt, e = q & (B - 1), j = k + (q >> lb);   // e fastest: B consecutive rows of one column
          v[u] = (q < totr && e < nb) ? abs2(A[(i0 + e) + (size_t)j * lda]) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int q = q0 + u * c.nt, e = q & (B - 1), j = k + (q >> lb);
          if (q < totr && e < nb) rbuf[(size_t)e * rst + j] = v[u];
        }
      }
      cta_sync();
#ifndef STAB_EMU
      {
        // The staging loads of the NEXT block are pure latency (HBM round trips with nothing to overlap): pull its
        // 128-byte lines into L2 now, under this block's decisions.  No registers, no shared memory.
        const int i1 = (i0 + B < l) ? i0 + B : k;             // the first block of the next sweep after the last one
        const int nb1 = (l - i1 < B) ? (l - i1) : B;
        const int lines_c = (l * 16 + 127) >> 7;              // lines per column (rows 0..l-1)
        for (int q = c.tid; q < nb1 * lines_c; q += c.nt) {
          const int e = q / lines_c, ln = q - e * lines_c;
          asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(A + (size_t)(i1 + e) * lda) + ((size_t)ln << 7)));
        }
        for (int j = k + c.tid; j < n; j += c.nt) {           // rows i1..i1+nb1-1 of column j: one or two lines
          const char* q = reinterpret_cast<const char*>(A + i1 + (size_t)j * lda);
          asm volatile("prefetch.global.L2 [%0];" ::"l"(q));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(q + (nb1 - 1) * 16));
        }
      }
#endif
      // (a) every index of the block in parallel, one warp each: its sums and maxima over the rows / columns OUTSIDE the
      //     block, whose factors cannot change while the block is decided
      for (int e = c.wid; e < nb; e += c.nw) {
        const double* cb = cbuf + (size_t)e * n;
        const double* rb = rbuf + (size_t)e * rst;
        double cs = 0.0, rs = 0.0, cam = 0.0, ram = 0.0;
        for (int r = c.lane; r < l; r += c.ws) {
          if (r >= i0 && r < i0 + nb) continue;
          const double w = fi[r];                           // row factors chosen so far (1 for r < k)
          const double v = cb[r] * (w * w);
          if (r >= k) cs += v;
          cam = fmax(cam, v);
        }
        for (int j = k + c.lane; j < n; j += c.ws) {
          if (j >= i0 && j < i0 + nb) continue;
          const double w = fs[j];                           // column factors chosen so far (1 for j >= l)
          const double v = rb[j] * (w * w);
          if (j < l) rs += v;
          ram = fmax(ram, v);
        }
        cs = warp_sum(cs); rs = warp_sum(rs); cam = warp_max(cam); ram = warp_max(ram);
        if (c.lane == 0) { part[4 * e] = cs; part[4 * e + 1] = rs; part[4 * e + 2] = cam; part[4 * e + 3] = ram; }
      }
      cta_sync();
      // (b) the decisions in order (Gauss-Seidel): only the nb in-block terms are weighted with the factors of the moment
      if (c.wid == 0) {
        for (int e = 0; e < nb; ++e) {
          const int i = i0 + e;
          const double* cb = cbuf + (size_t)e * n;
          const double* rb = rbuf + (size_t)e * rst;
          // column i: c = ||A(k:l, i)||_2, ca = max |A(0:l, i)| ; row i: r = ||A(i, k:l)||_2, ra = max |A(i, k:n)|
          double cs = 0.0, rs = 0.0, cam = 0.0, ram = 0.0;
          for (int q = c.lane; q < nb; q += c.ws) {
            const int r = i0 + q;                             // k <= r < l
            double w = fi[r];
            double v = cb[r] * (w * w);
            cs += v; cam = fmax(cam, v);
            w = fs[r];
            v = rb[r] * (w * w);
            rs += v; ram = fmax(ram, v);
          }
          cs = warp_sum(cs); rs = warp_sum(rs); cam = warp_max(cam); ram = warp_max(ram);
          cs += part[4 * e]; rs += part[4 * e + 1]; cam = fmax(cam, part[4 * e + 2]); ram = fmax(ram, part[4 * e + 3]);
          const double sc = fs[i], isc = fi[i];
          cs *= sc * sc; cam *= sc * sc; rs *= isc * isc; ram *= isc * isc;   // this index's own factors (exact)
          double ca = sqrt(cam), ra = sqrt(ram);
          double cn = sqrt(cs), rn = sqrt(rs);
          if (cn == 0.0 || rn == 0.0) continue;
          double g = rn / radix, f = 1.0, sum = cn + rn;
          while (cn < g && fmax(f, fmax(cn, ca)) < sfmax2 && fmin(rn, fmin(g, ra)) > sfmin2) {
            f *= radix; cn *= radix; ca *= radix; rn /= radix; g /= radix; ra /= radix;
          }
          g = cn / radix;
          while (g >= rn && fmax(rn, ra) < sfmax2 && fmin(fmin(f, cn), fmin(g, ca)) > sfmin2) {
            f /= radix; cn /= radix; g /= radix; ca /= radix; rn *= radix; ra *= radix;
          }
          if ((cn + rn) >= factor * sum) continue;
          if (f < 1.0 && sc < 1.0 && f * sc <= sfmin1) continue;
          if (f > 1.0 && sc > 1.0 && sc >= sfmax1 / f) continue;
          warp_sync();
          if (c.lane == 0) { fs[i] = sc * f; fi[i] = isc / f; *flag = 1; }
          warp_sync();
        }
      }
      cta_sync();
    }
    noconv = (*flag != 0);
#ifdef STAB_EMU_TRACE
    fprintf(stderr, "balance sweep %d k=%d l=%d noconv=%d\n", guard, k, l, (int)noconv);
#endif
    cta_sync();
  }
  // apply: column j in [k,l) scaled by fs[j] over rows 0..l-1, row r in [k,l) by fi[r] over columns k..n-1
  for (int i = k + c.tid; i < l; i += c.nt) scale[i] = fs[i];
  for (int j0 = k; j0 < n; j0 += 4) {
    for (int r = c.tid; r < l; r += c.nt) {
      const double fr = (r >= k) ? fi[r] : 1.0;
      cplx v[4]; double w[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = j0 + u;
        w[u] = (j < n) ? ((j < l) ? fs[j] : 1.0) * fr : 1.0;
        if (j < n && w[u] != 1.0) v[u] = A[r + (size_t)j * lda];
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = j0 + u;
        if (j < n && w[u] != 1.0) A[r + (size_t)j * lda] = v[u] * w[u];
      }
    }
  }
  cta_sync();
  ilo_out = k;
  ihi_out = l - 1;
}

}  // namespace stab
