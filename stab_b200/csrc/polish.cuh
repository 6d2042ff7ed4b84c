// polish.cuh -- stage (4) of the north star, batched: polish one mode per sweep point by shift-invert
// (residual) inverse iteration on the operator polynomial of that point (GPU only).
//
//     P(lambda) = M0 + lambda M1 + lambda^2 M2
//       temporal (temporal.f90:622-752):  M0 = A0, M1 = -B0, M2 = 0         A0 x = omega B0 x
//       spatial  (spatial.f90:681-1016):  M0 = C0, M1 = C1,  M2 = C2        (C0 + alpha C1 + alpha^2 C2) x = 0,
//                                         x = the bottom half of the companion eigenvector the reference computes
//
// The reference itself does not polish (its tool for that, `shoot`, is external: README.md:3-7,
// thesis/TStest/run.sh:30); getevec only selects a mode of the full spectrum (getevec.f90:154-222).  What a sweep
// needs is ONE mode per point at a fraction of a full eigensolve:
//   k_polish_form     all points: M0, M1[, M2] and K = P(sigma_p), element-wise from the node coefficients
//   lu_run(...)       batched blocked LU of K with DMMA rank-32 updates (lu_blocked.cuh), no right-hand side
//   k_polish_iterate  one CTA per point, all iterations without a host round trip:
//                       y_k = M_k x;  lambda = root of x^H P(lambda) x nearest the current value (Rayleigh functional);
//                       r = P(lambda) x;  stop when |r| <= tol (|y0| + |lambda||y1| + |lambda|^2|y2|);
//                       x <- x - K^-1 r  (Neumaier's residual inverse iteration; the first two steps use the plain
//                       inverse-iteration direction K^-1 P'(lambda) x, which does not need a good start vector);
//                     K^-1 is a REPLAY of the blocked factorization on one vector (cta_lu_replay): the panels'
//                     interchanges, unit-lower 32 x 32 solves and L21 updates in order, then the blocked U solve.
// Algorithmic work per point: (8/3) n^3 flops for the LU + per iteration 16 n^2 B per M_k GEMV and 16 n^2 B for the
// replay (HBM bound).
#pragma once
#include "common.cuh"
#include "assemble.cuh"
#include "lu_blocked.cuh"

#ifndef STAB_EMU
namespace stab {

struct PolishBatch {
  int n, kind;                    // kind 1 temporal, 2 spatial
  cplx* M0; cplx* M1; cplx* M2;   // n x n per point (M2 unused for kind 1)
  cplx* K; size_t mstride;        // P(sigma), factored in place by lu_run
  const int* ipiv;                // n per point
  const int* info_lu;             // per point: 0 or index of the first zero pivot + 1
  const cplx* sigma;              // per point
  cplx* x;                        // n per point: start vector in, scaled eigenvector out
  double* out4;                   // per point: lambda.re, lambda.im, residual, iterations (negative: singular K)
};

// coef: the node coefficients of k_node_coef_temporal (apply_b0inv = 0) / k_node_coef_spatial; b0blk: B0's 5 x 5 blocks
__global__ void k_polish_form(GridDev g, const cplx* coef, const cplx* b0blk, PolishBatch pb) {
  const int n = pb.n, p = blockIdx.y;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)n * n) return;
  const int r = (int)(idx % n), c = (int)(idx / n);
  const int i = r / 5, e = r % 5, j = c / 5, v = c % 5;
  const cplx sg = pb.sigma[p];
  const size_t o = (size_t)p * pb.mstride + idx;
  if (pb.kind == 1) {
    const cplx* cf = coef + ((size_t)p * g.ny + i) * 75;
    const cplx a = op_element(cf, cf + 25, cf + 50, g, i, e, j, v);
    const cplx b = (i == j) ? b0blk[((size_t)p * g.ny + i) * 25 + e * 5 + v] : mk(0.0, 0.0);
    pb.M0[o] = a; pb.M1[o] = -b;
    pb.K[o] = a - sg * b;
  } else {
    const cplx* cf = coef + ((size_t)p * g.ny + i) * 150;
    const cplx c0 = op_element(cf, cf + 25, cf + 50, g, i, e, j, v);
    const int k = e * 5 + v;
    cplx b1 = cf[75 + k] * g.D1[i + (size_t)j * g.ny];
    cplx b2 = mk(0.0, 0.0);
    if (i == j) { b1 += cf[100 + k]; b2 = cf[125 + k]; }
    pb.M0[o] = c0; pb.M1[o] = -b1; pb.M2[o] = -b2;              // the reference's signs (spatial.f90:982-983)
    pb.K[o] = c0 - sg * b1 - (sg * sg) * b2;
  }
}

// v <- K^-1 v with the factors lu_run left in K: L below the diagonal in the row order of ITS panel (columns left of a
// panel are never interchanged), U on and above it, ipiv = absolute pivot rows.  v lives in shared memory.
SD_DEV void cta_lu_replay(const Cta& c, const cplx* K, int n, const int* ipiv, cplx* v) {
  for (int j0 = 0; j0 < n; j0 += LU_NB) {
    const int jb = min(LU_NB, n - j0), r0 = j0 + jb;
    if (c.tid == 0) {
      for (int s = 0; s < jb; ++s) {
        const int pr = ipiv[j0 + s];
        if (pr != j0 + s) { const cplx t = v[j0 + s]; v[j0 + s] = v[pr]; v[pr] = t; }
      }
    }
    cta_sync();
    if (c.wid == 0) {                                          // unit-lower jb x jb solve, lane = row
      cplx vi = (c.lane < jb) ? v[j0 + c.lane] : mk(0.0, 0.0);
      for (int k = 0; k < jb - 1; ++k) {
        const cplx vk = mk(__shfl_sync(0xffffffffu, vi.re, k), __shfl_sync(0xffffffffu, vi.im, k));
        if (c.lane > k && c.lane < jb) fms_acc(vi, K[(j0 + c.lane) + (size_t)(j0 + k) * n], vk);
      }
      if (c.lane < jb) v[j0 + c.lane] = vi;
    }
    cta_sync();
    for (int r = r0 + c.tid; r < n; r += c.nt) {
      cplx a = v[r];
      for (int k = 0; k < jb; ++k) fms_acc(a, K[r + (size_t)(j0 + k) * n], v[j0 + k]);
      v[r] = a;
    }
    cta_sync();
  }
  const int nblk = (n + LU_NB - 1) / LU_NB;
  for (int b = nblk - 1; b >= 0; --b) {
    const int i0 = b * LU_NB, bs = min(LU_NB, n - i0);
    if (c.wid == 0) {                                          // upper bs x bs solve, last row first
      cplx vi = (c.lane < bs) ? v[i0 + c.lane] : mk(0.0, 0.0);
      for (int k = bs - 1; k >= 0; --k) {
        cplx xk = mk(0.0, 0.0);
        if (c.lane == k) { xk = cdiv(vi, K[(i0 + k) + (size_t)(i0 + k) * n]); vi = xk; }
        xk = mk(__shfl_sync(0xffffffffu, xk.re, k), __shfl_sync(0xffffffffu, xk.im, k));
        if (c.lane < k) fms_acc(vi, K[(i0 + c.lane) + (size_t)(i0 + k) * n], xk);
      }
      if (c.lane < bs) v[i0 + c.lane] = vi;
    }
    cta_sync();
    for (int r = c.tid; r < i0; r += c.nt) {
      cplx a = v[r];
      for (int k = 0; k < bs; ++k) fms_acc(a, K[r + (size_t)(i0 + k) * n], v[i0 + k]);
      v[r] = a;
    }
    cta_sync();
  }
}

// root of a + b z + c z^2 nearest z0 (c may vanish: the linear pencil)
SD_DEV cplx nearest_root(cplx a, cplx b, cplx c, cplx z0) {
  if (cabs1(c) <= 1.0e-300 * (cabs1(a) + cabs1(b))) return -cdiv(a, b);
  const cplx disc = csqrt_(b * b - 4.0 * (a * c));
  // the two roots without cancellation: q = -(b + sgn disc)/2, z1 = q/c, z2 = a/q
  const double sg = (b.re * disc.re + b.im * disc.im) >= 0.0 ? 1.0 : -1.0;
  const cplx q = -0.5 * (b + sg * disc);
  if (is_zero(q)) return mk(0.0, 0.0);
  const cplx z1 = cdiv(q, c), z2 = cdiv(a, q);
  return (abs2(z1 - z0) <= abs2(z2 - z0)) ? z1 : z2;
}

// dynamic shared memory: 160 doubles (reductions) + 4 n complex (x, y0 | r, y1, y2)
__global__ void __launch_bounds__(512) k_polish_iterate(PolishBatch pb, int max_iters, double tol) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* red = reinterpret_cast<double*>(smem_raw);
  cplx* x = reinterpret_cast<cplx*>(smem_raw + 160 * sizeof(double));
  const int n = pb.n, p = blockIdx.x;
  cplx* y0 = x + n; cplx* y1 = y0 + n; cplx* y2 = y1 + n;
  Cta c = make_cta(red);
  const bool quad = pb.kind == 2;
  const cplx* M0 = pb.M0 + (size_t)p * pb.mstride;
  const cplx* M1 = pb.M1 + (size_t)p * pb.mstride;
  const cplx* M2 = quad ? pb.M2 + (size_t)p * pb.mstride : nullptr;
  const cplx* K = pb.K + (size_t)p * pb.mstride;
  const int* ipiv = pb.ipiv + (size_t)p * n;
  cplx* xg = pb.x + (size_t)p * n;
  double* out = pb.out4 + 4 * (size_t)p;
  if (pb.info_lu[p] != 0) {                                   // exactly singular P(sigma): sigma IS an eigenvalue to working precision
    if (c.tid == 0) { out[0] = pb.sigma[p].re; out[1] = pb.sigma[p].im; out[2] = 0.0; out[3] = -(double)pb.info_lu[p]; }
    return;
  }
  double nx = 0.0;
  for (int r = c.tid; r < n; r += c.nt) { x[r] = xg[r]; nx += abs2(xg[r]); }
  nx = cta_sum(c, nx);
  nx = (nx > 0.0 && nx == nx) ? 1.0 / sqrt(nx) : 0.0;
  for (int r = c.tid; r < n; r += c.nt) x[r] = (nx > 0.0) ? x[r] * nx : mk(1.0 / sqrt((double)n), 0.0);
  cta_sync();
  cplx lam = pb.sigma[p];
  double resid = 1.0;
  int it = 0;
  for (; it < max_iters; ++it) {
    // y_k = M_k x, one pass over the columns (x broadcast from shared memory, rows coalesced)
    cplx a = mk(0.0, 0.0), b = mk(0.0, 0.0), cc = mk(0.0, 0.0);
    double n0 = 0.0, n1 = 0.0, n2 = 0.0, dm = 0.0;
    for (int r = c.tid; r < n; r += c.nt) {
      cplx s0 = mk(0.0, 0.0), s1 = mk(0.0, 0.0), s2 = mk(0.0, 0.0);
      if (quad) {
#pragma unroll 4
        for (int j = 0; j < n; ++j) { const cplx xj = x[j]; const size_t o = r + (size_t)j * n; fma_acc(s0, M0[o], xj); fma_acc(s1, M1[o], xj); fma_acc(s2, M2[o], xj); }
      } else {
#pragma unroll 4
        for (int j = 0; j < n; ++j) { const cplx xj = x[j]; const size_t o = r + (size_t)j * n; fma_acc(s0, M0[o], xj); fma_acc(s1, M1[o], xj); }
      }
      y0[r] = s0; y1[r] = s1; y2[r] = s2;
      fma_acc_conj(a, x[r], s0); fma_acc_conj(b, x[r], s1); fma_acc_conj(cc, x[r], s2);
      n0 += abs2(s0); n1 += abs2(s1); n2 += abs2(s2);
    }
    a = cta_sum(c, a); b = cta_sum(c, b); cc = cta_sum(c, cc);
    cta_sum4(c, n0, n1, n2, dm);
    lam = nearest_root(a, b, cc, lam);
    const cplx lam2 = lam * lam;
    double rr = 0.0;
    for (int r = c.tid; r < n; r += c.nt) {
      cplx rv = y0[r]; fma_acc(rv, lam, y1[r]); fma_acc(rv, lam2, y2[r]);
      rr += abs2(rv);
      // direction: residual inverse iteration from the third step on, plain inverse iteration K^-1 P'(lambda) x before
      cplx w = y1[r]; fma_acc(w, 2.0 * lam, y2[r]);
      y0[r] = (it >= 2) ? rv : w;
    }
    rr = cta_sum(c, rr);
    const double la = cabs(lam);
    resid = sqrt(rr) / (sqrt(n0) + la * sqrt(n1) + la * la * sqrt(n2));
    if (resid < tol) { ++it; break; }                         // uniform: every thread holds the same reductions
    cta_sync();
    cta_lu_replay(c, K, n, ipiv, y0);
    double nz = 0.0;
    for (int r = c.tid; r < n; r += c.nt) {
      const cplx z = (it >= 2) ? x[r] - y0[r] : y0[r];
      y0[r] = z; nz += abs2(z);
    }
    nz = 1.0 / sqrt(cta_sum(c, nz));
    for (int r = c.tid; r < n; r += c.nt) x[r] = y0[r] * nz;
    cta_sync();
  }
  // scale like temporal.f90:867-879: the first entry of maximum modulus becomes 1
  double best = -1.0; int bi = 0;
  for (int r = c.tid; r < n; r += c.nt) { const double m = cabs(x[r]); if (m > best) { best = m; bi = r; } }
  cta_argmax(c, best, bi);
  const cplx sc = x[bi];
  cta_sync();
  for (int r = c.tid; r < n; r += c.nt) xg[r] = cdiv(x[r], sc);
  if (c.tid == 0) { out[0] = lam.re; out[1] = lam.im; out[2] = resid; out[3] = (double)it; }
}

}  // namespace stab
#endif
