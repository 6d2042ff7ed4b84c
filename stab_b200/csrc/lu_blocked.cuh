// lu_blocked.cuh -- batched blocked LU reduce of the spatial problem (GPU only), stage (2) of the pipeline:
//     X = C0^-1 [ -C1 | -C2 ]        (ZGETRF + 2 x ZGETRS, spatial.f90:978-1004)
// C0 (n x n) is factored by a right-looking blocked LU with partial pivoting (panel width 32) and the 2n right-hand
// side columns -- the top half of the companion buffer, spatial.f90:1007-1008 -- ride along as extra trailing columns,
// so the forward substitution is part of the rank-32 updates; the back substitution U X = Y is blocked the same way.
// Per panel:
//   k_lu_panel      one CTA per matrix: ZGETF2 on the n-j0 x 32 panel (IZAMAX pivot = first maximum of |re|+|im|,
//                   reciprocal scaling, rank-1 updates inside the panel); the pivot rows are interchanged inside the panel only
//   k_lu_swap_trsm_warp  one warp per trailing / right-hand-side column: the panel's row interchanges (ZLASWP), composed
//                   into one gather by k_lu_panel, and the unit-lower 32 x 32 triangular solve that turns the pivot rows
//                   into the U12 block row (ZTRSM 'L','L','N','U')
//   k_lu_gemm<0>    DMMA rank-32 update  [C22 | B2] -= L21 [U12 | B1]   (gemm_pipe.cuh, persistent cp.async ring)
// Back substitution, last block row first:
//   k_lu_back_trsm  X_R = U_RR^-1 Y_R (one thread per right-hand-side column, true complex divisions as ZTRSM)
//   k_lu_gemm<1>    DMMA rank-32 update  Y(0:i0, :) -= U(0:i0, R) X_R
// The columns to the LEFT of a panel are not interchanged: L is never used again (the forward substitution has already
// happened), only U and the transformed right-hand side are.  Row order does not enter the arithmetic of any entry,
// so X is the same as with LAPACK's explicit interchanges.
// Algorithmic work: (8/3) n^3 [factor] + 8 n^3 [forward] + 8 n^3 [backward] real flops per matrix.
#pragma once
#include "common.cuh"
#include "gemm_pipe.cuh"

#ifndef STAB_EMU
namespace stab {

constexpr int LU_NB = 32;
constexpr int LU_TRSM_THREADS = 128;
constexpr int LU_PERM = 3 * LU_NB + 1;  // src[32] | nd | dpos[32] | dsrc[32]
constexpr int LU_SWAP_WARPS = 8;

struct LuBatch {
  cplx* C; size_t cstride; int n;       // C0, n x n, ld n
  cplx* B; size_t bstride; int ldb;     // right-hand side: rows 0..n-1 of an ldb x nrhs block
  int nrhs;
  int* ipiv;                            // n per matrix (absolute row positions, ZGETRF convention, 0-based)
  int* perm;                            // LU_PERM ints per matrix: the current panel's interchanges as one gather (k_lu_panel)
  int* info;                            // per matrix: 0 or (index of the first exactly-zero pivot) + 1
};

__global__ void __launch_bounds__(512) k_lu_panel(LuBatch lb, int j0) {
  __shared__ double red[64];
  __shared__ cplx prow[LU_NB];
  Cta c = make_cta(red);
  const int p = blockIdx.x, n = lb.n;
  cplx* C = lb.C + (size_t)p * lb.cstride;
  int* ipiv = lb.ipiv + (size_t)p * n;
  const int jb = min(LU_NB, n - j0);
  for (int s = 0; s < jb; ++s) {
    const int j = j0 + s;
    cplx* col = C + (size_t)j * n;
    double best = -1.0; int bi = 0x7fffffff;
    for (int r = j + c.tid; r < n; r += c.nt) {
      const double m = cabs1(col[r]);
      if (m > best) { best = m; bi = r; }
    }
    cta_argmax(c, best, bi);
    if (!(best > 0.0) || bi >= n) { best = 0.0; bi = j; }   // a column of zeros -- or of NaNs, which no comparison selects
    if (c.tid == 0) ipiv[j] = bi;
    if (best == 0.0) {                                   // exactly singular column: ZGETF2 records it and goes on
      if (c.tid == 0 && lb.info[p] == 0) lb.info[p] = j + 1;
      cta_sync();
      continue;
    }
    if (c.tid < jb) {                                    // interchange inside the panel, pivot row to shared memory
      cplx* e = C + (size_t)(j0 + c.tid) * n;
      const cplx a = e[bi];
      if (bi != j) { e[bi] = e[j]; e[j] = a; }
      prow[c.tid] = a;
    }
    cta_sync();
    const cplx rp = cdiv(mk(1.0, 0.0), prow[s]);
    for (int r = j + 1 + c.tid; r < n; r += c.nt) {
      const cplx l = col[r] * rp;
      col[r] = l;
#pragma unroll 8
      for (int q = s + 1; q < jb; ++q) {
        cplx* e = C + (size_t)(j0 + q) * n + r;
        cplx v = *e;
        fms_acc(v, l, prow[q]);
        *e = v;
      }
    }
    cta_sync();
  }
  // The panel's interchanges (ZLASWP, applied in order) composed into ONE gather for the columns outside the panel:
  // position j0+s receives the entry of row src[s]; the rows outside the top block that were displaced receive dsrc -> dpos.
  if (c.tid == 0) {
    int top[LU_NB], dpos[LU_NB], dval[LU_NB], nd = 0;
    const int r0 = j0 + jb;
    for (int s = 0; s < LU_NB; ++s) top[s] = j0 + s;
    for (int s = 0; s < jb; ++s) {
      const int pr = ipiv[j0 + s];
      if (pr == j0 + s) continue;
      if (pr < r0) { const int t = top[s]; top[s] = top[pr - j0]; top[pr - j0] = t; }
      else {
        int q = 0;
        while (q < nd && dpos[q] != pr) ++q;
        if (q == nd) { dpos[nd] = pr; dval[nd] = pr; ++nd; }
        const int t = top[s]; top[s] = dval[q]; dval[q] = t;
      }
    }
    int* pm = lb.perm + (size_t)p * LU_PERM;
    for (int s = 0; s < LU_NB; ++s) { pm[s] = top[s]; pm[LU_NB + 1 + s] = s < nd ? dpos[s] : 0; pm[2 * LU_NB + 1 + s] = s < nd ? dval[s] : 0; }
    pm[LU_NB] = nd;
  }
}

// One WARP per trailing / right-hand-side column (lane = row of the block row): gather the interchanged rows (all loads
// of a column in flight at once), write the displaced rows back, then the unit-lower solve U12 = L11^-1 (P A12) as a
// column-oriented forward substitution over lanes (row k broadcast by shuffle, lane l > k eliminates with L11(l,k), which
// it keeps in registers).  grid (column groups, matrices), LU_SWAP_WARPS warps per CTA.
__global__ void __launch_bounds__(LU_SWAP_WARPS * 32) k_lu_swap_trsm_warp(LuBatch lb, int j0, int cols_per_cta) {
  __shared__ cplx sL[LU_NB][LU_NB + 1];
  __shared__ int spm[LU_PERM];
  const int p = blockIdx.y, n = lb.n, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int jb = min(LU_NB, n - j0), r0 = j0 + jb, ntrail = n - r0, ncols = ntrail + lb.nrhs;
  cplx* C = lb.C + (size_t)p * lb.cstride;
  for (int e = tid; e < LU_NB * LU_NB; e += LU_SWAP_WARPS * 32) {
    const int i = e % LU_NB, k = e / LU_NB;
    sL[i][k] = (i < jb && k < i) ? C[(j0 + i) + (size_t)(j0 + k) * n] : mk(0.0, 0.0);
  }
  for (int e = tid; e < LU_PERM; e += LU_SWAP_WARPS * 32) spm[e] = lb.perm[(size_t)p * LU_PERM + e];
  __syncthreads();
  cplx Lrow[LU_NB];
#pragma unroll
  for (int k = 0; k < LU_NB; ++k) Lrow[k] = sL[lane][k];
  const int src = spm[lane], nd = spm[LU_NB], dpos = spm[LU_NB + 1 + lane], dsrc = spm[2 * LU_NB + 1 + lane];
  const int q0 = blockIdx.x * cols_per_cta, q1 = min(ncols, q0 + cols_per_cta);
  for (int q = q0 + wid; q < q1; q += LU_SWAP_WARPS) {
    cplx* col = (q < ntrail) ? C + (size_t)(r0 + q) * n : lb.B + (size_t)p * lb.bstride + (size_t)(q - ntrail) * lb.ldb;
    cplx v = (lane < jb) ? col[src] : mk(0.0, 0.0);
    cplx d = (lane < nd) ? col[dsrc] : mk(0.0, 0.0);
    __syncwarp();
    if (lane < nd) col[dpos] = d;
#pragma unroll
    for (int k = 0; k < LU_NB - 1; ++k) {
      const cplx uk = mk(__shfl_sync(0xffffffffu, v.re, k), __shfl_sync(0xffffffffu, v.im, k));
      if (lane > k) fms_acc(v, Lrow[k], uk);
    }
    if (lane < jb) col[j0 + lane] = v;
  }
}

// X_R = U_RR^-1 Y_R for the block row R = [i0, i0 + bs), one thread per right-hand-side column
__global__ void __launch_bounds__(LU_TRSM_THREADS) k_lu_back_trsm(LuBatch lb, int i0, int bs) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cplx* sU = reinterpret_cast<cplx*>(smem_raw);                       // [LU_NB][LU_TRSM_THREADS]
  cplx* sT = sU + LU_NB * LU_TRSM_THREADS;                            // [LU_NB][LU_NB] upper block, column-major
  const int p = blockIdx.y, n = lb.n, tid = threadIdx.x;
  const cplx* C = lb.C + (size_t)p * lb.cstride;
  for (int e = tid; e < bs * bs; e += LU_TRSM_THREADS) {
    const int i = e % bs, k = e / bs;
    sT[i + k * LU_NB] = C[(i0 + i) + (size_t)(i0 + k) * n];
  }
  __syncthreads();
  const int q = blockIdx.x * LU_TRSM_THREADS + tid;
  if (q >= lb.nrhs) return;
  cplx* col = lb.B + (size_t)p * lb.bstride + (size_t)q * lb.ldb;
#define LU_U(s) sU[(s) * LU_TRSM_THREADS + tid]
  for (int s = 0; s < bs; ++s) LU_U(s) = col[i0 + s];
  for (int i = bs - 1; i >= 0; --i) {
    const cplx x = cdiv(LU_U(i), sT[i + i * LU_NB]);
    LU_U(i) = x;
    for (int k = 0; k < i; ++k) fms_acc(LU_U(k), sT[k + i * LU_NB], x);
  }
  for (int s = 0; s < bs; ++s) col[i0 + s] = LU_U(s);
#undef LU_U
}

// MODE 0: trailing update of panel j0 (matl = 2 mat + part; part 0: C0's trailing columns, part 1: right-hand side)
// MODE 1: back-substitution update above the block row [j0, j0 + jb)
template <int MODE>
struct LuProb {
  LuBatch lb; int j0;
  SD_DEV GemmProb operator()(int matl) const {
    GemmProb q;
    const int n = lb.n, jb = min(LU_NB, n - j0);
    q.lsi = 1; q.lsl = n; q.rsl = 1; q.K = jb;
    if (MODE == 0) {
      const int mat = matl >> 1, part = matl & 1, r0 = j0 + jb;
      cplx* C = lb.C + (size_t)mat * lb.cstride;
      cplx* B = lb.B + (size_t)mat * lb.bstride;
      q.m = n - r0;
      q.L = C + r0 + (size_t)j0 * n;
      if (part == 0) { q.R = C + j0 + (size_t)r0 * n; q.rsj = n; q.nc = n - r0; q.C = C + r0 + (size_t)r0 * n; q.ldc = n; }
      else           { q.R = B + j0; q.rsj = lb.ldb; q.nc = lb.nrhs; q.C = B + r0; q.ldc = lb.ldb; }
    } else {
      cplx* C = lb.C + (size_t)matl * lb.cstride;
      cplx* B = lb.B + (size_t)matl * lb.bstride;
      q.m = j0;
      q.L = C + (size_t)j0 * n;
      q.R = B + j0; q.rsj = lb.ldb; q.nc = lb.nrhs; q.C = B; q.ldc = lb.ldb;
    }
    return q;
  }
};

template <int MODE>
__global__ void __launch_bounds__(GEMM_THREADS, 2) k_lu_gemm(LuBatch lb, int j0, int tiles_i, int tiles_j, int nmat) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* smem = reinterpret_cast<double*>(smem_raw);
  LuProb<MODE> pf{lb, j0};
  //            TM  TN  SUB   CONJL  CONJR  LKFAST RKFAST
  gemm_pipe_run<64, 32, true, false, false, false, true>(pf, tiles_i, tiles_j, nmat, smem);
}

}  // namespace stab
#endif
