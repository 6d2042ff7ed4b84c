// kernels.cuh -- __global__ entry points (sm_100a).  Each eigen-stage kernel is "one CTA per sweep
// point"; the assembly kernels are grid-wide streaming writers.
#pragma once
#include "common.cuh"
#include "tables.cuh"
#include "assemble.cuh"
#include "balance.cuh"
#include "hessenberg.cuh"
#include "gemm_pipe.cuh"
#include "evec.cuh"
#include "lu.cuh"
#include "hess_blocked.cuh"
#include "launch.h"

namespace stab {

struct SweepDev {            // per-point sweep values (device arrays; Re/Ma may be null)
  const cplx* s1;            // alpha (temporal) or omega (spatial)
  const cplx* s2;            // beta
  const double* Re;
  const double* Ma;
};

SD_DEV Phys point_phys(const Phys& base, const SweepDev& sw, int p) {
  Phys q = base;
  if (sw.Re) q.Re = sw.Re[p];
  if (sw.Ma) q.Ma = sw.Ma[p];
  q.navier = !(q.Re >= 1.0e98 || q.Re == 0.0);        // temporal.f90:88-91
  return q;
}

// ---- stage 1a: node coefficients ---------------------------------------------------------------
// coef layout temporal: [p][node][3][25]; spatial: [p][node][6][25] (C0.c1, C0.c2, C0.c0, C1.c1, C1.c0, C2.c0)
__global__ void k_node_coef_temporal(GridDev g, Phys base, SweepDev sw, int p0, int apply_b0inv, cplx* coef, cplx* b0blk) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int p = blockIdx.y;
  if (i >= g.ny) return;
  Phys ph = point_phys(base, sw, p0 + p);
  NodeIn q = load_node(g, i);
  Tables t;
  node_tables_temporal(q, ph, t);
  PointTemporal pt; pt.alpha = sw.s1[p0 + p]; pt.beta = sw.s2[p0 + p];
  NodeCoef3 o;
  node_coef_temporal(t, i, g.ny, g.wallt, g.deta[i], g.d2eta[i], pt, apply_b0inv != 0, o);
  cplx* dst = coef + ((size_t)p * g.ny + i) * 75;
  for (int k = 0; k < 25; ++k) { dst[k] = o.c1[k]; dst[25 + k] = o.c2[k]; dst[50 + k] = o.c0[k]; }
  if (b0blk) node_b0_temporal(t, i, g.ny, g.wallt, b0blk + ((size_t)p * g.ny + i) * 25);
}

__global__ void k_node_coef_spatial(GridDev g, Phys base, SweepDev sw, int p0, cplx* coef) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int p = blockIdx.y;
  if (i >= g.ny) return;
  Phys ph = point_phys(base, sw, p0 + p);
  NodeIn q = load_node(g, i);
  Tables t;
  node_tables_spatial(q, ph, t);
  PointSpatial pt; pt.omega = sw.s1[p0 + p]; pt.beta = sw.s2[p0 + p];
  NodeCoefSpatial o;
  node_coef_spatial(t, i, g.ny, g.wallt, g.top, g.deta[i], g.d2eta[i], pt, o);
  cplx* dst = coef + ((size_t)p * g.ny + i) * 150;
  for (int k = 0; k < 25; ++k) {
    dst[k] = o.C0.c1[k]; dst[25 + k] = o.C0.c2[k]; dst[50 + k] = o.C0.c0[k];
    dst[75 + k] = o.C1c1[k]; dst[100 + k] = o.C1c0[k]; dst[125 + k] = o.C2c0[k];
  }
}

// ---- stage 1b: dense operator streaming writers -------------------------------------------------
// One thread per matrix row r = 5 i + e; a CTA covers 128 rows and JT node-columns (5 JT matrix
// columns).  For a fixed column consecutive threads write consecutive rows: full 16-byte
// coalesced stores; D1/D2 are read through the read-only path (L1/L2 resident, 128 KB each).
// Algorithmic traffic: 16 n^2 bytes written per matrix; HBM-bound.
constexpr int ASM_ROWS = 128;
constexpr int ASM_JT = 8;

__global__ void __launch_bounds__(ASM_ROWS) k_assemble_temporal(GridDev g, const cplx* __restrict__ coef, cplx* __restrict__ M, size_t mstride) {
  const int n = 5 * g.ny, ny = g.ny;
  const int r = blockIdx.x * ASM_ROWS + threadIdx.x;
  const int p = blockIdx.z;
  if (r >= n) return;
  const int i = r / 5, e = r - 5 * i;
  const cplx* cf = coef + ((size_t)p * ny + i) * 75 + e * 5;
  cplx c1[5], c2[5], c0[5];
#pragma unroll
  for (int v = 0; v < 5; ++v) { c1[v] = cf[v]; c2[v] = cf[25 + v]; c0[v] = cf[50 + v]; }
  cplx* out = M + (size_t)p * mstride + r;
  const int j0 = blockIdx.y * ASM_JT;
  const bool wallrow = (g.wallt == 2 && i == ny - 1);
#pragma unroll 2
  for (int j = j0; j < j0 + ASM_JT && j < ny; ++j) {
    const double d1 = __ldg(g.D1 + i + (size_t)j * ny);
    const double d2 = __ldg(g.D2 + i + (size_t)j * ny);
    const double d2w = wallrow ? __ldg(g.Dt2w + j) : d2;
#pragma unroll
    for (int v = 0; v < 5; ++v) {
      cplx a = c1[v] * d1 + c2[v] * ((v == 4) ? d2w : d2);
      if (i == j) a += c0[v];
      out[(size_t)(5 * j + v) * n] = a;
    }
  }
}

// Spatial: C (n x n, ldc = n) <- C0 ; companion (2n x 2n, ld 2n): top <- [-C1 | -C2], bottom <- [I | 0]
__global__ void __launch_bounds__(ASM_ROWS) k_assemble_spatial(GridDev g, const cplx* __restrict__ coef, cplx* __restrict__ C, size_t cstride,
                                                               cplx* __restrict__ B, size_t bstride) {
  const int n = 5 * g.ny, ny = g.ny, n2 = 2 * n;
  const int r = blockIdx.x * ASM_ROWS + threadIdx.x;
  const int p = blockIdx.z;
  if (r >= n) return;
  const int i = r / 5, e = r - 5 * i;
  const cplx* cf = coef + ((size_t)p * ny + i) * 150 + e * 5;
  cplx c1[5], c2[5], c0[5], k1[5], k0[5], q0[5];
#pragma unroll
  for (int v = 0; v < 5; ++v) { c1[v] = cf[v]; c2[v] = cf[25 + v]; c0[v] = cf[50 + v]; k1[v] = cf[75 + v]; k0[v] = cf[100 + v]; q0[v] = cf[125 + v]; }
  cplx* outC = C + (size_t)p * cstride + r;
  cplx* outB = B + (size_t)p * bstride + r;
  const int j0 = blockIdx.y * ASM_JT;
  const bool wallrow = (g.wallt == 2 && i == ny - 1);
  const cplx zero = mk(0.0, 0.0);
  for (int j = j0; j < j0 + ASM_JT && j < ny; ++j) {
    const double d1 = __ldg(g.D1 + i + (size_t)j * ny);
    const double d2 = __ldg(g.D2 + i + (size_t)j * ny);
    const double d2w = wallrow ? __ldg(g.Dt2w + j) : d2;
#pragma unroll
    for (int v = 0; v < 5; ++v) {
      const int col = 5 * j + v;
      cplx a = c1[v] * d1 + c2[v] * ((v == 4) ? d2w : d2);
      cplx b1 = k1[v] * d1;
      cplx b2 = zero;
      if (i == j) { a += c0[v]; b1 += k0[v]; b2 = q0[v]; }
      outC[(size_t)col * n] = a;
      outB[(size_t)col * n2] = b1;                          // -C1
      outB[(size_t)(n + col) * n2] = b2;                    // -C2
      outB[(size_t)col * n2 + n] = (col == r) ? mk(1.0, 0.0) : zero;   // identity block
      outB[(size_t)(n + col) * n2 + n] = zero;
    }
  }
}

// element-wise inspection kernel: A0 (no B0^-1) and B0, or C0/C1/C2 with the reference's signs
__global__ void k_inspect_temporal(GridDev g, const cplx* coef, const cplx* b0blk, cplx* A0, cplx* B0) {
  const int n = 5 * g.ny;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)n * n) return;
  const int r = (int)(idx % n), c = (int)(idx / n);
  const int i = r / 5, e = r % 5, j = c / 5, v = c % 5;
  const cplx* cf = coef + (size_t)i * 75;
  A0[idx] = op_element(cf, cf + 25, cf + 50, g, i, e, j, v);
  B0[idx] = (i == j) ? b0blk[(size_t)i * 25 + e * 5 + v] : mk(0.0, 0.0);
}

__global__ void k_inspect_spatial(GridDev g, const cplx* coef, cplx* C0, cplx* C1, cplx* C2) {
  const int n = 5 * g.ny;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)n * n) return;
  const int r = (int)(idx % n), c = (int)(idx / n);
  const int i = r / 5, e = r % 5, j = c / 5, v = c % 5;
  const cplx* cf = coef + (size_t)i * 150;
  C0[idx] = op_element(cf, cf + 25, cf + 50, g, i, e, j, v);
  const int k = e * 5 + v;
  cplx b1 = cf[75 + k] * g.D1[i + (size_t)j * g.ny];
  cplx b2 = mk(0.0, 0.0);
  if (i == j) { b1 += cf[100 + k]; b2 = cf[125 + k]; }
  C1[idx] = -b1;
  C2[idx] = -b2;
}

// ---- stage 2 (spatial): LU reduce ---------------------------------------------------------------
__global__ void k_lu(cplx* C, size_t cstride, int n, cplx* B, size_t bstride, int nrhs, int ldb, int* info) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* red = reinterpret_cast<double*>(smem_raw);
  cplx* sl = reinterpret_cast<cplx*>(smem_raw + 160 * sizeof(double));
  Cta c = make_cta(red);
  const int p = blockIdx.x;
  int r = cta_lu_solve(c, C + (size_t)p * cstride, n, n, B + (size_t)p * bstride, nrhs, ldb, sl);
  if (threadIdx.x == 0) info[p] = r;
}

// ---- stage 3a: balance --------------------------------------------------------------------------
SD_HD int balance_block(int n) { const int b = (80 * 1024) / (16 * n); return b >= 8 ? 8 : (b >= 4 ? 4 : (b >= 2 ? 2 : 1)); }   // power of two
__global__ void __launch_bounds__(256, 2) k_balance(cplx* A, size_t astride, int n, double* scale, int* cnt, int* ilohi, int bal_b) {
  __shared__ double red[160];
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Cta c = make_cta(red);
  const int p = blockIdx.x;
  int ilo, ihi;
  cta_balance(c, A + (size_t)p * astride, n, n, scale + (size_t)p * n, cnt + (size_t)p * n, reinterpret_cast<double*>(smem_raw), bal_b, ilo, ihi);
  if (threadIdx.x == 0) { ilohi[2 * p] = ilo; ilohi[2 * p + 1] = ihi; }
}
// orders above 640 (spatial companion, Ny = 256): one CTA of 512 threads per SM, 8-index blocks while they fit
SD_HD int balance_block_wide(int n) { return (size_t)144 * n <= 200 * 1024 ? 8 : balance_block(n); }
__global__ void __launch_bounds__(512, 1) k_balance_wide(cplx* A, size_t astride, int n, double* scale, int* cnt, int* ilohi, int bal_b) {
  __shared__ double red[160];
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Cta c = make_cta(red);
  const int p = blockIdx.x;
  int ilo, ihi;
  cta_balance(c, A + (size_t)p * astride, n, n, scale + (size_t)p * n, cnt + (size_t)p * n, reinterpret_cast<double*>(smem_raw), bal_b, ilo, ihi);
  if (threadIdx.x == 0) { ilohi[2 * p] = ilo; ilohi[2 * p + 1] = ihi; }
}

// ---- stage 3b: Hessenberg -----------------------------------------------------------------------
__global__ void k_hessenberg(cplx* A, size_t astride, int n, const int* ilohi, cplx* tau) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* red = reinterpret_cast<double*>(smem_raw);
  cplx* sv = reinterpret_cast<cplx*>(smem_raw + 160 * sizeof(double));
  cplx* sy = sv + n;
  Cta c = make_cta(red);
  const int p = blockIdx.x;
  cta_hessenberg(c, A + (size_t)p * astride, n, n, ilohi[2 * p], ilohi[2 * p + 1], tau + (size_t)p * n, sv, sy);
}

// ---- stage 3b': batched blocked Hessenberg (hess_blocked.cuh) -------------------------------------
__global__ void __launch_bounds__(512) k_hb_panel_step(HessBatch hb, int panel, int j) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* red = reinterpret_cast<double*>(smem_raw);
  cplx* sb = reinterpret_cast<cplx*>(smem_raw + 160 * sizeof(double));
  cplx* sw = sb + hb.n;
  cplx* st = sw + HB_NB;
  cplx* sc = st + HB_NB;
  Cta c = make_cta(red);
  cta_hb_panel_step(c, hb, hb.mat0 + blockIdx.x, panel, j, red, sb, sw, st, sc);
}

__global__ void __launch_bounds__(HB_GEMV_ROWS) k_hb_gemv(HessBatch hb, int panel, int j) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Cta c = make_cta(nullptr);
  if (blockIdx.y == HB_CHUNKS) {                           // the extra chunk index: one CTA per matrix takes the V^H v dot products
    if (blockIdx.x == 0) cta_hb_vdots(c, hb, hb.mat0 + blockIdx.z, panel, j);
    return;
  }
  cta_hb_gemv(c, hb, hb.mat0 + blockIdx.z, panel, j, blockIdx.x, blockIdx.y, reinterpret_cast<cplx*>(smem_raw));
}

template <int PHASE, bool USE_MMA>
__global__ void __launch_bounds__(GEMM_THREADS, 2) k_hb_gemm(HessBatch hb, int panel) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Cta c = make_cta(nullptr);
  cta_hb_gemm<PHASE, USE_MMA>(c, hb, hb.mat0 + blockIdx.z, panel, blockIdx.x, blockIdx.y, reinterpret_cast<double*>(smem_raw));
}

#ifndef STAB_EMU
// ---- pipelined DMMA GEMM path (gemm_pipe.cuh): plain operands, V materialised per panel ----------
// Vx(r, l) = V(r, l) of the panel (zeros above / beyond the reflector, explicit one); grid (row blocks, batch).
// want bit 0: also VT = Vx T (the operand of Y_top = A_top (V T) and of the back-transformation X -= (V T)(V^H X));
// want bit 1: also VTh = Vx T^H (the operand of the left update A -= (V T^H)(V^H A)).  T is upper triangular, so a row
// costs 528 complex FMAs per product, from registers, with coalesced loads and stores -- this replaces the passes that
// applied T to Y_top and to W (one thread per column with stride-32 accesses: 9.8 ms per 296 points at n = 640).
__global__ void __launch_bounds__(128) k_hb_vx(HessBatch hb, int panel, int want) {
  __shared__ cplx sT[HB_NB * HB_NB];
  const int mat = hb.mat0 + blockIdx.y, n = hb.n;
  const int ilo = hb.ilohi[2 * mat], ihi = hb.ilohi[2 * mat + 1];
  const int k = ilo + panel * HB_NB;
  if (k >= ihi) return;
  if (want) {
    const cplx* T = hb.T + ((size_t)mat * hb.P + panel) * HB_NB * HB_NB;
    for (int e = threadIdx.x; e < HB_NB * HB_NB; e += blockDim.x) sT[e] = T[e];
    __syncthreads();
  }
  const cplx* A = hb.A + (size_t)mat * hb.astride;
  cplx* Vx = hb.Vx + (size_t)mat * n * HB_NB;
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  cplx v[HB_NB];
#pragma unroll
  for (int l = 0; l < HB_NB; ++l) { v[l] = hb_v(A, n, k, ihi, r, l); Vx[r + (size_t)l * n] = v[l]; }
  const bool zero = (r <= k) || (r > ihi);
  if (want & 1) {
    cplx* VT = hb.VT + (size_t)mat * n * HB_NB;
#pragma unroll
    for (int m = 0; m < HB_NB; ++m) {
      cplx s = mk(0.0, 0.0);
      if (!zero) {
#pragma unroll
        for (int l = 0; l <= m; ++l) fma_acc(s, v[l], sT[l + m * HB_NB]);
      }
      VT[r + (size_t)m * n] = s;
    }
  }
  if (want & 2) {
    cplx* VTh = hb.VTh + (size_t)mat * n * HB_NB;
#pragma unroll
    for (int m = 0; m < HB_NB; ++m) {
      cplx s = mk(0.0, 0.0);
      if (!zero) {
#pragma unroll
        for (int l = m; l < HB_NB; ++l) fma_acc_conj(s, sT[m + l * HB_NB], v[l]);       // conj(T(m,l)) v(l)
      }
      VTh[r + (size_t)m * n] = s;
    }
  }
}

enum PipePhase { PP_YTOP = 0, PP_RIGHT_TRAIL, PP_RIGHT_PANEL, PP_LEFT_W, PP_LEFT_UPD, PP_BT_W, PP_BT_UPD, PP_RIGHT_TOP, PP_S, PP_FUSED_UPD };

template <int PHASE>
struct HbProb {
  HessBatch hb; cplx* X; size_t xstride; int panel;
  SD_DEV GemmProb operator()(int matl) const {
    GemmProb q;
    const int mat = hb.mat0 + matl, n = hb.n, lda = n;
    const int ilo = hb.ilohi[2 * mat], ihi = hb.ilohi[2 * mat + 1];
    const int k = ilo + panel * HB_NB;
    q.m = 0; q.nc = 0; q.K = 0; q.L = nullptr; q.R = nullptr; q.C = nullptr; q.ldc = 1; q.lsi = q.lsl = q.rsl = q.rsj = 0;
    if (k >= ihi) return q;
    cplx* A = hb.A + (size_t)mat * hb.astride;
    cplx* Y = hb.Y + (size_t)mat * n * HB_NB;
    cplx* W = hb.W + (size_t)mat * n * HB_NB;
    const cplx* Vx = hb.Vx + (size_t)mat * n * HB_NB;
    const cplx* VT = hb.VT ? hb.VT + (size_t)mat * n * HB_NB : Vx;       // T folded into the operand (then no pass applies T afterwards)
    const cplx* VTh = hb.VTh ? hb.VTh + (size_t)mat * n * HB_NB : Vx;
    if (PHASE == PP_YTOP) {                 // Y(0:k+1, :) = A(0:k+1, k+1:ihi+1) V(k+1:ihi+1, :)
      q.m = k + 1; q.nc = HB_NB; q.K = ihi - k;
      q.L = A + (size_t)(k + 1) * lda; q.lsi = 1; q.lsl = lda;
      q.R = VT + (k + 1); q.rsl = 1; q.rsj = n;
      q.C = Y; q.ldc = n;
    } else if (PHASE == PP_RIGHT_TRAIL) {   // A(0:ihi+1, k+NB:ihi+1) -= Y V(k+NB:ihi+1, :)^H
      q.m = ihi + 1; q.nc = ihi + 1 - (k + HB_NB); q.K = HB_NB;
      q.L = Y; q.lsi = 1; q.lsl = n;
      q.R = Vx + (k + HB_NB); q.rsl = n; q.rsj = 1;
      q.C = A + (size_t)(k + HB_NB) * lda; q.ldc = lda;
    } else if (PHASE == PP_RIGHT_PANEL) {   // A(0:k+1, k+1:k+NB) -= Y(0:k+1, :) V(k+1:k+NB, :)^H
      q.m = k + 1; q.nc = HB_NB - 1; if (q.nc > n - (k + 1)) q.nc = n - (k + 1); q.K = HB_NB;
      q.L = Y; q.lsi = 1; q.lsl = n;
      q.R = Vx + (k + 1); q.rsl = n; q.rsj = 1;
      q.C = A + (size_t)(k + 1) * lda; q.ldc = lda;
    } else if (PHASE == PP_LEFT_W) {        // W (NB x nc) = V^H A(k+1:ihi+1, k+NB:n)
      q.m = HB_NB; q.nc = n - (k + HB_NB); q.K = ihi - k;
      q.L = Vx + (k + 1); q.lsi = n; q.lsl = 1;
      q.R = A + (k + 1) + (size_t)(k + HB_NB) * lda; q.rsl = 1; q.rsj = lda;
      q.C = W; q.ldc = HB_NB;
    } else if (PHASE == PP_LEFT_UPD) {      // A(k+1:ihi+1, k+NB:n) -= V W
      q.m = ihi - k; q.nc = n - (k + HB_NB); q.K = HB_NB;
      q.L = VTh + (k + 1); q.lsi = 1; q.lsl = n;
      q.R = W; q.rsl = 1; q.rsj = HB_NB;
      q.C = A + (k + 1) + (size_t)(k + HB_NB) * lda; q.ldc = lda;
    } else if (PHASE == PP_RIGHT_TOP) {     // A(0:k+1, k+NB:ihi+1) -= Y(0:k+1, :) V(k+NB:ihi+1, :)^H   (rows above the panel only)
      q.m = k + 1; q.nc = ihi + 1 - (k + HB_NB); q.K = HB_NB;
      q.L = Y; q.lsi = 1; q.lsl = n;
      q.R = Vx + (k + HB_NB); q.rsl = n; q.rsj = 1;
      q.C = A + (size_t)(k + HB_NB) * lda; q.ldc = lda;
    } else if (PHASE == PP_S) {             // S (NB x NB) = V^H Y(k+1:ihi+1, :)
      q.m = HB_NB; q.nc = HB_NB; q.K = ihi - k;
      q.L = Vx + (k + 1); q.lsi = n; q.lsl = 1;
      q.R = Y + (k + 1); q.rsl = 1; q.rsj = n;
      q.C = hb.S + (size_t)mat * HB_NB * HB_NB; q.ldc = HB_NB;
    } else if (PHASE == PP_FUSED_UPD) {     // A(k+1:ihi+1, k+NB:n) -= Y Vh + V W   (right and left update in one pass)
      q.m = ihi - k; q.nc = n - (k + HB_NB); q.K = HB_NB;
      q.L = Y + (k + 1); q.lsi = 1; q.lsl = n;
      q.R = hb.Vh + (size_t)mat * n * HB_NB; q.rsl = 1; q.rsj = HB_NB;
      q.L2 = Vx + (k + 1); q.R2 = W; q.K2 = HB_NB;
      q.C = A + (k + 1) + (size_t)(k + HB_NB) * lda; q.ldc = lda;
    } else if (PHASE == PP_BT_W) {          // W (NB x n) = V^H X(k+1:ihi+1, :)
      cplx* Xm = X + (size_t)mat * xstride;
      q.m = HB_NB; q.nc = n; q.K = ihi - k;
      q.L = Vx + (k + 1); q.lsi = n; q.lsl = 1;
      q.R = Xm + (k + 1); q.rsl = 1; q.rsj = n;
      q.C = W; q.ldc = HB_NB;
    } else {                                // X(k+1:ihi+1, :) -= V W
      cplx* Xm = X + (size_t)mat * xstride;
      q.m = ihi - k; q.nc = n; q.K = HB_NB;
      q.L = VT + (k + 1); q.lsi = 1; q.lsl = n;
      q.R = W; q.rsl = 1; q.rsj = HB_NB;
      q.C = Xm + (k + 1); q.ldc = n;
    }
    return q;
  }
};

template <int PHASE>
__global__ void __launch_bounds__(GEMM_THREADS, 2) k_gemm_pipe(HessBatch hb, cplx* X, size_t xstride, int panel, int tiles_i, int tiles_j, int nmat) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* smem = reinterpret_cast<double*>(smem_raw);
  HbProb<PHASE> pf{hb, X, xstride, panel};
  //                                    TM  TN  SUB    CONJL  CONJR  LKFAST RKFAST
  if (PHASE == PP_YTOP)                                   gemm_pipe_run<64, 32, false, false, false, false, true>(pf, tiles_i, tiles_j, nmat, smem);
  else if (PHASE == PP_RIGHT_TRAIL || PHASE == PP_RIGHT_PANEL || PHASE == PP_RIGHT_TOP) gemm_pipe_run<64, 32, true, false, true, false, false>(pf, tiles_i, tiles_j, nmat, smem);
  else if (PHASE == PP_LEFT_W || PHASE == PP_BT_W || PHASE == PP_S) gemm_pipe_run<32, 64, false, true, false, true, true>(pf, tiles_i, tiles_j, nmat, smem);
  else                                                    gemm_pipe_run<64, 32, true, false, false, false, true>(pf, tiles_i, tiles_j, nmat, smem);
}
#endif

__global__ void k_hb_ytop_T(HessBatch hb, int panel) {
  Cta c = make_cta(nullptr);
  cta_hb_ytop_T(c, hb, hb.mat0 + blockIdx.y, panel, blockIdx.x);
}

__global__ void k_hb_w_T(HessBatch hb, int panel) {
  Cta c = make_cta(nullptr);
  const int mat = hb.mat0 + blockIdx.y, n = hb.n;
  const int ilo = hb.ilohi[2 * mat], ihi = hb.ilohi[2 * mat + 1];
  const int k = ilo + panel * HB_NB;
  if (k >= ihi) return;
  cta_hb_w_T(c, hb.T + ((size_t)mat * hb.P + panel) * HB_NB * HB_NB, hb.W + (size_t)mat * n * HB_NB, n - (k + HB_NB), blockIdx.x, true);
}

__global__ void k_hb_w_T_fused(HessBatch hb, int panel) {
  __shared__ cplx sS[HB_NB * HB_NB], sT[HB_NB * HB_NB];
  Cta c = make_cta(nullptr);
  const int mat = hb.mat0 + blockIdx.y;
  const int ilo = hb.ilohi[2 * mat], ihi = hb.ilohi[2 * mat + 1];
  const int k = ilo + panel * HB_NB;
  if (k >= ihi) return;
  for (int e = c.tid; e < HB_NB * HB_NB; e += c.nt) {
    sS[e] = hb.S[(size_t)mat * HB_NB * HB_NB + e];
    sT[e] = hb.T[((size_t)mat * hb.P + panel) * HB_NB * HB_NB + e];
  }
  cta_sync();
  cta_hb_w_T_fused(c, hb, mat, panel, blockIdx.x, sS, sT);
}

// ---- stage 3c: prepare the QR operand: Hq := upper Hessenberg part of A (zeros below), plus the
// infinity norm of H for the inverse-iteration tolerances.  Hq may alias A (eigenvalues-only path).
__global__ void k_prep_qr(const cplx* A, size_t astride, cplx* Hq, size_t hstride, int n, double* hnorm, int* blkend) {
  __shared__ double red[160];
  Cta c = make_cta(red);
  const int p = blockIdx.x;
  const cplx* a = A + (size_t)p * astride;
  cplx* h = Hq + (size_t)p * hstride;
  double mx = 0.0, bad = 0.0;                           // bad: a NaN somewhere (fmax would drop it)
  for (int r = c.tid; r < n; r += c.nt) {
    double s = 0.0;
    for (int j = (r > 0 ? r - 1 : 0); j < n; ++j) s += cabs(a[r + (size_t)j * n]);
    if (!(s == s)) bad = 1.0;
    mx = fmax(mx, s);
  }
  cta_max2(c, mx, bad);
  if (c.tid == 0) {
    hnorm[p] = (bad != 0.0) ? (mx - mx) / (mx - mx) : mx;   // NaN marks the matrix for the QR and eigenvector stages
    // diagonal blocks of H separated by exactly-zero subdiagonals (ZHSEIN's KL..KR with FROMQR):
    // blkend[i] = last index of the block containing i
    int end = n - 1;
    for (int i = n - 1; i >= 0; --i) {
      if (i < n - 1 && is_zero(a[(i + 1) + (size_t)i * n])) end = i;
      blkend[(size_t)p * n + i] = end;
    }
  }
  for (int j = 0; j < n; ++j)
    for (int r = c.tid; r < n; r += c.nt) {
      cplx v = a[r + (size_t)j * n];
      if (r > j + 1) v = mk(0.0, 0.0);
      else if (h == a) continue;
      h[r + (size_t)j * n] = v;
    }
}

// ---- stage 4: shifted QR: qr.cu (its own translation unit, launch.h) -------------------------------

// ---- stage 5: sort (stable, ascending imaginary part) --------------------------------------------
// mode 1 (temporal): key = Im(w).  mode 2 (spatial): alpha = 1/lambda (0 if lambda == 0), key = Im(alpha)
// (spatial.f90:1065-1084).  mode 0: no sort (generic solver).
// out: sorted eigenvalue (omega or alpha); lam: the matrix eigenvalue in the same order, perturbed
// like ZHSEIN does for (nearly) coincident values so inverse iteration yields independent vectors.
__global__ void k_sort(const cplx* w, int n, int mode, const double* hnorm, const int* blkend, cplx* out, cplx* lam, int* kr) {
  const int p = blockIdx.x;
  const cplx* wp = w + (size_t)p * n;
  cplx* op = out + (size_t)p * n;
  cplx* lp = lam + (size_t)p * n;
  const int* be = blkend + (size_t)p * n;
  const double eps3 = fmax(SD_ULP * hnorm[p], SD_SAFMIN * ((double)n / SD_ULP));
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    cplx wi = wp[i];
    cplx vi = wi;
    if (mode == 2) vi = is_zero(wi) ? mk(0.0, 0.0) : cdiv(mk(1.0, 0.0), wi);
    int rank = i;
    if (mode != 0) {
      rank = 0;
      const double key = vi.im;
      for (int j = 0; j < n; ++j) {
        cplx wj = wp[j];
        double kj = wj.im;
        if (mode == 2) kj = is_zero(wj) ? 0.0 : cdiv(mk(1.0, 0.0), wj).im;
        if (kj < key || (kj == key && j < i)) ++rank;
      }
    }
    int close = 0;   // earlier eigenvalues of the same diagonal block closer than eps3 (ZHSEIN)
    for (int j = 0; j < i; ++j) if (be[j] == be[i] && cabs1(wp[j] - wi) < eps3) ++close;
    op[rank] = vi;
    lp[rank] = wi + mk(eps3 * close, 0.0);
    kr[(size_t)p * n + rank] = be[i];
  }
}

// ---- stage 6: eigenvectors ----------------------------------------------------------------------
// grid (chunks, batch); each warp takes eigen-indices e = chunk*warps + wid, += chunks*warps
__global__ void k_evec(const cplx* Hh, size_t hstride, int n, const int* ilohi, const cplx* tau, const double* scale,
                       const cplx* lam, const int* kr, const double* hnorm, int scale_rows, cplx* V, size_t vstride, int* vinfo,
                       const int* badsel, int raw) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Cta w = make_cta(nullptr);
  const int p = blockIdx.y;
  cplx* cvec = reinterpret_cast<cplx*>(smem_raw) + (size_t)w.wid * 2 * n;
  cplx* yvec = cvec + n;
  unsigned char* flag = smem_raw + (size_t)w.nw * 2 * n * sizeof(cplx) + (size_t)w.wid * n;
  const int ilo = ilohi[2 * p], ihi = ilohi[2 * p + 1];
  int bad = 0;
  for (int e = blockIdx.x * w.nw + w.wid; e < n; e += gridDim.x * w.nw) {
    if (badsel && badsel[(size_t)p * n + e] == 0) continue;     // fallback pass: only vectors the fast kernel rejected
    bad += warp_eigvec(w, Hh + (size_t)p * hstride, n, n, ilo, ihi, tau + (size_t)p * n, scale + (size_t)p * n,
                       lam[(size_t)p * n + e], kr[(size_t)p * n + e], hnorm[p], scale_rows, cvec, yvec, flag,
                       V + (size_t)p * vstride + (size_t)e * n, raw != 0);
  }
  if (bad && w.lane == 0) atomicAdd(vinfo + p, bad);
}

// back-transformation GEMM phases (hess_blocked.cuh) and the T multiply
template <int PHASE, bool USE_MMA>
__global__ void __launch_bounds__(GEMM_THREADS, 2) k_bt_gemm(HessBatch hb, cplx* X, size_t xstride, int panel) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Cta c = make_cta(nullptr);
  cta_bt_gemm<PHASE, USE_MMA>(c, hb, X, xstride, hb.mat0 + blockIdx.z, panel, blockIdx.x, blockIdx.y, reinterpret_cast<double*>(smem_raw));
}

__global__ void k_bt_w_T(HessBatch hb, int panel) {
  Cta c = make_cta(nullptr);
  const int mat = hb.mat0 + blockIdx.y, n = hb.n;
  const int ilo = hb.ilohi[2 * mat], ihi = hb.ilohi[2 * mat + 1];
  if (ilo + panel * HB_NB >= ihi) return;
  cta_hb_w_T(c, hb.T + ((size_t)mat * hb.P + panel) * HB_NB * HB_NB, hb.W + (size_t)mat * n * HB_NB, n, blockIdx.x, false);
}

// ZGEBAK + ZGEEV normalisation [+ temporal.f90:867-879 scaling] of every column; warp per column
__global__ void k_vec_finalize(cplx* V, size_t vstride, int n, const int* ilohi, const double* scale, int scale_rows) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Cta w = make_cta(nullptr);
  const int p = blockIdx.y;
  cplx* x = reinterpret_cast<cplx*>(smem_raw) + (size_t)w.wid * n;
  const int ilo = ilohi[2 * p], ihi = ilohi[2 * p + 1];
  for (int e = blockIdx.x * w.nw + w.wid; e < n; e += gridDim.x * w.nw) {
    cplx* col = V + (size_t)p * vstride + (size_t)e * n;
    for (int r = w.lane; r < n; r += 32) x[r] = col[r];
    __syncwarp();
    warp_gebak(w, n, ilo, ihi, scale + (size_t)p * n, x);
    warp_normalize_zgeev(w, n, x);
    if (scale_rows > 0) warp_scale_maxabs(w, scale_rows, x);
    for (int r = w.lane; r < n; r += 32) col[r] = x[r];
    __syncwarp();
  }
}

}  // namespace stab
