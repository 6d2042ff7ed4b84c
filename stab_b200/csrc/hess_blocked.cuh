// hess_blocked.cuh -- batched BLOCKED Householder reduction to upper Hessenberg form,
// H = Q^H A Q, the ZGEHRD stage of the ZGEEV that the reference calls (temporal.f90:803,
// spatial.f90:1043).  Output layout is ZGEHRD's (H above, reflectors below the subdiagonal,
// tau[]), plus the triangular factors T of every panel (kept for the back-transformation).
//
// B200 design: the whole batch of sweep points advances in lock step through a fixed schedule of
// small kernels (captured once per plan), so every step is spread over all 148 SMs:
//   per column   k_hb_panel_step (one CTA per matrix: thin O(n*nb) updates, Householder vector)
//                k_hb_gemv       (grid = row tiles x column chunks x batch): y = A(k+1:, c+1:) v
//                                -> THE memory-bound kernel: 16 (ihi-k)(ihi-c) bytes per matrix
//   per panel    k_hb_gemm<...>  FP64 tensor-core (DMMA) rank-nb updates through gemm.cuh
// Algorithm: LAPACK's ZLAHR2/ZGEHRD recurrences (Y = A V T, delayed right update A -= Y V^H,
// left update with ZLARFB), restated with masks instead of LAPACK's in-place unit-diagonal
// save/restore; per-matrix (ilo, ihi) from the balancing stage are honoured inside the kernels.
#pragma once
#include "common.cuh"
#include "gemm.cuh"
#include "hessenberg.cuh"   // cta_zlarfg

namespace stab {

constexpr int HB_NB = 32;         // panel width
constexpr int HB_CHUNKS = 4;      // column chunks of the GEMV (deterministic split-K)
constexpr int HB_GEMV_ROWS = 128; // rows per GEMV CTA
constexpr int HB_DU = 5;          // loads in flight per lane in the panel-step dot products

struct HessBatch {
  cplx* A; size_t astride; int n;
  const int* ilohi;     // 2 per matrix (0-based, inclusive)
  cplx* tau;            // n per matrix
  cplx* Y;              // n x NB per matrix
  cplx* T;              // NB x NB per panel, P panels per matrix
  cplx* Ypart;          // n x CHUNKS per matrix
  cplx* W;              // NB x n per matrix
  int P;
  int mat0;             // first matrix of this launch group (the batch is split over two streams)
  cplx* Vx;             // n x NB per matrix: the current panel's V with explicit ones / zeros (pipelined GEMM path)
  cplx* VT = nullptr;   // n x NB per matrix: Vx T   (pipelined path: folds the triangular factor into the GEMM operand,
  cplx* VTh = nullptr;  // n x NB per matrix: Vx T^H  so that no separate pass applies T to Y_top / W)
  cplx* tv = nullptr;   // NB per matrix: t = V(:,0:j)^H v_j of the column whose GEMV is running (cta_hb_vdots)
  cplx* S = nullptr;    // NB x NB per matrix: V^H Y of the current panel (fused trailing update)
  cplx* Vh = nullptr;   // NB x n per matrix: conj(V(k+NB+j, :)) as a plain NB x nc operand (fused trailing update)
};

SD_DEV cplx hb_v(const cplx* A, int lda, int k, int ihi, int r, int l) {   // V(r, l) of the panel starting at k
  const int cc = k + l;
  if (r > ihi || r <= cc) return mk(0.0, 0.0);
  if (r == cc + 1) return mk(1.0, 0.0);
  return A[r + (size_t)cc * lda];
}

// One CTA per matrix.  Call j = 0..NB-1 before the GEMV of column j; call j = NB after the last
// GEMV of the panel.  smem: 160 doubles (reductions) + n + 3*NB complex.
SD_DEV void cta_hb_panel_step(const Cta& c, const HessBatch& hb, int mat, int panel, int j, double* red, cplx* sb, cplx* sw, cplx* st, cplx* sc) {
  const int n = hb.n, lda = n;
  cplx* A = hb.A + (size_t)mat * hb.astride;
  const int ilo = hb.ilohi[2 * mat], ihi = hb.ilohi[2 * mat + 1];
  const int k = ilo + panel * HB_NB;
  cplx* tau = hb.tau + (size_t)mat * n;
  if (panel == 0 && j == 0) {
    for (int q = c.tid; q < n; q += c.nt) tau[q] = mk(0.0, 0.0);
    cta_sync();
  }
  if (k >= ihi) return;
  cplx* Y = hb.Y + (size_t)mat * n * HB_NB;
  cplx* T = hb.T + ((size_t)mat * hb.P + panel) * HB_NB * HB_NB;
  const cplx* Yp = hb.Ypart + (size_t)mat * n * HB_CHUNKS;
  if (j == 0) {
    for (int q = c.tid; q < HB_NB * HB_NB; q += c.nt) T[q] = mk(0.0, 0.0);
    cta_sync();
  }
  // ---- finish column jp = j-1: Y(:,jp), T(:,jp) (ZLAHR2's post-GEMV part), fused with the first update of
  // column `col` (b -= Y(:,0:j) conj(V(col,0:j))^T): ONE pass over the rows of Y serves both ----
  const int col = k + j;
  const bool have_col = (j < HB_NB) && (col < n);
  cplx* acol = A + (size_t)col * lda;
  const int nr = ihi - k;                                   // rows k+1..ihi, local index r-(k+1)
  if (j > 0) {
    const int jp = j - 1, cp = k + jp;
    const bool fin = cp < ihi;
    const cplx taup = fin ? tau[cp] : mk(0.0, 0.0);
    if (fin) {                                             // t = V(:,0:jp)^H v_jp: computed beside the GEMV (cta_hb_vdots)
      const cplx* tv = hb.tv + (size_t)mat * HB_NB;
      for (int l = c.tid; l < jp; l += c.nt) st[l] = tv[l];
    }
    if (have_col)
      for (int l = c.tid; l < j; l += c.nt) sc[l] = conj(hb_v(A, lda, k, ihi, col, l));
    cta_sync();
    for (int r = k + 1 + c.tid; r <= ihi; r += c.nt) {
      cplx yjp = mk(0.0, 0.0);
      cplx b = have_col ? acol[r] : mk(0.0, 0.0);
      if (fin) {                                            // Y(:,jp) = tau (A v - Y(:,0:jp) t)
        cplx s = mk(0.0, 0.0);
        for (int ch = 0; ch < HB_CHUNKS; ++ch) s += Yp[r + (size_t)ch * n];
#pragma unroll 4
        for (int l = 0; l < jp; ++l) {
          const cplx y = Y[r + (size_t)l * n];
          s -= y * st[l];
          b -= y * sc[l];
        }
        yjp = taup * s;
      } else if (have_col) {
        for (int l = 0; l < jp; ++l) b -= Y[r + (size_t)l * n] * sc[l];
      }
      Y[r + (size_t)jp * n] = yjp;
      if (have_col) { b -= yjp * sc[jp]; sb[r - (k + 1)] = b; }   // b -= Y(:,0:j) conj(V(col,0:j))^T
    }
    if (fin) {
      for (int l = c.tid; l <= jp; l += c.nt) {            // T(0:jp,jp) = -tau T(0:jp,0:jp) t ; T(jp,jp) = tau
        if (l == jp) { T[l + jp * HB_NB] = taup; continue; }
        cplx s = mk(0.0, 0.0);
        for (int m = l; m < jp; ++m) fma_acc(s, T[l + m * HB_NB], st[m]);
        T[l + jp * HB_NB] = -(taup * s);
      }
    }
    cta_sync();
  }
  if (!have_col) return;
  // ---- column `col`: the panel's previous reflectors from the left, then generate H(col) ----
  if (j > 0) {
    for (int l = c.wid; l < j; l += c.nw) {                 // w = V^H b
      const cplx* vl = A + (size_t)(k + l) * lda;
      cplx s = mk(0.0, 0.0);
      for (int r0 = k + l + 1 + c.lane; r0 <= ihi; r0 += HB_DU * c.ws) {
        cplx a[HB_DU];
#pragma unroll
        for (int u = 0; u < HB_DU; ++u) {
          const int r = r0 + u * c.ws;
          a[u] = (r <= ihi) ? ((r == k + l + 1) ? mk(1.0, 0.0) : vl[r]) : mk(0.0, 0.0);
        }
#pragma unroll
        for (int u = 0; u < HB_DU; ++u) {
          const int r = r0 + u * c.ws;
          if (r <= ihi) fma_acc_conj(s, a[u], sb[r - (k + 1)]);
        }
      }
      s = warp_sum(s);
      if (c.lane == 0) st[l] = s;
    }
    cta_sync();
    for (int l = c.tid; l < j; l += c.nt) {                 // w = T^H w
      cplx s = mk(0.0, 0.0);
      for (int m = 0; m <= l; ++m) fma_acc_conj(s, T[m + l * HB_NB], st[m]);
      sw[l] = s;
    }
    cta_sync();
    for (int q = c.tid; q < nr; q += c.nt) {                // b -= V w
      const int r = k + 1 + q;
      cplx s = sb[q];
      const int lmax = (q + 1 < j) ? q + 1 : j;              // reflector l reaches rows r > k + l, i.e. l <= q
#pragma unroll 4
      for (int l = 0; l < lmax; ++l) {
        const cplx vr = (l == q) ? mk(1.0, 0.0) : A[r + (size_t)(k + l) * lda];
        s -= vr * sw[l];
      }
      acol[r] = s;
    }
    cta_sync();
  }
  if (col < ihi) {
    cplx alpha = acol[col + 1];
    cta_sync();
    cplx tj = cta_zlarfg(c, ihi - col, alpha, acol + col + 2);
    cta_sync();
    if (c.tid == 0) { tau[col] = tj; acol[col + 1] = alpha; }
  }
}

// y_chunk(r) = sum_{cc in chunk} A(r, cc) v(cc), r in the row tile; v = reflector of column `col`.
// grid: (row tiles, HB_CHUNKS, batch); block HB_GEMV_ROWS threads; sv: (n) complex shared.
SD_DEV void cta_hb_gemv(const Cta& c, const HessBatch& hb, int mat, int panel, int j, int rowtile, int chunk, cplx* sv) {
  const int n = hb.n, lda = n;
  const cplx* A = hb.A + (size_t)mat * hb.astride;
  const int ilo = hb.ilohi[2 * mat], ihi = hb.ilohi[2 * mat + 1];
  const int k = ilo + panel * HB_NB, col = k + j;
  if (k >= ihi || col >= ihi) return;
  const int r0 = k + 1 + rowtile * HB_GEMV_ROWS;
  if (r0 > ihi) return;
  const int ncols = ihi - col;                              // columns col+1..ihi
  const int per = (ncols + HB_CHUNKS - 1) / HB_CHUNKS;
  const int c0 = col + 1 + chunk * per;
  int c1 = c0 + per; if (c1 > ihi + 1) c1 = ihi + 1;
  const cplx* vcol = A + (size_t)col * lda;
  for (int q = c0 + c.tid; q < c1; q += c.nt) sv[q - c0] = (q == col + 1) ? mk(1.0, 0.0) : vcol[q];
  cta_sync();
  cplx* Yp = hb.Ypart + (size_t)mat * n * HB_CHUNKS + (size_t)chunk * n;
  for (int r = r0 + c.tid; r <= ihi && r < r0 + HB_GEMV_ROWS; r += c.nt) {
    cplx a0 = mk(0.0, 0.0), a1 = a0, a2 = a0, a3 = a0;
    const cplx* ap = A + r;
    int cc = c0;
#ifdef STAB_EMU
    for (; cc + 3 < c1; cc += 4) {
      const cplx x0 = ap[(size_t)cc * lda], x1 = ap[(size_t)(cc + 1) * lda], x2 = ap[(size_t)(cc + 2) * lda], x3 = ap[(size_t)(cc + 3) * lda];
      fma_acc(a0, x0, sv[cc - c0]); fma_acc(a1, x1, sv[cc + 1 - c0]); fma_acc(a2, x2, sv[cc + 2 - c0]); fma_acc(a3, x3, sv[cc + 3 - c0]);
    }
    for (; cc < c1; ++cc) fma_acc(a0, ap[(size_t)cc * lda], sv[cc - c0]);
#else
    const unsigned long long pol = l2_policy_evict_first();   // the trailing matrix streams through once: keep the panels in L2
    for (; cc + 3 < c1; cc += 4) {
      const cplx x0 = ld_stream(ap + (size_t)cc * lda, pol), x1 = ld_stream(ap + (size_t)(cc + 1) * lda, pol),
                 x2 = ld_stream(ap + (size_t)(cc + 2) * lda, pol), x3 = ld_stream(ap + (size_t)(cc + 3) * lda, pol);
      fma_acc(a0, x0, sv[cc - c0]); fma_acc(a1, x1, sv[cc + 1 - c0]); fma_acc(a2, x2, sv[cc + 2 - c0]); fma_acc(a3, x3, sv[cc + 3 - c0]);
    }
    for (; cc < c1; ++cc) fma_acc(a0, ld_stream(ap + (size_t)cc * lda, pol), sv[cc - c0]);
#endif
    Yp[r] = (a0 + a1) + (a2 + a3);
  }
}

// t = V(:,0:j)^H v_j for the column whose GEMV is in flight (ZLAHR2's  T(0:j,j) = -tau T V^H v  and  Y(:,j) -= Y t  need
// it only AFTER the GEMV): one extra CTA per matrix in the GEMV launch, so that this pass over V runs in the shadow of the
// bandwidth-bound GEMV instead of in the latency-bound panel step.  One warp per column l, HB_DU loads in flight per lane.
SD_DEV void cta_hb_vdots(const Cta& c, const HessBatch& hb, int mat, int panel, int j) {
  const int n = hb.n, lda = n;
  const cplx* A = hb.A + (size_t)mat * hb.astride;
  const int ilo = hb.ilohi[2 * mat], ihi = hb.ilohi[2 * mat + 1];
  const int k = ilo + panel * HB_NB, cp = k + j;
  if (k >= ihi || cp >= ihi) return;
  cplx* tv = hb.tv + (size_t)mat * HB_NB;
  const cplx* vcol = A + (size_t)cp * lda;
  for (int l = c.wid; l < j; l += c.nw) {
    const cplx* vl = A + (size_t)(k + l) * lda;
    cplx s = mk(0.0, 0.0);
    for (int r0 = cp + 1 + c.lane; r0 <= ihi; r0 += HB_DU * c.ws) {
      cplx a[HB_DU], b[HB_DU];
#pragma unroll
      for (int u = 0; u < HB_DU; ++u) {
        const int r = r0 + u * c.ws;
        a[u] = (r <= ihi) ? vl[r] : mk(0.0, 0.0);
        b[u] = (r <= ihi) ? ((r == cp + 1) ? mk(1.0, 0.0) : vcol[r]) : mk(0.0, 0.0);
      }
#pragma unroll
      for (int u = 0; u < HB_DU; ++u) fma_acc_conj(s, a[u], b[u]);
    }
    s = warp_sum(s);
    if (c.lane == 0) tv[l] = s;
  }
}

// ---- per-panel level-3 phases --------------------------------------------------------------------
enum HbPhase { HB_YTOP = 0, HB_RIGHT_TRAIL = 1, HB_RIGHT_PANEL = 2, HB_LEFT_W = 3, HB_LEFT_UPD = 4 };

// One tile of one phase for one matrix.  Returns without work when the tile is out of range.
template <int PHASE, bool USE_MMA>
SD_DEV void cta_hb_gemm(const Cta& c, const HessBatch& hb, int mat, int panel, int ti, int tj, double* smem) {
  const int n = hb.n, lda = n;
  cplx* A = hb.A + (size_t)mat * hb.astride;
  const int ilo = hb.ilohi[2 * mat], ihi = hb.ilohi[2 * mat + 1];
  const int k = ilo + panel * HB_NB;
  if (k >= ihi) return;
  cplx* Y = hb.Y + (size_t)mat * n * HB_NB;
  cplx* W = hb.W + (size_t)mat * n * HB_NB;
  if (PHASE == HB_YTOP) {            // Y(0:k+1, :) = A(0:k+1, k+1:ihi+1) V(k+1:ihi+1, :)        (then * T, separate kernel)
    const int m = k + 1, nc = HB_NB, K = ihi - k;
    if (ti * 64 >= m || tj > 0) return;
    OpL_ColMajor L{A + (size_t)(k + 1) * lda, lda};
    OpR_V R{VBlock{A, lda, k, ihi, k + 1}};
    cta_gemm_tile<64, 32, false, USE_MMA>(c, smem, ti * 64, 0, m, nc, K, L, R, Y, n);
  } else if (PHASE == HB_RIGHT_TRAIL) {   // A(0:ihi+1, k+NB:ihi+1) -= Y V(k+NB:ihi+1, :)^H
    const int m = ihi + 1, nc = ihi + 1 - (k + HB_NB);
    if (nc <= 0 || ti * 64 >= m || tj * 64 >= nc) return;
    OpL_ColMajor L{Y, n};
    OpR_VH R{VBlock{A, lda, k, ihi, k + HB_NB}};
    cta_gemm_tile<64, 64, true, USE_MMA>(c, smem, ti * 64, tj * 64, m, nc, HB_NB, L, R, A + (size_t)(k + HB_NB) * lda, lda);
  } else if (PHASE == HB_RIGHT_PANEL) {   // A(0:k+1, k+1:k+NB) -= Y(0:k+1, :) V(k+1:k+NB, :)^H
    const int m = k + 1;
    int nc = HB_NB - 1; if (nc > n - (k + 1)) nc = n - (k + 1);
    if (ti * 64 >= m || tj > 0 || nc <= 0) return;
    OpL_ColMajor L{Y, n};
    OpR_VH R{VBlock{A, lda, k, ihi, k + 1}};
    cta_gemm_tile<64, 32, true, USE_MMA>(c, smem, ti * 64, 0, m, nc, HB_NB, L, R, A + (size_t)(k + 1) * lda, lda);
  } else if (PHASE == HB_LEFT_W) {        // W(NB x ncF) = V^H A(k+1:ihi+1, k+NB:n)            (then T^H *, separate kernel)
    const int nc = n - (k + HB_NB), K = ihi - k;
    if (nc <= 0 || ti > 0 || tj * 64 >= nc) return;
    OpL_VH L{VBlock{A, lda, k, ihi, k + 1}};
    OpR_ColMajor R{A + (k + 1) + (size_t)(k + HB_NB) * lda, lda};
    cta_gemm_tile<32, 64, false, USE_MMA>(c, smem, 0, tj * 64, HB_NB, nc, K, L, R, W, HB_NB);
  } else {                                // A(k+1:ihi+1, k+NB:n) -= V W
    const int m = ihi - k, nc = n - (k + HB_NB);
    if (nc <= 0 || ti * 64 >= m || tj * 64 >= nc) return;
    OpL_V L{VBlock{A, lda, k, ihi, k + 1}};
    OpR_ColMajor R{W, HB_NB};
    cta_gemm_tile<64, 64, true, USE_MMA>(c, smem, ti * 64, tj * 64, m, nc, HB_NB, L, R, A + (k + 1) + (size_t)(k + HB_NB) * lda, lda);
  }
}

// ---- back-transformation of the eigenvectors: X <- Q X, Q = Q_0 Q_1 ... (ZUNMHR), one panel ----
// X(k+1:ihi+1, :) -= V (T (V^H X(k+1:ihi+1, :))) for panels in DECREASING order.
enum BtPhase { BT_W = 0, BT_UPD = 1 };
template <int PHASE, bool USE_MMA>
SD_DEV void cta_bt_gemm(const Cta& c, const HessBatch& hb, cplx* Xall, size_t xstride, int mat, int panel, int ti, int tj, double* smem) {
  const int n = hb.n, lda = n;
  const cplx* A = hb.A + (size_t)mat * hb.astride;
  cplx* X = Xall + (size_t)mat * xstride;
  const int ilo = hb.ilohi[2 * mat], ihi = hb.ilohi[2 * mat + 1];
  const int k = ilo + panel * HB_NB;
  if (k >= ihi) return;
  cplx* W = hb.W + (size_t)mat * n * HB_NB;
  if (PHASE == BT_W) {
    if (ti > 0 || tj * 64 >= n) return;
    OpL_VH L{VBlock{A, lda, k, ihi, k + 1}};
    OpR_ColMajor R{X + (k + 1), n};
    cta_gemm_tile<32, 64, false, USE_MMA>(c, smem, 0, tj * 64, HB_NB, n, ihi - k, L, R, W, HB_NB);
  } else {
    const int m = ihi - k;
    if (ti * 64 >= m || tj * 64 >= n) return;
    OpL_V L{VBlock{A, lda, k, ihi, k + 1}};
    OpR_ColMajor R{W, HB_NB};
    cta_gemm_tile<64, 64, true, USE_MMA>(c, smem, ti * 64, tj * 64, m, n, HB_NB, L, R, X + (k + 1), n);
  }
}

// Y(0:k+1, :) *= T  (upper triangular, right multiplication); thread per row
SD_DEV void cta_hb_ytop_T(const Cta& c, const HessBatch& hb, int mat, int panel, int rowblock) {
  const int n = hb.n;
  const int ilo = hb.ilohi[2 * mat], ihi = hb.ilohi[2 * mat + 1];
  const int k = ilo + panel * HB_NB;
  if (k >= ihi) return;
  cplx* Y = hb.Y + (size_t)mat * n * HB_NB;
  const cplx* T = hb.T + ((size_t)mat * hb.P + panel) * HB_NB * HB_NB;
  const int r = rowblock * c.nt + c.tid;
  if (r <= k) {
    cplx y[HB_NB];
    for (int l = 0; l < HB_NB; ++l) y[l] = Y[r + (size_t)l * n];
    for (int l = HB_NB - 1; l >= 0; --l) {
      cplx s = mk(0.0, 0.0);
      for (int m = 0; m <= l; ++m) fma_acc(s, y[m], T[m + l * HB_NB]);
      Y[r + (size_t)l * n] = s;
    }
  }
}

// W = op(T) W for the NB x ncols block W (ld NB); conjT: T^H (Hessenberg left update), else T
// (back-transformation).  thread per column.
SD_DEV void cta_hb_w_T(const Cta& c, const cplx* T, cplx* W, int ncols, int colblock, bool conjT) {
  const int j = colblock * c.nt + c.tid;
  if (j >= ncols) return;
  cplx w[HB_NB];
  for (int l = 0; l < HB_NB; ++l) w[l] = W[l + (size_t)j * HB_NB];
  if (conjT) {
    for (int l = HB_NB - 1; l >= 0; --l) {      // (T^H w)[l] = sum_{m<=l} conj(T[m,l]) w[m]
      cplx s = mk(0.0, 0.0);
      for (int m = 0; m <= l; ++m) fma_acc_conj(s, T[m + l * HB_NB], w[m]);
      W[l + (size_t)j * HB_NB] = s;
    }
  } else {
    for (int l = 0; l < HB_NB; ++l) {           // (T w)[l] = sum_{m>=l} T[l,m] w[m]
      cplx s = mk(0.0, 0.0);
      for (int m = l; m < HB_NB; ++m) fma_acc(s, T[l + m * HB_NB], w[m]);
      W[l + (size_t)j * HB_NB] = s;
    }
  }
}

// Fused trailing update (hess_mode 5): column j of the trailing block (global column c = k + NB + j).
//   w0 = V^H A_old(:, c)   (from the W-product on the NOT yet right-updated matrix)
//   W(:, j)  = T^H ( w0 - S conj(V(c, :))^T ),   S = V^H Y      -- equals T^H V^H (A_old - Y V^H)(:, c)
//   Vh(:, j) = conj(V(c, :))^T                                  -- the right update's operand as a plain NB x nc matrix
SD_DEV void cta_hb_w_T_fused(const Cta& c, const HessBatch& hb, int mat, int panel, int colblock, const cplx* sS, const cplx* sT) {
  const int n = hb.n;
  const int ilo = hb.ilohi[2 * mat], ihi = hb.ilohi[2 * mat + 1];
  const int k = ilo + panel * HB_NB;
  const int ncols = n - (k + HB_NB);
  const int j = colblock * c.nt + c.tid;
  if (j >= ncols) return;
  const cplx* A = hb.A + (size_t)mat * hb.astride;
  cplx* W = hb.W + (size_t)mat * n * HB_NB;
  cplx* Vh = hb.Vh + (size_t)mat * n * HB_NB;
  const int col = k + HB_NB + j;
  cplx w[HB_NB], v[HB_NB];
  for (int l = 0; l < HB_NB; ++l) { w[l] = W[l + (size_t)j * HB_NB]; v[l] = conj(hb_v(A, n, k, ihi, col, l)); }
  for (int l = 0; l < HB_NB; ++l) Vh[l + (size_t)j * HB_NB] = v[l];
  if (col <= ihi)
    for (int i = 0; i < HB_NB; ++i) {
      cplx s = w[i];
      for (int l = 0; l < HB_NB; ++l) fms_acc(s, sS[i + l * HB_NB], v[l]);
      w[i] = s;
    }
  for (int l = HB_NB - 1; l >= 0; --l) {        // (T^H w)[l] = sum_{m<=l} conj(T[m,l]) w[m]
    cplx s = mk(0.0, 0.0);
    for (int m = 0; m <= l; ++m) fma_acc_conj(s, sT[m + l * HB_NB], w[m]);
    W[l + (size_t)j * HB_NB] = s;
  }
}

}  // namespace stab
