// gemm.cuh -- CTA-tile complex GEMM on the FP64 tensor cores:
//     C (m x nc)  =  [C -]  L (m x K) * R (K x nc)           (complex double, C column-major)
// This is the one level-3 building block of the package: the trailing updates of the blocked
// Hessenberg reduction (ZGEHRD's ZGEMM / ZLARFB calls inside the ZGEEV of temporal.f90:803,
// spatial.f90:1043), the Y_top = A V T product of ZLAHR2, and the back-transformation of the
// eigenvectors (ZUNMHR-like) all go through it with different operand functors.
//
// Tensor-core mapping (sm_100a has no tcgen05 kind for f64; FP64 tensor math is the register
// fragment DMMA `mma.sync.m8n8k4.f64`).  A complex product is a real one on the 2x2 embedding
// [[a,-b],[b,a]]:  with C stored interleaved (re,im) we compute the TRANSPOSED real product
//     D[j, 2i+p] += sum_kk  Rst[kk, j] * Lemb[2i+p, kk],     kk = 2l+q,
//     Rst[2l+q, j] = (re,im)[q] of R(l,j),   Lemb[2i+p, 2l+q] = {re, -im; im, re}[p][q] of L(i,l)
// so that MMA rows <-> 8 columns j of C and MMA columns <-> 4 complex rows i of C: the two
// accumulator registers of a thread are exactly (re, im) of ONE complex entry of C, and both
// operand fragments are single 8-byte shared-memory loads (conflict-free with the +8 padding).
// Operand functors apply conjugation / the implicit unit-lower-triangular structure of the
// Householder block V at tile-load time, so the tiles in shared memory hold effective values.
#pragma once
#include "common.cuh"

namespace stab {

constexpr int GEMM_KC = 32;       // K chunk staged in shared memory
constexpr int GEMM_THREADS = 256; // 8 warps

template <int TM, int TN>
struct GemmCfg {
  static constexpr int SLD = 2 * TM + 8;   // doubles per k-row of the L tile
  static constexpr int SRD = 2 * TN + 8;
  static constexpr size_t smem_bytes = sizeof(double) * GEMM_KC * (SLD + SRD);
  static constexpr int WR = TM / 16;       // warp grid: WR row groups of 16 rows
  static constexpr int WC = 8 / WR;        //            WC column groups
  static constexpr int CPW = TN / WC;      // columns per warp
  static constexpr int MT = CPW / 8;       // MMA tiles along columns per warp
  static constexpr int NT = 4;             // MMA tiles along rows per warp (16 rows / 4)
};

#ifndef STAB_EMU
SD_DEV void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}
#endif

// L functor: cplx operator()(int i, int l) const, static constexpr bool kfast (true: memory runs
// fastest along l).  R functor: cplx operator()(int l, int j) const, kfast likewise.
// Computes the TM x TN tile at (i0, j0); rows >= m / columns >= nc are masked.
// SUB: C -= L*R, else C = L*R.   smem: GemmCfg<TM,TN>::smem_bytes, 16-byte aligned.
template <int TM, int TN, bool SUB, bool USE_MMA, class LOp, class ROp>
SD_DEV void cta_gemm_tile(const Cta& c, double* smem, int i0, int j0, int m, int nc, int K, const LOp& L, const ROp& R,
                          cplx* C, int ldc) {
  typedef GemmCfg<TM, TN> G;
  double* sL = smem;
  double* sR = smem + GEMM_KC * G::SLD;
#ifdef STAB_EMU
  constexpr bool mma = false;
  constexpr int NACC = TM * TN;
#else
  constexpr bool mma = USE_MMA;
  constexpr int NACC = USE_MMA ? G::MT * G::NT : (TM * TN) / GEMM_THREADS;
#endif
  cplx acc[NACC];
#pragma unroll
  for (int q = 0; q < NACC; ++q) acc[q] = mk(0.0, 0.0);
#ifndef STAB_EMU
  const int lane = c.lane, g = lane >> 2, tg = lane & 3;
  const int rowbase = (c.wid % G::WR) * 16, colbase = (c.wid / G::WR) * G::CPW;
  const int lsel = tg >> 1, qsel = tg & 1;
  const int bx = (g & 1) ^ qsel;                 // which component of L the B fragment needs
  const bool bneg = (((g & 1) == 0) && (qsel == 1)) != (mma && SUB);   // SUB: accumulate C - L*R directly (L negated)
  if (mma && SUB) {
    // the accumulators start from C: its loads overlap the operand tile loads instead of trailing the MMAs
#pragma unroll
    for (int mt = 0; mt < G::MT; ++mt)
#pragma unroll
      for (int nt = 0; nt < G::NT; ++nt) {
        const int i = i0 + rowbase + nt * 4 + tg, j = j0 + colbase + mt * 8 + g;
        if (i < m && j < nc) acc[mt * G::NT + nt] = C[i + (size_t)j * ldc];
      }
  }
#endif
  for (int k0 = 0; k0 < K; k0 += GEMM_KC) {
#ifndef STAB_EMU
    {
      // all loads of both operand tiles are issued before the first shared-memory store (memory-level parallelism:
      // the tile load is one DRAM round trip, not one per element)
      constexpr int NLD = (TM * GEMM_KC) / GEMM_THREADS, NRD = (TN * GEMM_KC) / GEMM_THREADS;
      cplx vl[NLD], vr[NRD];
#pragma unroll
      for (int u = 0; u < NLD; ++u) {
        const int idx = c.tid + u * GEMM_THREADS;
        int i, l;
        // kfast operands: 4 consecutive k (64 B of memory) x 8 rows per warp -> full sectors AND conflict-free smem stores
        if (LOp::kfast) { l = (idx & 3) + 4 * (idx / (4 * TM)); i = (idx >> 2) % TM; } else { i = idx % TM; l = idx / TM; }
        vl[u] = mk(0.0, 0.0);
        if (i0 + i < m && k0 + l < K) vl[u] = L(i0 + i, k0 + l);
      }
#pragma unroll
      for (int u = 0; u < NRD; ++u) {
        const int idx = c.tid + u * GEMM_THREADS;
        int j, l;
        if (ROp::kfast) { l = (idx & 3) + 4 * (idx / (4 * TN)); j = (idx >> 2) % TN; } else { j = idx % TN; l = idx / TN; }
        vr[u] = mk(0.0, 0.0);
        if (j0 + j < nc && k0 + l < K) vr[u] = R(k0 + l, j0 + j);
      }
#pragma unroll
      for (int u = 0; u < NLD; ++u) {
        const int idx = c.tid + u * GEMM_THREADS;
        int i, l;
        if (LOp::kfast) { l = (idx & 3) + 4 * (idx / (4 * TM)); i = (idx >> 2) % TM; } else { i = idx % TM; l = idx / TM; }
        sL[l * G::SLD + 2 * i] = vl[u].re; sL[l * G::SLD + 2 * i + 1] = vl[u].im;
      }
#pragma unroll
      for (int u = 0; u < NRD; ++u) {
        const int idx = c.tid + u * GEMM_THREADS;
        int j, l;
        if (ROp::kfast) { l = (idx & 3) + 4 * (idx / (4 * TN)); j = (idx >> 2) % TN; } else { j = idx % TN; l = idx / TN; }
        sR[l * G::SRD + 2 * j] = vr[u].re; sR[l * G::SRD + 2 * j + 1] = vr[u].im;
      }
    }
#else
    for (int idx = c.tid; idx < TM * GEMM_KC; idx += c.nt) {
      int i, l;
      if (LOp::kfast) { l = (idx & 3) + 4 * (idx / (4 * TM)); i = (idx >> 2) % TM; } else { i = idx % TM; l = idx / TM; }
      cplx v = mk(0.0, 0.0);
      if (i0 + i < m && k0 + l < K) v = L(i0 + i, k0 + l);
      sL[l * G::SLD + 2 * i] = v.re; sL[l * G::SLD + 2 * i + 1] = v.im;
    }
    for (int idx = c.tid; idx < TN * GEMM_KC; idx += c.nt) {
      int j, l;
      if (ROp::kfast) { l = (idx & 3) + 4 * (idx / (4 * TN)); j = (idx >> 2) % TN; } else { j = idx % TN; l = idx / TN; }
      cplx v = mk(0.0, 0.0);
      if (j0 + j < nc && k0 + l < K) v = R(k0 + l, j0 + j);
      sR[l * G::SRD + 2 * j] = v.re; sR[l * G::SRD + 2 * j + 1] = v.im;
    }
#endif
    cta_sync();
    if (mma) {
#ifndef STAB_EMU
#pragma unroll 4
      for (int ks = 0; ks < GEMM_KC / 2; ++ks) {
        const double* rl = sR + (2 * ks + lsel) * G::SRD + qsel;
        const double* ll = sL + (2 * ks + lsel) * G::SLD + bx;
        double a[G::MT], b[G::NT];
#pragma unroll
        for (int mt = 0; mt < G::MT; ++mt) a[mt] = rl[2 * (colbase + mt * 8 + g)];
#pragma unroll
        for (int nt = 0; nt < G::NT; ++nt) {
          double v = ll[2 * (rowbase + nt * 4 + (g >> 1))];
          b[nt] = bneg ? -v : v;
        }
#pragma unroll
        for (int mt = 0; mt < G::MT; ++mt)
#pragma unroll
          for (int nt = 0; nt < G::NT; ++nt) dmma884(acc[mt * G::NT + nt].re, acc[mt * G::NT + nt].im, a[mt], b[nt]);
      }
#endif
    } else {
      int q = 0;
      for (int idx = c.tid; idx < TM * TN; idx += c.nt, ++q) {
        const int i = idx % TM, j = idx / TM;
        cplx s = acc[q];
        for (int l = 0; l < GEMM_KC; ++l)
          fma_acc(s, mk(sL[l * G::SLD + 2 * i], sL[l * G::SLD + 2 * i + 1]), mk(sR[l * G::SRD + 2 * j], sR[l * G::SRD + 2 * j + 1]));
        acc[q] = s;
      }
    }
    cta_sync();
  }
  if (mma) {
#ifndef STAB_EMU
#pragma unroll
    for (int mt = 0; mt < G::MT; ++mt)
#pragma unroll
      for (int nt = 0; nt < G::NT; ++nt) {
        const int i = i0 + rowbase + nt * 4 + tg, j = j0 + colbase + mt * 8 + g;
        if (i < m && j < nc) {
          C[i + (size_t)j * ldc] = acc[mt * G::NT + nt];
        }
      }
#endif
  } else {
    int q = 0;
    for (int idx = c.tid; idx < TM * TN; idx += c.nt, ++q) {
      const int i = i0 + idx % TM, j = j0 + idx / TM;
      if (i < m && j < nc) {
        cplx* p = C + i + (size_t)j * ldc;
        *p = SUB ? (*p - acc[q]) : acc[q];
      }
    }
  }
}

// ---- operand functors --------------------------------------------------------------------------
// plain column-major matrix P(a, b) = p[a + b*ld]
struct OpL_ColMajor {          // L(i,l) = P(i,l): memory fastest along i
  const cplx* p; int ld;
  static constexpr bool kfast = false;
  SD_DEV cplx operator()(int i, int l) const { return p[i + (size_t)l * ld]; }
};
struct OpR_ColMajor {          // R(l,j) = P(l,j): memory fastest along l
  const cplx* p; int ld;
  static constexpr bool kfast = true;
  SD_DEV cplx operator()(int l, int j) const { return p[l + (size_t)j * ld]; }
};

// The Householder block of one panel, stored ZGEHRD-style inside A: reflector jj of the panel that
// starts at column k lives in column c = k+jj, has an implicit 1 at row c+1, its tail in rows
// c+2..ihi, zeros elsewhere.  r0 = global row of local index 0.
struct VBlock {
  const cplx* A; int lda, k, ihi, r0;
  SD_DEV cplx at(int rloc, int jj) const {
    const int r = r0 + rloc, cc = k + jj;
    if (r > ihi || r <= cc) return mk(0.0, 0.0);
    if (r == cc + 1) return mk(1.0, 0.0);
    return A[r + (size_t)cc * lda];
  }
};
struct OpL_V {                 // L(i,l) = V(i,l)
  VBlock v; static constexpr bool kfast = false;
  SD_DEV cplx operator()(int i, int l) const { return v.at(i, l); }
};
struct OpL_VH {                // L(i,l) = conj(V(l,i))       (V^H, reduction over rows)
  VBlock v; static constexpr bool kfast = true;
  SD_DEV cplx operator()(int i, int l) const { return conj(v.at(l, i)); }
};
struct OpR_V {                 // R(l,j) = V(l,j)
  VBlock v; static constexpr bool kfast = true;
  SD_DEV cplx operator()(int l, int j) const { return v.at(l, j); }
};
struct OpR_VH {                // R(l,j) = conj(V(j,l))
  VBlock v; static constexpr bool kfast = false;
  SD_DEV cplx operator()(int l, int j) const { return conj(v.at(j, l)); }
};

}  // namespace stab
