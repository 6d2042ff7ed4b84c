// gemm_pipe.cuh -- persistent, software-pipelined complex GEMM on the FP64 tensor cores (GPU only).
//
//     C (m x nc)  =  [C -]  op(L) (m x K) * op(R) (K x nc)        complex double, column-major C
//
// Same tensor-core mapping as gemm.cuh (DMMA m8n8k4 on the transposed 2x2 real embedding, the two
// accumulator registers of a thread are (re, im) of one entry of C), different data movement:
//   * every operand is a PLAIN strided matrix (the Householder block V is materialised once per
//     panel with its implicit ones and zeros, k_hb_vx), so the 32-deep operand chunks travel
//     global -> shared memory by cp.async (LDGSTS, 16 B = one complex entry, zero-filled out of
//     range) with no register staging; conjugation is a sign flip at fragment-load time;
//   * a CTA is persistent: it walks a list of (tile, K-chunk) work items with a two-stage
//     shared-memory ring -- the chunk of item i+1 is in flight while item i is multiplied;
//   * for the rank-32 updates (SUB) the C tile of the NEXT tile is prefetched into registers
//     while the current tile is multiplied, so neither the operand nor the C round trip to HBM/L2
//     is exposed; one barrier per work item.
// Two CTAs of 256 threads per SM (2 x 2 x 53 KB of shared memory, <= 128 registers).
#pragma once
#include "common.cuh"
#include "gemm.cuh"

#ifndef STAB_EMU
namespace stab {

struct GemmProb {
  const cplx* L; long long lsi, lsl;   // L(i,l) = L[i*lsi + l*lsl]
  const cplx* R; long long rsl, rsj;   // R(l,j) = R[l*rsl + j*rsj]
  cplx* C; int ldc;
  int m, nc, K;                        // m <= 0 or nc <= 0 or K <= 0: nothing to do
  // optional second operand pair with the same strides: C -= L R + L2 R2 in ONE pass over C (the K-chunks of the
  // second pair follow those of the first in the work list)
  const cplx* L2 = nullptr; const cplx* R2 = nullptr; int K2 = 0;
};

template <int TM, int TN>
struct PipeCfg {
  static constexpr int SLD = 2 * TM + 8;   // doubles per k-row of the L tile
  static constexpr int SRD = 2 * TN + 8;
  static constexpr int STAGE = GEMM_KC * (SLD + SRD);          // doubles per ring stage
  static constexpr size_t smem_bytes = sizeof(double) * 2 * STAGE;
  static constexpr int WR = TM / 16;       // warp grid: WR row groups of 16 rows
  static constexpr int WC = 8 / WR;        //            WC column groups
  static constexpr int CPW = TN / WC;      // columns per warp
  static constexpr int MT = CPW / 8;       // MMA tiles along columns per warp
  static constexpr int NT = 4;             // MMA tiles along rows per warp (16 rows / 4)
  static constexpr int NACC = MT * NT;
};

SD_DEV void cp_async16_zfill(void* smem, const void* gmem, bool valid) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(sa), "l"(gmem), "r"(sz));
}
SD_DEV void cp_async_commit_() { asm volatile("cp.async.commit_group;\n" ::); }
SD_DEV void cp_async_wait_all_() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// Problem factory: Prob(mat) -> GemmProb.  The work list of a CTA is
//   tile t = blockIdx.x, blockIdx.x + gridDim.x, ...  of  tiles_i x tiles_j x nmat  (i fastest),
// each tile followed through its ceil(K / 32) chunks.
template <int TM, int TN, bool SUB, bool CONJL, bool CONJR, bool LKFAST, bool RKFAST, class ProbF>
SD_DEV void gemm_pipe_run(const ProbF& probf, int tiles_i, int tiles_j, int nmat, double* smem) {
  typedef PipeCfg<TM, TN> G;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int g = lane >> 2, tg = lane & 3;
  const int rowbase = (wid % G::WR) * 16, colbase = (wid / G::WR) * G::CPW;
  const int lsel = tg >> 1, qsel = tg & 1;
  const int bx = (g & 1) ^ qsel;                              // component of L the B fragment needs
  // sign of the B fragment: the embedding {re,-im; im,re} (conj: {re,im; -im,re}); SUB accumulates C - L*R
  const bool bneg = (CONJL ? (((g & 1) == 1) && (qsel == 0)) : (((g & 1) == 0) && (qsel == 1))) != SUB;
  const bool aneg = CONJR && (qsel == 1);

  struct Item { int t, chunk, nchunks, nch1, i0, j0; GemmProb p; bool valid; };
  const int total = tiles_i * tiles_j * nmat;
  auto load_tile = [&](Item& it) {                             // advance it.t to the next tile with work
    it.valid = false;
    while (it.t < total) {
      const int ti = it.t % tiles_i, rest = it.t / tiles_i;
      const int tj = rest % tiles_j, mat = rest / tiles_j;
      it.p = probf(mat);
      it.i0 = ti * TM; it.j0 = tj * TN;
      if (it.p.m > 0 && it.p.nc > 0 && it.p.K > 0 && it.i0 < it.p.m && it.j0 < it.p.nc) {
        it.nch1 = (it.p.K + GEMM_KC - 1) / GEMM_KC;
        it.nchunks = it.nch1 + ((it.p.L2 && it.p.K2 > 0) ? (it.p.K2 + GEMM_KC - 1) / GEMM_KC : 0); it.chunk = 0; it.valid = true;
        return;
      }
      it.t += gridDim.x;
    }
  };
  auto advance = [&](Item& it) {
    if (it.chunk + 1 < it.nchunks) { it.chunk += 1; return; }
    it.t += gridDim.x;
    load_tile(it);
  };
  constexpr int NLD = (TM * GEMM_KC) / GEMM_THREADS, NRD = (TN * GEMM_KC) / GEMM_THREADS;
  static_assert(NLD + NRD <= GEMM_KC / 2 && G::NACC <= GEMM_KC / 2, "one load slice per k-step");
  static_assert(GEMM_THREADS % (4 * TM) == 0 && GEMM_THREADS % (4 * TN) == 0 && (GEMM_THREADS / 4) % TM == 0 && (GEMM_THREADS / 4) % TN == 0,
                "a thread's slices differ only in k");
  // ---- operand loader.  Slice u of a thread copies element (i, l0 + u DL) of the L chunk (u < NLD) or (l0 + u DL, j) of the
  // R chunk: i / j and l0 are per-thread constants and DL = 256 / TM (256 / TN) is a compile-time step, so everything that
  // depends on the work item -- base pointers of the tile and chunk, bounds -- is computed ONCE per item (prep_ld, sliced
  // into the k-loop of the previous item) and a slice is a pointer plus u steps, a compare and the cp.async.  (Round 1
  // recomputed the 64-bit address arithmetic in every slice: 58 of the 76 instructions between two DMMA groups.)
  constexpr int DLL = GEMM_THREADS / TM, DLR = GEMM_THREADS / TN;
  const int iL = LKFAST ? (tid >> 2) % TM : tid % TM;
  const int l0L = LKFAST ? (tid & 3) + 4 * (tid / (4 * TM)) : tid / TM;
  const int jR = RKFAST ? (tid >> 2) % TN : tid % TN;
  const int l0R = RKFAST ? (tid & 3) + 4 * (tid / (4 * TN)) : tid / TN;
  const unsigned dstL0 = (unsigned)__cvta_generic_to_shared(smem + l0L * G::SLD + 2 * iL);                       // stage 0, slice 0
  const unsigned dstR0 = (unsigned)__cvta_generic_to_shared(smem + GEMM_KC * G::SLD + l0R * G::SRD + 2 * jR);
  struct Ld { const cplx* pL; const cplx* pR; long long sL, sR; int kL, kR; };     // kL / kR: k entries left from l0 (<= 0: none, also when the row / column is out of range)
  auto prep_ld = [&](const Item& it, Ld& d) {
    const bool second = it.chunk >= it.nch1;
    const int k0 = (second ? it.chunk - it.nch1 : it.chunk) * GEMM_KC;
    const int Kc = second ? it.p.K2 : it.p.K;
    const cplx* Lb = second ? it.p.L2 : it.p.L;
    const cplx* Rb = second ? it.p.R2 : it.p.R;
    d.kL = (it.i0 + iL < it.p.m) ? Kc - k0 - l0L : 0;
    d.kR = (it.j0 + jR < it.p.nc) ? Kc - k0 - l0R : 0;
    d.pL = Lb + (long long)(it.i0 + iL) * it.p.lsi + (long long)(k0 + l0L) * it.p.lsl;
    d.pR = Rb + (long long)(k0 + l0R) * it.p.rsl + (long long)(it.j0 + jR) * it.p.rsj;
    d.sL = DLL * it.p.lsl; d.sR = DLR * it.p.rsl;
  };
  auto issue_one = [&](const Ld& d, int stage, int u) {          // u is a compile-time constant at every call site
    if (u < NLD) {
      const bool ok = u * DLL < d.kL;
      const cplx* src = d.pL + (ok ? u * d.sL : 0);
      const unsigned dst = dstL0 + (unsigned)((stage * G::STAGE + u * DLL * G::SLD) * sizeof(double));
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(ok ? 16 : 0));
    } else {
      const int v = u - NLD;
      const bool ok = v * DLR < d.kR;
      const cplx* src = d.pR + (ok ? v * d.sR : 0);
      const unsigned dst = dstR0 + (unsigned)((stage * G::STAGE + v * DLR * G::SRD) * sizeof(double));
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(ok ? 16 : 0));
    }
  };
  // ---- the C tile of a thread: entry q = (mt, nt) sits at pC + nt 4 + mt 8 ldc
  struct Cd { cplx* pC; long long ldc8; int rrem, crem; };
  auto prep_c = [&](const Item& it, Cd& d) {
    const int i = it.i0 + rowbase + tg, j = it.j0 + colbase + g;
    d.pC = it.p.C + i + (long long)j * it.p.ldc;
    d.ldc8 = 8LL * it.p.ldc;
    d.rrem = it.p.m - i; d.crem = it.p.nc - j;
  };
  auto load_c_one = [&](const Cd& d, cplx (&dst)[G::NACC], int q) {
    const int mt = q / G::NT, nt = q - mt * G::NT;
    dst[q] = (nt * 4 < d.rrem && mt * 8 < d.crem) ? d.pC[nt * 4 + mt * d.ldc8] : mk(0.0, 0.0);
  };
  // the sign of a thread's B fragment is a per-thread constant: flip the sign bit on the integer pipe
  // (a DADD would queue behind the DMMAs on the FP64 pipe)
  const int bmask = bneg ? (int)0x80000000 : 0, amask = aneg ? (int)0x80000000 : 0;

  Item cur; cur.t = blockIdx.x; load_tile(cur);
  if (!cur.valid) return;
  cplx acc[G::NACC], cpre[G::NACC];
  Ld lnxt; Cd ccur, cnxt;
  prep_c(cur, ccur);
#pragma unroll
  for (int q = 0; q < G::NACC; ++q) cpre[q] = mk(0.0, 0.0);
  if (SUB) {
#pragma unroll
    for (int q = 0; q < G::NACC; ++q) load_c_one(ccur, cpre, q);
  }
  prep_ld(cur, lnxt);
#pragma unroll
  for (int u = 0; u < NLD + NRD; ++u) issue_one(lnxt, 0, u);
  cp_async_commit_();
  Item nxt = cur; advance(nxt);
  if (nxt.valid) { prep_ld(nxt, lnxt); prep_c(nxt, cnxt); } else { cnxt = ccur; }
  Item nn = nxt;
  Ld lnn = lnxt; Cd cnn = cnxt;
  int stage = 0;
  while (cur.valid) {
    cp_async_wait_all_();
    __syncthreads();                       // chunk of `cur` visible to all; everyone has left the other stage
    if (cur.chunk == 0) {
#pragma unroll
      for (int q = 0; q < G::NACC; ++q) acc[q] = SUB ? cpre[q] : mk(0.0, 0.0);
    }
    const bool pre_c = SUB && nxt.valid && nxt.chunk == 0;
    const double* sL = smem + stage * G::STAGE;
    const double* sR = sL + GEMM_KC * G::SLD;
    // the k-steps of this item, with the loads of the NEXT item (operand chunk by cp.async, C tile into registers)
    // and the work-list advance sliced in between: they issue in the shadow of the queued DMMAs
#pragma unroll
    for (int ks = 0; ks < GEMM_KC / 2; ++ks) {
      const double* rl = sR + (2 * ks + lsel) * G::SRD + qsel;
      const double* ll = sL + (2 * ks + lsel) * G::SLD + bx;
      double a[G::MT], b[G::NT];
#pragma unroll
      for (int mt = 0; mt < G::MT; ++mt) {
        const double v = rl[2 * (colbase + mt * 8 + g)];
        a[mt] = CONJR ? __hiloint2double(__double2hiint(v) ^ amask, __double2loint(v)) : v;
      }
#pragma unroll
      for (int nt = 0; nt < G::NT; ++nt) {
        const double v = ll[2 * (rowbase + nt * 4 + (g >> 1))];
        b[nt] = __hiloint2double(__double2hiint(v) ^ bmask, __double2loint(v));
      }
#pragma unroll
      for (int mt = 0; mt < G::MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < G::NT; ++nt) dmma884(acc[mt * G::NT + nt].re, acc[mt * G::NT + nt].im, a[mt], b[nt]);
      if (nxt.valid && ks < NLD + NRD) issue_one(lnxt, stage ^ 1, ks);
      if (pre_c && ks < G::NACC) load_c_one(cnxt, cpre, ks);
      if (ks == GEMM_KC / 2 - 3) { nn = nxt; if (nxt.valid) advance(nn); }
      if (ks == GEMM_KC / 2 - 2) { if (nn.valid) prep_ld(nn, lnn); }
      if (ks == GEMM_KC / 2 - 1) { if (nn.valid) prep_c(nn, cnn); }
    }
    cp_async_commit_();
    if (cur.chunk + 1 == cur.nchunks) {
#pragma unroll
      for (int mt = 0; mt < G::MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < G::NT; ++nt)
          if (nt * 4 < ccur.rrem && mt * 8 < ccur.crem) ccur.pC[nt * 4 + mt * ccur.ldc8] = acc[mt * G::NT + nt];
    }
    cur = nxt; ccur = cnxt;
    nxt = nn; lnxt = lnn; cnxt = cnn;
    stage ^= 1;
  }
}

}  // namespace stab
#endif
