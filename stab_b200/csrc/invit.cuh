// invit.cuh -- right eigenvectors of the Hessenberg matrix by shift-invert inverse iteration,
// register-resident variant (GPU only).  Stage (4)/(6) of the pipeline: replaces ZTREVC of the
// ZGEEV('N','V') the reference calls (temporal.f90:803, spatial.f90:1043) by the ZHSEIN/ZLAEIN
// idea -- one solve with (H - lambda I) per eigenvalue, started from a vector of size eps.
//
// One WARP per eigenvalue, 8 eigenvalues per CTA advancing block by block through the elimination
//   for k = m-1 .. 1:  combine Hessenberg column k-1 with the carried column (column operations
//                      from the bottom row up, with column interchanges -- see evec.cuh)
// * the carried column c and the right-hand side y live in REGISTERS: lane l owns rows
//   r = 32 s + l (s = 0..NS-1), so every update is two complex FMAs on registers;
// * the Hessenberg columns are staged ONCE per CTA through shared memory in double-buffered
//   8-column blocks by bulk copies on the TMA engine (cp.async.bulk + mbarrier transaction counts,
//   tma.cuh), shared by the 8 warps: global/L2 traffic is (n^2/2 * 16 B) per 8 eigenvalues;
// * the pivots travel by warp shuffles; the multipliers overwrite the dead entries of c.
// Two forms of the step loop: the per-step form (InvitSlot, round 1; validation switch evec_mode 3)
// and the panel / bulk form (InvitGroup, the default; described at its definition).  Orders above
// 640 use two warps per eigenvalue (k_invit2).
// The vectors come out in the Hessenberg basis (zero below the diagonal block's end kr);
// the back-transformation by Q is a tensor-core GEMM (k_gemm_pipe<BT_*>), then k_vec_finalize.
#pragma once
#include "common.cuh"
#include "tma.cuh"

#ifndef STAB_EMU
namespace stab {

constexpr int INVIT_WARPS = 8;
constexpr int INVIT_CB = 8;        // columns per staged block
constexpr int INVIT_PAD = 512;     // slack behind the staging buffers (the branch-free boundary update reads whole 32-row slots)

SD_DEV cplx shfl_c(cplx v, int src) { return mk(__shfl_sync(0xffffffffu, v.re, src), __shfl_sync(0xffffffffu, v.im, src)); }

// One elimination step k with the boundary slot index SB = k >> 5 known at compile time: rows < 32*SB
// (slots 0..SB-1, or 0..SB-2 when k is a multiple of 32) are updated branch-free on registers.
// pivot decision of step k: ak = H(k, k-1) against the carried column's entry in row k
SD_DEV void invit_pivot(cplx ak, cplx cdiag, cplx ydiag, double eps3, bool& sw, cplx& mq, cplx& yk) {
  sw = cabs1(ak) > cabs1(cdiag);
  cplx piv = sw ? ak : cdiag;
  if (is_zero(piv)) piv = mk(eps3, 0.0);
  // 1/piv = conj(piv)/|piv|^2: pivots are O(eps3)..O(||H||), their squares are far from the
  // overflow/underflow thresholds, so Smith's dependent divisions are not needed here
  const double rd = __drcp_rn(fma(piv.re, piv.re, piv.im * piv.im));
  const cplx inv = mk(piv.re * rd, -piv.im * rd);
  yk = ydiag * inv;
  mq = (sw ? cdiag : ak) * inv;
}

template <int NS, int SB>
SD_DEV void invit_apply(int k, int lane, const cplx* __restrict__ acol, cplx lm, bool sw, cplx mq, cplx yk, cplx (&c)[NS], cplx (&y)[NS],
                        unsigned& flags, cplx& cdiag, cplx& ydiag) {
  cplx cnext = mk(0.0, 0.0), ynext = mk(0.0, 0.0);
  const bool edge = (k & 31) == 0;                   // row k-1 lives in slot SB-1 (lane 31)
  // (1) boundary slot(s): diagonal shift, pivot capture, multiplier store
#pragma unroll
  for (int s = (SB > 0 ? SB - 1 : 0); s <= SB; ++s) {
    if (s == SB || edge) {
      const int r = 32 * s + lane;
      if (r < k) {
        cplx a = acol[r];
        if (r == k - 1) a -= lm;
        const cplx cr = c[s];
        const cplx u = sw ? a : cr;
        const cplx cn = sw ? (cr - mq * a) : (a - mq * cr);
        const cplx yn = y[s] - yk * u;
        c[s] = cn; y[s] = yn;
        if (r == k - 1) { cnext = cn; ynext = yn; }
      } else if (r == k) {
        c[s] = mq; y[s] = yk;
        if (sw) flags |= (1u << s);
      }
    }
  }
  cdiag = shfl_c(cnext, (k - 1) & 31);
  ydiag = shfl_c(ynext, (k - 1) & 31);
  // (2) the bulk: slots that hold only rows < k-1
  constexpr int NP = SB > 0 ? SB - 1 : 0;            // always plain
  if (sw) {
#pragma unroll
    for (int s = 0; s < NP; ++s) {
      const cplx a = acol[32 * s + lane];
      fms_acc(y[s], yk, a); fms_acc(c[s], mq, a);
    }
    if (SB > 0 && !edge) {
      const cplx a = acol[32 * (SB - 1) + lane];
      fms_acc(y[SB > 0 ? SB - 1 : 0], yk, a); fms_acc(c[SB > 0 ? SB - 1 : 0], mq, a);
    }
  } else {
#pragma unroll
    for (int s = 0; s < NP; ++s) {
      cplx a = acol[32 * s + lane];
      const cplx cr = c[s];
      fms_acc(y[s], yk, cr); fms_acc(a, mq, cr);
      c[s] = a;
    }
    if (SB > 0 && !edge) {
      cplx a = acol[32 * (SB - 1) + lane];
      const cplx cr = c[SB > 0 ? SB - 1 : 0];
      fms_acc(y[SB > 0 ? SB - 1 : 0], yk, cr); fms_acc(a, mq, cr);
      c[SB > 0 ? SB - 1 : 0] = a;
    }
  }
}

template <int NS, int SB>
SD_DEV void invit_step(int k, int lane, const cplx* __restrict__ acol, cplx lm, double eps3, cplx (&c)[NS], cplx (&y)[NS],
                       unsigned& flags, cplx& cdiag, cplx& ydiag) {
  bool sw; cplx mq, yk;
  invit_pivot(acol[k], cdiag, ydiag, eps3, sw, mq, yk);
  invit_apply<NS, SB>(k, lane, acol, lm, sw, mq, yk, c, y, flags, cdiag, ydiag);
}

// The 32 steps k = 32*SB+31 .. 32*SB (four staged 8-column blocks), then recurse to SB-1.
template <int NS, int SB>
struct InvitSlot {
  template <class Prefetch>
  SD_DEV static void run(int n, int m, bool live, int lane, const cplx* __restrict__ H, cplx* sH, cplx lm, double eps3,
                         cplx (&c)[NS], cplx (&y)[NS], unsigned& flags, cplx& cdiag, cplx& ydiag, int& buf, Prefetch& prefetch) {
    for (int bq = 3; bq >= 0; --bq) {
      const int B = 4 * SB + bq;
      if (8 * B > n - 1) continue;                       // block above the matrix (uniform)
      prefetch.wait(buf);                                // block B landed (transaction barrier of its buffer)
      __syncthreads();                                   // everyone left block B+1
      if (B > 0) prefetch(B - 1, buf ^ 1);
      const cplx* tile = sH + (size_t)buf * INVIT_CB * n;
      for (int q = INVIT_CB - 1; q >= 0; --q) {
        const int k = 8 * B + q;
        if (k > n - 1 || k < 1) continue;
        if (!live || k > m - 1) continue;
        invit_step<NS, SB>(k, lane, tile + (size_t)q * n, lm, eps3, c, y, flags, cdiag, ydiag);
      }
      buf ^= 1;
    }
    InvitSlot<NS, SB - 1>::run(n, m, live, lane, H, sH, lm, eps3, c, y, flags, cdiag, ydiag, buf, prefetch);
  }
};
template <int NS>
struct InvitSlot<NS, -1> {
  template <class Prefetch>
  SD_DEV static void run(int, int, bool, int, const cplx*, cplx*, cplx, double, cplx (&)[NS], cplx (&)[NS], unsigned&, cplx&, cplx&, int&, Prefetch&) {}
};


// ---------------------------------------------------------------------------------------------------------------
// Round 2, panel / bulk form of the same elimination (k_invit<NS, 1>, the default).  The per-step kernel above issues,
// for every step, the pivot chain (compare, reciprocal, two complex products, boundary update, shuffle: ~200 dependent
// cycles and a lane-divergent boundary branch) and then that step's bulk FMAs; with 8 warps per SM (the carried column
// and the right-hand side fill the register file) nothing hides the chain: FP64 pipe 24 % busy.  Here the 8 steps of a
// staged block are split:
//   panel : the 8 pivot decisions and the update of the BOUNDARY slot only (rows 32 SB .. k-1, branch-free: every lane
//           computes the update, selects keep the finished rows), recording (mq, yk) per step in shared memory and the
//           interchange bits in a register;
//   bulk  : the 8 recorded steps applied to the slots above the boundary -- one shared-memory load and 8 DFMAs per slot
//           and step, no dependence on the pivot chain, warp-uniform branch on the interchange bit.
// The pivot chain of a step depends only on the boundary slot, which the panel keeps current, so the bulk of a block may
// trail its panel; at a slot edge (k = 32 SB: row k-1 is lane 31 of slot SB-1) the trailing steps are first applied to
// slot SB-1 alone, then the edge step runs on that slot.  Every entry sees the same operations in the same order as in
// the per-step kernel.
struct InvitRec { cplx mq, yk; };

// 1/d for d > 0 in the normal range: the 2^-23 seed of MUFU.RCP64H and one cubically convergent correction x (1 + e + e^2),
// e = 1 - d x: relative error ~2^-69 before rounding -- three dependent FMAs instead of the five (and the range check with
// its slow-path call) of the correctly rounded __drcp_rn.  The reciprocal sits on the pivot chain of every step.
SD_DEV double rcp_fast(double d) {
  double x;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
  const double e = fma(-d, x, 1.0);
  const double e2 = fma(e, e, e);
  return fma(x, e2, x);
}

// pivot decision of the panel: as invit_pivot, with the dependent chain shortened -- the zero test comes from the CABS1
// values the comparison already has, and the numerators are multiplied by conj(piv) beside the reciprocal of |piv|^2
// instead of after it
SD_DEV void invit_pivot_pb(cplx ak, cplx cdiag, cplx ydiag, double eps3, bool& sw, cplx& mq, cplx& yk) {
  const double ca = cabs1(ak), cc = cabs1(cdiag);
  sw = ca > cc;
  const bool zero = !sw && cc == 0.0;                    // the chosen pivot is zero only if both candidates are
  cplx piv = sw ? ak : cdiag;
  if (zero) piv = mk(eps3, 0.0);
  const cplx num = sw ? cdiag : ak;
  const double rd = rcp_fast(fma(piv.re, piv.re, piv.im * piv.im));
  const cplx tm = mulc(num, piv), ty = mulc(ydiag, piv);
  mq = tm * rd;
  yk = ty * rd;
}

// one panel step k (k & 31 != 0) of slot group SB: pivot, boundary slot SB, next carried diagonal.
// Branch-free and with nothing but two FMA levels between the multiplier and the shuffle that feeds the next pivot: the
// operands of   c' = P - mq Q,  y' = Y - yk Q   are chosen per lane class BEFORE the pivot chain delivers mq --
//   rows below the pivot row : (P, Q) = interchanged ? (c, a) : (a, c),  Y = y      the elimination proper
//   the pivot row            : (P, Q, Y) = (0, -1, 0)                               c' = mq, y' = yk exactly: the multiplier is parked
//   finished rows            : (P, Q, Y) = (c, 0, y)                                unchanged
// and the interchange bit, known early in the chain, picks between the two variants.
template <int NS, int SB>
SD_DEV void invit_panel_step(int k, int lane, cplx ak, cplx a, cplx lm, double eps3, cplx (&c)[NS], cplx (&y)[NS], unsigned& flags,
                             cplx& cdiag, cplx& ydiag, bool& sw, cplx& mq, cplx& yk) {
  const int kk = k & 31;
  const bool below = lane < kk, at = lane == kk;
  if (lane == kk - 1) a -= lm;
  const cplx cr = c[SB], yr = y[SB];
  const cplx unit = mk(at ? -1.0 : 0.0, 0.0);
  const cplx qa = below ? a : unit, qc = below ? cr : unit;
  const cplx pa = at ? mk(0.0, 0.0) : cr, pc = below ? a : pa;
  const cplx y0 = at ? mk(0.0, 0.0) : yr;
  invit_pivot_pb(ak, cdiag, ydiag, eps3, sw, mq, yk);
  const cplx q = sw ? qa : qc, p = sw ? pa : pc;
  cplx cn = p, yn = y0;
  fms_acc(cn, mq, q);
  fms_acc(yn, yk, q);
  c[SB] = cn; y[SB] = yn;
  flags |= (at && sw) ? (1u << SB) : 0u;
  cdiag = shfl_c(cn, kk - 1);
  ydiag = shfl_c(yn, kk - 1);
}

// the edge step k = 32 SB (SB >= 1): all of slot SB-1 is above the pivot row, lane 0 of slot SB takes the multiplier
template <int NS, int SB>
SD_DEV void invit_panel_edge(int lane, cplx ak, cplx a, cplx lm, double eps3, cplx (&c)[NS], cplx (&y)[NS], unsigned& flags,
                             cplx& cdiag, cplx& ydiag, bool& sw, cplx& mq, cplx& yk) {
  constexpr int S1 = SB > 0 ? SB - 1 : 0;
  invit_pivot_pb(ak, cdiag, ydiag, eps3, sw, mq, yk);
  if (lane == 31) a -= lm;
  const cplx cr = c[S1], yr = y[S1];
  const cplx q = sw ? a : cr, p = sw ? cr : a;
  cplx cn = p, yn = yr;
  fms_acc(cn, mq, q);
  fms_acc(yn, yk, q);
  c[S1] = cn; y[S1] = yn;
  if (lane == 0) {
    c[SB] = mq; y[SB] = yk;
    if (sw) flags |= (1u << SB);
  }
  cdiag = shfl_c(cn, 31);
  ydiag = shfl_c(yn, 31);
}

// one recorded step applied to slots S0 .. S1-1 (all rows above the boundary)
template <int NS, int S0, int S1>
SD_DEV void invit_bulk_step(int lane, const cplx* __restrict__ acol, bool sw, cplx mq, cplx yk, cplx (&c)[NS], cplx (&y)[NS]) {
  if (sw) {
#pragma unroll
    for (int s = S0; s < S1; ++s) {
      const cplx a = acol[32 * s + lane];
      fms_acc(y[s], yk, a); fms_acc(c[s], mq, a);
    }
  } else {
#pragma unroll
    for (int s = S0; s < S1; ++s) {
      cplx a = acol[32 * s + lane];
      const cplx cr = c[s];
      fms_acc(y[s], yk, cr); fms_acc(a, mq, cr);
      c[s] = a;
    }
  }
}

template <int NS, int SB, int CBK = INVIT_CB>
struct InvitGroup {
  template <class Prefetch>
  SD_DEV static void run(int n, int m, bool live, int lane, cplx* sH, InvitRec* rec, cplx lm, double eps3,
                         cplx (&c)[NS], cplx (&y)[NS], unsigned& flags, cplx& cdiag, cplx& ydiag, int& buf, Prefetch& prefetch) {
    constexpr int S1 = SB > 0 ? SB - 1 : 0;
#pragma unroll 1
    for (int bq = 32 / CBK - 1; bq >= 0; --bq) {
      const int B = (32 / CBK) * SB + bq;
      if (CBK * B > n - 1) continue;                       // block above the matrix (uniform)
      prefetch.wait(buf);                                // block B landed (transaction barrier of its buffer)
      __syncthreads();                                   // everyone left block B+1
      if (B > 0) prefetch(B - 1, buf ^ 1);
      const cplx* tile = sH + (size_t)buf * CBK * n;
      const int k0 = CBK * B;
      const int qhi = live ? min(CBK - 1, m - 1 - k0) : -1;   // steps k0+qhi .. k0+qlo of this block are this eigenvalue's
      const bool edge = SB > 0 && bq == 0;               // q = 0 is the slot-edge step k = 32 SB
      const int qlo = (B == 0 || edge) ? 1 : 0;          // k = 0 is no step; the edge step runs apart
      unsigned swm = 0u;
      // the boundary slot is read whole; rows past the end of the matrix are clamped to the last row (their values are never
      // selected), so that the load cannot reach into the other buffer while its bulk copies are in flight
      const int rb = min(32 * SB + lane, n - 1);
      // ---- panel ----
      if (qhi >= qlo) {
        cplx ak = tile[(size_t)qhi * n + k0 + qhi], a = tile[(size_t)qhi * n + rb];
#pragma unroll 2
        for (int q = qhi; q >= qlo; --q) {
          const int k = k0 + q;
          const int qn = q > qlo ? q - 1 : q;            // next step's entries: loaded ahead of this step's pivot chain
          const cplx ak_n = tile[(size_t)qn * n + k0 + qn], a_n = tile[(size_t)qn * n + rb];
          bool sw; cplx mq, yk;
          invit_panel_step<NS, SB>(k, lane, ak, a, lm, eps3, c, y, flags, cdiag, ydiag, sw, mq, yk);
          if (sw) swm |= 1u << q;
          if (lane == 0) { rec[q].mq = mq; rec[q].yk = yk; }
          ak = ak_n; a = a_n;
        }
      }
      __syncwarp();
      // ---- bulk ----
      if (edge) {
        if (qhi >= 0) {
#pragma unroll 1
          for (int q = qhi; q >= 1; --q)                 // bring slot SB-1 up to date, then the edge step on it
            invit_bulk_step<NS, S1, SB>(lane, tile + (size_t)q * n, (swm >> q) & 1u, rec[q].mq, rec[q].yk, c, y);
          {
            bool sw; cplx mq, yk;
            invit_panel_edge<NS, SB>(lane, tile[k0], tile[32 * S1 + lane], lm, eps3, c, y, flags, cdiag, ydiag, sw, mq, yk);
            if (sw) swm |= 1u;
            if (lane == 0) { rec[0].mq = mq; rec[0].yk = yk; }
          }
          __syncwarp();
#pragma unroll 1
          for (int q = qhi; q >= 0; --q)
            invit_bulk_step<NS, 0, S1>(lane, tile + (size_t)q * n, (swm >> q) & 1u, rec[q].mq, rec[q].yk, c, y);
        }
      } else {
#pragma unroll 1
        for (int q = qhi; q >= qlo; --q)
          invit_bulk_step<NS, 0, SB>(lane, tile + (size_t)q * n, (swm >> q) & 1u, rec[q].mq, rec[q].yk, c, y);
      }
      __syncwarp();                                      // records read before the next block's panel rewrites them
      buf ^= 1;
    }
    InvitGroup<NS, SB - 1, CBK>::run(n, m, live, lane, sH, rec, lm, eps3, c, y, flags, cdiag, ydiag, buf, prefetch);
  }
};
template <int NS, int CBK>
struct InvitGroup<NS, -1, CBK> {
  template <class Prefetch>
  SD_DEV static void run(int, int, bool, int, cplx*, InvitRec*, cplx, double, cplx (&)[NS], cplx (&)[NS], unsigned&, cplx&, cplx&, int&, Prefetch&) {}
};


// x = T_{m-1} ... T_1 y, the first-order recurrence that undoes the column operations:
//   k = 1 .. m-1:  t = y_k - m_k prev;  interchanged ? (x_{k-1} = t, prev kept) : (x_{k-1} = prev, prev = t);   x_{m-1} = prev
// as a warp-wide scan instead of one lane walking all rows (7 % of the kernel): a row acts on `prev` as the affine map
// prev -> A prev + B with (A, B) = (1, 0) or (-m_k, y_k); lane l composes the maps of its chunk of R consecutive rows
// (R odd: the chunks start in 8 different 16-byte bank groups, so the strided shared-memory reads are conflict free), five
// shuffle steps give every lane the state entering its chunk, and a second pass over the chunk produces x.  A lane's first
// store hits the last row its left neighbour still reads, so that one store waits for the warp.  |m_k| <= sqrt 2 (partial
// pivoting), in practice the products decay; a product that overflows ends in a NaN vector, which the caller's growth test
// hands to the retry kernel.
SD_DEV void invit_recurrence_scan(int m, int lane, cplx* wy, const cplx* wc, const unsigned char* wf, cplx prev0) {
  const int R = ((m - 1 + 31) / 32) | 1;
  const int k0 = 1 + lane * R, k1 = min(k0 + R, m);
  cplx A = mk(1.0, 0.0), B = mk(0.0, 0.0);
#pragma unroll 4
  for (int k = k0; k < k1; ++k) {
    if (wf[k] == 0) {
      const cplx mm = wc[k];
      cplx nb = wy[k];
      fms_acc(nb, mm, B);
      B = nb;
      A = -(mm * A);
    }
  }
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const cplx Ae = mk(__shfl_up_sync(0xffffffffu, A.re, d), __shfl_up_sync(0xffffffffu, A.im, d));
    const cplx Be = mk(__shfl_up_sync(0xffffffffu, B.re, d), __shfl_up_sync(0xffffffffu, B.im, d));
    if (lane >= d) { fma_acc(B, A, Be); A = A * Ae; }
  }
  cplx Ax = mk(__shfl_up_sync(0xffffffffu, A.re, 1), __shfl_up_sync(0xffffffffu, A.im, 1));
  cplx Bx = mk(__shfl_up_sync(0xffffffffu, B.re, 1), __shfl_up_sync(0xffffffffu, B.im, 1));
  if (lane == 0) { Ax = mk(1.0, 0.0); Bx = mk(0.0, 0.0); }
  cplx prev = Bx;
  fma_acc(prev, Ax, prev0);                                 // the state entering this lane's chunk
  cplx first = mk(0.0, 0.0);
#pragma unroll 4
  for (int k = k0; k < k1; ++k) {
    cplx t = wy[k];
    fms_acc(t, wc[k], prev);
    const bool f = wf[k] != 0;
    const cplx fin = f ? t : prev;
    prev = f ? prev : t;
    if (k == k0) first = fin; else wy[k - 1] = fin;
  }
  __syncwarp();                                             // the left neighbour has read its last row
  if (k0 < k1) wy[k0 - 1] = first;
  if (k1 == m && (k0 < k1 || (m == 1 && lane == 0))) wy[m - 1] = prev;   // the lane holding the last row (m = 1: no row at all)
}

// grid: (ceil(n / (8*rounds)), batch), block 256.  smem: 2 * INVIT_CB * n complex + INVIT_WARPS * n bytes (+ INVIT_PAD).  n <= 32 NS.
// PB = 1: panel / bulk form (default), PB = 0: the per-step form (validation switch evec_mode 3).
template <int NS, int PB>
__global__ void __launch_bounds__(INVIT_WARPS * 32, 1)
k_invit(const cplx* __restrict__ Hh, size_t hstride, int n, const cplx* __restrict__ lam, const int* __restrict__ kr,
        const double* __restrict__ hnorm, cplx* __restrict__ Y, size_t ystride, int* __restrict__ bad, int rounds) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cplx* sH = reinterpret_cast<cplx*>(smem_raw);            // [2][INVIT_CB][n]
  const int p = blockIdx.y;
  const cplx* H = Hh + (size_t)p * hstride;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const double smlnum = SD_SAFMIN * ((double)n / SD_ULP);
  const double eps3 = fmax(SD_ULP * hnorm[p], smlnum);
  const double growto = 0.1 / sqrt((double)n);
  __shared__ uint64_t bars[2];                              // transaction barriers of the two staging buffers
  __shared__ InvitRec recs[PB ? INVIT_WARPS * INVIT_CB : 1];  // (mq, yk) of the current block's steps, per warp
  constexpr unsigned NISS = PB ? INVIT_WARPS : 1;          // arrivals per staged block (panel / bulk form: one issuing lane per warp)
  if (threadIdx.x == 0) { mbar_init(&bars[0], NISS); mbar_init(&bars[1], NISS); mbar_fence_init(); }
  __syncthreads();
  unsigned par = 0u;                                        // phase parity of the two barriers (bits 0, 1), tracked by every thread

  for (int rd = 0; rd < rounds; ++rd) {
    const int e = (blockIdx.x * rounds + rd) * INVIT_WARPS + wid;
    const bool live = e < n;
    const int m = live ? kr[(size_t)p * n + e] + 1 : 0;     // leading block order
    const cplx lm = live ? lam[(size_t)p * n + e] : mk(0.0, 0.0);
    cplx c[NS], y[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) { c[s] = mk(0.0, 0.0); y[s] = mk(0.0, 0.0); }
    unsigned flags = 0u;
    cplx cdiag = mk(0.0, 0.0), ydiag = mk(0.0, 0.0);

    // block B covers steps k = 8B+7 .. 8B, i.e. columns k-1 = 8B+6 .. 8B-1, rows 0..k of each;
    // tile column q <-> step k = 8B+q.  Every column is one contiguous run of (k+1) entries: ONE thread issues one bulk
    // copy (TMA engine, cp.async.bulk) per column with completion on the buffer's transaction barrier -- no per-thread
    // 16-byte cp.async and address arithmetic (that loop was 6 % of the kernel's instructions).
    struct Stager {
      const cplx* H; cplx* sH; uint64_t* bars; unsigned* parity; int n;
      __device__ void operator()(int B, int bufi) const {
        if (PB != 0) {                                       // lane 0 of warp q issues column q: no warp carries the whole loop
          if ((threadIdx.x & 31) != 0) return;
          const int q = threadIdx.x >> 5, k = 8 * B + q;
          const bool ok = k >= 1 && k <= n - 1;
          const unsigned bytes = ok ? (unsigned)(k + 1) * 16u : 0u;
          fence_async_smem();
          mbar_arrive_expect_tx(bars + bufi, bytes);
          if (ok) bulk_g2s(sH + ((size_t)bufi * INVIT_CB + q) * n, H + (size_t)(k - 1) * n, bytes, bars + bufi);
          return;
        }
        if (threadIdx.x != 0) return;
        cplx* dst = sH + (size_t)bufi * INVIT_CB * n;
        unsigned bytes = 0;
        for (int q = 0; q < INVIT_CB; ++q) {
          const int k = 8 * B + q;
          if (k >= 1 && k <= n - 1) bytes += (unsigned)(k + 1) * 16u;
        }
        fence_async_smem();                                  // the buffer may have been written by generic stores (parked vectors)
        mbar_arrive_expect_tx(bars + bufi, bytes);
        for (int q = 0; q < INVIT_CB; ++q) {
          const int k = 8 * B + q;
          if (k < 1 || k > n - 1) continue;
          bulk_g2s(dst + (size_t)q * n, H + (size_t)(k - 1) * n, (unsigned)(k + 1) * 16u, bars + bufi);
        }
      }
      __device__ void wait(int bufi) const { mbar_wait(bars + bufi, (*parity >> bufi) & 1u); *parity ^= 1u << bufi; }
    };
    Stager prefetch{H, sH, bars, &par, n};
    if (live) {                                             // start: carried column = column m-1 of H - lam I
      const cplx* hc = H + (size_t)(m - 1) * n;
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        const int r = 32 * s + lane;
        if (r < m) {
          cplx a = hc[r];
          if (r == m - 1) a -= lm;
          c[s] = a; y[s] = mk(eps3, 0.0);
        }
      }
      cdiag = hc[m - 1] - lm;
      ydiag = mk(eps3, 0.0);
    }
    __syncthreads();                                        // previous round finished with both buffers
    int buf = 0;
    prefetch((n - 1) >> 3, 0);
    if constexpr (PB != 0) InvitGroup<NS, NS - 1>::run(n, m, live, lane, sH, recs + wid * INVIT_CB, lm, eps3, c, y, flags, cdiag, ydiag, buf, prefetch);
    else InvitSlot<NS, NS - 1>::run(n, m, live, lane, H, sH, lm, eps3, c, y, flags, cdiag, ydiag, buf, prefetch);
    // ---- k = 0, then x = T_{m-1} ... T_1 y: a first-order recurrence in k.  The warp parks y, the multipliers and
    // the interchange flags in the (now idle) staging buffers and ONE lane runs the recurrence from shared memory
    // (~25 dependent cycles per row) instead of three warp shuffles per row ----
    __syncthreads();                                        // every warp has finished reading the staged columns
    int isbad = 0;
    if (live) {
      cplx* wy = sH + (size_t)wid * 2 * n;
      cplx* wc = wy + n;
      unsigned char* wf = reinterpret_cast<unsigned char*>(sH + (size_t)2 * INVIT_CB * n) + (size_t)wid * n;
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        const int r = 32 * s + lane;
        if (r < m) { wy[r] = y[s]; wc[r] = c[s]; wf[r] = (unsigned char)((flags >> s) & 1u); }
      }
      __syncwarp();
      if constexpr (PB != 0) {
        cplx piv = cdiag;
        if (is_zero(piv)) piv = mk(eps3, 0.0);
        invit_recurrence_scan(m, lane, wy, wc, wf, cdiv(ydiag, piv));
      } else
      if (lane == 0) {
        cplx piv = cdiag;
        if (is_zero(piv)) piv = mk(eps3, 0.0);
        cplx prev = cdiv(ydiag, piv);                       // current value of y[k-1]
        int k = 1;
        for (; k + 3 < m; k += 4) {                         // loads of four rows in flight, recurrence in order
          const cplx y0 = wy[k], y1 = wy[k + 1], y2 = wy[k + 2], y3 = wy[k + 3];
          const cplx m0 = wc[k], m1 = wc[k + 1], m2 = wc[k + 2], m3 = wc[k + 3];
          const bool f0 = wf[k] != 0, f1 = wf[k + 1] != 0, f2 = wf[k + 2] != 0, f3 = wf[k + 3] != 0;
          cplx t, fin;
          t = y0 - m0 * prev; fin = f0 ? t : prev; prev = f0 ? prev : t; wy[k - 1] = fin;
          t = y1 - m1 * prev; fin = f1 ? t : prev; prev = f1 ? prev : t; wy[k] = fin;
          t = y2 - m2 * prev; fin = f2 ? t : prev; prev = f2 ? prev : t; wy[k + 1] = fin;
          t = y3 - m3 * prev; fin = f3 ? t : prev; prev = f3 ? prev : t; wy[k + 2] = fin;
        }
        for (; k < m; ++k) {
          const cplx t = wy[k] - wc[k] * prev;
          const bool f = wf[k] != 0;
          const cplx fin = f ? t : prev;                    // final x[k-1]
          prev = f ? prev : t;
          wy[k - 1] = fin;
        }
        wy[m - 1] = prev;                                   // last row gets the carried value
      }
      __syncwarp();
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        const int r = 32 * s + lane;
        if (r < m) y[s] = wy[r];
      }
      double vn = 0.0;
#pragma unroll
      for (int s = 0; s < NS; ++s)
        if (32 * s + lane < m) vn += cabs1(y[s]);
      vn = warp_sum(vn);
      if (!(vn == vn) || vn > 1.0e300 || vn < growto) isbad = 1;
      cplx* out = Y + (size_t)p * ystride + (size_t)e * n;
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        const int r = 32 * s + lane;
        if (r < n) out[r] = (r < m) ? y[s] : mk(0.0, 0.0);
      }
      if (lane == 0) bad[(size_t)p * n + e] = isbad;
    }
  }
}


// ---------------------------------------------------------------------------------------------------------------
// Orders 32 NSH < n <= 64 NSH (the spatial companion problem at Ny = 128, n = 1280; Ny = 256 temporal): TWO warps per
// eigenvalue.  Warp F owns rows [0, R0), warp P rows [R0, 2 R0), R0 = 32 NSH; both keep their part of the carried column
// and of the right-hand side in registers, exactly as above.
//   steps k >= R0 (the pivot row is P's): P takes the pivot decision, publishes (sw, mq, yk) through shared memory, both
//                 warps meet at a 64-thread named barrier and apply the step to their rows -- F's rows are all above
//                 the boundary, so its update is the plain bulk form (at k = R0 it also shifts the diagonal entry of
//                 row R0-1 and takes over the carried diagonal);
//   steps k <  R0: P's rows are finished (they hold multipliers); F runs the single-warp step.
// The published values are double buffered by the parity of k, so P may run one step ahead of F.  Four eigenvalues per
// CTA; the Hessenberg columns are staged in 4-column blocks (2 x 4 x n x 16 B = 160 KB at n = 1280).
constexpr int INVIT2_CB = 4;
constexpr int INVIT2_PAIRS = 4;

SD_DEV void pair_barrier(int pi) { asm volatile("bar.sync %0, 64;" ::"r"(pi + 1) : "memory"); }

struct Invit2Pub { cplx mq, yk; int sw; int pad; };

template <int NSH>
SD_DEV void invit_follow(bool last, int lane, const cplx* __restrict__ acol, cplx lm, bool sw, cplx mq, cplx yk,
                         cplx (&c)[NSH], cplx (&y)[NSH], cplx& cdiag, cplx& ydiag) {
  if (sw) {
#pragma unroll
    for (int s = 0; s < NSH; ++s) {
      cplx a = acol[32 * s + lane];
      if (s == NSH - 1 && last && lane == 31) a -= lm;
      fms_acc(y[s], yk, a); fms_acc(c[s], mq, a);
    }
  } else {
#pragma unroll
    for (int s = 0; s < NSH; ++s) {
      cplx a = acol[32 * s + lane];
      if (s == NSH - 1 && last && lane == 31) a -= lm;
      const cplx cr = c[s];
      fms_acc(y[s], yk, cr); fms_acc(a, mq, cr);
      c[s] = a;
    }
  }
  if (last) { cdiag = shfl_c(c[NSH - 1], 31); ydiag = shfl_c(y[NSH - 1], 31); }
}

// the 32 steps k = 32 SBG + 31 .. 32 SBG (eight staged 4-column blocks), then recurse to SBG-1
template <int NSH, int SBG>
struct Invit2Slot {
  template <class Prefetch>
  SD_DEV static void run(int n, int m, bool live, int lane, int role, int pi, cplx* sH, Invit2Pub* pub, cplx lm, double eps3,
                         cplx (&c)[NSH], cplx (&y)[NSH], unsigned& flags, cplx& cdiag, cplx& ydiag, int& buf, Prefetch& prefetch) {
    constexpr int R0 = 32 * NSH;
    for (int bq = 32 / INVIT2_CB - 1; bq >= 0; --bq) {
      const int B = (32 / INVIT2_CB) * SBG + bq;
      if (INVIT2_CB * B > n - 1) continue;                 // block above the matrix (uniform)
      prefetch.wait(buf);                                  // block B landed (transaction barrier of its buffer)
      __syncthreads();                                     // everyone left block B+1
      if (B > 0) prefetch(B - 1, buf ^ 1);
      const cplx* tile = sH + (size_t)buf * INVIT2_CB * n;
      for (int q = INVIT2_CB - 1; q >= 0; --q) {
        const int k = INVIT2_CB * B + q;
        if (k > n - 1 || k < 1) continue;
        if (!live || k > m - 1) continue;                  // same for both warps of a pair
        const cplx* acol = tile + (size_t)q * n;
        if (SBG >= NSH) {                                  // pivot row in P's range
          Invit2Pub* pb = pub + 2 * pi + (k & 1);
          if (role == 1) {
            bool sw; cplx mq, yk;
            invit_pivot_pb(acol[k], cdiag, ydiag, eps3, sw, mq, yk);
            if (lane == 0) { pb->mq = mq; pb->yk = yk; pb->sw = sw ? 1 : 0; }
            pair_barrier(pi);
            invit_apply<NSH, (SBG >= NSH ? SBG - NSH : 0)>(k - R0, lane, acol + R0, lm, sw, mq, yk, c, y, flags, cdiag, ydiag);
          } else {
            pair_barrier(pi);
            const bool sw = pb->sw != 0;
            const cplx mq = pb->mq, yk = pb->yk;
            invit_follow<NSH>(k == R0, lane, acol, lm, sw, mq, yk, c, y, cdiag, ydiag);
          }
        } else if (role == 0) {
          invit_step<NSH, (SBG < NSH ? SBG : 0)>(k, lane, acol, lm, eps3, c, y, flags, cdiag, ydiag);
        }
      }
      buf ^= 1;
    }
    Invit2Slot<NSH, SBG - 1>::run(n, m, live, lane, role, pi, sH, pub, lm, eps3, c, y, flags, cdiag, ydiag, buf, prefetch);
  }
};
template <int NSH>
struct Invit2Slot<NSH, -1> {
  template <class Prefetch>
  SD_DEV static void run(int, int, bool, int, int, int, cplx*, Invit2Pub*, cplx, double, cplx (&)[NSH], cplx (&)[NSH], unsigned&, cplx&, cplx&,
                         int&, Prefetch&) {}
};

// ---------------------------------------------------------------------------------------------------------------
// Two warps per eigenvalue in the panel / bulk form (k_invit2<NSH, 1>, the default for 640 < n <= 1280).  While the pivot
// row is in P's range (steps k >= R0) P runs the PANEL of a staged 4-column block on its boundary slot exactly as the
// one-warp kernel does (local row numbers k - R0), publishes the four (mq, yk) records and the interchange bits through
// shared memory, and the pair meets ONCE per block (the per-step form met once per step); then P applies the recorded
// steps to its slots below the boundary and F to all of its slots.  At k = R0 the row above the pivot is F's last one:
// F takes that step -- P hands over the carried diagonal, F updates its last slot, returns the multiplier that belongs to
// P's row R0 (second meeting of that block) and applies the step to its other slots.  Below R0 P's rows are finished and
// F runs the one-warp panel / bulk groups; P only keeps the staging protocol going.
struct Invit2Mail { InvitRec rec[INVIT2_CB]; InvitRec edge; cplx cdiag, ydiag; unsigned swm; int esw; int pad[2]; };

// the step k = R0 on F's last slot: as invit_panel_edge, the multiplier goes back to P instead of into an own register
template <int NSH>
SD_DEV void invit_cross_edge(int lane, cplx ak, cplx a, cplx lm, double eps3, cplx (&c)[NSH], cplx (&y)[NSH], cplx& cdiag, cplx& ydiag,
                             bool& sw, cplx& mq, cplx& yk) {
  invit_pivot_pb(ak, cdiag, ydiag, eps3, sw, mq, yk);
  if (lane == 31) a -= lm;
  const cplx cr = c[NSH - 1], yr = y[NSH - 1];
  const cplx q = sw ? a : cr, p = sw ? cr : a;
  cplx cn = p, yn = yr;
  fms_acc(cn, mq, q);
  fms_acc(yn, yk, q);
  c[NSH - 1] = cn; y[NSH - 1] = yn;
  cdiag = shfl_c(cn, 31);
  ydiag = shfl_c(yn, 31);
}

template <int NSH, int SBL>                                  // SBL: P's local slot group, steps k = R0 + 32 SBL + 31 .. R0 + 32 SBL
struct Invit2Group {
  template <class Prefetch>
  SD_DEV static void run(int n, int m, bool live, int lane, int role, int pi, cplx* sH, Invit2Mail* mail, cplx lm, double eps3,
                         cplx (&c)[NSH], cplx (&y)[NSH], unsigned& flags, cplx& cdiag, cplx& ydiag, int& buf, Prefetch& prefetch) {
    constexpr int R0 = 32 * NSH;
    constexpr int S1 = SBL > 0 ? SBL - 1 : 0;
    constexpr int NBLK = 32 / INVIT2_CB;
#pragma unroll 1
    for (int bq = NBLK - 1; bq >= 0; --bq) {
      const int B = NBLK * (NSH + SBL) + bq;
      if (INVIT2_CB * B > n - 1) continue;                 // block above the matrix (uniform)
      prefetch.wait(buf);
      __syncthreads();                                     // everyone left block B+1
      if (B > 0) prefetch(B - 1, buf ^ 1);
      const cplx* tile = sH + (size_t)buf * INVIT2_CB * n;
      const int k0 = INVIT2_CB * B;
      const int qhi = live ? min(INVIT2_CB - 1, m - 1 - k0) : -1;   // the same for both warps of a pair
      const bool edgeP = SBL > 0 && bq == 0;               // q = 0 is P's own slot edge
      const bool edgeF = SBL == 0 && bq == 0;              // q = 0 is the step k = R0, taken by F
      const int qlo = (edgeP || edgeF) ? 1 : 0;
      if (role == 1) {
        unsigned swm = 0u;
        if (qhi >= qlo) {
          const int rb = min(R0 + 32 * SBL + lane, n - 1);  // as in InvitGroup: never past the matrix
          cplx ak = tile[(size_t)qhi * n + k0 + qhi], a = tile[(size_t)qhi * n + rb];
#pragma unroll 1
          for (int q = qhi; q >= qlo; --q) {
            const int qn = q > qlo ? q - 1 : q;
            const cplx ak_n = tile[(size_t)qn * n + k0 + qn], a_n = tile[(size_t)qn * n + rb];
            bool sw; cplx mq, yk;
            invit_panel_step<NSH, SBL>(k0 + q - R0, lane, ak, a, lm, eps3, c, y, flags, cdiag, ydiag, sw, mq, yk);
            if (sw) swm |= 1u << q;
            if (lane == 0) { mail->rec[q].mq = mq; mail->rec[q].yk = yk; }
            ak = ak_n; a = a_n;
          }
        }
        __syncwarp();
        if (edgeP && qhi >= 0) {                           // bring P's slot SBL-1 up to date, then the edge step on it
#pragma unroll 1
          for (int q = qhi; q >= 1; --q)
            invit_bulk_step<NSH, S1, SBL>(lane, tile + (size_t)q * n + R0, (swm >> q) & 1u, mail->rec[q].mq, mail->rec[q].yk, c, y);
          bool sw; cplx mq, yk;
          invit_panel_edge<NSH, SBL>(lane, tile[k0], tile[R0 + 32 * S1 + lane], lm, eps3, c, y, flags, cdiag, ydiag, sw, mq, yk);
          if (sw) swm |= 1u;
          if (lane == 0) { mail->rec[0].mq = mq; mail->rec[0].yk = yk; }
        }
        if (lane == 0) { mail->swm = swm; mail->cdiag = cdiag; mail->ydiag = ydiag; }
        pair_barrier(pi);                                  // (A) the block's records are published
        if (edgeP) {
#pragma unroll 1
          for (int q = qhi; q >= 0; --q)
            invit_bulk_step<NSH, 0, S1>(lane, tile + (size_t)q * n + R0, (swm >> q) & 1u, mail->rec[q].mq, mail->rec[q].yk, c, y);
        } else {
#pragma unroll 1
          for (int q = qhi; q >= qlo; --q)
            invit_bulk_step<NSH, 0, SBL>(lane, tile + (size_t)q * n + R0, (swm >> q) & 1u, mail->rec[q].mq, mail->rec[q].yk, c, y);
        }
        if (edgeF) {
          pair_barrier(pi);                                // (B) F has taken the step k = R0
          if (qhi >= 0 && lane == 0) {                     // its multiplier belongs to row R0: slot 0, lane 0
            c[0] = mail->edge.mq; y[0] = mail->edge.yk;
            if (mail->esw) flags |= 1u;
          }
        }
      } else {
        pair_barrier(pi);                                  // (A)
        const unsigned swm = mail->swm;
#pragma unroll 1
        for (int q = qhi; q >= (edgeF ? 1 : 0); --q)       // P's own slot edge is an ordinary step for F's rows
          invit_bulk_step<NSH, 0, NSH>(lane, tile + (size_t)q * n, (swm >> q) & 1u, mail->rec[q].mq, mail->rec[q].yk, c, y);
        if (edgeF) {
          bool sw = false; cplx mq = mk(0.0, 0.0), yk = mk(0.0, 0.0);
          if (qhi >= 0) {
            cdiag = mail->cdiag; ydiag = mail->ydiag;      // P's carried diagonal after its last step
            invit_cross_edge<NSH>(lane, tile[k0], tile[32 * (NSH - 1) + lane], lm, eps3, c, y, cdiag, ydiag, sw, mq, yk);
            if (lane == 0) { mail->edge.mq = mq; mail->edge.yk = yk; mail->esw = sw ? 1 : 0; }
          }
          pair_barrier(pi);                                // (B)
          if (qhi >= 0) invit_bulk_step<NSH, 0, NSH - 1>(lane, tile, sw, mq, yk, c, y);
        }
      }
      buf ^= 1;
    }
    Invit2Group<NSH, SBL - 1>::run(n, m, live, lane, role, pi, sH, mail, lm, eps3, c, y, flags, cdiag, ydiag, buf, prefetch);
  }
};
template <int NSH>
struct Invit2Group<NSH, -1> {
  template <class Prefetch>
  SD_DEV static void run(int, int, bool, int, int, int, cplx*, Invit2Mail*, cplx, double, cplx (&)[NSH], cplx (&)[NSH], unsigned&, cplx&, cplx&,
                         int&, Prefetch&) {}
};

// grid: (ceil(n / (4*rounds)), batch), block 256.  smem: 2 * INVIT2_CB * n complex + INVIT2_PAIRS * n bytes.  n <= 64 NSH.
// PB = 1: panel / bulk form (default), PB = 0: the per-step form (evec_mode 3).
template <int NSH, int PB>
__global__ void __launch_bounds__(INVIT2_PAIRS * 64, 1)
k_invit2(const cplx* __restrict__ Hh, size_t hstride, int n, const cplx* __restrict__ lam, const int* __restrict__ kr,
         const double* __restrict__ hnorm, cplx* __restrict__ Y, size_t ystride, int* __restrict__ bad, int rounds) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ Invit2Pub pub[2 * INVIT2_PAIRS];
  __shared__ Invit2Mail mails[PB ? INVIT2_PAIRS : 1];
  __shared__ InvitRec recs2[PB ? INVIT2_PAIRS * INVIT2_CB : 1];      // F's own records below R0 (one-warp groups)
  __shared__ double vnorm[2 * INVIT2_PAIRS];
  constexpr int R0 = 32 * NSH;
  cplx* sH = reinterpret_cast<cplx*>(smem_raw);            // [2][INVIT2_CB][n]
  const int p = blockIdx.y;
  const cplx* H = Hh + (size_t)p * hstride;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, pi = wid >> 1, role = wid & 1;
  const int rbase = role * R0;                             // first row this warp owns
  const double smlnum = SD_SAFMIN * ((double)n / SD_ULP);
  const double eps3 = fmax(SD_ULP * hnorm[p], smlnum);
  const double growto = 0.1 / sqrt((double)n);
  __shared__ uint64_t bars[2];                              // transaction barriers of the two staging buffers
  if (threadIdx.x == 0) { mbar_init(&bars[0], INVIT2_CB); mbar_init(&bars[1], INVIT2_CB); mbar_fence_init(); }   // one arrival per issuing lane
  __syncthreads();
  unsigned par = 0u;

  for (int rd = 0; rd < rounds; ++rd) {
    const int e = (blockIdx.x * rounds + rd) * INVIT2_PAIRS + pi;
    const bool live = e < n;
    const int m = live ? kr[(size_t)p * n + e] + 1 : 0;     // leading block order
    const cplx lm = live ? lam[(size_t)p * n + e] : mk(0.0, 0.0);
    cplx c[NSH], y[NSH];
#pragma unroll
    for (int s = 0; s < NSH; ++s) { c[s] = mk(0.0, 0.0); y[s] = mk(0.0, 0.0); }
    unsigned flags = 0u;
    cplx cdiag = mk(0.0, 0.0), ydiag = mk(0.0, 0.0);

    // block B covers steps k = 4B+3 .. 4B, i.e. columns k-1, rows 0..k of each; tile column q <-> step k = 4B+q; one
    // bulk copy (TMA engine) per column, issued by one thread, completion on the buffer's transaction barrier
    struct Stager {
      const cplx* H; cplx* sH; uint64_t* bars; unsigned* parity; int n;
      __device__ void operator()(int B, int bufi) const {
        if ((threadIdx.x & 31) != 0 || (threadIdx.x >> 5) >= INVIT2_CB) return;   // lane 0 of warp q issues column q
        const int q = threadIdx.x >> 5, k = INVIT2_CB * B + q;
        const bool ok = k >= 1 && k <= n - 1;
        const unsigned bytes = ok ? (unsigned)(k + 1) * 16u : 0u;
        fence_async_smem();
        mbar_arrive_expect_tx(bars + bufi, bytes);
        if (ok) bulk_g2s(sH + ((size_t)bufi * INVIT2_CB + q) * n, H + (size_t)(k - 1) * n, bytes, bars + bufi);
      }
      __device__ void wait(int bufi) const { mbar_wait(bars + bufi, (*parity >> bufi) & 1u); *parity ^= 1u << bufi; }
    };
    Stager prefetch{H, sH, bars, &par, n};
    if (live) {                                             // start: carried column = column m-1 of H - lam I
      const cplx* hc = H + (size_t)(m - 1) * n;
#pragma unroll
      for (int s = 0; s < NSH; ++s) {
        const int r = rbase + 32 * s + lane;
        if (r < m) {
          cplx a = hc[r];
          if (r == m - 1) a -= lm;
          c[s] = a; y[s] = mk(eps3, 0.0);
        }
      }
      cdiag = hc[m - 1] - lm;
      ydiag = mk(eps3, 0.0);
    }
    __syncthreads();                                        // previous round finished with both buffers
    int buf = 0;
    prefetch((n - 1) / INVIT2_CB, 0);
    if constexpr (PB != 0) {
      Invit2Group<NSH, NSH - 1>::run(n, m, live, lane, role, pi, sH, mails + pi, lm, eps3, c, y, flags, cdiag, ydiag, buf, prefetch);
      // below R0: F alone (P only waits, synchronises and issues its staging copy)
      InvitGroup<NSH, NSH - 1, INVIT2_CB>::run(n, m, live && role == 0, lane, sH, recs2 + pi * INVIT2_CB, lm, eps3, c, y, flags, cdiag, ydiag, buf,
                                               prefetch);
    } else {
      Invit2Slot<NSH, 2 * NSH - 1>::run(n, m, live, lane, role, pi, sH, pub, lm, eps3, c, y, flags, cdiag, ydiag, buf, prefetch);
    }
    // ---- x = T_{m-1} ... T_1 y: the first-order recurrence of the single-warp kernel, run by lane 0 of warp F over
    // the rows of both warps (parked in the idle staging buffers) ----
    __syncthreads();                                        // every warp has finished reading the staged columns
    if (live) {
      cplx* wy = sH + (size_t)pi * 2 * n;
      cplx* wc = wy + n;
      unsigned char* wf = reinterpret_cast<unsigned char*>(sH + (size_t)2 * INVIT2_CB * n) + (size_t)pi * n;
#pragma unroll
      for (int s = 0; s < NSH; ++s) {
        const int r = rbase + 32 * s + lane;
        if (r < m) { wy[r] = y[s]; wc[r] = c[s]; wf[r] = (unsigned char)((flags >> s) & 1u); }
      }
      pair_barrier(pi);
      if (role == 0) {                                      // warp F: the recurrence as a warp-wide scan (invit_recurrence_scan)
        cplx piv = cdiag;
        if (is_zero(piv)) piv = mk(eps3, 0.0);
        invit_recurrence_scan(m, lane, wy, wc, wf, cdiv(ydiag, piv));
      }
      pair_barrier(pi);
#pragma unroll
      for (int s = 0; s < NSH; ++s) {
        const int r = rbase + 32 * s + lane;
        if (r < m) y[s] = wy[r];
      }
      double vn = 0.0;
#pragma unroll
      for (int s = 0; s < NSH; ++s)
        if (rbase + 32 * s + lane < m) vn += cabs1(y[s]);
      vn = warp_sum(vn);
      if (lane == 0) vnorm[wid] = vn;
      pair_barrier(pi);
      vn = vnorm[2 * pi] + vnorm[2 * pi + 1];
      const int isbad = (!(vn == vn) || vn > 1.0e300 || vn < growto) ? 1 : 0;
      cplx* out = Y + (size_t)p * ystride + (size_t)e * n;
#pragma unroll
      for (int s = 0; s < NSH; ++s) {
        const int r = rbase + 32 * s + lane;
        if (r < n) out[r] = (r < m) ? y[s] : mk(0.0, 0.0);
      }
      if (role == 0 && lane == 0) bad[(size_t)p * n + e] = isbad;
    }
  }
}

}  // namespace stab
#endif
