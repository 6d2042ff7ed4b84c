// hqr.cuh -- eigenvalues of a complex upper Hessenberg matrix by shifted QR, one CTA per matrix.
// This is the ZHSEQR stage of the ZGEEV the reference calls (temporal.f90:803, spatial.f90:1043),
// eigenvalues only (the Schur form is not kept; eigenvectors come from inverse iteration, evec.cuh).
//
// Design (B200-first, not LAPACK's blocking):
//   * small-bulge MULTISHIFT sweeps: a chain of ns single-shift bulges (2-element reflectors,
//     spaced two rows apart) is chased down the active block;
//   * the chain is chased inside a W x W diagonal WINDOW held in shared memory (all bulges step
//     simultaneously: reflector generation / left application / right application are three
//     barrier-separated phases), the reflectors of the pass are recorded, and then streamed over
//     the off-window slabs straight from L2/HBM -- thread per column (left slab) or per row
//     (right slab), bulge-major order with a register carry, so every slab element is loaded and
//     stored once per bulge;
//   * active blocks that fit the window are finished entirely in shared memory by a single-shift
//     Wilkinson QR (ZLAHQR logic: conservative Ahues-Tisseur deflation test, exceptional shifts);
//     the same routine provides the ns shifts (eigenvalues of the trailing ns x ns block).
// Eigenvalues only => all updates are restricted to the active block [L, I].
#pragma once
#include "common.cuh"
#include "hessenberg.cuh"   // cta_zlarfg

namespace stab {

#ifndef LINE_ROT
#define LINE_ROT 4          // rotations (consecutive bulges) per line job of chase_tiles (2 and 8 measured: no better)
#endif

// A bulge step is a complex plane rotation G = [c s; -conj(s) c] (c real, ZLARTG's form) rather than ZLAHQR's
// 2-element Householder reflector: the same unitary similarity up to a phase, but 12 FMA-class operations per
// updated pair in FOUR independent chains of depth 3 (the reflector form is 14 operations in two chains of depth
// 7), and ONE rsqrt on the critical path of a bulge step instead of an rsqrt followed by a reciprocal.  The QR
// kernel is bound by the latency of dependent FP64 operations (ncu: a third of the stall samples are fixed-latency
// waits), so chain depth is what this buys: the slab pipeline in isolation runs 10 % faster (scratch/mb/slab_mb.cu).
struct Rot { cplx s; double c; double pad; };

// Slot of bulge b's reflector inside one time step of the record.  The slab pipeline splits the 16 stages over a lane
// pair (even lane: stages 0-7, odd lane: 8-15); with the natural order the two lanes of a pair read addresses 256 B apart
// -- the same shared-memory banks, a 2-way conflict on every reflector load (ncu: 47 % of the slab kernel's shared
// wavefronts were conflict replays).  Interleaving the halves puts them 32 B apart: different banks, one wavefront.
SD_HD int rec_slot(int ns, int b) { return ns == 16 ? (((b & 7) << 1) | (b >> 3)) : b; }

struct Grp { int tid, nt; bool warp; };
SD_DEV void grp_sync(const Grp& g) {
#ifndef STAB_EMU
  if (g.warp) __syncwarp(); else __syncthreads();
#else
  (void)g;
#endif
}

SD_DEV Rot rot_identity() { Rot r; r.s = mk(0.0, 0.0); r.c = 1.0; r.pad = 0.0; return r; }
SD_DEV bool rot_is_identity(const Rot& r) { return r.s.re == 0.0 && r.s.im == 0.0; }

// G [f; g] = [r; 0]: on return f := r, returns G.  d = 1/sqrt(|f|^2 (|f|^2+|g|^2)), c = |f|^2 d, s = f conj(g) d,
// r = f (|f|^2+|g|^2) d  (LAPACK 3.10 ZLARTG's unscaled branch).  The operands are scaled by a power of two only when
// their squares could leave the safe range; |f|^2 underflowing to zero is the f = 0 case (c = 0, a swap).
SD_DEV Rot lartg2(cplx& f, cplx g) {
  Rot r = rot_identity();
  if (is_zero(g)) return r;
  const double mx = fmax(cabs1(f), cabs1(g));
  double sc = 1.0;                                          // power-of-two scale
  if (mx > 1.0e70 || mx < 1.0e-70) { const int e = ilogb(mx); sc = ldexp(1.0, -e); }
  const cplx a = mk(f.re * sc, f.im * sc), b = mk(g.re * sc, g.im * sc);
  const double f2 = fma(a.re, a.re, a.im * a.im), g2 = fma(b.re, b.re, b.im * b.im);
  if (f2 == 0.0) {
#ifdef STAB_EMU
    const double ig = 1.0 / sqrt(g2);
#else
    const double ig = rsqrt(g2);
#endif
    r.c = 0.0; r.s = mk(b.re * ig, -b.im * ig);
    f = mk(g2 * ig / sc, 0.0);
    return r;
  }
  const double h2 = f2 + g2;
#ifdef STAB_EMU
  const double d = 1.0 / sqrt(f2 * h2);
#else
  const double d = rsqrt(f2 * h2);                         // one MUFU + Newton: no divide, no sqrt on the critical path
#endif
  r.c = f2 * d;
  const cplx fd = mk(a.re * d, a.im * d);
  r.s = mk(fma(fd.re, b.re, fd.im * b.im), fma(fd.im, b.re, -(fd.re * b.im)));   // f conj(g) d
  const double q = h2 * d;                                  // sqrt(h2) / |f|: scale free
  f = mk(f.re * q, f.im * q);
  return r;
}

SD_DEV void apply_left(const Rot& r, cplx& x1, cplx& x2) {     // [x1;x2] := G [x1;x2]   (12 FMA-class ops, 4 chains of depth 3)
  const cplx a = x1, b = x2;
  x1.re = fma(-r.s.im, b.im, fma(r.s.re, b.re, r.c * a.re));
  x1.im = fma(r.s.im, b.re, fma(r.s.re, b.im, r.c * a.im));
  x2.re = fma(-r.s.im, a.im, fma(-r.s.re, a.re, r.c * b.re));
  x2.im = fma(r.s.im, a.re, fma(-r.s.re, a.im, r.c * b.im));
}
SD_DEV void apply_right(const Rot& r, cplx& x1, cplx& x2) {    // [x1 x2] := [x1 x2] G^H
  const cplx a = x1, b = x2;
  x1.re = fma(r.s.im, b.im, fma(r.s.re, b.re, r.c * a.re));
  x1.im = fma(-r.s.im, b.re, fma(r.s.re, b.im, r.c * a.im));
  x2.re = fma(r.s.im, a.im, fma(-r.s.re, a.re, r.c * b.re));
  x2.im = fma(-r.s.im, a.re, fma(-r.s.re, a.im, r.c * b.im));
}

// ---------------------------------------------------------------------------------------------
// Chase of a bulge chain on a matrix S held in shared memory.  S(r,c) = S[r + c*lds] maps to
// global indices (g0+r, g0+c).  Bulge b at time t sits at k = L + t - 2b and is active while
// 0 <= t-2b <= I-1-L.  Left applications cover local columns <= chi, right applications local
// rows >= rlo.  `rec` (optional) receives the reflectors, (t-ta)*ns + b.
// ---------------------------------------------------------------------------------------------
SD_DEV void chase(const Grp& g, cplx* S, int lds, int g0, int rlo, int chi, int L, int I,
                  const cplx* shifts, int ns, int ta, int tb, Rot* rec, Rot* cur, long long* prof = nullptr) {
  const int smax = I - 1 - L;
#ifndef STAB_EMU
  long long pc0 = (prof && g.tid == 0) ? clock64() : 0;
#define CHASE_PROF(i) do { if (prof && g.tid == 0) { long long t1_ = clock64(); prof[i] += t1_ - pc0; pc0 = t1_; } } while (0)
#else
#define CHASE_PROF(i)
#endif
  for (int t = ta; t < tb; ++t) {
    for (int b = g.tid; b < ns; b += g.nt) {
      const int s = t - 2 * b;
      Rot r = rot_identity();
      if (s >= 0 && s <= smax) {
        const int kl = L + s - g0;
        if (s == 0) {
          cplx x1 = S[kl + kl * lds] - shifts[b];
          cplx x2 = S[kl + 1 + kl * lds];
          r = lartg2(x1, x2);
        } else {
          cplx x1 = S[kl + (kl - 1) * lds];
          cplx x2 = S[kl + 1 + (kl - 1) * lds];
          r = lartg2(x1, x2);
          S[kl + (kl - 1) * lds] = x1;
          S[kl + 1 + (kl - 1) * lds] = mk(0.0, 0.0);
        }
      }
      cur[b] = r;
      if (rec) rec[(t - ta) * ns + rec_slot(ns, b)] = r;
    }
    grp_sync(g);
    CHASE_PROF(10);
    // left / right applications: one thread per (column | row, bulge group); the bulges of one
    // time step touch disjoint row / column pairs, so a thread's updates are independent: load all
    // operands, then all the arithmetic, then all the stores (ILP instead of one latency chain per bulge)
    const int ncol = chi + 1;
    constexpr int NPT = 8, NH = 4;                          // bulges per thread, processed in halves (register budget: 128)
    const int ngrp = (ns + NPT - 1) / NPT;
    if (ncol * ngrp <= g.nt) {
      const int col = g.tid % ncol, grp = g.tid / ncol;     // also the row index in the right phase
      const bool mine = g.tid < ncol * ngrp;
#pragma unroll
      for (int h = 0; h < NPT / NH; ++h) {
        cplx x1[NH], x2[NH]; Rot rf[NH]; int kk[NH];
#pragma unroll
        for (int u = 0; u < NH; ++u) {
          const int b = grp + (h * NH + u) * ngrp, s = t - 2 * b;
          kk[u] = -1;
          rf[u] = rot_identity(); x1[u] = mk(0.0, 0.0); x2[u] = mk(0.0, 0.0);
          if (mine && b < ns && s >= 0 && s <= smax) {
            const int kl = L + s - g0;
            if (col >= kl) { kk[u] = kl; rf[u] = cur[b]; x1[u] = S[kl + col * lds]; x2[u] = S[kl + 1 + col * lds]; }
          }
        }
#pragma unroll
        for (int u = 0; u < NH; ++u) apply_left(rf[u], x1[u], x2[u]);
#pragma unroll
        for (int u = 0; u < NH; ++u)
          if (kk[u] >= 0) { S[kk[u] + col * lds] = x1[u]; S[kk[u] + 1 + col * lds] = x2[u]; }
      }
      grp_sync(g);
      CHASE_PROF(11);
      const int row = col;
#pragma unroll
      for (int h = 0; h < NPT / NH; ++h) {
        cplx x1[NH], x2[NH]; Rot rf[NH]; int kk[NH];
#pragma unroll
        for (int u = 0; u < NH; ++u) {
          const int b = grp + (h * NH + u) * ngrp, s = t - 2 * b;
          kk[u] = -1;
          rf[u] = rot_identity(); x1[u] = mk(0.0, 0.0); x2[u] = mk(0.0, 0.0);
          if (mine && row >= rlo && b < ns && s >= 0 && s <= smax) {
            const int kl = L + s - g0;
            int rmax = kl + 2; if (rmax > I - g0) rmax = I - g0;
            if (row <= rmax) { kk[u] = kl; rf[u] = cur[b]; x1[u] = S[row + kl * lds]; x2[u] = S[row + (kl + 1) * lds]; }
          }
        }
#pragma unroll
        for (int u = 0; u < NH; ++u) apply_right(rf[u], x1[u], x2[u]);
#pragma unroll
        for (int u = 0; u < NH; ++u)
          if (kk[u] >= 0) { S[row + kk[u] * lds] = x1[u]; S[row + (kk[u] + 1) * lds] = x2[u]; }
      }
      grp_sync(g);
    } else {
      // generic path (small thread groups / many bulges per thread)
      int ng = g.nt / ncol; if (ng < 1) ng = 1; if (ng > ns) ng = ns;
      for (int idx = g.tid; idx < ncol * ng; idx += g.nt) {
        const int col = idx % ncol, grp = idx / ncol;
        for (int b = grp; b < ns; b += ng) {
          const int s = t - 2 * b;
          if (s < 0 || s > smax) continue;
          const int kl = L + s - g0;
          if (col < kl) continue;
          const Rot r = cur[b];
          cplx x1 = S[kl + col * lds], x2 = S[kl + 1 + col * lds];
          apply_left(r, x1, x2);
          S[kl + col * lds] = x1; S[kl + 1 + col * lds] = x2;
        }
      }
      grp_sync(g);
      CHASE_PROF(11);
      for (int idx = g.tid; idx < ncol * ng; idx += g.nt) {
        const int row = idx % ncol, grp = idx / ncol;
        if (row < rlo) continue;
        for (int b = grp; b < ns; b += ng) {
          const int s = t - 2 * b;
          if (s < 0 || s > smax) continue;
          const int kl = L + s - g0;
          int rmax = kl + 2; if (rmax > I - g0) rmax = I - g0;
          if (row > rmax) continue;
          const Rot r = cur[b];
          cplx x1 = S[row + kl * lds], x2 = S[row + (kl + 1) * lds];
          apply_right(r, x1, x2);
          S[row + kl * lds] = x1; S[row + (kl + 1) * lds] = x2;
        }
      }
      grp_sync(g);
    }
    CHASE_PROF(12);
  }
}

// ---------------------------------------------------------------------------------------------
// Chase of a multishift chain inside the shared-memory window, ONE barrier per time step.
// The work of a time step is partitioned into jobs that touch disjoint entries, so the left and
// right applications need no barrier between them (a left and a right application commute; the
// per-entry order left-then-right of `chase` is kept):
//   * bulge job b (one thread per bulge): the 3x2 block rows k..k+2, cols k..k+1 of bulge b
//     (left on both columns, right on the three rows) and -- from the two entries it just
//     produced in registers -- the reflector of the NEXT step, written to the other half of the
//     double-buffered `cur`; bulges entering at the next step are generated by the same thread;
//   * tile job (b, b'), b' ahead of b: rows (k_b, k_b+1) x cols (k_b', k_b'+1): left with
//     reflector b on both columns, right with reflector b' on both rows;
//   * line jobs: columns right of the chain (left applications only) and rows above it (right
//     applications only), four consecutive bulges (8 contiguous entries) per job.
// S is the wsz x wsz window with global origin g0 (g0 <= L).  Everything inside the window is updated: columns right
// of the active block and rows above it too (they exist only when the caller wants the Schur form of the window, not
// just its eigenvalues), and the right applications extend upwards to local row rtop <= 0 -- the rows rtop..-1 of the
// same array hold a matrix that accumulates the rotations (the stacked [Z; T] layout of the deflation window).
// cur holds 2*ns reflectors.  rec receives reflector (t, b) at (t-ta)*ns + b.
// ---------------------------------------------------------------------------------------------
SD_DEV Rot zero_refl() { Rot r = rot_identity(); return r; }

template <int NSC>
SD_DEV void chase_tiles(const Grp& g, cplx* S, int lds, int g0, int wsz, int L, int I,
                        const cplx* shifts, int ns_rt, int ta, int tb, Rot* rec, Rot* cur, int rtop = 0) {
  const int ns = NSC ? NSC : ns_rt;                          // compile-time shift count when known (index arithmetic by shifts)
  const int smax = I - 1 - L;
  const int kb0 = L - g0;                                   // local position of a bulge at s = 0
  const int ilast = I - g0;                                 // last local row / column of the active block
  // prologue: reflectors of step ta (as `chase` generates them at the start of a step)
  {
    Rot* cb = cur + (ta & 1) * ns;
    for (int b = g.tid; b < ns; b += g.nt) {
      const int s = ta - 2 * b;
      Rot r = zero_refl();
      if (s >= 0 && s <= smax) {
        const int kl = kb0 + s;
        if (s == 0) {
          cplx x1 = S[kl + kl * lds] - shifts[b];
          cplx x2 = S[kl + 1 + kl * lds];
          r = lartg2(x1, x2);
        } else {
          cplx x1 = S[kl + (kl - 1) * lds];
          cplx x2 = S[kl + 1 + (kl - 1) * lds];
          r = lartg2(x1, x2);
          S[kl + (kl - 1) * lds] = x1;
          S[kl + 1 + (kl - 1) * lds] = mk(0.0, 0.0);
        }
      }
      cb[b] = r;
      rec[rec_slot(ns, b)] = r;
    }
  }
  grp_sync(g);
  const int nbt = (g.nt >= 64) ? 32 : 0;                    // threads reserved for the bulge jobs (first warp)
  const int ntile = ns * (ns - 1) / 2;
  const float inv_ns = 1.0f / (float)ns;                    // run-time shift count: no integer division per tile job
  int jj0 = (g.tid - nbt) - (ntile % (g.nt - nbt)); if (jj0 < 0) jj0 += g.nt - nbt;   // first line job of this thread (loop invariant)
#ifdef STAB_CHASE_MB
  long long mb_t0_ = clock64();
#endif
#ifdef STAB_TL
#define TL(w, k) do { if (threadIdx.x == (w) * 32 && g_tl_n[w] < 4000) { g_tl[w][g_tl_n[w]++] = (unsigned)clock() | 0u; g_tl_k[w][g_tl_n[w] - 1] = (k); } } while (0)
#else
#define TL(w, k)
#endif
  for (int t = ta; t < tb; ++t) {
    TL(0, 0); TL(1, 0);
    const Rot* cb = cur + (t & 1) * ns;
    Rot* nb = cur + ((t + 1) & 1) * ns;
    const bool more = (t + 1 < tb);
    // active bulges at time t: blo..bhi (s = t - 2b in [0, smax])
    int bhi = t >> 1; if (bhi > ns - 1) bhi = ns - 1;
    int blo = (t - smax + 1) >> 1; if (blo < 0) blo = 0;     // ceil((t - smax)/2) for t - smax >= 0
    // ---- bulge jobs ------------------------------------------------------------------------
#ifndef STAB_MB_NOBULGE
    if (g.tid < (nbt ? ns : g.nt)) {
      for (int b = g.tid; b < ns; b += (nbt ? ns : g.nt)) {
        const int s = t - 2 * b;
        Rot rn = zero_refl();
        if (s >= 0 && s <= smax) {
          const int k = kb0 + s;
          const bool has3 = (k + 2 <= ilast);
          cplx a00 = S[k + k * lds], a10 = S[k + 1 + k * lds];
          cplx a01 = S[k + (k + 1) * lds], a11 = S[k + 1 + (k + 1) * lds];
          cplx a20 = mk(0.0, 0.0), a21 = mk(0.0, 0.0);
          if (has3) { a20 = S[k + 2 + k * lds]; a21 = S[k + 2 + (k + 1) * lds]; }
          const Rot r = cb[b];
          apply_left(r, a00, a10); apply_left(r, a01, a11);
          apply_right(r, a00, a01); apply_right(r, a10, a11);
          if (has3) apply_right(r, a20, a21);
          if (more && s + 1 <= smax) {                       // reflector of step t+1 from column k, rows k+1, k+2
            rn = lartg2(a10, a20);
            a20 = mk(0.0, 0.0);
          }
          S[k + k * lds] = a00; S[k + 1 + k * lds] = a10;
          S[k + (k + 1) * lds] = a01; S[k + 1 + (k + 1) * lds] = a11;
          if (has3) { S[k + 2 + k * lds] = a20; S[k + 2 + (k + 1) * lds] = a21; }
        } else if (s == -1 && more && smax >= 0) {           // bulge b enters at step t+1
          cplx x1 = S[kb0 + kb0 * lds] - shifts[b];
          cplx x2 = S[kb0 + 1 + kb0 * lds];
          rn = lartg2(x1, x2);
        }
        nb[b] = rn;
        if (more) rec[(t + 1 - ta) * ns + rec_slot(ns, b)] = rn;
      }
    }
#endif
    TL(0, 1);
    // ---- tile and line jobs -----------------------------------------------------------------
    if (blo <= bhi && (nbt == 0 || g.tid >= nbt)) {
      const int klo = kb0 + t - 2 * blo;                     // position of the leading active bulge
      const int khi = kb0 + t - 2 * bhi;                     // position of the trailing active bulge
      const int c0 = klo + 2;                                // first left-only column
      int nL = wsz - c0; if (nL < 0) nL = 0;
      const int nR = khi - rtop;                             // right-only rows rtop .. khi-1
      constexpr int LR = (NSC == 1) ? 1 : ((NSC == 2) ? 2 : LINE_ROT);   // rotations per line job: no idle slots in the short chains of the small QR
      const int nq = (ns + LR - 1) / LR;
      const int w0 = g.tid - nbt, wn = g.nt - nbt;
#ifndef STAB_MB_NOTILE
      for (int j = w0; j < ntile; j += wn) {
        const int p = NSC ? j / ns : (int)(((float)j + 0.5f) * inv_ns), i = j - p * ns;   // j < 2^10: exact
        int b, bp;
        if (i <= p) { b = p + 1; bp = i; } else { b = ns - 1 - p; bp = i - p - 1; }
        if (bp < blo || b > bhi) continue;
        const int kr = kb0 + t - 2 * b, kc = kb0 + t - 2 * bp;
        cplx* q0 = S + kr + kc * lds;
        cplx a00 = q0[0], a10 = q0[1], a01 = q0[lds], a11 = q0[lds + 1];
        const Rot rl = cb[b], rr = cb[bp];
        apply_left(rl, a00, a10); apply_left(rl, a01, a11);
        apply_right(rr, a00, a01); apply_right(rr, a10, a11);
        q0[0] = a00; q0[1] = a10; q0[lds] = a01; q0[lds + 1] = a11;
      }
#endif
      // Line jobs continue the job numbering after the tiles, so the threads without a tile take the first lines.  The
      // column jobs (left applications) of every quarter come first, then the row jobs: a warp holds one kind except at
      // the single boundary, and since [x1 x2] G^H is G(c, conj s) applied to [x1; x2], both kinds run the SAME code with
      // the sign of Im s and the element stride selected per job -- no divergent branch.  All loads of a job are issued
      // before its arithmetic (four independent rotations), all stores after it.
      const int nCol = nL * nq, nlj = nCol + nR * nq;
#ifdef STAB_MB_NOLINE
      if (false)
#endif
      for (int u = jj0; u < nlj; u += wn) {
        const bool iscol = u < nCol;
        int v = iscol ? u : u - nCol;
        const int len = iscol ? nL : nR;
        int q = 0;
        while (v >= len) { v -= len; ++q; }                  // q = v / len without a division (nq is small); consecutive threads: consecutive lines
        cplx* base = iscol ? S + (c0 + v) * lds : S + (v + rtop);
        const int str = iscol ? 1 : lds;
        const double sgn = iscol ? 1.0 : -1.0;
        cplx x1[LR], x2[LR]; Rot r[LR]; bool on[LR];
#pragma unroll
        for (int e = 0; e < LR; ++e) {
          const int b = LR * q + e;
          on[e] = b >= blo && b <= bhi;
          const int k = kb0 + t - 2 * (on[e] ? b : blo);
          x1[e] = mk(0.0, 0.0); x2[e] = mk(0.0, 0.0);
          if (on[e]) { x1[e] = base[k * str]; x2[e] = base[(k + 1) * str]; }   // predicated loads: an idle slot must not read entries another job is writing
          r[e] = cb[on[e] ? b : blo];
          r[e].s.im *= sgn;
        }
#pragma unroll
        for (int e = 0; e < LR; ++e) apply_left(r[e], x1[e], x2[e]);
#pragma unroll
        for (int e = 0; e < LR; ++e)
          if (on[e]) { const int k = kb0 + t - 2 * (LR * q + e); base[k * str] = x1[e]; base[(k + 1) * str] = x2[e]; }
      }
    }
#ifdef STAB_CHASE_MB
    if ((g.tid & 31) == 0) { long long now_ = clock64(); g_chase_busy[g.tid >> 5] += now_ - mb_t0_; }
#endif
    TL(1, 1);
    grp_sync(g);
#ifdef STAB_CHASE_MB
    mb_t0_ = clock64();
#endif
  }
}

// ZLAHQR's small-subdiagonal test at (k, k-1) on a matrix accessed through `at(r,c)`;
// lo/hi bound the neighbours consulted when both diagonal entries vanish.
template <class At>
SD_DEV bool negligible_subdiag(const At& at, int k, int lo, int hi, double smlnum) {
  cplx hkk1 = at(k, k - 1);
  double a1 = cabs1(hkk1);
  if (a1 <= smlnum) return true;
  cplx hkk = at(k, k), hk1k1 = at(k - 1, k - 1);
  double tst = cabs1(hk1k1) + cabs1(hkk);
  if (tst == 0.0) {
    if (k - 2 >= lo) tst += cabs1(at(k - 1, k - 2));
    if (k + 1 <= hi) tst += cabs1(at(k + 1, k));
  }
  if (a1 <= SD_ULP * tst) {
    double b1 = cabs1(at(k - 1, k));
    double ab = fmax(a1, b1), ba = fmin(a1, b1);
    double d1 = cabs1(hkk), d2 = cabs1(hk1k1 - hkk);
    double aa = fmax(d1, d2), bb = fmin(d1, d2);
    double s = aa + ab;
    if (ba * (ab / s) <= fmax(smlnum, SD_ULP * (bb * (aa / s)))) return true;
  }
  return false;
}

struct SmemAt { const cplx* S; int lds; SD_DEV cplx operator()(int r, int c) const { return S[r + c * lds]; } };
struct GlobAt { const cplx* H; int ldh; SD_DEV cplx operator()(int r, int c) const { return H[r + (size_t)c * ldh]; } };

// Wilkinson shift of ZLAHQR from the trailing 2x2 of the active block ending at I
SD_DEV cplx wilkinson_shift(cplx h11, cplx h12, cplx h21, cplx h22) {
  cplx t = h22;
  cplx u = csqrt_(h12) * csqrt_(h21);
  double s = cabs1(u);
  if (s != 0.0) {
    cplx x = 0.5 * (h11 - t);
    double sx = cabs1(x);
    s = fmax(s, sx);
    cplx xs = mk(x.re / s, x.im / s), us = mk(u.re / s, u.im / s);
    cplx y = s * csqrt_(xs * xs + us * us);
    if (sx > 0.0) {
      if ((x.re / sx) * y.re + (x.im / sx) * y.im < 0.0) y = -y;
    }
    t = t - u * cdiv(u, x + y);
  }
  return t;
}

// Both eigenvalues of [h11 h12; h21 h22], l1 the one closer to h22 (Wilkinson's choice): mid +- sqrt(dlt^2 + h12 h21) with
// ONE complex square root built on rsqrt -- no hypot, no division (each a ~250-cycle dependent sequence on the one
// thread every other thread of the group is waiting for; the LAPACK-style wilkinson_shift above costs ~5000 cycles).
// The cancellation of the direct formula perturbs a shift by O(eps |h11 - h22|), which the iteration does not notice.
SD_DEV void eig2x2_fast(cplx h11, cplx h12, cplx h21, cplx h22, cplx& l1, cplx& l2) {
  const cplx mid = mk(0.5 * (h11.re + h22.re), 0.5 * (h11.im + h22.im));
  const cplx dlt = mk(0.5 * (h11.re - h22.re), 0.5 * (h11.im - h22.im));
  const cplx disc = dlt * dlt + h12 * h21;
  const double a2 = fma(disc.re, disc.re, disc.im * disc.im);
  cplx r;
  if (!(a2 > 1.0e-280 && a2 < 1.0e280)) {
    r = csqrt_(disc);                                       // out of the safe range of the squares (or zero / NaN)
  } else {
#ifdef STAB_EMU
    const double m = sqrt(a2);
    const double t = 0.5 * (m + fabs(disc.re)), it = 1.0 / sqrt(t);
#else
    const double m = a2 * rsqrt(a2);
    const double t = 0.5 * (m + fabs(disc.re)), it = rsqrt(t);
#endif
    const double sr = t * it, si = 0.5 * disc.im * it;
    r = (disc.re >= 0.0) ? mk(sr, si) : mk(fabs(si), disc.im >= 0.0 ? sr : -sr);
  }
  // l - h22 = dlt +- r: take the smaller one
  const double dot = dlt.re * r.re + dlt.im * r.im;
  if (dot > 0.0) r = -r;
  l1 = mid + r; l2 = mid - r;
}

struct SmallCtl { int L; int pad; cplx shift[2]; };

// Early termination of the Schur factorisation of a deflation window (smem_hqr<true>): the eigenvalues converge from
// the bottom of the window upwards, and column j of Z is final once position j has converged, so the deflation test of
// aggressive early deflation -- |s| |Z(0,j)| <= max(smlnum, ulp |T(j,j)|) -- is taken the moment position j converges.
// Without reordering the deflation ends at the first failure; the factorisation then goes on only for the `want`
// eigenvalues the next sweep uses as shifts, and not at all when nd >= nd_skip (another deflation step follows).
struct AedStop { double s1, smlnum; int nd_skip, want; };
#ifdef STAB_EMU_COUNT
static long long g_emu_cnt[8];
#endif

SD_DEV void grp_amax(int* p, int v) {              // shared-memory integer max (a plain max under emulation)
#ifdef STAB_EMU
  if (v > *p) *p = v;
#else
  atomicMax(p, v);
#endif
}

// All eigenvalues of the m x m upper Hessenberg matrix S (shared memory) by shifted QR.  A QR iteration chases a chain
// of TWO bulges whose shifts are the two eigenvalues of the trailing 2 x 2 block (one bulge with ZLAHQR's Wilkinson
// shift on blocks of order < 4, and for the exceptional shifts): half the time steps of single-shift sweeps, each of
// them with one barrier (chase_tiles).  The deflation scan is done by the whole group, one subdiagonal per thread.
// WANTT: S is taken to its Schur form -- the rotations act on the whole m x m matrix, not just the active block -- and
// are accumulated from the right into the zrows rows stored directly above S in the same array (rows -zrows..-1).
// Returns the number of eigenvalues that failed to converge (0 = success); converged ones are in wout[...], for
// failures wout holds the current diagonal.
template <bool WANTT>
SD_DEV int smem_hqr(const Grp& g, cplx* S, int lds, int m, int zrows, cplx* wout, SmallCtl* ctl, Rot* cur, Rot* rec,
                    const AedStop* stop = nullptr, int* ns_out = nullptr, int* istop_out = nullptr) {
  const double smlnum = SD_SAFMIN * ((double)m / SD_ULP);
  int I = m - 1;
  int its = 0;
  const int itmax = 30 * (m > 10 ? m : 10);
  int total = 0;
  SmemAt at; at.S = S; at.lds = lds;
  int ns_und = -1, ilow = -1;                  // undeflated count once the deflation test has failed; stop when I <= ilow
  while (I > ilow) {
    if (g.tid == 0) ctl->L = 0;
    grp_sync(g);
    for (int k = 1 + g.tid; k <= I; k += g.nt)
      if (negligible_subdiag(at, k, 0, m - 1, smlnum)) grp_amax(&ctl->L, k);
    grp_sync(g);
    const int L = ctl->L;
    if (L >= I) {
      if (g.tid == 0) {
        if (L > 0) S[L + (L - 1) * lds] = mk(0.0, 0.0);
        wout[I] = S[I + I * lds];
      }
      if (WANTT && stop && ns_und < 0) {       // deflation test of the eigenvalue that has just converged (uniform)
        double foo = cabs1(S[I + I * lds]);
        if (foo == 0.0) foo = stop->s1;
        if (!(stop->s1 * cabs1(S[-zrows + I * lds]) <= fmax(stop->smlnum, SD_ULP * foo))) {
          ns_und = I + 1;
          const int nd = m - ns_und;
          ilow = (nd >= stop->nd_skip) ? I : ns_und - 1 - stop->want;
          if (ilow < -1) ilow = -1;
        }
      }
      I -= 1; its = 0;
      grp_sync(g);
      continue;
    }
    if (total++ >= itmax) {
      // give up: report the diagonal for what is left
      for (int k = g.tid; k <= I; k += g.nt) wout[k] = S[k + k * lds];
      grp_sync(g);
      return I + 1;
    }
    its += 1;
    const bool exceptional = (its == 10 || its == 20);
    const int nsh = (!exceptional && I - L + 1 >= 4) ? 2 : 1;
    if (g.tid == 0) {
      if (L > 0) S[L + (L - 1) * lds] = mk(0.0, 0.0);
      cplx t;
      if (its == 10) {
        t = mk(0.75 * fabs(S[L + 1 + L * lds].re), 0.0) + S[L + L * lds];
      } else if (its == 20) {
        t = mk(0.75 * fabs(S[I + (I - 1) * lds].re), 0.0) + S[I + I * lds];
      } else {
        eig2x2_fast(S[I - 1 + (I - 1) * lds], S[I - 1 + I * lds], S[I + (I - 1) * lds], S[I + I * lds], t, ctl->shift[1]);
      }
      ctl->shift[0] = t;
    }
    grp_sync(g);
    const int T = (I - L) + 2 * nsh - 2;
#ifdef STAB_EMU_COUNT
    g_emu_cnt[WANTT ? 0 : 2] += 1; g_emu_cnt[WANTT ? 1 : 3] += T; if (nsh == 1) g_emu_cnt[4] += 1;
#endif
#ifdef STAB_MB_COUNT
    if (g.tid == 0 && blockIdx.x == 0) { g_steps[0] += 1; g_steps[1] += T; }
#endif
    // compile-time chain length: the index arithmetic folds and a line job carries exactly the chain's rotations
    if (nsh == 2) {
      if (WANTT) chase_tiles<2>(g, S, lds, 0, m, L, I, ctl->shift, nsh, 0, T, rec, cur, -zrows);
      else chase_tiles<2>(g, S + L + (size_t)L * lds, lds, L, I - L + 1, L, I, ctl->shift, nsh, 0, T, rec, cur, 0);
    } else {
      if (WANTT) chase_tiles<1>(g, S, lds, 0, m, L, I, ctl->shift, nsh, 0, T, rec, cur, -zrows);
      else chase_tiles<1>(g, S + L + (size_t)L * lds, lds, L, I - L + 1, L, I, ctl->shift, nsh, 0, T, rec, cur, 0);
    }
  }
  if (ns_out) *ns_out = ns_und < 0 ? 0 : ns_und;
  if (istop_out) *istop_out = I;
  return 0;
}

struct HqrSmem {
  cplx* win;     // W * ldw; also the stacked [Z; T] arrays of the deflation window: (2 nw + 1) * nw <= W * ldw
  int ldw, W;
  int nw, nibble; // deflation window (0: no aggressive early deflation) and ZLAQR0's NIBBLE (percent)
  Rot* rec;     // steps_max * ns_max
  int steps_max, ns_max;
  Rot* cur;     // 2 * ns_max (double-buffered by chase_tiles)
  cplx* shifts;  // ns_max
  cplx* sm;      // ns_max * (ns_max + 1)   trailing block for the shift computation
  SmallCtl* ctl;
  long long* prof;   // optional cycle counters (debug), in SHARED memory (copied out by the kernel at the end): 0 scan, 1 shifts, 2 window io, 3 chase, 4 slabs, 5 -, 6 small blocks, 7 #sweeps, 8 #passes, 9 sweeps (outer clock), 10 AED write-back + slab, 11 #AED, 12 #deflated by AED, 13 AED Schur, 14 AED count + restore, 15 AED (outer clock)
};

#if defined(STAB_EMU)
#define HQR_PROF_START()
#define HQR_PROF(i)
#define HQR_COUNT(i)
#else
#define HQR_PROF_START() long long pt0_ = (sh.prof && c.tid == 0) ? clock64() : 0
#define HQR_PROF(i) do { if (sh.prof && c.tid == 0) { long long t1_ = clock64(); sh.prof[i] += t1_ - pt0_; pt0_ = t1_; } } while (0)
#define HQR_COUNT(i) do { if (sh.prof && c.tid == 0) sh.prof[i] += 1; } while (0)
#endif

// ---------------------------------------------------------------------------------------------
// Off-window (slab) update of ONE line -- a column of the left slab (positions = rows, stride 1)
// or a row of the right slab (positions = columns, stride ldh) -- with the reflectors recorded
// for the time steps [ta, tb) of a chain of NSB bulges, as a register-resident systolic pipeline:
// bulge b is a stage that holds one element (st[b]), consumes the element its predecessor
// emitted one time step earlier (pipe[b]; stage 0 reads memory), applies reflector (t, b) to the
// pair, emits the upper element to stage b+1 (the last stage writes memory) and keeps the lower
// one.  Every element of the line is loaded once and stored once per pass; all 2*NSB in-flight
// elements live in registers (static indexing: the stage loop is unrolled).  Stage b at time t
// acts on positions (L+s, L+s+1), s = t-2b, for 0 <= s <= smax; s = -1 is its start-up
// (absorb the first element), s = smax+1 its flush.  STEADY: every stage is active at every
// step of the pass (no start-up / flush inside), so the mode tests vanish.
// ---------------------------------------------------------------------------------------------
constexpr int SLAB_PF = 4;   // input prefetch distance (time steps)

template <int NSB, bool RIGHT, bool STEADY>
SD_NOINLINE void slab_line(cplx* base, size_t stride, int L, int smax, int ta, int tb, const Rot* rec) {
  cplx st[NSB], pipe[NSB];
#pragma unroll
  for (int b = 0; b < NSB; ++b) {                         // prologue: elements in flight at time ta
    const int s = ta - 2 * b;
    st[b] = mk(0.0, 0.0); pipe[b] = mk(0.0, 0.0);
    if (STEADY || (s >= 0 && s <= smax + 1)) st[b] = base[(size_t)(L + s) * stride];
    if (b >= 1 && (STEADY || (s + 1 >= 0 && s + 1 <= smax + 1))) pipe[b] = base[(size_t)(L + s + 1) * stride];
  }
  cplx inq[SLAB_PF];                                      // stage 0's input: position L+t+1 at time t
#pragma unroll
  for (int u = 0; u < SLAB_PF; ++u) {
    const int t = ta + u;
    inq[u] = (t < tb && t <= smax) ? base[(size_t)(L + t + 1) * stride] : mk(0.0, 0.0);
  }
  for (int t0 = ta; t0 < tb; t0 += SLAB_PF) {
#pragma unroll
    for (int u = 0; u < SLAB_PF; ++u) {
      const int t = t0 + u;
      if (t < tb) {
        const cplx xin = inq[u];
        {
          const int tn = t + SLAB_PF;
          if (tn < tb && tn <= smax) inq[u] = base[(size_t)(L + tn + 1) * stride];
        }
        const Rot* rt = rec + (size_t)(t - ta) * NSB;
#pragma unroll
        for (int bb = 0; bb < NSB; ++bb) {
          const int b = NSB - 1 - bb;                     // descending: consume pipe[b] before stage b-1 refills it
          const int s = t - 2 * b;
          if (STEADY || (s >= 0 && s <= smax)) {
            cplx x1 = st[b];
            cplx x2 = (b == 0) ? xin : pipe[b];
            const Rot r = rt[rec_slot(NSB, b)];
            if (RIGHT) apply_right(r, x1, x2); else apply_left(r, x1, x2);   // c = 1, s = 0 is an exact identity
            if (b == NSB - 1) base[(size_t)(L + s) * stride] = x1; else pipe[b + 1 < NSB ? b + 1 : b] = x1;
            st[b] = x2;
          } else if (s == -1) {
            if (b >= 1) st[b] = pipe[b];
          } else if (s == smax + 1) {
            if (b == NSB - 1) base[(size_t)(L + s) * stride] = st[b]; else pipe[b + 1 < NSB ? b + 1 : b] = st[b];
          }
        }
      }
    }
  }
#pragma unroll
  for (int b = 0; b < NSB; ++b) {                         // epilogue: park the elements still in flight
    const int s = tb - 2 * b;
    if (STEADY || (s >= 0 && s <= smax + 1)) base[(size_t)(L + s) * stride] = st[b];
    if (b >= 1 && (STEADY || (s + 1 >= 0 && s + 1 <= smax + 1))) base[(size_t)(L + s + 1) * stride] = pipe[b];
  }
}

#ifndef STAB_EMU
// The same pipeline split over a LANE PAIR (even lane: stages 0..NSB/2-1 and the input stream, odd
// lane: stages NSB/2..NSB-1 and the output stream); the element crossing the split travels by one
// warp shuffle per time step.  Halves the registers per thread (two CTAs per SM) at equal work.
//
// Register roles alternate with the parity of the step instead of being copied: a stage's resident
// element lives in A[bl] and its input arrives in B[bl]; after the update the emitted value is
// WRITTEN (as the result of the arithmetic, not copied) to A[bl+1], which the next stage has just
// vacated, and the kept value stays in B[bl] -- so at the next step B holds the resident elements
// and A the inputs.  One step:
template <int NL, bool RIGHT, bool STEADY>
SD_DEV void slab_pair_step(cplx (&A)[NL], cplx (&B)[NL], cplx& outp, const Rot* __restrict__ rt, int part, int b0, int t,
                           int L, int smax, cplx* base, size_t stride, bool act) {
#pragma unroll
  for (int bb = 0; bb < NL; ++bb) {
    const int bl = NL - 1 - bb, b = b0 + bl, s = t - 2 * b;    // descending: stage bl+1 has vacated A[bl+1]
    if (STEADY || (s >= 0 && s <= smax)) {
      cplx x1 = A[bl];
      cplx x2 = B[bl];
      const Rot r = rt[2 * bl];                                  // rec_slot: stage b0 + bl lives at 2 bl + part
      if (RIGHT) apply_right(r, x1, x2); else apply_left(r, x1, x2);   // c = 1, s = 0 is an exact identity
      if (bl == NL - 1) { if (part == 1) { if (act) base[(size_t)(L + s) * stride] = x1; } else outp = x1; }
      else A[bl + 1 < NL ? bl + 1 : bl] = x1;
      B[bl] = x2;
    } else if (s == smax + 1) {                                 // flush the resident element
      if (bl == NL - 1) { if (part == 1) { if (act) base[(size_t)(L + s) * stride] = A[bl]; } else outp = A[bl]; }
      else A[bl + 1 < NL ? bl + 1 : bl] = A[bl];
    }                                                           // s == -1: the arrived element (B[bl]) becomes resident by the role swap
  }
}

template <int NSB, bool RIGHT, bool STEADY>
// Every lane of the warp runs the schedule (the shuffles take the constant full mask: a computed mask costs a
// MATCH / REDUX / VOTE / divergence-check sequence per time step); a lane pair without a line (`act` false, only in
// the last unit of a slab) computes on zeros and never touches memory.
SD_NOINLINE void slab_line_pair(cplx* base, size_t stride, int L, int smax, int ta, int tb, const Rot* rec, bool act, int part) {
  constexpr unsigned mask = 0xffffffffu;
  constexpr int NL = NSB / 2;
  static_assert(SLAB_PF % 2 == 0, "the role alternation needs an even unroll");
  const int b0 = part * NL;
  cplx P[NL], Q[NL];                                        // at even (t - ta): P resident, Q input; odd: swapped
#pragma unroll
  for (int bl = 0; bl < NL; ++bl) {                       // prologue: elements in flight at time ta
    const int b = b0 + bl, s = ta - 2 * b;
    P[bl] = mk(0.0, 0.0); Q[bl] = mk(0.0, 0.0);
    if (act && (STEADY || (s >= 0 && s <= smax + 1))) P[bl] = base[(size_t)(L + s) * stride];
    if (act && b >= 1 && (STEADY || (s + 1 >= 0 && s + 1 <= smax + 1))) Q[bl] = base[(size_t)(L + s + 1) * stride];
  }
  const bool feed = act && part == 0;
  cplx inq[SLAB_PF];
#pragma unroll
  for (int u = 0; u < SLAB_PF; ++u) {
    const int t = ta + u;
    inq[u] = (feed && t < tb && t <= smax) ? base[(size_t)(L + t + 1) * stride] : mk(0.0, 0.0);
  }
  cplx outp = mk(0.0, 0.0);                               // emission of this lane's last stage at the previous step
  for (int t0 = ta; t0 < tb; t0 += SLAB_PF) {
#pragma unroll
    for (int u = 0; u < SLAB_PF; ++u) {
      const int t = t0 + u;
      if (t < tb) {
        const cplx xin = inq[u];
        {
          const int tn = t + SLAB_PF;
          if (feed && tn < tb && tn <= smax) inq[u] = base[(size_t)(L + tn + 1) * stride];
        }
        const cplx got = mk(__shfl_xor_sync(mask, outp.re, 1), __shfl_xor_sync(mask, outp.im, 1));
        const Rot* rt = rec + (size_t)(t - ta) * NSB + part;
        if ((u & 1) == 0) {
          if (part == 0) Q[0] = xin; else if (t > ta) Q[0] = got;
          slab_pair_step<NL, RIGHT, STEADY>(P, Q, outp, rt, part, b0, t, L, smax, base, stride, act);
        } else {
          if (part == 0) P[0] = xin; else P[0] = got;
          slab_pair_step<NL, RIGHT, STEADY>(Q, P, outp, rt, part, b0, t, L, smax, base, stride, act);
        }
      }
    }
  }
  const cplx got = mk(__shfl_xor_sync(mask, outp.re, 1), __shfl_xor_sync(mask, outp.im, 1));
  if (((tb - ta) & 1) == 0) {
    if (part == 1 && tb > ta) Q[0] = got;
#pragma unroll
    for (int bl = 0; bl < NL; ++bl) {                     // epilogue: park the elements still in flight
      const int b = b0 + bl, s = tb - 2 * b;
      if (act && (STEADY || (s >= 0 && s <= smax + 1))) base[(size_t)(L + s) * stride] = P[bl];
      if (act && b >= 1 && (STEADY || (s + 1 >= 0 && s + 1 <= smax + 1))) base[(size_t)(L + s + 1) * stride] = Q[bl];
    }
  } else {
    if (part == 1) P[0] = got;
#pragma unroll
    for (int bl = 0; bl < NL; ++bl) {
      const int b = b0 + bl, s = tb - 2 * b;
      if (act && (STEADY || (s >= 0 && s <= smax + 1))) base[(size_t)(L + s) * stride] = Q[bl];
      if (act && b >= 1 && (STEADY || (s + 1 >= 0 && s + 1 <= smax + 1))) base[(size_t)(L + s + 1) * stride] = P[bl];
    }
  }
}
#endif

template <int NSB>
SD_DEV void slabs_stream(const Cta& c, cplx* H, int ldh, int L, int I, int g0, int g1, int ta, int tb, const Rot* rec) {
  const int smax = I - 1 - L;
  const bool steady = (ta >= 2 * (NSB - 1)) && (tb - 1 <= smax);
#ifndef STAB_EMU
  {
    // work unit = 16 lines (one warp of lane pairs); the units of BOTH slabs are dealt round-robin to the
    // warps, so a warp never idles through a whole round because the other slab's line count is ragged
    const int part = c.tid & 1, lp = c.lane >> 1;
    const int nleft = I - g1, nright = g0 - L;              // left slab columns (g1, I], right slab rows [L, g0)
    const int UL = (nleft + 15) >> 4, UR = (nright + 15) >> 4;
    const int p0 = L + (ta > 2 * (NSB - 1) ? ta - 2 * (NSB - 1) : 0);     // first position this pass touches
    int p1 = L + tb + 1; if (p1 > I) p1 = I;                             // last one
    for (int u = c.wid; u < UL + UR; u += c.nw) {
#ifndef STAB_QR_NO_L2PF
      {
        // The 2 x NSB elements a line has in flight are loaded at the start of its unit and nothing hides that latency
        // (ncu: 20 % of the slab samples wait on global loads): pull the NEXT unit of this warp into L2 now, so that its
        // prologue and its streamed inputs are L2 hits by the time they are issued.  No registers, no shared memory.
        const int u2 = u + c.nw;
        if (u2 < UL) {                                                   // 16 columns x (p1 - p0 + 1) rows, contiguous per column
          const int li2 = (u2 << 4) + (c.lane >> 1);
          if (li2 < nleft) {
            const char* q = reinterpret_cast<const char*>(H + (size_t)(g1 + 1 + li2) * ldh + p0);
            const int nb = (p1 - p0 + 1) * 16;
            for (int o = (c.lane & 1) * 128; o < nb; o += 256) asm volatile("prefetch.global.L2 [%0];" ::"l"(q + o));
          }
        } else if (u2 < UL + UR) {                                       // 16 rows (256 B) of every touched column
          const int r0 = L + ((u2 - UL) << 4);
          for (int col = p0 + (c.lane >> 1); col <= p1; col += 16) {
            const char* q = reinterpret_cast<const char*>(H + r0 + (size_t)col * ldh) + (c.lane & 1) * 128;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(q));
          }
        }
      }
#endif
      if (u < UL) {
        const int li = (u << 4) + lp;
        const bool act = li < nleft;
        cplx* line = H + (size_t)(g1 + 1 + (act ? li : 0)) * ldh;
        if (steady) slab_line_pair<NSB, false, true>(line, 1, L, smax, ta, tb, rec, act, part);
        else slab_line_pair<NSB, false, false>(line, 1, L, smax, ta, tb, rec, act, part);
      } else {
        const int li = ((u - UL) << 4) + lp;
        const bool act = li < nright;
        cplx* line = H + (L + (act ? li : 0));
        if (steady) slab_line_pair<NSB, true, true>(line, (size_t)ldh, L, smax, ta, tb, rec, act, part);
        else slab_line_pair<NSB, true, false>(line, (size_t)ldh, L, smax, ta, tb, rec, act, part);
      }
    }
  }
#else
  // left slab: rows of the window, columns (g1, I]; one thread per column
  for (int col = g1 + 1 + c.tid; col <= I; col += c.nt) {
    if (steady) slab_line<NSB, false, true>(H + (size_t)col * ldh, 1, L, smax, ta, tb, rec);
    else slab_line<NSB, false, false>(H + (size_t)col * ldh, 1, L, smax, ta, tb, rec);
  }
  // right slab: rows [L, g0), columns of the chain; one thread per row
  for (int row = L + c.tid; row < g0; row += c.nt) {
    if (steady) slab_line<NSB, true, true>(H + row, (size_t)ldh, L, smax, ta, tb, rec);
    else slab_line<NSB, true, false>(H + row, (size_t)ldh, L, smax, ta, tb, rec);
  }
#endif
}

// One multishift sweep over the active block [L, I] of the global Hessenberg matrix H.
SD_DEV void sweep_multishift(const Cta& c, const HqrSmem& sh, cplx* H, int ldh, int L, int I, int ns) {
  Grp g; g.tid = c.tid; g.nt = c.nt; g.warp = false;
  const int W = sh.W, ldw = sh.ldw;
  const int T = I - L + 2 * ns - 2;
  int ta = 0;
  HQR_PROF_START();
  HQR_COUNT(7);
  while (ta < T) {
    HQR_COUNT(8);
    int tb, g0;
    if (ta == 0) {
      int first = W - 2; if (first > sh.steps_max) first = sh.steps_max;
      tb = first; g0 = L;
    } else {
      int adv = W - 2 * ns - 1; if (adv > sh.steps_max) adv = sh.steps_max;
      tb = ta + adv; g0 = L + ta - 2 * (ns - 1) - 1;
    }
    if (tb > T) tb = T;
    int g1 = L + tb + 1; if (g1 > I) g1 = I;
    const int wsz = g1 - g0 + 1;
    // load the window (upper Hessenberg part plus the two sub-diagonals that can hold bulges)
    for (int q = c.tid; q < wsz * wsz; q += c.nt) {
      const int col = q / wsz, row = q - col * wsz;
      sh.win[row + col * ldw] = (row <= col + 2) ? H[(g0 + row) + (size_t)(g0 + col) * ldh] : mk(0.0, 0.0);
    }
    cta_sync();
    HQR_PROF(2);
    if (ns == 16) chase_tiles<16>(g, sh.win, ldw, g0, wsz, L, I, sh.shifts, ns, ta, tb, sh.rec, sh.cur);
    else chase_tiles<0>(g, sh.win, ldw, g0, wsz, L, I, sh.shifts, ns, ta, tb, sh.rec, sh.cur);
    HQR_PROF(3);
    for (int q = c.tid; q < wsz * wsz; q += c.nt) {
      const int col = q / wsz, row = q - col * wsz;
      if (row <= col + 2) H[(g0 + row) + (size_t)(g0 + col) * ldh] = sh.win[row + col * ldw];
    }
    HQR_PROF(2);
    const int smax = I - 1 - L;
    if (ns == 16) {
      cta_sync();                                   // window stores precede slab reads of neighbouring entries? (disjoint) -- rec is final
      slabs_stream<16>(c, H, ldh, L, I, g0, g1, ta, tb, sh.rec);
      cta_sync();
      HQR_PROF(4);
      ta = tb;
      continue;
    }
    // left slab: rows [g0, g1], columns (g1, I]; thread per column, bulge-major with carry
    for (int col = g1 + 1 + c.tid; col <= I; col += c.nt) {
      cplx* hc = H + (size_t)col * ldh;
      for (int b = 0; b < ns; ++b) {
        int t0 = ta > 2 * b ? ta : 2 * b;
        int t1 = tb - 1; if (t1 > 2 * b + smax) t1 = 2 * b + smax;
        if (t0 > t1) continue;
        int k = L + t0 - 2 * b;
        cplx x = hc[k];
        for (int t = t0; t <= t1; ++t, ++k) {
          cplx y = hc[k + 1];
          const Rot r = sh.rec[(t - ta) * ns + rec_slot(ns, b)];
          if (!rot_is_identity(r)) apply_left(r, x, y);
          hc[k] = x;
          x = y;
        }
        hc[k] = x;
      }
    }
    cta_sync();
    HQR_PROF(4);
    // right slab: rows [L, g0), columns of the chain; thread per row
    for (int row = L + c.tid; row < g0; row += c.nt) {
      cplx* hr = H + row;
      for (int b = 0; b < ns; ++b) {
        int t0 = ta > 2 * b ? ta : 2 * b;
        int t1 = tb - 1; if (t1 > 2 * b + smax) t1 = 2 * b + smax;
        if (t0 > t1) continue;
        int k = L + t0 - 2 * b;
        cplx x = hr[(size_t)k * ldh];
        for (int t = t0; t <= t1; ++t, ++k) {
          cplx y = hr[(size_t)(k + 1) * ldh];
          const Rot r = sh.rec[(t - ta) * ns + rec_slot(ns, b)];
          if (!rot_is_identity(r)) apply_right(r, x, y);
          hr[(size_t)k * ldh] = x;
          x = y;
        }
        hr[(size_t)k * ldh] = x;
      }
    }
    cta_sync();
    HQR_PROF(5);
    ta = tb;
  }
}

// ---------------------------------------------------------------------------------------------
// Aggressive early deflation (Braman, Byers, Mathias; ZLAQR3's logic for eigenvalues only) on the trailing jw x jw
// window of the active block [L, I], kw = I-jw+1 > L, with s = H(kw, kw-1):
//   1. T = Schur form of the window, Z its Schur vectors (smem_hqr<true> on the stacked [Z; T] shared-memory array);
//   2. in the basis Z the window couples to the rest of the block only through the spike s conj(Z(0, :)); the
//      trailing eigenvalues whose spike entries are negligible (|s| |Z(0,j)| <= max(smlnum, ulp |T(j,j)|)) are
//      deflated -- nd of them, counted from the bottom up to the first one that fails (no reordering of the Schur form:
//      on the reference's operators the count without ZTREXC is within a few per cent of the count with it), which
//      lets step 1 stop early (AedStop): only the deflated eigenvalues and the next sweep's shifts are converged;
//   3. if nd > 0 the undeflated part (ns = jw - nd) returns to Hessenberg form: a reflector takes the spike to a multiple
//      of e1, ZGEHD2 on the ns x ns block (in shared memory, Z accumulating), H(kw, kw-1) = s conj(Z(0,0)), and the
//      columns of the block above the window are multiplied by Z(:, 0:ns) (aed_slab).  The coupling block T(0:ns, ns:jw)
//      is dropped: eigenvalues only.
// The diagonal of T(0:ns, 0:ns) (before step 3) are the shifts of the next sweep (ZLAQR0), written to sh.shifts --
// nshift of them (0: none usable, the caller computes its own); fewer than nshift_want distinct ones are cycled.
// Returns nd; the deflated eigenvalues are stored in w[kw+ns .. I].  H is untouched when nd == 0.
// ---------------------------------------------------------------------------------------------
template <int KH>
SD_NOINLINE void aed_slab(int tid, int nt, cplx* H, int ldh, int L, int kw, int jw, int ns, const cplx* Zs, int lds) {
  // H(L:kw-1, kw:kw+ns-1) = H(L:kw-1, kw:kw+jw-1) Z(:, 0:ns-1), in place: a row is held in registers, half of it per
  // lane of a lane pair (KH entries each, 2 KH >= jw), the partial sums meet by one shuffle per output; a store of
  // output j happens after the shuffle, i.e. after BOTH lanes have loaded (and consumed) their whole half row.
#ifdef STAB_EMU
  (void)tid; (void)nt;
  cplx row[2 * KH];
  for (int r = L; r < kw; ++r) {
    for (int k = 0; k < jw; ++k) row[k] = H[r + (size_t)(kw + k) * ldh];
    for (int j = 0; j < ns; ++j) {
      cplx acc = mk(0.0, 0.0);
      for (int k = 0; k < jw; ++k) fma_acc(acc, row[k], Zs[k + j * lds]);
      H[r + (size_t)(kw + j) * ldh] = acc;
    }
  }
#else
  const int part = tid & 1, k0 = part * KH;
  const int rows_per_round = nt >> 1;
  for (int r0 = L; r0 < kw; r0 += rows_per_round) {
    const int r = r0 + (tid >> 1);
    const bool act = r < kw;
    cplx a[KH];
#pragma unroll
    for (int i = 0; i < KH; ++i) a[i] = (act && k0 + i < jw) ? H[r + (size_t)(kw + k0 + i) * ldh] : mk(0.0, 0.0);
    for (int j = 0; j < ns; j += 2) {
      const int j1 = (j + 1 < ns) ? j + 1 : j;
      const cplx* z0 = Zs + k0 + j * lds;
      const cplx* z1 = Zs + k0 + j1 * lds;
      cplx acc0 = mk(0.0, 0.0), acc1 = mk(0.0, 0.0);
#pragma unroll
      for (int i = 0; i < KH; ++i) { fma_acc(acc0, a[i], z0[i]); fma_acc(acc1, a[i], z1[i]); }
      acc0.re += __shfl_xor_sync(0xffffffffu, acc0.re, 1); acc0.im += __shfl_xor_sync(0xffffffffu, acc0.im, 1);
      acc1.re += __shfl_xor_sync(0xffffffffu, acc1.re, 1); acc1.im += __shfl_xor_sync(0xffffffffu, acc1.im, 1);
      if (act) {
        if (part == 0) H[r + (size_t)(kw + j) * ldh] = acc0;
        else if (j + 1 < ns) H[r + (size_t)(kw + j + 1) * ldh] = acc1;
      }
    }
  }
#endif
}

// One Householder step of the return to Hessenberg form on the stacked [Z; T] array: P = I - tau v v^H acts on the
// indices j0 .. ns-1 (v in shared memory, v[0] = 1): T := P^H T P on the leading ns x ns block, Z := Z P.
SD_DEV void aed_reflect(const Cta& c, cplx* Zs, int lds, int jw, int ns, int j0, const cplx* v, cplx tau) {
  const int len = ns - j0;
  cplx* Ts = Zs + jw;
  // right application on the stacked rows (Z: all jw rows, T: rows 0..ns-1), columns j0..ns-1; a thread owns a row
  for (int q = c.tid; q < jw + ns; q += c.nt) {
    cplx* row = Zs + q + (size_t)j0 * lds;
    cplx x0 = mk(0.0, 0.0), x1 = mk(0.0, 0.0);
    int i = 0;
    for (; i + 1 < len; i += 2) { fma_acc(x0, row[(size_t)i * lds], v[i]); fma_acc(x1, row[(size_t)(i + 1) * lds], v[i + 1]); }
    if (i < len) fma_acc(x0, row[(size_t)i * lds], v[i]);
    const cplx x = (x0 + x1) * tau;
    for (i = 0; i < len; ++i) fms_acc(row[(size_t)i * lds], x, conj(v[i]));
  }
  cta_sync();
  // left application on T(j0:ns-1, j0':ns-1); a thread owns a column.  (For j0 >= 1 the columns left of j0 are already
  // reduced: column j0-1 holds the reflector's source and is set by the caller.)
  const cplx tauc = conj(tau);
  for (int col = j0 + c.tid; col < ns; col += c.nt) {
    cplx* cp = Ts + j0 + (size_t)col * lds;
    cplx y0 = mk(0.0, 0.0), y1 = mk(0.0, 0.0);
    int i = 0;
    for (; i + 1 < len; i += 2) { fma_acc_conj(y0, v[i], cp[i]); fma_acc_conj(y1, v[i + 1], cp[i + 1]); }
    if (i < len) fma_acc_conj(y0, v[i], cp[i]);
    const cplx y = (y0 + y1) * tauc;
    for (i = 0; i < len; ++i) fms_acc(cp[i], v[i], y);
  }
  cta_sync();
}

SD_DEV int aed(const Cta& c, const HqrSmem& sh, cplx* H, int ldh, int L, int I, int jw, double smlnum, cplx* w,
               int nshift_want, int* nshift) {
  Grp g; g.tid = c.tid; g.nt = c.nt; g.warp = false;
  const int kw = I - jw + 1;
  const int lds = 2 * jw + 1;
  cplx* Zs = sh.win;                  // Z(r, c) = Zs[r + c lds]
  cplx* Ts = sh.win + jw;             // T(r, c) = Ts[r + c lds]
  const cplx s = (kw > L) ? H[kw + (size_t)(kw - 1) * ldh] : mk(0.0, 0.0);
  HQR_PROF_START();
  HQR_COUNT(11);
  for (int q = c.tid; q < jw * jw; q += c.nt) {
    const int col = q / jw, row = q - col * jw;
    Ts[row + col * lds] = (row <= col + 1) ? H[(kw + row) + (size_t)(kw + col) * ldh] : mk(0.0, 0.0);
    Zs[row + col * lds] = mk(row == col ? 1.0 : 0.0, 0.0);
  }
  cta_sync();
  // Schur factorisation with the deflation test taken as the eigenvalues converge (AedStop): on return ns is the
  // undeflated count and the positions istop+1 .. jw-1 are triangular (the block 0 .. istop is still Hessenberg)
  AedStop st; st.s1 = cabs1(s); st.smlnum = smlnum; st.nd_skip = (sh.nibble * sh.nw) / 100 + 1; st.want = nshift_want;
  int ns = 0, istop = -1;
  const int bad = smem_hqr<true>(g, Ts, lds, jw, jw, sh.sm, sh.ctl, sh.cur, sh.rec, &st, &ns, &istop);
  HQR_PROF(13);
  *nshift = 0;
  if (bad) return 0;                  // the window did not reach Schur form (never seen): a plain sweep follows
  const int nd = jw - ns;
  // ---- shifts of the next sweep: the converged, undeflated eigenvalues (positions istop+1 .. ns-1)
  const int navail = ns - 1 - istop;
  if (navail >= 2) {
    for (int b = c.tid; b < nshift_want; b += c.nt) { const int j = ns - 1 - (b % navail); sh.shifts[b] = Ts[j + j * lds]; }
    *nshift = nshift_want;
  }
  if (nd == 0) { cta_sync(); HQR_PROF(10); return 0; }
  if (sh.prof && c.tid == 0) sh.prof[12] += nd;
  for (int j = ns + c.tid; j < jw; j += c.nt) w[kw + j] = Ts[j + j * lds];
  if (ns == 0) {                      // the whole window deflated
    if (c.tid == 0 && kw > L) H[kw + (size_t)(kw - 1) * ldh] = mk(0.0, 0.0);
    cta_sync();
    HQR_PROF(10);
    return nd;
  }
  cta_sync();                         // shifts and eigenvalues are read before T changes
  // ---- return of T(0:ns, 0:ns) + spike to Hessenberg form
  if (ns > 1 && !is_zero(s)) {
    cplx* v = sh.sm;
    for (int i = c.tid; i < ns; i += c.nt) v[i] = conj(Zs[i * lds]);
    cta_sync();
    cplx beta = v[0];
    cta_sync();
    cplx tau = cta_zlarfg(c, ns, beta, v + 1);
    cta_sync();
    if (c.tid == 0) v[0] = mk(1.0, 0.0);
    cta_sync();
    if (!is_zero(tau)) aed_reflect(c, Zs, lds, jw, ns, 0, v, tau);
    for (int j = 0; j + 2 < ns; ++j) {              // ZGEHD2: reflector from T(j+1:ns-1, j)
      cplx* col = Ts + (size_t)j * lds;
      cplx alpha = col[j + 1];
      cta_sync();
      tau = cta_zlarfg(c, ns - j - 1, alpha, col + j + 2);
      cta_sync();
      for (int i = 1 + c.tid; i < ns - j - 1; i += c.nt) { v[i] = col[j + 1 + i]; col[j + 1 + i] = mk(0.0, 0.0); }
      if (c.tid == 0) { v[0] = mk(1.0, 0.0); col[j + 1] = alpha; }
      cta_sync();
      if (!is_zero(tau)) aed_reflect(c, Zs, lds, jw, ns, j + 1, v, tau);
    }
  }
  HQR_PROF(14);
  // ---- the reduced window back in place (with explicit zeros on the second subdiagonal, which window loads read)
  if (c.tid == 0) {
    if (kw > L) H[kw + (size_t)(kw - 1) * ldh] = s * conj(Zs[0]);
    H[(kw + ns) + (size_t)(kw + ns - 1) * ldh] = mk(0.0, 0.0);
  }
  for (int q = c.tid; q < ns * ns; q += c.nt) {
    const int col = q / ns, row = q - col * ns;
    if (row <= col + 2) H[(kw + row) + (size_t)(kw + col) * ldh] = (row <= col + 1) ? Ts[row + col * lds] : mk(0.0, 0.0);
  }
  // ---- the block above the window: H(L:kw-1, kw:kw+ns-1) = H(L:kw-1, kw:I) Z(:, 0:ns-1)
  if (jw <= 32) aed_slab<16>(c.tid, c.nt, H, ldh, L, kw, jw, ns, Zs, lds);
  else aed_slab<23>(c.tid, c.nt, H, ldh, L, kw, jw, ns, Zs, lds);
  cta_sync();
  HQR_PROF(10);
  return nd;
}

// Eigenvalues of the Hessenberg matrix H (entries below the first subdiagonal must be zero) on
// [ilo, ihi]; entries outside are read off the diagonal.  Returns 0 or the number of unconverged
// eigenvalues (LAPACK-style info > 0).
SD_DEV int cta_hqr(const Cta& c, const HqrSmem& sh, cplx* H, int n, int ldh, int ilo, int ihi, cplx* w) {
  Grp g; g.tid = c.tid; g.nt = c.nt; g.warp = false;
  for (int j = c.tid; j < n; j += c.nt)
    if (j < ilo || j > ihi) w[j] = H[j + (size_t)j * ldh];
  const int nh = ihi - ilo + 1;
  const double smlnum = SD_SAFMIN * ((double)nh / SD_ULP);
  GlobAt at; at.H = H; at.ldh = ldh;
  int I = ihi;
  int stagn = 0;
  long total = 0;
  const long itmax = 30L * (nh > 10 ? nh : 10);
  int info = 0;
  cta_sync();
  HQR_PROF_START();
  while (I >= ilo) {
    HQR_PROF(9);
    // largest k in (ilo, I] whose subdiagonal is negligible
    int best = ilo;
    for (int k = ilo + 1 + c.tid; k <= I; k += c.nt)
      if (negligible_subdiag(at, k, ilo, ihi, smlnum)) best = k;
    const int L = cta_max_i(c, best);
    HQR_PROF(0);
    if (L > ilo && c.tid == 0) H[L + (size_t)(L - 1) * ldh] = mk(0.0, 0.0);
    if (L == I) {
      if (c.tid == 0) w[I] = H[I + (size_t)I * ldh];
      I -= 1; stagn = 0;
      cta_sync();
      continue;
    }
    const int m = I - L + 1;
    if (m <= sh.W) {
      // finish this block in shared memory
      for (int q = c.tid; q < m * m; q += c.nt) {
        const int col = q / m, row = q - col * m;
        sh.win[row + col * sh.ldw] = (row <= col + 1) ? H[(L + row) + (size_t)(L + col) * ldh] : mk(0.0, 0.0);
      }
      cta_sync();
      int bad = 0;
      if (m <= 40) {                           // small: one warp with warp-level barriers beats CTA barriers
        if (c.wid == 0) {
          Grp gw; gw.tid = c.lane; gw.nt = c.ws; gw.warp = true;
          bad = smem_hqr<false>(gw, sh.win, sh.ldw, m, 0, w + L, sh.ctl, sh.cur, sh.rec);
          if (c.lane == 0) sh.ctl->pad = bad;
        }
        cta_sync();
        bad = sh.ctl->pad;
      } else {
        bad = smem_hqr<false>(g, sh.win, sh.ldw, m, 0, w + L, sh.ctl, sh.cur, sh.rec);
      }
      if (bad) info += bad;
      I = L - 1; stagn = 0;
      cta_sync();
      HQR_PROF(6);
      continue;
    }
    if (total++ >= itmax) { info += I - ilo + 1; for (int k = ilo + c.tid; k <= I; k += c.nt) w[k] = H[k + (size_t)k * ldh]; break; }
    // number of shifts for this block
    int ns = sh.ns_max;
    if (m < 4 * ns) ns = m / 4;
    if (ns < 2) ns = 2;
    stagn += 1;
    int have_shifts = 0;
    if (sh.nw > 0 && stagn % 6 != 0) {
      // aggressive early deflation on the trailing window; while it keeps deflating more than NIBBLE per cent of the
      // window no sweep is spent (ZLAQR0's rule)
      const int nd = aed(c, sh, H, ldh, L, I, sh.nw, smlnum, w, ns, &have_shifts);
      HQR_PROF(15);
      if (nd > 0) {
        I -= nd; stagn = 0;
        const int m2 = I - L + 1;
        if (100 * nd > sh.nibble * sh.nw || m2 <= sh.W) continue;
        int ns2 = sh.ns_max;
        if (m2 < 4 * ns2) ns2 = m2 / 4;
        if (ns2 < 2) ns2 = 2;
        if (ns2 != ns) { ns = ns2; have_shifts = 0; }        // the chain shortens with the block: recompute the shifts
      }
    }
    if (have_shifts) {
      cta_sync();
    } else if (stagn % 6 == 0) {
      // exceptional shifts (ZLAQR0 style): diagonal entry plus 0.75 |subdiagonal|
      for (int b = c.tid; b < ns; b += c.nt) {
        const int k = I - b;
        sh.shifts[b] = H[k + (size_t)k * ldh] + mk(0.75 * cabs1(H[k + (size_t)(k - 1) * ldh]), 0.0);
      }
      cta_sync();
    } else {
      const int k0 = I - ns + 1;
      const int lds = sh.ns_max + 1;
      for (int q = c.tid; q < ns * ns; q += c.nt) {
        const int col = q / ns, row = q - col * ns;
        sh.sm[row + col * lds] = (row <= col + 1) ? H[(k0 + row) + (size_t)(k0 + col) * ldh] : mk(0.0, 0.0);
      }
      cta_sync();
      if (c.wid == 0) {                        // a 16x16 problem: one warp, warp-level barriers
        Grp gw; gw.tid = c.lane; gw.nt = c.ws; gw.warp = true;
        smem_hqr<false>(gw, sh.sm, lds, ns, 0, sh.shifts, sh.ctl, sh.cur, sh.rec);
      }
      cta_sync();
    }
    HQR_PROF(1);
    sweep_multishift(c, sh, H, ldh, L, I, ns);
  }
  cta_sync();
  return info;
}

}  // namespace stab
