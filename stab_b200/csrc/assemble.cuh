// assemble.cuh -- operator assembly for one sweep point.
//
// Stage 1a (node coefficients): combine the 11 real tables of a node with the point's
// (alpha, beta) [temporal] or (omega, beta) [spatial] and the grid metrics into complex 5x5
// coefficient blocks c1, c2, c0 such that every dense operator row-block is
//       Op[(i,e),(j,v)] = c1_i[e,v] * D1[i,j] + c2_i[e,v] * D2*[i,j] + delta_ij * c0_i[e,v]
// where D2* is the second-derivative Chebyshev matrix, except at the wall node for the
// temperature column (v = 4) when wallt = 2, where the adiabatic operator row Dt2 replaces it.
// Boundary rows (Dirichlet rows, first-order continuity rows, wall energy row) are encoded in
// the coefficients, so stage 1b is one uniform streaming kernel.
//
// Temporal: the reference forms A0 and the block-diagonal B0 = i*G (temporal.f90:630-752) and
// calls ZGESV to get B0^-1 A0 (temporal.f90:774).  B0's 5x5 diagonal blocks are "diagonal plus
// one entry (5,1)", so the LU solve collapses into a per-node 5x5 forward/back substitution that
// we apply to the coefficients: the kernel writes M = B0^-1 A0 directly.
// Spatial: C0, C1, C2 of spatial.f90:687-959 are written side by side as the augmented matrix
// [C0 | -C1 | -C2] (signs of spatial.f90:982-983) that the LU stage consumes.
#pragma once
#include "tables.cuh"

namespace stab {

struct PointTemporal { cplx alpha, beta; };
struct PointSpatial { cplx omega, beta; };

// c arrays: 25 complex each, index e*5+v
struct NodeCoef3 { cplx c1[25], c2[25], c0[25]; };

// Solve (i*Gb) X = Cin for one 5x5 block where Gb = diag(g) + g40 at (4,0): what ZGESV's LU does
// on this block (no pivoting needed while |g40| <= |g00|; the result is the same either way).
SD_HD void apply_b0_inverse(const double g[5], double g40, cplx* c /*25, in place*/) {
  for (int v = 0; v < 5; ++v) {
    cplx y4 = c[4 * 5 + v] - g40 * c[0 * 5 + v];
    c[4 * 5 + v] = y4;
  }
  for (int e = 0; e < 5; ++e)
    for (int v = 0; v < 5; ++v) {
      cplx x = c[e * 5 + v];                       // x / (i*g) = (x.im/g, -x.re/g)
      c[e * 5 + v] = mk(x.im / g[e], -x.re / g[e]);
    }
}

// temporal.f90:604-618 + row selection of :630-752, then B0^-1
SD_HD void node_coef_temporal(const Tables& t, int i, int ny, int wallt, double deta, double d2eta,
                              PointTemporal pt, bool apply_b0inv, NodeCoef3& o) {
  const cplx im = mk(0.0, 1.0);
  const cplx al = pt.alpha, be = pt.beta;
  const cplx al2 = al * al, albe = al * be, be2 = be * be;
  const cplx ial = im * al, ibe = im * be;
  cplx Dh[25], Bh[25];
  double Vy[25];
  for (int k = 0; k < 25; ++k) {
    Dh[k] = mk(t.D[k], 0.0) + ial * t.A[k] + ibe * t.C[k] + al2 * t.Vxx[k] + albe * t.Vxz[k] + be2 * t.Vzz[k];
    cplx b = mk(t.B[k], 0.0) - ial * t.Vxy[k] - ibe * t.Vyz[k];
    Bh[k] = b * deta - mk(t.Vyy[k] * d2eta, 0.0);
    Vy[k] = t.Vyy[k] * (deta * deta);
  }
  const cplx zero = mk(0.0, 0.0);
  for (int k = 0; k < 25; ++k) { o.c1[k] = zero; o.c2[k] = zero; o.c0[k] = zero; }
  const bool top = (i == 0), wall = (i == ny - 1);
  if (!top && !wall) {
    for (int k = 0; k < 25; ++k) { o.c1[k] = Bh[k]; o.c2[k] = mk(-Vy[k], 0.0); o.c0[k] = Dh[k]; }
  } else {
    for (int v = 0; v < 5; ++v) { o.c1[v] = Bh[v]; o.c0[v] = Dh[v]; }         // continuity, first-derivative form
    if (wall && wallt == 2) {
      for (int v = 0; v < 4; ++v) {                                            // quirk q4: D1 multiplies Vyy here
        o.c1[20 + v] = Bh[20 + v] - mk(Vy[20 + v], 0.0);
        o.c0[20 + v] = Dh[20 + v];
      }
      o.c2[24] = mk(-Vy[24], 0.0);                                             // * Dt2 wall row; Dt1 wall row == 0
      o.c0[24] = Dh[24];
    }
  }
  if (apply_b0inv) {
    double g[5] = {1.0, 1.0, 1.0, 1.0, 1.0};
    double g40 = 0.0;
    if (!top && !wall) {
      for (int e = 0; e < 5; ++e) g[e] = t.G[e * 5 + e];
      g40 = t.G[20];
    } else if (wall && wallt == 2) {
      g[4] = t.G[24];
      g40 = t.G[20];
    }
    apply_b0_inverse(g, g40, o.c1);
    apply_b0_inverse(g, g40, o.c2);
    apply_b0_inverse(g, g40, o.c0);
  }
}

// B0 5x5 diagonal block of node i (temporal.f90:630-662), for the inspection entry point
SD_HD void node_b0_temporal(const Tables& t, int i, int ny, int wallt, cplx* b /*25*/) {
  for (int k = 0; k < 25; ++k) b[k] = mk(0.0, 0.0);
  const bool top = (i == 0), wall = (i == ny - 1);
  if (!top && !wall) {
    for (int k = 0; k < 25; ++k) b[k] = mk(0.0, t.G[k]);
  } else {
    for (int e = 0; e < 5; ++e) b[e * 5 + e] = mk(0.0, 1.0);
    if (wall && wallt == 2)
      for (int v = 0; v < 5; ++v) b[20 + v] = mk(0.0, t.G[20 + v]);
  }
}

struct NodeCoefSpatial { NodeCoef3 C0; cplx C1c1[25], C1c0[25], C2c0[25]; };

// spatial.f90:687-959; C1 and C2 are stored negated (spatial.f90:982-983)
SD_HD void node_coef_spatial(const Tables& t, int i, int ny, int wallt, int topflag, double deta, double d2eta,
                             PointSpatial pt, NodeCoefSpatial& o) {
  const cplx im = mk(0.0, 1.0), zero = mk(0.0, 0.0);
  const cplx be = pt.beta, om = pt.omega;
  const cplx ibe = im * be, be2 = be * be, iom = im * om;
  cplx Dh[25], Bh[25], Dh1[25], Bh1[25];
  double Vy[25];
  for (int k = 0; k < 25; ++k) {
    Dh[k] = mk(t.D[k], 0.0) + ibe * t.C[k] + be2 * t.Vzz[k];
    cplx b = mk(t.B[k], 0.0) - ibe * t.Vyz[k];
    Bh[k] = b * deta - mk(t.Vyy[k] * d2eta, 0.0);
    Vy[k] = t.Vyy[k] * (deta * deta);
    Dh1[k] = im * t.A[k] + be * t.Vxz[k];
    Bh1[k] = (mk(0.0, -1.0) * t.Vxy[k]) * deta;
  }
  for (int k = 0; k < 25; ++k) {
    o.C0.c1[k] = zero; o.C0.c2[k] = zero; o.C0.c0[k] = zero;
    o.C1c1[k] = zero; o.C1c0[k] = zero; o.C2c0[k] = zero;
  }
  const bool top = (i == 0), wall = (i == ny - 1);
  if (!top && !wall) {
    for (int k = 0; k < 25; ++k) {
      o.C0.c1[k] = Bh[k]; o.C0.c2[k] = mk(-Vy[k], 0.0); o.C0.c0[k] = Dh[k] - iom * t.G[k];
      o.C1c1[k] = -Bh1[k]; o.C1c0[k] = -Dh1[k];
      o.C2c0[k] = mk(-t.Vxx[k], 0.0);
    }
    return;
  }
  if (top) {
    if (topflag == 1) {
      for (int v = 0; v < 5; ++v) { o.C0.c1[v] = Bh[v]; o.C0.c0[v] = Dh[v]; o.C1c1[v] = -Bh1[v]; o.C1c0[v] = -Dh1[v]; }
      o.C0.c0[0] -= iom;
    } else {
      o.C0.c0[0] = mk(-1.0, 0.0);
    }
    for (int e = 1; e < 5; ++e) o.C0.c0[e * 5 + e] = mk(-1.0, 0.0);
    return;
  }
  // wall
  for (int v = 0; v < 5; ++v) { o.C0.c1[v] = Bh[v]; o.C0.c0[v] = Dh[v]; o.C1c1[v] = -Bh1[v]; o.C1c0[v] = -Dh1[v]; }
  o.C0.c0[0] -= iom;
  for (int e = 1; e < 4; ++e) o.C0.c0[e * 5 + e] = mk(-1.0, 0.0);
  if (wallt == 0) {
    o.C0.c0[24] = mk(-1.0, 0.0);
  } else {  // wallt == 2: energy equation with the adiabatic operators Dt1 (zero wall row) / Dt2
    for (int v = 0; v < 4; ++v) {
      o.C0.c1[20 + v] = Bh[20 + v]; o.C0.c2[20 + v] = mk(-Vy[20 + v], 0.0); o.C0.c0[20 + v] = Dh[20 + v] - iom * t.G[20 + v];
      o.C1c1[20 + v] = -Bh1[20 + v]; o.C1c0[20 + v] = -Dh1[20 + v];
    }
    o.C0.c2[24] = mk(-Vy[24], 0.0);
    o.C0.c0[24] = Dh[24] - iom * t.G[24];
    o.C1c0[24] = -Dh1[24];
    for (int v = 0; v < 5; ++v) o.C2c0[20 + v] = mk(-t.Vxx[20 + v], 0.0);
  }
}

// Device-resident per-batch inputs shared by all points
struct GridDev {
  int ny;
  int wallt, top;
  const double* D1;      // ny x ny column-major: D1[i + j*ny]
  const double* D2;
  const double* Dt2w;    // ny: wall row of Dt2 (wallt=2), else == D2 wall row
  const double* deta;
  const double* d2eta;
  const double* vm;      // ny x 5 column-major (node, dof)
  const double* g2vm;    // ny x 5, already mapped to y-space
  const double* g22vm;
  const double* h5;      // ny x 5 (h, dhds, dhdr, dhdsr, dhdrr) or nullptr for flat
};

SD_HD NodeIn load_node(const GridDev& g, int i) {
  NodeIn q;
  const int ny = g.ny;
  q.rho = g.vm[i]; q.u1 = g.vm[i + ny]; q.u2 = g.vm[i + 2 * ny]; q.u3 = g.vm[i + 3 * ny]; q.T = g.vm[i + 4 * ny];
  for (int k = 0; k < 5; ++k) { q.g2[k] = g.g2vm[i + k * ny]; q.g22[k] = g.g22vm[i + k * ny]; }
  if (g.h5) {
    q.h = g.h5[i]; q.dhds = g.h5[i + ny]; q.dhdr = g.h5[i + 2 * ny]; q.dhdsr = g.h5[i + 3 * ny]; q.dhdrr = g.h5[i + 4 * ny];
  } else {
    q.h = 1.0; q.dhds = 0.0; q.dhdr = 0.0; q.dhdsr = 0.0; q.dhdrr = 0.0;
  }
  return q;
}

// One dense element from a coefficient triple.  `r` = row (i*5+e), `c` = column (j*5+v).
SD_HD cplx op_element(const cplx* c1, const cplx* c2, const cplx* c0, const GridDev& g, int i, int e, int j, int v) {
  const int ny = g.ny;
  const int k = e * 5 + v;
  double d1 = g.D1[i + j * ny];
  double d2 = (g.wallt == 2 && i == ny - 1 && v == 4) ? g.Dt2w[j] : g.D2[i + j * ny];
  cplx a = c1[k] * d1 + c2[k] * d2;
  if (i == j) a += c0[k];
  return a;
}

}  // namespace stab
