// tables.cuh -- per-collocation-node coefficient tables of the linearised compressible
// Navier-Stokes operator
//     G q_t + A q_x + B q_y + C q_z + D q = Vxx q_xx + Vxy q_xy + Vyy q_yy + Vxz q_xz + Vyz q_yz + Vzz q_zz
// for q = (rho, u, v, w, T) about a parallel mean flow.  One call fills the eleven 5x5 real
// tables of ONE node; the tables are independent of (alpha, beta, omega).
//
// Reference behaviour restated (not copied): temporal.f90:206-598 (flat plate, temporal form of
// the energy equation) and spatial.f90:223-673 (curvature metrics h(y), u2m == 0, gamma-scaled
// energy equation); material law getmat.f90:12-34 and its nondimensionalisation
// temporal.f90:262-277.  Index convention: T[e*5+v] = equation e (0 continuity, 1-3 momentum,
// 4 energy) acting on variable v.
#pragma once
#include "common.cuh"

namespace stab {

struct Phys {            // mirrors the scalar part of module stuff (stuff.f90:11-59)
  double Ma, Re, Pr, gamma, gamma1, cp, Te, rmue, rlme, cone, datmat[3];
  int mattyp, navier;
};

struct NodeIn {          // mean flow at one node, y-derivatives already in physical space
  double rho, u1, u2, u3, T;
  double g2[5];          // d/dy   of (rho,u,v,w,T)   (temporal.f90:174-179)
  double g22[5];         // d2/dy2 of (rho,u,v,w,T)
  double h, dhds, dhdr, dhdsr, dhdrr;   // curvature metrics (spatial.f90:110-125); 1,0,0,0,0 if flat
};

struct Tables {
  double G[25], A[25], B[25], C[25], D[25], Vxx[25], Vxy[25], Vyy[25], Vxz[25], Vyz[25], Vzz[25];
};

struct Material { double mu, dmu, d2mu, lm, dlm, d2lm, con, dcon, d2con; };

// getmat.f90:12-34 evaluated at the dimensional temperature t, then nondimensionalised by the edge
// values (temporal.f90:267-277 / spatial.f90:312-322)
SD_HD Material material_at(double tm, const Phys& p) {
  const double pt66 = 6.6666666666666666666e-1;   // stuff.f90:35
  double t = tm * p.Te;
  double d1 = p.datmat[0], d2 = p.datmat[1], d3 = p.datmat[2];
  double mu, dmu, d2mu;
  if (p.mattyp == 0) {
    mu = d1; dmu = 0.0; d2mu = 0.0;
  } else {
    double sq = sqrt(t / d2);
    mu = d1 * t / d2 * sq * (d2 + d3) / (t + d3);
    dmu = (d1 * (3.0 * d3 + t) * (d3 + d2) * sq) / (2.0 * ((d3 + t) * (d3 + t)) * d2);
    d2mu = (d1 * (3.0 * (d3 * d3) - 6.0 * d3 * t - t * t) * (d3 + d2)) /
           (4.0 * ((d3 + t) * (d3 + t) * (d3 + t)) * sq * (d2 * d2));
  }
  double con = mu * p.cp / p.Pr, dcon = dmu * p.cp / p.Pr, d2con = d2mu * p.cp / p.Pr;
  double lm = -pt66 * mu, dlm = -pt66 * dmu, d2lm = -pt66 * d2mu;
  Material m;
  m.mu = mu / p.rmue; m.dmu = dmu * p.Te / p.rmue; m.d2mu = d2mu * (p.Te * p.Te) / p.rmue;
  m.con = con / p.cone; m.dcon = dcon * p.Te / p.cone; m.d2con = d2con * (p.Te * p.Te) / p.cone;
  m.lm = lm / p.rlme; m.dlm = dlm * p.Te / p.rlme; m.d2lm = d2lm * (p.Te * p.Te) / p.rlme;
  return m;
}

SD_HD void tables_clear(Tables& t) {
  double* q = t.G;
  for (int i = 0; i < 11 * 25; ++i) q[i] = 0.0;
}

#define TB(X, e, v) t.X[(e) * 5 + (v)]

// ---------------------------------------------------------------------------------------------
// temporal form (temporal.f90:206-598)
// ---------------------------------------------------------------------------------------------
SD_HD void node_tables_temporal(const NodeIn& q, const Phys& p, Tables& t) {
  tables_clear(t);
  const double rho = q.rho, tm = q.T;
  const double um[3] = {q.u1, q.u2, q.u3};
  const double gam = p.gamma, gam1 = p.gamma1, Ma = p.Ma, Re = p.Re, Pr = p.Pr;
  const double gm2 = gam * (Ma * Ma);
  // only d/dy of the mean is non-zero (parallel flow): gum[k][d] = d u_k / d x_d
  double gum[3][3] = {{0, q.g2[1], 0}, {0, q.g2[2], 0}, {0, q.g2[3], 0}};
  double grho[3] = {0, q.g2[0], 0}, gt[3] = {0, q.g2[4], 0};
  double divum = gum[0][0] + gum[1][1] + gum[2][2];
  double fact = 1.0 / gm2;
  double gp[3];
  for (int k = 0; k < 3; ++k) gp[k] = fact * (grho[k] * tm + rho * gt[k]);
  double gdiv[3] = {0.0, q.g22[2], 0.0};                         // temporal.f90:239-241
  double S[3][3];
  for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) S[a][b] = 0.5 * (gum[a][b] + gum[b][a]);
  double Sjj[3] = {0.5 * q.g22[1], 0.5 * (q.g22[2] + q.g22[2]), 0.5 * q.g22[3]};   // :251-258
  Material m = material_at(tm, p);
  double gmu[3], gdmu[3], gcon[3], gdcon[3], glm[3], gdlm[3];
  for (int k = 0; k < 3; ++k) {
    gmu[k] = m.dmu * gt[k];   gdmu[k] = m.d2mu * gt[k];
    gcon[k] = m.dcon * gt[k]; gdcon[k] = m.d2con * gt[k];
    glm[k] = m.dlm * gt[k];   gdlm[k] = m.d2lm * gt[k];
  }
  double* T3[3] = {t.A, t.B, t.C};                                // first-derivative table of direction d
  double* Vd[3][3] = {{t.Vxx, t.Vxy, t.Vxz}, {t.Vxy, t.Vyy, t.Vyz}, {t.Vxz, t.Vyz, t.Vzz}};

  // continuity
  TB(G, 0, 0) = 1.0;
  for (int d = 0; d < 3; ++d) { T3[d][0] = um[d]; T3[d][1 + d] = rho; }
  TB(D, 0, 0) = divum; TB(D, 0, 1) = grho[0]; TB(D, 0, 2) = grho[1]; TB(D, 0, 3) = grho[2];

  // momentum in direction k -> equation e = 1 + k
  for (int k = 0; k < 3; ++k) {
    const int e = 1 + k;
    TB(G, e, e) = rho;
    for (int d = 0; d < 3; ++d) T3[d][e * 5 + e] = rho * um[d];
    T3[k][e * 5 + 0] = tm / gm2;
    T3[k][e * 5 + 4] = rho / gm2;
    TB(D, e, 0) = um[0] * gum[k][0] + um[1] * gum[k][1] + um[2] * gum[k][2] + gt[k] / gm2;
    TB(D, e, 1) = rho * gum[k][0]; TB(D, e, 2) = rho * gum[k][1]; TB(D, e, 3) = rho * gum[k][2];
    TB(D, e, 4) = grho[k] / gm2;
    if (p.navier) {
      fact = p.rlme / (p.rmue * Re);                              // bulk viscosity terms
      for (int d = 0; d < 3; ++d) T3[d][e * 5 + 1 + d] -= fact * glm[k];
      T3[k][e * 5 + 4] -= fact * m.dlm * divum;
      TB(D, e, 4) -= fact * (gdlm[k] * divum + m.dlm * gdiv[k]);
      for (int d = 0; d < 3; ++d) Vd[k][d][e * 5 + 1 + d] = fact * m.lm;
      fact = 1.0 / Re;                                            // shear viscosity terms
      for (int d = 0; d < 3; ++d) {
        if (d == k) {
          T3[k][e * 5 + e] -= fact * 2.0 * gmu[k];
        } else {
          T3[d][e * 5 + e] -= fact * gmu[d];
          T3[k][e * 5 + 1 + d] -= fact * gmu[d];
        }
        T3[d][e * 5 + 4] -= fact * m.dmu * 2.0 * S[k][d];
      }
      TB(D, e, 4) -= fact * 2.0 * (gdmu[0] * S[k][0] + gdmu[1] * S[k][1] + gdmu[2] * S[k][2] + m.dmu * Sjj[k]);
      for (int d = 0; d < 3; ++d) {
        if (d == k) {
          Vd[k][k][e * 5 + e] += fact * 2.0 * m.mu;
        } else {
          Vd[d][d][e * 5 + e] += fact * m.mu;
          Vd[k][d][e * 5 + 1 + d] += fact * m.mu;
        }
      }
    }
  }

  // energy (temporal.f90:530-598)
  TB(G, 4, 0) = -gam1 * tm / gam;
  TB(G, 4, 4) = rho / gam;
  for (int d = 0; d < 3; ++d) {
    T3[d][4 * 5 + 0] = -gam1 * um[d] * tm / gam;
    T3[d][4 * 5 + 4] = rho * um[d] / gam;
  }
  TB(D, 4, 0) = 1.0 / gam * (um[0] * gt[0] + um[1] * gt[1] + um[2] * gt[2]);
  for (int d = 0; d < 3; ++d) TB(D, 4, 1 + d) = rho * gt[d] - gam1 * (Ma * Ma) * gp[d];
  TB(D, 4, 4) = -gam1 / gam * (um[0] * grho[0] + um[1] * grho[1] + um[2] * grho[2]);
  if (p.navier) {
    fact = 1.0 / (Pr * Re);
    for (int d = 0; d < 3; ++d) T3[d][4 * 5 + 4] -= fact * (gcon[d] + m.dcon * gt[d]);
    TB(D, 4, 4) -= fact * (gdcon[0] * gt[0] + gdcon[1] * gt[1] + gdcon[2] * gt[2] + m.dcon * (0.0 + q.g22[4] + 0.0));
    TB(Vxx, 4, 4) = fact * m.con; TB(Vyy, 4, 4) = fact * m.con; TB(Vzz, 4, 4) = fact * m.con;
    fact = gam1 * (Ma * Ma) * p.rlme / (Re * p.rmue);
    for (int d = 0; d < 3; ++d) T3[d][4 * 5 + 1 + d] -= fact * 2.0 * m.lm * divum;
    TB(D, 4, 4) -= fact * m.dlm * divum * divum;
    fact = 2.0 * gam1 * (Ma * Ma) / Re;
    for (int d = 0; d < 3; ++d)
      for (int k = 0; k < 3; ++k) T3[d][4 * 5 + 1 + k] -= fact * 2.0 * m.mu * S[k][d];
    double ss = 0.0;
    for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) ss += S[a][b] * S[a][b];
    TB(D, 4, 4) -= fact * m.dmu * ss;
  }
}

// ---------------------------------------------------------------------------------------------
// spatial form with curvature metrics (spatial.f90:223-673); u2m == 0 (spatial.f90:149)
// ---------------------------------------------------------------------------------------------
SD_HD void node_tables_spatial(const NodeIn& q, const Phys& p, Tables& t) {
  tables_clear(t);
  const double rho = q.rho, u1 = q.u1, u2 = 0.0, u3 = q.u3, tm = q.T;
  const double h = q.h, dhds = q.dhds, dhdr = q.dhdr, dhdsr = q.dhdsr, dhdrr = q.dhdrr;
  const double h2 = h * h, h3 = h * h * h;
  const double gam = p.gamma, gam1 = p.gamma1, Ma = p.Ma, Re = p.Re, Pr = p.Pr;
  const double gm2 = gam * (Ma * Ma);
  double gum[3][3] = {{0, q.g2[1], 0}, {0, q.g2[2], 0}, {0, q.g2[3], 0}};
  double grho[3] = {0, q.g2[0], 0}, gt[3] = {0, q.g2[4], 0};
  const double* g22 = q.g22;
  double divum = (gum[0][0] + u2 * dhdr) / h + gum[1][1] + gum[2][2];
  double g1div = -dhds / h3 * (gum[0][0] + u2 * dhdr) + 1.0 / h2 * (0.0 + gum[1][0] * dhdr + u2 * dhdsr) + 1.0 / h * (0.0 + 0.0);
  double g2div = -dhdr / h2 * (gum[0][0] + u2 * dhdr) + 1.0 / h * (0.0 + gum[1][1] * dhdr + u2 * dhdrr) + (g22[2] + 0.0);
  double g3div = 1.0 / h * (0.0 + gum[1][2] * dhdr) + 0.0 + 0.0;
  double S[3][3];
  S[0][0] = (gum[0][0] + u2 * dhdr) / h;
  S[0][1] = 0.5 * ((gum[1][0] - u1 * dhdr) / h + gum[0][1]);
  S[0][2] = 0.5 * (gum[2][0] / h + gum[0][2]);
  S[1][0] = S[0][1];
  S[1][1] = gum[1][1];
  S[1][2] = 0.5 * (gum[2][1] + gum[1][2]);
  S[2][0] = S[0][2];
  S[2][1] = S[1][2];
  S[2][2] = gum[2][2];
  double S1jj = -0.5 * (dhdr * dhdr + dhdrr * h) / h2 * u1 + 0.5 * dhdr * gum[0][1] / h + 0.5 * g22[1] -
                dhds * gum[0][0] / h3 + 0.0 / h2 + 0.5 * 0.0 + (h * dhdsr - dhdr * dhds) / h3 * u2 +
                3.0 * dhdr * gum[1][0] / (2.0 * h2) + 0.5 * 0.0 / h + 0.5 * 0.0 / h;
  double S2jj = 0.5 * (dhdr * dhds - h * dhdsr) / h3 * u1 - 3.0 * dhdr * gum[0][0] / (2.0 * h2) + 0.5 * 0.0 / h -
                dhdr * dhdr * u2 / h2 + dhdr * gum[1][1] / h + g22[2] - 0.5 * dhds * gum[1][0] / h3 +
                0.5 * 0.0 / h2 + 0.5 * 0.0 + 0.5 * 0.0;
  double S3jj = 0.5 * 0.0 / h + 0.5 * 0.0 + 0.5 * dhdr * gum[1][2] / h + 0.5 * dhdr * gum[2][1] / h + 0.5 * g22[3] +
                0.5 * 0.0 / h2 + 0.0 - 0.5 * dhds * gum[2][0] / h3;
  double LapT = 1.0 / h * (-dhds / h2 * gt[0] + 1.0 / h * 0.0 + h * g22[4] + gt[1] * dhdr + h * 0.0);
  Material m = material_at(tm, p);
  const double mu = m.mu, dmu = m.dmu, lm = m.lm, dlm = m.dlm, con = m.con, dcon = m.dcon;
  double g1mu = dmu * gt[0], g2mu = dmu * gt[1], g3mu = dmu * gt[2];
  double g1dmu = m.d2mu * gt[0], g2dmu = m.d2mu * gt[1], g3dmu = m.d2mu * gt[2];
  double g1con = dcon * gt[0], g2con = dcon * gt[1], g3con = dcon * gt[2];
  double g1dcon = m.d2con * gt[0], g2dcon = m.d2con * gt[1], g3dcon = m.d2con * gt[2];
  double g1lm = dlm * gt[0], g2lm = dlm * gt[1], g3lm = dlm * gt[2];
  double g1dlm = m.d2lm * gt[0], g2dlm = m.d2lm * gt[1], g3dlm = m.d2lm * gt[2];
  double fact;

  // continuity
  TB(G, 0, 0) = 1.0;
  TB(A, 0, 0) = u1 / h; TB(A, 0, 1) = rho / h;
  TB(B, 0, 0) = u2; TB(B, 0, 2) = rho;
  TB(C, 0, 0) = u3; TB(C, 0, 3) = rho;
  TB(D, 0, 0) = divum; TB(D, 0, 1) = grho[0] / h; TB(D, 0, 2) = grho[1] + rho * dhdr / h; TB(D, 0, 3) = grho[2];
  // x1 momentum
  TB(G, 1, 1) = rho;
  TB(A, 1, 0) = tm / (h * gm2); TB(A, 1, 1) = rho * u1 / h; TB(A, 1, 4) = rho / (h * gm2);
  TB(B, 1, 1) = rho * u2;
  TB(C, 1, 1) = rho * u3;
  TB(D, 1, 0) = u1 / h * (gum[0][0] + u2 * dhdr) + u2 * gum[0][1] + u3 * gum[0][2] + gt[0] / (h * gm2);
  TB(D, 1, 1) = rho * (gum[0][0] + u2 * dhdr) / h;
  TB(D, 1, 2) = rho * (gum[0][1] + u1 * dhdr / h);
  TB(D, 1, 3) = rho * gum[0][2];
  TB(D, 1, 4) = grho[0] / (h * gm2);
  if (p.navier) {
    fact = p.rlme / (p.rmue * Re);
    TB(A, 1, 1) -= fact * (g1lm / h2 - lm / h3 * dhds);
    TB(A, 1, 2) -= fact * lm / h2 * dhdr;
    TB(A, 1, 4) -= fact * dlm * divum / h;
    TB(B, 1, 2) -= fact * (g1lm / h);
    TB(C, 1, 3) -= fact * (g1lm / h);
    TB(D, 1, 2) -= fact * (g1lm * dhdr / h2 - lm / h3 * dhds * dhdr + lm / h2 * dhdsr);
    TB(D, 1, 4) -= fact * (g1dlm * divum / h + dlm * g1div);
    TB(Vxx, 1, 1) = fact * lm / h2;
    TB(Vxy, 1, 2) = fact * lm / h;
    TB(Vxz, 1, 3) = fact * lm / h;
    fact = 1.0 / Re;
    TB(A, 1, 1) -= fact * (2.0 * g1mu / h2 - 2.0 * mu * dhds / h3);
    TB(A, 1, 2) -= fact * (g2mu / h + mu * 3.0 * dhdr / h2);
    TB(A, 1, 3) -= fact * g3mu / h;
    TB(A, 1, 4) -= fact * dmu * 2.0 * S[0][0] / h;
    TB(B, 1, 1) -= fact * (g2mu + mu * dhdr / h);
    TB(B, 1, 4) -= fact * dmu * 2.0 * S[0][1];
    TB(C, 1, 1) -= fact * g3mu;
    TB(C, 1, 4) -= fact * dmu * 2.0 * S[0][2];
    TB(D, 1, 1) -= fact * (g2mu / h * (-dhdr) - mu * (dhdr * dhdr + dhdrr * h) / h2);
    TB(D, 1, 2) -= fact * (2.0 * g1mu / h2 * dhdr + 2.0 * mu * (dhdsr * h - dhds * dhdr) / h3);
    TB(D, 1, 4) -= fact * 2.0 * (g1dmu / h * S[0][0] + g2dmu * S[0][1] + g3dmu * S[0][2] + dmu * S1jj);
    TB(Vxx, 1, 1) += fact * 2.0 * mu / h2;
    TB(Vxy, 1, 2) += fact * mu / h;
    TB(Vyy, 1, 1) += fact * mu;
    TB(Vxz, 1, 3) += fact * mu / h;
    TB(Vzz, 1, 1) += fact * mu;
  }
  // x2 momentum
  TB(G, 2, 2) = rho;
  TB(A, 2, 2) = rho * u1 / h;
  TB(B, 2, 0) = tm / gm2; TB(B, 2, 2) = rho * u2; TB(B, 2, 4) = rho / gm2;
  TB(C, 2, 2) = rho * u3;
  TB(D, 2, 0) = u1 / h * (gum[1][0] - u1 * dhdr) + u2 * gum[1][1] + u3 * gum[1][2] + gt[1] / gm2;
  TB(D, 2, 1) = rho * (gum[1][0] - 2.0 * u1 * dhdr) / h;
  TB(D, 2, 2) = rho * gum[1][1];
  TB(D, 2, 3) = rho * gum[1][2];
  TB(D, 2, 4) = grho[1] / gm2;
  if (p.navier) {
    fact = p.rlme / (p.rmue * Re);
    TB(A, 2, 1) -= fact * (g2lm / h - lm * dhdr / h2);
    TB(B, 2, 2) -= fact * (g2lm + lm * dhdr / h);
    TB(B, 2, 4) -= fact * dlm * divum;
    TB(C, 2, 3) -= fact * g2lm;
    TB(D, 2, 2) -= fact * (g2lm / h * dhdr - lm * dhdr / h2 * dhdr + lm / h * dhdrr);
    TB(D, 2, 4) -= fact * (g2dlm * divum + dlm * g2div);
    TB(Vxy, 2, 1) = fact * lm / h;
    TB(Vyy, 2, 2) = fact * lm;
    TB(Vyz, 2, 3) = fact * lm;
    fact = 1.0 / Re;
    TB(A, 2, 1) += fact * mu * 3.0 * dhdr / h2;
    TB(A, 2, 2) -= fact * (g1mu / h2 - mu * dhds / h3);
    TB(A, 2, 4) -= fact * dmu * 2.0 * S[1][0] / h;
    TB(B, 2, 1) -= fact * g1mu / h;
    TB(B, 2, 2) -= fact * (2.0 * g2mu + 2.0 * mu * dhdr / h);
    TB(B, 2, 3) -= fact * g3mu;
    TB(B, 2, 4) -= fact * dmu * 2.0 * S[1][1];
    TB(C, 2, 2) -= fact * g3mu;
    TB(C, 2, 4) -= fact * dmu * 2.0 * S[1][2];
    TB(D, 2, 1) -= fact * (g1mu / h2 * (-dhdr) + mu * (dhds * dhdr - h * dhdsr) / h3);
    TB(D, 2, 2) += fact * 2.0 * mu * (dhdr * dhdr) / h2;
    TB(D, 2, 4) -= fact * 2.0 * (g1dmu / h * S[1][0] + g2dmu * S[1][1] + g3dmu * S[1][2] + dmu * S2jj);
    TB(Vxx, 2, 2) += fact * mu / h2;
    TB(Vxy, 2, 1) += fact * mu / h;
    TB(Vyy, 2, 2) += fact * 2.0 * mu;
    TB(Vyz, 2, 3) += fact * mu;
    TB(Vzz, 2, 2) += fact * mu;
  }
  // x3 momentum
  TB(G, 3, 3) = rho;
  TB(A, 3, 3) = rho * u1 / h;
  TB(B, 3, 3) = rho * u2;
  TB(C, 3, 0) = tm / gm2; TB(C, 3, 3) = rho * u3; TB(C, 3, 4) = rho / gm2;
  TB(D, 3, 0) = u1 * gum[2][0] / h + u2 * gum[2][1] + u3 * gum[2][2] + gt[2] / gm2;
  TB(D, 3, 1) = rho * gum[2][0] / h;
  TB(D, 3, 2) = rho * gum[2][1];
  TB(D, 3, 3) = rho * gum[2][2];
  TB(D, 3, 4) = grho[2] / gm2;
  if (p.navier) {
    fact = p.rlme / (p.rmue * Re);
    TB(A, 3, 1) -= fact * g3lm / h;
    TB(B, 3, 2) -= fact * g3lm;
    TB(C, 3, 3) -= fact * g3lm;
    TB(C, 3, 2) -= fact * lm / h * dhdr;
    TB(C, 3, 4) -= fact * dlm * divum;
    TB(D, 3, 2) -= fact * (g2lm / h * dhdr);
    TB(D, 3, 4) -= fact * (g3dlm * divum + dlm * g3div);
    TB(Vxz, 3, 1) = fact * lm / h;
    TB(Vyz, 3, 2) = fact * lm;
    TB(Vzz, 3, 3) = fact * lm;
    fact = 1.0 / Re;
    TB(A, 3, 3) -= fact * (g1mu / h2 - mu * dhds / h3);
    TB(A, 3, 4) -= fact * dmu * 2.0 * S[2][0] / h;
    TB(B, 3, 3) -= fact * (g2mu + mu * dhdr / h);
    TB(B, 3, 4) -= fact * dmu * 2.0 * S[2][1];
    TB(C, 3, 1) -= fact * g1mu / h;
    TB(C, 3, 2) -= fact * (g2mu + mu * dhdr / h);
    TB(C, 3, 3) -= fact * 2.0 * g3mu;
    TB(C, 3, 4) -= fact * dmu * 2.0 * S[2][2];
    TB(D, 3, 4) -= fact * 2.0 * (g1dmu / h * S[2][0] + g2dmu * S[2][1] + g3dmu * S[2][2] + dmu * S3jj);
    TB(Vxx, 3, 3) += fact * mu / h2;
    TB(Vyy, 3, 3) += fact * mu;
    TB(Vxz, 3, 1) += fact * mu / h;
    TB(Vyz, 3, 2) += fact * mu;
    TB(Vzz, 3, 3) += fact * 2.0 * mu;
  }
  // energy
  TB(G, 4, 4) = rho;
  TB(A, 4, 1) = rho * gam1 * tm / h; TB(A, 4, 4) = rho * u1 / h;
  TB(B, 4, 2) = rho * gam1 * tm; TB(B, 4, 4) = rho * u2;
  TB(C, 4, 3) = rho * gam1 * tm; TB(C, 4, 4) = rho * u3;
  TB(D, 4, 0) = u1 / h * gt[0] + u2 * gt[1] + u3 * gt[2] + gam1 * tm * divum;
  TB(D, 4, 1) = rho * gt[0] / h;
  TB(D, 4, 2) = rho * gt[1] + rho * gam1 * tm * dhdr / h;
  TB(D, 4, 3) = rho * gt[2];
  TB(D, 4, 4) = rho * gam1 * divum;
  if (p.navier) {
    fact = gam / (Pr * Re);
    TB(A, 4, 4) -= fact * (g1con / h2 + dcon * gt[0] / h2 - con * dhds / h3);
    TB(B, 4, 4) -= fact * (g2con + dcon * gt[1] + con * dhdr / h);
    TB(C, 4, 4) -= fact * (g3con + dcon * gt[2]);
    TB(D, 4, 4) -= fact * (g1dcon * gt[0] / h2 + g2dcon * gt[1] + g3dcon * gt[2] + dcon * LapT);
    TB(Vxx, 4, 4) = fact * con / h2;
    TB(Vyy, 4, 4) = fact * con;
    TB(Vzz, 4, 4) = fact * con;
    fact = gam * gam1 * (Ma * Ma) * p.rlme / (Re * p.rmue);
    TB(A, 4, 1) -= fact * 2.0 * lm * divum / h;
    TB(B, 4, 2) -= fact * 2.0 * lm * divum;
    TB(C, 4, 3) -= fact * 2.0 * lm * divum;
    TB(D, 4, 2) -= fact * 2.0 * lm * divum * dhdr / h;
    TB(D, 4, 4) -= fact * dlm * divum * divum;
    fact = gam * gam1 * (Ma * Ma) / Re;
    for (int k = 0; k < 3; ++k) {
      TB(A, 4, 1 + k) -= fact * 4.0 * mu * S[k][0] / h;
      TB(B, 4, 1 + k) -= fact * 4.0 * mu * S[k][1];
      TB(C, 4, 1 + k) -= fact * 4.0 * mu * S[k][2];
    }
    TB(D, 4, 1) += fact * 4.0 * mu * S[1][0] * dhdr / h;
    TB(D, 4, 2) -= fact * 4.0 * mu * S[0][0] * dhdr / h;
    double ss = 0.0;
    for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) ss += S[a][b] * S[a][b];
    TB(D, 4, 4) -= fact * 2.0 * dmu * ss;
  }
}
#undef TB

}  // namespace stab
