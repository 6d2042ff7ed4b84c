// host_math.cpp -- host-side pieces of the stab hot path: everything that is O(ny) or O(ny^2),
// runs once per profile, and whose rounding feeds every operator entry (grid, spline, Chebyshev
// matrix, curvature metrics, edge properties), plus sweep enumeration and the on-disk formats.
// Compiled with -ffp-contract=off so the operation order written here is the operation order
// executed (the reference is built without FMA contraction assumptions, gcc.mak:9-13).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <string>
#include <vector>
#include "../../include/stabgpu.h"

extern "C" {

void stabgpu_params_default(stabgpu_params* p) {
  std::memset(p, 0, sizeof(*p));
  p->ny = 1; p->ider = 1; p->ievec = 1;
  p->gamma = 1.4; p->gamma1 = 0.4; p->cp = 1003.1;      // stuff.f90:45
  p->Pr = 1.0;
  p->Te = 1.0; p->rmue = 1.0; p->rlme = 1.0; p->cone = 1.0;
  p->datmat[0] = 1.0; p->datmat[1] = 0.0; p->datmat[2] = 0.0;
}

// getmat.f90:40-75 (scalar flavour) evaluated at the edge temperature
static void sgetmat(const stabgpu_params* p, double t, double* mu, double* lm, double* con) {
  const double pt66 = 6.6666666666666666666e-1;
  double d1 = p->datmat[0], d2 = p->datmat[1], d3 = p->datmat[2];
  double rmu;
  if (p->mattyp == 0) rmu = d1;
  else rmu = d1 * t / d2 * std::sqrt(t / d2) * (d2 + d3) / (t + d3);
  *mu = rmu; *con = rmu * p->cp / p->Pr; *lm = -pt66 * rmu;
}

int stabgpu_edge_properties(stabgpu_params* p, double T0) {
  if (p->mattyp == 1) {
    // input.f90:25 evaluates Te with Ma not yet read (module default 0): Te = T0
    p->Te = T0;
    p->datmat[0] = 1.715336725523065e-05; p->datmat[1] = 273.0; p->datmat[2] = 110.4;
  } else {
    p->Te = 1.0;                                       // never set by the reference; multiplies zeros
    p->datmat[0] = 1.0; p->datmat[1] = 0.0; p->datmat[2] = 0.0;
  }
  sgetmat(p, p->Te, &p->rmue, &p->rlme, &p->cone);     // input.f90:43
  return 0;
}

int stabgpu_sgengrid(int ny, double yi, double ymax, double* y, double* eta, double* deta, double* d2eta) {
  if (ny < 2 || yi == 0.0) return 1;                   // tanh map (Yi = 0) reads stdin per call: unsupported
  const double pi = 3.1415926535897932385;             // stuff.f90:40
  const double dth = pi / (double)(ny - 1);
  for (int i = 0; i < ny; ++i) eta[i] = std::cos((double)i * dth);
  if (ymax == 0.0) {                                   // algebraic semi-infinite map, sgengrid.f90:28-38
    const double L = yi;
    for (int i = 0; i < ny; ++i) {
      deta[i] = ((eta[i] - 1.0) * (eta[i] - 1.0)) / (2.0 * L);
      d2eta[i] = ((eta[i] - 1.0) * (eta[i] - 1.0) * (eta[i] - 1.0)) / (2.0 * (L * L));
      y[i] = (eta[i] != 1.0) ? L * (1.0 + eta[i]) / (1.0 - eta[i]) : 1.0e99;
    }
  } else {                                             // Streett's finite map, sgengrid.f90:39-45
    for (int i = 0; i < ny; ++i) {
      double q = 2.0 * yi + 1.0 - eta[i];
      y[i] = ymax * yi * (1.0 + eta[i]) / (1.0 + 2.0 * yi - eta[i]);
      deta[i] = (q * q) / (2.0 * ymax * yi * (yi + 1.0));
      double d = ymax * yi * (yi + 1.0);
      d2eta[i] = -0.5 * (q * q * q) / (d * d);
    }
  }
  return 0;
}

int stabgpu_chebyd(int N, double* D) {
  const int m = N + 1;
  const double pi = std::acos(-1.0);
  std::vector<double> x(m), a(m);
  for (int j = 0; j < m; ++j) x[j] = std::cos(pi * (double)j / (double)N);
  for (int j = 0; j < m; ++j) {
    double aj = 1.0, dj = 0.0;
    for (int k = 0; k < m; ++k)
      if (k != j) aj = aj * (x[j] - x[k]);
    for (int k = 0; k < m; ++k)
      if (k != j) dj = dj + 1.0 / (x[j] - x[k]);
    a[j] = aj;
    D[j + (size_t)j * m] = dj;
  }
  for (int j = 0; j < m; ++j)
    for (int k = 0; k < m; ++k)
      if (k != j) D[j + (size_t)k * m] = a[j] / (a[k] * (x[j] - x[k]));
  return 0;
}

int stabgpu_spline(int n, const double* x, const double* y, double* fdp) {
  if (n < 4) return 1;
  std::vector<double> a(n, 0.0), b(n, 0.0), c(n, 0.0), r(n, 0.0);
  c[0] = x[1] - x[0];
  for (int i = 1; i < n - 1; ++i) {
    c[i] = x[i + 1] - x[i];
    a[i] = c[i - 1];
    b[i] = 2.0 * (a[i] + c[i]);
    r[i] = 6.0 * ((y[i + 1] - y[i]) / c[i] - (y[i] - y[i - 1]) / c[i - 1]);
  }
  b[1] = b[1] + 1.0 * c[0];
  b[n - 2] = b[n - 2] + 1.0 * c[n - 2];
  for (int i = 2; i < n - 1; ++i) {
    double t = a[i] / b[i - 1];
    b[i] = b[i] - t * c[i - 1];
    r[i] = r[i] - t * r[i - 1];
  }
  fdp[n - 2] = r[n - 2] / b[n - 2];
  for (int i = 2; i < n - 1; ++i) {
    int k = n - 1 - i;
    fdp[k] = (r[k] - c[k] * fdp[k + 1]) / b[k];
  }
  fdp[0] = 1.0 * fdp[1];
  fdp[n - 1] = 1.0 * fdp[n - 2];
  return 0;
}

int stabgpu_speval(int n, const double* x, const double* y, const double* fdp, double xx, double* f) {
  int i = 0;
  for (; i < n - 1; ++i)
    if (xx <= x[i + 1]) break;
  if (i >= n - 1) return 1;                            // the reference would index out of bounds here
  double dxm = xx - x[i], dxp = x[i + 1] - xx, del = x[i + 1] - x[i];
  *f = fdp[i] * dxp * (dxp * dxp / del - del) / 6.0 + fdp[i + 1] * dxm * (dxm * dxm / del - del) / 6.0 +
       y[i] * dxp / del + y[i + 1] * dxm / del;
  return 0;
}

int stabgpu_getmean_table(int nrows, const double* table, int ny, const double* y, double* vm) {
  std::vector<double> ym(nrows), vt(nrows), vs(nrows);
  for (int j = 0; j < nrows; ++j) ym[j] = table[(size_t)j * 6];
  const double ymaxm = ym[nrows - 1];
  for (int k = 0; k < 5; ++k) {
    for (int j = 0; j < nrows; ++j) vt[j] = (k == 2) ? 0.0 : table[(size_t)j * 6 + 1 + k];   // getmean.f90:75
    if (stabgpu_spline(nrows, ym.data(), vt.data(), vs.data())) return 1;
    for (int j = 0; j < ny; ++j) {
      if (y[j] <= ymaxm) {
        if (stabgpu_speval(nrows, ym.data(), vt.data(), vs.data(), y[j], &vm[j + (size_t)k * ny])) return 2;
      } else {
        vm[j + (size_t)k * ny] = vt[nrows - 1];        // constant beyond the table, getmean.f90:103-108
      }
    }
  }
  return 0;
}

int stabgpu_read_profile(const char* path, int* nrows, double* table, int max_rows) {
  FILE* f = std::fopen(path, "r");
  if (!f) return 1;
  char line[1024];
  int n = 0;
  while (std::fgets(line, sizeof line, f)) {
    if (line[0] == '#') continue;                      // getmean.f90:47-50
    for (char* q = line; *q; ++q) if (*q == 'D' || *q == 'd') *q = 'E';
    double v[6];
    if (std::sscanf(line, "%lf %lf %lf %lf %lf %lf", v, v + 1, v + 2, v + 3, v + 4, v + 5) != 6) continue;
    if (n >= max_rows) { std::fclose(f); return 2; }
    for (int k = 0; k < 6; ++k) table[(size_t)n * 6 + k] = v[k];
    ++n;
  }
  std::fclose(f);
  *nrows = n;
  return n >= 4 ? 0 : 3;
}

static void matmul_sq(int n, const double* A, const double* B, double* C) {   // C = A*B, column-major
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < n; ++i) {
      double s = 0.0;
      for (int k = 0; k < n; ++k) s += A[i + (size_t)k * n] * B[k + (size_t)j * n];
      C[i + (size_t)j * n] = s;
    }
}

int stabgpu_mean_gradients(int ny, int wallt, const double* vm, const double* deta, const double* d2eta,
                           double* D1, double* D2, double* Dt2w, double* g2vm, double* g22vm) {
  stabgpu_chebyd(ny - 1, D1);
  matmul_sq(ny, D1, D1, D2);                           // temporal.f90:136
  const int w = ny - 1;
  for (int j = 0; j < ny; ++j) {
    if (wallt == 2) {                                  // Dt1 = D1 with zero wall row; Dt2 = D1*Dt1 (:137-140)
      double s = 0.0;
      for (int k = 0; k < ny; ++k) s += D1[w + (size_t)k * ny] * ((k == w) ? 0.0 : D1[k + (size_t)j * ny]);
      Dt2w[j] = s;
    } else {
      Dt2w[j] = D2[w + (size_t)j * ny];
    }
  }
  if (g2vm && g22vm && vm) {
    for (int k = 0; k < 5; ++k)
      for (int i = 0; i < ny; ++i) {
        double s1 = 0.0, s2 = 0.0;
        for (int j = 0; j < ny; ++j) {
          s1 += D1[i + (size_t)j * ny] * vm[j + (size_t)k * ny];
          s2 += D2[i + (size_t)j * ny] * vm[j + (size_t)k * ny];
        }
        g22vm[i + (size_t)k * ny] = s2 * (deta[i] * deta[i]) + s1 * d2eta[i];     // temporal.f90:176
        g2vm[i + (size_t)k * ny] = s1 * deta[i];                                  // temporal.f90:177
      }
  }
  return 0;
}

// circh.f90:35-188.  The scalar prologue is evaluated exactly as written there (including the
// branches that are dead for s = 0), then the O(ny) metric formulas.
int stabgpu_circh(double* x_inout, int ny, const double* r, double* h5) {
  double* h = h5; double* dhds = h5 + ny; double* dhdr = h5 + 2 * ny; double* dhdsr = h5 + 3 * ny; double* dhdrr = h5 + 4 * ny;
  if (*x_inout == -1.0) {                               // flat plate: no curvature terms (circh.f90:41-46)
    for (int i = 0; i < ny; ++i) { h[i] = 1.0; dhds[i] = dhdr[i] = dhdsr[i] = dhdrr[i] = 0.0; }
    return 0;
  }
  // The reference evaluates the body-fitted metrics of a circular cylinder of radius R = x at the arc length s = 0 only
  // (circh.f90:47-49 overwrites x with 0 after taking the radius) through its general curve machinery (circh.f90:51-188:
  // local tangent and normal, their first and second arc-length derivatives, branches on the steeper coordinate).  On a
  // circle that machinery reduces to constant curvature 1/R:
  //     h = 1 + r / R,   dh/dr = 1 / R,   dh/ds = d2h/(ds dr) = d2h/dr2 = 0.
  // The literal evaluation (oracle/stab_oracle.py::circh restates it line by line) differs from these values by at most
  // 1 ulp in h, 2e-16 relative in dh/dr and 3e-20 absolute in the three vanishing terms, which come out as rounding
  // residue of cos(pi/2) (tests/test_host_cabi.py::test_circh_matches_oracle, test_circh_closed_form_against_literal).
  const double radius = std::fabs(*x_inout);            // the literal evaluation depends on R only through R^2 and |R|
  if (!(radius > 0.0) || !std::isfinite(radius)) return 1;   // R = 0 divides by zero in the reference
  *x_inout = 0.0;
  const double curv = 1.0 / radius;
  for (int i = 0; i < ny; ++i) {
    h[i] = 1.0 + r[i] / radius;
    dhdr[i] = curv;
    dhds[i] = dhdsr[i] = dhdrr[i] = 0.0;
  }
  return 0;
}

static int nint_(double v) { return (int)(v >= 0.0 ? std::floor(v + 0.5) : -std::floor(-v + 0.5)); }

int stabgpu_mtemporal_points(double amin, double amax, double ainc, double bmin, double bmax, double binc,
                             double* alpha_r, double* beta_r, int max_pts) {
  // mtemporal.f90:25 divides by the increments unguarded (a zero increment is a floating exception there); here a zero
  // or non-finite increment, or a count beyond 10^7 points, is refused with -1 instead of an undefined int conversion
  const double qa = (amax - amin) / ainc, qb = (bmax - bmin) / binc;
  if (!(ainc != 0.0) || !(binc != 0.0) || !std::isfinite(qa) || !std::isfinite(qb) || std::fabs(qa) > 1.0e7 || std::fabs(qb) > 1.0e7) return -1;
  int na = nint_(qa); if (na < 1) na = 1;                          // mtemporal.f90:25 (no +1)
  int nb = nint_(qb); if (nb < 1) nb = 1;
  if ((long long)na * nb > 10000000LL) return -1;
  int n = 0;
  for (int ia = 1; ia <= na; ++ia)
    for (int ib = 1; ib <= nb; ++ib) {
      if (n < max_pts) { alpha_r[n] = amin + (double)(ia - 1) * ainc; beta_r[n] = bmin + (double)(ib - 1) * binc; }
      ++n;
    }
  return n;
}

int stabgpu_mspatial_points(double omin, double omax, double oinc, double bmin, double bmax, double binc,
                            double* omega_r, double* beta_r, int max_pts) {
  if (oinc == 0.0) oinc = 1.0;                                     // mspatial.f90:68-69
  if (binc == 0.0) binc = 1.0;
  const double qo = (omax - omin) / oinc, qb = (bmax - bmin) / binc;
  if (!std::isfinite(qo) || !std::isfinite(qb) || std::fabs(qo) > 1.0e7 || std::fabs(qb) > 1.0e7) return -1;
  const int no = nint_(qo) + 1, nb = nint_(qb) + 1;
  if ((long long)no * nb > 10000000LL) return -1;
  int n = 0;
  for (int io = 0; io < no; ++io)
    for (int ib = 0; ib < nb; ++ib) {
      if (n < max_pts) { omega_r[n] = omin + (double)io * oinc; beta_r[n] = bmin + (double)ib * binc; }
      ++n;
    }
  return n;
}

void stabgpu_shard_range(int npts, int rank, int world, int* lo, int* hi) {
  *lo = (int)(((long long)npts * rank) / world);
  *hi = (int)(((long long)npts * (rank + 1)) / world);
}

static void put_rec(FILE* f, const void* data, int32_t bytes) {
  std::fwrite(&bytes, 4, 1, f);
  std::fwrite(data, 1, (size_t)bytes, f);
  std::fwrite(&bytes, 4, 1, f);
}

int stabgpu_write_eig_file(const char* path, const stabgpu_params* p, int itype, int ind,
                           const double* omega, const double* alpha, const double* beta, double x,
                           const double* y, const double* eta, const double* deta, const double* d2eta,
                           const double* eig, const double* evec) {
  FILE* f = std::fopen(path, "wb");
  if (!f) return 1;
  const int ny = p->ny;
  const int nmax = (itype == 1 ? 1 : 2) * STABGPU_NDOF * ny;
  int32_t r1[10] = {ind, ny, STABGPU_NDOF, itype, p->ievec, p->curve, p->top, p->wall, p->wallt, p->ider ? 1 : 0};
  put_rec(f, r1, 40);
  double r2[9] = {omega[0], omega[1], alpha[0], alpha[1], beta[0], beta[1], p->Re, p->Ma, p->Pr};
  put_rec(f, r2, 72);
  std::vector<double> r3;
  r3.push_back(x);
  r3.insert(r3.end(), y, y + ny); r3.insert(r3.end(), eta, eta + ny);
  r3.insert(r3.end(), deta, deta + ny); r3.insert(r3.end(), d2eta, d2eta + ny);
  r3.push_back(p->yi); r3.push_back(p->ymax);
  put_rec(f, r3.data(), (int32_t)(r3.size() * 8));
  put_rec(f, eig, (int32_t)(16 * nmax));
  if (evec) {
    // 16 n^2 exceeds a 32-bit record marker only for n > 11585 -- far above any Ny here
    put_rec(f, evec, (int32_t)((size_t)16 * nmax * nmax));
  }
  std::fclose(f);
  return 0;
}

}  // extern "C"
