// stabgpu_cli -- C++ host harness over the C ABI (include/stabgpu.h): the reference's front end for the
// hot path where no Fortran compiler is available.  Reads the positional stdin deck of `stab`
// (input.f90:15-122, stab.f90:46-92): itype 1 temporal, 2 spatial, 7 temporal (alpha,beta) sweep
// (mtemporal.f90:20-39), 8 spatial (omega,beta) sweep over the stations of `delta.dat` (mspatial.f90:20-96);
// the mean profile `profile.<ind>` (and `first.<ind>`, `second.<ind>` when ider=0, getmean2.f90:26-187) from the
// working directory (getmean.f90:27-111); solves on the GPUs of the box (ONE batched call per sweep / station,
// sharded over every visible device by stabgpu_init_multi; STABGPU_DEVICES=k limits it) and writes the
// reference's unformatted records: `evec.dat` (single point) or `eig.<iver>` per sweep point
// (temporal.f90:883-890, spatial.f90:1120-1126).
//   usage:  stabgpu_cli < temporal.inp          (same decks as `stab < temporal.inp`)
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>
#include "../include/stabgpu.h"

static std::vector<double> next_numbers(std::istream& in) {           // list-directed read of one deck line
  std::string line;
  while (std::getline(in, line)) {
    size_t bang = line.find('!');
    if (bang != std::string::npos) line = line.substr(0, bang);
    for (char& c : line) if (c == ',' || c == 'd' || c == 'D') c = (c == ',') ? ' ' : 'e';
    std::istringstream ss(line);
    std::vector<double> v; double x;
    while (ss >> x) v.push_back(x);
    if (!v.empty()) return v;
  }
  std::fprintf(stderr, "stabgpu_cli: unexpected end of deck\n");
  std::exit(1);
}

static void die(const char* what) { std::fprintf(stderr, "stabgpu_cli: %s: %s\n", what, stabgpu_last_error()); std::exit(1); }

struct Mean {                                                         // grid + mean flow on the grid for one profile index
  std::vector<double> y, eta, deta, d2eta, vm, g2, g22, h5;
  double x_out = 0.0;
};

static std::vector<double> table_on_grid(const char* base, int ind, int ny, const std::vector<double>& y) {
  static std::vector<double> table((size_t)200000 * 6);
  char name[96]; std::snprintf(name, sizeof name, "%s.%d", base, ind);
  int nrows = 0;
  if (stabgpu_read_profile(name, &nrows, table.data(), 200000)) { std::fprintf(stderr, "stabgpu_cli: cannot read %s\n", name); std::exit(1); }
  std::vector<double> out((size_t)ny * 5);
  if (stabgpu_getmean_table(nrows, table.data(), ny, y.data(), out.data())) die("getmean");
  return out;
}

// sgengrid + getmean [+ getmean2 when ider = 0] [+ circh when spatial and curve = 2] (temporal.f90:95-103, spatial.f90:96-125)
static Mean load_mean(const stabgpu_params& p, int ind, bool spatial, double x) {
  Mean m;
  const int ny = p.ny;
  m.y.resize(ny); m.eta.resize(ny); m.deta.resize(ny); m.d2eta.resize(ny);
  if (stabgpu_sgengrid(ny, p.yi, p.ymax, m.y.data(), m.eta.data(), m.deta.data(), m.d2eta.data())) die("sgengrid (the tanh map is not supported)");
  m.vm = table_on_grid("profile", ind, ny, m.y);
  if (p.ider == 0) { m.g2 = table_on_grid("first", ind, ny, m.y); m.g22 = table_on_grid("second", ind, ny, m.y); }
  m.x_out = x;
  if (spatial) {
    if (p.curve == 2) { m.h5.resize((size_t)ny * 5); double xr = x; stabgpu_circh(&xr, ny, m.y.data(), m.h5.data()); m.x_out = xr; }
    else if (p.curve != 0) { std::fprintf(stderr, "stabgpu_cli: curve=%d is not supported\n", p.curve); std::exit(1); }
  }
  return m;
}

static const double* opt(const std::vector<double>& v) { return v.empty() ? nullptr : v.data(); }

int main() {
  stabgpu_params p; stabgpu_params_default(&p);
  std::istream& in = std::cin;
  p.mattyp = (int)next_numbers(in)[0];
  double T0 = 0.0;
  if (p.mattyp == 1) T0 = next_numbers(in)[0];
  { auto v = next_numbers(in); p.Ma = v[0]; p.Re = v[1]; p.Pr = v[2]; }
  { auto v = next_numbers(in); p.ny = (int)v[0]; p.yi = v[1]; p.ymax = v[2]; }
  p.ievec = (int)next_numbers(in)[0];
  p.ider = next_numbers(in)[0] == 0.0 ? 0 : 1;
  { auto v = next_numbers(in); p.top = (int)v[0]; p.wall = (int)v[1]; p.wallt = (int)v[2]; p.curve = (int)v[3]; }
  const int itype = (int)next_numbers(in)[0];
  stabgpu_edge_properties(&p, T0);
  double s1[2] = {0, 0}, s2[2] = {0, 0};                               // alpha|omega, beta
  int ind = 0;                                                         // stab.f90: read for itype 1-6 only; mtemporal(ind) runs with 0
  double x = 0.0;
  if (itype == 1 || itype == 2) {
    { auto v = next_numbers(in); s1[0] = v[0]; s1[1] = v[1]; }
    { auto v = next_numbers(in); s2[0] = v[0]; s2[1] = v[1]; }
    ind = (int)next_numbers(in)[0];
    if (itype == 2) { x = next_numbers(in)[0]; p.x = x; }
  }
  const int ny = p.ny, n = STABGPU_NDOF * ny;

  {
    int want = 0, got = 0;
    if (const char* e = std::getenv("STABGPU_DEVICES")) want = std::atoi(e);
    if (stabgpu_init_multi(want, &got)) die("init");
  }
  const double zero2[2] = {0, 0};
  if (itype == 1) {
    Mean m = load_mean(p, ind, false, x);
    std::vector<double> omg((size_t)2 * n), evec((size_t)2 * n * n);
    int info = 0;
    if (stabgpu_temporal_batch(&p, m.vm.data(), opt(m.g2), opt(m.g22), m.deta.data(), m.d2eta.data(), 1, s1, s2, nullptr, nullptr, 1,
                               omg.data(), evec.data(), &info)) die("temporal_batch");
    if (info != 0) { std::fprintf(stderr, "Error in eigensolver: info = %d\n", info); return 1; }   // temporal.f90:776-785,806-809
    if (stabgpu_write_eig_file("evec.dat", &p, 1, ind, zero2, s1, s2, x, m.y.data(), m.eta.data(), m.deta.data(), m.d2eta.data(),
                               omg.data(), evec.data())) die("write evec.dat");
    std::printf(" temporal: %d eigenvalues written to evec.dat\n", n);
  } else if (itype == 2) {
    Mean m = load_mean(p, ind, true, x);
    const int N = 2 * n;
    std::vector<double> alp((size_t)2 * N), evec(p.ievec == 1 ? (size_t)2 * N * N : 0);
    int info = 0;
    if (stabgpu_spatial_batch(&p, m.vm.data(), opt(m.g2), opt(m.g22), m.deta.data(), m.d2eta.data(), opt(m.h5), 1, s1, s2,
                              nullptr, nullptr, p.ievec == 1, alp.data(), p.ievec == 1 ? evec.data() : nullptr, &info)) die("spatial_batch");
    if (info != 0) std::fprintf(stderr, "WARNING: eigensolver info = %d\n", info);                   // spatial.f90:1050-1056: warn and continue
    if (stabgpu_write_eig_file("evec.dat", &p, 2, ind, s1, zero2, s2, m.x_out, m.y.data(), m.eta.data(), m.deta.data(), m.d2eta.data(),
                               alp.data(), p.ievec == 1 ? evec.data() : nullptr)) die("write evec.dat");
    std::printf(" spatial: %d eigenvalues written to evec.dat\n", N);
  } else if (itype == 7) {                                              // mtemporal.f90:20-39; `temporal` always computes and writes the vectors
    auto a = next_numbers(in), b = next_numbers(in);                    // (temporal.f90:803,883-890), here they follow the deck's ievec
    Mean m = load_mean(p, ind, false, x);
    const int npts = stabgpu_mtemporal_points(a[0], a[1], a[2], b[0], b[1], b[2], nullptr, nullptr, 0);
    if (npts < 0) { std::fprintf(stderr, "stabgpu_cli: mtemporal: zero or non-finite increment\n"); return 1; }
    std::vector<double> ar(npts), br(npts), al((size_t)2 * npts, 0.0), be((size_t)2 * npts, 0.0);
    stabgpu_mtemporal_points(a[0], a[1], a[2], b[0], b[1], b[2], ar.data(), br.data(), npts);
    for (int k = 0; k < npts; ++k) { al[2 * k] = ar[k]; be[2 * k] = br[k]; }
    const bool vec = p.ievec == 1;
    const size_t per = (size_t)2 * n * n;
    int chunk = npts;                                                   // with vectors: at most ~4 GB of host memory per batched call
    if (vec) { const size_t c = ((size_t)4 << 30) / (per * sizeof(double)); chunk = (int)(c < 1 ? 1 : (c < (size_t)npts ? c : (size_t)npts)); }
    std::vector<double> omg((size_t)2 * n * chunk), evec(vec ? per * chunk : 0);
    std::vector<int> info(chunk);
    for (int k0 = 0; k0 < npts; k0 += chunk) {
      const int m1 = npts - k0 < chunk ? npts - k0 : chunk;
      if (stabgpu_temporal_batch(&p, m.vm.data(), opt(m.g2), opt(m.g22), m.deta.data(), m.d2eta.data(), m1, &al[2 * k0], &be[2 * k0], nullptr,
                                 nullptr, vec ? 1 : 0, omg.data(), vec ? evec.data() : nullptr, info.data())) die("temporal_batch");
      for (int k = 0; k < m1; ++k) {
        const int iver = k0 + k + 1;
        if (iver >= 10000) { std::fprintf(stderr, "Error in MakeName:  iver too large\n"); return 1; }  // mtemporal.f90:53-76
        char fn[64]; std::snprintf(fn, sizeof fn, "eig.%d", iver);
        std::printf(" %4d alpha = %13.6e beta = %13.6e info = %d\n", iver, ar[k0 + k], br[k0 + k], info[k]);
        if (info[k] != 0) { std::fprintf(stderr, "Error in eigensolver: info = %d\n", info[k]); return 1; }   // temporal.f90:806-809 stops
        if (stabgpu_write_eig_file(fn, &p, 1, ind, zero2, &al[2 * (k0 + k)], &be[2 * (k0 + k)], x, m.y.data(), m.eta.data(), m.deta.data(),
                                   m.d2eta.data(), &omg[(size_t)2 * n * k], vec ? &evec[per * k] : nullptr)) die("write eig file");
      }
    }
  } else if (itype == 8) {                                              // mspatial.f90:20-96
    auto o = next_numbers(in), b = next_numbers(in), ii = next_numbers(in);
    const int dtype = (int)next_numbers(in)[0];
    if (dtype != 0) { std::fprintf(stderr, "stabgpu_cli: mspatial dtype=1 (finite differences) is outside the supported path\n"); return 1; }
    int ind1 = (int)ii[0], ind2 = (int)ii[1], ind_inc = (int)ii[2];
    if (ind_inc == 0) ind_inc = 1;
    std::vector<double> xb, delta;
    {
      std::ifstream f("delta.dat");
      if (!f) { std::fprintf(stderr, "stabgpu_cli: cannot open delta.dat\n"); return 1; }
      std::string line;
      while (std::getline(f, line)) {
        size_t q = line.find_first_not_of(" \t");
        if (q == std::string::npos || line[q] == '#') continue;
        for (char& c : line) if (c == ',' || c == 'd' || c == 'D') c = (c == ',') ? ' ' : 'e';
        std::istringstream ss(line); double u, v;
        if (ss >> u >> v) { xb.push_back(u); delta.push_back(v); }
      }
    }
    if (ind1 < 1 || ind2 > (int)xb.size()) {
      std::fprintf(stderr, "%d %d 1 %d\nERROR: illegal index...check delta.dat\n", ind1, ind2, (int)xb.size());   // mspatial.f90:46-50
      return 1;
    }
    const int npts = stabgpu_mspatial_points(o[0], o[1], o[2], b[0], b[1], b[2], nullptr, nullptr, 0);
    if (npts < 1) { std::fprintf(stderr, "stabgpu_cli: mspatial: bad sweep range\n"); return 1; }
    std::vector<double> orr(npts), br(npts), om((size_t)2 * npts, 0.0), be((size_t)2 * npts, 0.0);
    stabgpu_mspatial_points(o[0], o[1], o[2], b[0], b[1], b[2], orr.data(), br.data(), npts);
    for (int k = 0; k < npts; ++k) { om[2 * k] = orr[k]; be[2 * k] = br[k]; }
    const int N = 2 * n;
    const bool vec = p.ievec == 1;
    const size_t per = (size_t)2 * N * N;
    int chunk = npts;
    if (vec) { const size_t c = ((size_t)4 << 30) / (per * sizeof(double)); chunk = (int)(c < 1 ? 1 : (c < (size_t)npts ? c : (size_t)npts)); }
    std::vector<double> alp((size_t)2 * N * chunk), evec(vec ? per * chunk : 0);
    std::vector<int> info(chunk);
    int iver = 0;
    for (int st = ind1; st <= ind2; st += ind_inc) {
      stabgpu_params q = p;
      q.x = xb[st - 1]; q.yi = 2.0 * delta[st - 1];                      // mspatial.f90:77-79
      Mean m = load_mean(q, st, true, q.x);
      for (int k0 = 0; k0 < npts; k0 += chunk) {
        const int m1 = npts - k0 < chunk ? npts - k0 : chunk;
        if (stabgpu_spatial_batch(&q, m.vm.data(), opt(m.g2), opt(m.g22), m.deta.data(), m.d2eta.data(), opt(m.h5), m1, &om[2 * k0], &be[2 * k0],
                                  nullptr, nullptr, vec ? 1 : 0, alp.data(), vec ? evec.data() : nullptr, info.data())) die("spatial_batch");
        for (int k = 0; k < m1; ++k) {
          ++iver;
          if (iver >= 10000) { std::fprintf(stderr, "Error in MakeName:  iver too large\n"); return 1; }
          char fn[64]; std::snprintf(fn, sizeof fn, "eig.%d", iver);
          std::printf(" %4d %4d x = %13.6e Yi = %13.6e omega = %13.6e beta = %13.6e info = %d\n", st, iver, q.x, q.yi, orr[k0 + k], br[k0 + k], info[k]);
          if (info[k] != 0) std::fprintf(stderr, "WARNING: eigensolver info = %d\n", info[k]);       // spatial.f90:1050-1056
          if (stabgpu_write_eig_file(fn, &q, 2, st, &om[2 * (k0 + k)], zero2, &be[2 * (k0 + k)], m.x_out, m.y.data(), m.eta.data(), m.deta.data(),
                                     m.d2eta.data(), &alp[(size_t)2 * N * k], vec ? &evec[per * k] : nullptr)) die("write eig file");
        }
      }
    }
  } else {
    std::fprintf(stderr, "stabgpu_cli: itype = %d is outside the supported path (1, 2, 7, 8)\n", itype);
    return 1;
  }
  stabgpu_finalize();
  return 0;
}
