// stabgpu_cli -- C++ host harness over the C ABI (include/stabgpu.h): the reference's front end for the
// hot path where no Fortran compiler is available.  Reads the positional stdin deck of `stab`
// (input.f90:15-122, stab.f90:46-92; itype 1 temporal, 2 spatial, 7 temporal (alpha,beta) sweep), the mean
// profile `profile.<ind>` from the working directory (getmean.f90:27-111), solves on the GPU and writes the
// reference's unformatted records: `evec.dat` (single point) or `eig.<iver>` per sweep point
// (temporal.f90:883-890, spatial.f90:1120-1126, mtemporal.f90:25-39).
//   usage:  stabgpu_cli < temporal.inp          (same decks as `stab < temporal.inp`)
#include <cmath>
#include <cstdio>
#include <cstdlib>
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
  if (itype == 1 || itype == 2) {
    { auto v = next_numbers(in); s1[0] = v[0]; s1[1] = v[1]; }
    { auto v = next_numbers(in); s2[0] = v[0]; s2[1] = v[1]; }
  }
  const int ind = (int)next_numbers(in)[0];
  double x = 0.0;
  if (itype == 2) { x = next_numbers(in)[0]; p.x = x; }
  if (p.ider == 0) { std::fprintf(stderr, "stabgpu_cli: ider=0 (getmean2) is served by the library API, not by this harness\n"); return 1; }

  const int ny = p.ny, n = STABGPU_NDOF * ny;
  std::vector<double> y(ny), eta(ny), deta(ny), d2eta(ny), vm((size_t)ny * 5), table((size_t)200000 * 6), h5;
  if (stabgpu_sgengrid(ny, p.yi, p.ymax, y.data(), eta.data(), deta.data(), d2eta.data())) die("sgengrid (tanh map is not supported)");
  char pname[64]; std::snprintf(pname, sizeof pname, "profile.%d", ind);
  int nrows = 0;
  if (stabgpu_read_profile(pname, &nrows, table.data(), 200000)) { std::fprintf(stderr, "stabgpu_cli: cannot read %s\n", pname); return 1; }
  if (stabgpu_getmean_table(nrows, table.data(), ny, y.data(), vm.data())) die("getmean");
  if (stabgpu_init(-1)) die("init");

  const double zero2[2] = {0, 0};
  if (itype == 1) {
    std::vector<double> omg((size_t)2 * n), evec((size_t)2 * n * n);
    int info = 0;
    if (stabgpu_temporal_batch(&p, vm.data(), nullptr, nullptr, deta.data(), d2eta.data(), 1, s1, s2, nullptr, nullptr, 1,
                               omg.data(), evec.data(), &info)) die("temporal_batch");
    if (info != 0) { std::fprintf(stderr, "Error in eigensolver: info = %d\n", info); return 1; }   // temporal.f90:776-785,806-809
    if (stabgpu_write_eig_file("evec.dat", &p, 1, ind, zero2, s1, s2, x, y.data(), eta.data(), deta.data(), d2eta.data(),
                               omg.data(), evec.data())) die("write evec.dat");
    std::printf(" temporal: %d eigenvalues written to evec.dat\n", n);
  } else if (itype == 2) {
    double xr = x;
    if (p.curve == 2) { h5.resize((size_t)ny * 5); stabgpu_circh(&xr, ny, y.data(), h5.data()); }
    else if (p.curve != 0) { std::fprintf(stderr, "stabgpu_cli: curve=%d is not supported\n", p.curve); return 1; }
    const int N = 2 * n;
    std::vector<double> alp((size_t)2 * N), evec(p.ievec == 1 ? (size_t)2 * N * N : 0);
    int info = 0;
    if (stabgpu_spatial_batch(&p, vm.data(), nullptr, nullptr, deta.data(), d2eta.data(), h5.empty() ? nullptr : h5.data(), 1, s1, s2,
                              nullptr, nullptr, p.ievec == 1, alp.data(), p.ievec == 1 ? evec.data() : nullptr, &info)) die("spatial_batch");
    if (info != 0) std::fprintf(stderr, "WARNING: eigensolver info = %d\n", info);                   // spatial.f90:1050-1056: warn and continue
    if (stabgpu_write_eig_file("evec.dat", &p, 2, ind, s1, zero2, s2, xr, y.data(), eta.data(), deta.data(), d2eta.data(),
                               alp.data(), p.ievec == 1 ? evec.data() : nullptr)) die("write evec.dat");
    std::printf(" spatial: %d eigenvalues written to evec.dat\n", N);
  } else if (itype == 7) {                                              // mtemporal.f90:20-39
    auto a = next_numbers(in), b = next_numbers(in);
    const int npts = stabgpu_mtemporal_points(a[0], a[1], a[2], b[0], b[1], b[2], nullptr, nullptr, 0);
    std::vector<double> ar(npts), br(npts), al((size_t)2 * npts, 0.0), be((size_t)2 * npts, 0.0);
    stabgpu_mtemporal_points(a[0], a[1], a[2], b[0], b[1], b[2], ar.data(), br.data(), npts);
    for (int k = 0; k < npts; ++k) { al[2 * k] = ar[k]; be[2 * k] = br[k]; }
    std::vector<double> omg((size_t)2 * n * npts);
    std::vector<int> info(npts);
    if (stabgpu_temporal_batch(&p, vm.data(), nullptr, nullptr, deta.data(), d2eta.data(), npts, al.data(), be.data(), nullptr, nullptr, 0,
                               omg.data(), nullptr, info.data())) die("temporal_batch");
    for (int k = 0; k < npts; ++k) {
      if (k + 1 >= 10000) { std::fprintf(stderr, "Error in MakeName:  iver too large\n"); return 1; }  // mtemporal.f90:53-76
      char fn[64]; std::snprintf(fn, sizeof fn, "eig.%d", k + 1);
      std::printf(" %4d alpha = %13.6e beta = %13.6e info = %d\n", k + 1, ar[k], br[k], info[k]);
      if (stabgpu_write_eig_file(fn, &p, 1, ind, zero2, &al[2 * k], &be[2 * k], x, y.data(), eta.data(), deta.data(), d2eta.data(),
                                 &omg[(size_t)2 * n * k], nullptr)) die("write eig file");
    }
  } else {
    std::fprintf(stderr, "stabgpu_cli: itype = %d is outside the supported path (1, 2, 7)\n", itype);
    return 1;
  }
  stabgpu_finalize();
  return 0;
}
