// emu.cpp -- single-threaded host TRACE of the CTA-level kernel source (TEST INFRASTRUCTURE ONLY).
// Compiles the same .cuh files as the CUDA build with -DSTAB_EMU, where a CTA degenerates to one
// thread and barriers to no-ops, so the algorithmic logic (balancing, Hessenberg, QR, ...) can be
// unit-tested in a container without a GPU.  Not linked into libstabgpu.so, exports only emu_*.
#include <vector>
#include <cstring>
#include <cstdlib>
#include <algorithm>
#include "../../stab_b200/csrc/common.cuh"
#include "../../stab_b200/csrc/tables.cuh"
#include "../../stab_b200/csrc/assemble.cuh"
#include "../../stab_b200/csrc/balance.cuh"
#include "../../stab_b200/csrc/hessenberg.cuh"
#include "../../stab_b200/csrc/hqr.cuh"

using namespace stab;

extern "C" {

int emu_balance(cplx* A, int n, double* scale, int* ilo, int* ihi) {
  std::vector<double> red(256);
  std::vector<int> cnt(n);
  Cta c = make_cta(red.data());
  const int bb = 4;
  std::vector<double> wsp(balance_wsp_doubles(n, bb));
  cta_balance(c, A, n, n, scale, cnt.data(), wsp.data(), bb, *ilo, *ihi);
  return 0;
}

int emu_hessenberg(cplx* A, int n, int ilo, int ihi, cplx* tau) {
  std::vector<double> red(256);
  std::vector<cplx> sv(n), sy(n);
  Cta c = make_cta(red.data());
  cta_hessenberg(c, A, n, n, ilo, ihi, tau, sv.data(), sy.data());
  return 0;
}

// H must be upper Hessenberg with explicit zeros below the subdiagonal
int emu_hqr(cplx* H, int n, int ilo, int ihi, cplx* w, int W, int ns_max, int steps_max, int nw, int nibble) {
  std::vector<double> red(256);
  Cta c = make_cta(red.data());
  HqrSmem sh;
  sh.W = W; sh.ldw = W + 1; sh.nw = nw; sh.nibble = nibble;
  std::vector<cplx> win((size_t)W * (W + 1)), shifts(ns_max), sm((size_t)ns_max * (ns_max + 1) > 64 ? (size_t)ns_max * (ns_max + 1) : 64);
  std::vector<Rot> rec(std::max((size_t)steps_max * ns_max, 2 * ((size_t)W + 2))), cur(2 * ns_max);
  SmallCtl ctl;
  sh.win = win.data(); sh.rec = rec.data(); sh.steps_max = steps_max; sh.ns_max = ns_max;
  sh.cur = cur.data(); sh.shifts = shifts.data(); sh.sm = sm.data(); sh.ctl = &ctl; sh.prof = nullptr;
  return cta_hqr(c, sh, H, n, n, ilo, ihi, w);
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// assembly / LU / eigenvector traces
// ---------------------------------------------------------------------------------------------
#include "../../include/stabgpu.h"
#include "../../stab_b200/csrc/evec.cuh"
#include "../../stab_b200/csrc/lu.cuh"

namespace {
struct HostGrid {
  std::vector<double> D1, D2, Dt2w, g2, g22;
  GridDev g;
};
void make_grid(HostGrid& hg, const stabgpu_params* p, const double* vm, const double* g2vm, const double* g22vm,
               const double* deta, const double* d2eta, const double* h5) {
  const int ny = p->ny;
  hg.D1.resize((size_t)ny * ny); hg.D2.resize((size_t)ny * ny); hg.Dt2w.resize(ny); hg.g2.resize(ny * 5); hg.g22.resize(ny * 5);
  stabgpu_mean_gradients(ny, p->wallt, vm, deta, d2eta, hg.D1.data(), hg.D2.data(), hg.Dt2w.data(), hg.g2.data(), hg.g22.data());
  if (!p->ider) { std::memcpy(hg.g2.data(), g2vm, sizeof(double) * ny * 5); std::memcpy(hg.g22.data(), g22vm, sizeof(double) * ny * 5); }
  hg.g.ny = ny; hg.g.wallt = p->wallt; hg.g.top = p->top;
  hg.g.D1 = hg.D1.data(); hg.g.D2 = hg.D2.data(); hg.g.Dt2w = hg.Dt2w.data(); hg.g.deta = deta; hg.g.d2eta = d2eta;
  hg.g.vm = vm; hg.g.g2vm = hg.g2.data(); hg.g.g22vm = hg.g22.data(); hg.g.h5 = h5;
}
Phys phys_from(const stabgpu_params* p) {
  Phys q;
  q.Ma = p->Ma; q.Re = p->Re; q.Pr = p->Pr; q.gamma = p->gamma; q.gamma1 = p->gamma1; q.cp = p->cp;
  q.Te = p->Te; q.rmue = p->rmue; q.rlme = p->rlme; q.cone = p->cone;
  for (int k = 0; k < 3; ++k) q.datmat[k] = p->datmat[k];
  q.mattyp = p->mattyp;
  q.navier = !(p->Re >= 1.0e98 || p->Re == 0.0);
  return q;
}
}  // namespace

extern "C" {

// M = B0^-1 A0 (apply_b0inv=1) or A0 (0); B0 optional
int emu_temporal_matrix(const stabgpu_params* p, const double* vm, const double* g2vm, const double* g22vm, const double* deta,
                        const double* d2eta, const double* alpha, const double* beta, int apply_b0inv, cplx* M, cplx* B0) {
  HostGrid hg; make_grid(hg, p, vm, g2vm, g22vm, deta, d2eta, nullptr);
  const int ny = p->ny, n = 5 * ny;
  Phys ph = phys_from(p);
  PointTemporal pt; pt.alpha = mk(alpha[0], alpha[1]); pt.beta = mk(beta[0], beta[1]);
  std::vector<NodeCoef3> co(ny);
  std::vector<cplx> blk((size_t)ny * 25);
  for (int i = 0; i < ny; ++i) {
    NodeIn q = load_node(hg.g, i);
    Tables t; node_tables_temporal(q, ph, t);
    node_coef_temporal(t, i, ny, p->wallt, deta[i], d2eta[i], pt, apply_b0inv != 0, co[i]);
    node_b0_temporal(t, i, ny, p->wallt, &blk[(size_t)i * 25]);
  }
  for (int c = 0; c < n; ++c)
    for (int r = 0; r < n; ++r) {
      const int i = r / 5, e = r % 5, j = c / 5, v = c % 5;
      M[r + (size_t)c * n] = op_element(co[i].c1, co[i].c2, co[i].c0, hg.g, i, e, j, v);
      if (B0) B0[r + (size_t)c * n] = (i == j) ? blk[(size_t)i * 25 + e * 5 + v] : mk(0.0, 0.0);
    }
  return 0;
}

// C0, C1, C2 with the reference's signs; optionally the companion matrix after the LU reduction
int emu_spatial_matrices(const stabgpu_params* p, const double* vm, const double* g2vm, const double* g22vm, const double* deta,
                         const double* d2eta, const double* h5, const double* omega, const double* beta,
                         cplx* C0, cplx* C1, cplx* C2, cplx* companion, int* lu_info) {
  HostGrid hg; make_grid(hg, p, vm, g2vm, g22vm, deta, d2eta, h5);
  const int ny = p->ny, n = 5 * ny, n2 = 2 * n;
  Phys ph = phys_from(p);
  PointSpatial pt; pt.omega = mk(omega[0], omega[1]); pt.beta = mk(beta[0], beta[1]);
  std::vector<NodeCoefSpatial> co(ny);
  for (int i = 0; i < ny; ++i) {
    NodeIn q = load_node(hg.g, i);
    Tables t; node_tables_spatial(q, ph, t);
    node_coef_spatial(t, i, ny, p->wallt, p->top, deta[i], d2eta[i], pt, co[i]);
  }
  std::vector<cplx> Cw((size_t)n * n);
  if (companion) for (size_t k = 0; k < (size_t)n2 * n2; ++k) companion[k] = mk(0.0, 0.0);
  for (int c = 0; c < n; ++c)
    for (int r = 0; r < n; ++r) {
      const int i = r / 5, e = r % 5, j = c / 5, v = c % 5, k = e * 5 + v;
      cplx a = op_element(co[i].C0.c1, co[i].C0.c2, co[i].C0.c0, hg.g, i, e, j, v);
      cplx b1 = co[i].C1c1[k] * hg.D1[i + (size_t)j * ny];
      cplx b2 = mk(0.0, 0.0);
      if (i == j) { b1 += co[i].C1c0[k]; b2 = co[i].C2c0[k]; }
      Cw[r + (size_t)c * n] = a;
      if (C0) C0[r + (size_t)c * n] = a;
      if (C1) C1[r + (size_t)c * n] = -b1;
      if (C2) C2[r + (size_t)c * n] = -b2;
      if (companion) {
        companion[r + (size_t)c * n2] = b1;
        companion[r + (size_t)(n + c) * n2] = b2;
        if (r == c) companion[(n + r) + (size_t)c * n2] = mk(1.0, 0.0);
      }
    }
  if (companion) {
    std::vector<double> red(256);
    std::vector<cplx> sl(n);
    Cta c = make_cta(red.data());
    int info = cta_lu_solve(c, Cw.data(), n, n, companion, n2, n2, sl.data());
    if (lu_info) *lu_info = info;
  }
  return 0;
}

// Eigenvectors of the ORIGINAL matrix from (Hessenberg+reflectors, tau, scale, ilo, ihi) and eigenvalues lam
// lam[e] must be the eigenvalue found at index position e of the Hessenberg matrix (ZHSEQR order)
int emu_evec(const cplx* Hh, int n, int ilo, int ihi, const cplx* tau, const double* scale, const cplx* lam, int nlam,
             double hnorm, int scale_rows, cplx* V) {
  std::vector<int> blkend(n);
  {
    int end = n - 1;
    for (int i = n - 1; i >= 0; --i) {
      if (i < n - 1 && is_zero(Hh[(i + 1) + (size_t)i * n])) end = i;
      blkend[i] = end;
    }
  }
  Cta w = make_cta(nullptr);
  std::vector<cplx> c(n), y(n);
  std::vector<unsigned char> flag(n);
  int bad = 0;
  for (int e = 0; e < nlam; ++e)
    bad += warp_eigvec(w, Hh, n, n, ilo, ihi, tau, scale, lam[e], blkend[e], hnorm, scale_rows, c.data(), y.data(), flag.data(), V + (size_t)e * n);
  return bad;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// blocked Hessenberg reduction: the batched kernel schedule of stabgpu.cu traced for one matrix
// ---------------------------------------------------------------------------------------------
#include "../../stab_b200/csrc/hess_blocked.cuh"

extern "C" int emu_hess_blocked(cplx* A, int n, int ilo, int ihi, cplx* tau, cplx* Tout /* P*NB*NB or null */) {
  std::vector<double> red(256);
  Cta c = make_cta(red.data());
  const int P = (n - 1 + HB_NB - 1) / HB_NB;
  std::vector<cplx> Y((size_t)n * HB_NB), T((size_t)P * HB_NB * HB_NB), Yp((size_t)n * HB_CHUNKS), W((size_t)n * HB_NB);
  std::vector<cplx> sb(n), sw(HB_NB), st(HB_NB), sv(n);
  std::vector<cplx> scv(HB_NB);
  std::vector<double> smem(GemmCfg<64, 64>::smem_bytes / sizeof(double));
  int ilohi[2] = {ilo, ihi};
  HessBatch hb{A, (size_t)n * n, n, ilohi, tau, Y.data(), T.data(), Yp.data(), W.data(), P, 0};
  std::vector<cplx> tv(HB_NB);
  hb.tv = tv.data();
  const int tiles = (n + 63) / 64;
  for (int p = 0; p < P; ++p) {
    for (int j = 0; j < HB_NB; ++j) {
      cta_hb_panel_step(c, hb, 0, p, j, red.data(), sb.data(), sw.data(), st.data(), scv.data());
      for (int rt = 0; rt * HB_GEMV_ROWS < n; ++rt)
        for (int ch = 0; ch < HB_CHUNKS; ++ch) cta_hb_gemv(c, hb, 0, p, j, rt, ch, sv.data());
      cta_hb_vdots(c, hb, 0, p, j);                       // runs beside the GEMV on the device
    }
    cta_hb_panel_step(c, hb, 0, p, HB_NB, red.data(), sb.data(), sw.data(), st.data(), scv.data());
    for (int ti = 0; ti < tiles; ++ti) cta_hb_gemm<HB_YTOP, false>(c, hb, 0, p, ti, 0, smem.data());
    for (int r = 0; r < n; ++r) cta_hb_ytop_T(c, hb, 0, p, r);
    for (int ti = 0; ti < tiles; ++ti)
      for (int tj = 0; tj < tiles; ++tj) cta_hb_gemm<HB_RIGHT_TRAIL, false>(c, hb, 0, p, ti, tj, smem.data());
    for (int ti = 0; ti < tiles; ++ti) cta_hb_gemm<HB_RIGHT_PANEL, false>(c, hb, 0, p, ti, 0, smem.data());
    for (int tj = 0; tj < tiles; ++tj) cta_hb_gemm<HB_LEFT_W, false>(c, hb, 0, p, 0, tj, smem.data());
    {
      const int k = ilo + p * HB_NB;
      if (k < ihi)
        for (int j = 0; j < n - (k + HB_NB); ++j) cta_hb_w_T(c, T.data() + (size_t)p * HB_NB * HB_NB, W.data(), n - (k + HB_NB), j, true);
    }
    for (int ti = 0; ti < tiles; ++ti)
      for (int tj = 0; tj < tiles; ++tj) cta_hb_gemm<HB_LEFT_UPD, false>(c, hb, 0, p, ti, tj, smem.data());
  }
  if (Tout) std::memcpy(Tout, T.data(), sizeof(cplx) * T.size());
  return P;
}
