// stabgpu.cu -- C-ABI implementation (include/stabgpu.h): device-memory plans, stage launches,
// host<->device staging.  No CPU fallback: every compute entry point requires a CUDA device.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include "../../include/stabgpu.h"
#include "kernels.cuh"
#include "lu_blocked.cuh"
#include "polish.cuh"

using namespace stab;

namespace {

thread_local std::string g_err;
int g_device = -1;            // primary device (plans, single-matrix entry points)
int g_sm_count = 148;
bool g_inited = false;

// Devices the batch entry points shard over (stabgpu_init: one; stabgpu_init_multi: up to all of the box).  Every
// device keeps its own cached plan and its own pinned staging ring; a batch call runs one host worker thread per device.
struct StageRing {
  static constexpr int NSLOT = 4;
  size_t slot_bytes = 0;
  char* buf[NSLOT] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev[NSLOT] = {nullptr, nullptr, nullptr, nullptr};
};
struct DevCtx { int device = 0; stabgpu_plan* cached = nullptr; StageRing ring; };
std::vector<DevCtx> g_devs;
int g_stage_threads = 0;      // host threads that copy one staged chunk into the caller's (pageable) array; 0: host cores / devices, 2..8
                              // (296 points with vectors, 16 host cores: 4 threads 970, 8 threads 987 eigensolves/s)
int g_pin_mode = 1;           // 1: pageable destinations go through the pinned staging ring; 0: plain cudaMemcpyAsync into them
struct Tuning { int W = 64, ns = 16, qr_threads = 256, hess_threads = 512; int qr_steps = 32;   /* two CTAs per SM: 2 x 89 KB, 128 registers */ int qr_nw = -1, qr_nibble = 14; /* deflation window of the QR kernel (0: classic deflation only; -1: by order, 32 up to 640 and 44 above: 174 -> 167 ms per 148 matrices at N = 1280) and ZLAQR0's NIBBLE */ int hess_streams = 1; int hess_graph = 1; /* replay the Hessenberg stage as one CUDA graph (0: individual launches) */ int evec_mode = 1; int lu_mode = 1; /* 1: blocked LU with DMMA updates, 0: v1 one-CTA kernel */ int hess_mode = 1; /* 0: v1 unblocked CTA kernel, 1: batched blocked + DMMA, 2: blocked, scalar GEMM */ } g_tune;

int fail(const std::string& m) { g_err = m; return 1; }
bool g_qrprof_on = false;
long long* g_qrprof_dev = nullptr;

#define CU(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      g_err = std::string(#call) + ": " + cudaGetErrorString(e_) + " (" __FILE__ ":" + std::to_string(__LINE__) + ")"; \
      return 1;                                                                                    \
    }                                                                                              \
  } while (0)

int ensure_init() {
  if (g_inited) return cudaSetDevice(g_device) == cudaSuccess ? 0 : fail("libstabgpu: cudaSetDevice failed");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail("libstabgpu: no CUDA device available (there is no CPU fallback for the hot path)");
  if (g_device < 0) {
    int cur = 0;
    CU(cudaGetDevice(&cur));
    g_device = cur;
  }
  if (g_device >= ndev) return fail("libstabgpu: device index out of range");
  CU(cudaSetDevice(g_device));
  { cudaDeviceProp pr; if (cudaGetDeviceProperties(&pr, g_device) == cudaSuccess && pr.multiProcessorCount > 0) g_sm_count = pr.multiProcessorCount; }
  if (g_devs.empty()) { DevCtx d; d.device = g_device; g_devs.push_back(d); }
  g_inited = true;
  return 0;
}

template <class T>
struct DBuf {
  T* p = nullptr;
  size_t n = 0;
  int alloc(size_t count) {
    release();
    if (count == 0) return 0;
    cudaError_t e = cudaMalloc(&p, count * sizeof(T));
    if (e != cudaSuccess) { p = nullptr; g_err = std::string("cudaMalloc failed: ") + cudaGetErrorString(e); return 1; }
    n = count;
    return 0;
  }
  void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
  ~DBuf() { release(); }
};

Phys phys_from(const stabgpu_params* p) {
  Phys q;
  q.Ma = p->Ma; q.Re = p->Re; q.Pr = p->Pr; q.gamma = p->gamma; q.gamma1 = p->gamma1; q.cp = p->cp;
  q.Te = p->Te; q.rmue = p->rmue; q.rlme = p->rlme; q.cone = p->cone;
  for (int k = 0; k < 3; ++k) q.datmat[k] = p->datmat[k];
  q.mattyp = p->mattyp;
  q.navier = !(p->Re >= 1.0e98 || p->Re == 0.0);
  return q;
}

enum Stage { ST_ASM = 0, ST_LU, ST_BAL, ST_HESS, ST_PREP, ST_QR, ST_SORT, ST_EVEC, ST_N };

}  // namespace

struct stabgpu_plan {
  int device = 0;          // the CUDA device that owns every buffer, stream and event of this plan
  int kind = 0;            // 1 temporal, 2 spatial, 3 generic matrices
  stabgpu_params prm;
  int ny = 0, n = 0, N = 0; // n = 5 ny, N = order of the eigenproblem (n or 2n)
  int cap = 0, npts = 0;
  bool cap_limited = false;   // cap was clipped by device memory
  int want_vectors = 0;
  cudaStream_t stream = nullptr, stream2 = nullptr;
  cudaEvent_t evFork = nullptr, evJoin = nullptr;
  cudaEvent_t evSub[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  double* evec_host = nullptr;            // when set (batch C-ABI calls): eigenvectors are copied to the host per sub-batch, overlapped
  bool evec_staged = false;               // the destination is pageable: run_eigvecs only records the sub-batch events, the
                                          // batch driver drains the vectors through the device's pinned staging ring
  int nsub = 0, sub_m0[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};   // sub-batch boundaries of the last eigenvector stage
  std::vector<cudaEvent_t> evA, evB;
  bool stop_after_lu = false;             // debug: assembly + LU reduce only (stabgpu_debug_spatial_reduce)
  bool prof_hess = false;                 // record events around every Hessenberg kernel class (bench breakdown)
  cudaGraphExec_t hess_graph = nullptr;   // the ~1400 launches of the blocked reduction, captured on the first execute
  int hess_graph_mode = -1; long long hess_graph_launches = 0;
  std::vector<cudaEvent_t> pev; size_t pev_n = 0; std::vector<int> pev_cls;
  float hess_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // panel_step, gemv, gemm, other, invit, back-transformation GEMM, finalize, -
  // grid / profile
  DBuf<double> vm, g2, g22, deta, d2eta, D1, D2, Dt2w, h5;
  bool has_h5 = false;
  // sweep values
  DBuf<cplx> s1, s2;
  DBuf<double> Re, Ma;
  bool has_Re = false, has_Ma = false;
  // work
  DBuf<cplx> coef, A, C, Hq, V, tau, w, eig, lam;
  DBuf<cplx> hbY, hbT, hbYp, hbW, hbVx, hbS, hbVh, hbTv, hbVT, hbVTh;   // blocked Hessenberg workspaces (Vx: the panel's V with explicit ones / zeros)
  int hbP = 0;
  DBuf<double> scale, hnorm;
  DBuf<int> cnt, ilohi, info_lu, info_qr, info_v, blkend, kr, vbad, lu_perm;
  cudaEvent_t ev[ST_N + 1] = {};
  float ms[ST_N] = {};
  long long launches = 0;
  GridDev grid() const {
    GridDev g;
    g.ny = ny; g.wallt = prm.wallt; g.top = prm.top;
    g.D1 = D1.p; g.D2 = D2.p; g.Dt2w = Dt2w.p; g.deta = deta.p; g.d2eta = d2eta.p;
    g.vm = vm.p; g.g2vm = g2.p; g.g22vm = g22.p; g.h5 = has_h5 ? h5.p : nullptr;
    return g;
  }
};

namespace {

size_t per_point_bytes(int kind, int n, int N, int ny, int want_vectors) {
  size_t b = 0;
  b += (size_t)N * N * 16;                           // A (operand of the eigensolve)
  if (kind == 2) b += (size_t)n * n * 16;            // C0
  if (want_vectors) b += 2 * (size_t)N * N * 16;     // Hq + V
  b += (size_t)ny * 150 * 16;                        // coefficients
  b += (size_t)N * (16 * 4 + 8 + 4 * 3) + 64;
  b += (size_t)N * 16 * (7 * HB_NB + HB_CHUNKS) + 3 * 16 * HB_NB * HB_NB;   // blocked Hessenberg: Y, W, Vx, Vh, T, S, Ypart
  return b;
}

int plan_alloc(stabgpu_plan* pl, int max_pts) {
  size_t freeb = 0, totalb = 0;
  CU(cudaMemGetInfo(&freeb, &totalb));
  size_t ppb = per_point_bytes(pl->kind, pl->n, pl->N, pl->ny, pl->want_vectors);
  size_t capmem = (size_t)(0.85 * (double)freeb) / ppb;
  int cap = max_pts;
  if ((size_t)cap > capmem) cap = (int)capmem;
  if (cap > 65535) cap = 65535;
  if (cap < 1) return fail("libstabgpu: not enough device memory for a single point");
  pl->cap_limited = cap < max_pts;
  pl->cap = cap;
  const int N = pl->N, n = pl->n, ny = pl->ny;
  if (pl->s1.alloc(cap) || pl->s2.alloc(cap) || pl->Re.alloc(cap) || pl->Ma.alloc(cap)) return 1;
  if (pl->kind != 3 && pl->coef.alloc((size_t)cap * ny * (pl->kind == 1 ? 75 : 150))) return 1;
  if (pl->A.alloc((size_t)cap * N * N)) return 1;
  if (pl->kind == 2 && (pl->C.alloc((size_t)cap * n * n) || pl->lu_perm.alloc((size_t)cap * LU_PERM))) return 1;
  if (pl->want_vectors && (pl->Hq.alloc((size_t)cap * N * N) || pl->V.alloc((size_t)cap * N * N))) return 1;
  if (pl->tau.alloc((size_t)cap * N) || pl->w.alloc((size_t)cap * N) || pl->eig.alloc((size_t)cap * N) || pl->lam.alloc((size_t)cap * N)) return 1;
  if (pl->scale.alloc((size_t)cap * N) || pl->hnorm.alloc(cap) || pl->cnt.alloc((size_t)cap * N)) return 1;
  if (pl->blkend.alloc((size_t)cap * N) || pl->kr.alloc((size_t)cap * N) || pl->vbad.alloc((size_t)cap * N)) return 1;
  pl->hbP = (N - 1 + HB_NB - 1) / HB_NB;
  if (pl->hbY.alloc((size_t)cap * N * HB_NB) || pl->hbT.alloc((size_t)cap * pl->hbP * HB_NB * HB_NB) ||
      pl->hbYp.alloc((size_t)cap * N * HB_CHUNKS) || pl->hbW.alloc((size_t)cap * N * HB_NB) || pl->hbVx.alloc((size_t)cap * N * HB_NB) ||
      pl->hbS.alloc((size_t)cap * HB_NB * HB_NB) || pl->hbVh.alloc((size_t)cap * N * HB_NB) || pl->hbTv.alloc((size_t)cap * HB_NB) ||
      pl->hbVT.alloc((size_t)cap * N * HB_NB) || pl->hbVTh.alloc((size_t)cap * N * HB_NB)) return 1;
  if (pl->ilohi.alloc(2 * (size_t)cap) || pl->info_lu.alloc(cap) || pl->info_qr.alloc(cap) || pl->info_v.alloc(cap)) return 1;
  CU(cudaStreamCreate(&pl->stream));
  CU(cudaStreamCreate(&pl->stream2));
  CU(cudaEventCreateWithFlags(&pl->evFork, cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&pl->evJoin, cudaEventDisableTiming));
  pl->evA.resize(pl->hbP); pl->evB.resize(pl->hbP);
  for (int i = 0; i < pl->hbP; ++i) {
    CU(cudaEventCreateWithFlags(&pl->evA[i], cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&pl->evB[i], cudaEventDisableTiming));
  }
  for (int i = 0; i <= ST_N; ++i) CU(cudaEventCreate(&pl->ev[i]));
  return 0;
}

template <int PHASE>
int launch_hb_gemm(stabgpu_plan* pl, const HessBatch& hb, int nmat, cudaStream_t s, int panel, int ti, int tj, size_t smem, bool mma) {
  dim3 grid(ti, tj, nmat);
  if (mma) {
    CU(cudaFuncSetAttribute(k_hb_gemm<PHASE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_hb_gemm<PHASE, true><<<grid, GEMM_THREADS, smem, s>>>(hb, panel);
  } else {
    CU(cudaFuncSetAttribute(k_hb_gemm<PHASE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_hb_gemm<PHASE, false><<<grid, GEMM_THREADS, smem, s>>>(hb, panel);
  }
  CU(cudaGetLastError());
  pl->launches += 1;
  return 0;
}


// pipelined DMMA GEMM (gemm_pipe.cuh): persistent CTAs, two per SM
template <int PHASE>
int launch_pipe(stabgpu_plan* pl, const HessBatch& hb, cplx* X, size_t xstride, int nmat, cudaStream_t s, int panel, int ti, int tj) {
  if (ti < 1 || tj < 1 || nmat < 1) return 0;
  const size_t smem = (PHASE == PP_LEFT_W || PHASE == PP_BT_W) ? PipeCfg<32, 64>::smem_bytes : PipeCfg<64, 32>::smem_bytes;
  CU(cudaFuncSetAttribute(k_gemm_pipe<PHASE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  long long total = (long long)ti * tj * nmat;
  int grid = 2 * g_sm_count; if ((long long)grid > total) grid = (int)total;
  k_gemm_pipe<PHASE><<<grid, GEMM_THREADS, smem, s>>>(hb, X, xstride, panel, ti, tj, nmat);
  CU(cudaGetLastError());
  pl->launches += 1;
  return 0;
}
int launch_vx(stabgpu_plan* pl, const HessBatch& hb, int nmat, cudaStream_t s, int panel, int want = 0) {
  k_hb_vx<<<dim3((hb.n + 127) / 128, nmat), 128, 0, s>>>(hb, panel, want);
  CU(cudaGetLastError());
  pl->launches += 1;
  return 0;
}

// profile marks: an event after each kernel class (0 panel_step, 1 gemv, 2 gemm/other level-3, 3 start marker)
static int hmark(stabgpu_plan* pl, cudaStream_t s, int cls) {
  if (!pl->prof_hess) return 0;
  if (pl->pev_n == pl->pev.size()) { cudaEvent_t e; CU(cudaEventCreate(&e)); pl->pev.push_back(e); pl->pev_cls.push_back(cls); }
  pl->pev_cls[pl->pev_n] = cls;
  CU(cudaEventRecord(pl->pev[pl->pev_n++], s));
  return 0;
}

// one panel of the blocked reduction for the matrices [hb.mat0, hb.mat0 + nmat) on stream s;
// phase 1: the column loop (panel steps + HBM-bound GEMVs), phase 2: the tensor-core block updates
int hess_panel(stabgpu_plan* pl, const HessBatch& hb, int nmat, cudaStream_t s, int p, bool mma, int phase) {
  const int N = pl->N;
  const size_t sm_step = 160 * sizeof(double) + ((size_t)N + 3 * HB_NB) * sizeof(cplx);
  const int pst = g_tune.hess_threads;                      // threads of the panel-step CTA
  const size_t sm_gemv = (size_t)N * sizeof(cplx);
  const size_t sm64 = GemmCfg<64, 64>::smem_bytes, sm6432 = GemmCfg<64, 32>::smem_bytes, sm3264 = GemmCfg<32, 64>::smem_bytes;
  const int k0 = p * HB_NB;                                  // smallest possible panel start (ilo = 0)
  const int rows_max = N - 1 - k0;                           // rows k+1..ihi
  if (rows_max <= 0) return 0;
  const int trail_max = N - (k0 + HB_NB);                    // columns k+NB..n-1
  if (phase == 1) {
    dim3 ggemv((rows_max + HB_GEMV_ROWS - 1) / HB_GEMV_ROWS, HB_CHUNKS + 1, nmat);   // + the V^H v dot products of the same column
    for (int j = 0; j < HB_NB; ++j) {
      k_hb_panel_step<<<nmat, pst, sm_step, s>>>(hb, p, j);
      if (hmark(pl, s, 0)) return 1;
      k_hb_gemv<<<ggemv, HB_GEMV_ROWS, sm_gemv, s>>>(hb, p, j);
      if (hmark(pl, s, 1)) return 1;
    }
    k_hb_panel_step<<<nmat, pst, sm_step, s>>>(hb, p, HB_NB);
    if (hmark(pl, s, 0)) return 1;
    CU(cudaGetLastError());
    pl->launches += 2 * HB_NB + 1;
    return 0;
  }
  const int tm = (N + 63) / 64;
  if (g_tune.hess_mode == 5) {               // as mode 1, with the right and left trailing updates fused into one rank-64 pass
    if (launch_vx(pl, hb, nmat, s, p)) return 1;
    if (launch_pipe<PP_YTOP>(pl, hb, nullptr, 0, nmat, s, p, tm, 1)) return 1;
    k_hb_ytop_T<<<dim3((N + 127) / 128, nmat), 128, 0, s>>>(hb, p);
    pl->launches += 1;
    if (trail_max > 0) {
      if (launch_pipe<PP_S>(pl, hb, nullptr, 0, nmat, s, p, 1, 1)) return 1;
      if (launch_pipe<PP_LEFT_W>(pl, hb, nullptr, 0, nmat, s, p, 1, (trail_max + 63) / 64)) return 1;     // on the matrix BEFORE the right update
      k_hb_w_T_fused<<<dim3((trail_max + 127) / 128, nmat), 128, 0, s>>>(hb, p);
      pl->launches += 1;
      if (launch_pipe<PP_RIGHT_TOP>(pl, hb, nullptr, 0, nmat, s, p, tm, (trail_max + 31) / 32)) return 1;
    }
    if (launch_pipe<PP_RIGHT_PANEL>(pl, hb, nullptr, 0, nmat, s, p, tm, 1)) return 1;
    if (trail_max > 0 && launch_pipe<PP_FUSED_UPD>(pl, hb, nullptr, 0, nmat, s, p, (rows_max + 63) / 64, (trail_max + 31) / 32)) return 1;
    if (hmark(pl, s, 2)) return 1;
    CU(cudaGetLastError());
    return 0;
  }
  if (g_tune.hess_mode == 1) {               // pipelined path: plain operands (V, V T, V T^H materialised), persistent cp.async ring
    if (launch_vx(pl, hb, nmat, s, p, 3)) return 1;
    if (launch_pipe<PP_YTOP>(pl, hb, nullptr, 0, nmat, s, p, tm, 1)) return 1;                 // Y_top = A_top (V T)
    if (trail_max > 0 && launch_pipe<PP_RIGHT_TRAIL>(pl, hb, nullptr, 0, nmat, s, p, tm, (trail_max + 31) / 32)) return 1;
    if (launch_pipe<PP_RIGHT_PANEL>(pl, hb, nullptr, 0, nmat, s, p, tm, 1)) return 1;
    if (trail_max > 0) {
      if (launch_pipe<PP_LEFT_W>(pl, hb, nullptr, 0, nmat, s, p, 1, (trail_max + 63) / 64)) return 1;   // W = V^H A
      if (launch_pipe<PP_LEFT_UPD>(pl, hb, nullptr, 0, nmat, s, p, (rows_max + 63) / 64, (trail_max + 31) / 32)) return 1;   // A -= (V T^H) W
    }
    if (hmark(pl, s, 2)) return 1;
    CU(cudaGetLastError());
    return 0;
  }
  if (launch_hb_gemm<HB_YTOP>(pl, hb, nmat, s, p, tm, 1, sm6432, mma)) return 1;
  k_hb_ytop_T<<<dim3((N + 127) / 128, nmat), 128, 0, s>>>(hb, p);
  pl->launches += 1;
  if (trail_max > 0) {
    const int tn = (trail_max + 63) / 64;
    if (launch_hb_gemm<HB_RIGHT_TRAIL>(pl, hb, nmat, s, p, tm, tn, sm64, mma)) return 1;
  }
  if (launch_hb_gemm<HB_RIGHT_PANEL>(pl, hb, nmat, s, p, tm, 1, sm6432, mma)) return 1;
  if (trail_max > 0) {
    const int tn = (trail_max + 63) / 64;
    if (launch_hb_gemm<HB_LEFT_W>(pl, hb, nmat, s, p, 1, tn, sm3264, mma)) return 1;
    k_hb_w_T<<<dim3((trail_max + 127) / 128, nmat), 128, 0, s>>>(hb, p);
    pl->launches += 1;
    if (launch_hb_gemm<HB_LEFT_UPD>(pl, hb, nmat, s, p, (rows_max + 63) / 64, tn, sm64, mma)) return 1;
  }
  if (hmark(pl, s, 2)) return 1;
  CU(cudaGetLastError());
  return 0;
}

// Stage 3b: A <- Hessenberg form + reflectors (ZGEHRD layout), tau, and the panel factors T.
// The batch is split in two halves that run the same kernel schedule on two streams: the
// HBM-bound GEMV of one half overlaps the latency-bound panel step / tensor-core updates of the other.
int run_hessenberg(stabgpu_plan* pl) {
  const int N = pl->N, np = pl->npts;
  const size_t st = (size_t)N * N;
  cudaStream_t s = pl->stream;
  if (g_tune.hess_mode == 0) {
    size_t sm = 160 * sizeof(double) + 2 * (size_t)N * sizeof(cplx);
    CU(cudaFuncSetAttribute(k_hessenberg, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    k_hessenberg<<<np, g_tune.hess_threads, sm, s>>>(pl->A.p, st, N, pl->ilohi.p, pl->tau.p);
    CU(cudaGetLastError());
    pl->launches += 1;
    return 0;
  }
  const bool mma = g_tune.hess_mode == 1 || g_tune.hess_mode == 3;   // 1: pipelined DMMA kernels, 3: the tile-per-CTA DMMA kernels
  HessBatch hb{pl->A.p, st, N, pl->ilohi.p, pl->tau.p, pl->hbY.p, pl->hbT.p, pl->hbYp.p, pl->hbW.p, pl->hbP, 0, pl->hbVx.p,
               g_tune.hess_mode == 1 ? pl->hbVT.p : nullptr, g_tune.hess_mode == 1 ? pl->hbVTh.p : nullptr, pl->hbTv.p, pl->hbS.p, pl->hbVh.p};
  pl->pev_n = 0;
  if (hmark(pl, s, 3)) return 1;
  const int half = (g_tune.hess_streams >= 2 && np >= 16) ? (np + 1) / 2 : np;
  HessBatch hb2 = hb; hb2.mat0 = half;
  if (half < np) { CU(cudaEventRecord(pl->evFork, s)); CU(cudaStreamWaitEvent(pl->stream2, pl->evFork, 0)); }
  // software pipeline over the two halves: the GEMV phases (memory bound) of A and B never overlap each
  // other, each overlaps the tensor-core phase of the other half
  const bool two = half < np;
  // One stream, no profiling marks: the schedule depends only on (N, npts, hess_mode) -- every data-dependent quantity
  // (ilo/ihi, the panel extents) is read on the device -- so it is captured once per plan and replayed as a CUDA graph.
  const bool graphed = g_tune.hess_graph && !two && !pl->prof_hess;
  if (graphed && pl->hess_graph && pl->hess_graph_mode == g_tune.hess_mode * 65536 + g_tune.hess_threads) {
    CU(cudaGraphLaunch(pl->hess_graph, s));
    pl->launches += pl->hess_graph_launches;
    return 0;
  }
  const long long launches0 = pl->launches;
  if (graphed) {
    if (pl->hess_graph) { cudaGraphExecDestroy(pl->hess_graph); pl->hess_graph = nullptr; }
    CU(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
  }
  int rc = 0;
  for (int p = 0; p < pl->hbP && !rc; ++p) {
    if (!two) {
      rc = hess_panel(pl, hb, half, s, p, mma, 1) || hess_panel(pl, hb, half, s, p, mma, 2);
      continue;
    }
    if (two && p > 0) CU(cudaStreamWaitEvent(s, pl->evB[p - 1], 0));
    if (hess_panel(pl, hb, half, s, p, mma, 1)) return 1;
    if (two) {
      CU(cudaEventRecord(pl->evA[p], s));
      CU(cudaStreamWaitEvent(pl->stream2, pl->evA[p], 0));
    }
    if (hess_panel(pl, hb, half, s, p, mma, 2)) return 1;
    if (two) {
      if (hess_panel(pl, hb2, np - half, pl->stream2, p, mma, 1)) return 1;
      CU(cudaEventRecord(pl->evB[p], pl->stream2));
      if (hess_panel(pl, hb2, np - half, pl->stream2, p, mma, 2)) return 1;
    }
  }
  if (graphed) {
    cudaGraph_t gr = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(s, &gr);
    if (rc || ce != cudaSuccess) { if (gr) cudaGraphDestroy(gr); return rc ? rc : fail(cudaGetErrorString(ce)); }
    const cudaError_t ci = cudaGraphInstantiate(&pl->hess_graph, gr, 0);
    cudaGraphDestroy(gr);
    if (ci != cudaSuccess) { pl->hess_graph = nullptr; return fail(cudaGetErrorString(ci)); }
    pl->hess_graph_mode = g_tune.hess_mode * 65536 + g_tune.hess_threads;
    pl->hess_graph_launches = pl->launches - launches0;
    CU(cudaGraphLaunch(pl->hess_graph, s));
    return 0;
  }
  if (rc) return rc;
  if (half < np) { CU(cudaEventRecord(pl->evJoin, pl->stream2)); CU(cudaStreamWaitEvent(s, pl->evJoin, 0)); }
  return 0;
}

template <int PHASE>
int launch_bt_gemm(stabgpu_plan* pl, const HessBatch& hb, int nmat, int panel, int ti, int tj, size_t smem) {
  dim3 grid(ti, tj, nmat);
  CU(cudaFuncSetAttribute(k_bt_gemm<PHASE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_bt_gemm<PHASE, true><<<grid, GEMM_THREADS, smem, pl->stream>>>(hb, pl->V.p, (size_t)pl->N * pl->N, panel);
  CU(cudaGetLastError());
  pl->launches += 1;
  return 0;
}

// Stage 6: right eigenvectors.  evec_mode 1 (default, N <= 1280 and blocked Hessenberg factors available):
// register-resident inverse iteration -> GEMM back-transformation -> finalize; otherwise the v1 warp kernel.
// The batch is processed in sub-batches; when the caller registered a host destination (the batch C-ABI calls), the
// finished vectors of sub-batch i travel to the host on the copy stream while sub-batch i+1 is computed.
int run_eigvecs(stabgpu_plan* pl, int scale_rows) {
  const int N = pl->N, np = pl->npts;
  const size_t st = (size_t)N * N;
  cudaStream_t s = pl->stream;
  CU(cudaMemsetAsync(pl->info_v.p, 0, sizeof(int) * np, s));
  int warps = 8;
  const size_t per_warp = 2 * (size_t)N * sizeof(cplx) + (size_t)N;
  while (warps > 1 && warps * per_warp > 200 * 1024) warps >>= 1;
  const size_t sm_old = warps * per_warp;
  if (sm_old > 227 * 1024) return fail("libstabgpu: matrix too large for the eigenvector kernel");
  CU(cudaFuncSetAttribute(k_evec, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_old));
  const bool fast = (g_tune.evec_mode == 1 || g_tune.evec_mode == 3) && g_tune.hess_mode != 0 && N <= 1280;   // 3: per-step inverse iteration (validation)
  const bool to_host = pl->evec_host != nullptr;
  const int nsub = !to_host ? 1 : (np >= 8 * 37 ? 8 : (np >= 4 * 37 ? 4 : (np >= 2 ? 2 : 1)));   // >= 37 matrices per sub-batch keep the kernels full
  pl->nsub = nsub;
  for (int sb = 0; sb <= nsub; ++sb) pl->sub_m0[sb] = (int)((long long)np * sb / nsub);
  if (hmark(pl, s, 7)) return 1;               // start marker of the eigenvector breakdown (class 7: not reported)
  for (int sb = 0; sb < nsub; ++sb) {
    const int m0 = pl->sub_m0[sb], m1 = pl->sub_m0[sb + 1], cnt = m1 - m0;
    if (cnt <= 0) continue;
    int chunks = 1;
    while (chunks * cnt < 2 * 148 && chunks * warps < N) chunks *= 2;
    dim3 grid_old(chunks, cnt);
    if (!fast) {
      // the reflectors live in A (ZGEHRD layout); the QR ran on the copy Hq
      k_evec<<<grid_old, warps * 32, sm_old, s>>>(pl->A.p + (size_t)m0 * st, st, N, pl->ilohi.p + 2 * m0, pl->tau.p + (size_t)m0 * N,
                                                  pl->scale.p + (size_t)m0 * N, pl->lam.p + (size_t)m0 * N, pl->kr.p + (size_t)m0 * N,
                                                  pl->hnorm.p + m0, scale_rows, pl->V.p + (size_t)m0 * st, st, pl->info_v.p + m0, nullptr, 0);
      CU(cudaGetLastError());
      pl->launches += 1;
    } else {
      const int rounds = 4;
      CU(launch_invit(pl->A.p + (size_t)m0 * st, st, N, pl->lam.p + (size_t)m0 * N, pl->kr.p + (size_t)m0 * N, pl->hnorm.p + m0,
                      pl->V.p + (size_t)m0 * st, st, pl->vbad.p + (size_t)m0 * N, rounds, cnt, g_tune.evec_mode == 3, s));
      pl->launches += 1;
      // vectors the fast kernel rejected (no growth / overflow): ZLAEIN's retry vectors, v1 kernel, Hessenberg basis
      k_evec<<<grid_old, warps * 32, sm_old, s>>>(pl->A.p + (size_t)m0 * st, st, N, pl->ilohi.p + 2 * m0, pl->tau.p + (size_t)m0 * N,
                                                  pl->scale.p + (size_t)m0 * N, pl->lam.p + (size_t)m0 * N, pl->kr.p + (size_t)m0 * N,
                                                  pl->hnorm.p + m0, 0, pl->V.p + (size_t)m0 * st, st, pl->info_v.p + m0,
                                                  pl->vbad.p + (size_t)m0 * N, 1);
      CU(cudaGetLastError());
      pl->launches += 1;
      if (hmark(pl, s, 4)) return 1;
      HessBatch hb{pl->A.p, st, N, pl->ilohi.p, pl->tau.p, pl->hbY.p, pl->hbT.p, pl->hbYp.p, pl->hbW.p, pl->hbP, m0, pl->hbVx.p,
                   (g_tune.hess_mode == 1 || g_tune.hess_mode == 5) ? pl->hbVT.p : nullptr};
      const int tn = (N + 63) / 64;
      for (int p = pl->hbP - 1; p >= 0; --p) {
        const int rows_max = N - 1 - p * HB_NB;
        if (rows_max <= 0) continue;
        if (g_tune.hess_mode == 1 || g_tune.hess_mode == 5) {
          if (launch_vx(pl, hb, cnt, s, p, 1)) return 1;                                                   // V and V T
          if (launch_pipe<PP_BT_W>(pl, hb, pl->V.p, st, cnt, s, p, 1, tn)) return 1;                        // W = V^H X
          if (launch_pipe<PP_BT_UPD>(pl, hb, pl->V.p, st, cnt, s, p, (rows_max + 63) / 64, (N + 31) / 32)) return 1;   // X -= (V T) W
          continue;
        }
        if (launch_bt_gemm<BT_W>(pl, hb, cnt, p, 1, tn, GemmCfg<32, 64>::smem_bytes)) return 1;
        k_bt_w_T<<<dim3((N + 127) / 128, cnt), 128, 0, s>>>(hb, p);
        pl->launches += 1;
        if (launch_bt_gemm<BT_UPD>(pl, hb, cnt, p, (rows_max + 63) / 64, tn, GemmCfg<64, 64>::smem_bytes)) return 1;
      }
      if (hmark(pl, s, 5)) return 1;
      {
        const size_t smf = 8 * (size_t)N * sizeof(cplx);
        CU(cudaFuncSetAttribute(k_vec_finalize, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smf));
        int ch = 1;
        while (ch * cnt < 4 * 148 && ch * 8 < N) ch *= 2;
        k_vec_finalize<<<dim3(ch, cnt), 256, smf, s>>>(pl->V.p + (size_t)m0 * st, st, N, pl->ilohi.p + 2 * m0, pl->scale.p + (size_t)m0 * N, scale_rows);
        CU(cudaGetLastError());
        pl->launches += 1;
        if (hmark(pl, s, 6)) return 1;
      }
    }
    if (to_host) {                             // D2H of this sub-batch overlaps the next one
      if (!pl->evSub[sb]) CU(cudaEventCreateWithFlags(&pl->evSub[sb], cudaEventDisableTiming));
      CU(cudaEventRecord(pl->evSub[sb], s));
      if (!pl->evec_staged) {                  // pinned destination: straight into the caller's array
        CU(cudaStreamWaitEvent(pl->stream2, pl->evSub[sb], 0));
        CU(cudaMemcpyAsync(pl->evec_host + 2 * (size_t)m0 * st, pl->V.p + (size_t)m0 * st, sizeof(cplx) * (size_t)cnt * st,
                           cudaMemcpyDeviceToHost, pl->stream2));
      }
    }
  }
  if (to_host && !pl->evec_staged) {           // the compute stream's completion covers the copies as well
    CU(cudaEventRecord(pl->evJoin, pl->stream2));
    CU(cudaStreamWaitEvent(s, pl->evJoin, 0));
  }
  return 0;
}

// Batched blocked LU of lb.C with the lb.nrhs right-hand-side columns riding along (lu_blocked.cuh): factor + forward
// substitution, then the blocked back substitution.  nrhs = 0: factorization only (the polish path replays it).
int lu_run(const LuBatch& lb, int np, cudaStream_t s, long long* launches) {
  const int n = lb.n, N = lb.nrhs;
  CU(cudaMemsetAsync(lb.info, 0, sizeof(int) * np, s));
  const size_t sm_trsm = sizeof(cplx) * (LU_NB * LU_TRSM_THREADS + LU_NB * LU_NB);
  const size_t sm_gemm = PipeCfg<64, 32>::smem_bytes;
  CU(cudaFuncSetAttribute(k_lu_back_trsm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_trsm));
  CU(cudaFuncSetAttribute(k_lu_gemm<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_gemm));
  CU(cudaFuncSetAttribute(k_lu_gemm<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_gemm));
  auto gemm_grid = [&](long long total) { long long g = 2LL * g_sm_count; return (int)(g < total ? g : total); };
  for (int j0 = 0; j0 < n; j0 += LU_NB) {
    const int jb = (n - j0 < LU_NB) ? n - j0 : LU_NB, r0 = j0 + jb, ncols = (n - r0) + N;
    k_lu_panel<<<np, 512, 0, s>>>(lb, j0);
    *launches += 1;
    if (ncols > 0) {
      int cpc = 64;                                              // columns per CTA: enough CTAs to fill the GPU, few enough to amortise the L11 load
      while (cpc > LU_SWAP_WARPS && (long long)((ncols + cpc - 1) / cpc) * np < 4LL * g_sm_count) cpc >>= 1;
      k_lu_swap_trsm_warp<<<dim3((ncols + cpc - 1) / cpc, np), LU_SWAP_WARPS * 32, 0, s>>>(lb, j0, cpc);
      *launches += 1;
    }
    if (r0 < n) {
      const int wide = N > n - r0 ? N : n - r0;
      const int ti = (n - r0 + 63) / 64, tj = (wide + 31) / 32;
      k_lu_gemm<0><<<gemm_grid((long long)ti * tj * 2 * np), GEMM_THREADS, sm_gemm, s>>>(lb, j0, ti, tj, 2 * np);
      *launches += 1;
    }
    CU(cudaGetLastError());
  }
  if (N == 0) return 0;
  const int nblk = (n + LU_NB - 1) / LU_NB;
  for (int b = nblk - 1; b >= 0; --b) {
    const int i0 = b * LU_NB, bs = (n - i0 < LU_NB) ? n - i0 : LU_NB;
    k_lu_back_trsm<<<dim3((N + LU_TRSM_THREADS - 1) / LU_TRSM_THREADS, np), LU_TRSM_THREADS, sm_trsm, s>>>(lb, i0, bs);
    *launches += 1;
    if (i0 > 0) {
      const int ti = (i0 + 63) / 64, tj = (N + 31) / 32;
      k_lu_gemm<1><<<gemm_grid((long long)ti * tj * np), GEMM_THREADS, sm_gemm, s>>>(lb, i0, ti, tj, np);
      *launches += 1;
    }
    CU(cudaGetLastError());
  }
  return 0;
}

// stage 2 (spatial): blocked LU reduce  A(0:n, :) <- C0^-1 A(0:n, :)
int run_lu_blocked(stabgpu_plan* pl) {
  const int n = pl->n, N = pl->N;
  LuBatch lb{pl->C.p, (size_t)n * n, n, pl->A.p, (size_t)N * N, N, N, pl->cnt.p, pl->lu_perm.p, pl->info_lu.p};   // ipiv lives in the balancing stage's counter array
  return lu_run(lb, pl->npts, pl->stream, &pl->launches);
}

// the eigen-pipeline on pl->A (npts matrices of order N): balance -> Hessenberg -> QR -> sort [-> vectors]
int run_eigen(stabgpu_plan* pl, int sort_mode, int scale_rows) {
  const int N = pl->N, np = pl->npts;
  const size_t st = (size_t)N * N;
  cudaStream_t s = pl->stream;
  if (N > 640) {
    const int bb = balance_block_wide(N);
    const size_t smb = balance_wsp_doubles(N, bb) * sizeof(double);
    CU(cudaFuncSetAttribute(k_balance_wide, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smb));
    k_balance_wide<<<np, 512, smb, s>>>(pl->A.p, st, N, pl->scale.p, pl->cnt.p, pl->ilohi.p, bb);
  } else {
    const int bb = balance_block(N);
    const size_t smb = balance_wsp_doubles(N, bb) * sizeof(double);
    CU(cudaFuncSetAttribute(k_balance, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smb));
    k_balance<<<np, 256, smb, s>>>(pl->A.p, st, N, pl->scale.p, pl->cnt.p, pl->ilohi.p, bb);
  }
  CU(cudaGetLastError());
  CU(cudaEventRecord(pl->ev[ST_BAL + 1], s));
  if (run_hessenberg(pl)) return 1;
  CU(cudaEventRecord(pl->ev[ST_HESS + 1], s));
  cplx* Hq = pl->want_vectors ? pl->Hq.p : pl->A.p;
  k_prep_qr<<<np, 256, 0, s>>>(pl->A.p, st, Hq, st, N, pl->hnorm.p, pl->blkend.p);
  CU(cudaGetLastError());
  CU(cudaEventRecord(pl->ev[ST_PREP + 1], s));
  {
    HqrLaunch q; q.W = g_tune.W; q.ns_max = g_tune.ns; q.steps_max = g_tune.qr_steps; q.nw = g_tune.qr_nw >= 0 ? g_tune.qr_nw : (N > 640 ? 44 : 32); q.nibble = g_tune.qr_nibble;
    if (q.nw >= q.W || q.nw > 45 || (2 * q.nw + 1) * q.nw > q.W * (q.W + 1)) q.nw = 0;     // the window must fit the shared-memory tile
    if (q.W - 2 < 2 * q.ns_max - 1 || q.steps_max < 2 * q.ns_max - 1 || q.W - 2 * q.ns_max - 1 < 4)
      return fail("libstabgpu: invalid QR tuning (window too small for the shift count)");
    long long* prof = nullptr;
    if (g_qrprof_on) {
      if (!g_qrprof_dev) CU(cudaMalloc(&g_qrprof_dev, 16 * sizeof(long long)));
      CU(cudaMemsetAsync(g_qrprof_dev, 0, 16 * sizeof(long long), s));
      prof = g_qrprof_dev;
    }
    CU(launch_hqr(Hq, st, N, pl->ilohi.p, pl->w.p, pl->info_qr.p, q, prof, pl->hnorm.p, np, g_tune.qr_threads, s));
  }
  CU(cudaEventRecord(pl->ev[ST_QR + 1], s));
  k_sort<<<np, 256, 0, s>>>(pl->w.p, N, sort_mode, pl->hnorm.p, pl->blkend.p, pl->eig.p, pl->lam.p, pl->kr.p);
  CU(cudaGetLastError());
  CU(cudaEventRecord(pl->ev[ST_SORT + 1], s));
  pl->launches += 4;
  if (pl->want_vectors && run_eigvecs(pl, scale_rows)) return 1;
  CU(cudaEventRecord(pl->ev[ST_EVEC + 1], s));
  return 0;
}

int upload_grid(stabgpu_plan* pl, const stabgpu_params* p, const double* vm, const double* g2vm, const double* g22vm,
                const double* deta, const double* d2eta, const double* h5) {
  const int ny = p->ny;
  std::vector<double> D1((size_t)ny * ny), D2((size_t)ny * ny), Dt2w(ny), g2((size_t)ny * 5), g22((size_t)ny * 5);
  stabgpu_mean_gradients(ny, p->wallt, vm, deta, d2eta, D1.data(), D2.data(), Dt2w.data(), g2.data(), g22.data());
  if (!p->ider) {                                    // getmean2 path: analytic derivatives supplied (temporal.f90:99-103)
    if (!g2vm || !g22vm) return fail("libstabgpu: ider=0 requires g2vm and g22vm");
    std::memcpy(g2.data(), g2vm, sizeof(double) * ny * 5);
    std::memcpy(g22.data(), g22vm, sizeof(double) * ny * 5);
  }
  if (pl->vm.alloc((size_t)ny * 5) || pl->g2.alloc((size_t)ny * 5) || pl->g22.alloc((size_t)ny * 5) || pl->deta.alloc(ny) ||
      pl->d2eta.alloc(ny) || pl->D1.alloc((size_t)ny * ny) || pl->D2.alloc((size_t)ny * ny) || pl->Dt2w.alloc(ny))
    return 1;
  CU(cudaMemcpy(pl->vm.p, vm, sizeof(double) * ny * 5, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(pl->g2.p, g2.data(), sizeof(double) * ny * 5, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(pl->g22.p, g22.data(), sizeof(double) * ny * 5, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(pl->deta.p, deta, sizeof(double) * ny, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(pl->d2eta.p, d2eta, sizeof(double) * ny, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(pl->D1.p, D1.data(), sizeof(double) * ny * ny, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(pl->D2.p, D2.data(), sizeof(double) * ny * ny, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(pl->Dt2w.p, Dt2w.data(), sizeof(double) * ny, cudaMemcpyHostToDevice));
  pl->has_h5 = false;
  if (h5) {
    if (pl->h5.alloc((size_t)ny * 5)) return 1;
    CU(cudaMemcpy(pl->h5.p, h5, sizeof(double) * ny * 5, cudaMemcpyHostToDevice));
    pl->has_h5 = true;
  }
  return 0;
}

int check_params(const stabgpu_params* p) {
  if (!p) return fail("libstabgpu: null params");
  if (p->ny < 4 || p->ny > 512) return fail("libstabgpu: ny out of range [4,512]");
  if (p->wallt != 0 && p->wallt != 2) return fail("Illegal value of wallt");   // temporal.f90:659-661
  if (p->mattyp != 0 && p->mattyp != 1) return fail("libstabgpu: mattyp must be 0 or 1");
  if (p->curve == 1) return fail("libstabgpu: curve=1 (calch) is out of scope; supply metrics through h5");
  return 0;
}

}  // namespace

namespace {

void ring_release(StageRing& r) {
  for (int i = 0; i < StageRing::NSLOT; ++i) {
    if (r.buf[i]) cudaFreeHost(r.buf[i]);
    if (r.ev[i]) cudaEventDestroy(r.ev[i]);
    r.buf[i] = nullptr; r.ev[i] = nullptr;
  }
  r.slot_bytes = 0;
}

void release_devices() {                       // cached plans and staging rings of every device
  for (auto& d : g_devs) {
    cudaSetDevice(d.device);
    if (d.cached) { stabgpu_plan* pl = d.cached; d.cached = nullptr; stabgpu_plan_destroy(pl); }
    ring_release(d.ring);
  }
  g_devs.clear();
}

int use_plan(const stabgpu_plan* pl) {        // the current device is per host thread
  CU(cudaSetDevice(pl->device));
  return 0;
}

}  // namespace

extern "C" {

const char* stabgpu_last_error(void) { return g_err.c_str(); }

int stabgpu_init(int device) {
  if (g_inited) release_devices();             // a cached plan lives on the device it was created on
  g_inited = false;
  g_device = device;
  return ensure_init();
}

int stabgpu_init_multi(int max_devices, int* ndev_used) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail("libstabgpu: no CUDA device available (there is no CPU fallback for the hot path)");
  if (max_devices > 0 && max_devices < ndev) ndev = max_devices;
  if (g_inited) release_devices();
  g_inited = false;
  g_device = 0;
  if (ensure_init()) return 1;
  g_devs.clear();
  for (int d = 0; d < ndev; ++d) { DevCtx c; c.device = d; g_devs.push_back(c); }
  if (ndev_used) *ndev_used = ndev;
  return 0;
}

int stabgpu_device_count(void) { return g_inited ? (int)g_devs.size() : 0; }

int stabgpu_set_host_staging(int pin_mode, int copy_threads) {
  if (pin_mode == 0 || pin_mode == 1) g_pin_mode = pin_mode;
  if (copy_threads > 0) g_stage_threads = copy_threads > 32 ? 32 : copy_threads;
  return 0;
}

/* page-lock a caller array once (e.g. the Fortran evec array after its allocate) so that the batch calls copy into it
 * directly; the staging ring is used for any destination that is not page-locked */
int stabgpu_host_register(void* ptr, size_t bytes) {
  if (ensure_init()) return 1;
  CU(cudaHostRegister(ptr, bytes, cudaHostRegisterPortable));
  return 0;
}
int stabgpu_host_unregister(void* ptr) {
  CU(cudaHostUnregister(ptr));
  return 0;
}

int stabgpu_finalize(void) {
  if (g_inited) {
    for (auto& d : g_devs) { cudaSetDevice(d.device); cudaDeviceSynchronize(); }
    release_devices();
    cudaSetDevice(g_device);
  }
  g_inited = false;
  return 0;
}

int stabgpu_device_info(char* name, int name_len, int* sm_count, double* mem_gb) {
  if (ensure_init()) return 1;
  cudaDeviceProp pr;
  CU(cudaGetDeviceProperties(&pr, g_device));
  if (name && name_len > 0) { std::strncpy(name, pr.name, name_len - 1); name[name_len - 1] = 0; }
  if (sm_count) *sm_count = pr.multiProcessorCount;
  if (mem_gb) *mem_gb = (double)pr.totalGlobalMem / 1.0e9;
  return 0;
}

/* debug: cycle counters of the QR kernel for matrix 0 of the last run (not part of the public header) */
int stabgpu_debug_qr_profile(int enable, long long* out16) {
  g_qrprof_on = enable != 0;
  if (out16 && g_qrprof_dev) {
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy(out16, g_qrprof_dev, 16 * sizeof(long long), cudaMemcpyDeviceToHost));
  }
  return 0;
}

int stabgpu_set_hess_mode(int mode) { g_tune.hess_mode = mode; return 0; }
int stabgpu_debug_set_qr_steps(int steps) { if (steps > 0) g_tune.qr_steps = steps; return 0; }
int stabgpu_debug_set_qr_aed(int nw, int nibble) { if (nw >= 0) g_tune.qr_nw = nw; if (nibble >= 0) g_tune.qr_nibble = nibble; return 0; }
int stabgpu_set_qr_deflation(int window, int nibble) { return stabgpu_debug_set_qr_aed(window, nibble); }
int stabgpu_set_evec_mode(int mode) { g_tune.evec_mode = mode; return 0; }
int stabgpu_debug_set_hess_graph(int on) { g_tune.hess_graph = on ? 1 : 0; return 0; }
int stabgpu_set_lu_mode(int mode) { g_tune.lu_mode = mode; return 0; }

int stabgpu_set_tuning(int qr_window, int qr_shifts, int qr_threads, int hess_threads) {
  if (qr_window > 0) g_tune.W = qr_window;
  if (qr_shifts > 0) g_tune.ns = qr_shifts;
  if (qr_threads > 0) g_tune.qr_threads = qr_threads;
  if (hess_threads > 0) {
    if (hess_threads > 512 || hess_threads % 32) return fail("libstabgpu: hess_threads must be a multiple of 32 up to 512 (the panel-step kernel's launch bound)");
    g_tune.hess_threads = hess_threads;
  }
  return 0;
}

// plan on the calling thread's CURRENT device (the batch workers set their own before they get here)
static int plan_create_here(stabgpu_plan** out, int kind, const stabgpu_params* p, const double* vm, const double* g2vm,
                            const double* g22vm, const double* deta, const double* d2eta, const double* h5, int max_pts,
                            int want_vectors) {
  if (kind != 1 && kind != 2) return fail("libstabgpu: plan kind must be 1 (temporal) or 2 (spatial)");
  if (check_params(p)) return 1;
  if (!vm || !deta || !d2eta || max_pts < 1) return fail("libstabgpu: bad argument");
  stabgpu_plan* pl = new stabgpu_plan();
  CU(cudaGetDevice(&pl->device));              // the calling thread's current device (a batch worker sets its own)
  pl->kind = kind; pl->prm = *p; pl->ny = p->ny; pl->n = 5 * p->ny; pl->N = (kind == 1 ? 1 : 2) * pl->n;
  pl->want_vectors = want_vectors ? 1 : 0;
  if (upload_grid(pl, p, vm, g2vm, g22vm, deta, d2eta, h5) || plan_alloc(pl, max_pts)) {
    const std::string keep = g_err;
    stabgpu_plan_destroy(pl);                  // frees the streams and events plan_alloc may already have created
    g_err = keep;
    return 1;
  }
  *out = pl;
  return 0;
}

int stabgpu_plan_create(stabgpu_plan** out, int kind, const stabgpu_params* p, const double* vm, const double* g2vm,
                        const double* g22vm, const double* deta, const double* d2eta, const double* h5, int max_pts,
                        int want_vectors) {
  if (ensure_init()) return 1;                 // selects the primary device for this thread
  return plan_create_here(out, kind, p, vm, g2vm, g22vm, deta, d2eta, h5, max_pts, want_vectors);
}

int stabgpu_plan_upload(stabgpu_plan* pl, int npts, const double* s1, const double* s2, const double* Re_pt, const double* Ma_pt) {
  if (!pl || npts < 1 || npts > pl->cap || !s1 || !s2) return fail("libstabgpu: plan_upload bad argument");
  if (use_plan(pl)) return 1;
  pl->npts = npts;
  CU(cudaMemcpyAsync(pl->s1.p, s1, sizeof(cplx) * npts, cudaMemcpyHostToDevice, pl->stream));
  CU(cudaMemcpyAsync(pl->s2.p, s2, sizeof(cplx) * npts, cudaMemcpyHostToDevice, pl->stream));
  pl->has_Re = Re_pt != nullptr; pl->has_Ma = Ma_pt != nullptr;
  if (Re_pt) CU(cudaMemcpyAsync(pl->Re.p, Re_pt, sizeof(double) * npts, cudaMemcpyHostToDevice, pl->stream));
  if (Ma_pt) CU(cudaMemcpyAsync(pl->Ma.p, Ma_pt, sizeof(double) * npts, cudaMemcpyHostToDevice, pl->stream));
  CU(cudaStreamSynchronize(pl->stream));
  return 0;
}

int stabgpu_plan_enqueue(stabgpu_plan* pl) {
  if (!pl || pl->npts < 1) return fail("libstabgpu: plan_execute without uploaded points");
  if (use_plan(pl)) return 1;
  const int np = pl->npts, ny = pl->ny, n = pl->n, N = pl->N;
  cudaStream_t s = pl->stream;
  pl->launches = 0;
  GridDev g = pl->grid();
  Phys ph = phys_from(&pl->prm);
  SweepDev sw; sw.s1 = pl->s1.p; sw.s2 = pl->s2.p; sw.Re = pl->has_Re ? pl->Re.p : nullptr; sw.Ma = pl->has_Ma ? pl->Ma.p : nullptr;
  CU(cudaEventRecord(pl->ev[0], s));
  dim3 cgrid((ny + 63) / 64, np);
  dim3 agrid((n + ASM_ROWS - 1) / ASM_ROWS, (ny + ASM_JT - 1) / ASM_JT, np);
  if (pl->kind == 1) {
    k_node_coef_temporal<<<cgrid, 64, 0, s>>>(g, ph, sw, 0, 1, pl->coef.p, nullptr);
    CU(cudaGetLastError());
    k_assemble_temporal<<<agrid, ASM_ROWS, 0, s>>>(g, pl->coef.p, pl->A.p, (size_t)N * N);
    CU(cudaGetLastError());
    CU(cudaEventRecord(pl->ev[ST_ASM + 1], s));
    CU(cudaEventRecord(pl->ev[ST_LU + 1], s));
    CU(cudaMemsetAsync(pl->info_lu.p, 0, sizeof(int) * np, s));
    pl->launches += 2;
    if (run_eigen(pl, 1, n)) return 1;
  } else {
    k_node_coef_spatial<<<cgrid, 64, 0, s>>>(g, ph, sw, 0, pl->coef.p);
    CU(cudaGetLastError());
    k_assemble_spatial<<<agrid, ASM_ROWS, 0, s>>>(g, pl->coef.p, pl->C.p, (size_t)n * n, pl->A.p, (size_t)N * N);
    CU(cudaGetLastError());
    CU(cudaEventRecord(pl->ev[ST_ASM + 1], s));
    if (g_tune.lu_mode == 1) {
      if (run_lu_blocked(pl)) return 1;
    } else {
      size_t sm = 160 * sizeof(double) + (size_t)n * sizeof(cplx);
      k_lu<<<np, 512, sm, s>>>(pl->C.p, (size_t)n * n, n, pl->A.p, (size_t)N * N, N, N, pl->info_lu.p);
      CU(cudaGetLastError());
      pl->launches += 1;
    }
    CU(cudaEventRecord(pl->ev[ST_LU + 1], s));
    pl->launches += 2;
    if (pl->stop_after_lu) return 0;
    if (run_eigen(pl, 2, 0)) return 1;
  }
  return 0;
}

int stabgpu_plan_wait(stabgpu_plan* pl) {
  if (!pl) return fail("libstabgpu: null plan");
  if (use_plan(pl)) return 1;
  CU(cudaStreamSynchronize(pl->stream));
  if (pl->prof_hess && pl->pev_n > 1) {
    for (int c = 0; c < 8; ++c) pl->hess_ms[c] = 0.f;
    for (size_t i = 1; i < pl->pev_n; ++i) {
      float t = 0.f;
      cudaEventElapsedTime(&t, pl->pev[i - 1], pl->pev[i]);
      pl->hess_ms[pl->pev_cls[i] < 8 ? pl->pev_cls[i] : 3] += t;
    }
  }
  for (int i = 0; i < ST_N; ++i) {
    float t = 0.f;
    cudaEventElapsedTime(&t, pl->ev[i], pl->ev[i + 1]);
    pl->ms[i] = t;
  }
  return 0;
}

int stabgpu_plan_execute(stabgpu_plan* pl) {
  if (stabgpu_plan_enqueue(pl)) return 1;
  return stabgpu_plan_wait(pl);
}

int stabgpu_plan_download(stabgpu_plan* pl, double* eig, double* evec, int* info) {
  if (!pl || pl->npts < 1) return fail("libstabgpu: plan_download without results");
  if (use_plan(pl)) return 1;
  const int np = pl->npts, N = pl->N;
  if (eig) CU(cudaMemcpy(eig, pl->eig.p, sizeof(cplx) * (size_t)np * N, cudaMemcpyDeviceToHost));
  if (evec) {
    if (!pl->want_vectors) return fail("libstabgpu: plan was created without eigenvectors");
    CU(cudaMemcpy(evec, pl->V.p, sizeof(cplx) * (size_t)np * N * N, cudaMemcpyDeviceToHost));
  }
  if (info) {
    std::vector<int> a(np), b(np);
    CU(cudaMemcpy(a.data(), pl->info_lu.p, sizeof(int) * np, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(b.data(), pl->info_qr.p, sizeof(int) * np, cudaMemcpyDeviceToHost));
    for (int i = 0; i < np; ++i) info[i] = a[i] ? a[i] : b[i];
  }
  return 0;
}

int stabgpu_plan_stage_times(stabgpu_plan* pl, float* ms) {
  if (!pl) return 1;
  for (int i = 0; i < ST_N; ++i) ms[i] = pl->ms[i];
  return 0;
}

long long stabgpu_plan_launch_count(stabgpu_plan* pl) { return pl ? pl->launches : 0; }

void* stabgpu_plan_stream(stabgpu_plan* pl) { return pl ? (void*)pl->stream : nullptr; }
int stabgpu_plan_capacity(stabgpu_plan* pl) { return pl ? pl->cap : 0; }

/* per-kernel-class breakdown of the Hessenberg stage: enable, execute once, read ms[4] = panel_step, gemv, gemm, other */
int stabgpu_plan_profile_eigvec(stabgpu_plan* pl, float* ms3) {
  if (!pl) return 1;
  if (ms3) for (int c = 0; c < 3; ++c) ms3[c] = pl->hess_ms[4 + c];
  return 0;
}

int stabgpu_plan_profile_hessenberg(stabgpu_plan* pl, int enable, float* ms4) {
  if (!pl) return 1;
  pl->prof_hess = enable != 0;
  if (ms4) for (int c = 0; c < 4; ++c) ms4[c] = pl->hess_ms[c];
  return 0;
}

int stabgpu_plan_ilohi(stabgpu_plan* pl, int* ilohi) {
  if (!pl || pl->npts < 1 || !ilohi) return fail("libstabgpu: plan_ilohi bad argument");
  if (use_plan(pl)) return 1;
  CU(cudaMemcpy(ilohi, pl->ilohi.p, sizeof(int) * 2 * pl->npts, cudaMemcpyDeviceToHost));
  return 0;
}

void* stabgpu_plan_eig_dev(stabgpu_plan* pl) { return pl ? (void*)pl->eig.p : nullptr; }

int stabgpu_plan_destroy(stabgpu_plan* pl) {
  if (!pl) return 0;
  for (auto& d : g_devs) if (d.cached == pl) d.cached = nullptr;
  int prev = 0;
  const bool switched = cudaGetDevice(&prev) == cudaSuccess && prev != pl->device && cudaSetDevice(pl->device) == cudaSuccess;
  if (pl->hess_graph) cudaGraphExecDestroy(pl->hess_graph);
  if (pl->stream) cudaStreamDestroy(pl->stream);
  if (pl->stream2) cudaStreamDestroy(pl->stream2);
  if (pl->evFork) cudaEventDestroy(pl->evFork);
  if (pl->evJoin) cudaEventDestroy(pl->evJoin);
  for (int i = 0; i < 8; ++i) if (pl->evSub[i]) cudaEventDestroy(pl->evSub[i]);
  for (auto e : pl->evA) cudaEventDestroy(e);
  for (auto e : pl->evB) cudaEventDestroy(e);
  for (auto e : pl->pev) cudaEventDestroy(e);
  for (int i = 0; i <= ST_N; ++i) if (pl->ev[i]) cudaEventDestroy(pl->ev[i]);
  delete pl;                                   // DBuf members: cudaFree on the owning device
  if (switched) cudaSetDevice(prev);
  return 0;
}

// The batch entry points keep ONE plan (device workspace) per device alive between calls: a sweep driver calls
// them repeatedly with the same problem shape, and cudaMalloc of GBs per call would dominate.

}  // extern "C"

namespace {

void parallel_memcpy(void* dst, const void* src, size_t bytes) {
  int want = g_stage_threads;
  if (want <= 0) {
    const int hw = (int)std::thread::hardware_concurrency(), nd = g_devs.empty() ? 1 : (int)g_devs.size();
    want = std::min(8, std::max(2, hw / nd));
  }
  const int nt = (bytes < ((size_t)4 << 20)) ? 1 : want;
  if (nt <= 1) { std::memcpy(dst, src, bytes); return; }
  std::vector<std::thread> th;
  const size_t part = ((bytes / nt) + 4095) & ~(size_t)4095;
  for (int i = 1; i < nt; ++i) {
    const size_t off = (size_t)i * part;
    if (off >= bytes) break;
    const size_t len = std::min(part, bytes - off);
    th.emplace_back([=] { std::memcpy((char*)dst + off, (const char*)src + off, len); });
  }
  std::memcpy(dst, src, std::min(part, bytes));
  for (auto& t : th) t.join();
}

int ring_ensure(StageRing& r, size_t want) {
  if (r.slot_bytes >= want) return 0;
  ring_release(r);
  for (int i = 0; i < StageRing::NSLOT; ++i) {
    CU(cudaHostAlloc((void**)&r.buf[i], want, cudaHostAllocPortable));
    CU(cudaEventCreateWithFlags(&r.ev[i], cudaEventDisableTiming));
  }
  r.slot_bytes = want;
  return 0;
}

// Pageable destination: the vectors of each finished sub-batch travel device -> pinned ring slot (DMA, copy stream) ->
// caller array (host threads).  Up to NSLOT-1 chunks are in flight on the copy engine while one is copied out, and all of
// it runs under the kernels of the later sub-batches, which are already enqueued on the compute stream.
int drain_vectors_staged(stabgpu_plan* pl, StageRing& ring, double* dst_host) {
  const size_t st = (size_t)pl->N * pl->N, mat_bytes = st * sizeof(cplx);
  size_t per = ((size_t)64 << 20) / mat_bytes; if (per < 1) per = 1;
  if (ring_ensure(ring, per * mat_bytes)) return 1;
  struct Pend { char* dst; size_t bytes; bool live; } pend[StageRing::NSLOT] = {};
  auto finish = [&](int slot) -> int {
    if (!pend[slot].live) return 0;
    CU(cudaEventSynchronize(ring.ev[slot]));
    parallel_memcpy(pend[slot].dst, ring.buf[slot], pend[slot].bytes);
    pend[slot].live = false;
    return 0;
  };
  long long k = 0;
  for (int sb = 0; sb < pl->nsub; ++sb) {
    const int m0 = pl->sub_m0[sb], m1 = pl->sub_m0[sb + 1];
    if (m1 <= m0) continue;
    CU(cudaStreamWaitEvent(pl->stream2, pl->evSub[sb], 0));
    for (int m = m0; m < m1; m += (int)per, ++k) {
      const int slot = (int)(k % StageRing::NSLOT);
      if (finish(slot)) return 1;
      const int cnt = std::min((int)per, m1 - m);
      CU(cudaMemcpyAsync(ring.buf[slot], pl->V.p + (size_t)m * st, (size_t)cnt * mat_bytes, cudaMemcpyDeviceToHost, pl->stream2));
      CU(cudaEventRecord(ring.ev[slot], pl->stream2));
      pend[slot].dst = (char*)dst_host + (size_t)m * mat_bytes; pend[slot].bytes = (size_t)cnt * mat_bytes; pend[slot].live = true;
    }
  }
  for (long long q = k; q < k + StageRing::NSLOT; ++q) if (finish((int)(q % StageRing::NSLOT))) return 1;
  return 0;
}

bool host_is_pinned(const void* ptr) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, ptr) != cudaSuccess) { cudaGetLastError(); return false; }
  return at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged;
}

struct BatchArgs {
  int kind; const stabgpu_params* p; const double *vm, *g2vm, *g22vm, *deta, *d2eta, *h5;
  const double *s1, *s2, *Re_pt, *Ma_pt; int want_vectors; double *eig, *evec; int* info; bool evec_pinned;
};

// points [lo, hi) of the call on device slot `slot`; runs on its own host thread when the call is sharded
int batch_on_device(int slot, const BatchArgs& a, int lo, int hi) {
  DevCtx& dc = g_devs[slot];
  CU(cudaSetDevice(dc.device));
  const int npts = hi - lo, want_vectors = a.want_vectors;
  stabgpu_plan* pl = dc.cached;
  if (pl && (pl->kind != a.kind || pl->ny != a.p->ny || pl->want_vectors != (want_vectors ? 1 : 0) || (pl->cap < npts && !pl->cap_limited))) {
    dc.cached = nullptr;
    stabgpu_plan_destroy(pl);
    pl = nullptr;
  }
  if (!pl) {
    if (plan_create_here(&pl, a.kind, a.p, a.vm, a.g2vm, a.g22vm, a.deta, a.d2eta, a.h5, npts, want_vectors)) return 1;
    dc.cached = pl;
  } else {
    pl->prm = *a.p;
    if (upload_grid(pl, a.p, a.vm, a.g2vm, a.g22vm, a.deta, a.d2eta, a.h5)) return 1;
  }
  const int N = pl->N;
  const bool staged = want_vectors && !a.evec_pinned && g_pin_mode == 1;
  int rc = 0;
  for (int p0 = lo; p0 < hi && !rc; p0 += pl->cap) {
    int m = hi - p0; if (m > pl->cap) m = pl->cap;
    rc = stabgpu_plan_upload(pl, m, a.s1 + 2 * (size_t)p0, a.s2 + 2 * (size_t)p0, a.Re_pt ? a.Re_pt + p0 : nullptr, a.Ma_pt ? a.Ma_pt + p0 : nullptr);
    double* dst = want_vectors ? a.evec + 2 * (size_t)p0 * N * N : nullptr;
    pl->evec_host = dst;                       // D2H of the vectors overlaps the eigenvector stage
    pl->evec_staged = staged;
    if (!rc) rc = stabgpu_plan_enqueue(pl);
    if (!rc && staged) rc = drain_vectors_staged(pl, dc.ring, dst);
    if (!rc) rc = stabgpu_plan_wait(pl);
    pl->evec_host = nullptr; pl->evec_staged = false;
    if (!rc) rc = stabgpu_plan_download(pl, a.eig + 2 * (size_t)p0 * N, nullptr, a.info ? a.info + p0 : nullptr);
  }
  return rc;
}

// run fn(slot, lo, hi) for the contiguous shards of stabgpu_shard_range on the devices of g_devs, one host thread per
// device (slot 0 on the calling thread); the first error message is handed to the caller's stabgpu_last_error()
template <class F>
int shard_over_devices(int npts, F fn) {
  const int ndev = (int)g_devs.size() < npts ? (int)g_devs.size() : npts;
  if (ndev <= 1) { const int rc = fn(0, 0, npts); cudaSetDevice(g_device); return rc; }
  std::vector<int> rcs(ndev, 0);
  std::vector<std::string> errs(ndev);
  std::vector<std::thread> th;
  auto work = [&](int d) {
    int lo = 0, hi = 0;
    stabgpu_shard_range(npts, d, ndev, &lo, &hi);
    g_err.clear();
    rcs[d] = hi > lo ? fn(d, lo, hi) : 0;
    if (rcs[d]) errs[d] = g_err;
  };
  for (int d = 1; d < ndev; ++d) th.emplace_back(work, d);
  work(0);
  for (auto& t : th) t.join();
  cudaSetDevice(g_device);
  for (int d = 0; d < ndev; ++d)
    if (rcs[d]) return fail("device " + std::to_string(g_devs[d].device) + ": " + errs[d]);
  return 0;
}

}  // namespace

static int batch_common(int kind, const stabgpu_params* p, const double* vm, const double* g2vm, const double* g22vm,
                        const double* deta, const double* d2eta, const double* h5, int npts, const double* s1,
                        const double* s2, const double* Re_pt, const double* Ma_pt, int want_vectors, double* eig,
                        double* evec, int* info) {
  if (npts < 1 || !s1 || !s2 || !eig) return fail("libstabgpu: bad argument");
  if (want_vectors && !evec) return fail("libstabgpu: want_vectors set but evec is NULL");
  if (ensure_init()) return 1;
  if (check_params(p)) return 1;
  if (!vm || !deta || !d2eta) return fail("libstabgpu: bad argument");
  BatchArgs a{kind, p, vm, g2vm, g22vm, deta, d2eta, h5, s1, s2, Re_pt, Ma_pt, want_vectors, eig, evec, info,
              want_vectors ? host_is_pinned(evec) : true};
  return shard_over_devices(npts, [&](int slot, int lo, int hi) { return batch_on_device(slot, a, lo, hi); });
}

// Batched polish (polish.cuh) of the points [lo, hi) on device slot `slot`
struct PolishArgs {
  int kind; const stabgpu_params* p; const double *vm, *g2vm, *g22vm, *deta, *d2eta, *h5;
  const double *s1, *s2, *Re_pt, *Ma_pt, *sigma, *x0; int max_iters; double tol;
  double *lambda, *x, *resid; int* iters;
};

static int polish_on_device(int slot, const PolishArgs& a, int lo, int hi) {
  CU(cudaSetDevice(g_devs[slot].device));
  const stabgpu_params* p = a.p;
  const int ny = p->ny, n = 5 * ny, kind = a.kind;
  const size_t st = (size_t)n * n;
  stabgpu_plan* pl = new stabgpu_plan();       // grid / profile holder only
  CU(cudaGetDevice(&pl->device));
  pl->kind = kind; pl->prm = *p; pl->ny = ny; pl->n = n; pl->N = n;
  struct Guard { stabgpu_plan* pl; ~Guard() { stabgpu_plan_destroy(pl); } } guard{pl};
  if (upload_grid(pl, p, a.vm, a.g2vm, a.g22vm, a.deta, a.d2eta, a.h5)) return 1;
  size_t freeb = 0, totalb = 0;
  CU(cudaMemGetInfo(&freeb, &totalb));
  const size_t per_pt = ((kind == 2 ? 4 : 3) * st + (size_t)ny * 175 + 3 * (size_t)n) * sizeof(cplx) + 4096;
  int cap = (int)std::min<size_t>((size_t)(hi - lo), (size_t)(0.8 * (double)freeb) / per_pt);
  if (cap < 1) return fail("libstabgpu: not enough device memory to polish a single point");
  DBuf<cplx> coef, blk, M0, M1, M2, K, sv1, sv2, sg, xv, dummy;
  DBuf<double> Re, Ma, out4;
  DBuf<int> ipiv, perm, info;
  if (coef.alloc((size_t)cap * ny * (kind == 1 ? 75 : 150)) || (kind == 1 && blk.alloc((size_t)cap * ny * 25)) || M0.alloc(cap * st) ||
      M1.alloc(cap * st) || (kind == 2 && M2.alloc(cap * st)) || K.alloc(cap * st) || sv1.alloc(cap) || sv2.alloc(cap) || sg.alloc(cap) ||
      xv.alloc((size_t)cap * n) || dummy.alloc(16) || Re.alloc(cap) || Ma.alloc(cap) || out4.alloc((size_t)cap * 4) ||
      ipiv.alloc((size_t)cap * n) || perm.alloc((size_t)cap * LU_PERM) || info.alloc(cap)) return 1;
  cudaStream_t s = nullptr;
  CU(cudaStreamCreate(&s));
  struct SGuard { cudaStream_t s; ~SGuard() { cudaStreamDestroy(s); } } sguard{s};
  const size_t smem = 160 * sizeof(double) + 4 * (size_t)n * sizeof(cplx);
  CU(cudaFuncSetAttribute(k_polish_iterate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  GridDev g = pl->grid();
  Phys ph = phys_from(p);
  std::vector<cplx> xh;
  std::vector<double> oh;
  long long launches = 0;
  for (int p0 = lo; p0 < hi; p0 += cap) {
    const int m = std::min(cap, hi - p0);
    CU(cudaMemcpyAsync(sv1.p, a.s1 + 2 * (size_t)p0, sizeof(cplx) * m, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(sv2.p, a.s2 + 2 * (size_t)p0, sizeof(cplx) * m, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(sg.p, a.sigma + 2 * (size_t)p0, sizeof(cplx) * m, cudaMemcpyHostToDevice, s));
    if (a.Re_pt) CU(cudaMemcpyAsync(Re.p, a.Re_pt + p0, sizeof(double) * m, cudaMemcpyHostToDevice, s));
    if (a.Ma_pt) CU(cudaMemcpyAsync(Ma.p, a.Ma_pt + p0, sizeof(double) * m, cudaMemcpyHostToDevice, s));
    if (a.x0) {
      CU(cudaMemcpyAsync(xv.p, a.x0 + 2 * (size_t)p0 * n, sizeof(cplx) * (size_t)m * n, cudaMemcpyHostToDevice, s));
    } else {
      xh.assign((size_t)m * n, mk(1.0, 0.0));
      CU(cudaMemcpyAsync(xv.p, xh.data(), sizeof(cplx) * (size_t)m * n, cudaMemcpyHostToDevice, s));
    }
    SweepDev sw; sw.s1 = sv1.p; sw.s2 = sv2.p; sw.Re = a.Re_pt ? Re.p : nullptr; sw.Ma = a.Ma_pt ? Ma.p : nullptr;
    dim3 cgrid((ny + 63) / 64, m);
    if (kind == 1) k_node_coef_temporal<<<cgrid, 64, 0, s>>>(g, ph, sw, 0, 0, coef.p, blk.p);
    else k_node_coef_spatial<<<cgrid, 64, 0, s>>>(g, ph, sw, 0, coef.p);
    CU(cudaGetLastError());
    LuBatch lb{K.p, st, n, dummy.p, 0, n, 0, ipiv.p, perm.p, info.p};
    PolishBatch pb{n, kind, M0.p, M1.p, M2.p, K.p, st, ipiv.p, info.p, sg.p, xv.p, out4.p};
    k_polish_form<<<dim3((unsigned)((st + 255) / 256), m), 256, 0, s>>>(g, coef.p, blk.p, pb);
    CU(cudaGetLastError());
    if (lu_run(lb, m, s, &launches)) return 1;
    k_polish_iterate<<<m, 512, smem, s>>>(pb, a.max_iters, a.tol);
    CU(cudaGetLastError());
    oh.resize((size_t)m * 4);
    CU(cudaMemcpyAsync(oh.data(), out4.p, sizeof(double) * 4 * m, cudaMemcpyDeviceToHost, s));
    if (a.x) CU(cudaMemcpyAsync(a.x + 2 * (size_t)p0 * n, xv.p, sizeof(cplx) * (size_t)m * n, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    for (int k = 0; k < m; ++k) {
      a.lambda[2 * (size_t)(p0 + k)] = oh[4 * k]; a.lambda[2 * (size_t)(p0 + k) + 1] = oh[4 * k + 1];
      if (a.resid) a.resid[p0 + k] = oh[4 * k + 2];
      if (a.iters) a.iters[p0 + k] = (int)oh[4 * k + 3];
    }
  }
  return 0;
}

extern "C" {

/* Stage (4) of the north star, batched over sweep points: see polish.cuh and include/stabgpu.h */
int stabgpu_polish_batch(int kind, const stabgpu_params* p, const double* vm, const double* g2vm, const double* g22vm,
                         const double* deta, const double* d2eta, const double* h5, int npts, const double* s1,
                         const double* s2, const double* Re_pt, const double* Ma_pt, const double* sigma, const double* x0,
                         int max_iters, double tol, double* lambda, double* x, double* resid, int* iters) {
  if (kind != 1 && kind != 2) return fail("libstabgpu: polish kind must be 1 (temporal) or 2 (spatial)");
  if (npts < 1 || !vm || !deta || !d2eta || !s1 || !s2 || !sigma || !lambda) return fail("libstabgpu: polish bad argument");
  if (ensure_init()) return 1;
  if (check_params(p)) return 1;
  if (5 * p->ny > 1280) return fail("libstabgpu: polish supports ny <= 256");
  PolishArgs a{kind, p, vm, g2vm, g22vm, deta, d2eta, kind == 2 ? h5 : nullptr, s1, s2, Re_pt, Ma_pt, sigma, x0,
               max_iters < 1 ? 12 : max_iters, tol <= 0.0 ? 1e-13 : tol, lambda, x, resid, iters};
  return shard_over_devices(npts, [&](int slot, int lo, int hi) { return polish_on_device(slot, a, lo, hi); });
}

int stabgpu_temporal_batch(const stabgpu_params* p, const double* vm, const double* g2vm, const double* g22vm,
                           const double* deta, const double* d2eta, int npts, const double* alpha, const double* beta,
                           const double* Re_pt, const double* Ma_pt, int want_vectors, double* omg, double* evec, int* info) {
  return batch_common(1, p, vm, g2vm, g22vm, deta, d2eta, nullptr, npts, alpha, beta, Re_pt, Ma_pt, want_vectors, omg, evec, info);
}

int stabgpu_spatial_batch(const stabgpu_params* p, const double* vm, const double* g2vm, const double* g22vm,
                          const double* deta, const double* d2eta, const double* h5, int npts, const double* omega,
                          const double* beta, const double* Re_pt, const double* Ma_pt, int want_vectors, double* alp,
                          double* evec, int* info) {
  return batch_common(2, p, vm, g2vm, g22vm, deta, d2eta, h5, npts, omega, beta, Re_pt, Ma_pt, want_vectors, alp, evec, info);
}

int stabgpu_zgeev_batch(int n, int batch, const double* A, int want_vectors, double* w, double* V, int* info) {
  if (ensure_init()) return 1;
  if (n < 1 || batch < 1 || !A || !w) return fail("libstabgpu: bad argument");
  if (want_vectors && !V) return fail("libstabgpu: want_vectors set but V is NULL");
  if (n == 1) {                                        // ZGEEV's quick return: the entry is the eigenvalue, the vector is 1
    for (int b = 0; b < batch; ++b) {
      w[2 * b] = A[2 * b]; w[2 * b + 1] = A[2 * b + 1];
      if (want_vectors) { V[2 * b] = 1.0; V[2 * b + 1] = 0.0; }
      if (info) info[b] = 0;
    }
    return 0;
  }
  stabgpu_plan* pl = new stabgpu_plan();
  cudaGetDevice(&pl->device);
  pl->kind = 3; pl->ny = 0; pl->n = n; pl->N = n; pl->want_vectors = want_vectors ? 1 : 0;
  int rc = plan_alloc(pl, batch);
  for (int p0 = 0; p0 < batch && !rc; p0 += pl->cap) {
    int m = batch - p0; if (m > pl->cap) m = pl->cap;
    pl->npts = m;
    const size_t st = (size_t)n * n;
    rc = (cudaMemcpy(pl->A.p, A + 2 * (size_t)p0 * st, sizeof(cplx) * m * st, cudaMemcpyHostToDevice) != cudaSuccess);
    if (rc) { fail("libstabgpu: H2D copy failed"); break; }
    cudaEventRecord(pl->ev[0], pl->stream);
    cudaEventRecord(pl->ev[ST_ASM + 1], pl->stream);
    cudaEventRecord(pl->ev[ST_LU + 1], pl->stream);
    pl->launches = 0;
    rc = run_eigen(pl, 0, 0);
    if (!rc) rc = (cudaStreamSynchronize(pl->stream) != cudaSuccess);
    if (rc) { if (g_err.empty()) fail("libstabgpu: eigen pipeline failed"); break; }
    if (cudaMemsetAsync(pl->info_lu.p, 0, sizeof(int) * m, pl->stream) != cudaSuccess || cudaStreamSynchronize(pl->stream) != cudaSuccess) {
      rc = fail("libstabgpu: clearing the LU status failed");
      break;
    }
    rc = stabgpu_plan_download(pl, w + 2 * (size_t)p0 * n, want_vectors ? V + 2 * (size_t)p0 * st : nullptr, info ? info + p0 : nullptr);
  }
  {
    cudaError_t e = cudaGetLastError();
    if (!rc && e != cudaSuccess) rc = fail(std::string("libstabgpu: ") + cudaGetErrorString(e));
  }
  stabgpu_plan_destroy(pl);
  return rc;
}

int stabgpu_debug_stages(int n, const double* A, double* balanced, double* scale, int* ilo, int* ihi, double* hess, double* tau) {
  if (ensure_init()) return 1;
  if (n < 1 || !A) return fail("libstabgpu: bad argument");
  stabgpu_plan* pl = new stabgpu_plan();
  struct Guard { stabgpu_plan* pl; ~Guard() { stabgpu_plan_destroy(pl); } } guard{pl};
  CU(cudaGetDevice(&pl->device));
  pl->kind = 3; pl->n = n; pl->N = n; pl->want_vectors = 0;
  if (plan_alloc(pl, 1)) return 1;
  const size_t st = (size_t)n * n;
  cudaStream_t s = pl->stream;
  CU(cudaMemcpy(pl->A.p, A, sizeof(cplx) * st, cudaMemcpyHostToDevice));
  {
    const int bb = balance_block(n);
    const size_t smb = balance_wsp_doubles(n, bb) * sizeof(double);
    CU(cudaFuncSetAttribute(k_balance, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smb));
    k_balance<<<1, 256, smb, s>>>(pl->A.p, st, n, pl->scale.p, pl->cnt.p, pl->ilohi.p, bb);
    CU(cudaGetLastError());
  }
  CU(cudaStreamSynchronize(s));
  int lh[2];
  CU(cudaMemcpy(lh, pl->ilohi.p, sizeof(lh), cudaMemcpyDeviceToHost));
  if (ilo) *ilo = lh[0];
  if (ihi) *ihi = lh[1];
  if (balanced) CU(cudaMemcpy(balanced, pl->A.p, sizeof(cplx) * st, cudaMemcpyDeviceToHost));
  if (scale) CU(cudaMemcpy(scale, pl->scale.p, sizeof(double) * n, cudaMemcpyDeviceToHost));
  pl->npts = 1;
  if (run_hessenberg(pl)) return 1;
  CU(cudaStreamSynchronize(s));
  if (hess) CU(cudaMemcpy(hess, pl->A.p, sizeof(cplx) * st, cudaMemcpyDeviceToHost));
  if (tau) CU(cudaMemcpy(tau, pl->tau.p, sizeof(cplx) * n, cudaMemcpyDeviceToHost));
  CU(cudaGetLastError());
  return 0;
}

static int inspect_common(int kind, const stabgpu_params* p, const double* vm, const double* g2vm, const double* g22vm,
                          const double* deta, const double* d2eta, const double* h5, const double* s1, const double* s2,
                          double* o0, double* o1, double* o2) {
  if (ensure_init()) return 1;
  if (check_params(p)) return 1;
  if (!vm || !deta || !d2eta || !s1 || !s2 || !o0 || !o1) return fail("libstabgpu: bad argument");
  stabgpu_plan* pl = new stabgpu_plan();
  struct Guard { stabgpu_plan* pl; ~Guard() { stabgpu_plan_destroy(pl); } } guard{pl};
  CU(cudaGetDevice(&pl->device));
  pl->kind = kind; pl->prm = *p; pl->ny = p->ny; pl->n = 5 * p->ny; pl->N = pl->n; pl->want_vectors = 0;
  if (upload_grid(pl, p, vm, g2vm, g22vm, deta, d2eta, h5)) return 1;
  const int ny = p->ny, n = pl->n;
  const size_t st = (size_t)n * n;
  DBuf<cplx> coef, blk, m0, m1, m2, sv1, sv2;
  if (coef.alloc((size_t)ny * 150) || blk.alloc((size_t)ny * 25) || m0.alloc(st) || m1.alloc(st) || m2.alloc(st) || sv1.alloc(1) || sv2.alloc(1)) return 1;
  CU(cudaMemcpy(sv1.p, s1, sizeof(cplx), cudaMemcpyHostToDevice));
  CU(cudaMemcpy(sv2.p, s2, sizeof(cplx), cudaMemcpyHostToDevice));
  GridDev g = pl->grid();
  Phys ph = phys_from(p);
  SweepDev sw; sw.s1 = sv1.p; sw.s2 = sv2.p; sw.Re = nullptr; sw.Ma = nullptr;
  dim3 cgrid((ny + 63) / 64, 1);
  const int eb = (int)((st + 255) / 256);
  if (kind == 1) {
    k_node_coef_temporal<<<cgrid, 64>>>(g, ph, sw, 0, 0, coef.p, blk.p);
    CU(cudaGetLastError());
    k_inspect_temporal<<<eb, 256>>>(g, coef.p, blk.p, m0.p, m1.p);
  } else {
    k_node_coef_spatial<<<cgrid, 64>>>(g, ph, sw, 0, coef.p);
    CU(cudaGetLastError());
    k_inspect_spatial<<<eb, 256>>>(g, coef.p, m0.p, m1.p, m2.p);
  }
  CU(cudaGetLastError());
  CU(cudaDeviceSynchronize());
  CU(cudaMemcpy(o0, m0.p, sizeof(cplx) * st, cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(o1, m1.p, sizeof(cplx) * st, cudaMemcpyDeviceToHost));
  if (o2) CU(cudaMemcpy(o2, m2.p, sizeof(cplx) * st, cudaMemcpyDeviceToHost));
  return 0;
}

int stabgpu_temporal_assemble(const stabgpu_params* p, const double* vm, const double* g2vm, const double* g22vm,
                              const double* deta, const double* d2eta, const double* alpha, const double* beta,
                              double* A0, double* B0) {
  return inspect_common(1, p, vm, g2vm, g22vm, deta, d2eta, nullptr, alpha, beta, A0, B0, nullptr);
}

int stabgpu_spatial_assemble(const stabgpu_params* p, const double* vm, const double* g2vm, const double* g22vm,
                             const double* deta, const double* d2eta, const double* h5, const double* omega,
                             const double* beta, double* C0, double* C1, double* C2) {
  return inspect_common(2, p, vm, g2vm, g22vm, deta, d2eta, h5, omega, beta, C0, C1, C2);
}

/* debug / parity: the reduced spatial operator [M1 | M2] = C0^-1 [-C1 | -C2] (n x 2n, column-major) of one point, as the
 * LU stage leaves it in the top half of the companion matrix (spatial.f90:978-1008) */
int stabgpu_debug_spatial_reduce(const stabgpu_params* p, const double* vm, const double* g2vm, const double* g22vm,
                                 const double* deta, const double* d2eta, const double* h5, const double* omega,
                                 const double* beta, double* M, int* info) {
  stabgpu_plan* pl = nullptr;
  if (stabgpu_plan_create(&pl, 2, p, vm, g2vm, g22vm, deta, d2eta, h5, 1, 0)) return 1;
  int rc = stabgpu_plan_upload(pl, 1, omega, beta, nullptr, nullptr);
  pl->stop_after_lu = true;
  if (!rc) rc = stabgpu_plan_enqueue(pl);
  if (!rc) rc = cudaStreamSynchronize(pl->stream) != cudaSuccess;
  const int n = pl->n, N = pl->N;
  if (!rc) rc = cudaMemcpy2D(M, sizeof(cplx) * n, pl->A.p, sizeof(cplx) * N, sizeof(cplx) * n, N, cudaMemcpyDeviceToHost) != cudaSuccess;
  if (!rc && info) rc = cudaMemcpy(info, pl->info_lu.p, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess;
  stabgpu_plan_destroy(pl);
  return rc ? fail("libstabgpu: debug_spatial_reduce failed") : 0;
}

int stabgpu_temporal_polish(const stabgpu_params* p, const double* vm, const double* g2vm, const double* g22vm,
                            const double* deta, const double* d2eta, const double* alpha, const double* beta,
                            const double* sigma, const double* x0, int max_iters, double tol, double* lambda, double* x,
                            double* resid, int* iters) {
  if (!alpha || !beta || !sigma || !lambda) return fail("libstabgpu: polish bad argument");
  int it = 0;
  if (stabgpu_polish_batch(1, p, vm, g2vm, g22vm, deta, d2eta, nullptr, 1, alpha, beta, nullptr, nullptr, sigma, x0,
                           max_iters < 1 ? 8 : max_iters, tol, lambda, x, resid, &it)) return 1;
  if (iters) *iters = it;
  if (it < 0) return fail("libstabgpu: polish: A0 - sigma B0 is exactly singular (sigma is an eigenvalue to working precision; perturb it)");
  return 0;
}

}  // extern "C"
