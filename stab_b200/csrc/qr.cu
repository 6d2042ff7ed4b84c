// qr.cu -- translation unit of the shifted-QR kernel (hqr.cuh): stage 4 of the eigen pipeline, the ZHSEQR of the ZGEEV
// the reference calls (temporal.f90:803, spatial.f90:1043).
#include "launch.h"
#include "hqr.cuh"

namespace stab {

// scratch `sm`: the trailing block of the classic shift computation; the eigenvalues / Householder vector of the deflation window
static size_t hqr_sm_entries(const HqrLaunch& q) { size_t a = (size_t)q.ns_max * (q.ns_max + 1); return a > 64 ? a : 64; }

// reflector record: a pass of the multishift chain, or a whole two-bulge QR iteration on a block that fits the window
static size_t hqr_rec_entries(const HqrLaunch& q) { size_t a = (size_t)q.steps_max * q.ns_max, b = 2 * ((size_t)q.W + 2); return a > b ? a : b; }

size_t hqr_smem_bytes(const HqrLaunch& q) {
  size_t b = 160 * sizeof(double);
  b += (size_t)q.W * (q.W + 1) * sizeof(cplx);
  b += hqr_rec_entries(q) * sizeof(Rot);
  b += (size_t)2 * q.ns_max * sizeof(Rot);
  b += (size_t)q.ns_max * sizeof(cplx);
  b += hqr_sm_entries(q) * sizeof(cplx);
  b += sizeof(SmallCtl);
  return b;
}

__global__ void __launch_bounds__(256, 2) k_hqr(cplx* Hq, size_t hstride, int n, const int* ilohi, cplx* w, int* info, HqrLaunch q, long long* prof, const double* hnorm) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned char* sp = smem_raw;
  double* red = reinterpret_cast<double*>(sp); sp += 160 * sizeof(double);
  HqrSmem sh;
  sh.W = q.W; sh.ldw = q.W + 1; sh.ns_max = q.ns_max; sh.steps_max = q.steps_max; sh.nw = q.nw; sh.nibble = q.nibble;
  sh.win = reinterpret_cast<cplx*>(sp); sp += (size_t)q.W * (q.W + 1) * sizeof(cplx);
  sh.rec = reinterpret_cast<Rot*>(sp); { size_t a = (size_t)q.steps_max * q.ns_max, b = 2 * ((size_t)q.W + 2); sp += (a > b ? a : b) * sizeof(Rot); }
  sh.cur = reinterpret_cast<Rot*>(sp); sp += (size_t)2 * q.ns_max * sizeof(Rot);
  sh.shifts = reinterpret_cast<cplx*>(sp); sp += (size_t)q.ns_max * sizeof(cplx);
  sh.sm = reinterpret_cast<cplx*>(sp); { size_t a = (size_t)q.ns_max * (q.ns_max + 1); sp += (a > 64 ? a : 64) * sizeof(cplx); }
  sh.ctl = reinterpret_cast<SmallCtl*>(sp);
  Cta c = make_cta(red);
  const int p = blockIdx.x;
  __shared__ long long sprof[16];
  sh.prof = (prof && p == 0) ? sprof : nullptr;
  if (sh.prof && threadIdx.x < 16) sprof[threadIdx.x] = 0;
  __syncthreads();
  // A matrix with a NaN or an infinity (bad sweep value, overflowed operator) never deflates: report it as ZHSEQR's
  // "failed to converge" at once (info = n) instead of iterating to the limit; the other points of the batch go on.
  const double hn = hnorm[p];
  if (!(hn == hn) || hn > 1.0e300) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) w[(size_t)p * n + i] = mk(hn - hn, hn - hn);   // NaN
    if (threadIdx.x == 0) info[p] = n;
    return;
  }
  int r = cta_hqr(c, sh, Hq + (size_t)p * hstride, n, n, ilohi[2 * p], ilohi[2 * p + 1], w + (size_t)p * n);
  if (threadIdx.x == 0) info[p] = r;
  if (sh.prof && threadIdx.x < 16) prof[threadIdx.x] = sprof[threadIdx.x];
}

cudaError_t launch_hqr(cplx* Hq, size_t hstride, int n, const int* ilohi, cplx* w, int* info, HqrLaunch q, long long* prof,
                       const double* hnorm, int nmat, int threads, cudaStream_t s) {
  const size_t sm = hqr_smem_bytes(q);
  cudaError_t e = cudaFuncSetAttribute(k_hqr, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  if (e != cudaSuccess) return e;
  k_hqr<<<nmat, threads, sm, s>>>(Hq, hstride, n, ilohi, w, info, q, prof, hnorm);
  return cudaGetLastError();
}

}  // namespace stab
