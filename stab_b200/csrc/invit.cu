// invit.cu -- translation unit of the inverse-iteration kernels (invit.cuh): stage 6 of the eigen pipeline, the
// eigenvector half of the ZGEEV('N','V') the reference calls (temporal.f90:803, spatial.f90:1043).
#include <cstdlib>
#include "launch.h"
#include "invit.cuh"

namespace stab {

template <int NS, int PB>
static cudaError_t run1v(const cplx* Hh, size_t hstride, int n, const cplx* lam, const int* kr, const double* hnorm, cplx* Y, size_t ystride,
                         int* bad, int rounds, int nmat, cudaStream_t s) {
  const size_t sm = 2 * (size_t)INVIT_CB * n * sizeof(cplx) + (size_t)INVIT_WARPS * n + INVIT_PAD;
  cudaError_t e = cudaFuncSetAttribute(k_invit<NS, PB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  if (e != cudaSuccess) return e;
  dim3 grid((n + INVIT_WARPS * rounds - 1) / (INVIT_WARPS * rounds), nmat);
  k_invit<NS, PB><<<grid, INVIT_WARPS * 32, sm, s>>>(Hh, hstride, n, lam, kr, hnorm, Y, ystride, bad, rounds);
  return cudaGetLastError();
}
template <int NS>
static cudaError_t run1(const cplx* Hh, size_t hstride, int n, const cplx* lam, const int* kr, const double* hnorm, cplx* Y, size_t ystride,
                        int* bad, int rounds, int nmat, int per_step, cudaStream_t s) {
  return per_step ? run1v<NS, 0>(Hh, hstride, n, lam, kr, hnorm, Y, ystride, bad, rounds, nmat, s)
                  : run1v<NS, 1>(Hh, hstride, n, lam, kr, hnorm, Y, ystride, bad, rounds, nmat, s);
}

// orders 640 < n <= 1280: two warps per eigenvalue
template <int NSH, int PB>
static cudaError_t run2v(const cplx* Hh, size_t hstride, int n, const cplx* lam, const int* kr, const double* hnorm, cplx* Y, size_t ystride,
                         int* bad, int rounds, int nmat, cudaStream_t s) {
  const size_t sm = 2 * (size_t)INVIT2_CB * n * sizeof(cplx) + (size_t)INVIT2_PAIRS * n + INVIT_PAD;
  cudaError_t e = cudaFuncSetAttribute(k_invit2<NSH, PB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  if (e != cudaSuccess) return e;
  dim3 grid((n + INVIT2_PAIRS * rounds - 1) / (INVIT2_PAIRS * rounds), nmat);
  k_invit2<NSH, PB><<<grid, INVIT2_PAIRS * 64, sm, s>>>(Hh, hstride, n, lam, kr, hnorm, Y, ystride, bad, rounds);
  return cudaGetLastError();
}
template <int NSH>
static cudaError_t run2(const cplx* Hh, size_t hstride, int n, const cplx* lam, const int* kr, const double* hnorm, cplx* Y, size_t ystride,
                        int* bad, int rounds, int nmat, int per_step, cudaStream_t s) {
  return per_step ? run2v<NSH, 0>(Hh, hstride, n, lam, kr, hnorm, Y, ystride, bad, rounds, nmat, s)
                  : run2v<NSH, 1>(Hh, hstride, n, lam, kr, hnorm, Y, ystride, bad, rounds, nmat, s);
}

cudaError_t launch_invit(const cplx* Hh, size_t hstride, int n, const cplx* lam, const int* kr, const double* hnorm, cplx* Y,
                         size_t ystride, int* bad, int rounds, int nmat, int per_step, cudaStream_t s) {
  static const bool env_per_step = getenv("STAB_INVIT_PER_STEP") != nullptr;   // measurement override of the validation switch
  if (env_per_step) per_step = 1;
  if (n <= 128) return run1<4>(Hh, hstride, n, lam, kr, hnorm, Y, ystride, bad, rounds, nmat, per_step, s);
  if (n <= 256) return run1<8>(Hh, hstride, n, lam, kr, hnorm, Y, ystride, bad, rounds, nmat, per_step, s);
  if (n <= 384) return run1<12>(Hh, hstride, n, lam, kr, hnorm, Y, ystride, bad, rounds, nmat, per_step, s);
  if (n <= 512) return run1<16>(Hh, hstride, n, lam, kr, hnorm, Y, ystride, bad, rounds, nmat, per_step, s);
  if (n <= 640) return run1<20>(Hh, hstride, n, lam, kr, hnorm, Y, ystride, bad, rounds, nmat, per_step, s);
  if (n <= 1024) return run2<16>(Hh, hstride, n, lam, kr, hnorm, Y, ystride, bad, rounds, nmat, per_step, s);
  if (n <= 1280) return run2<20>(Hh, hstride, n, lam, kr, hnorm, Y, ystride, bad, rounds, nmat, per_step, s);
  return cudaErrorInvalidValue;
}

}  // namespace stab
