// launch.h -- host-side launchers of the kernels that live in their own translation units (qr.cu, invit.cu).
// The QR and inverse-iteration kernels are the largest device functions of the library; compiled separately, an edit
// to one of them leaves the code generation of every other kernel untouched (in one translation unit ptxas / NVVM
// heuristics moved unrelated kernels by +-5 %), and the three units build in parallel.
#pragma once
#include <cuda_runtime.h>
#include "common.cuh"

namespace stab {

struct HqrLaunch { int W, ns_max, steps_max; int nw, nibble; };   // nw: deflation window (0 = classic deflation only)
size_t hqr_smem_bytes(const HqrLaunch& q);
// stage 4: eigenvalues of nmat upper Hessenberg matrices (hqr.cuh); `threads` per CTA (256), one CTA per matrix
cudaError_t launch_hqr(cplx* Hq, size_t hstride, int n, const int* ilohi, cplx* w, int* info, HqrLaunch q, long long* prof,
                       const double* hnorm, int nmat, int threads, cudaStream_t s);
// stage 6: right eigenvectors of the Hessenberg matrices by register-resident inverse iteration (invit.cuh), n <= 1280;
// picks the one-warp (n <= 640) or two-warp kernel and its register-slot count from n; per_step != 0 selects the per-step
// form of the one-warp kernel (validation switch) instead of the panel / bulk form
cudaError_t launch_invit(const cplx* Hh, size_t hstride, int n, const cplx* lam, const int* kr, const double* hnorm, cplx* Y,
                         size_t ystride, int* bad, int rounds, int nmat, int per_step, cudaStream_t s);

}  // namespace stab
