// common.cuh -- shared device/host helpers for the stab_b200 kernels.
//
// All kernels in this package are "one CTA (or one grid) per sweep point" FP64 kernels.  The
// CTA-level algorithms (balancing, Hessenberg reduction, shifted QR, inverse iteration, LU) are
// written against a tiny execution-context abstraction so that the *same source* can also be
// traced single-threaded on the host (-DSTAB_EMU, tests/emu/) for logic tests in a GPU-less
// container.  The emulation build is test infrastructure: it is never linked into libstabgpu.so
// and the product has no CPU path.
#pragma once
#include <math.h>
#include <stdint.h>
#include <float.h>

#ifdef STAB_EMU
#define SD_HD inline
#define SD_DEV inline
#define SD_SHARED static
#define SD_NOINLINE inline
#else
#include <cuda_runtime.h>
#define SD_HD __host__ __device__ __forceinline__
#define SD_DEV __device__ __forceinline__
#define SD_SHARED __shared__
#define SD_NOINLINE __device__ __noinline__
#endif

namespace stab {

// ---------------------------------------------------------------------------------------------
// complex double, layout-compatible with double2 / C99 double _Complex / Fortran complex*16
// ---------------------------------------------------------------------------------------------
struct alignas(16) cplx {
  double re, im;
};

SD_HD cplx mk(double r, double i) { cplx z; z.re = r; z.im = i; return z; }
SD_HD cplx operator+(cplx a, cplx b) { return mk(a.re + b.re, a.im + b.im); }
SD_HD cplx operator-(cplx a, cplx b) { return mk(a.re - b.re, a.im - b.im); }
SD_HD cplx operator-(cplx a) { return mk(-a.re, -a.im); }
SD_HD cplx operator*(cplx a, cplx b) { return mk(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re); }
SD_HD cplx operator*(double s, cplx a) { return mk(s * a.re, s * a.im); }
SD_HD cplx operator*(cplx a, double s) { return mk(s * a.re, s * a.im); }
SD_HD cplx& operator+=(cplx& a, cplx b) { a.re += b.re; a.im += b.im; return a; }
SD_HD cplx& operator-=(cplx& a, cplx b) { a.re -= b.re; a.im -= b.im; return a; }
SD_HD cplx conj(cplx a) { return mk(a.re, -a.im); }
SD_HD bool is_zero(cplx a) { return a.re == 0.0 && a.im == 0.0; }
SD_HD double cabs1(cplx a) { return fabs(a.re) + fabs(a.im); }   // LAPACK CABS1
SD_HD double abs2(cplx a) { return a.re * a.re + a.im * a.im; }
SD_HD double cabs(cplx a) { return hypot(a.re, a.im); }
// a * conj(b)
SD_HD cplx mulc(cplx a, cplx b) { return mk(a.re * b.re + a.im * b.im, a.im * b.re - a.re * b.im); }
// acc += a*b
SD_HD void fma_acc(cplx& acc, cplx a, cplx b) {
  acc.re = fma(a.re, b.re, acc.re); acc.re = fma(-a.im, b.im, acc.re);
  acc.im = fma(a.re, b.im, acc.im); acc.im = fma(a.im, b.re, acc.im);
}
// acc += conj(a)*b
SD_HD void fma_acc_conj(cplx& acc, cplx a, cplx b) {
  acc.re = fma(a.re, b.re, acc.re); acc.re = fma(a.im, b.im, acc.re);
  acc.im = fma(a.re, b.im, acc.im); acc.im = fma(-a.im, b.re, acc.im);
}
// acc -= a*b  (four FMAs)
SD_HD void fms_acc(cplx& acc, cplx a, cplx b) {
  acc.re = fma(-a.re, b.re, acc.re); acc.re = fma(a.im, b.im, acc.re);
  acc.im = fma(-a.re, b.im, acc.im); acc.im = fma(-a.im, b.re, acc.im);
}
// robust complex division (Smith), as LAPACK ZLADIV in spirit
SD_HD cplx cdiv(cplx a, cplx b) {
  if (fabs(b.im) <= fabs(b.re)) {
    double r = b.im / b.re, d = b.re + b.im * r;
    return mk((a.re + a.im * r) / d, (a.im - a.re * r) / d);
  } else {
    double r = b.re / b.im, d = b.im + b.re * r;
    return mk((a.re * r + a.im) / d, (a.im * r - a.re) / d);
  }
}
SD_HD cplx csqrt_(cplx z) {
  double m = cabs(z);
  if (m == 0.0) return mk(0.0, 0.0);
  double sr = sqrt(0.5 * (m + fabs(z.re)));
  double si = 0.5 * z.im / sr;
  if (z.re >= 0.0) return mk(sr, si);
  return mk(fabs(si), z.im >= 0.0 ? sr : -sr);
}

// Streaming load of one complex entry: no L1 allocation, evict-first in L2.  For operands that are read once per launch and
// are far larger than the L2 (the trailing matrix of the Hessenberg GEMV: 1.9 GB per column step), so that the small
// panels other kernels re-read every step (Y, V, T: ~200 MB) are not flushed out of the 126 MB L2 by the stream.
#ifdef STAB_EMU
SD_HD cplx ld_stream(const cplx* p) { return *p; }
#else
SD_DEV unsigned long long l2_policy_evict_first() {
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
SD_DEV cplx ld_stream(const cplx* p, unsigned long long pol) {
  cplx v;
  asm("ld.global.L1::no_allocate.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(v.re), "=d"(v.im) : "l"(p), "l"(pol));
  return v;
}
#endif

// machine constants as LAPACK's DLAMCH reports them
#define SD_EPS   1.1102230246251565e-16      /* DLAMCH('E') = 2^-53 */
#define SD_ULP   2.2204460492503131e-16      /* DLAMCH('P') = eps*base */
#define SD_SAFMIN 2.2250738585072014e-308    /* DLAMCH('S') */

// ---------------------------------------------------------------------------------------------
// CTA execution context
// ---------------------------------------------------------------------------------------------
struct Cta {
  int tid;      // thread index in the CTA
  int nt;       // threads in the CTA
  int lane;     // tid % warp
  int wid;      // tid / warp
  int nw;       // warps in the CTA
  int ws;       // warp size (32; 1 under emulation)
  double* red;  // shared scratch for reductions: >= 4*(nw+1) doubles
};

#ifdef STAB_EMU
SD_DEV void cta_sync() {}
SD_DEV void warp_sync() {}
SD_DEV Cta make_cta(double* red) { Cta c; c.tid = 0; c.nt = 1; c.lane = 0; c.wid = 0; c.nw = 1; c.ws = 1; c.red = red; return c; }
SD_DEV double warp_sum(double v) { return v; }
SD_DEV double warp_max(double v) { return v; }
SD_DEV double warp_bcast(double v, int) { return v; }
SD_DEV int warp_bcast_i(int v, int) { return v; }
#else
SD_DEV void cta_sync() { __syncthreads(); }
SD_DEV void warp_sync() { __syncwarp(); }
SD_DEV Cta make_cta(double* red) {
  Cta c; c.tid = threadIdx.x; c.nt = blockDim.x; c.lane = threadIdx.x & 31; c.wid = threadIdx.x >> 5;
  c.nw = (blockDim.x + 31) >> 5; c.ws = 32; c.red = red; return c;
}
SD_DEV double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
SD_DEV double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
SD_DEV double warp_bcast(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
SD_DEV int warp_bcast_i(int v, int src) { return __shfl_sync(0xffffffffu, v, src); }
#endif

SD_DEV cplx warp_sum(cplx v) { return mk(warp_sum(v.re), warp_sum(v.im)); }

// Block-wide sum of up to 4 doubles at once; result returned to every thread.  Two barriers.
SD_DEV void cta_sum4(const Cta& c, double& a, double& b, double& d, double& e) {
#ifdef STAB_EMU
  (void)c; (void)a; (void)b; (void)d; (void)e;
#else
  a = warp_sum(a); b = warp_sum(b); d = warp_sum(d); e = warp_sum(e);
  cta_sync();  // protect red[] from a previous use
  if (c.lane == 0) { c.red[4 * c.wid] = a; c.red[4 * c.wid + 1] = b; c.red[4 * c.wid + 2] = d; c.red[4 * c.wid + 3] = e; }
  cta_sync();
  double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
  for (int w = 0; w < c.nw; ++w) { s0 += c.red[4 * w]; s1 += c.red[4 * w + 1]; s2 += c.red[4 * w + 2]; s3 += c.red[4 * w + 3]; }
  a = s0; b = s1; d = s2; e = s3;
#endif
}
SD_DEV double cta_sum(const Cta& c, double v) { double b = 0, d = 0, e = 0; cta_sum4(c, v, b, d, e); return v; }
SD_DEV cplx cta_sum(const Cta& c, cplx v) { double d = 0, e = 0; cta_sum4(c, v.re, v.im, d, e); return v; }

// Block-wide max of two doubles.
SD_DEV void cta_max2(const Cta& c, double& a, double& b) {
#ifdef STAB_EMU
  (void)c; (void)a; (void)b;
#else
  a = warp_max(a); b = warp_max(b);
  cta_sync();
  if (c.lane == 0) { c.red[2 * c.wid] = a; c.red[2 * c.wid + 1] = b; }
  cta_sync();
  double m0 = c.red[0], m1 = c.red[1];
  for (int w = 1; w < c.nw; ++w) { m0 = fmax(m0, c.red[2 * w]); m1 = fmax(m1, c.red[2 * w + 1]); }
  a = m0; b = m1;
#endif
}

// Block-wide integer max / min (used for "largest index with property").
SD_DEV int cta_max_i(const Cta& c, int v) {
#ifdef STAB_EMU
  (void)c; return v;
#else
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
  int* ri = reinterpret_cast<int*>(c.red);
  cta_sync();
  if (c.lane == 0) ri[c.wid] = v;
  cta_sync();
  int m = ri[0];
  for (int w = 1; w < c.nw; ++w) m = max(m, ri[w]);
  return m;
#endif
}
SD_DEV int cta_min_i(const Cta& c, int v) { return -cta_max_i(c, -v); }

// argmax with first-index tie-break: returns the smallest index among the maxima of `val`
SD_DEV void cta_argmax(const Cta& c, double& val, int& idx) {
#ifdef STAB_EMU
  (void)c; (void)val; (void)idx;
#else
  for (int o = 16; o > 0; o >>= 1) {
    double ov = __shfl_xor_sync(0xffffffffu, val, o);
    int oi = __shfl_xor_sync(0xffffffffu, idx, o);
    if (ov > val || (ov == val && oi < idx)) { val = ov; idx = oi; }
  }
  int* ri = reinterpret_cast<int*>(c.red + 2 * c.nw);
  cta_sync();
  if (c.lane == 0) { c.red[c.wid] = val; ri[c.wid] = idx; }
  cta_sync();
  double bv = c.red[0]; int bi = ri[0];
  for (int w = 1; w < c.nw; ++w) { double ov = c.red[w]; int oi = ri[w]; if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; } }
  val = bv; idx = bi;
#endif
}

}  // namespace stab
