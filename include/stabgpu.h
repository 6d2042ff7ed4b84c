/* stabgpu.h -- C ABI of libstabgpu: the B200-native replacement for the hot path of sscollis/stab
 * (Chebyshev-collocation operator assembly + dense complex eigensolve, batched over sweep points).
 *
 * The reference has no FFI of its own for this path: `temporal` / `spatial` are Fortran subroutines
 * that talk through module `stuff` (stuff.f90:11-59).  The seam defined here is the narrowest data
 * cut of that path: in = grid metrics + mean profile on the grid + scalar parameters +
 * sweep values, out = sorted eigenvalues [+ eigenvectors] + per-point status.  Every entry point
 * names the reference code it replaces.  INTEGRATION.md shows the ISO_C_BINDING side.
 *
 * Conventions
 *   - all pointers are HOST pointers owned by the caller unless the name says `_dev`;
 *   - arrays are column-major, `double` or interleaved complex double (binary compatible with
 *     Fortran real(c_double) / complex(c_double_complex) and C99 double _Complex);
 *   - return value 0 = call accepted, nonzero = whole-call failure (stabgpu_last_error());
 *   - info[p] mirrors LAPACK per point: 0 ok, >0 numerical failure (singular LU pivot index or the
 *     number of unconverged eigenvalues), so the Fortran caller keeps its stop/warn behaviour
 *     (temporal.f90:776-785,806-809; spatial.f90:1050-1056);
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef STABGPU_H
#define STABGPU_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define STABGPU_NDOF 5            /* stuff.f90:54 */

/* Mirrors the scalar run state of module stuff (stuff.f90:11-59) that the hot path reads. */
typedef struct stabgpu_params {
  int ny;            /* collocation points */
  int mattyp;        /* 0 constant mu, 1 Sutherland            (getmat.f90:12-26) */
  int wallt;         /* 0: T'=0 at the wall, 2: adiabatic       (temporal.f90:653-662) */
  int top;           /* spatial only: 1 continuity at infinity  (spatial.f90:702,794-798) */
  int curve;         /* 0 flat, 2 circular arc (metrics supplied by the caller in h5) */
  int ider;          /* 1: differentiate the mean with D1/D2, 0: g2vm/g22vm supplied */
  int ievec;         /* recorded in output files only */
  int wall;          /* recorded in output files only (never tested by the reference) */
  double Ma, Re, Pr;
  double gamma, gamma1, cp;        /* 1.4, 0.4, 1003.1 (stuff.f90:45) */
  double Te, rmue, rlme, cone;     /* edge state (input.f90:43) -- see stabgpu_edge_properties */
  double datmat[3];                /* material constants (input.f90:29-36) */
  double yi, ymax, x;
} stabgpu_params;

/* ---- lifetime --------------------------------------------------------------------------------- */
int stabgpu_init(int device);                 /* ONE device: selects the CUDA device; <0: current device */
/* The drop-in case: the reference is one serial process that walks the sweep point by point
 * (mtemporal.f90:29-39, mspatial.f90:77-96).  After this call every stabgpu_*_batch / stabgpu_polish_batch call shards
 * its points over the first min(max_devices, visible) GPUs of the box -- contiguous shards of stabgpu_shard_range, one
 * host worker thread, one cached plan and one pinned staging ring per device -- and writes into the caller's single
 * host arrays.  max_devices <= 0: all visible devices.  Results are bit-identical to the single-device call. */
int stabgpu_init_multi(int max_devices, int* ndev_used);
int stabgpu_device_count(void);               /* devices the batch calls shard over (0 before init) */
/* Host staging of the eigenvector output.  A page-locked destination (cudaHostAlloc / stabgpu_host_register) receives
 * the vectors by direct DMA under the eigenvector stage; a pageable one (a Fortran `allocate`, malloc, numpy) is served
 * through a pinned staging ring inside the library (4 x 64 MB per device, DMA -> ring -> caller array by `copy_threads`
 * host threads; default: host cores / devices, between 2 and 8), so the overlap survives.  pin_mode 0 disables the ring (plain cudaMemcpyAsync into pageable memory);
 * values < 0 / <= 0 keep the current setting. */
int stabgpu_set_host_staging(int pin_mode, int copy_threads);
int stabgpu_host_register(void* ptr, size_t bytes);    /* page-lock a caller array once (cudaHostRegister, portable) */
int stabgpu_host_unregister(void* ptr);
int stabgpu_finalize(void);
const char* stabgpu_last_error(void);
int stabgpu_device_info(char* name, int name_len, int* sm_count, double* mem_gb);
/* tuning knobs of the QR stage (window size, shifts per sweep, threads); 0 keeps the default */
int stabgpu_set_tuning(int qr_window, int qr_shifts, int qr_threads, int hess_threads);
/* Aggressive early deflation of the QR stage (the ZLAQR3 step of the ZHSEQR inside the reference's ZGEEV): deflation window
 * (default by order: 32 up to 640, 44 above; 0 = classic ZLAHQR-style deflation only, the round-1 algorithm -- kept as a validation switch) and ZLAQR0's
 * NIBBLE in per cent (default 14).  Negative values keep the current setting. */
int stabgpu_set_qr_deflation(int window, int nibble);
/* Hessenberg stage variant: 1 (default) batched blocked reduction with DMMA tensor-core updates; 2 the same with a
 * scalar-FMA GEMM (validation of the tensor-core path); 0 the unblocked one-CTA-per-matrix kernel of v1; 5 as 1 with the right and
 * left trailing updates fused into one rank-64 pass (measured slower, kept as a validated variant) */
int stabgpu_set_hess_mode(int mode);
/* eigenvector stage variant: 1 (default) register-resident inverse iteration in panel / bulk form + tensor-core back-transformation;
 * 3 the same with the per-step inverse iteration of round 1 (validation of the panel / bulk form); 0 the v1 warp kernel */
int stabgpu_set_evec_mode(int mode);
/* spatial LU reduce variant (ZGETRF + 2 x ZGETRS, spatial.f90:978-1004): 1 (default) blocked LU with DMMA rank-32
 * updates; 0 the v1 one-CTA-per-matrix kernel */
int stabgpu_set_lu_mode(int mode);

/* ---- host-side pieces of the path (pure C++, no device) ------------------------------------------ */
void stabgpu_params_default(stabgpu_params* p);                       /* stuff.f90 initial values */
int  stabgpu_edge_properties(stabgpu_params* p, double T0);           /* input.f90:19-44 (incl. quirk q1) */
int  stabgpu_sgengrid(int ny, double yi, double ymax,                 /* sgengrid.f90:15-45 */
                      double* y, double* eta, double* deta, double* d2eta);
int  stabgpu_chebyd(int N, double* D /* (N+1)x(N+1) col-major */);    /* chebyd.f90:2-60 */
int  stabgpu_spline(int n, const double* x, const double* y, double* fdp);                 /* spline.f90:2-47 */
int  stabgpu_speval(int n, const double* x, const double* y, const double* fdp, double xx, double* f); /* spline.f90:49-75 */
/* getmean.f90:27-111: table = nrows x 6 row-major (y rho u v w T), v is forced to 0; vm = ny x 5 col-major */
int  stabgpu_getmean_table(int nrows, const double* table, int ny, const double* y, double* vm);
int  stabgpu_read_profile(const char* path, int* nrows, double* table, int max_rows);      /* getmean.f90:40-80 */
/* temporal.f90:135-179: D1, D2 (ny x ny), wall row of Dt2, and the mapped mean gradients */
int  stabgpu_mean_gradients(int ny, int wallt, const double* vm, const double* deta, const double* d2eta,
                            double* D1, double* D2, double* Dt2w, double* g2vm, double* g22vm);
/* circh.f90:35-188 (curve=2): x_inout is the radius on entry and 0 on return; h5 = ny x 5 col-major */
int  stabgpu_circh(double* x_inout, int ny, const double* y, double* h5);

/* ---- the hot path ------------------------------------------------------------------------------ */
/* Replaces temporal.f90:95-879 for npts points (alpha[p], beta[p]) sharing one mean profile:
 * assembly of A0,B0, B0^-1 A0 (ZGESV), eigenvalues [+ right vectors] (ZGEEV), stable sort by Im,
 * max-|.| scaling of vectors.  omg: n x npts, evec: n x n x npts (NULL unless want_vectors), n = 5 ny.
 * g2vm/g22vm: used only when p->ider == 0.  Re_pt / Ma_pt: optional per-point overrides (NULL). */
int stabgpu_temporal_batch(const stabgpu_params* p, const double* vm, const double* g2vm, const double* g22vm,
                           const double* deta, const double* d2eta,
                           int npts, const double* alpha /* 2*npts */, const double* beta /* 2*npts */,
                           const double* Re_pt, const double* Ma_pt, int want_vectors,
                           double* omg /* 2*n*npts */, double* evec /* 2*n*n*npts or NULL */, int* info);

/* Replaces spatial.f90:96-1084 for npts points (omega[p], beta[p]): C0,C1,C2, LU reduction
 * (ZGETRF + 2 ZGETRS), 2n x 2n companion, ZGEEV, alpha = 1/lambda, stable sort by Im(alpha).
 * h5: ny x 5 curvature metrics (NULL = flat).  alp: 2n x npts; evec: 2n x 2n x npts (not rescaled). */
int stabgpu_spatial_batch(const stabgpu_params* p, const double* vm, const double* g2vm, const double* g22vm,
                          const double* deta, const double* d2eta, const double* h5,
                          int npts, const double* omega, const double* beta,
                          const double* Re_pt, const double* Ma_pt, int want_vectors,
                          double* alp, double* evec, int* info);

/* Generic batched complex eigensolver on caller-supplied matrices: the ZGEEV('N','V'|'N') of
 * temporal.f90:803 / spatial.f90:1043 in isolation (roofline probes, SURVEY 8d config Cr).
 * A: n x n x batch (not modified); w: n x batch in ZGEEV-like (unsorted) order; V: n x n x batch or NULL. */
int stabgpu_zgeev_batch(int n, int batch, const double* A, int want_vectors, double* w, double* V, int* info);

/* Stage (4) of the north star, batched: polish ONE mode per sweep point by shift-invert (residual) inverse iteration
 * on the operator polynomial of that point, P(l) = A0 - l B0 (kind 1, temporal.f90:622-752; l = omega) or
 * P(l) = C0 + l C1 + l^2 C2 (kind 2, spatial.f90:681-1016; l = alpha, x = the bottom half of the companion eigenvector).
 * New functionality: the reference polishes with the external `shoot` (README.md:3-7, thesis/TStest/run.sh:30); getevec only
 * selects a mode of the full spectrum (getevec.f90:154-222).  Per point: P(sigma_p) is factored once by the batched blocked LU
 * (DMMA rank-32 updates), then every iteration is three n x n matrix-vector products and one replay of the factorization.
 *   s1/s2: (alpha|omega, beta) per point as in the batch calls; sigma: the shift per point (e.g. the neighbouring point's
 *   eigenvalue); x0: n x npts start vectors or NULL; h5: spatial curvature metrics or NULL.
 *   lambda: polished eigenvalue per point; x: n x npts eigenvectors scaled as temporal.f90:867-879 (or NULL);
 *   resid: |P(lambda) x| / (|M0 x| + |lambda||M1 x| + |lambda|^2 |M2 x|); iters: iterations used, or -(k+1) when the
 *   k-th pivot of P(sigma) is exactly zero (sigma is an eigenvalue to working precision).  Shards over the devices of
 *   stabgpu_init_multi like the batch calls. */
int stabgpu_polish_batch(int kind, const stabgpu_params* p, const double* vm, const double* g2vm, const double* g22vm,
                         const double* deta, const double* d2eta, const double* h5,
                         int npts, const double* s1, const double* s2, const double* Re_pt, const double* Ma_pt,
                         const double* sigma /* 2*npts */, const double* x0 /* 2*n*npts or NULL */,
                         int max_iters, double tol,
                         double* lambda /* 2*npts */, double* x /* 2*n*npts or NULL */, double* resid, int* iters);
/* one temporal point through the same path (kept for the round-1 callers) */
int stabgpu_temporal_polish(const stabgpu_params* p, const double* vm, const double* g2vm, const double* g22vm,
                            const double* deta, const double* d2eta, const double* alpha, const double* beta,
                            const double* sigma, const double* x0, int max_iters, double tol,
                            double* lambda, double* x, double* resid, int* iters);

/* ---- inspection entry points (parity tests of the individual stages) --------------------------- */
/* A0, B0 of temporal.f90:622-752 for one point (n x n each). */
int stabgpu_temporal_assemble(const stabgpu_params* p, const double* vm, const double* g2vm, const double* g22vm,
                              const double* deta, const double* d2eta, const double* alpha, const double* beta,
                              double* A0, double* B0);
/* C0, C1, C2 of spatial.f90:681-959 for one point (n x n each, signs as in the reference). */
int stabgpu_spatial_assemble(const stabgpu_params* p, const double* vm, const double* g2vm, const double* g22vm,
                             const double* deta, const double* d2eta, const double* h5,
                             const double* omega, const double* beta, double* C0, double* C1, double* C2);
/* Balanced matrix, scale, ilo/ihi (0-based), Hessenberg form + reflectors, tau for one n x n matrix. */
int stabgpu_debug_stages(int n, const double* A, double* balanced, double* scale, int* ilo, int* ihi,
                         double* hess, double* tau);

/* ---- device-resident plan API (what the batch calls are built from; used by bench.py) ---------- */
typedef struct stabgpu_plan stabgpu_plan;
/* kind: 1 temporal, 2 spatial.  The plan owns device copies of the profile/grid and workspaces for
 * up to max_pts points per wave. */
int stabgpu_plan_create(stabgpu_plan** plan, int kind, const stabgpu_params* p, const double* vm,
                        const double* g2vm, const double* g22vm, const double* deta, const double* d2eta,
                        const double* h5, int max_pts, int want_vectors);
int stabgpu_plan_upload(stabgpu_plan* plan, int npts, const double* s1 /* alpha|omega */, const double* s2 /* beta */,
                        const double* Re_pt, const double* Ma_pt);              /* H2D of the sweep values */
int stabgpu_plan_execute(stabgpu_plan* plan);                                    /* kernels only: enqueue + wait */
int stabgpu_plan_enqueue(stabgpu_plan* plan);                                    /* launches every kernel of one pass on the plan's stream and returns */
int stabgpu_plan_wait(stabgpu_plan* plan);                                       /* waits for the enqueued pass; fills the stage times */
int stabgpu_plan_download(stabgpu_plan* plan, double* eig, double* evec, int* info); /* D2H */
int stabgpu_plan_stage_times(stabgpu_plan* plan, float* ms /* 8 floats */);      /* CUDA-event time per stage of the last execute */
long long stabgpu_plan_launch_count(stabgpu_plan* plan);
/* measurement aid: with enable=1 the next execute records a CUDA event after every kernel of the Hessenberg stage; ms4 receives
 * the time per kernel class of the last profiled execute: panel step, GEMV (HBM bound), tensor-core block updates, other */
/* (ilo, ihi) of ZGEBAL per point of the last execute, 0-based inclusive: 2*npts ints (sizes the algorithmic work) */
int stabgpu_plan_ilohi(stabgpu_plan* plan, int* ilohi);
int stabgpu_plan_profile_hessenberg(stabgpu_plan* plan, int enable, float* ms4);
int stabgpu_plan_profile_eigvec(stabgpu_plan* plan, float* ms3);   /* inverse iteration, back-transformation GEMMs, finalize (same profiled execute) */                         /* kernels launched by the last execute */
void* stabgpu_plan_stream(stabgpu_plan* plan);                                   /* the cudaStream_t the plan launches on (for external CUDA-event timing) */
void* stabgpu_plan_eig_dev(stabgpu_plan* plan);                                  /* DEVICE pointer: sorted eigenvalues, N x npts complex (for the multi-GPU result gather) */
int stabgpu_plan_capacity(stabgpu_plan* plan);                                   /* points per wave that fit the device workspace */
int stabgpu_plan_destroy(stabgpu_plan* plan);

/* ---- sweep drivers and file formats (host) ------------------------------------------------------ */
/* mtemporal.f90:25-39: point enumeration (upper end excluded, quirk q6). Returns npts, or -1 for a zero / non-finite
 * increment or more than 10^7 points (the reference divides by the increment unguarded). */
int stabgpu_mtemporal_points(double amin, double amax, double ainc, double bmin, double bmax, double binc,
                             double* alpha_r, double* beta_r, int max_pts);
/* mspatial.f90:68-96: upper end included. */
int stabgpu_mspatial_points(double omin, double omax, double oinc, double bmin, double bmax, double binc,
                            double* omega_r, double* beta_r, int max_pts);
/* contiguous shard [lo,hi) of npts points for rank r of w */
void stabgpu_shard_range(int npts, int rank, int world, int* lo, int* hi);
/* gfortran sequential-unformatted eigensystem file (temporal.f90:883-890 / spatial.f90:1120-1126) */
int stabgpu_write_eig_file(const char* path, const stabgpu_params* p, int itype, int ind,
                           const double* omega, const double* alpha, const double* beta, double x,
                           const double* y, const double* eta, const double* deta, const double* d2eta,
                           const double* eig, const double* evec /* NULL: record 5 omitted */);

#ifdef __cplusplus
}
#endif
#endif /* STABGPU_H */
