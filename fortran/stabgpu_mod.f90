!> ISO_C_BINDING interface to libstabgpu (include/stabgpu.h) for stab's Fortran drivers.
!> Replaces the inside of temporal (temporal.f90:95-879) and spatial (spatial.f90:96-1084);
!> see INTEGRATION.md for the patched call sites.  Not compiled in this repository's build image
!> (no Fortran compiler); field order and types mirror `struct stabgpu_params` exactly.
module stabgpu_mod
  use iso_c_binding
  implicit none

  type, bind(C) :: stabgpu_params
     integer(c_int) :: ny, mattyp, wallt, top, curve, ider, ievec, wall
     real(c_double) :: Ma, Re, Pr
     real(c_double) :: gamma, gamma1, cp
     real(c_double) :: Te, rmue, rlme, cone
     real(c_double) :: datmat(3)
     real(c_double) :: yi, ymax, x
  end type stabgpu_params

  interface
     integer(c_int) function stabgpu_init(device) bind(C, name='stabgpu_init')
       import :: c_int
       integer(c_int), value :: device
     end function stabgpu_init

     !> the drop-in case: one serial Fortran process, every GPU of the box behind each *_batch call
     integer(c_int) function stabgpu_init_multi(max_devices, ndev_used) bind(C, name='stabgpu_init_multi')
       import :: c_int
       integer(c_int), value       :: max_devices          ! <= 0: all visible devices
       integer(c_int), intent(out) :: ndev_used
     end function stabgpu_init_multi

     !> optional: page-lock a caller array once (c_loc(evec), bytes) so the eigenvectors arrive by direct DMA;
     !> without it a pageable destination is served through the library's pinned staging ring
     integer(c_int) function stabgpu_host_register(ptr, bytes) bind(C, name='stabgpu_host_register')
       import :: c_int, c_ptr, c_size_t
       type(c_ptr), value       :: ptr
       integer(c_size_t), value :: bytes
     end function stabgpu_host_register

     !> deflation window of the QR stage (default 32; 0 = classic deflation only) and ZLAQR0's NIBBLE in per cent (default 14)
     integer(c_int) function stabgpu_set_qr_deflation(window, nibble) bind(C, name='stabgpu_set_qr_deflation')
       import :: c_int
       integer(c_int), value :: window, nibble
     end function stabgpu_set_qr_deflation

     integer(c_int) function stabgpu_host_unregister(ptr) bind(C, name='stabgpu_host_unregister')
       import :: c_int, c_ptr
       type(c_ptr), value :: ptr
     end function stabgpu_host_unregister

     integer(c_int) function stabgpu_finalize() bind(C, name='stabgpu_finalize')
       import :: c_int
     end function stabgpu_finalize

     integer(c_int) function stabgpu_temporal_batch(p, vm, g2vm, g22vm, deta, d2eta, npts, alpha, beta, &
                                                    Re_pt, Ma_pt, want_vectors, omg, evec, info)       &
                                                    bind(C, name='stabgpu_temporal_batch')
       import :: c_int, c_double, c_double_complex, c_ptr, stabgpu_params
       type(stabgpu_params), intent(in) :: p
       real(c_double), intent(in)  :: vm(*), deta(*), d2eta(*)
       type(c_ptr), value          :: g2vm, g22vm, Re_pt, Ma_pt
       integer(c_int), value       :: npts, want_vectors
       complex(c_double_complex), intent(in)  :: alpha(*), beta(*)
       complex(c_double_complex), intent(out) :: omg(*)
       type(c_ptr), value          :: evec
       integer(c_int), intent(out) :: info(*)
     end function stabgpu_temporal_batch

     integer(c_int) function stabgpu_spatial_batch(p, vm, g2vm, g22vm, deta, d2eta, h5, npts, omega, beta, &
                                                   Re_pt, Ma_pt, want_vectors, alp, evec, info)            &
                                                   bind(C, name='stabgpu_spatial_batch')
       import :: c_int, c_double, c_double_complex, c_ptr, stabgpu_params
       type(stabgpu_params), intent(in) :: p
       real(c_double), intent(in)  :: vm(*), deta(*), d2eta(*)
       type(c_ptr), value          :: g2vm, g22vm, h5, Re_pt, Ma_pt
       integer(c_int), value       :: npts, want_vectors
       complex(c_double_complex), intent(in)  :: omega(*), beta(*)
       complex(c_double_complex), intent(out) :: alp(*)
       type(c_ptr), value          :: evec
       integer(c_int), intent(out) :: info(*)
     end function stabgpu_spatial_batch

     !> stage (4): one mode per sweep point polished from a shift sigma(p); kind 1 temporal (omega), 2 spatial (alpha)
     integer(c_int) function stabgpu_polish_batch(kind, p, vm, g2vm, g22vm, deta, d2eta, h5, npts, s1, s2, Re_pt, Ma_pt, &
                                                  sigma, x0, max_iters, tol, lambda, x, resid, iters)                   &
                                                  bind(C, name='stabgpu_polish_batch')
       import :: c_int, c_double, c_double_complex, c_ptr, stabgpu_params
       integer(c_int), value       :: kind
       type(stabgpu_params), intent(in) :: p
       real(c_double), intent(in)  :: vm(*), deta(*), d2eta(*)
       type(c_ptr), value          :: g2vm, g22vm, h5, Re_pt, Ma_pt, x0, x   ! c_null_ptr when absent
       integer(c_int), value       :: npts, max_iters
       complex(c_double_complex), intent(in)  :: s1(*), s2(*), sigma(*)
       real(c_double), value       :: tol
       complex(c_double_complex), intent(out) :: lambda(*)
       real(c_double), intent(out) :: resid(*)
       integer(c_int), intent(out) :: iters(*)
     end function stabgpu_polish_batch

     integer(c_int) function stabgpu_temporal_polish(p, vm, g2vm, g22vm, deta, d2eta, alpha, beta, sigma, x0, &
                                                     max_iters, tol, lambda, x, resid, iters)                 &
                                                     bind(C, name='stabgpu_temporal_polish')
       import :: c_int, c_double, c_double_complex, c_ptr, stabgpu_params
       type(stabgpu_params), intent(in) :: p
       real(c_double), intent(in)  :: vm(*), deta(*), d2eta(*)
       type(c_ptr), value          :: g2vm, g22vm, x0
       complex(c_double_complex), intent(in)  :: alpha, beta, sigma
       integer(c_int), value       :: max_iters
       real(c_double), value       :: tol
       complex(c_double_complex), intent(out) :: lambda, x(*)
       real(c_double), intent(out) :: resid
       integer(c_int), intent(out) :: iters
     end function stabgpu_temporal_polish
  end interface
end module stabgpu_mod
