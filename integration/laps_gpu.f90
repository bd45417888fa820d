! laps_gpu.f90 -- ISO_C_BINDING interface of the laps_b200 C ABI (include/laps_b200.h, ABI version 5).
! Drop this file into src_compressible/ (or any of the other three source trees) and patch mhd.f90 as
! INTEGRATION.md section 2 describes.  Not compiled in this repository's image (no Fortran compiler);
! the same ABI with the same array layouts is exercised through laps_b200/capi.py.
module laps_gpu
  use iso_c_binding
  implicit none
  integer(c_int), parameter :: LAPS_ABI_VERSION = 6, LAPS_PEER_BLOB_BYTES = 256

  type, bind(C) :: laps_params          ! field order = include/laps_b200.h
    integer(c_int32_t) :: abi_version
    integer(c_int32_t) :: nx, ny, nz
    real(c_double)     :: Lx, Ly, Lz
    real(c_double)     :: adiabatic_index
    integer(c_int32_t) :: if_resis, if_resis_exp
    real(c_double)     :: resistivity
    integer(c_int32_t) :: if_visc, if_visc_exp
    real(c_double)     :: viscosity
    integer(c_int32_t) :: if_conserve_background
    real(c_double)     :: cfl
    integer(c_int32_t) :: dealias_option
    real(c_double)     :: afx, afy, afz
    integer(c_int32_t) :: if_AEB, if_corotating
    real(c_double)     :: radius0, Ur0, corotating_angle
    integer(c_int32_t) :: if_hall
    real(c_double)     :: ion_inertial_length
    integer(c_int32_t) :: rank, nranks
    integer(c_int32_t) :: device
    integer(c_int32_t) :: ndim                  ! 3 (or 0): 3D trees; 2: src_compressible/2D (nz = 1)
    integer(c_int32_t) :: if_z_radial           ! 2D/mhd.f90:44
    integer(c_int32_t) :: if_limit_dt_increase  ! 2D/mhd.f90:23
    integer(c_int32_t) :: incompressible        ! 1: src_incompressible (uu(8) = pressure)
    real(c_double)     :: rho0                  ! src_incompressible/mhdinit.f90:15
    integer(c_int32_t) :: if_external_force     ! 2D/mhd.f90:43 (&pert), 2D compressible tree only
  end type

  type, bind(C) :: laps_extents
    integer(c_int32_t) :: nx, ny, nz, nxh, z_offset, z_size, y_offset, y_size, y_stride
  end type

  type(c_ptr) :: gpu = c_null_ptr        ! the handle

  interface
    integer(c_int) function laps_create(p, h) bind(C, name='laps_create')
      import; type(laps_params), intent(in) :: p; type(c_ptr), intent(out) :: h
    end function
    integer(c_int) function laps_destroy(h) bind(C, name='laps_destroy')
      import; type(c_ptr), value :: h
    end function
    type(c_ptr) function laps_last_error(h) bind(C, name='laps_last_error')
      import; type(c_ptr), value :: h
    end function
    integer(c_int) function laps_get_extents(h, e) bind(C, name='laps_get_extents')
      import; type(c_ptr), value :: h; type(laps_extents), intent(out) :: e
    end function
    integer(c_int) function laps_export_peer_blob(h, blob) bind(C, name='laps_export_peer_blob')
      import; type(c_ptr), value :: h; character(kind=c_char) :: blob(*)
    end function
    integer(c_int) function laps_import_peer_blobs(h, blobs) bind(C, name='laps_import_peer_blobs')
      import; type(c_ptr), value :: h; character(kind=c_char) :: blobs(*)
    end function
    integer(c_int) function laps_set_primitive(h, uu) bind(C, name='laps_set_primitive')
      import; type(c_ptr), value :: h; real(c_double), intent(in) :: uu(*)
    end function
    integer(c_int) function laps_set_time(h, t) bind(C, name='laps_set_time')
      import; type(c_ptr), value :: h; real(c_double), value :: t
    end function
    integer(c_int) function laps_vardt(h, dt) bind(C, name='laps_vardt')
      import; type(c_ptr), value :: h; real(c_double), intent(inout) :: dt
    end function
    integer(c_int) function laps_evolve(h) bind(C, name='laps_evolve')
      import; type(c_ptr), value :: h
    end function
    integer(c_int) function laps_max_divb(h, v) bind(C, name='laps_max_divb')
      import; type(c_ptr), value :: h; real(c_double), intent(out) :: v
    end function
    integer(c_int) function laps_rms(h, out19) bind(C, name='laps_rms')
      import; type(c_ptr), value :: h; real(c_double), intent(out) :: out19(19)
    end function
    integer(c_int) function laps_get_state(h, uu, uu_prim) bind(C, name='laps_get_state')
      import; type(c_ptr), value :: h; real(c_double) :: uu(*), uu_prim(*)
    end function
    integer(c_int) function laps_get_output(h, out8, primitive) bind(C, name='laps_get_output')
      import; type(c_ptr), value :: h; real(c_double) :: out8(*); integer(c_int32_t), value :: primitive
    end function
    integer(c_int) function laps_step(h, time, dt) bind(C, name='laps_step')
      import; type(c_ptr), value :: h; real(c_double), intent(inout) :: time, dt
    end function
    integer(c_int) function laps_set_primitive_modes(h, nmodes, k, coef, background) bind(C, name='laps_set_primitive_modes')
      import; type(c_ptr), value :: h; integer(c_int32_t), value :: nmodes
      integer(c_int32_t), intent(in) :: k(3,*); complex(c_double_complex), intent(in) :: coef(nmodes,7)
      real(c_double), intent(in) :: background(8)
    end function
    ! incompressible driver only (src_incompressible/mhd.f90:157-161,620-732)
    integer(c_int) function laps_max_divv(h, v) bind(C, name='laps_max_divv')
      import; type(c_ptr), value :: h; real(c_double), intent(out) :: v
    end function
    integer(c_int) function laps_max_div_real(h, out2) bind(C, name='laps_max_div_real')
      import; type(c_ptr), value :: h; real(c_double), intent(out) :: out2(2)
    end function
    integer(c_int) function laps_get_rho0(h, rho0) bind(C, name='laps_get_rho0')
      import; type(c_ptr), value :: h; real(c_double), intent(out) :: rho0
    end function
    integer(c_int) function laps_check_nan(h, is_nan) bind(C, name='laps_check_nan')
      import; type(c_ptr), value :: h; integer(c_int32_t), intent(out) :: is_nan
    end function
    integer(c_int) function laps_set_external_force(h, force_local) bind(C, name='laps_set_external_force')
      import; type(c_ptr), value :: h; real(c_double), intent(in) :: force_local(*)
    end function
    ! ---- optional: fixed time step, synchronisation, diagnostics beyond the reference's, unit-level parity, measurement
    integer(c_int) function laps_rkt_init(h, dt) bind(C, name='laps_rkt_init')          ! rktmod.f90:15-32 with a given dt
      import; type(c_ptr), value :: h; real(c_double), value :: dt
    end function
    integer(c_int) function laps_sync(h) bind(C, name='laps_sync')
      import; type(c_ptr), value :: h
    end function
    integer(c_int) function laps_get_stream(h, stream) bind(C, name='laps_get_stream')
      import; type(c_ptr), value :: h; type(c_ptr), intent(out) :: stream
    end function
    integer(c_int) function laps_connect_local(handles, nranks) bind(C, name='laps_connect_local')   ! one process driving several GPUs
      import; type(c_ptr), intent(inout) :: handles(*); integer(c_int32_t), value :: nranks
    end function
    integer(c_int) function laps_invariants(h, out3) bind(C, name='laps_invariants')    ! mean energy, mean u.B, max |k.B^|
      import; type(c_ptr), value :: h; real(c_double), intent(out) :: out3(3)
    end function
    integer(c_int) function laps_get_spectral(h, uu_fourier_local) bind(C, name='laps_get_spectral')
      import; type(c_ptr), value :: h; complex(c_double_complex), intent(out) :: uu_fourier_local(*)
    end function
    integer(c_int) function laps_fft_forward(h, real_fields, nfields, spec_out) bind(C, name='laps_fft_forward')   ! fftw.f90:42-71
      import; type(c_ptr), value :: h; real(c_double), intent(in) :: real_fields(*); integer(c_int32_t), value :: nfields
      complex(c_double_complex), intent(out) :: spec_out(*)
    end function
    integer(c_int) function laps_fft_inverse(h, spec_in, nfields, real_out) bind(C, name='laps_fft_inverse')      ! fftw.f90:73-103
      import; type(c_ptr), value :: h; complex(c_double_complex), intent(in) :: spec_in(*); integer(c_int32_t), value :: nfields
      real(c_double), intent(out) :: real_out(*)
    end function
    integer(c_int) function laps_transpose_yz_indexmap(h, pairs) bind(C, name='laps_transpose_yz_indexmap')    ! parallel.f90:273-297
      import; type(c_ptr), value :: h; integer(c_int64_t), intent(out) :: pairs(2,*)
    end function
    integer(c_int) function laps_transpose_zy_indexmap(h, pairs) bind(C, name='laps_transpose_zy_indexmap')    ! parallel.f90:300-324
      import; type(c_ptr), value :: h; integer(c_int64_t), intent(out) :: pairs(2,*)
    end function
    integer(c_int) function laps_get_pruning(h, nkx, kymax, nky_local) bind(C, name='laps_get_pruning')
      import; type(c_ptr), value :: h; integer(c_int32_t), intent(out) :: nkx, kymax, nky_local
    end function
    integer(c_int) function laps_get_pruning_counts(h, live_columns, live_modes) bind(C, name='laps_get_pruning_counts')
      import; type(c_ptr), value :: h; integer(c_int64_t), intent(out) :: live_columns, live_modes
    end function
    integer(c_int) function laps_get_field_counts(h, nf, ni, spec_rows) bind(C, name='laps_get_field_counts')
      import; type(c_ptr), value :: h; integer(c_int32_t), intent(out) :: nf, ni, spec_rows
    end function
    integer(c_int) function laps_last_step_ms(h, ms, launches) bind(C, name='laps_last_step_ms')
      import; type(c_ptr), value :: h; real(c_float), intent(out) :: ms; integer(c_int32_t), intent(out) :: launches
    end function
    integer(c_int) function laps_set_profiling(h, on) bind(C, name='laps_set_profiling')
      import; type(c_ptr), value :: h; integer(c_int32_t), value :: on
    end function
    integer(c_int) function laps_get_output_async(h, out_local, primitive) bind(C, name='laps_get_output_async')
      import; type(c_ptr), value :: h; real(c_double), intent(out) :: out_local(*); integer(c_int32_t), value :: primitive
    end function
    integer(c_int) function laps_output_wait(h) bind(C, name='laps_output_wait')
      import; type(c_ptr), value :: h
    end function
    integer(c_int) function laps_get_profile_bytes(h, bytes, cap, count) bind(C, name='laps_get_profile_bytes')
      import; type(c_ptr), value :: h; real(c_double), intent(out) :: bytes(*); integer(c_int32_t), value :: cap
      integer(c_int32_t), intent(out) :: count
    end function
    integer(c_int) function laps_get_footprint(h, device_bytes) bind(C, name='laps_get_footprint')
      import; type(c_ptr), value :: h; integer(c_int64_t), intent(out) :: device_bytes
    end function
    integer(c_int) function laps_set_tune(h, name, value) bind(C, name='laps_set_tune')
      import; type(c_ptr), value :: h; character(kind=c_char), intent(in) :: name(*); integer(c_int32_t), value :: value
    end function
    integer(c_int) function laps_get_profile(h, names, ms, cap, count) bind(C, name='laps_get_profile')
      import; type(c_ptr), value :: h; character(kind=c_char), intent(out) :: names(32,*); real(c_float), intent(out) :: ms(*)
      integer(c_int32_t), value :: cap; integer(c_int32_t), intent(out) :: count
    end function
  end interface
end module laps_gpu
