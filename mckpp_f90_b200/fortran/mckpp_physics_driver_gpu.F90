! mckpp_physics_driver_gpu.F90 -- drop-in replacement of the reference's
! src/mckpp_physics_driver_mod.F90 that runs the column physics on B200 through the
! C ABI of include/kpp_gpu.h.
!
! NOT COMPILED IN THIS REPOSITORY'S IMAGE: there is no Fortran compiler here (no
! gfortran/flang/nvfortran).  It is delivered as source for a maintainer of
! aosprey/mckpp-f90 (see INTEGRATION.md).  The same C entry points are exercised by the
! Python ctypes host (mckpp_f90_b200/capi.py) and the C++ host (host/), which use the
! identical memory image.
!
! Public interface kept from the reference:
!   SUBROUTINE mckpp_physics_driver()          (no arguments; works on the module globals
!                                               kpp_3d_fields / kpp_const_fields)
! New:
!   SUBROUTINE mckpp_physics_gpu_initialize()  call once after mckpp_initialize_fields
!   SUBROUTINE mckpp_physics_gpu_pull(id, a)   bring a member back before it is read on the host
!   SUBROUTINE mckpp_physics_gpu_pack(out_id, a)            one xios_send_field block, packed on the device
!   SUBROUTINE mckpp_physics_gpu_clim_records / _clim_blend  climatology time interpolation on the device
!   SUBROUTINE mckpp_physics_gpu_finalize()
!
! Build: add -DMCKPP_GPU and link libkpp_gpu.so; REAL must be 8 bytes
! (-fdefault-real-8 / -s real64, as both shipped configs already do).
MODULE mckpp_physics_driver_mod

  USE, INTRINSIC :: iso_c_binding
  USE mckpp_data_fields, ONLY: kpp_3d_fields, kpp_const_fields
  USE mckpp_parameters
  USE mckpp_time_control, ONLY: ntime
  USE mckpp_log_messages, ONLY: mckpp_print_warning, mckpp_print_error, max_message_len
  USE mckpp_abort_mod, ONLY: mckpp_abort
  USE mckpp_timer, ONLY: mckpp_start_timer, mckpp_stop_timer

  IMPLICIT NONE

  ! ---- mirror of the C structs (include/kpp_gpu.h)
  TYPE, BIND(C) :: kpp_dims
    INTEGER(c_int32_t) :: npts, nz, nztmax, nsflxs, njdt, maxmodeadv
  END TYPE kpp_dims

  TYPE, BIND(C) :: kpp_consts
    REAL(c_double) :: dto, grav, vonk, sice, hmixtolfrac, iso_thresh
    INTEGER(c_int32_t) :: itermax, iso_bot, dt_uvdamp, LKPP, LRI, LDD, L_SSref, &
        L_RELAX_SST, L_RELAX_CALCONLY, L_FCORR, L_FCORR_WITHZ, L_SFCORR, L_SFCORR_WITHZ, &
        L_RELAX_SAL, L_RELAX_OCNT, L_NO_FREEZE, L_NO_ISOTHERM, L_DAMP_CURR, L_VARY_BOTTOM_TEMP, &
        have_ocnT_file, have_sal_file, numerics, reserved
  END TYPE kpp_consts

  TYPE, BIND(C) :: kpp_step_report
    INTEGER(c_int32_t) :: ntime, n_active, n_long_iter, n_reint, n_reint_fail, n_reset, &
        n_pivot_zero, n_iter_cap, max_iter, n_handed_over
    INTEGER(c_int64_t) :: sum_iter
    REAL(c_float) :: kernel_ms, reserved2
  END TYPE kpp_step_report

  ! field ids: enum kpp_field_id, in declaration order (include/kpp_gpu.h)
  INTEGER(c_int), PARAMETER :: KPP_F_U=0, KPP_F_X=1, KPP_F_US=2, KPP_F_XS=3, KPP_F_HMIXD=4, &
      KPP_F_OLD=5, KPP_F_NEW=6, KPP_F_HMIX=7, KPP_F_KMIX=8, KPP_F_TREF=9, KPP_F_UREF=10, &
      KPP_F_VREF=11, KPP_F_SSURF=12, KPP_F_SREF=13, KPP_F_SSREF=14, KPP_F_F=15, KPP_F_OCDEPTH=16, &
      KPP_F_JERLOV=17, KPP_F_L_OCEAN=18, KPP_F_RUN_PHYSICS=19, KPP_F_SFLUX=20, KPP_F_U_INIT=21, &
      KPP_F_RELAX_SST=22, KPP_F_SST0=23, KPP_F_FCORR_TWOD=24, KPP_F_FCORR=25, KPP_F_RELAX_SAL=26, &
      KPP_F_RELAX_OCNT=27, KPP_F_SAL_CLIM=28, KPP_F_OCNT_CLIM=29, KPP_F_FCORR_WITHZ=30, &
      KPP_F_SFCORR_WITHZ=31, KPP_F_BOTTOM_TEMP=32, KPP_F_NMODEADV=33, KPP_F_MODEADV=34, &
      KPP_F_ADVECTION=35, KPP_F_FREEZE_FLAG=36, KPP_F_RESET_FLAG=37, KPP_F_DAMPU_FLAG=38, &
      KPP_F_DAMPV_FLAG=39, KPP_F_RHO=40, KPP_F_CP=41, KPP_F_BUOY=42, KPP_F_RIG=43, KPP_F_DBLOC=44, &
      KPP_F_SHSQ=45, KPP_F_DIFM=46, KPP_F_DIFS=47, KPP_F_DIFT=48, KPP_F_GHAT=49, KPP_F_WU=50, &
      KPP_F_WX=51, KPP_F_WXNT=52, KPP_F_TINC_FCORR=53, KPP_F_SINC_FCORR=54, KPP_F_OCNTCORR=55, &
      KPP_F_SCORR=56, KPP_F_SWFRAC=57, KPP_F_SWDK_OPT=58

  INTEGER(c_int), PARAMETER :: KPP_E_PIVOT_ZERO = -4

  ! output sets packed on the device (enum kpp_out_id; SURVEY 8 f2)
  INTEGER(c_int), PARAMETER :: KPP_OUT_U=0, KPP_OUT_V=1, KPP_OUT_T=2, KPP_OUT_S=3, KPP_OUT_B=4, &
      KPP_OUT_WU=5, KPP_OUT_WV=6, KPP_OUT_WT=7, KPP_OUT_WS=8, KPP_OUT_WB=9, KPP_OUT_WTNT=10, &
      KPP_OUT_DIFM=11, KPP_OUT_DIFT=12, KPP_OUT_DIFS=13, KPP_OUT_RHO=14, KPP_OUT_CP=15, &
      KPP_OUT_SCORR=16, KPP_OUT_RIG=17, KPP_OUT_DBLOC=18, KPP_OUT_SHSQ=19, KPP_OUT_TINC_FCORR=20, &
      KPP_OUT_FCORR_Z=21, KPP_OUT_SINC_FCORR=22, KPP_OUT_HMIX=23, KPP_OUT_FCORR=24, &
      KPP_OUT_TAUX_IN=25, KPP_OUT_TAUY_IN=26, KPP_OUT_SOLAR_IN=27, KPP_OUT_NSOLAR_IN=28, &
      KPP_OUT_PMINUSE_IN=29, KPP_OUT_FREEZE_FLAG=30, KPP_OUT_COMP_FLAG=31, KPP_OUT_DAMPU_FLAG=32, &
      KPP_OUT_DAMPV_FLAG=33, KPP_OUT_R_UVEL=34, KPP_OUT_R_VVEL=35, KPP_OUT_R_T=36, KPP_OUT_R_S=37, &
      KPP_OUT_R_CP=38, KPP_OUT_R_RHO=39, KPP_OUT_R_HMIX=40, KPP_OUT_R_KMIX=41, KPP_OUT_R_SREF=42, &
      KPP_OUT_R_SSREF=43, KPP_OUT_R_SSURF=44, KPP_OUT_R_TREF=45, KPP_OUT_R_OLD=46, &
      KPP_OUT_R_NEW=47, KPP_OUT_R_US=48, KPP_OUT_R_VS=49, KPP_OUT_R_TS=50, KPP_OUT_R_SS=51, &
      KPP_OUT_R_HMIXD=52

  INTERFACE
    INTEGER(c_int) FUNCTION kpp_gpu_create(dims, consts, zm, hm, dm, tri, wmt, wst, device, h) &
        BIND(C, name="kpp_gpu_create")
      IMPORT :: c_int, c_double, c_ptr, kpp_dims, kpp_consts
      TYPE(kpp_dims), INTENT(IN) :: dims
      TYPE(kpp_consts), INTENT(IN) :: consts
      REAL(c_double), INTENT(IN) :: zm(*), hm(*), dm(*), tri(*), wmt(*), wst(*)
      INTEGER(c_int), VALUE :: device
      TYPE(c_ptr), INTENT(OUT) :: h
    END FUNCTION
    INTEGER(c_int) FUNCTION kpp_gpu_destroy(h) BIND(C, name="kpp_gpu_destroy")
      IMPORT :: c_int, c_ptr
      TYPE(c_ptr), VALUE :: h
    END FUNCTION
    ! assumed-size dummies: a contiguous ALLOCATABLE component is passed by base address, so
    ! kpp_3d_fields%X etc. go through unchanged (no TARGET attribute, no copy)
    INTEGER(c_int) FUNCTION kpp_gpu_upload_r8(h, id, a, bytes) BIND(C, name="kpp_gpu_upload_field")
      IMPORT :: c_int, c_ptr, c_double, c_size_t
      TYPE(c_ptr), VALUE :: h
      INTEGER(c_int), VALUE :: id
      REAL(c_double), INTENT(IN) :: a(*)
      INTEGER(c_size_t), VALUE :: bytes
    END FUNCTION
    INTEGER(c_int) FUNCTION kpp_gpu_upload_i4(h, id, a, bytes) BIND(C, name="kpp_gpu_upload_field")
      IMPORT :: c_int, c_ptr, c_int32_t, c_size_t
      TYPE(c_ptr), VALUE :: h
      INTEGER(c_int), VALUE :: id
      INTEGER(c_int32_t), INTENT(IN) :: a(*)
      INTEGER(c_size_t), VALUE :: bytes
    END FUNCTION
    INTEGER(c_int) FUNCTION kpp_gpu_download_r8(h, id, a, bytes) BIND(C, name="kpp_gpu_download_field")
      IMPORT :: c_int, c_ptr, c_double, c_size_t
      TYPE(c_ptr), VALUE :: h
      INTEGER(c_int), VALUE :: id
      REAL(c_double), INTENT(INOUT) :: a(*)
      INTEGER(c_size_t), VALUE :: bytes
    END FUNCTION
    INTEGER(c_int) FUNCTION kpp_gpu_download_i4(h, id, a, bytes) BIND(C, name="kpp_gpu_download_field")
      IMPORT :: c_int, c_ptr, c_int32_t, c_size_t
      TYPE(c_ptr), VALUE :: h
      INTEGER(c_int), VALUE :: id
      INTEGER(c_int32_t), INTENT(INOUT) :: a(*)
      INTEGER(c_size_t), VALUE :: bytes
    END FUNCTION
    INTEGER(c_int) FUNCTION kpp_gpu_upload_forcing(h, sflux6) BIND(C, name="kpp_gpu_upload_forcing")
      IMPORT :: c_int, c_ptr, c_double
      TYPE(c_ptr), VALUE :: h
      REAL(c_double), INTENT(IN) :: sflux6(*)
    END FUNCTION
    INTEGER(c_int) FUNCTION kpp_gpu_init_vmix(h) BIND(C, name="kpp_gpu_init_vmix")
      IMPORT :: c_int, c_ptr
      TYPE(c_ptr), VALUE :: h
    END FUNCTION
    INTEGER(c_int) FUNCTION kpp_gpu_step(h, nt) BIND(C, name="kpp_gpu_step")
      IMPORT :: c_int, c_ptr
      TYPE(c_ptr), VALUE :: h
      INTEGER(c_int), VALUE :: nt
    END FUNCTION
    INTEGER(c_int) FUNCTION kpp_gpu_pack_output(h, out_id, a, bytes) BIND(C, name="kpp_gpu_pack_output")
      IMPORT :: c_int, c_ptr, c_double, c_size_t
      TYPE(c_ptr), VALUE :: h
      INTEGER(c_int), VALUE :: out_id
      REAL(c_double), INTENT(OUT) :: a(*)
      INTEGER(c_size_t), VALUE :: bytes
    END FUNCTION
    INTEGER(c_int) FUNCTION kpp_gpu_upload_clim_record(h, id, which, a, bytes) &
        BIND(C, name="kpp_gpu_upload_clim_record")
      IMPORT :: c_int, c_ptr, c_double, c_size_t
      TYPE(c_ptr), VALUE :: h
      INTEGER(c_int), VALUE :: id, which
      REAL(c_double), INTENT(IN) :: a(*)
      INTEGER(c_size_t), VALUE :: bytes
    END FUNCTION
    INTEGER(c_int) FUNCTION kpp_gpu_blend_clim(h, id, prev_weight, next_weight) BIND(C, name="kpp_gpu_blend_clim")
      IMPORT :: c_int, c_ptr, c_double
      TYPE(c_ptr), VALUE :: h
      INTEGER(c_int), VALUE :: id
      REAL(c_double), VALUE :: prev_weight, next_weight
    END FUNCTION
    INTEGER(c_int) FUNCTION kpp_gpu_sync(h, rep) BIND(C, name="kpp_gpu_sync")
      IMPORT :: c_int, c_ptr, kpp_step_report
      TYPE(c_ptr), VALUE :: h
      TYPE(kpp_step_report), INTENT(OUT) :: rep
    END FUNCTION
  END INTERFACE

  TYPE(c_ptr), SAVE :: gpu = c_null_ptr
  INTEGER(c_size_t), PARAMETER :: r8 = 8_c_size_t, i4 = 4_c_size_t

CONTAINS

  PURE INTEGER(c_int32_t) FUNCTION l2i(l)
    LOGICAL, INTENT(IN) :: l
    l2i = MERGE(1_c_int32_t, 0_c_int32_t, l)
  END FUNCTION l2i

  ! Create the device mirror and push everything the physics reads.  Call after
  ! mckpp_initialize_fields (which has filled zm/hm/dm, wmt/wst, tri and the initial state;
  ! the reference's own initial vmix may be skipped: kpp_gpu_init_vmix does the same on device).
  SUBROUTINE mckpp_physics_gpu_initialize(device)
    INTEGER, INTENT(IN), OPTIONAL :: device
    TYPE(kpp_dims) :: d
    TYPE(kpp_consts) :: k
    INTEGER(c_int) :: rc, dev
    INTEGER(c_int32_t), ALLOCATABLE :: mask(:)
    CHARACTER(LEN=28) :: routine = "MCKPP_PHYSICS_GPU_INITIALIZE"

    dev = 0
    IF (PRESENT(device)) dev = device
    d%npts = npts; d%nz = nz; d%nztmax = nztmax; d%nsflxs = nsflxs; d%njdt = njdt; d%maxmodeadv = maxmodeadv
    k%dto = kpp_const_fields%dto; k%grav = kpp_const_fields%grav; k%vonk = kpp_const_fields%vonk
    k%sice = kpp_const_fields%sice; k%hmixtolfrac = hmixtolfrac; k%iso_thresh = kpp_const_fields%iso_thresh
    k%itermax = itermax; k%iso_bot = kpp_const_fields%iso_bot; k%dt_uvdamp = kpp_const_fields%dt_uvdamp
    k%LKPP = l2i(kpp_const_fields%LKPP); k%LRI = l2i(kpp_const_fields%LRI); k%LDD = l2i(kpp_const_fields%LDD)
    k%L_SSref = l2i(kpp_const_fields%L_SSref); k%L_RELAX_SST = l2i(kpp_const_fields%L_RELAX_SST)
    k%L_RELAX_CALCONLY = l2i(kpp_const_fields%L_RELAX_CALCONLY); k%L_FCORR = l2i(kpp_const_fields%L_FCORR)
    k%L_FCORR_WITHZ = l2i(kpp_const_fields%L_FCORR_WITHZ); k%L_SFCORR = l2i(kpp_const_fields%L_SFCORR)
    k%L_SFCORR_WITHZ = l2i(kpp_const_fields%L_SFCORR_WITHZ); k%L_RELAX_SAL = l2i(kpp_const_fields%L_RELAX_SAL)
    k%L_RELAX_OCNT = l2i(kpp_const_fields%L_RELAX_OCNT); k%L_NO_FREEZE = l2i(kpp_const_fields%L_NO_FREEZE)
    k%L_NO_ISOTHERM = l2i(kpp_const_fields%L_NO_ISOTHERM); k%L_DAMP_CURR = l2i(kpp_const_fields%L_DAMP_CURR)
    k%L_VARY_BOTTOM_TEMP = l2i(kpp_const_fields%L_VARY_BOTTOM_TEMP)
    k%have_ocnT_file = l2i(kpp_const_fields%ocnT_file .NE. 'none')
    k%have_sal_file = l2i(kpp_const_fields%sal_file .NE. 'none')
    k%numerics = 0          ! strict: same roundings as the CPU build
    k%reserved = 0

    rc = kpp_gpu_create(d, k, kpp_const_fields%zm, kpp_const_fields%hm, kpp_const_fields%dm, &
                        kpp_const_fields%tri, kpp_const_fields%wmt, kpp_const_fields%wst, dev, gpu)
    IF (rc .NE. 0) THEN
      CALL mckpp_print_error(routine, "kpp_gpu_create failed (no CUDA device? there is no CPU fallback)")
      CALL mckpp_abort()
    END IF

    ! REAL members: passed whole, the library moves the part the physics touches
    rc = kpp_gpu_upload_r8(gpu, KPP_F_U, kpp_3d_fields%U, r8*SIZE(kpp_3d_fields%U))
    rc = kpp_gpu_upload_r8(gpu, KPP_F_X, kpp_3d_fields%X, r8*SIZE(kpp_3d_fields%X))
    rc = kpp_gpu_upload_r8(gpu, KPP_F_US, kpp_3d_fields%Us, r8*SIZE(kpp_3d_fields%Us))
    rc = kpp_gpu_upload_r8(gpu, KPP_F_XS, kpp_3d_fields%Xs, r8*SIZE(kpp_3d_fields%Xs))
    rc = kpp_gpu_upload_r8(gpu, KPP_F_HMIXD, kpp_3d_fields%hmixd, r8*SIZE(kpp_3d_fields%hmixd))
    rc = kpp_gpu_upload_r8(gpu, KPP_F_U_INIT, kpp_3d_fields%U_init, r8*SIZE(kpp_3d_fields%U_init))
    rc = kpp_gpu_upload_r8(gpu, KPP_F_SSURF, kpp_3d_fields%Ssurf, r8*SIZE(kpp_3d_fields%Ssurf))
    rc = kpp_gpu_upload_r8(gpu, KPP_F_SREF, kpp_3d_fields%Sref, r8*SIZE(kpp_3d_fields%Sref))
    rc = kpp_gpu_upload_r8(gpu, KPP_F_SSREF, kpp_3d_fields%SSref, r8*SIZE(kpp_3d_fields%SSref))
    rc = kpp_gpu_upload_r8(gpu, KPP_F_F, kpp_3d_fields%f, r8*SIZE(kpp_3d_fields%f))
    rc = kpp_gpu_upload_r8(gpu, KPP_F_OCDEPTH, kpp_3d_fields%ocdepth, r8*SIZE(kpp_3d_fields%ocdepth))
    rc = kpp_gpu_upload_r8(gpu, KPP_F_SFLUX, kpp_3d_fields%sflux, r8*SIZE(kpp_3d_fields%sflux))
    rc = kpp_gpu_upload_r8(gpu, KPP_F_RELAX_SST, kpp_3d_fields%relax_sst, r8*SIZE(kpp_3d_fields%relax_sst))
    rc = kpp_gpu_upload_r8(gpu, KPP_F_SST0, kpp_3d_fields%SST0, r8*SIZE(kpp_3d_fields%SST0))
    rc = kpp_gpu_upload_r8(gpu, KPP_F_FCORR_TWOD, kpp_3d_fields%fcorr_twod, r8*SIZE(kpp_3d_fields%fcorr_twod))
    rc = kpp_gpu_upload_r8(gpu, KPP_F_FCORR, kpp_3d_fields%fcorr, r8*SIZE(kpp_3d_fields%fcorr))
    rc = kpp_gpu_upload_r8(gpu, KPP_F_RELAX_SAL, kpp_3d_fields%relax_sal, r8*SIZE(kpp_3d_fields%relax_sal))
    rc = kpp_gpu_upload_r8(gpu, KPP_F_RELAX_OCNT, kpp_3d_fields%relax_ocnT, r8*SIZE(kpp_3d_fields%relax_ocnT))
    rc = kpp_gpu_upload_r8(gpu, KPP_F_SAL_CLIM, kpp_3d_fields%sal_clim, r8*SIZE(kpp_3d_fields%sal_clim))
    rc = kpp_gpu_upload_r8(gpu, KPP_F_OCNT_CLIM, kpp_3d_fields%ocnT_clim, r8*SIZE(kpp_3d_fields%ocnT_clim))
    rc = kpp_gpu_upload_r8(gpu, KPP_F_FCORR_WITHZ, kpp_3d_fields%fcorr_withz, r8*SIZE(kpp_3d_fields%fcorr_withz))
    rc = kpp_gpu_upload_r8(gpu, KPP_F_SFCORR_WITHZ, kpp_3d_fields%sfcorr_withz, r8*SIZE(kpp_3d_fields%sfcorr_withz))
    rc = kpp_gpu_upload_r8(gpu, KPP_F_BOTTOM_TEMP, kpp_3d_fields%bottom_temp, r8*SIZE(kpp_3d_fields%bottom_temp))
    rc = kpp_gpu_upload_r8(gpu, KPP_F_ADVECTION, kpp_3d_fields%advection, r8*SIZE(kpp_3d_fields%advection))
    rc = kpp_gpu_upload_r8(gpu, KPP_F_FREEZE_FLAG, kpp_3d_fields%freeze_flag, r8*SIZE(kpp_3d_fields%freeze_flag))
    ! INTEGER members
    rc = kpp_gpu_upload_i4(gpu, KPP_F_OLD, kpp_3d_fields%old, i4*SIZE(kpp_3d_fields%old))
    rc = kpp_gpu_upload_i4(gpu, KPP_F_NEW, kpp_3d_fields%new, i4*SIZE(kpp_3d_fields%new))
    rc = kpp_gpu_upload_i4(gpu, KPP_F_JERLOV, kpp_3d_fields%jerlov, i4*SIZE(kpp_3d_fields%jerlov))
    rc = kpp_gpu_upload_i4(gpu, KPP_F_NMODEADV, kpp_3d_fields%nmodeadv, i4*SIZE(kpp_3d_fields%nmodeadv))
    rc = kpp_gpu_upload_i4(gpu, KPP_F_MODEADV, kpp_3d_fields%modeadv, i4*SIZE(kpp_3d_fields%modeadv))
    ! LOGICAL members are converted (default LOGICAL is 4 bytes but its bit pattern is processor dependent)
    ALLOCATE(mask(npts))
    mask = MERGE(1_c_int32_t, 0_c_int32_t, kpp_3d_fields%l_ocean)
    rc = kpp_gpu_upload_i4(gpu, KPP_F_L_OCEAN, mask, i4*SIZE(mask))
    mask = MERGE(1_c_int32_t, 0_c_int32_t, kpp_3d_fields%run_physics)
    rc = kpp_gpu_upload_i4(gpu, KPP_F_RUN_PHYSICS, mask, i4*SIZE(mask))
    DEALLOCATE(mask)

    IF (.NOT. kpp_const_fields%L_RESTART) THEN
      ! replaces the per-column loop of MCKPP_INITIALIZE_OCEAN_MODEL (initialize_ocean.F90:54-104)
      rc = kpp_gpu_init_vmix(gpu)
      CALL mckpp_physics_gpu_pull_state()
    END IF
  END SUBROUTINE mckpp_physics_gpu_initialize


  ! The reference's entry point, unchanged signature (physics_driver_mod.F90:15).
  SUBROUTINE mckpp_physics_driver()
    TYPE(kpp_step_report) :: rep
    INTEGER(c_int) :: rc
    CHARACTER(LEN=20) :: routine = "MCKPP_PHYSICS_DRIVER"
    CHARACTER(LEN=max_message_len) :: message

    CALL mckpp_start_timer("KPP Physics (GPU)")
    ! what mckpp_fluxes just wrote: sflux(:,1:6,5,0) is one contiguous block of 6*npts REALs
    rc = kpp_gpu_upload_forcing(gpu, kpp_3d_fields%sflux(:,1:6,5,0))
    rc = kpp_gpu_step(gpu, INT(ntime, c_int))
    rc = kpp_gpu_sync(gpu, rep)
    CALL mckpp_stop_timer("KPP Physics (GPU)")

    ! same warnings / fatal errors as the CPU path, from the per-column status word
    IF (rep%n_long_iter .GT. 0) THEN
      WRITE(message,*) 'long iteration at timestep', ntime, ' on ', rep%n_long_iter, ' points'
      CALL mckpp_print_warning("MCKPP_PHYSICS_OCNSTEP", message)
    END IF
    IF (rep%n_reint_fail .GT. 0) THEN
      WRITE(message,*) 'Failed to find a reasonable solution in the semi-implicit integration after ', &
          10, ' iterations on ', rep%n_reint_fail, ' points'
      CALL mckpp_print_warning("MCKPP_PHYSICS_OCNSTEP", message)
    END IF
    IF (rep%n_reset .GT. 0) THEN
      WRITE(message,*) 'Resetting ', rep%n_reset, ' points (climatology / initial currents)'
      CALL mckpp_print_warning("MCKPP_PHSYICS_OVERRIDE_CHECK_PROFILE", message)
    END IF
    IF (rc .EQ. KPP_E_PIVOT_ZERO) THEN
      CALL mckpp_print_error("MCKPP_PHSYICS_SOLVER_TRIDMAT", "Algorithm for solving tridiag matrix failed.")
      CALL mckpp_abort()
    ELSE IF (rc .NE. 0) THEN
      CALL mckpp_print_error(routine, "GPU column step failed")
      CALL mckpp_abort()
    END IF

    ! members the next host stages read every step (coupling / output scalars)
    CALL mckpp_physics_gpu_pull_scalars()
  END SUBROUTINE mckpp_physics_driver


  SUBROUTINE mckpp_physics_gpu_pull_scalars()
    INTEGER(c_int) :: rc
    rc = kpp_gpu_download_r8(gpu, KPP_F_HMIX, kpp_3d_fields%hmix, r8*SIZE(kpp_3d_fields%hmix))
    rc = kpp_gpu_download_r8(gpu, KPP_F_KMIX, kpp_3d_fields%kmix, r8*SIZE(kpp_3d_fields%kmix))
    rc = kpp_gpu_download_r8(gpu, KPP_F_TREF, kpp_3d_fields%Tref, r8*SIZE(kpp_3d_fields%Tref))
    rc = kpp_gpu_download_r8(gpu, KPP_F_UREF, kpp_3d_fields%uref, r8*SIZE(kpp_3d_fields%uref))
    rc = kpp_gpu_download_r8(gpu, KPP_F_VREF, kpp_3d_fields%vref, r8*SIZE(kpp_3d_fields%vref))
    rc = kpp_gpu_download_r8(gpu, KPP_F_SSURF, kpp_3d_fields%Ssurf, r8*SIZE(kpp_3d_fields%Ssurf))
    rc = kpp_gpu_download_r8(gpu, KPP_F_RESET_FLAG, kpp_3d_fields%reset_flag, r8*SIZE(kpp_3d_fields%reset_flag))
    rc = kpp_gpu_download_r8(gpu, KPP_F_FREEZE_FLAG, kpp_3d_fields%freeze_flag, r8*SIZE(kpp_3d_fields%freeze_flag))
    rc = kpp_gpu_download_r8(gpu, KPP_F_DAMPU_FLAG, kpp_3d_fields%dampu_flag, r8*SIZE(kpp_3d_fields%dampu_flag))
    rc = kpp_gpu_download_r8(gpu, KPP_F_DAMPV_FLAG, kpp_3d_fields%dampv_flag, r8*SIZE(kpp_3d_fields%dampv_flag))
    rc = kpp_gpu_download_r8(gpu, KPP_F_FCORR, kpp_3d_fields%fcorr, r8*SIZE(kpp_3d_fields%fcorr))
    rc = kpp_gpu_download_i4(gpu, KPP_F_OLD, kpp_3d_fields%old, i4*SIZE(kpp_3d_fields%old))
    rc = kpp_gpu_download_i4(gpu, KPP_F_NEW, kpp_3d_fields%new, i4*SIZE(kpp_3d_fields%new))
  END SUBROUTINE mckpp_physics_gpu_pull_scalars


  ! prognostic state: before mckpp_restart_control / 3-D output steps
  SUBROUTINE mckpp_physics_gpu_pull_state()
    INTEGER(c_int) :: rc
    CALL mckpp_physics_gpu_pull_scalars()
    rc = kpp_gpu_download_r8(gpu, KPP_F_U, kpp_3d_fields%U, r8*SIZE(kpp_3d_fields%U))
    rc = kpp_gpu_download_r8(gpu, KPP_F_X, kpp_3d_fields%X, r8*SIZE(kpp_3d_fields%X))
    rc = kpp_gpu_download_r8(gpu, KPP_F_US, kpp_3d_fields%Us, r8*SIZE(kpp_3d_fields%Us))
    rc = kpp_gpu_download_r8(gpu, KPP_F_XS, kpp_3d_fields%Xs, r8*SIZE(kpp_3d_fields%Xs))
    rc = kpp_gpu_download_r8(gpu, KPP_F_HMIXD, kpp_3d_fields%hmixd, r8*SIZE(kpp_3d_fields%hmixd))
  END SUBROUTINE mckpp_physics_gpu_pull_state


  ! any other member by id, e.g. before mckpp_output_control sends it to XIOS:
  !   CALL mckpp_physics_gpu_pull(KPP_F_DIFM, kpp_3d_fields%difm)
  SUBROUTINE mckpp_physics_gpu_pull(id, a)
    INTEGER(c_int), INTENT(IN) :: id
    REAL(c_double), INTENT(INOUT), CONTIGUOUS :: a(..)
    INTEGER(c_int) :: rc
    SELECT RANK (a)
    RANK (1); rc = kpp_gpu_download_r8(gpu, id, a, r8*SIZE(a))
    RANK (2); rc = kpp_gpu_download_r8(gpu, id, a, r8*SIZE(a))
    RANK (3); rc = kpp_gpu_download_r8(gpu, id, a, r8*SIZE(a))
    RANK (4); rc = kpp_gpu_download_r8(gpu, id, a, r8*SIZE(a))
    END SELECT
  END SUBROUTINE mckpp_physics_gpu_pull


! One xios_send_field of mckpp_xios_diagnostic_output / mckpp_xios_restart_output
  ! (mckpp_xios_io.F90:72-207, 406-431), packed on the device in the shape that call sends:
  !   CALL mckpp_physics_gpu_pack(KPP_OUT_DIFM, temp_2d);  CALL xios_send_field("difm", temp_2d)
  ! replaces  temp_2d(:,1)=0.0; temp_2d(:,2:NZP1)=kpp_3d_fields%difm(:,1:NZ)  and the pull of difm.
  SUBROUTINE mckpp_physics_gpu_pack(out_id, a)
    INTEGER(c_int), INTENT(IN) :: out_id
    REAL(c_double), INTENT(OUT) :: a(..)
    INTEGER(c_int) :: rc
    rc = kpp_gpu_pack_output(gpu, out_id, a, r8*SIZE(a))
    IF (rc /= 0) THEN
      CALL mckpp_print_error("mckpp_physics_gpu_pack", "kpp_gpu_pack_output failed")
      CALL mckpp_abort()
    END IF
  END SUBROUTINE mckpp_physics_gpu_pack

  ! Device side of MCKPP_BOUNDARY_INTERPOLATE_TEMP / _SAL (mckpp_boundary_interpolate.F90:14-123).
  ! The host still reads prev_ocnT / next_ocnT and computes the weights (:27-52); instead of
  !   kpp_3d_fields%ocnT_clim=next_ocnT*next_weight+prev_ocnT*prev_weight          (:60)
  ! it hands the two records over once per bracket and lets the device blend them:
  !   IF (bracket_moved) CALL mckpp_physics_gpu_clim_records(KPP_F_OCNT_CLIM, prev_ocnT, next_ocnT)
  !   CALL mckpp_physics_gpu_clim_blend(KPP_F_OCNT_CLIM, prev_weight, next_weight)
  SUBROUTINE mckpp_physics_gpu_clim_records(id, prev_rec, next_rec)
    INTEGER(c_int), INTENT(IN) :: id
    REAL(c_double), INTENT(IN) :: prev_rec(:,:), next_rec(:,:)
    INTEGER(c_int) :: rc
    rc = kpp_gpu_upload_clim_record(gpu, id, 0_c_int, prev_rec, r8*SIZE(prev_rec))
    IF (rc == 0) rc = kpp_gpu_upload_clim_record(gpu, id, 1_c_int, next_rec, r8*SIZE(next_rec))
    IF (rc /= 0) THEN
      CALL mckpp_print_error("mckpp_physics_gpu_clim_records", "kpp_gpu_upload_clim_record failed")
      CALL mckpp_abort()
    END IF
  END SUBROUTINE mckpp_physics_gpu_clim_records

  SUBROUTINE mckpp_physics_gpu_clim_blend(id, prev_weight, next_weight)
    INTEGER(c_int), INTENT(IN) :: id
    REAL(c_double), INTENT(IN) :: prev_weight, next_weight
    INTEGER(c_int) :: rc
    rc = kpp_gpu_blend_clim(gpu, id, prev_weight, next_weight)
    IF (rc /= 0) THEN
      CALL mckpp_print_error("mckpp_physics_gpu_clim_blend", "kpp_gpu_blend_clim failed")
      CALL mckpp_abort()
    END IF
  END SUBROUTINE mckpp_physics_gpu_clim_blend

  SUBROUTINE mckpp_physics_gpu_finalize()
    INTEGER(c_int) :: rc
    IF (C_ASSOCIATED(gpu)) rc = kpp_gpu_destroy(gpu)
    gpu = c_null_ptr
  END SUBROUTINE mckpp_physics_gpu_finalize

END MODULE mckpp_physics_driver_mod
