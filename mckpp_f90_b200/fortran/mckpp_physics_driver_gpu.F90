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
! Drop-in: same module name, same public  SUBROUTINE mckpp_physics_driver()  (no arguments; works on
! the module globals kpp_3d_fields / kpp_const_fields).  NOTHING else of the host changes -- not the
! main program, not mckpp_xios_io: the device mirror is created on the first call (the host's own
! mckpp_initialize_fields has filled kpp_3d_fields by then, initial vmix included), and after
! every step the members the host reads next are brought back into kpp_3d_fields:
!   MCKPP_GPU_PULL=all      (default) every member mckpp_fields_1dto3d writes: the host sees exactly
!                           what the CPU path leaves, whatever it does next (costs a PCIe transfer of
!                           the whole set per step)
!   MCKPP_GPU_PULL=active   only the members behind the XIOS fields that are active in iodef.xml
!                           (xios_field_is_active), plus the restart set on restart steps
!   MCKPP_GPU_PULL=scalars  only the per-column scalars; a host that wants 3-D members calls
!                           mckpp_physics_gpu_pull / _pull_state / _pack itself
! Ancillaries that mckpp_boundary_update re-reads on the host are pushed again with the same
! MOD(ntime-1, ndtupd*) conditions (mckpp_boundary_update_mod.F90:42-102).
! MCKPP_GPU_NGPUS=n partitions the columns over n GPUs of the node inside the library
! (kpp_gpu_create_multi); the host stays one rank.
!
! Optional entry points for a host that wants more than the drop-in:
!   mckpp_physics_gpu_initialize(device, ngpus)   explicit creation (e.g. to skip the CPU's initial vmix)
!   mckpp_physics_gpu_pull(id, a) / _pull_state   bring a member back on demand
!   mckpp_physics_gpu_pack(out_id, a)             one xios_send_field block, packed on the device
!   mckpp_physics_gpu_ring_*                      asynchronous output: the copy of step n overlaps step n+1
!   mckpp_physics_gpu_clim_records / _clim_blend  climatology time interpolation on the device
!   mckpp_physics_gpu_finalize()
!
! Build: link libkpp_gpu.so; REAL must be 8 bytes (-fdefault-real-8 / -s real64, as both shipped
! configs already do).
MODULE mckpp_physics_driver_mod

  USE, INTRINSIC :: iso_c_binding
  USE mckpp_data_fields, ONLY: kpp_3d_fields, kpp_const_fields
  USE mckpp_parameters
  USE mckpp_time_control, ONLY: ntime
  USE mckpp_log_messages, ONLY: mckpp_print_warning, mckpp_print_error, max_message_len
  USE mckpp_abort_mod, ONLY: mckpp_abort
  USE mckpp_timer, ONLY: mckpp_start_timer, mckpp_stop_timer
  USE xios, ONLY: xios_field_is_active

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
    INTEGER(c_int) FUNCTION kpp_gpu_create_multi(dims, consts, zm, hm, dm, tri, wmt, wst, ngpus, devices, h) &
        BIND(C, name="kpp_gpu_create_multi")
      IMPORT :: c_int, c_double, c_ptr, kpp_dims, kpp_consts
      TYPE(kpp_dims), INTENT(IN) :: dims
      TYPE(kpp_consts), INTENT(IN) :: consts
      REAL(c_double), INTENT(IN) :: zm(*), hm(*), dm(*), tri(*), wmt(*), wst(*)
      INTEGER(c_int), VALUE :: ngpus
      TYPE(c_ptr), VALUE :: devices           ! c_null_ptr: devices 0..ngpus-1
      TYPE(c_ptr), INTENT(OUT) :: h
    END FUNCTION
    TYPE(c_ptr) FUNCTION kpp_gpu_last_error(h) BIND(C, name="kpp_gpu_last_error")
      IMPORT :: c_ptr
      TYPE(c_ptr), VALUE :: h
    END FUNCTION
    INTEGER(c_size_t) FUNCTION c_strlen(s) BIND(C, name="strlen")
      IMPORT :: c_ptr, c_size_t
      TYPE(c_ptr), VALUE :: s
    END FUNCTION
    INTEGER(c_int) FUNCTION kpp_gpu_output_ring_create(h, out_ids, n_ids, depth) BIND(C, name="kpp_gpu_output_ring_create")
      IMPORT :: c_int, c_ptr, c_int32_t
      TYPE(c_ptr), VALUE :: h
      INTEGER(c_int32_t), INTENT(IN) :: out_ids(*)
      INTEGER(c_int), VALUE :: n_ids, depth
    END FUNCTION
    INTEGER(c_size_t) FUNCTION kpp_gpu_output_ring_offset(h, idx) BIND(C, name="kpp_gpu_output_ring_offset")
      IMPORT :: c_int, c_ptr, c_size_t
      TYPE(c_ptr), VALUE :: h
      INTEGER(c_int), VALUE :: idx
    END FUNCTION
    INTEGER(c_int) FUNCTION kpp_gpu_output_ring_submit(h, slot) BIND(C, name="kpp_gpu_output_ring_submit")
      IMPORT :: c_int, c_ptr
      TYPE(c_ptr), VALUE :: h
      INTEGER(c_int), INTENT(OUT) :: slot
    END FUNCTION
    INTEGER(c_int) FUNCTION kpp_gpu_output_ring_wait(h, slot, host) BIND(C, name="kpp_gpu_output_ring_wait")
      IMPORT :: c_int, c_ptr
      TYPE(c_ptr), VALUE :: h
      INTEGER(c_int), VALUE :: slot
      TYPE(c_ptr), INTENT(OUT) :: host
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
  INTEGER, PARAMETER :: LAST_MEMBER_ID = 58
  ! which members of kpp_3d_fields are brought back after every step (see MCKPP_GPU_PULL above)
  LOGICAL, SAVE :: pull_member(0:LAST_MEMBER_ID) = .FALSE.
  INTEGER, SAVE :: pull_mode = 0            ! 0 all, 1 active, 2 scalars

CONTAINS

  PURE INTEGER(c_int32_t) FUNCTION l2i(l)
    LOGICAL, INTENT(IN) :: l
    l2i = MERGE(1_c_int32_t, 0_c_int32_t, l)
  END FUNCTION l2i

  ! Every return code of the library is checked: on failure print the library's own message and stop
  ! the way the reference does (MCKPP_ABORT), instead of carrying on with a stale or zero device state.
  SUBROUTINE check(rc, what)
    INTEGER(c_int), INTENT(IN) :: rc
    CHARACTER(LEN=*), INTENT(IN) :: what
    TYPE(c_ptr) :: cmsg
    CHARACTER(KIND=c_char), POINTER :: cs(:)
    CHARACTER(LEN=max_message_len) :: message
    INTEGER :: i, n
    IF (rc .EQ. 0) RETURN
    message = ''
    cmsg = kpp_gpu_last_error(gpu)
    IF (C_ASSOCIATED(cmsg)) THEN
      n = MIN(INT(c_strlen(cmsg)), max_message_len)
      CALL C_F_POINTER(cmsg, cs, [n])
      DO i = 1, n
        message(i:i) = cs(i)
      END DO
    END IF
    CALL mckpp_print_error("MCKPP_PHYSICS_GPU", what // " failed: " // TRIM(message))
    CALL mckpp_abort()
  END SUBROUTINE check

  ! Create the device mirror and push everything the physics reads.  Called by the first
  ! mckpp_physics_driver() -- by then mckpp_initialize_fields has filled zm/hm/dm, wmt/wst, tri, the
  ! initial state and the initial vmix results -- or explicitly by a host that wants to choose the
  ! device(s) or let the device do the initial vmix (init_vmix_on_device).
  SUBROUTINE mckpp_physics_gpu_initialize(device, ngpus, init_vmix_on_device)
    INTEGER, INTENT(IN), OPTIONAL :: device, ngpus
    LOGICAL, INTENT(IN), OPTIONAL :: init_vmix_on_device
    TYPE(kpp_dims) :: d
    TYPE(kpp_consts) :: k
    INTEGER(c_int) :: dev, ng
    INTEGER(c_int32_t), ALLOCATABLE :: mask(:)
    CHARACTER(LEN=32) :: env
    INTEGER :: elen, estat

    dev = 0
    IF (PRESENT(device)) dev = device
    ng = 1
    IF (PRESENT(ngpus)) ng = ngpus
    CALL GET_ENVIRONMENT_VARIABLE("MCKPP_GPU_NGPUS", env, elen, estat)
    IF (estat .EQ. 0 .AND. elen .GT. 0) READ(env(1:elen), *) ng
    CALL GET_ENVIRONMENT_VARIABLE("MCKPP_GPU_PULL", env, elen, estat)
    pull_mode = 0
    IF (estat .EQ. 0 .AND. elen .GT. 0) THEN
      IF (env(1:elen) .EQ. 'active') pull_mode = 1
      IF (env(1:elen) .EQ. 'scalars') pull_mode = 2
    END IF

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

    IF (ng .GT. 1) THEN
      CALL check(kpp_gpu_create_multi(d, k, kpp_const_fields%zm, kpp_const_fields%hm, kpp_const_fields%dm, &
                 kpp_const_fields%tri, kpp_const_fields%wmt, kpp_const_fields%wst, ng, c_null_ptr, gpu), &
                 "kpp_gpu_create_multi (no CUDA device? there is no CPU fallback)")
    ELSE
      CALL check(kpp_gpu_create(d, k, kpp_const_fields%zm, kpp_const_fields%hm, kpp_const_fields%dm, &
                 kpp_const_fields%tri, kpp_const_fields%wmt, kpp_const_fields%wst, dev, gpu), &
                 "kpp_gpu_create (no CUDA device? there is no CPU fallback)")
    END IF

    ! REAL members: passed whole, the library moves the part the physics touches
    CALL push_r8(KPP_F_U, kpp_3d_fields%U, SIZE(kpp_3d_fields%U))
    CALL push_r8(KPP_F_X, kpp_3d_fields%X, SIZE(kpp_3d_fields%X))
    CALL push_r8(KPP_F_US, kpp_3d_fields%Us, SIZE(kpp_3d_fields%Us))
    CALL push_r8(KPP_F_XS, kpp_3d_fields%Xs, SIZE(kpp_3d_fields%Xs))
    CALL push_r8(KPP_F_HMIXD, kpp_3d_fields%hmixd, SIZE(kpp_3d_fields%hmixd))
    CALL push_r8(KPP_F_HMIX, kpp_3d_fields%hmix, SIZE(kpp_3d_fields%hmix))
    CALL push_r8(KPP_F_KMIX, kpp_3d_fields%kmix, SIZE(kpp_3d_fields%kmix))
    CALL push_r8(KPP_F_U_INIT, kpp_3d_fields%U_init, SIZE(kpp_3d_fields%U_init))
    CALL push_r8(KPP_F_SSURF, kpp_3d_fields%Ssurf, SIZE(kpp_3d_fields%Ssurf))
    CALL push_r8(KPP_F_SREF, kpp_3d_fields%Sref, SIZE(kpp_3d_fields%Sref))
    CALL push_r8(KPP_F_SSREF, kpp_3d_fields%SSref, SIZE(kpp_3d_fields%SSref))
    CALL push_r8(KPP_F_F, kpp_3d_fields%f, SIZE(kpp_3d_fields%f))
    CALL push_r8(KPP_F_OCDEPTH, kpp_3d_fields%ocdepth, SIZE(kpp_3d_fields%ocdepth))
    CALL push_r8(KPP_F_SFLUX, kpp_3d_fields%sflux, SIZE(kpp_3d_fields%sflux))
    CALL push_r8(KPP_F_RELAX_SST, kpp_3d_fields%relax_sst, SIZE(kpp_3d_fields%relax_sst))
    CALL push_r8(KPP_F_FCORR, kpp_3d_fields%fcorr, SIZE(kpp_3d_fields%fcorr))
    CALL push_r8(KPP_F_RELAX_SAL, kpp_3d_fields%relax_sal, SIZE(kpp_3d_fields%relax_sal))
    CALL push_r8(KPP_F_RELAX_OCNT, kpp_3d_fields%relax_ocnT, SIZE(kpp_3d_fields%relax_ocnT))
    CALL push_r8(KPP_F_ADVECTION, kpp_3d_fields%advection, SIZE(kpp_3d_fields%advection))
    CALL push_r8(KPP_F_FREEZE_FLAG, kpp_3d_fields%freeze_flag, SIZE(kpp_3d_fields%freeze_flag))
    CALL push_r8(KPP_F_TINC_FCORR, kpp_3d_fields%tinc_fcorr, SIZE(kpp_3d_fields%tinc_fcorr))
    CALL push_r8(KPP_F_WXNT, kpp_3d_fields%wXNT, SIZE(kpp_3d_fields%wXNT))
    CALL push_r8(KPP_F_SWFRAC, kpp_3d_fields%swfrac, SIZE(kpp_3d_fields%swfrac))
    CALL push_r8(KPP_F_SWDK_OPT, kpp_3d_fields%swdk_opt, SIZE(kpp_3d_fields%swdk_opt))
    CALL push_ancillaries(.TRUE.)
    ! INTEGER members
    CALL push_i4(KPP_F_OLD, kpp_3d_fields%old, SIZE(kpp_3d_fields%old))
    CALL push_i4(KPP_F_NEW, kpp_3d_fields%new, SIZE(kpp_3d_fields%new))
    CALL push_i4(KPP_F_JERLOV, kpp_3d_fields%jerlov, SIZE(kpp_3d_fields%jerlov))
    CALL push_i4(KPP_F_NMODEADV, kpp_3d_fields%nmodeadv, SIZE(kpp_3d_fields%nmodeadv))
    CALL push_i4(KPP_F_MODEADV, kpp_3d_fields%modeadv, SIZE(kpp_3d_fields%modeadv))
    ! LOGICAL members are converted (default LOGICAL is 4 bytes but its bit pattern is processor dependent)
    ALLOCATE(mask(npts))
    mask = MERGE(1_c_int32_t, 0_c_int32_t, kpp_3d_fields%l_ocean)
    CALL push_i4(KPP_F_L_OCEAN, mask, SIZE(mask))
    mask = MERGE(1_c_int32_t, 0_c_int32_t, kpp_3d_fields%run_physics)
    CALL push_i4(KPP_F_RUN_PHYSICS, mask, SIZE(mask))
    DEALLOCATE(mask)

    IF (PRESENT(init_vmix_on_device)) THEN
      IF (init_vmix_on_device .AND. .NOT. kpp_const_fields%L_RESTART) THEN
        ! replaces the per-column loop of MCKPP_INITIALIZE_OCEAN_MODEL (initialize_ocean.F90:54-104)
        CALL check(kpp_gpu_init_vmix(gpu), "kpp_gpu_init_vmix")
        CALL mckpp_physics_gpu_pull_state()
      END IF
    END IF
    CALL select_pulls()
  END SUBROUTINE mckpp_physics_gpu_initialize


  SUBROUTINE push_r8(id, a, n)
    INTEGER(c_int), INTENT(IN) :: id
    REAL(c_double), INTENT(IN) :: a(*)
    INTEGER, INTENT(IN) :: n
    CALL check(kpp_gpu_upload_r8(gpu, id, a, r8*INT(n, c_size_t)), "kpp_gpu_upload_field")
  END SUBROUTINE push_r8

  SUBROUTINE push_i4(id, a, n)
    INTEGER(c_int), INTENT(IN) :: id
    INTEGER(c_int32_t), INTENT(IN) :: a(*)
    INTEGER, INTENT(IN) :: n
    CALL check(kpp_gpu_upload_i4(gpu, id, a, i4*INT(n, c_size_t)), "kpp_gpu_upload_field")
  END SUBROUTINE push_i4

  SUBROUTINE pull_r8(id, a, n)
    INTEGER(c_int), INTENT(IN) :: id
    REAL(c_double), INTENT(INOUT) :: a(*)
    INTEGER, INTENT(IN) :: n
    CALL check(kpp_gpu_download_r8(gpu, id, a, r8*INT(n, c_size_t)), "kpp_gpu_download_field")
  END SUBROUTINE pull_r8

  SUBROUTINE pull_i4(id, a, n)
    INTEGER(c_int), INTENT(IN) :: id
    INTEGER(c_int32_t), INTENT(INOUT) :: a(*)
    INTEGER, INTENT(IN) :: n
    CALL check(kpp_gpu_download_i4(gpu, id, a, i4*INT(n, c_size_t)), "kpp_gpu_download_field")
  END SUBROUTINE pull_i4


  ! Ancillaries the host re-reads from file in mckpp_boundary_update: pushed at creation (everything)
  ! and afterwards under the very conditions of mckpp_boundary_update_mod.F90:42-102, evaluated with
  ! the same ntime (the main program calls mckpp_boundary_update before the physics, for nt /= 1).
  SUBROUTINE push_ancillaries(everything)
    LOGICAL, INTENT(IN) :: everything
    LOGICAL :: upd
    IF (.NOT. everything .AND. ntime .EQ. 1) RETURN
    upd = everything
    IF (.NOT. upd) upd = kpp_const_fields%L_UPD_CLIMSST .AND. MOD(ntime-1, kpp_const_fields%ndtupdsst) .EQ. 0
    IF (upd) CALL push_r8(KPP_F_SST0, kpp_3d_fields%SST0, SIZE(kpp_3d_fields%SST0))
    upd = everything
    IF (.NOT. upd) upd = kpp_const_fields%L_UPD_FCORR .AND. MOD(ntime-1, kpp_const_fields%ndtupdfcorr) .EQ. 0
    IF (upd) THEN
      CALL push_r8(KPP_F_FCORR_WITHZ, kpp_3d_fields%fcorr_withz, SIZE(kpp_3d_fields%fcorr_withz))
      CALL push_r8(KPP_F_FCORR_TWOD, kpp_3d_fields%fcorr_twod, SIZE(kpp_3d_fields%fcorr_twod))
    END IF
    upd = everything
    IF (.NOT. upd) upd = kpp_const_fields%L_UPD_SFCORR .AND. MOD(ntime-1, kpp_const_fields%ndtupdsfcorr) .EQ. 0
    IF (upd) CALL push_r8(KPP_F_SFCORR_WITHZ, kpp_3d_fields%sfcorr_withz, SIZE(kpp_3d_fields%sfcorr_withz))
    upd = everything
    IF (.NOT. upd) upd = kpp_const_fields%L_UPD_BOTTOM_TEMP .AND. MOD(ntime-1, kpp_const_fields%ndtupdbottom) .EQ. 0
    IF (upd) CALL push_r8(KPP_F_BOTTOM_TEMP, kpp_3d_fields%bottom_temp, SIZE(kpp_3d_fields%bottom_temp))
    upd = everything
    IF (.NOT. upd .AND. kpp_const_fields%L_UPD_SAL) THEN
      IF (kpp_const_fields%L_INTERP_SAL) THEN
        upd = MOD(ntime-1, kpp_const_fields%ndt_interp_sal) .EQ. 0
      ELSE
        upd = MOD(ntime-1, kpp_const_fields%ndtupdsal) .EQ. 0
      END IF
    END IF
    IF (upd) CALL push_r8(KPP_F_SAL_CLIM, kpp_3d_fields%sal_clim, SIZE(kpp_3d_fields%sal_clim))
    upd = everything
    IF (.NOT. upd .AND. kpp_const_fields%L_UPD_OCNT) THEN
      IF (kpp_const_fields%L_INTERP_OCNT) THEN
        upd = MOD(ntime-1, kpp_const_fields%ndt_interp_ocnt) .EQ. 0
      ELSE
        upd = MOD(ntime-1, kpp_const_fields%ndtupdocnt) .EQ. 0
      END IF
    END IF
    IF (upd) CALL push_r8(KPP_F_OCNT_CLIM, kpp_3d_fields%ocnT_clim, SIZE(kpp_3d_fields%ocnT_clim))
  END SUBROUTINE push_ancillaries


  ! Which members come back after every step.
  SUBROUTINE select_pulls()
    pull_member = .FALSE.
    IF (pull_mode .EQ. 0) THEN
      ! everything mckpp_fields_1dto3d writes (types_transfer.F90:199-327)
      pull_member((/ KPP_F_U, KPP_F_X, KPP_F_US, KPP_F_XS, KPP_F_HMIXD, KPP_F_RHO, KPP_F_CP, KPP_F_BUOY, &
                     KPP_F_RIG, KPP_F_DBLOC, KPP_F_SHSQ, KPP_F_DIFM, KPP_F_DIFS, KPP_F_DIFT, KPP_F_GHAT, &
                     KPP_F_WU, KPP_F_WX, KPP_F_WXNT, KPP_F_TINC_FCORR, KPP_F_SINC_FCORR, KPP_F_OCNTCORR, &
                     KPP_F_SCORR, KPP_F_SWFRAC, KPP_F_SWDK_OPT /)) = .TRUE.
    ELSE IF (pull_mode .EQ. 1) THEN
      ! the members behind the fields XIOS will actually write (mckpp_xios_io.F90:72-207)
      IF (xios_field_is_active("u") .OR. xios_field_is_active("v")) pull_member(KPP_F_U) = .TRUE.
      IF (xios_field_is_active("T") .OR. xios_field_is_active("S")) pull_member(KPP_F_X) = .TRUE.
      IF (xios_field_is_active("B")) pull_member(KPP_F_BUOY) = .TRUE.
      IF (xios_field_is_active("wu") .OR. xios_field_is_active("wv")) pull_member(KPP_F_WU) = .TRUE.
      IF (xios_field_is_active("wT") .OR. xios_field_is_active("wS") .OR. xios_field_is_active("wB")) &
        pull_member(KPP_F_WX) = .TRUE.
      IF (xios_field_is_active("wTnt")) pull_member(KPP_F_WXNT) = .TRUE.
      IF (xios_field_is_active("difm")) pull_member(KPP_F_DIFM) = .TRUE.
      IF (xios_field_is_active("dift")) pull_member(KPP_F_DIFT) = .TRUE.
      IF (xios_field_is_active("difs")) pull_member(KPP_F_DIFS) = .TRUE.
      IF (xios_field_is_active("rho")) pull_member(KPP_F_RHO) = .TRUE.
      IF (xios_field_is_active("cp")) pull_member(KPP_F_CP) = .TRUE.
      IF (xios_field_is_active("scorr")) pull_member(KPP_F_SCORR) = .TRUE.
      IF (xios_field_is_active("Rig")) pull_member(KPP_F_RIG) = .TRUE.
      IF (xios_field_is_active("dbloc")) pull_member(KPP_F_DBLOC) = .TRUE.
      IF (xios_field_is_active("Shsq")) pull_member(KPP_F_SHSQ) = .TRUE.
      IF (xios_field_is_active("tinc_fcorr")) pull_member(KPP_F_TINC_FCORR) = .TRUE.
      IF (xios_field_is_active("fcorr_z")) pull_member(KPP_F_OCNTCORR) = .TRUE.
      IF (xios_field_is_active("sinc_fcorr")) pull_member(KPP_F_SINC_FCORR) = .TRUE.
    END IF
  END SUBROUTINE select_pulls


  ! The reference's entry point, unchanged signature (physics_driver_mod.F90:15).
  SUBROUTINE mckpp_physics_driver()
    TYPE(kpp_step_report) :: rep
    INTEGER(c_int) :: rc
    INTEGER :: id
    LOGICAL :: restart_step
    CHARACTER(LEN=max_message_len) :: message

    IF (.NOT. C_ASSOCIATED(gpu)) CALL mckpp_physics_gpu_initialize()

    CALL mckpp_start_timer("KPP Physics (GPU)")
    CALL push_ancillaries(.FALSE.)
    ! what mckpp_fluxes just wrote: sflux(:,1:6,5,0) is one contiguous block of 6*npts REALs
    CALL check(kpp_gpu_upload_forcing(gpu, kpp_3d_fields%sflux(:,1:6,5,0)), "kpp_gpu_upload_forcing")
    CALL check(kpp_gpu_step(gpu, INT(ntime, c_int)), "kpp_gpu_step")
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
    END IF
    CALL check(rc, "kpp_gpu_sync")

    ! members the next host stages read: the per-column scalars always, 3-D members per MCKPP_GPU_PULL,
    ! the restart set when mckpp_restart_control is about to write (mckpp_xios_control.F90:69-71)
    CALL mckpp_start_timer("KPP GPU device-to-host")
    CALL mckpp_physics_gpu_pull_scalars()
    restart_step = MOD(ntime, kpp_const_fields%ndt_per_restart) .EQ. 0 .OR. ntime .EQ. kpp_const_fields%num_timesteps
    DO id = 0, LAST_MEMBER_ID
      IF (pull_member(id)) CALL pull_by_id(INT(id, c_int))
    END DO
    IF (restart_step .AND. pull_mode .NE. 0) THEN
      CALL mckpp_physics_gpu_pull_state()
      IF (.NOT. pull_member(KPP_F_CP)) CALL pull_by_id(KPP_F_CP)
      IF (.NOT. pull_member(KPP_F_RHO)) CALL pull_by_id(KPP_F_RHO)
    END IF
    CALL mckpp_stop_timer("KPP GPU device-to-host")
  END SUBROUTINE mckpp_physics_driver


  SUBROUTINE pull_by_id(id)
    INTEGER(c_int), INTENT(IN) :: id
    SELECT CASE (id)
    CASE (KPP_F_U); CALL pull_r8(id, kpp_3d_fields%U, SIZE(kpp_3d_fields%U))
    CASE (KPP_F_X); CALL pull_r8(id, kpp_3d_fields%X, SIZE(kpp_3d_fields%X))
    CASE (KPP_F_US); CALL pull_r8(id, kpp_3d_fields%Us, SIZE(kpp_3d_fields%Us))
    CASE (KPP_F_XS); CALL pull_r8(id, kpp_3d_fields%Xs, SIZE(kpp_3d_fields%Xs))
    CASE (KPP_F_HMIXD); CALL pull_r8(id, kpp_3d_fields%hmixd, SIZE(kpp_3d_fields%hmixd))
    CASE (KPP_F_RHO); CALL pull_r8(id, kpp_3d_fields%rho, SIZE(kpp_3d_fields%rho))
    CASE (KPP_F_CP); CALL pull_r8(id, kpp_3d_fields%cp, SIZE(kpp_3d_fields%cp))
    CASE (KPP_F_BUOY); CALL pull_r8(id, kpp_3d_fields%buoy, SIZE(kpp_3d_fields%buoy))
    CASE (KPP_F_RIG); CALL pull_r8(id, kpp_3d_fields%Rig, SIZE(kpp_3d_fields%Rig))
    CASE (KPP_F_DBLOC); CALL pull_r8(id, kpp_3d_fields%dbloc, SIZE(kpp_3d_fields%dbloc))
    CASE (KPP_F_SHSQ); CALL pull_r8(id, kpp_3d_fields%Shsq, SIZE(kpp_3d_fields%Shsq))
    CASE (KPP_F_DIFM); CALL pull_r8(id, kpp_3d_fields%difm, SIZE(kpp_3d_fields%difm))
    CASE (KPP_F_DIFS); CALL pull_r8(id, kpp_3d_fields%difs, SIZE(kpp_3d_fields%difs))
    CASE (KPP_F_DIFT); CALL pull_r8(id, kpp_3d_fields%dift, SIZE(kpp_3d_fields%dift))
    CASE (KPP_F_GHAT); CALL pull_r8(id, kpp_3d_fields%ghat, SIZE(kpp_3d_fields%ghat))
    CASE (KPP_F_WU); CALL pull_r8(id, kpp_3d_fields%wU, SIZE(kpp_3d_fields%wU))
    CASE (KPP_F_WX); CALL pull_r8(id, kpp_3d_fields%wX, SIZE(kpp_3d_fields%wX))
    CASE (KPP_F_WXNT); CALL pull_r8(id, kpp_3d_fields%wXNT, SIZE(kpp_3d_fields%wXNT))
    CASE (KPP_F_TINC_FCORR); CALL pull_r8(id, kpp_3d_fields%tinc_fcorr, SIZE(kpp_3d_fields%tinc_fcorr))
    CASE (KPP_F_SINC_FCORR); CALL pull_r8(id, kpp_3d_fields%sinc_fcorr, SIZE(kpp_3d_fields%sinc_fcorr))
    CASE (KPP_F_OCNTCORR); CALL pull_r8(id, kpp_3d_fields%ocnTcorr, SIZE(kpp_3d_fields%ocnTcorr))
    CASE (KPP_F_SCORR); CALL pull_r8(id, kpp_3d_fields%scorr, SIZE(kpp_3d_fields%scorr))
    CASE (KPP_F_SWFRAC); CALL pull_r8(id, kpp_3d_fields%swfrac, SIZE(kpp_3d_fields%swfrac))
    CASE (KPP_F_SWDK_OPT); CALL pull_r8(id, kpp_3d_fields%swdk_opt, SIZE(kpp_3d_fields%swdk_opt))
    END SELECT
  END SUBROUTINE pull_by_id


  SUBROUTINE mckpp_physics_gpu_pull_scalars()
    CALL pull_r8(KPP_F_HMIX, kpp_3d_fields%hmix, SIZE(kpp_3d_fields%hmix))
    CALL pull_r8(KPP_F_KMIX, kpp_3d_fields%kmix, SIZE(kpp_3d_fields%kmix))
    CALL pull_r8(KPP_F_TREF, kpp_3d_fields%Tref, SIZE(kpp_3d_fields%Tref))
    CALL pull_r8(KPP_F_UREF, kpp_3d_fields%uref, SIZE(kpp_3d_fields%uref))
    CALL pull_r8(KPP_F_VREF, kpp_3d_fields%vref, SIZE(kpp_3d_fields%vref))
    CALL pull_r8(KPP_F_SSURF, kpp_3d_fields%Ssurf, SIZE(kpp_3d_fields%Ssurf))
    CALL pull_r8(KPP_F_RESET_FLAG, kpp_3d_fields%reset_flag, SIZE(kpp_3d_fields%reset_flag))
    CALL pull_r8(KPP_F_FREEZE_FLAG, kpp_3d_fields%freeze_flag, SIZE(kpp_3d_fields%freeze_flag))
    CALL pull_r8(KPP_F_DAMPU_FLAG, kpp_3d_fields%dampu_flag, SIZE(kpp_3d_fields%dampu_flag))
    CALL pull_r8(KPP_F_DAMPV_FLAG, kpp_3d_fields%dampv_flag, SIZE(kpp_3d_fields%dampv_flag))
    CALL pull_r8(KPP_F_FCORR, kpp_3d_fields%fcorr, SIZE(kpp_3d_fields%fcorr))
    CALL pull_i4(KPP_F_OLD, kpp_3d_fields%old, SIZE(kpp_3d_fields%old))
    CALL pull_i4(KPP_F_NEW, kpp_3d_fields%new, SIZE(kpp_3d_fields%new))
  END SUBROUTINE mckpp_physics_gpu_pull_scalars


  ! prognostic state: before mckpp_restart_control / 3-D output steps
  SUBROUTINE mckpp_physics_gpu_pull_state()
    CALL mckpp_physics_gpu_pull_scalars()
    CALL pull_by_id(KPP_F_U)
    CALL pull_by_id(KPP_F_X)
    CALL pull_by_id(KPP_F_US)
    CALL pull_by_id(KPP_F_XS)
    CALL pull_by_id(KPP_F_HMIXD)
  END SUBROUTINE mckpp_physics_gpu_pull_state


  ! any other member by id, e.g. before mckpp_output_control sends it to XIOS:
  !   CALL mckpp_physics_gpu_pull(KPP_F_DIFM, kpp_3d_fields%difm)
  SUBROUTINE mckpp_physics_gpu_pull(id, a)
    INTEGER(c_int), INTENT(IN) :: id
    REAL(c_double), INTENT(INOUT), CONTIGUOUS :: a(..)
    SELECT RANK (a)
    RANK (1); CALL pull_r8(id, a, SIZE(a))
    RANK (2); CALL pull_r8(id, a, SIZE(a))
    RANK (3); CALL pull_r8(id, a, SIZE(a))
    RANK (4); CALL pull_r8(id, a, SIZE(a))
    END SELECT
  END SUBROUTINE mckpp_physics_gpu_pull


  ! One xios_send_field of mckpp_xios_diagnostic_output / mckpp_xios_restart_output
  ! (mckpp_xios_io.F90:72-207, 406-431), packed on the device in the shape that call sends:
  !   CALL mckpp_physics_gpu_pack(KPP_OUT_DIFM, temp_2d);  CALL xios_send_field("difm", temp_2d)
  ! replaces  temp_2d(:,1)=0.0; temp_2d(:,2:NZP1)=kpp_3d_fields%difm(:,1:NZ)  and the pull of difm.
  ! (An assumed-rank dummy may only be passed on inside SELECT RANK -- F2018 C839.)
  SUBROUTINE mckpp_physics_gpu_pack(out_id, a)
    INTEGER(c_int), INTENT(IN) :: out_id
    REAL(c_double), INTENT(INOUT), CONTIGUOUS :: a(..)
    SELECT RANK (a)
    RANK (1); CALL pack_block(out_id, a, SIZE(a))
    RANK (2); CALL pack_block(out_id, a, SIZE(a))
    END SELECT
  END SUBROUTINE mckpp_physics_gpu_pack

  SUBROUTINE pack_block(out_id, a, n)
    INTEGER(c_int), INTENT(IN) :: out_id
    REAL(c_double), INTENT(INOUT) :: a(*)
    INTEGER, INTENT(IN) :: n
    CALL check(kpp_gpu_pack_output(gpu, out_id, a, r8*INT(n, c_size_t)), "kpp_gpu_pack_output")
  END SUBROUTINE pack_block


  ! Asynchronous output (optional, needs a host that reads its output blocks from the ring instead
  ! of kpp_3d_fields): the device->host copy of step n overlaps the kernels of step n+1.
  !   CALL mckpp_physics_gpu_ring_create((/ KPP_OUT_T, KPP_OUT_S, KPP_OUT_HMIX /), 2)      once
  !   after mckpp_physics_driver():   CALL mckpp_physics_gpu_ring_submit(slot)
  !   one step later:                 CALL mckpp_physics_gpu_ring_block(slot_prev, 1, temp_2d_ptr)   ! "T"
  !                                   CALL xios_send_field("T", temp_2d_ptr)
  SUBROUTINE mckpp_physics_gpu_ring_create(out_ids, depth)
    INTEGER(c_int32_t), INTENT(IN) :: out_ids(:)
    INTEGER, INTENT(IN) :: depth
    IF (.NOT. C_ASSOCIATED(gpu)) CALL mckpp_physics_gpu_initialize()
    CALL check(kpp_gpu_output_ring_create(gpu, out_ids, INT(SIZE(out_ids), c_int), INT(depth, c_int)), &
               "kpp_gpu_output_ring_create")
  END SUBROUTINE mckpp_physics_gpu_ring_create

  SUBROUTINE mckpp_physics_gpu_ring_submit(slot)
    INTEGER(c_int), INTENT(OUT) :: slot
    CALL check(kpp_gpu_output_ring_submit(gpu, slot), "kpp_gpu_output_ring_submit")
  END SUBROUTINE mckpp_physics_gpu_ring_submit

  ! block `idx` (1-based position in out_ids) of a finished slot as a (npts, rows) array
  SUBROUTINE mckpp_physics_gpu_ring_block(slot, idx, rows, block)
    INTEGER(c_int), INTENT(IN) :: slot
    INTEGER, INTENT(IN) :: idx, rows
    REAL(c_double), POINTER, INTENT(OUT) :: block(:,:)
    TYPE(c_ptr) :: base
    INTEGER(c_size_t) :: off
    REAL(c_double), POINTER :: flat(:)
    CALL check(kpp_gpu_output_ring_wait(gpu, slot, base), "kpp_gpu_output_ring_wait")
    off = kpp_gpu_output_ring_offset(gpu, INT(idx-1, c_int)) / r8
    CALL C_F_POINTER(base, flat, [off + INT(npts, c_size_t)*INT(rows, c_size_t)])
    block(1:npts, 1:rows) => flat(off+1 : off + INT(npts, c_size_t)*INT(rows, c_size_t))
  END SUBROUTINE mckpp_physics_gpu_ring_block

  ! Device side of MCKPP_BOUNDARY_INTERPOLATE_TEMP / _SAL (mckpp_boundary_interpolate.F90:14-123).
  ! The host still reads prev_ocnT / next_ocnT and computes the weights (:27-52); instead of
  !   kpp_3d_fields%ocnT_clim=next_ocnT*next_weight+prev_ocnT*prev_weight          (:60)
  ! it hands the two records over once per bracket and lets the device blend them:
  !   IF (bracket_moved) CALL mckpp_physics_gpu_clim_records(KPP_F_OCNT_CLIM, prev_ocnT, next_ocnT)
  !   CALL mckpp_physics_gpu_clim_blend(KPP_F_OCNT_CLIM, prev_weight, next_weight)
  SUBROUTINE mckpp_physics_gpu_clim_records(id, prev_rec, next_rec)
    INTEGER(c_int), INTENT(IN) :: id
    REAL(c_double), INTENT(IN), CONTIGUOUS :: prev_rec(:,:), next_rec(:,:)
    CALL check(kpp_gpu_upload_clim_record(gpu, id, 0_c_int, prev_rec, r8*SIZE(prev_rec)), "kpp_gpu_upload_clim_record")
    CALL check(kpp_gpu_upload_clim_record(gpu, id, 1_c_int, next_rec, r8*SIZE(next_rec)), "kpp_gpu_upload_clim_record")
  END SUBROUTINE mckpp_physics_gpu_clim_records

  SUBROUTINE mckpp_physics_gpu_clim_blend(id, prev_weight, next_weight)
    INTEGER(c_int), INTENT(IN) :: id
    REAL(c_double), INTENT(IN) :: prev_weight, next_weight
    CALL check(kpp_gpu_blend_clim(gpu, id, prev_weight, next_weight), "kpp_gpu_blend_clim")
  END SUBROUTINE mckpp_physics_gpu_clim_blend

  SUBROUTINE mckpp_physics_gpu_finalize()
    INTEGER(c_int) :: rc
    IF (C_ASSOCIATED(gpu)) rc = kpp_gpu_destroy(gpu)
    gpu = c_null_ptr
  END SUBROUTINE mckpp_physics_gpu_finalize

END MODULE mckpp_physics_driver_mod
