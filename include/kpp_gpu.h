/*
 * kpp_gpu.h -- C ABI of the B200 (sm_100a) MC-KPP column-physics library.
 *
 * This is the drop-in boundary for ONE path of aosprey/mckpp-f90: the
 * per-timestep column physics
 *
 *     CALL mckpp_physics_driver()            src/mckpp_ocean_model_3D.F90:58
 *       -> mckpp_fields_3dto1d               src/mckpp_types_transfer.F90:15
 *       -> mckpp_physics_ocnstep             src/mckpp_physics_ocnstep_mod.F90:43
 *       -> mckpp_physics_overrides_check_profile   src/mckpp_physics_overrides.F90:42
 *       -> mckpp_fields_1dto3d               src/mckpp_types_transfer.F90:199
 *       -> mckpp_physics_overrides_bottomtemp      src/mckpp_physics_overrides.F90:12
 *
 * plus the initial per-column vmix of MCKPP_INITIALIZE_OCEAN_MODEL
 * (src/mckpp_initialize_ocean.F90:54-104).
 *
 * Everything crossing the boundary is plain C: pointers, sizes, int32 and
 * double.  Host arrays have exactly the shape and column-major element order
 * of the corresponding member of the reference's `kpp_3d_fields`
 * (src/mckpp_data_fields.F90:353-447) with REAL = 8 bytes (-fdefault-real-8),
 * INTEGER and LOGICAL = 4 bytes, so a Fortran host passes `kpp_3d_fields%X`
 * etc. by address through ISO_C_BINDING.  Columns are the
 * fastest index on both sides; the device keeps the same structure-of-arrays
 * with a padded leading dimension.
 *
 * All functions return 0 on success or a negative KPP_E_* code; none throws.
 * There is no CPU fallback: without a CUDA device every call fails.
 */
#ifndef KPP_GPU_H
#define KPP_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KPP_GPU_ABI_VERSION 1

/* error codes */
#define KPP_OK              0
#define KPP_E_INVALID      -1   /* bad argument / size mismatch / unsupported switch */
#define KPP_E_CUDA         -2   /* CUDA runtime error (see kpp_gpu_last_error) */
#define KPP_E_NODEVICE     -3   /* no CUDA device: there is no CPU fallback */
#define KPP_E_PIVOT_ZERO   -4   /* tridmat zero pivot: the reference calls MCKPP_ABORT (solvers.F90:140-149) */
#define KPP_E_NOMEM        -5

/* mckpp_parameters subset (src/mckpp_parameters.F90:4-59) */
typedef struct kpp_dims {
    int32_t npts;        /* columns owned by this handle (nx*ny, or this rank's block) */
    int32_t nz;          /* layers; nzp1 = nz+1 grid points */
    int32_t nztmax;      /* >= nz+1; extent of difm/difs/dift/wU/wX/wXNT/ghat host arrays */
    int32_t nsflxs;      /* 9  (initialize_namelist_mod.F90:37) */
    int32_t njdt;        /* 1  (:38) */
    int32_t maxmodeadv;  /* 6  (:40) */
} kpp_dims;

/* kpp_const_type scalars and switches read by the path
 * (src/mckpp_data_fields.F90:187-346; defaults initialize_namelist_mod.F90:27-47,92-119) */
typedef struct kpp_consts {
    double dto;          /* ocean timestep (s) */
    double grav, vonk, sice;
    double hmixtolfrac;  /* mckpp_parameters */
    double iso_thresh;
    int32_t itermax;     /* mckpp_parameters */
    int32_t iso_bot;
    int32_t dt_uvdamp;
    int32_t LKPP;        /* must be 1: with LKPP=.FALSE. the reference leaves hmix/kmix undefined */
    int32_t LRI, LDD, L_SSref;
    int32_t L_RELAX_SST, L_RELAX_CALCONLY, L_FCORR, L_FCORR_WITHZ;
    int32_t L_SFCORR, L_SFCORR_WITHZ, L_RELAX_SAL, L_RELAX_OCNT;
    int32_t L_NO_FREEZE, L_NO_ISOTHERM, L_DAMP_CURR, L_VARY_BOTTOM_TEMP;
    int32_t have_ocnT_file;  /* ocnT_file .ne. 'none' (overrides.F90:57) */
    int32_t have_sal_file;   /* sal_file  .ne. 'none' */
    /* numerics variant of the kernels:
     *   0 = strict: no FMA contraction, every divide a true IEEE divide, same
     *       operation order as the reference's x86-64 gfortran build; differs
     *       from it only through exp() (CUDA libdevice vs glibc).
     *   1 = fast: FMA contraction on and shared reciprocals; tolerance-only parity. */
    int32_t numerics;
    int32_t reserved;
} kpp_consts;

/* Fields of kpp_3d_fields that can be uploaded / downloaded.  The host buffer
 * is always the WHOLE Fortran array of that member (shape in the comment,
 * column-major); the library moves only the part the physics touches. */
typedef enum kpp_field_id {
    KPP_F_U = 0,        /* U(npts,nzp1,2)            in/out */
    KPP_F_X,            /* X(npts,nzp1,2)            in/out */
    KPP_F_US,           /* Us(npts,nzp1,2,0:1)       in/out */
    KPP_F_XS,           /* Xs(npts,nzp1,2,0:1)       in/out */
    KPP_F_HMIXD,        /* hmixd(npts,0:1)           in/out */
    KPP_F_OLD,          /* old(npts)        INTEGER  in/out */
    KPP_F_NEW,          /* new(npts)        INTEGER  in/out */
    KPP_F_HMIX,         /* hmix(npts)                out (in for restart) */
    KPP_F_KMIX,         /* kmix(npts)       REAL     out */
    KPP_F_TREF,         /* Tref(npts)                out */
    KPP_F_UREF,         /* uref(npts)                out */
    KPP_F_VREF,         /* vref(npts)                out */
    KPP_F_SSURF,        /* Ssurf(npts)               in/out */
    KPP_F_SREF,         /* Sref(npts)                in */
    KPP_F_SSREF,        /* SSref(npts)               in */
    KPP_F_F,            /* f(npts)                   in */
    KPP_F_OCDEPTH,      /* ocdepth(npts)             in */
    KPP_F_JERLOV,       /* jerlov(npts)     INTEGER  in; 1..5, anything else is rejected */
    KPP_F_L_OCEAN,      /* l_ocean(npts)    LOGICAL  in */
    KPP_F_RUN_PHYSICS,  /* run_physics(npts) LOGICAL in */
    KPP_F_SFLUX,        /* sflux(npts,nsflxs,5,0:njdt)  in; only (:,1:6,5,0) is moved */
    KPP_F_U_INIT,       /* U_init(npts,nzp1,2)       in */
    KPP_F_RELAX_SST,    /* relax_sst(npts)           in */
    KPP_F_SST0,         /* SST0(npts)                in */
    KPP_F_FCORR_TWOD,   /* fcorr_twod(npts)          in */
    KPP_F_FCORR,        /* fcorr(npts)               in/out */
    KPP_F_RELAX_SAL,    /* relax_sal(npts)           in */
    KPP_F_RELAX_OCNT,   /* relax_ocnT(npts)          in */
    KPP_F_SAL_CLIM,     /* sal_clim(npts,nzp1)       in */
    KPP_F_OCNT_CLIM,    /* ocnT_clim(npts,nzp1)      in */
    KPP_F_FCORR_WITHZ,  /* fcorr_withz(npts,nzp1)    in */
    KPP_F_SFCORR_WITHZ, /* sfcorr_withz(npts,nzp1)   in */
    KPP_F_BOTTOM_TEMP,  /* bottom_temp(npts)         in */
    KPP_F_NMODEADV,     /* nmodeadv(npts,2) INTEGER  in; only (:,2) is moved */
    KPP_F_MODEADV,      /* modeadv(npts,maxmodeadv,2) INTEGER in; only (:,:,2); a mode > 7 among the first
                         * nmodeadv(:,2) entries of a column makes the next kpp_gpu_step fail (solvers.F90:320) */
    KPP_F_ADVECTION,    /* advection(npts,maxmodeadv,2) in; only (:,:,2) */
    KPP_F_FREEZE_FLAG,  /* freeze_flag(npts)         in/out */
    KPP_F_RESET_FLAG,   /* reset_flag(npts)          out */
    KPP_F_DAMPU_FLAG,   /* dampu_flag(npts)          out */
    KPP_F_DAMPV_FLAG,   /* dampv_flag(npts)          out */
    /* diagnostics written back by 1dto3d */
    KPP_F_RHO,          /* rho(npts,0:nzp1tmax)      out; rows 0:nzp1 */
    KPP_F_CP,           /* cp(npts,0:nzp1tmax)       out; rows 0:nzp1 */
    KPP_F_BUOY,         /* buoy(npts,nzp1tmax)       out; rows 1:nzp1 */
    KPP_F_RIG,          /* Rig(npts,nzp1)            out; rows 1:nz */
    KPP_F_DBLOC,        /* dbloc(npts,nz)            out */
    KPP_F_SHSQ,         /* Shsq(npts,nzp1)           out; rows 1:nz */
    KPP_F_DIFM,         /* difm(npts,0:nztmax)       out; rows 0:nzp1 */
    KPP_F_DIFS,         /* difs(npts,0:nztmax)       out; rows 0:nzp1 */
    KPP_F_DIFT,         /* dift(npts,0:nztmax)       out; rows 0:nzp1 */
    KPP_F_GHAT,         /* ghat(npts,nztmax)         out; rows 1:nz */
    KPP_F_WU,           /* wU(npts,0:nztmax,3)       out; (:,0:nz,1:2) */
    KPP_F_WX,           /* wX(npts,0:nztmax,3)       out; (:,0:nz,1:3) */
    KPP_F_WXNT,         /* wXNT(npts,0:nztmax,2)     out; (:,0:nz,1) */
    KPP_F_TINC_FCORR,   /* tinc_fcorr(npts,nzp1)     out */
    KPP_F_SINC_FCORR,   /* sinc_fcorr(npts,nzp1)     out */
    KPP_F_OCNTCORR,     /* ocnTcorr(npts,nzp1)       out */
    KPP_F_SCORR,        /* scorr(npts,nzp1)          out */
    KPP_F_SWFRAC,       /* swfrac(npts,nzp1)         out (filled at ntime<=1, bldepth_mod.F90:113) */
    KPP_F_SWDK_OPT,     /* swdk_opt(npts,0:nz)       out (filled at ntime<=1, fluxes_mod.F90:103) */
    /* extra diagnostics that are locals / 1-D-only members in the reference */
    KPP_F_DIAG_ITER,    /* int32(npts): final `iter` of ocnstep (ocnstep_mod.F90:54,154) */
    KPP_F_DIAG_NREINT,  /* int32(npts): reset_flag before check_profile (ocnstep_mod.F90:228) */
    KPP_F_DIAG_STATUS,  /* int32(npts): KPP_ST_* bits */
    KPP_F_DIAG_TALPHA,  /* double(npts,0:nzp1): kpp_1d_fields%talpha */
    KPP_F_DIAG_SBETA,   /* double(npts,0:nzp1): kpp_1d_fields%sbeta  */
    KPP_F__COUNT
} kpp_field_id;

/* per-column status bits (KPP_F_DIAG_STATUS, kpp_gpu_get_status) */
#define KPP_ST_LONG_ITER   1   /* 'long iteration' warning        ocnstep_mod.F90:184-191 */
#define KPP_ST_REINT_FAIL  2   /* 'Failed to find a reasonable solution' ocnstep_mod.F90:229-236 */
#define KPP_ST_RESET       4   /* check_profile reset             overrides.F90:57-78 */
#define KPP_ST_PIVOT_ZERO  8   /* tridmat bet == 0                solvers.F90:140 */
#define KPP_ST_ITER_CAP   16   /* safety cap itermax + KPP_ITER_CAP_EXTRA on the goto-45 loop */
#define KPP_ST_ISO_RESET  32   /* isothermal reset                overrides.F90:116-120 */
#define KPP_ST_BAD_OLDNEW 64   /* 'Dodgy value of old/new'        ocnstep_mod.F90:93-102 */
#define KPP_ITER_CAP_EXTRA 1000

/* what kpp_gpu_sync reports about the step just finished */
typedef struct kpp_step_report {
    int32_t ntime;
    int32_t n_active;        /* columns with run_physics */
    int32_t n_long_iter;
    int32_t n_reint;         /* columns integrated more than once (instability trap) */
    int32_t n_reint_fail;
    int32_t n_reset;
    int32_t n_pivot_zero;
    int32_t n_iter_cap;
    int32_t max_iter;
    int32_t n_handed_over;   /* columns whose iteration the cooperative kernel finished (kpp_gpu_set_pass_budget) */
    int64_t sum_iter;        /* sum of final iter over active columns */
    float   kernel_ms;       /* device time of the step's kernels (CUDA events on the handle's stream) */
    float   reserved2;
} kpp_step_report;

typedef struct kpp_handle kpp_handle;

int kpp_gpu_abi_version(void);
/* 1 if the given numerics variant evaluates exp() with the host libm's own algorithm and
 * table (bit-identical to the CPU build's exp), 0 if it uses CUDA's exp (<= 1 ulp away) */
int kpp_gpu_exp_is_host_libm(int numerics);
int kpp_gpu_device_count(void);
const char *kpp_gpu_strerror(int code);
const char *kpp_gpu_last_error(const kpp_handle *h);

/* Create a handle on CUDA device `device` for npts columns.
 * zm(nzp1), hm(nzp1), dm(0:nz)   vertical grid    initialize_geography_mod.F90:43-74
 * tri(0:nztmax,0:1,1)            initialize_ocean.F90:34-43
 * wmt, wst (0:891,0:49)          physics_lookup_mod.F90:42-64
 * The library derives its per-level tables (Jerlov swfrac/swdk tables with
 * glibc exp as the reference's `ntime<=1` fill would, reference-integral
 * weights, dto/hm, ...) from these on the host at creation. */
int kpp_gpu_create(const kpp_dims *dims, const kpp_consts *consts,
                   const double *zm, const double *hm, const double *dm, const double *tri,
                   const double *wmt, const double *wst, int device, kpp_handle **out);
/* Multi-GPU inside the library: ONE host process, the reference's single-rank host
 * (mckpp_xios_control.F90:25; the column loop of mckpp_physics_driver_mod.F90:27-46 covers all npts),
 * `ngpus` devices of one node.  The npts columns are split into contiguous blocks of
 * ceil(npts/ngpus) columns (rounded up to whole 32-column tiles), one block, one stream and one
 * device mirror per GPU.  Every other call of this header takes the returned handle unchanged and
 * fans out: uploads/downloads/packed outputs move each device's block to or from its own column
 * slice of the host's whole arrays (disjoint slices, so there is no gather and no collective --
 * columns never exchange data), step/init launch on all devices at once, kpp_gpu_sync waits for
 * all of them and adds the reports up (kernel_ms = the slowest device).  Results are bit-identical
 * to a single-device handle: a column's arithmetic does not depend on its neighbours.
 * ngpus <= 0: all visible devices.  devices: ngpus device indices, or NULL for 0..ngpus-1. */
int kpp_gpu_create_multi(const kpp_dims *dims, const kpp_consts *consts,
                         const double *zm, const double *hm, const double *dm, const double *tri,
                         const double *wmt, const double *wst, int ngpus, const int *devices, kpp_handle **out);
int kpp_gpu_num_parts(const kpp_handle *h);          /* devices behind the handle (1 for kpp_gpu_create) */
/* which device holds which block of columns: device index, first column (0-based), column count */
int kpp_gpu_part_columns(const kpp_handle *h, int part, int *device, int *col0, int *ncols);
int kpp_gpu_destroy(kpp_handle *h);

/* host -> device / device -> host of one member of kpp_3d_fields; `bytes` must
 * equal the size of the whole host array (checked).  Asynchronous on the
 * handle's stream when the host buffer is pinned; upload/download/step are
 * ordered on that stream. */
int kpp_gpu_upload_field(kpp_handle *h, int field_id, const void *host, size_t bytes);
int kpp_gpu_download_field(kpp_handle *h, int field_id, void *host, size_t bytes);
/* same, but only enqueued on the handle's stream: the host buffer (pinned, for a truly
 * asynchronous copy) is valid after the next kpp_gpu_sync / kpp_gpu_download_field */
int kpp_gpu_download_field_async(kpp_handle *h, int field_id, void *host, size_t bytes);
size_t kpp_gpu_field_host_bytes(const kpp_handle *h, int field_id);
const char *kpp_gpu_field_name(int field_id);

/* per-step forcing: sflux6 = 6 rows of npts doubles = sflux(:,1:6,5,0)
 * (what mckpp_fluxes fills, fluxes_mod.F90:63-70) */
int kpp_gpu_upload_forcing(kpp_handle *h, const double *sflux6);

/* the forcing map of mckpp_fluxes on the device (src/mckpp_fluxes_mod.F90:56-72,
 * l_rest = .FALSE.).  The host passes the eight raw flux fields (npts doubles each) exactly as
 * mckpp_read_fluxes / the built-in constants deliver them; the device fills sflux(:,1:6,5,0)
 * for the ocean points (l_ocean), including the taux = 1e-10 guard when both stresses are zero.
 * flsn, el: kpp_const_fields%FLSN, %EL.  The l_rest = .TRUE. branch of mckpp_fluxes (all fluxes
 * zero: fluxes_mod.F90:51-55) is not provided: upload zeros, or skip the call and upload sflux. */
int kpp_gpu_upload_fluxes(kpp_handle *h, const double *taux, const double *tauy, const double *swf,
                          const double *lwf, const double *lhf, const double *shf, const double *rain,
                          const double *snow, double flsn, double el);

/* Forcing staged on the device ahead of time (a GPU-resident coupler, or a host that
 * uploads a forcing interval at once): reserve `nslots` buffers, fill slot i with the
 * same 6 x npts block as kpp_gpu_upload_forcing, then select which slot the next
 * kpp_gpu_step reads.  slot = -1 selects the buffer kpp_gpu_upload_forcing writes. */
int kpp_gpu_reserve_forcing_slots(kpp_handle *h, int nslots);
int kpp_gpu_upload_forcing_slot(kpp_handle *h, int slot, const double *sflux6);
int kpp_gpu_select_forcing_slot(kpp_handle *h, int slot);

/* number of this library's kernels launched on the handle since creation */
long long kpp_gpu_launch_count(const kpp_handle *h);

/* the per-column loop of MCKPP_INITIALIZE_OCEAN_MODEL (L_INITFLAG vmix at ntime=0,
 * initial diagnostic fluxes, old/new, hmixd, Us, Xs).  initialize_ocean.F90:54-104 */
int kpp_gpu_init_vmix(kpp_handle *h);

/* mckpp_physics_driver for timestep `ntime` (1-based), asynchronous. */
int kpp_gpu_step(kpp_handle *h, int ntime);

/* wait for the stream; fills the report of the LAST step; returns KPP_E_PIVOT_ZERO if any column
 * hit the tridiagonal zero pivot (the reference aborts there) in ANY step since the previous
 * kpp_gpu_sync -- steps may be queued without a sync in between, the condition is kept. */
int kpp_gpu_sync(kpp_handle *h, kpp_step_report *report);
/* Scheduling knob, no effect on results.  A column that has not converged after `budget`
 * passes of an integration (the reference iterates up to itermax = 200 where most columns need
 * 6) is handed from the one-thread-per-column kernel to a cooperative kernel that spreads one
 * column over a whole CTA, so that a few slow columns do not hold the step.  0 = never hand
 * over; -1 = every column runs in the cooperative kernel from its first pass (small domains: a
 * thread per column would leave the GPU empty).  Default 6, or -1 for domains of at most
 * KPP_SMALL_DOMAIN_COLUMNS columns (environment KPP_PASS_BUDGET overrides it at kpp_gpu_create). */
#define KPP_SMALL_DOMAIN_COLUMNS 1536
int kpp_gpu_set_pass_budget(kpp_handle *h, int budget);
/* Scheduling knob, no effect on results.  A column that stops converging needs ~6 ms in the
 * cooperative kernel (200 passes, the reference's itermax, one after the other) however small the
 * domain is.  With this mode on, the columns a step hands over are finished on a second stream and do
 * their following steps in the cooperative kernel on a third one, while the step kernel already runs
 * the next step of all other columns (a column's step n+1 depends on nothing but its own step n).
 * Everything that reads or writes device state through this header -- kpp_gpu_sync included -- first
 * joins the streams, so a host that syncs after every step behaves as before; a host that queues
 * steps (forcing slots resident on the device, outputs through the ring) no longer waits for the
 * stragglers.  Off by default; ignored while L_VARY_BOTTOM_TEMP is set or the pass budget is <= 0. */
int kpp_gpu_set_async_stragglers(kpp_handle *h, int on);
int kpp_gpu_get_status(kpp_handle *h, int32_t *status /* npts */);
/* Debugging aid: with KPP_GUARD=1 in the environment at kpp_gpu_create every device array of the handle sits
 * between two 64 KB canary zones; this returns how many zones a kernel has written into (0 = clean), or a
 * negative error.  (A stand-in where compute-sanitizer cannot be used; it sees stray writes, not reads.) */
int kpp_gpu_debug_check_guards(kpp_handle *h);

/* ---- the output sets of the host I/O layer, packed on the device -------------
 * Each id is one xios_send_field of mckpp_xios_diagnostic_output (xios_io.F90:72-207) or
 * mckpp_xios_restart_output (:406-431), delivered in the shape that call sends: a dense
 * double(npts, rows) block, column index fastest.  The temp_2d reshuffles the reference does on
 * the host every output step happen on the device: S = X(:,k,2)+Sref (:94-97); difm/dift/difs
 * shifted down one level under a zero top row (:120-133); dbloc padded with a zero bottom row
 * (:148-150); REAL(old), REAL(new) (:425-426); Us/Vs/Ts/Ss = one component of both saved time
 * levels (:427-430).  "cplwght" (:186-192) and "time" (:412) are host data and not provided. */
typedef enum kpp_out_id {
    KPP_OUT_U = 0,      /* "u"          U(:,:,1)                 rows nzp1 */
    KPP_OUT_V,          /* "v"          U(:,:,2) */
    KPP_OUT_T,          /* "T"          X(:,:,1) */
    KPP_OUT_S,          /* "S"          X(:,k,2)+Sref(:) */
    KPP_OUT_B,          /* "B"          buoy(:,1:NZP1) */
    KPP_OUT_WU,         /* "wu"         wU(:,0:NZ,1) */
    KPP_OUT_WV,         /* "wv"         wU(:,0:NZ,2) */
    KPP_OUT_WT,         /* "wT"         wX(:,0:NZ,1) */
    KPP_OUT_WS,         /* "wS"         wX(:,0:NZ,2) */
    KPP_OUT_WB,         /* "wB"         wX(:,0:NZ,NSP1) */
    KPP_OUT_WTNT,       /* "wTnt"       wXNT(:,0:NZ,1) */
    KPP_OUT_DIFM,       /* "difm"       (:,1)=0, (:,2:NZP1)=difm(:,1:NZ) */
    KPP_OUT_DIFT,       /* "dift" */
    KPP_OUT_DIFS,       /* "difs" */
    KPP_OUT_RHO,        /* "rho"        rho(:,1:NZP1) */
    KPP_OUT_CP,         /* "cp"         cp(:,1:NZP1) */
    KPP_OUT_SCORR,      /* "scorr" */
    KPP_OUT_RIG,        /* "Rig"        rows 1:NZ; row NZP1 is never written by the physics: 0 */
    KPP_OUT_DBLOC,      /* "dbloc"      (:,1:NZ)=dbloc, (:,NZP1)=0 */
    KPP_OUT_SHSQ,       /* "Shsq"       like Rig */
    KPP_OUT_TINC_FCORR, /* "tinc_fcorr" */
    KPP_OUT_FCORR_Z,    /* "fcorr_z"    ocnTcorr */
    KPP_OUT_SINC_FCORR, /* "sinc_fcorr" */
    KPP_OUT_HMIX,       /* "hmix"       rows 1 from here on */
    KPP_OUT_FCORR,      /* "fcorr" */
    KPP_OUT_TAUX_IN,    /* "taux_in"    sflux(:,1,5,0) */
    KPP_OUT_TAUY_IN,    /* "tauy_in"    sflux(:,2,5,0) */
    KPP_OUT_SOLAR_IN,   /* "solar_in"   sflux(:,3,5,0) */
    KPP_OUT_NSOLAR_IN,  /* "nsolar_in"  sflux(:,4,5,0) */
    KPP_OUT_PMINUSE_IN, /* "PminusE_in" sflux(:,6,5,0) */
    KPP_OUT_FREEZE_FLAG,/* "freeze_flag" */
    KPP_OUT_COMP_FLAG,  /* "comp_flag"  reset_flag */
    KPP_OUT_DAMPU_FLAG, /* "dampu_flag" */
    KPP_OUT_DAMPV_FLAG, /* "dampv_flag" */
    KPP_OUT_R_UVEL,     /* restart "uvel"  U(:,:,1)          rows nzp1 */
    KPP_OUT_R_VVEL,     /* restart "vvel"  U(:,:,2) */
    KPP_OUT_R_T,        /* restart "T"     X(:,:,1) */
    KPP_OUT_R_S,        /* restart "S"     X(:,:,2) (no Sref) */
    KPP_OUT_R_CP,       /* restart "CP"    cp(:,1:NZP1) */
    KPP_OUT_R_RHO,      /* restart "rho"   rho(:,1:NZP1) */
    KPP_OUT_R_HMIX,     /* restart "hmix"                    rows 1 */
    KPP_OUT_R_KMIX,     /* restart "kmix" */
    KPP_OUT_R_SREF,     /* restart "Sref" */
    KPP_OUT_R_SSREF,    /* restart "SSref" */
    KPP_OUT_R_SSURF,    /* restart "Ssurf" */
    KPP_OUT_R_TREF,     /* restart "Tref" */
    KPP_OUT_R_OLD,      /* restart "old"   REAL(old) */
    KPP_OUT_R_NEW,      /* restart "new"   REAL(new) */
    KPP_OUT_R_US,       /* restart "Us"    Us(:,:,1,0:1)     rows 2*nzp1 */
    KPP_OUT_R_VS,       /* restart "Vs"    Us(:,:,2,0:1) */
    KPP_OUT_R_TS,       /* restart "Ts"    Xs(:,:,1,0:1) */
    KPP_OUT_R_SS,       /* restart "Ss"    Xs(:,:,2,0:1) */
    KPP_OUT_R_HMIXD,    /* restart "hmixd" hmixd(:,0:1)      rows 2 */
    KPP_OUT__COUNT
} kpp_out_id;
#define KPP_OUT__FIRST_RESTART KPP_OUT_R_UVEL

const char *kpp_gpu_output_name(int out_id);                 /* the XIOS field id of that send */
int kpp_gpu_output_rows(const kpp_handle *h, int out_id);    /* rows of the block (negative: error) */
/* Pack on the device, copy to `host` (npts*rows doubles).  The _async form only enqueues on the
 * handle's stream (pinned host memory for a true overlap); kpp_gpu_sync completes it. */
int kpp_gpu_pack_output(kpp_handle *h, int out_id, double *host, size_t bytes);
int kpp_gpu_pack_output_async(kpp_handle *h, int out_id, double *host, size_t bytes);

/* ---- asynchronous output ring -------------------------------------------------------------------
 * The reference's main loop calls mckpp_output_control() after EVERY physics step
 * (mckpp_ocean_model_3D.F90:62; mckpp_xios_control.F90:52-57) and mckpp_xios_diagnostic_output hands
 * XIOS up to 34 blocks each time (mckpp_xios_io.F90:72-207).  Pulled synchronously that set is a PCIe
 * transfer several times longer than the step.  The ring delivers a chosen set of kpp_out_id blocks
 * through `depth` pinned host slots: submit (after kpp_gpu_step) packs the blocks into device
 * staging on the step's stream -- a device-to-device copy -- and moves them to the host on a second
 * stream, so the copy of step n overlaps the kernels of step n+1.  A slot is laid out as the blocks
 * in the order given at creation, each dense double(npts, rows) as kpp_gpu_pack_output delivers it;
 * it is valid from kpp_gpu_output_ring_wait until it is submitted again `depth` submits later.
 * Works on multi-GPU handles (every device fills its column slice of the same host slot). */
int kpp_gpu_output_ring_create(kpp_handle *h, const int32_t *out_ids, int n_ids, int depth);
size_t kpp_gpu_output_ring_slot_bytes(const kpp_handle *h);
size_t kpp_gpu_output_ring_offset(const kpp_handle *h, int index);   /* byte offset of out_ids[index] in a slot */
int kpp_gpu_output_ring_submit(kpp_handle *h, int *slot);            /* enqueue only */
int kpp_gpu_output_ring_wait(kpp_handle *h, int slot, double **host);
int kpp_gpu_output_ring_destroy(kpp_handle *h);

/* ---- climatology time interpolation on the device ------------------------------
 * MCKPP_BOUNDARY_INTERPOLATE_TEMP / _SAL (boundary_interpolate.F90:14-123) read the two records
 * that bracket `time` and set  clim = next*next_weight + prev*prev_weight  over npts x NZP1.
 * The host keeps reading the files (and computes the weights, hostinit.boundary_interp_weights
 * mirrors :27-36,51-52 with the reference's INTEGER truncations); the two records stay
 * resident on the device and are re-blended into KPP_F_OCNT_CLIM / KPP_F_SAL_CLIM every
 * ndt_interp steps without another upload until the bracket moves.
 * id = KPP_F_OCNT_CLIM or KPP_F_SAL_CLIM; which = 0 (prev) / 1 (next); record = double(npts,nzp1). */
int kpp_gpu_upload_clim_record(kpp_handle *h, int id, int which, const double *record, size_t bytes);
int kpp_gpu_blend_clim(kpp_handle *h, int id, double prev_weight, double next_weight);

/* pinned host memory helpers (for asynchronous, full-rate PCIe copies) */
int kpp_gpu_host_alloc(void **ptr, size_t bytes);
int kpp_gpu_host_free(void *ptr);

/* unit-test entry points: device evaluation of single routines on n points */
int kpp_gpu_test_eos(int device, int numerics, int n, const double *S, const double *T, const double *P,
                     double *sig0, double *alpha, double *beta, double *cp);
int kpp_gpu_test_wscale(kpp_handle *h, int n, const double *sigma, const double *hbl, const double *ustar,
                        const double *bfsfc, double *wm, double *ws);
int kpp_gpu_test_swfrac(int device, int numerics, int n, const double *z, const int32_t *jerlov, double *out);
/* the cooperative kernel's split division (reciprocal taken off the dependency chain) next to
 * the plain a/b: plain[i] = a/b, split[i] = the split form, ok[i] = 1.0 where the split form
 * claims validity (it must then equal plain bit for bit) */
int kpp_gpu_test_div(int device, int numerics, int n, const double *a, const double *b, double *plain, double *split,
                     double *ok);

#ifdef __cplusplus
}
#endif
#endif /* KPP_GPU_H */
