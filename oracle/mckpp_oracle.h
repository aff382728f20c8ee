/*
 * mckpp_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C, scalar, one column at a time) of the MC-KPP
 * per-column physics timestep of aosprey/mckpp-f90.  It is the checker the
 * CUDA path is compared against; it is never shipped, never on the product
 * path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it.
 *
 * PARITY PINNING: the reference holds no tests and cannot be compiled in this
 * image (no Fortran compiler, no MPI/netCDF/XIOS).  The only known-answer
 * values in the reference tree are the three equation-of-state check values
 * (src/mckpp_physics_state_equations.F90:24-25,105-111); the oracle is pinned
 * on those (tests/test_oracle_kat.py).  Everything else is "parity unpinned":
 * pinned only by this literal transcription, which the independent CUDA
 * implementation cross-checks.
 *
 * Memory image: every array argument has exactly the shape and (column-major)
 * element order of the corresponding Fortran array of kpp_3D_type /
 * kpp_const_type (src/mckpp_data_fields.F90:353-447, 492-501, and
 * src/mckpp_initialize_namelist_mod.F90:87-89), REAL = 8 bytes
 * (-fdefault-real-8), INTEGER/LOGICAL = 4 bytes.
 */
#ifndef MCKPP_ORACLE_H
#define MCKPP_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* kpp_const_type subset read by the hot path + mckpp_parameters dimensions */
typedef struct orc_const {
    /* mckpp_parameters (src/mckpp_parameters.F90:4-59) */
    int32_t nz, nzp1, nztmax, nzp1tmax;
    int32_t npts;
    int32_t nsflxs, njdt, maxmodeadv;
    int32_t itermax;
    int32_t iso_bot, dt_uvdamp;
    /* LOGICALs (src/mckpp_data_fields.F90:262-323) */
    int32_t LKPP, LRI, LDD, L_SSref;
    int32_t L_RELAX_SST, L_RELAX_CALCONLY, L_FCORR, L_FCORR_WITHZ;
    int32_t L_SFCORR, L_SFCORR_WITHZ, L_RELAX_SAL, L_RELAX_OCNT;
    int32_t L_NO_FREEZE, L_NO_ISOTHERM, L_DAMP_CURR, L_VARY_BOTTOM_TEMP;
    int32_t have_ocnT_file;   /* ocnT_file .ne. 'none' */
    int32_t have_sal_file;    /* sal_file  .ne. 'none' */
    int32_t pad0;
    double hmixtolfrac;
    double dto, grav, vonk, sice, iso_thresh;
    /* arrays */
    const double *zm;   /* zm(nzp1)            */
    const double *hm;   /* hm(nzp1)            */
    const double *dm;   /* dm(0:nz)            */
    const double *tri;  /* tri(0:nztmax,0:1,1) */
    const double *wmt;  /* wmt(0:891,0:49)     */
    const double *wst;  /* wst(0:891,0:49)     */
} orc_const;

/* kpp_3D_type subset touched by 3dto1d / 1dto3d / bottomtemp.
 * Shapes in comments are the Fortran ones. */
typedef struct orc_3d {
    double *U;            /* (npts,nzp1,2)            */
    double *X;            /* (npts,nzp1,2)            */
    double *Rig;          /* (npts,nzp1)              */
    double *dbloc;        /* (npts,nz)                */
    double *Shsq;         /* (npts,nzp1)              */
    double *hmixd;        /* (npts,0:1)               */
    double *Us;           /* (npts,nzp1,2,0:1)        */
    double *Xs;           /* (npts,nzp1,2,0:1)        */
    double *rho;          /* (npts,0:nzp1tmax)        */
    double *cp;           /* (npts,0:nzp1tmax)        */
    double *buoy;         /* (npts,nzp1tmax)          */
    double *ocdepth;      /* (npts)                   */
    double *f;            /* (npts)                   */
    double *swfrac;       /* (npts,nzp1)              */
    double *swdk_opt;     /* (npts,0:nz)              */
    double *difm;         /* (npts,0:nztmax)          */
    double *difs;         /* (npts,0:nztmax)          */
    double *dift;         /* (npts,0:nztmax)          */
    double *wU;           /* (npts,0:nztmax,3)        */
    double *wX;           /* (npts,0:nztmax,3)        */
    double *wXNT;         /* (npts,0:nztmax,2)        */
    double *ghat;         /* (npts,nztmax)            */
    double *relax_sst;    /* (npts)                   */
    double *fcorr;        /* (npts)                   */
    double *SST0;         /* (npts)                   */
    double *fcorr_twod;   /* (npts)                   */
    double *tinc_fcorr;   /* (npts,nzp1)              */
    double *sinc_fcorr;   /* (npts,nzp1)              */
    double *fcorr_withz;  /* (npts,nzp1)              */
    double *sfcorr_withz; /* (npts,nzp1)              */
    double *advection;    /* (npts,maxmodeadv,2)      */
    double *relax_sal;    /* (npts)                   */
    double *scorr;        /* (npts,nzp1)              */
    double *relax_ocnT;   /* (npts)                   */
    double *ocnTcorr;     /* (npts,nzp1)              */
    double *sal_clim;     /* (npts,nzp1)              */
    double *ocnT_clim;    /* (npts,nzp1)              */
    double *hmix;         /* (npts)                   */
    double *kmix;         /* (npts)  REAL             */
    double *Tref;         /* (npts)                   */
    double *uref;         /* (npts)                   */
    double *vref;         /* (npts)                   */
    double *Ssurf;        /* (npts)                   */
    double *Sref;         /* (npts)                   */
    double *SSref;        /* (npts)                   */
    double *sflux;        /* (npts,nsflxs,5,0:njdt)   */
    double *freeze_flag;  /* (npts)                   */
    double *reset_flag;   /* (npts)                   */
    double *dampu_flag;   /* (npts)                   */
    double *dampv_flag;   /* (npts)                   */
    double *U_init;       /* (npts,nzp1,2)            */
    double *bottom_temp;  /* (npts)                   */
    int32_t *l_ocean;     /* (npts) LOGICAL           */
    int32_t *l_initflag;  /* (npts) LOGICAL           */
    int32_t *run_physics; /* (npts) LOGICAL           */
    int32_t *old;         /* (npts)                   */
    int32_t *new_;        /* (npts)                   */
    int32_t *jerlov;      /* (npts)                   */
    int32_t *nmodeadv;    /* (npts,2)                 */
    int32_t *modeadv;     /* (npts,maxmodeadv,2)      */
    /* --- oracle-only diagnostics (not in the reference's types; may be NULL):
     * locals of ocnstep/1-D-only fields that BASELINE.json wants compared */
    int32_t *diag_iter;      /* (npts) final `iter` of ocnstep (ocnstep_mod.F90:54) */
    int32_t *diag_nreint;    /* (npts) reset_flag BEFORE check_profile (ocnstep_mod.F90:228) */
    int32_t *diag_status;    /* (npts) bit flags, see ORC_ST_* */
    double  *diag_talpha;    /* (npts,0:nzp1) kpp_1d_fields%talpha */
    double  *diag_sbeta;     /* (npts,0:nzp1) kpp_1d_fields%sbeta  */
} orc_3d;

#define ORC_ST_LONG_ITER   1   /* iter > itermax+1 warning (ocnstep_mod.F90:184) */
#define ORC_ST_REINT_FAIL  2   /* reset_flag > comp_iter_max (ocnstep_mod.F90:229) */
#define ORC_ST_RESET       4   /* check_profile reset to climatology / U_init (overrides.F90:57-78) */
#define ORC_ST_PIVOT_ZERO  8   /* tridmat bet == 0 (solvers.F90:140) */
#define ORC_ST_ITER_CAP   16   /* safety cap on the goto-45 loop hit (oracle/GPU extension) */
#define ORC_ST_ISO_RESET  32   /* isothermal reset (overrides.F90:116-120) */
#define ORC_ST_BAD_OLDNEW 64   /* 'Dodgy value of old/new' (ocnstep_mod.F90:93-102) */

/* safety cap on the data-dependent goto-45 loop: iter may exceed itermax
 * while hmixn keeps growing (ocnstep_mod.F90:175-181).  The reference has no
 * cap; both the oracle and the GPU stop at itermax + ORC_ITER_CAP_EXTRA and
 * flag ORC_ST_ITER_CAP so a pathological column cannot hang a run. */
#define ORC_ITER_CAP_EXTRA 1000

/* comma-separated member names of orc_3d in declaration order (for binding checks) */
const char *orc_3d_member_names(void);
const char *orc_const_member_names(void);

/* mckpp_physics_driver (src/mckpp_physics_driver_mod.F90:15-73).
 * nthreads<=0: use omp default.  realloc_1d!=0: re-allocate the 1-D column
 * type on every call like mckpp_fields_3dto1d does (types_transfer.F90:24).
 * Returns 0, or -1 if any column hit the tridmat zero-pivot abort. */
int orc_physics_driver(const orc_const *c, orc_3d *s, int ntime, int nthreads, int realloc_1d);

/* the per-column loop of MCKPP_INITIALIZE_OCEAN_MODEL (src/mckpp_initialize_ocean.F90:54-104),
 * run with ntime = 0 as the reference does. */
int orc_initialize_ocean_model(const orc_const *c, orc_3d *s, int nthreads);

/* input builders restated from the reference's init code */
void orc_build_tri(int nz, int nztmax, double dto, const double *zm, const double *hm,
                   double *tri /* (0:nztmax,0:1,1), zero-filled first */);           /* initialize_ocean.F90:34-43 */
void orc_build_lookup(double vonk, double *wmt, double *wst /* (0:891,0:49) */);      /* physics_lookup_mod.F90:42-64 */
void orc_build_grid(int nz, double dmax, int l_stretchgrid, double dscale,
                    double *zm, double *hm, double *dm);                             /* initialize_geography_mod.F90:43-74 */
void orc_coriolis(int npts, const double *dlat, double *f);                          /* initialize_geography_mod.F90:78-88 */
/* mckpp_fluxes forcing map (src/mckpp_fluxes_mod.F90:59-70), l_rest = .FALSE. */
void orc_fluxes_map(int npts, int nsflxs, double flsn, double el,
                    double *taux, const double *tauy, const double *swf, const double *lwf,
                    const double *lhf, const double *shf, const double *rain, const double *snow,
                    const int32_t *l_ocean, double *sflux);

/* single-routine entry points for unit parity / KATs */
double orc_cpsw(double s, double t1, double p0);                                     /* state_equations.F90:7-58 */
void orc_abk80(double s, double t1, double p, double *alpha, double *beta,
               double *kappa, double *sig0, double *sig);                            /* state_equations.F90:133-190 */
void orc_wscale(const orc_const *c, double sigma, double hbl, double ustar, double bfsfc,
                double *wm, double *ws);                                             /* wscale_mod.F90:12-97 */
void orc_swfrac(double fact, double z, int jwtype, double *swdk);                    /* swfrac_mod.F90:49-79 */
double orc_swdk(double z, int j);                                                    /* fluxes_mod.F90:121-137 */
void orc_tridmat(const double *cu, const double *cc, const double *cl, const double *rhs,
                 const double *yo, int nzi, double *yn, int nztmax, int *pivot_zero);/* solvers.F90:112-161 */

/* second-reading support (oracle/second_reading.py): one kpp_1d_fields column driven from Python */
void *orc_col_new(const orc_const *c);
void orc_col_free(void *col);
void orc_col_load(void *col, const orc_const *c, const orc_3d *s, int point, int ntime);      /* 3dto1d */
void orc_col_store(void *col, const orc_const *c, orc_3d *s, int point, int iter_final, int nreint); /* 1dto3d */
void orc_col_vmix(void *col, const orc_const *c, double *hmix, int *kmix);
void orc_col_ocnint(void *col, const orc_const *c, int kmix, const double *Uo, const double *Xo);
void orc_col_check_profile(void *col, const orc_const *c);
double *orc_col_array(void *col, const char *name, int *lb, int *n);
double orc_col_get(void *col, const char *name);
void orc_col_set(void *col, const char *name, double v);
int orc_col_modeadv(void *col, int j, int i);

#ifdef __cplusplus
}
#endif
#endif
