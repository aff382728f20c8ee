// kpp_dev.h -- device-side argument block shared by kpp_api.cu and kpp_kernels.cu.
//
// Data layout in HBM: every per-column field is a structure of
// arrays with COLUMNS AS THE FAST INDEX, `field[row * ld + c]`, row = level (or
// component*levels + level), ld = npts rounded up to 32 columns so each warp's
// 32 x 8 B access to one row is one aligned 256 B segment.  This is the same
// element order as the reference's `kpp_3d_fields` members
// (src/mckpp_data_fields.F90:353-447, first extent npts), only padded.
#pragma once
#include <stddef.h>
#include <stdint.h>

struct KppDevArgs {
    // ---- dimensions
    int npts, ld, nz, nzp1, ntime, itermax, iso_bot, maxmodeadv;
    // ---- switches (kpp_const_type LOGICALs)
    int LRI, LDD, L_SSref, L_RELAX_SST, L_RELAX_CALCONLY, L_FCORR, L_FCORR_WITHZ, L_SFCORR,
        L_SFCORR_WITHZ, L_RELAX_SAL, L_RELAX_OCNT, L_NO_FREEZE, L_NO_ISOTHERM, L_DAMP_CURR,
        have_clim_files, pad_;
    // ---- scalars
    double dto, grav, vonk, sice, hmixtolfrac, iso_thresh;
    double cg;        // cstar*vonk*(cs*vonk*epsilon)**(1./3.)  blmix_mod.F90:62 (host pow)
    double Vtc;       // cv*sqrt(0.2/cs/epsilon)/vonk**2/Ricr   bldepth_mod.F90:91
    double uvdamp;    // dt_uvdamp*(86400./dto)                 ocnstep_mod.F90:322
    double dmNZ;      // dm(NZ)                                 ocnstep_mod.F90:211
    // ---- grid tables, all addressed with the FORTRAN index (slot 0 unused for 1-based ones)
    const double *zm;      // zm(1:nzp1)
    const double *hm;      // hm(1:nzp1)
    const double *dm;      // dm(0:nz)
    const double *tri0;    // tri(k,0,1)  k=0:nz
    const double *tri1;    // tri(k,1,1)
    const double *p0;      // (-zm(k))/10.            pressure in bars (state_equations.F90:33,402)
    const double *dzb;     // zm(k)-zm(k+1)           k=1:nz
    const double *dtoh;    // dto/hm(k)               k=1:nzp1
    const double *deltaz;  // 0.5*(hm(k)+hm(k+1))     k=1:nz   (ocnstep_mod.F90:243)
    const double *zint;    // -zm(k)+0.5*hm(k)        k=1:nz   (blmix_mod.F90:112)
    const double *zref;    // epsilon*zm(n)           n=1:nz   (verticalmixing_mod.F90:112)
    const double *wz0;     // AMAX1(zm(1),zref(n))
    const double *zrmz;    // zref(n)-zm(n)
    const int    *refoff;  // CSR offsets of the reference-integral trips of level n (n=1:nz+1)
    const double *refwz;   // wz  of each trip       (verticalmixing_mod.F90:120)
    const double *refdel;  // del of each trip       (:121)
    const double *swfrac_tab; // [5][nzp1+1]  MCKPP_PHYSICS_SWFRAC_OPT per Jerlov type (swfrac_mod.F90:36-41)
    const double *swdk_tab;   // [5][nz+1]    mckpp_fluxes_swdk(-dm(k), j)             (fluxes_mod.F90:103-108)
    const double2 *wtab;      // interleaved {wmt(i,j), wst(i,j)}, i fastest (0:891, 0:49)
    // ---- state / inputs (column-fastest SoA, leading dimension ld)
    double *U, *X, *Us, *Xs, *hmixd;
    int *old_, *new_;
    double *hmix, *kmix, *Tref, *uref, *vref, *Ssurf;
    const double *Sref, *SSref, *f, *ocdepth;
    const int *jerlov, *l_ocean, *run_physics;
    const double *sflux;      // 6 rows: sflux(:,1:6,5,0)
    const double *U_init, *relax_sst, *SST0, *fcorr_twod;
    double *fcorr;
    const double *relax_sal, *relax_ocnT, *sal_clim, *ocnT_clim, *fcorr_withz, *sfcorr_withz, *bottom_temp;
    const int *nmodeadv, *modeadv;
    const double *advection;
    double *freeze_flag, *reset_flag, *dampu_flag, *dampv_flag;
    // ---- diagnostics
    double *rho, *cp;         // rows 0:nzp1
    double *buoy;             // rows 1:nzp1 stored at row k-1
    double *Rig, *dbloc, *Shsq;  // rows k-1
    double *difm, *difs, *dift;  // rows 0:nzp1
    double *ghat;             // rows 1:nz at k-1
    double *wU;               // 2 comps x rows 0:nz
    double *wX;               // 3 comps x rows 0:nz
    double *wXNT;             // rows 0:nz
    double *tinc_fcorr, *sinc_fcorr, *ocnTcorr, *scorr;   // rows k-1, k=1:nzp1
    double *swfrac, *swdk_opt;
    int *diag_iter, *diag_nreint, *diag_status;
    double *talpha, *sbeta;   // rows 0:nzp1
    // ---- scratch (never crosses the ABI): tile-major records, see kpp_kernels.cu
    double *scr;              // [tile = column/32][level 0..nzp1][KPP_NF fields][32 lanes]
    // ---- straggler hand-over (kpp_step_kernel -> kpp_coop_kernel), see kpp_kernels.cu
    int pass_budget;          // passes of one integration the per-thread kernel runs itself; 0 = all of them
    int buoy_margin;          // the sweep stores buoyancy down to (last kbl + margin) only, see Tabs::kbuoy
    struct KppCont *cont;     // [npts] continuation record of a handed-over column
    int *cont_list;           // [npts] handed-over columns of this step
    int *cont_count;          // how many
    int *cont_next;           // next list entry to hand out: the cooperative CTAs fetch their columns one by one
    int *tile_counter;        // persistent step kernel: tiles handed out beyond every warp's first one
    int coop_expect;          // host's guess of the length of the next hand-over list (the last report's): sizes the cooperative launch
    int pad2_;
    // ---- asynchronous stragglers (kpp_gpu_set_async_stragglers; protocol in kpp_api.cu): a column the step
    // kernel hands over is finished on a second stream while the next step of everyone else already runs;
    // until the next join it does its steps in the cooperative kernel (the "lane") on a third stream
    int *in_lane;             // [ld] 1 = the step kernel skips the column (null: synchronous hand-over)
    int *lane_out_list;       // lane list of the NEXT step: a cooperative launch appends the columns it finished
    int *lane_out_count;      //   (null: do not append)
    int *pivot_sticky;        // zero pivots since the last kpp_gpu_sync (counted by the kernels themselves)
};

// loop state of a column whose iteration is continued by the cooperative kernel
struct KppCont {
    double hmixe, f;
    int iter, iconv, kmixe, nreint, status, pad_;
};

#define KPP_NF 20             // fields per scratch record

struct KppReportDev {
    int n_active, n_long_iter, n_reint, n_reint_fail, n_reset, n_pivot_zero, n_iter_cap, max_iter;
    int n_handed_over, pad_;
    long long sum_iter;
    // everything above is cleared before each step's report; this one only by kpp_gpu_sync, so a zero
    // pivot in a step that was queued behind others without a sync in between is still reported
    int pivot_sticky, pad2_;
};
#define KPP_REPORT_CLEAR_BYTES (offsetof(KppReportDev, pivot_sticky))
