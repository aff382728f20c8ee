/*
 * mckpp_oracle.c -- TEST INFRASTRUCTURE ONLY (see mckpp_oracle.h).
 *
 * Literal, scalar C restatement of the MC-KPP column-physics hot path of
 * aosprey/mckpp-f90.  Each function cites the reference file:line it follows.
 * Written to be read side by side with the Fortran: same loop order, same
 * operator association, same temporaries, same quirks.
 *
 * Build:  gcc -O2 -std=c11 -ffp-contract=off -fno-fast-math -fopenmp
 *   (mimics fcm-make-gfortran-local.cfg:5 on baseline x86-64: no FMA
 *    contraction, IEEE divide/sqrt, glibc exp/pow; REAL = double.)
 *
 * PARITY: pinned only on the reference's three EOS check values
 * (state_equations.F90:24-25,105-111); otherwise "parity unpinned".
 *
 * Indexing convention in this file: every array is addressed with its
 * FORTRAN index.  1-based Fortran dimensions get one unused leading slot in
 * the first dimension; 0-based ones map directly.
 */
#include "mckpp_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define NVEL 2
#define NSCLR 2
#define NSP1 3
#define NVP1 3

/* ------------------------------------------------------------------ */
/* kpp_1D_type (src/mckpp_data_fields.F90:104-184)                     */
/* ------------------------------------------------------------------ */
typedef struct col1d {
    double rhoh2o, ocdepth, f, relax_sst, fcorr, SST0, fcorr_twod;
    double relax_sal, relax_ocnT, hmix, kmix, Tref, uref, vref, Ssurf, Sref, SSref;
    double reset_flag, dampu_flag, dampv_flag, freeze_flag;
    double *U, *U_init, *X, *Rig, *dbloc, *Shsq, *hmixd, *Us, *Xs, *rho, *cp, *buoy;
    double *swfrac, *swdk_opt, *difm, *difs, *dift, *wU, *wX, *wXNT, *ghat;
    double *tinc_fcorr, *sinc_fcorr, *fcorr_withz, *sfcorr_withz, *advection;
    double *scorr, *ocnTcorr, *sal_clim, *ocnT_clim, *sflux, *talpha, *sbeta;
    int old, new_, jerlov, nmodeadv[3], point;
    int *modeadv;
    int l_ocean, l_initflag, comp_flag;
    /* oracle-only */
    int ntime;        /* module variable ntime (src/mckpp_time_control.F90:13) */
    int iter_final;   /* local `iter` of ocnstep */
    int nreint;       /* reset_flag before check_profile */
    int status;
    /* arguments and results of the last bldepth call (second-reading probes, oracle/second_reading.py) */
    double dbg_ustar, dbg_Bo, dbg_Bosol, dbg_hbl, dbg_bfsfc, dbg_stable, dbg_caseA;
    int dbg_kbl;
    /* dims for the index macros */
    int nz, nzp1, nztmax, nzp1tmax, nsflxs, njdt, maxmodeadv;
    /* scratch locals of ocnstep/ocnint/kppmix/tridmat (sized once) */
    double *Uo, *Xo, *Ux, *Xx;
    double *dVsq, *Ritop, *alphaDT, *betaDS;
    double *blmc;
    double *cu, *cc, *cl, *rhs, *diff, *gcap, *ntflx, *gam;
    double *zw; /* z121 weights: kppmix passes difs as `w` */
} col1d;

/* (k,l) with k=1..nzp1, l=1..2 */
#define U_(p,k,l)      ((p)->U[((l)-1)*((p)->nzp1+1) + (k)])
#define UI_(p,k,l)     ((p)->U_init[((l)-1)*((p)->nzp1+1) + (k)])
#define X_(p,k,l)      ((p)->X[((l)-1)*((p)->nzp1+1) + (k)])
#define US_(p,k,l,t)   ((p)->Us[(((t)*2 + (l)-1))*((p)->nzp1+1) + (k)])
#define XS_(p,k,l,t)   ((p)->Xs[(((t)*2 + (l)-1))*((p)->nzp1+1) + (k)])
#define UO_(p,k,l)     ((p)->Uo[((l)-1)*((p)->nzp1+1) + (k)])
#define XO_(p,k,l)     ((p)->Xo[((l)-1)*((p)->nzp1+1) + (k)])
#define UX_(p,k,l)     ((p)->Ux[((l)-1)*((p)->nzp1+1) + (k)])
#define XX_(p,k,l)     ((p)->Xx[((l)-1)*((p)->nzp1+1) + (k)])
/* (i,j) with i=0..nztmax, j=1.. */
#define WU_(p,i,j)     ((p)->wU[((j)-1)*((p)->nztmax+1) + (i)])
#define WX_(p,i,j)     ((p)->wX[((j)-1)*((p)->nztmax+1) + (i)])
#define WXNT_(p,i,j)   ((p)->wXNT[((j)-1)*((p)->nztmax+1) + (i)])
/* sflux(i,j,k) i=1..nsflxs, j=1..5, k=0..njdt */
#define SFLUX_(p,i,j,k) ((p)->sflux[(((k)*5 + (j)-1))*((p)->nsflxs) + (i)-1])
/* modeadv(j,i) j=1..maxmodeadv, i=1..2 ; advection(j,i) */
#define MODEADV_(p,j,i) ((p)->modeadv[((i)-1)*((p)->maxmodeadv) + (j)-1])
#define ADVEC_(p,j,i)   ((p)->advection[((i)-1)*((p)->maxmodeadv) + (j)-1])
/* blmc(ki,m) ki=1..km, m=1..3 */
#define BLMC_(p,ki,m)  ((p)->blmc[((m)-1)*((p)->nz+1) + (ki)])
/* const: tri(k,j,1) k=0..nztmax, j=0..1 ; wmt(i,j) i=0..891, j=0..49 */
#define TRI_(c,k,j)    ((c)->tri[(j)*((c)->nztmax+1) + (k)])
#define WMT_(c,i,j)    ((c)->wmt[(j)*892 + (i)])
#define WST_(c,i,j)    ((c)->wst[(j)*892 + (i)])
#define ZM_(c,k)       ((c)->zm[(k)-1])
#define HM_(c,k)       ((c)->hm[(k)-1])
#define DM_(c,k)       ((c)->dm[(k)])

static double *dalloc(size_t n) { return (double *)calloc(n ? n : 1, sizeof(double)); }

/* mckpp_allocate_1d_fields (src/mckpp_data_fields.F90:450-489) + ocnstep/ocnint locals */
static void col1d_alloc(col1d *p, const orc_const *c)
{
    memset(p, 0, sizeof(*p));
    p->nz = c->nz; p->nzp1 = c->nzp1; p->nztmax = c->nztmax; p->nzp1tmax = c->nzp1tmax;
    p->nsflxs = c->nsflxs; p->njdt = c->njdt; p->maxmodeadv = c->maxmodeadv;
    size_t n1 = (size_t)c->nzp1 + 1, nt = (size_t)c->nztmax + 1, ntt = (size_t)c->nzp1tmax + 1;
    p->U = dalloc(n1 * 2); p->U_init = dalloc(n1 * 2); p->X = dalloc(n1 * 2);
    p->Rig = dalloc(n1); p->dbloc = dalloc((size_t)c->nz + 1); p->Shsq = dalloc(n1);
    p->hmixd = dalloc(2); p->Us = dalloc(n1 * 4); p->Xs = dalloc(n1 * 4);
    p->rho = dalloc(ntt); p->cp = dalloc(ntt); p->buoy = dalloc(ntt);
    p->swfrac = dalloc(n1); p->swdk_opt = dalloc((size_t)c->nz + 1);
    p->difm = dalloc(nt); p->difs = dalloc(nt); p->dift = dalloc(nt);
    p->wU = dalloc(nt * NVP1); p->wX = dalloc(nt * NSP1); p->wXNT = dalloc(nt * NSCLR);
    p->ghat = dalloc(nt);
    p->tinc_fcorr = dalloc(n1); p->sinc_fcorr = dalloc(n1);
    p->fcorr_withz = dalloc(n1); p->sfcorr_withz = dalloc(n1);
    p->advection = dalloc((size_t)c->maxmodeadv * 2);
    p->scorr = dalloc(n1); p->ocnTcorr = dalloc(n1); p->sal_clim = dalloc(n1); p->ocnT_clim = dalloc(n1);
    p->sflux = dalloc((size_t)c->nsflxs * 5 * (c->njdt + 1));
    p->talpha = dalloc(ntt); p->sbeta = dalloc(ntt);
    p->modeadv = (int *)calloc((size_t)c->maxmodeadv * 2, sizeof(int));
    /* locals */
    p->Uo = dalloc(n1 * 2); p->Xo = dalloc(n1 * 2); p->Ux = dalloc(n1 * 2); p->Xx = dalloc(n1 * 2);
    p->dVsq = dalloc(n1); p->Ritop = dalloc(n1); p->alphaDT = dalloc(n1); p->betaDS = dalloc(n1);
    p->blmc = dalloc(((size_t)c->nz + 1) * 3);
    p->cu = dalloc(nt); p->cc = dalloc(nt); p->cl = dalloc(nt); p->rhs = dalloc(nt);
    p->diff = dalloc(nt); p->gcap = dalloc(nt); p->ntflx = dalloc(nt * NSCLR); p->gam = dalloc(nt);
    p->zw = dalloc(nt);
}

static void col1d_free(col1d *p)
{
    free(p->U); free(p->U_init); free(p->X); free(p->Rig); free(p->dbloc); free(p->Shsq);
    free(p->hmixd); free(p->Us); free(p->Xs); free(p->rho); free(p->cp); free(p->buoy);
    free(p->swfrac); free(p->swdk_opt); free(p->difm); free(p->difs); free(p->dift);
    free(p->wU); free(p->wX); free(p->wXNT); free(p->ghat);
    free(p->tinc_fcorr); free(p->sinc_fcorr); free(p->fcorr_withz); free(p->sfcorr_withz);
    free(p->advection); free(p->scorr); free(p->ocnTcorr); free(p->sal_clim); free(p->ocnT_clim);
    free(p->sflux); free(p->talpha); free(p->sbeta); free(p->modeadv);
    free(p->Uo); free(p->Xo); free(p->Ux); free(p->Xx);
    free(p->dVsq); free(p->Ritop); free(p->alphaDT); free(p->betaDS); free(p->blmc);
    free(p->cu); free(p->cc); free(p->cl); free(p->rhs); free(p->diff); free(p->gcap);
    free(p->ntflx); free(p->gam); free(p->zw);
    memset(p, 0, sizeof(*p));
}

/* Fortran intrinsics */
static inline double f_sign(double a, double b) { return copysign(fabs(a), b); }
static inline int f_int(double x) { return (int)x; } /* truncation toward zero */

/* ================================================================== */
/* Equation of state  (src/mckpp_physics_state_equations.F90)          */
/* ================================================================== */

/* MCKPP_CPSW  state_equations.F90:7-58 */
double orc_cpsw(double S, double T1, double P0)
{
    double T, P, SR, A, B, C, CP0, CP1, CP2;
    T = T1;
    if (T < -2.) T = -2.;
    P = P0 / 10.;
    SR = sqrt(fabs(S));
    A = (-1.38385E-3 * T + 0.1072763) * T - 7.643575;
    B = (5.148E-5 * T - 4.07718E-3) * T + 0.1770383;
    C = (((2.093236E-5 * T - 2.654387E-3) * T + 0.1412855) * T - 3.720283) * T + 4217.4;
    CP0 = (B * SR + A) * S + C;
    A = (((1.7168E-8 * T + 2.0357E-6) * T - 3.13885E-4) * T + 1.45747E-2) * T - 0.49592;
    B = (((2.2956E-11 * T - 4.0027E-9) * T + 2.87533E-7) * T - 1.08645E-5) * T + 2.4931E-4;
    C = ((6.136E-13 * T - 6.5637E-11) * T + 2.6380E-9) * T - 5.422E-8;
    CP1 = ((C * P + B) * P + A) * P;
    A = (((-2.9179E-10 * T + 2.5941E-8) * T + 9.802E-7) * T - 1.28315E-4) * T + 4.9247E-3;
    B = (3.122E-8 * T - 1.517E-6) * T - 1.2331E-4;
    A = (A + B * SR) * S;
    B = ((1.8448E-11 * T - 2.3905E-9) * T + 1.17054E-7) * T - 2.9558E-6;
    B = (B + 9.971E-8 * SR) * S;
    C = (3.513E-13 * T - 1.7682E-11) * T + 5.540E-10;
    C = (C - 1.4300E-12 * T * SR) * S;
    CP2 = ((C * P + B) * P + A) * P;
    return CP0 + CP1 + CP2;
}

/* the former COMMON block shared by Sig80/Bet80/Alf80/Kap80 */
typedef struct eos_common {
    double R1, R2, R3, R4, A, B, C, D, E, A1, B1, K, SR, P0, PK, Rho, Rho0, ABFac;
    int ABFlg;
} eos_common;

/* MCKPP_Sig80 state_equations.F90:371-476, with ENTRY MCKPP_BlkMod (:434).
 * entry_blkmod != 0 enters at the ENTRY statement. */
static void sig80(double S, double T, double P, int *KapFlg, double *Sig0, double *Sig,
                  eos_common *q, int entry_blkmod)
{
    double AW, BW, K0, KW;
    if (!entry_blkmod) {
        q->P0 = P / 10.0;
        q->SR = sqrt(fabs(S));
        *KapFlg = 0;
        q->R1 = ((((6.536332E-9 * T - 1.120083E-6) * T + 1.001685E-4) * T - 9.095290E-3)
                 * T + 6.793952E-2) * T - .157406;
        q->R2 = (((5.3875E-9 * T - 8.2467E-7) * T + 7.6438E-5) * T - 4.0899E-3) * T + 8.24493E-1;
        q->R3 = (-1.6546E-6 * T + 1.0227E-4) * T - 5.72466E-3;
        q->R4 = 4.8314E-4;
        *Sig0 = (q->R4 * S + q->R3 * q->SR + q->R2) * S + q->R1;
        q->Rho0 = 1000.0 + *Sig0;
        if (P == 0.0) {
            *Sig = *Sig0;
            q->Rho = q->Rho0;
            return;
        }
    }
    /* Entry MCKPP_BlkMod */
    if (*KapFlg) {
        q->P0 = P / 10.0;
        q->SR = sqrt(fabs(S));
    }
    q->B1 = (-5.3009E-4 * T + 1.6483E-2) * T + 7.944E-2;
    q->A1 = ((-6.1670E-5 * T + 1.09987E-2) * T - 0.603459) * T + 54.6746;
    KW = (((-5.155288E-5 * T + 1.360477E-2) * T - 2.327105) * T + 148.4206) * T + 19652.21;
    K0 = (q->B1 * q->SR + q->A1) * S + KW;
    if (P == 0.0) {
        q->K = K0;
        return;
    }
    q->E = (9.1697E-10 * T + 2.0816E-8) * T - 9.9348E-7;
    BW = (5.2787E-8 * T - 6.12293E-6) * T + 8.50935E-5;
    q->B = BW + q->E * S;
    q->D = 1.91075E-4;
    q->C = (-1.6078E-6 * T - 1.0981E-5) * T + 2.2838E-3;
    AW = ((-5.77905E-7 * T + 1.16092E-4) * T + 1.43713E-3) * T + 3.239908;
    q->A = (q->D * q->SR + q->C) * S + AW;
    q->K = (q->B * q->P0 + q->A) * q->P0 + K0;
    q->PK = q->P0 / q->K;
    if (*KapFlg) return;
    *Sig = (1000.0 * q->PK + *Sig0) / (1.0 - q->PK);
    q->Rho = 1000.0 + *Sig;
}

/* MCKPP_Bet80 state_equations.F90:206-240 */
static void bet80(double S, double T, double P, double *Beta, eos_common *q)
{
    double SR5, DRho, DK, DK0, DA, DB;
    (void)T;
    SR5 = q->SR * 1.5;
    DRho = q->R2 + SR5 * q->R3 + (S + S) * q->R4;
    if (P == 0) {
        *Beta = DRho / q->Rho;
        return;
    }
    DK0 = q->A1 + SR5 * q->B1;
    DA = q->C + SR5 * q->D;
    DB = q->E;
    DK = (DB * q->P0 + DA) * q->P0 + DK0;
    q->ABFac = q->Rho0 * q->P0 / ((q->K - q->P0) * (q->K - q->P0));
    q->ABFlg = 0;
    *Beta = DRho / (1. - q->PK) - q->ABFac * DK;
    *Beta = *Beta / q->Rho;
}

/* MCKPP_Alf80 state_equations.F90:244-317 */
static void alf80(double S, double T, double P, double *Alpha, eos_common *q)
{
    double AW, BW, K0, KW, Alph0, AlphaA, AlphB, AlphK;
    q->R1 = (((.3268166E-7 * T - .4480332e-5) * T + .3005055e-3) * T - .1819058E-1) * T + 6.793952E-2;
    q->R2 = ((.215500E-7 * T - .247401E-5) * T + .152876E-3) * T - 4.0899E-3;
    q->R3 = -.33092E-5 * T + 1.0227E-4;
    Alph0 = (q->R3 * q->SR + q->R2) * S + q->R1;
    if (P == 0.0) {
        *Alpha = -Alph0 / q->Rho;
        return;
    }
    q->B1 = -.106018E-2 * T + 1.6483E-2;
    q->A1 = (-.18501E-3 * T + .219974E-1) * T - 0.603459;
    KW = ((-.2062115E-3 * T + .4081431E-1) * T - .4654210E+1) * T + 148.4206;
    K0 = (q->B1 * q->SR + q->A1) * S + KW;
    q->E = .183394E-8 * T + 2.0816E-8;
    BW = .105574E-6 * T - 6.12293E-6;
    AlphB = BW + q->E * S;
    q->C = -.32156E-5 * T - 1.0981E-5;
    AW = (-.1733715E-5 * T + .232184E-3) * T + 1.43713E-3;
    AlphaA = q->C * S + AW;
    AlphK = (AlphB * q->P0 + AlphaA) * q->P0 + K0;
    if (q->ABFlg) {
        q->ABFac = q->Rho0 * q->P0 / ((q->K - q->P0) * (q->K - q->P0));
    }
    *Alpha = Alph0 / (1. - q->PK) - q->ABFac * AlphK;
    *Alpha = -*Alpha / q->Rho;
}

/* MCKPP_Kap80 state_equations.F90:336-367 (unreached by the hot path: vmix
 * passes Kappa=0; kept so the reference's kappa check values pin Sig80/BlkMod) */
static void kap80(double S, double T, double P, int *KapFlg, double *Kappa, eos_common *q)
{
    double DelK, dummy0 = 0, dummy1 = 0;
    if (P == 0) {
        *KapFlg = 1;
        sig80(S, T, P, KapFlg, &dummy0, &dummy1, q, 1);
        *Kappa = 1.0 / q->K;
        return;
    }
    if (*KapFlg) {
        sig80(S, T, P, KapFlg, &dummy0, &dummy1, q, 1);
    }
    DelK = q->A + (q->P0 + q->P0) * q->B;
    *Kappa = (1. - q->PK * DelK) / (q->K - q->P0);
}

/* MCKPP_ABK80 state_equations.F90:133-190 */
void orc_abk80(double S, double T1, double P, double *Alpha, double *Beta,
               double *Kappa, double *Sig0, double *Sig)
{
    eos_common q;
    double T;
    int KapFlg;
    memset(&q, 0, sizeof(q));
    T = T1;
    if (T < -2.) T = -2.;
    KapFlg = 1;
    q.ABFlg = 1;
    if (*Beta != 0) {
        sig80(S, T, P, &KapFlg, Sig0, Sig, &q, 0);
        bet80(S, T, P, Beta, &q);
    }
    if (*Alpha != 0) {
        if (KapFlg) {
            sig80(S, T, P, &KapFlg, Sig0, Sig, &q, 0);
        }
        alf80(S, T, P, Alpha, &q);
    }
    if (KapFlg) {
        *Sig = 0.;
        *Sig0 = 0.;
    }
    if (*Kappa != 0) {
        kap80(S, T, P, &KapFlg, Kappa, &q);
    }
}

/* ================================================================== */
/* Shortwave  (src/mckpp_physics_swfrac_mod.F90, src/mckpp_fluxes_mod.F90) */
/* ================================================================== */
static const double jw_rfac[5] = {0.58, 0.62, 0.67, 0.77, 0.78};
static const double jw_a1[5]   = {0.35, 0.6, 1.0, 1.5, 1.4};
static const double jw_a2[5]   = {23.0, 20.0, 17.0, 14.0, 7.9};

/* MCKPP_PHYSICS_SWFRAC_OPT swfrac_mod.F90:14-43 */
static void swfrac_opt(double fact, col1d *p, const orc_const *c)
{
    double rmin = -80., r1, r2;
    int l, j = p->jerlov - 1;
    for (l = 1; l <= c->nzp1; l++) {
        r1 = fmax(ZM_(c, l) * fact / jw_a1[j], rmin);
        r2 = fmax(ZM_(c, l) * fact / jw_a2[j], rmin);
        p->swfrac[l] = jw_rfac[j] * exp(r1) + (1. - jw_rfac[j]) * exp(r2);
    }
}

/* MCKPP_PHYSICS_SWFRAC swfrac_mod.F90:49-79 */
void orc_swfrac(double fact, double z, int jwtype, double *swdk)
{
    double rmin = -80., r1, r2;
    int j = jwtype - 1;
    r1 = fmax(z * fact / jw_a1[j], rmin);
    r2 = fmax(z * fact / jw_a2[j], rmin);
    *swdk = jw_rfac[j] * exp(r1) + (1. - jw_rfac[j]) * exp(r2);
}

/* mckpp_fluxes_swdk fluxes_mod.F90:121-137 */
double orc_swdk(double z, int j)
{
    return jw_rfac[j - 1] * exp(z / jw_a1[j - 1]) + (1.0 - jw_rfac[j - 1]) * exp(z / jw_a2[j - 1]);
}

/* mckpp_fluxes_ntflux fluxes_mod.F90:93-118 */
static void fluxes_ntflux(col1d *p, const orc_const *c)
{
    int k;
    if (p->ntime <= 1) {
        for (k = 0; k <= c->nz; k++)
            p->swdk_opt[k] = orc_swdk(-DM_(c, k), p->jerlov);
    }
    if (p->ntime >= 1) {
        for (k = 0; k <= c->nz; k++)
            WXNT_(p, k, 1) = -SFLUX_(p, 3, 5, 0) * p->swdk_opt[k] / (p->rho[0] * p->cp[0]);
    }
}

/* ================================================================== */
/* wscale  (src/mckpp_physics_verticalmixing_wscale_mod.F90:12-97)     */
/* ================================================================== */
void orc_wscale(const orc_const *c, double sigma, double hbl, double ustar, double bfsfc,
                double *wm, double *ws)
{
    int iz, izp1, ju, jup1;
    int ni = 890, nj = 48;
    double am, as, c1, c2, c3, cm, cs, epsln, fzfrac, ucube, udiff, ufrac,
        wam, was, wbm, wbs, zdiff, zetas, zfrac, zetam;
    double deltaz, deltau, zmin, zmax, umin, umax, zehat;

    zmin = -4.e-7; zmax = 0.0; umin = 0.0; umax = 0.04; epsln = 1.0e-20;
    c1 = 5.0; am = 1.257; cm = 8.380; c2 = 16.0; zetam = -0.2;
    as = -28.86; cs = 98.96; c3 = 16.0; zetas = -1.0;
    (void)am; (void)as; (void)c2; (void)c3; (void)cm; (void)cs; (void)epsln; (void)zetas; (void)zetam;

    deltaz = (zmax - zmin) / (ni + 1);
    deltau = (umax - umin) / (nj + 1);

    zehat = c->vonk * sigma * hbl * bfsfc;

    if (zehat <= zmax) {
        double q;
        zdiff = zehat - zmin;
        q = zdiff / deltaz;
        /* int() of an out-of-range REAL is undefined in Fortran; clamp first
         * (identical result inside the defined range) */
        if (q > 2.0e9) q = 2.0e9;
        if (q < -2.0e9) q = -2.0e9;
        iz = f_int(q);
        iz = iz < ni ? iz : ni;
        iz = iz > 0 ? iz : 0;
        izp1 = iz + 1;

        udiff = ustar - umin;
        q = udiff / deltau;
        if (q > 2.0e9) q = 2.0e9;
        if (q < -2.0e9) q = -2.0e9;
        ju = f_int(q);
        ju = ju < nj ? ju : nj;
        ju = ju > 0 ? ju : 0;
        jup1 = ju + 1;

        zfrac = zdiff / deltaz - (double)iz;
        ufrac = udiff / deltau - (double)ju;

        fzfrac = 1. - zfrac;
        wam = (fzfrac) * WMT_(c, iz, jup1) + zfrac * WMT_(c, izp1, jup1);
        wbm = (fzfrac) * WMT_(c, iz, ju) + zfrac * WMT_(c, izp1, ju);
        *wm = (1. - ufrac) * wbm + ufrac * wam;

        was = (fzfrac) * WST_(c, iz, jup1) + zfrac * WST_(c, izp1, jup1);
        wbs = (fzfrac) * WST_(c, iz, ju) + zfrac * WST_(c, izp1, ju);
        *ws = (1. - ufrac) * wbs + ufrac * was;
    } else {
        ucube = ustar * ustar * ustar;
        *wm = c->vonk * ustar * ucube / (ucube + c1 * zehat);
        *ws = *wm;
    }
}

/* ================================================================== */
/* z121  (src/mckpp_physics_verticalmixing_z121_mod.F90:7-45)          */
/* ================================================================== */
static void z121(int kmp1, double vlo, double vhi, double *V /*0:kmp1*/, double *w /*0:kmp1*/)
{
    double tmp, wait;
    int k, km;
    km = kmp1 - 1;
    w[0] = 0.0;
    w[kmp1] = 0.0;
    V[0] = 0.0;
    V[kmp1] = 0.0;
    for (k = 1; k <= km; k++) {
        if ((V[k] < vlo) || (V[k] > vhi))
            w[k] = 0.0;
        else
            w[k] = 1.0;
    }
    for (k = 1; k <= km; k++) {
        tmp = V[k];
        V[k] = w[k - 1] * V[0] + 2. * V[k] + w[k + 1] * V[k + 1];
        wait = w[k - 1] + 2.0 + w[k + 1];
        V[k] = V[k] / wait;
        V[0] = tmp;
    }
}

/* ================================================================== */
/* rimix  (src/mckpp_physics_verticalmixing_rimix_mod.F90:13-106)      */
/* ================================================================== */
static void rimix(int km, int kmp1, col1d *p, const orc_const *c)
{
    double Rigg, fri, fcon, ratio;
    double epsln, Riinfty, Ricon, difm0, difs0, difmiw, difsiw, difmcon, difscon, c1, c0;
    int ki, mRi, j;
    epsln = 1.e-16; Riinfty = 0.8; Ricon = -0.2; difm0 = 0.005; difs0 = 0.005;
    difmiw = 0.0001; difsiw = 0.00001; difmcon = 0.0000; difscon = 0.0000;
    c1 = 1.0; c0 = 0.0; mRi = 1;

    for (ki = 1; ki <= km; ki++) {
        p->Rig[ki] = p->dbloc[ki] * (ZM_(c, ki) - ZM_(c, ki + 1)) / (p->Shsq[ki] + epsln);
        p->dift[ki] = p->Rig[ki];
        p->difm[ki] = p->dift[ki];
    }
    for (j = 1; j <= mRi; j++)
        z121(kmp1, c0, Riinfty, p->difm, p->difs);

    for (ki = 1; ki <= km; ki++) {
        Rigg = fmax(p->dift[ki], Ricon);
        ratio = fmin((Ricon - Rigg) / Ricon, c1);
        fcon = (c1 - ratio * ratio);
        fcon = fcon * fcon * fcon;

        Rigg = fmax(p->difm[ki], c0);
        ratio = fmin(Rigg / Riinfty, c1);
        fri = (c1 - ratio * ratio);
        fri = fri * fri * fri;

        p->difm[ki] = (difmiw + fcon * difmcon + fri * difm0);
        p->difs[ki] = (difsiw + fcon * difscon + fri * difs0);
        p->dift[ki] = p->difs[ki];
    }
    p->difm[0] = c0;
    p->dift[0] = c0;
    p->difs[0] = c0;
}

/* ================================================================== */
/* ddmix  (src/mckpp_physics_verticalmixing_ddmix_mod.F90:12-52)       */
/* ================================================================== */
static void ddmix(int km, const double *alphaDT, const double *betaDS, col1d *p)
{
    int ki;
    double dsfmax, Rrho0, Rrho, diffdd, prandtl, t;
    Rrho0 = 1.9;
    dsfmax = 1.0e-4;
    for (ki = 1; ki <= km; ki++) {
        if ((alphaDT[ki] > betaDS[ki]) && (betaDS[ki] > 0.)) {
            Rrho = fmin(alphaDT[ki] / betaDS[ki], Rrho0);
            t = ((Rrho - 1) / (Rrho0 - 1));
            diffdd = 1.0 - t * t;
            diffdd = dsfmax * diffdd * diffdd * diffdd;
            p->dift[ki] = p->dift[ki] + diffdd * 0.8 / Rrho;
            p->difs[ki] = p->difs[ki] + diffdd;
        } else if ((alphaDT[ki] < 0.0) && (betaDS[ki] < 0.0) && (alphaDT[ki] < betaDS[ki])) {
            Rrho = alphaDT[ki] / betaDS[ki];
            diffdd = 1.5e-6 * 9.0 * 0.101 * exp(4.6 * exp(-0.54 * (1 / Rrho - 1)));
            prandtl = 0.15 * Rrho;
            if (Rrho > 0.5) prandtl = (1.85 - 0.85 / Rrho) * Rrho;
            p->dift[ki] = p->dift[ki] + diffdd;
            p->difs[ki] = p->difs[ki] + prandtl * diffdd;
        }
    }
}

/* ================================================================== */
/* bldepth  (src/mckpp_physics_verticalmixing_bldepth_mod.F90:32-203)  */
/* ================================================================== */
static void bldepth(int km, int kmp1, const double *dVsq, const double *Ritop, double ustar,
                    double Bo, double Bosol, double *hbl, double *bfsfc, double *stable,
                    double *caseA, int *kbl, double *Rib /*1:2*/, double *sigma, double *wm,
                    double *ws, col1d *p, const orc_const *c)
{
    double bvsq, cekman, cmonob, cs, cv, epsilon, fekman, fmonob, hbf, hekman, hmin, hmin2,
        hmonob, hri, Ricr, Vtc, Vtsq, epsln;
    int ka, ksave, ku, kl;
    double dmo[3], hek;

    epsln = 1.e-16; Ricr = 0.30; epsilon = 0.1; cekman = 0.7; cmonob = 1.0;
    cs = 98.96; cv = 1.6; hbf = 1.0;

    Vtc = cv * sqrt(0.2 / cs / epsilon) / (c->vonk * c->vonk) / Ricr;

    ka = 1;
    ku = 2;

    Rib[ka] = 0.0;
    dmo[ka] = -ZM_(c, kmp1);
    *kbl = km;
    *hbl = -ZM_(c, km);
    hek = cekman * ustar / (fabs(p->f) + epsln);

    for (kl = 2; kl <= km; kl++) {
        if (p->ntime <= 1 && kl == 2) {
            swfrac_opt(hbf, p, c);
        }
        if (*kbl >= km) {
            *caseA = -ZM_(c, kl);
            *bfsfc = Bo + Bosol * (1. - p->swfrac[kl]);
            *stable = 0.5 + f_sign(0.5, *bfsfc + epsln);
            *sigma = *stable * 1. + (1. - *stable) * epsilon;
        }
        orc_wscale(c, *sigma, *caseA, ustar, *bfsfc, wm, ws);

        if (*kbl >= km) {
            bvsq = 0.5 * (p->dbloc[kl - 1] / (ZM_(c, kl - 1) - ZM_(c, kl)) +
                          p->dbloc[kl] / (ZM_(c, kl) - ZM_(c, kl + 1)));
            Vtsq = -ZM_(c, kl) * *ws * sqrt(fabs(bvsq)) * Vtc;
            Rib[ku] = Ritop[kl] / (dVsq[kl] + Vtsq + epsln);
            Rib[ku] = fmax(Rib[ku], Rib[ka] + epsln);
            hri = -ZM_(c, kl - 1) + (ZM_(c, kl - 1) - ZM_(c, kl)) *
                                        (Ricr - Rib[ka]) / (Rib[ku] - Rib[ka]);

            fmonob = *stable * 1.0;
            dmo[ku] = cmonob * ustar * ustar * ustar / c->vonk / (fabs(*bfsfc) + epsln);
            dmo[ku] = fmonob * dmo[ku] - (1. - fmonob) * ZM_(c, kmp1);
            if (dmo[ku] <= (-ZM_(c, kl))) {
                hmonob = (dmo[ku] - dmo[ka]) / (ZM_(c, kl - 1) - ZM_(c, kl));
                hmonob = (dmo[ku] + hmonob * ZM_(c, kl)) / (1. - hmonob);
            } else {
                hmonob = -ZM_(c, kmp1);
            }

            fekman = *stable * 1.0;
            hekman = fekman * hek - (1. - fekman) * ZM_(c, kmp1);

            hmin = fmin(fmin(fmin(hri, hmonob), hekman), -p->ocdepth);
            if (hmin < -ZM_(c, kl)) {
                if (!p->l_initflag) {
                    if (hmin < -ZM_(c, kl - 1)) {
                        hmin2 = fmin(fmin(hri, hmonob), -p->ocdepth);
                        if (hmin2 < -ZM_(c, kl)) {
                            hmin = hmin2;
                        }
                    }
                }
                *hbl = hmin;
                *kbl = kl;
            }
        }
        ksave = ka;
        ka = ku;
        ku = ksave;
    }

    orc_swfrac(-1.0, *hbl, p->jerlov, bfsfc);

    *bfsfc = Bo + Bosol * (1. - *bfsfc);
    *stable = 0.5 + f_sign(0.5, *bfsfc);
    *bfsfc = *bfsfc + *stable * epsln;

    *caseA = 0.5 + f_sign(0.5, -ZM_(c, *kbl) - 0.5 * HM_(c, *kbl) - *hbl);
}

/* ================================================================== */
/* blmix  (src/mckpp_physics_verticalmixing_blmix_mod.F90:13-151)      */
/* ================================================================== */
static void blmix(int km, double ustar, double bfsfc, double hbl, double stable, double caseA,
                  int kbl, double *gat1, double *dat1, double *dkm1, double *sigma, double *wm,
                  double *ws, col1d *p, const orc_const *c)
{
    double a1, a2, a3, c1, cg, cs, cstar, delhat, difsh, difsp, difth, diftp, dvdzup, epsln, f1,
        Gm, Gs, dvdzdn, epsilon, Gt, R, visch, viscp, sig;
    int ki, kn;

    epsln = 1.e-20; epsilon = 0.1; c1 = 5.0; cs = 98.96; cstar = 5.0;

    cg = cstar * c->vonk * pow(cs * c->vonk * epsilon, 1. / 3.);

    *sigma = stable * 1.0 + (1. - stable) * epsilon;

    orc_wscale(c, *sigma, hbl, ustar, bfsfc, wm, ws);
    kn = f_int(caseA + epsln) * (kbl - 1) + (1 - f_int(caseA + epsln)) * kbl;

    delhat = 0.5 * HM_(c, kn) - ZM_(c, kn) - hbl;
    R = 1.0 - delhat / HM_(c, kn);
    dvdzup = (p->difm[kn - 1] - p->difm[kn]) / HM_(c, kn);
    dvdzdn = (p->difm[kn] - p->difm[kn + 1]) / HM_(c, kn + 1);
    viscp = 0.5 * ((1. - R) * (dvdzup + fabs(dvdzup)) + R * (dvdzdn + fabs(dvdzdn)));

    dvdzup = (p->difs[kn - 1] - p->difs[kn]) / HM_(c, kn);
    dvdzdn = (p->difs[kn] - p->difs[kn + 1]) / HM_(c, kn + 1);
    difsp = 0.5 * ((1. - R) * (dvdzup + fabs(dvdzup)) + R * (dvdzdn + fabs(dvdzdn)));

    dvdzup = (p->dift[kn - 1] - p->dift[kn]) / HM_(c, kn);
    dvdzdn = (p->dift[kn] - p->dift[kn + 1]) / HM_(c, kn + 1);
    diftp = 0.5 * ((1. - R) * (dvdzup + fabs(dvdzup)) + R * (dvdzdn + fabs(dvdzdn)));

    visch = p->difm[kn] + viscp * delhat;
    difsh = p->difs[kn] + difsp * delhat;
    difth = p->dift[kn] + diftp * delhat;

    f1 = stable * c1 * bfsfc / ((ustar * ustar) * (ustar * ustar) + epsln);
    gat1[1] = visch / hbl / (*wm + epsln);
    dat1[1] = -viscp / (*wm + epsln) + f1 * visch;
    dat1[1] = fmin(dat1[1], 0.);

    gat1[2] = difsh / hbl / (*ws + epsln);
    dat1[2] = -difsp / (*ws + epsln) + f1 * difsh;
    dat1[2] = fmin(dat1[2], 0.);

    gat1[3] = difth / hbl / (*ws + epsln);
    dat1[3] = -diftp / (*ws + epsln) + f1 * difth;
    dat1[3] = fmin(dat1[3], 0.);

    for (ki = 1; ki <= km; ki++) {
        sig = (-ZM_(c, ki) + 0.5 * HM_(c, ki)) / hbl;
        *sigma = stable * sig + (1. - stable) * fmin(sig, epsilon);
        orc_wscale(c, *sigma, hbl, ustar, bfsfc, wm, ws);

        sig = (-ZM_(c, ki) + 0.5 * HM_(c, ki)) / hbl;
        a1 = sig - 2.;
        a2 = 3. - 2. * sig;
        a3 = sig - 1.;

        Gm = a1 + a2 * gat1[1] + a3 * dat1[1];
        Gs = a1 + a2 * gat1[2] + a3 * dat1[2];
        Gt = a1 + a2 * gat1[3] + a3 * dat1[3];

        BLMC_(p, ki, 1) = hbl * *wm * sig * (1. + sig * Gm);
        BLMC_(p, ki, 2) = hbl * *ws * sig * (1. + sig * Gs);
        BLMC_(p, ki, 3) = hbl * *ws * sig * (1. + sig * Gt);

        p->ghat[ki] = (1. - stable) * cg / (*ws * hbl + epsln);
    }

    sig = -ZM_(c, kbl - 1) / hbl;
    *sigma = stable * sig + (1. - stable) * fmin(sig, epsilon);

    orc_wscale(c, *sigma, hbl, ustar, bfsfc, wm, ws);
    sig = -ZM_(c, kbl - 1) / hbl;
    a1 = sig - 2.;
    a2 = 3. - 2. * sig;
    a3 = sig - 1.;
    Gm = a1 + a2 * gat1[1] + a3 * dat1[1];
    Gs = a1 + a2 * gat1[2] + a3 * dat1[2];
    Gt = a1 + a2 * gat1[3] + a3 * dat1[3];
    dkm1[1] = hbl * *wm * sig * (1. + sig * Gm);
    dkm1[2] = hbl * *ws * sig * (1. + sig * Gs);
    dkm1[3] = hbl * *ws * sig * (1. + sig * Gt);
}

/* ================================================================== */
/* enhance  (src/mckpp_physics_verticalmixing_enhance_mod.F90:10-51)   */
/* ================================================================== */
static void enhance(int km, const double *dkm1, double hbl, int kbl, double caseA, col1d *p,
                    const orc_const *c)
{
    double dkmp5, dstar, delta;
    int ki;
    for (ki = 1; ki <= km - 1; ki++) {
        if (ki == (kbl - 1)) {
            delta = (hbl + ZM_(c, ki)) / (ZM_(c, ki) - ZM_(c, ki + 1));

            dkmp5 = caseA * p->difm[ki] + (1. - caseA) * BLMC_(p, ki, 1);
            dstar = ((1. - delta) * (1. - delta)) * dkm1[1] + (delta * delta) * dkmp5;
            BLMC_(p, ki, 1) = (1. - delta) * p->difm[ki] + delta * dstar;

            dkmp5 = caseA * p->difs[ki] + (1. - caseA) * BLMC_(p, ki, 2);
            dstar = ((1. - delta) * (1. - delta)) * dkm1[2] + (delta * delta) * dkmp5;
            BLMC_(p, ki, 2) = (1. - delta) * p->difs[ki] + delta * dstar;

            dkmp5 = caseA * p->dift[ki] + (1. - caseA) * BLMC_(p, ki, 3);
            dstar = ((1. - delta) * (1. - delta)) * dkm1[3] + (delta * delta) * dkmp5;
            BLMC_(p, ki, 3) = (1. - delta) * p->dift[ki] + delta * dstar;

            p->ghat[ki] = (1. - caseA) * p->ghat[ki];
        }
    }
}

/* ================================================================== */
/* kppmix  (src/mckpp_physics_verticalmixing_kppmix_mod.F90:25-126)    */
/* ================================================================== */
static void kppmix(int km, int kmp1, const double *dVsq, double ustar, double Bo, double Bosol,
                   const double *alphaDT, const double *betaDS, const double *Ritop, double *hbl,
                   int *kbl, col1d *p, const orc_const *c)
{
    int ki;
    double bfsfc = 0, ws = 0, wm = 0, caseA = 0, stable = 0, sigma = 0;
    double dkm1[4], gat1[4], dat1[4], Rib[3];

    for (ki = 0; ki <= km; ki++) {
        p->difm[ki] = 0.0;
        p->difs[ki] = 0.0;
        p->dift[ki] = 0.0;
    }
    if (c->LRI) rimix(km, kmp1, p, c);
    if (c->LDD) ddmix(km, alphaDT, betaDS, p);

    p->difm[kmp1] = p->difm[km];
    p->difs[kmp1] = p->difs[km];
    p->dift[kmp1] = p->dift[km];

    if (c->LKPP) {
        bldepth(km, kmp1, dVsq, Ritop, ustar, Bo, Bosol, hbl, &bfsfc, &stable, &caseA, kbl, Rib,
                &sigma, &wm, &ws, p, c);
        p->dbg_ustar = ustar; p->dbg_Bo = Bo; p->dbg_Bosol = Bosol; p->dbg_hbl = *hbl; p->dbg_bfsfc = bfsfc;
        p->dbg_stable = stable; p->dbg_caseA = caseA; p->dbg_kbl = *kbl;
        blmix(km, ustar, bfsfc, *hbl, stable, caseA, *kbl, gat1, dat1, dkm1, &sigma, &wm, &ws, p, c);
        enhance(km, dkm1, *hbl, *kbl, caseA, p, c);
        for (ki = 1; ki <= km; ki++) {
            if (ki < *kbl) {
                p->difm[ki] = BLMC_(p, ki, 1);
                p->difs[ki] = BLMC_(p, ki, 2);
                p->dift[ki] = BLMC_(p, ki, 3);
            } else {
                p->ghat[ki] = 0.;
            }
        }
    }
}

/* ================================================================== */
/* vmix  (src/mckpp_physics_verticalmixing_mod.F90:14-161)             */
/* ================================================================== */
static void verticalmixing(col1d *p, const orc_const *c, double *hmixn, int *kmixn)
{
    double B0, B0sol, ustar, rhob;
    double *dVsq = p->dVsq, *Ritop = p->Ritop, *alphaDT = p->alphaDT, *betaDS = p->betaDS;
    double epsilon, alpha, beta, exppr, sigma, sigma0, tau, zref, wz, bref, del, dlimit, vlimit;
    int k, n, kl, nz = c->nz, nzp1 = c->nzp1;

    epsilon = 0.1;

    alpha = 1.;
    beta = 1.;
    exppr = 0.0;
    sigma0 = 0;
    sigma = 0;
    orc_abk80(0.0, X_(p, 1, 1), -ZM_(c, 1), &alpha, &beta, &exppr, &sigma0, &sigma);
    p->rhoh2o = 1000. + sigma0;
    orc_abk80(c->sice, X_(p, 1, 1), -ZM_(c, 1), &alpha, &beta, &exppr, &sigma0, &sigma);
    rhob = 1000. + sigma0;

    for (k = 1; k <= nzp1; k++) {
        orc_abk80(X_(p, k, 2) + p->Sref, X_(p, k, 1), -ZM_(c, k), &alpha, &beta, &exppr, &sigma0, &sigma);
        p->rho[k] = 1000. + sigma0;
        p->cp[k] = orc_cpsw(X_(p, k, 2) + p->Sref, X_(p, k, 1), -ZM_(c, k));
        p->talpha[k] = alpha;
        p->sbeta[k] = beta;
        p->buoy[k] = -c->grav * sigma0 / 1000.;
    }
    p->rho[0] = p->rho[1];
    p->cp[0] = p->cp[1];
    p->talpha[0] = p->talpha[1];
    p->sbeta[0] = p->sbeta[1];

    fluxes_ntflux(p, c);

    WU_(p, 0, 1) = -SFLUX_(p, 1, 5, 0) / p->rho[0];
    WU_(p, 0, 2) = -SFLUX_(p, 2, 5, 0) / p->rho[0];
    tau = sqrt(SFLUX_(p, 1, 5, 0) * SFLUX_(p, 1, 5, 0) + SFLUX_(p, 2, 5, 0) * SFLUX_(p, 2, 5, 0)) + 1.e-16;
    ustar = sqrt(tau / p->rho[0]);

    WX_(p, 0, 1) = -SFLUX_(p, 4, 5, 0) / p->rho[0] / p->cp[0];

    WX_(p, 0, 2) = p->Ssurf * SFLUX_(p, 6, 5, 0) / p->rhoh2o +
                   (p->Ssurf - c->sice) * SFLUX_(p, 5, 5, 0) / rhob;

    B0 = -c->grav * (p->talpha[0] * WX_(p, 0, 1) - p->sbeta[0] * WX_(p, 0, 2));
    WX_(p, 0, NSP1) = -B0;
    B0sol = c->grav * p->talpha[0] * SFLUX_(p, 3, 5, 0) / (p->rho[0] * p->cp[0]);

    for (n = 1; n <= nz; n++) {
        alphaDT[n] = 0.5 * (p->talpha[n] + p->talpha[n + 1]) * (X_(p, n, 1) - X_(p, n + 1, 1));
        betaDS[n] = 0.5 * (p->sbeta[n] + p->sbeta[n + 1]) * (X_(p, n, 2) - X_(p, n + 1, 2));
    }

    for (n = 1; n <= nz; n++) {
        zref = epsilon * ZM_(c, n);
        wz = fmax(ZM_(c, 1), zref);
        p->uref = U_(p, 1, 1) * wz / zref;
        p->vref = U_(p, 1, 2) * wz / zref;
        bref = p->buoy[1] * wz / zref;
        for (kl = 1; kl <= nz; kl++) {
            if (zref >= ZM_(c, kl)) break; /* go to 126 */
            wz = fmin(ZM_(c, kl) - ZM_(c, kl + 1), ZM_(c, kl) - zref);
            del = 0.5 * wz / (ZM_(c, kl) - ZM_(c, kl + 1));
            p->uref = p->uref - wz * (U_(p, kl, 1) + del * (U_(p, kl + 1, 1) - U_(p, kl, 1))) / zref;
            p->vref = p->vref - wz * (U_(p, kl, 2) + del * (U_(p, kl + 1, 2) - U_(p, kl, 2))) / zref;
            bref = bref - wz * (p->buoy[kl] + del * (p->buoy[kl + 1] - p->buoy[kl])) / zref;
        }
        Ritop[n] = (zref - ZM_(c, n)) * (bref - p->buoy[n]);
        p->dbloc[n] = p->buoy[n] - p->buoy[n + 1];
        dVsq[n] = (p->uref - U_(p, n, 1)) * (p->uref - U_(p, n, 1)) +
                  (p->vref - U_(p, n, 2)) * (p->vref - U_(p, n, 2));
        p->Shsq[n] = (U_(p, n, 1) - U_(p, n + 1, 1)) * (U_(p, n, 1) - U_(p, n + 1, 1)) +
                     (U_(p, n, 2) - U_(p, n + 1, 2)) * (U_(p, n, 2) - U_(p, n + 1, 2));
    }

    kppmix(nz, nzp1, dVsq, ustar, B0, B0sol, alphaDT, betaDS, Ritop, hmixn, kmixn, p, c);

    dlimit = 0.00001;
    vlimit = 0.0001;
    for (k = nz; k <= nzp1; k++) {
        p->difm[k] = vlimit;
        p->difs[k] = dlimit;
        p->dift[k] = dlimit;
    }
    p->ghat[nz] = 0.0;
}

/* ================================================================== */
/* solvers  (src/mckpp_physics_solvers.F90)                            */
/* ================================================================== */

/* tridcof solvers.F90:14-44 (ind = 1) */
static void tridcof(const double *diff /*0:nzi*/, int nzi, double *cu, double *cc, double *cl,
                    const orc_const *c)
{
    int i;
    cu[1] = 0.;
    cc[1] = 1. + TRI_(c, 1, 1) * diff[1];
    cl[1] = -TRI_(c, 1, 1) * diff[1];
    for (i = 2; i <= nzi; i++) {
        cu[i] = -TRI_(c, i, 0) * diff[i - 1];
        cc[i] = 1. + TRI_(c, i, 1) * diff[i] + TRI_(c, i, 0) * diff[i - 1];
        cl[i] = -TRI_(c, i, 1) * diff[i];
    }
    cl[nzi] = 0.;
}

/* tridrhs solvers.F90:53-107.  h, yo: Fortran index 1..nzi+1 passed as base pointers
 * such that h[i], yo[i] are element i. */
static void tridrhs(int npd, const double *h, const double *yo, const double *ntflux /*0:nzi*/,
                    const double *diff /*0:nzi*/, const double *ghat /*1:nzi*/, double sturflux,
                    double ghatflux, double dto, int nzi, double *rhs, const orc_const *c)
{
    int i;
    double divflx;
    divflx = 1.0 / (double)npd;

    rhs[1] = yo[1] + dto / h[1] * (ghatflux * diff[1] * ghat[1] - sturflux * divflx + ntflux[1] - ntflux[0]);

    if (npd >= 2) {
        for (i = 2; i <= npd; i++) {
            rhs[i] = yo[i] + dto / h[i] * (ghatflux * diff[i] * ghat[i] - ghatflux * diff[i - 1] * ghat[i - 1]
                                           - sturflux * divflx + ntflux[i] - ntflux[i - 1]);
        }
    }
    for (i = npd + 1; i <= nzi - 1; i++) {
        rhs[i] = yo[i] + dto / h[i] * (ghatflux * (diff[i] * ghat[i] - diff[i - 1] * ghat[i - 1])
                                       + ntflux[i] - ntflux[i - 1]);
    }
    if (nzi > 1) {
        i = nzi;
        rhs[i] = yo[i] + dto / h[i] * (ghatflux * (diff[i] * ghat[i] - diff[i - 1] * ghat[i - 1])
                                       + ntflux[i] - ntflux[i - 1])
                 + yo[i + 1] * TRI_(c, i, 1) * diff[i];
    }
}

/* tridmat solvers.F90:112-161 (arrays addressed by Fortran index).
 * The reference aborts (STOP) on a zero pivot; the oracle flags it and
 * continues with the statement that follows the abort call (bet=1.E-12). */
static void tridmat_(const double *cu, const double *cc, const double *cl, const double *rhs,
                     const double *yo, int nzi, double *yn, double *gam, int *pivot_zero)
{
    double bet;
    int i;
    bet = cc[1];
    yn[1] = rhs[1] / bet;
    for (i = 2; i <= nzi; i++) {
        gam[i] = cl[i - 1] / bet;
        bet = cc[i] - cu[i] * gam[i];
        if (bet == 0.) {
            *pivot_zero = 1;
            bet = 1.E-12;
        }
        yn[i] = (rhs[i] - cu[i] * yn[i - 1]) / bet;
    }
    for (i = nzi - 1; i >= 1; i--) {
        yn[i] = yn[i] - gam[i + 1] * yn[i + 1];
    }
    yn[nzi + 1] = yo[nzi + 1];
}

/* public wrapper with 0-based C arrays (element i of the Fortran array at [i-1]) */
void orc_tridmat(const double *cu, const double *cc, const double *cl, const double *rhs,
                 const double *yo, int nzi, double *yn, int nztmax, int *pivot_zero)
{
    double *gam = dalloc((size_t)nztmax + 2);
    *pivot_zero = 0;
    tridmat_(cu - 1, cc - 1, cl - 1, rhs - 1, yo - 1, nzi, yn - 1, gam, pivot_zero);
    free(gam);
}

/* rhsmod solvers.F90:176-335 */
static void rhsmod(int jsclr, int mode, double A, double dto, int km, double dm, int nzi,
                   double *rhs, col1d *p, const orc_const *c)
{
    double am, fact = 0, delta, depth, dmax;
    int n, n1, n2, nzend;

    if (mode <= 0) return;
    am = A;

    if (mode == 1) {
        if (jsclr == 1) fact = dto * am / (p->rho[1] * p->cp[1]);
        if (jsclr == 2) fact = dto * am * 0.033;
        rhs[1] = rhs[1] + fact / HM_(c, 1);
    } else if (mode == 2) {
        delta = 0.0;
        for (n = 1; n <= km - 1; n++) delta = delta + HM_(c, n);
        for (n = 1; n <= km - 1; n++) {
            if (jsclr == 1) fact = dto * am / (p->rho[n] * p->cp[n]);
            if (jsclr == 2) fact = dto * am * 0.033;
            rhs[n] = rhs[n] + fact / delta;
        }
    } else if (mode == 3) {
        delta = 0.0;
        for (n = 1; n <= nzi; n++) delta = delta + HM_(c, n);
        for (n = 1; n <= nzi; n++) {
            if (jsclr == 1) fact = dto * am / (p->rho[n] * p->cp[n]);
            if (jsclr == 2) fact = dto * am * 0.033;
            rhs[n] = rhs[n] + fact / delta;
        }
    } else if (mode == 4) {
        nzend = nzi - 1;
        n1 = 0;
        do {
            n1 = n1 + 1;
        } while (n1 < c->nzp1 && ZM_(c, n1) >= -100.); /* 401 loop; bounded (reference runs off the array) */
        delta = 0.0;
        for (n = n1; n <= nzend; n++) delta = delta + HM_(c, n);
        for (n = n1; n <= nzend; n++) {
            if (jsclr == 1) fact = dto * am / (p->rho[n] * p->cp[n]);
            if (jsclr == 2) fact = dto * am * 0.033;
            rhs[n] = rhs[n] + fact / delta;
        }
    } else if (mode == 5) {
        if (jsclr == 1) fact = dto * am / (p->rho[nzi] * p->cp[nzi]);
        if (jsclr == 2) fact = dto * am * 0.033;
        rhs[nzi] = rhs[nzi] + fact / HM_(c, nzi);
    } else {
        n1 = 1; n2 = 0; delta = 0.0;
        if (mode == 6) {
            n1 = 1;
            depth = HM_(c, 1);
            dmax = dm - 0.5 * (HM_(c, km) + HM_(c, km - 1));
            delta = 0.0;
            for (n = n1; n <= nzi; n++) {
                n2 = n;
                delta = delta + HM_(c, n);
                depth = depth + HM_(c, n + 1);
                if (depth >= dmax) break;
            }
        } else if (mode == 7) {
            n1 = km - 1;
            depth = dm - 0.5 * HM_(c, km);
            dmax = 100.;
            delta = 0.0;
            for (n = n1; n <= nzi; n++) {
                n2 = n;
                delta = delta + HM_(c, n);
                depth = depth + HM_(c, n + 1);
                if (depth >= dmax) break;
            }
        } else {
            return; /* 'mode out of range' -> MCKPP_ABORT in the reference; inputs are validated upstream */
        }
        for (n = n1; n <= n2; n++) {
            if (jsclr == 1) fact = dto * am / (p->rho[n] * p->cp[n]);
            if (jsclr == 2) fact = dto * am * 0.033;
            rhs[n] = rhs[n] + fact / delta;
        }
    }
}

/* ================================================================== */
/* ocnint  (src/mckpp_physics_ocnint_mod.F90:19-221)                   */
/* ================================================================== */
static void ocnint(col1d *p, const orc_const *c, int intri, int kmixe)
{
    double *cu = p->cu, *cc = p->cc, *cl = p->cl, *rhs = p->rhs, *diff = p->diff, *gcap = p->gcap;
    double *ntflx = p->ntflx; /* ntflx(0:nztmax, nsclr) */
    int i, npd, imode, n, k, NZ = c->nz, NZP1 = c->nzp1, NZtmax = c->nztmax;
    double ftemp, ghatflux, sturflux, adv_mag;
    int adv_mode, pz = 0;
    const double *hm1 = c->hm - 1; /* hm1[i] == hm(i) */
    (void)intri;
#define NTFLX_(k, n) ntflx[((n)-1) * (NZtmax + 1) + (k)]

    ftemp = p->f;

    for (k = 0; k <= NZtmax; k++) diff[k] = p->difm[k];
    tridcof(diff, NZ, cu, cc, cl, c);

    rhs[1] = UO_(p, 1, 1) + c->dto * (ftemp * .5 * (UO_(p, 1, 2) + U_(p, 1, 2)) - WU_(p, 0, 1) / HM_(c, 1));
    for (i = 2; i <= NZ - 1; i++)
        rhs[i] = UO_(p, i, 1) + c->dto * ftemp * .5 * (UO_(p, i, 2) + U_(p, i, 2));
    i = NZ;
    rhs[i] = UO_(p, i, 1) + c->dto * ftemp * .5 * (UO_(p, i, 2) + U_(p, i, 2)) +
             TRI_(c, i, 1) * p->difm[i] * UO_(p, i + 1, 1);
    tridmat_(cu, cc, cl, rhs, &UO_(p, 0, 1), NZ, &U_(p, 0, 1), p->gam, &pz);

    rhs[1] = UO_(p, 1, 2) - c->dto * (ftemp * .5 * (UO_(p, 1, 1) + U_(p, 1, 1)) + WU_(p, 0, 2) / HM_(c, 1));
    for (i = 2; i <= NZ - 1; i++)
        rhs[i] = UO_(p, i, 2) - c->dto * ftemp * .5 * (UO_(p, i, 1) + U_(p, i, 1));
    i = NZ;
    rhs[i] = UO_(p, i, 2) - c->dto * ftemp * .5 * (UO_(p, i, 1) + U_(p, i, 1)) +
             TRI_(c, i, 1) * p->difm[i] * UO_(p, i + 1, 2);
    npd = 1;
    tridmat_(cu, cc, cl, rhs, &UO_(p, 0, 2), NZ, &U_(p, 0, 2), p->gam, &pz);

    ghatflux = WX_(p, 0, 1);
    sturflux = WX_(p, 0, 1);
    diff[0] = p->dift[0];
    NTFLX_(0, 1) = WXNT_(p, 0, 1);
    for (k = 1; k <= NZtmax; k++) {
        diff[k] = p->dift[k];
        gcap[k] = p->ghat[k];
        NTFLX_(k, 1) = WXNT_(p, k, 1);
    }
    tridcof(diff, NZ, cu, cc, cl, c);
    tridrhs(npd, hm1, &XO_(p, 0, 1), &NTFLX_(0, 1), diff, gcap, sturflux, ghatflux, c->dto, NZ, rhs, c);

    if (c->L_RELAX_SST && !c->L_FCORR_WITHZ && !c->L_FCORR) {
        if (p->relax_sst > 1.e-10) {
            if (!c->L_RELAX_CALCONLY) {
                rhs[1] = rhs[1] + c->dto * p->relax_sst * (p->SST0 - XO_(p, 1, 1)) * DM_(c, kmixe) / HM_(c, 1);
            }
            p->fcorr = p->relax_sst * (p->SST0 - XO_(p, 1, 1)) * DM_(c, kmixe) * p->rho[1] * p->cp[1];
        } else {
            p->fcorr = 0.0;
        }
    }

    if (c->L_FCORR && !c->L_RELAX_SST && !c->L_FCORR_WITHZ) {
        rhs[1] = rhs[1] + c->dto * p->fcorr_twod / (p->rho[1] * p->cp[1] * HM_(c, 1));
    }

    for (k = 1; k <= NZP1; k++) p->tinc_fcorr[k] = 0.;
    if (c->L_FCORR_WITHZ && !c->L_FCORR) {
        for (k = 1; k <= NZP1; k++)
            p->tinc_fcorr[k] = c->dto * p->fcorr_withz[k] / (p->rho[k] * p->cp[k]);
    }
    if (c->L_RELAX_OCNT) {
        for (k = 1; k <= NZP1; k++)
            p->tinc_fcorr[k] = p->tinc_fcorr[k] + c->dto * p->relax_ocnT * (p->ocnT_clim[k] - XO_(p, k, 1));
    }
    for (k = 1; k <= NZP1; k++) {
        rhs[k] = rhs[k] + p->tinc_fcorr[k];
        p->ocnTcorr[k] = p->tinc_fcorr[k] * p->rho[k] * p->cp[k] / c->dto;
    }

    tridmat_(cu, cc, cl, rhs, &XO_(p, 0, 1), NZ, &X_(p, 0, 1), p->gam, &pz);

    for (k = 0; k <= NZtmax; k++) diff[k] = p->difs[k];
    tridcof(diff, NZ, cu, cc, cl, c);
    for (n = 2; n <= NSCLR; n++) {
        for (k = 0; k <= NZtmax; k++) NTFLX_(k, n) = WXNT_(p, k, n);
        ghatflux = WX_(p, 0, n);
        sturflux = WX_(p, 0, n);
        tridrhs(npd, hm1, &XO_(p, 0, n), &NTFLX_(0, n), diff, gcap, sturflux, ghatflux, c->dto, NZ, rhs, c);

        for (imode = 1; imode <= p->nmodeadv[2]; imode++) {
            adv_mode = MODEADV_(p, imode, 2);
            adv_mag = ADVEC_(p, imode, 2);
            rhsmod(2, adv_mode, adv_mag, c->dto, kmixe, DM_(c, kmixe), NZ, rhs, p, c);
        }

        if (n == 2) {
            for (k = 1; k <= NZP1; k++) p->sinc_fcorr[k] = 0.;
            if (c->L_SFCORR_WITHZ && !c->L_SFCORR) {
                for (k = 1; k <= NZP1; k++) p->sinc_fcorr[k] = c->dto * p->sfcorr_withz[k];
            }
            if (c->L_RELAX_SAL) {
                for (k = 1; k <= NZP1; k++)
                    p->sinc_fcorr[k] = p->sinc_fcorr[k] + c->dto * p->relax_sal * (p->sal_clim[k] - XO_(p, k, n));
            }
            for (k = 1; k <= NZP1; k++) {
                rhs[k] = rhs[k] + p->sinc_fcorr[k];
                p->scorr[k] = p->sinc_fcorr[k] / c->dto;
            }
        }
        tridmat_(cu, cc, cl, rhs, &XO_(p, 0, n), NZ, &X_(p, 0, n), p->gam, &pz);
    }
    if (pz) p->status |= ORC_ST_PIVOT_ZERO;
#undef NTFLX_
}

/* ================================================================== */
/* ocnstep  (src/mckpp_physics_ocnstep_mod.F90:43-357)                 */
/* ================================================================== */
static void ocnstep(col1d *p, const orc_const *c)
{
    double hmixe = 0, hmixn = 0, tol;
    double Ui, dampU[3], lambda;
    int iter, iconv, kmixe = 0, kmixn = 0;
    double deltaz, a, b;
    int k, l, n, NZ = c->nz, NZP1 = c->nzp1;
    int comp_iter_max;
    double rmsd[5], rmsd_threshold[5];

    comp_iter_max = 10;
    rmsd_threshold[1] = 1; rmsd_threshold[2] = 1; rmsd_threshold[3] = 1; rmsd_threshold[4] = 1;
    lambda = 0.5;

    for (l = 1; l <= NVEL; l++) for (k = 1; k <= NZP1; k++) UO_(p, k, l) = U_(p, k, l);
    for (l = 1; l <= NSCLR; l++) for (k = 1; k <= NZP1; k++) XO_(p, k, l) = X_(p, k, l);
    p->comp_flag = 1;
    p->reset_flag = 0;
    p->dampu_flag = 0;
    p->dampv_flag = 0;
    iter = 0;

    while (p->comp_flag && p->reset_flag <= comp_iter_max) {
        for (k = 1; k <= NZP1; k++) {
            for (l = 1; l <= NVEL; l++) {
                if (p->old < 0 || p->old > 1) {
                    p->status |= ORC_ST_BAD_OLDNEW;
                    p->old = p->new_;
                }
                if (p->new_ < 0 || p->new_ > 1) {
                    p->status |= ORC_ST_BAD_OLDNEW;
                    p->new_ = p->old;
                }
                U_(p, k, l) = 2. * US_(p, k, l, p->new_) - US_(p, k, l, p->old);
                UX_(p, k, l) = U_(p, k, l);
            }
            for (l = 1; l <= NSCLR; l++) {
                X_(p, k, l) = 2. * XS_(p, k, l, p->new_) - XS_(p, k, l, p->old);
                XX_(p, k, l) = X_(p, k, l);
            }
        }

        iter = 0;
        iconv = 0;

        for (iter = 0; iter <= 2; iter++) {
            for (k = 1; k <= NZP1; k++) {
                for (l = 1; l <= NVEL; l++) {
                    U_(p, k, l) = lambda * UX_(p, k, l) + (1 - lambda) * U_(p, k, l);
                    UX_(p, k, l) = U_(p, k, l);
                }
                for (l = 1; l <= NSCLR; l++) {
                    X_(p, k, l) = lambda * XX_(p, k, l) + (1 - lambda) * X_(p, k, l);
                    XX_(p, k, l) = X_(p, k, l);
                }
            }
            verticalmixing(p, c, &hmixe, &kmixe);
            ocnint(p, c, 1, kmixe);
        }
        /* iter == 3 here (Fortran DO-variable after loop completion) */

        if (c->LKPP) {
            for (;;) { /* 45 continue */
                for (k = 1; k <= NZP1; k++) {
                    for (l = 1; l <= NVEL; l++) {
                        U_(p, k, l) = lambda * UX_(p, k, l) + (1 - lambda) * U_(p, k, l);
                        UX_(p, k, l) = U_(p, k, l);
                    }
                    for (l = 1; l <= NSCLR; l++) {
                        X_(p, k, l) = lambda * XX_(p, k, l) + (1 - lambda) * X_(p, k, l);
                        XX_(p, k, l) = X_(p, k, l);
                    }
                }
                verticalmixing(p, c, &hmixn, &kmixn);
                ocnint(p, c, 1, kmixn);
                iter = iter + 1;

                tol = c->hmixtolfrac * HM_(c, kmixn);
                if (kmixn == NZP1) tol = c->hmixtolfrac * HM_(c, NZ);
                if (fabs(hmixn - hmixe) > tol) {
                    iconv = 0;
                } else {
                    iconv = iconv + 1;
                }
                if (iconv < 3) {
                    if (iter < c->itermax) {
                        hmixe = hmixn;
                        kmixe = kmixn;
                        continue; /* goto 45 */
                    } else {
                        if (hmixn > hmixe) {
                            if (iter >= c->itermax + ORC_ITER_CAP_EXTRA) { /* oracle/GPU safety cap */
                                p->status |= ORC_ST_ITER_CAP;
                                break;
                            }
                            hmixe = hmixn;
                            kmixe = kmixn;
                            continue; /* goto 45 */
                        }
                    }
                }
                break;
            }
            if (iter > (c->itermax + 1)) {
                p->status |= ORC_ST_LONG_ITER;
            }
        }

        p->comp_flag = 0;
        for (k = 1; k <= NZ; k++) {
            if (fabs(U_(p, k, 1)) >= 10 || fabs(U_(p, k, 2)) >= 10 ||
                fabs(X_(p, k, 1) - X_(p, k + 1, 1)) >= 10) {
                p->comp_flag = 1;
                p->f = p->f * 1.01;
            }
        }
        if (!p->comp_flag) {
            rmsd[1] = rmsd[2] = rmsd[3] = rmsd[4] = 0.;
            for (k = 1; k <= NZP1; k++) {
                rmsd[1] = rmsd[1] + (U_(p, k, 1) - UO_(p, k, 1)) * (U_(p, k, 1) - UO_(p, k, 1)) * HM_(c, k) / DM_(c, NZ);
                rmsd[2] = rmsd[2] + (U_(p, k, 2) - UO_(p, k, 2)) * (U_(p, k, 2) - UO_(p, k, 2)) * HM_(c, k) / DM_(c, NZ);
                rmsd[3] = rmsd[3] + (X_(p, k, 1) - XO_(p, k, 1)) * (X_(p, k, 1) - XO_(p, k, 1)) * HM_(c, k) / DM_(c, NZ);
                rmsd[4] = rmsd[4] + (X_(p, k, 2) - XO_(p, k, 2)) * (X_(p, k, 2) - XO_(p, k, 2)) * HM_(c, k) / DM_(c, NZ);
            }
            for (k = 1; k <= 4; k++) {
                rmsd[k] = sqrt(rmsd[k]);
                if (rmsd[k] >= rmsd_threshold[k]) {
                    p->comp_flag = 1;
                    p->f = p->f * 1.01;
                }
            }
        }
        p->reset_flag = p->reset_flag + 1;
        if (p->reset_flag > comp_iter_max) {
            p->status |= ORC_ST_REINT_FAIL;
        }
    }

    for (k = 1; k <= NZ; k++) {
        deltaz = 0.5 * (HM_(c, k) + HM_(c, k + 1));
        for (n = 1; n <= NSCLR; n++) {
            WX_(p, k, n) = -p->difs[k] * ((X_(p, k, n) - X_(p, k + 1, n)) / deltaz - p->ghat[k] * WX_(p, 0, n));
        }
        if (c->LDD)
            WX_(p, k, 1) = -p->dift[k] * ((X_(p, k, 1) - X_(p, k + 1, 1)) / deltaz - p->ghat[k] * WX_(p, 0, 1));
        WX_(p, k, NSP1) = c->grav * (p->talpha[k] * WX_(p, k, 1) - p->sbeta[k] * WX_(p, k, 2));
        for (n = 1; n <= NVEL; n++) {
            WU_(p, k, n) = -p->difm[k] * (U_(p, k, n) - U_(p, k + 1, n)) / deltaz;
        }
    }
    /* energetics Eflx/Esnk/Ptke/Tmke (ocnstep_mod.F90:257-276) are computed into
     * locals and never stored: dead code, not restated. */

    p->hmix = hmixn;
    p->kmix = kmixn;
    p->uref = U_(p, 1, 1);
    p->vref = U_(p, 1, 2);
    p->Tref = X_(p, 1, 1);
    if (c->L_SSref) {
        p->Ssurf = p->SSref;
    } else {
        p->Ssurf = X_(p, 1, 2) + p->Sref;
    }

    if (c->L_DAMP_CURR) {
        dampU[1] = 0.; dampU[2] = 0.;
        for (k = 1; k <= NZP1; k++) {
            for (l = 1; l <= NVEL; l++) {
                a = 0.99 * fabs(U_(p, k, l));
                b = (U_(p, k, l) * U_(p, k, l)) / (c->dt_uvdamp * (86400. / c->dto));
                Ui = fmin(a, b);
                if (b < a) {
                    dampU[l] = dampU[l] + 1.0 / (double)NZP1;
                }
                U_(p, k, l) = U_(p, k, l) - f_sign(Ui, U_(p, k, l));
            }
        }
        p->dampu_flag = dampU[1];
        p->dampv_flag = dampU[2];
    }

    p->old = p->new_;
    p->new_ = 1 - p->old;
    p->hmixd[p->new_] = p->hmix;
    for (k = 1; k <= NZP1; k++) {
        for (l = 1; l <= NVEL; l++) US_(p, k, l, p->new_) = U_(p, k, l);
        for (l = 1; l <= NSCLR; l++) XS_(p, k, l, p->new_) = X_(p, k, l);
    }
    p->iter_final = iter;
}

/* ================================================================== */
/* check_profile  (src/mckpp_physics_overrides.F90:42-125)             */
/* ================================================================== */
static void check_profile(col1d *p, const orc_const *c)
{
    int z, j, l, NZP1 = c->nzp1;
    double dz_total, dtdz_total, dz;

    if (p->comp_flag && c->have_ocnT_file && c->have_sal_file) {
        for (z = 1; z <= NZP1; z++) X_(p, z, 1) = p->ocnT_clim[z];
        for (z = 1; z <= NZP1; z++) X_(p, z, 2) = p->sal_clim[z];
        for (l = 1; l <= NVEL; l++) for (z = 1; z <= NZP1; z++) U_(p, z, l) = UI_(p, z, l);
        p->reset_flag = 999;
        p->status |= ORC_ST_RESET;
    } else if (p->comp_flag) {
        for (l = 1; l <= NVEL; l++) for (z = 1; z <= NZP1; z++) U_(p, z, l) = UI_(p, z, l);
        p->reset_flag = 999;
        p->status |= ORC_ST_RESET;
    }

    if (p->l_ocean && c->L_NO_FREEZE) {
        for (z = 1; z <= NZP1; z++) {
            if (X_(p, z, 1) < -1.8) {
                p->tinc_fcorr[z] = p->tinc_fcorr[z] + (-1.8 - X_(p, z, 1));
                X_(p, z, 1) = -1.8;
                p->freeze_flag = p->freeze_flag + 1.0 / (double)NZP1;
            }
        }
    }

    if (p->l_ocean && c->L_NO_ISOTHERM) {
        dtdz_total = 0.;
        dz_total = 0.;
        for (j = 2; j <= c->iso_bot; j++) {
            dz = ZM_(c, j) - ZM_(c, j - 1);
            dtdz_total = dtdz_total + fabs((X_(p, j, 1) - X_(p, j - 1, 1))) * dz;
            dz_total = dz_total + dz;
        }
        dtdz_total = dtdz_total / dz_total;
        if (fabs(dtdz_total) < c->iso_thresh) {
            for (z = 1; z <= NZP1; z++) X_(p, z, 1) = p->ocnT_clim[z];
            for (z = 1; z <= NZP1; z++) X_(p, z, 2) = p->sal_clim[z];
            p->reset_flag = (-1.) * p->reset_flag;
            p->status |= ORC_ST_ISO_RESET;
        }
    } else {
        p->reset_flag = 0;
    }
}

/* ================================================================== */
/* 3-D <-> 1-D  (src/mckpp_types_transfer.F90:15-193, 199-327)         */
/* ================================================================== */
/* element (ipt, i2 [, i3 [, i4]]) of a column-major array whose first extent is npts */
#define G2(a, ipt, i2)                 (a)[(size_t)((ipt)-1) + (size_t)npts * (size_t)(i2)]
#define G3(a, ipt, i2, n2, i3)         (a)[(size_t)((ipt)-1) + (size_t)npts * ((size_t)(i2) + (size_t)(n2) * (size_t)(i3))]
#define G4(a, ipt, i2, n2, i3, n3, i4) (a)[(size_t)((ipt)-1) + (size_t)npts * ((size_t)(i2) + (size_t)(n2) * ((size_t)(i3) + (size_t)(n3) * (size_t)(i4)))]

static void fields_3dto1d(const orc_3d *s, int point, col1d *p, const orc_const *c)
{
    int i, j, k, npts = c->npts, NZ = c->nz, NZP1 = c->nzp1, NZtmax = c->nztmax, NZP1tmax = c->nzp1tmax;
    for (i = 1; i <= NZP1; i++) {
        for (j = 1; j <= NVEL; j++) {
            U_(p, i, j) = G3(s->U, point, i - 1, NZP1, j - 1);
            UI_(p, i, j) = G3(s->U_init, point, i - 1, NZP1, j - 1);
            for (k = 0; k <= 1; k++) US_(p, i, j, k) = G4(s->Us, point, i - 1, NZP1, j - 1, NVEL, k);
        }
        for (j = 1; j <= NSCLR; j++) {
            X_(p, i, j) = G3(s->X, point, i - 1, NZP1, j - 1);
            for (k = 0; k <= 1; k++) XS_(p, i, j, k) = G4(s->Xs, point, i - 1, NZP1, j - 1, NSCLR, k);
        }
        p->Rig[i] = G2(s->Rig, point, i - 1);
        p->Shsq[i] = G2(s->Shsq, point, i - 1);
        p->swfrac[i] = G2(s->swfrac, point, i - 1);
        p->tinc_fcorr[i] = G2(s->tinc_fcorr, point, i - 1);
        p->sinc_fcorr[i] = G2(s->sinc_fcorr, point, i - 1);
        p->fcorr_withz[i] = G2(s->fcorr_withz, point, i - 1);
        p->scorr[i] = G2(s->scorr, point, i - 1);
        p->sfcorr_withz[i] = G2(s->sfcorr_withz, point, i - 1);
        p->sal_clim[i] = G2(s->sal_clim, point, i - 1);
        p->ocnT_clim[i] = G2(s->ocnT_clim, point, i - 1);
        p->ocnTcorr[i] = G2(s->ocnTcorr, point, i - 1);
        if (i <= NZ) p->dbloc[i] = G2(s->dbloc, point, i - 1);
    }
    for (i = 0; i <= 1; i++) p->hmixd[i] = G2(s->hmixd, point, i);
    for (i = 0; i <= NZP1tmax; i++) {
        p->rho[i] = G2(s->rho, point, i);
        p->cp[i] = G2(s->cp, point, i);
        if (i > 0) p->buoy[i] = G2(s->buoy, point, i - 1);
        if (i <= NZtmax) {
            p->difm[i] = G2(s->difm, point, i);
            p->difs[i] = G2(s->difs, point, i);
            p->dift[i] = G2(s->dift, point, i);
            for (j = 1; j <= NVP1; j++) WU_(p, i, j) = G3(s->wU, point, i, NZtmax + 1, j - 1);
            for (j = 1; j <= NSP1; j++) WX_(p, i, j) = G3(s->wX, point, i, NZtmax + 1, j - 1);
            for (j = 1; j <= NSCLR; j++) WXNT_(p, i, j) = G3(s->wXNT, point, i, NZtmax + 1, j - 1);
            if (i > 0) p->ghat[i] = G2(s->ghat, point, i - 1);
        }
        if (i <= NZ) p->swdk_opt[i] = G2(s->swdk_opt, point, i);
    }
    for (i = 1; i <= 2; i++) {
        p->nmodeadv[i] = s->nmodeadv[(size_t)(point - 1) + (size_t)npts * (size_t)(i - 1)];
        for (j = 1; j <= c->maxmodeadv; j++) {
            MODEADV_(p, j, i) = s->modeadv[(size_t)(point - 1) + (size_t)npts * ((size_t)(j - 1) + (size_t)c->maxmodeadv * (size_t)(i - 1))];
            ADVEC_(p, j, i) = G3(s->advection, point, j - 1, c->maxmodeadv, i - 1);
        }
    }
    for (i = 1; i <= c->nsflxs; i++)
        for (j = 1; j <= 5; j++)
            for (k = 0; k <= c->njdt; k++)
                SFLUX_(p, i, j, k) = G4(s->sflux, point, i - 1, c->nsflxs, j - 1, 5, k);

    p->ocdepth = s->ocdepth[point - 1];
    p->l_ocean = s->l_ocean[point - 1];
    p->l_initflag = s->l_initflag[point - 1];
    p->f = s->f[point - 1];
    p->freeze_flag = s->freeze_flag[point - 1];
    p->relax_sst = s->relax_sst[point - 1];
    p->fcorr = s->fcorr[point - 1];
    p->fcorr_twod = s->fcorr_twod[point - 1];
    p->SST0 = s->SST0[point - 1];
    p->relax_sal = s->relax_sal[point - 1];
    p->relax_ocnT = s->relax_ocnT[point - 1];
    p->hmix = s->hmix[point - 1];
    p->kmix = s->kmix[point - 1];
    p->Tref = s->Tref[point - 1];
    p->uref = s->uref[point - 1];
    p->vref = s->vref[point - 1];
    p->Ssurf = s->Ssurf[point - 1];
    p->Sref = s->Sref[point - 1];
    p->SSref = s->SSref[point - 1];
    p->old = s->old[point - 1];
    p->new_ = s->new_[point - 1];
    p->jerlov = s->jerlov[point - 1];
    p->point = point;
}

static void fields_1dto3d(const col1d *p, int point, orc_3d *s, const orc_const *c)
{
    int i, j, k, npts = c->npts, NZ = c->nz, NZP1 = c->nzp1, NZtmax = c->nztmax, NZP1tmax = c->nzp1tmax;
    for (i = 1; i <= NZP1; i++) {
        for (j = 1; j <= NVEL; j++) {
            G3(s->U, point, i - 1, NZP1, j - 1) = U_(p, i, j);
            for (k = 0; k <= 1; k++) G4(s->Us, point, i - 1, NZP1, j - 1, NVEL, k) = US_(p, i, j, k);
        }
        for (j = 1; j <= NSCLR; j++) {
            G3(s->X, point, i - 1, NZP1, j - 1) = X_(p, i, j);
            for (k = 0; k <= 1; k++) G4(s->Xs, point, i - 1, NZP1, j - 1, NSCLR, k) = XS_(p, i, j, k);
        }
        G2(s->Rig, point, i - 1) = p->Rig[i];
        G2(s->Shsq, point, i - 1) = p->Shsq[i];
        G2(s->swfrac, point, i - 1) = p->swfrac[i];
        G2(s->tinc_fcorr, point, i - 1) = p->tinc_fcorr[i];
        G2(s->sinc_fcorr, point, i - 1) = p->sinc_fcorr[i];
        G2(s->fcorr_withz, point, i - 1) = p->fcorr_withz[i];
        G2(s->scorr, point, i - 1) = p->scorr[i];
        G2(s->sfcorr_withz, point, i - 1) = p->sfcorr_withz[i];
        G2(s->ocnTcorr, point, i - 1) = p->ocnTcorr[i];
        if (i <= NZ) G2(s->dbloc, point, i - 1) = p->dbloc[i];
    }
    for (i = 0; i <= 1; i++) G2(s->hmixd, point, i) = p->hmixd[i];
    for (i = 0; i <= NZP1tmax; i++) {
        G2(s->rho, point, i) = p->rho[i];
        G2(s->cp, point, i) = p->cp[i];
        if (i > 0) G2(s->buoy, point, i - 1) = p->buoy[i];
        if (i <= NZtmax) {
            G2(s->difm, point, i) = p->difm[i];
            G2(s->difs, point, i) = p->difs[i];
            G2(s->dift, point, i) = p->dift[i];
            for (j = 1; j <= NVP1; j++) G3(s->wU, point, i, NZtmax + 1, j - 1) = WU_(p, i, j);
            for (j = 1; j <= NSP1; j++) G3(s->wX, point, i, NZtmax + 1, j - 1) = WX_(p, i, j);
            for (j = 1; j <= NSCLR; j++) G3(s->wXNT, point, i, NZtmax + 1, j - 1) = WXNT_(p, i, j);
            if (i > 0) G2(s->ghat, point, i - 1) = p->ghat[i];
        }
        if (i <= NZ) G2(s->swdk_opt, point, i) = p->swdk_opt[i];
    }
    s->l_initflag[point - 1] = p->l_initflag;
    s->freeze_flag[point - 1] = p->freeze_flag;
    s->fcorr[point - 1] = p->fcorr;
    s->fcorr_twod[point - 1] = p->fcorr_twod;
    s->hmix[point - 1] = p->hmix;
    s->kmix[point - 1] = p->kmix;
    s->Tref[point - 1] = p->Tref;
    s->uref[point - 1] = p->uref;
    s->vref[point - 1] = p->vref;
    s->Ssurf[point - 1] = p->Ssurf;
    s->old[point - 1] = p->old;
    s->new_[point - 1] = p->new_;
    s->reset_flag[point - 1] = p->reset_flag;
    s->dampu_flag[point - 1] = p->dampu_flag;
    s->dampv_flag[point - 1] = p->dampv_flag;
    /* oracle-only diagnostics */
    if (s->diag_talpha)
        for (i = 0; i <= NZP1; i++) G2(s->diag_talpha, point, i) = p->talpha[i];
    if (s->diag_sbeta)
        for (i = 0; i <= NZP1; i++) G2(s->diag_sbeta, point, i) = p->sbeta[i];
}

/* mckpp_physics_overrides_bottomtemp  overrides.F90:12-24 */
static void overrides_bottomtemp(const orc_const *c, orc_3d *s)
{
    int ipt, npts = c->npts, NZP1 = c->nzp1;
    for (ipt = 1; ipt <= npts; ipt++) {
        G2(s->tinc_fcorr, ipt, NZP1 - 1) = s->bottom_temp[ipt - 1] - G3(s->X, ipt, NZP1 - 1, NZP1, 0);
        G2(s->ocnTcorr, ipt, NZP1 - 1) = G2(s->tinc_fcorr, ipt, NZP1 - 1) * G2(s->rho, ipt, NZP1) *
                                         G2(s->cp, ipt, NZP1) / c->dto;
        G3(s->X, ipt, NZP1 - 1, NZP1, 0) = s->bottom_temp[ipt - 1];
    }
}

/* ================================================================== */
/* mckpp_physics_driver  (src/mckpp_physics_driver_mod.F90:15-73)      */
/* ================================================================== */
int orc_physics_driver(const orc_const *c, orc_3d *s, int ntime, int nthreads, int realloc_1d)
{
    int any_pivot = 0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#else
    (void)nthreads;
#endif
#pragma omp parallel
    {
        col1d p;
        int ipt;
        int have = 0;
#pragma omp for schedule(dynamic)
        for (ipt = 1; ipt <= c->npts; ipt++) {
            if (s->run_physics[ipt - 1]) {
                if (realloc_1d || !have) {
                    if (have) col1d_free(&p);
                    col1d_alloc(&p, c);
                    have = 1;
                }
                p.status = 0;
                p.ntime = ntime;
                fields_3dto1d(s, ipt, &p, c);
                ocnstep(&p, c);
                p.nreint = (int)p.reset_flag;
                check_profile(&p, c);
                fields_1dto3d(&p, ipt, s, c);
                if (s->diag_iter) s->diag_iter[ipt - 1] = p.iter_final;
                if (s->diag_nreint) s->diag_nreint[ipt - 1] = p.nreint;
                if (s->diag_status) s->diag_status[ipt - 1] = p.status;
                if (p.status & ORC_ST_PIVOT_ZERO) {
#pragma omp atomic write
                    any_pivot = 1;
                }
            }
        }
        if (have) col1d_free(&p);
    }
    if (c->L_VARY_BOTTOM_TEMP) overrides_bottomtemp(c, s);
    return any_pivot ? -1 : 0;
}

/* per-column loop of MCKPP_INITIALIZE_OCEAN_MODEL  initialize_ocean.F90:54-104 */
int orc_initialize_ocean_model(const orc_const *c, orc_3d *s, int nthreads)
{
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#else
    (void)nthreads;
#endif
#pragma omp parallel
    {
        col1d p;
        int ipt, k, l, n, NZ = c->nz, NZP1 = c->nzp1;
        double hmix0 = 0, deltaz;
        int kmix0 = 0;
        col1d_alloc(&p, c);
#pragma omp for schedule(dynamic)
        for (ipt = 1; ipt <= c->npts; ipt++) {
            if (s->run_physics[ipt - 1]) {
                p.status = 0;
                p.ntime = 0; /* mckpp_initialize_time: ntime = 0 (time_control.F90:31) */
                fields_3dto1d(s, ipt, &p, c);
                p.l_initflag = 1;
                verticalmixing(&p, c, &hmix0, &kmix0);
                p.l_initflag = 0;
                p.hmix = hmix0;
                p.kmix = kmix0;
                p.Tref = X_(&p, 1, 1);
                for (k = 1; k <= NZ; k++) {
                    deltaz = 0.5 * (HM_(c, k) + HM_(c, k + 1));
                    for (n = 1; n <= NSCLR; n++) {
                        WX_(&p, k, n) = -p.difs[k] * ((X_(&p, k, n) - X_(&p, k + 1, n)) / deltaz - p.ghat[k] * WX_(&p, 0, n));
                    }
                    if (c->LDD)
                        WX_(&p, k, 1) = -p.dift[k] * ((X_(&p, k, 1) - X_(&p, k + 1, 1)) / deltaz - p.ghat[k] * WX_(&p, 0, 1));
                    WX_(&p, k, NSP1) = c->grav * (p.talpha[k] * WX_(&p, k, 1) - p.sbeta[k] * WX_(&p, k, 2));
                    for (n = 1; n <= NVEL; n++)
                        WU_(&p, k, n) = -p.difm[k] * (U_(&p, k, n) - U_(&p, k + 1, n)) / deltaz;
                }
                p.old = 0;
                p.new_ = 1;
                p.hmixd[0] = p.hmix;
                p.hmixd[1] = p.hmix;
                for (k = 1; k <= NZP1; k++) {
                    for (l = 1; l <= NVEL; l++) {
                        US_(&p, k, l, 0) = U_(&p, k, l);
                        US_(&p, k, l, 1) = U_(&p, k, l);
                    }
                    for (l = 1; l <= NSCLR; l++) {
                        XS_(&p, k, l, 0) = X_(&p, k, l);
                        XS_(&p, k, l, 1) = X_(&p, k, l);
                    }
                }
                /* 1dto3d copies reset/damp flags that ocnstep would have set;
                 * at init they hold what 3dto1d left: kpp_1d is INTENT(OUT), the
                 * values are undefined in the reference.  Keep the 3-D values. */
                p.reset_flag = s->reset_flag[ipt - 1];
                p.dampu_flag = s->dampu_flag[ipt - 1];
                p.dampv_flag = s->dampv_flag[ipt - 1];
                fields_1dto3d(&p, ipt, s, c);
            }
        }
        col1d_free(&p);
    }
    return 0;
}

/* ================================================================== */
/* input builders                                                      */
/* ================================================================== */

/* tri  initialize_ocean.F90:34-43 */
void orc_build_tri(int nz, int nztmax, double dto, const double *zm, const double *hm, double *tri)
{
    int k;
    double *dzb = dalloc((size_t)nz + 1);
#define T_(k, j) tri[(j) * (nztmax + 1) + (k)]
    for (k = 1; k <= nz; k++) dzb[k] = zm[k - 1] - zm[k];
    T_(0, 1) = dto / hm[0];
    T_(1, 1) = dto / hm[0] / dzb[1];
    for (k = 2; k <= nz; k++) {
        T_(k, 1) = dto / hm[k - 1] / dzb[k];
        T_(k, 0) = dto / hm[k - 1] / dzb[k - 1];
    }
#undef T_
    free(dzb);
}

/* mckpp_physics_lookup  physics_lookup_mod.F90:42-64 */
void orc_build_lookup(double vonk, double *wmt, double *wst)
{
    double zmin, zmax, umin, umax, usta, zeta, zehat, epsln, am, cm, c1, c2, zetam, as, cs, c3,
        zetas, deltau, deltaz;
    int i, j, ni, nj;
    ni = 890; nj = 48; epsln = 1.e-20; c1 = 5.0; zmin = -4.e-7; zmax = 0.0; umin = 0.0; umax = 0.04;
    am = 1.257; cm = 8.380; c2 = 16.0; zetam = -0.2; as = -28.86; cs = 98.96; c3 = 16.0; zetas = -1.0;
    deltaz = (zmax - zmin) / (ni + 1);
    deltau = (umax - umin) / (nj + 1);
    for (i = 0; i <= ni + 1; i++) {
        zehat = deltaz * (i) + zmin;
        for (j = 0; j <= nj + 1; j++) {
            usta = deltau * (j) + umin;
            zeta = zehat / (usta * usta * usta + epsln);
            if (zehat >= 0.) {
                wmt[j * 892 + i] = vonk * usta / (1. + c1 * zeta);
                wst[j * 892 + i] = wmt[j * 892 + i];
            } else {
                if (zeta > zetam)
                    wmt[j * 892 + i] = vonk * usta * pow(1. - c2 * zeta, 1. / 4.);
                else
                    wmt[j * 892 + i] = vonk * pow(am * (usta * usta * usta) - cm * zehat, 1. / 3.);
                if (zeta > zetas)
                    wst[j * 892 + i] = vonk * usta * pow(1. - c3 * zeta, 1. / 2.);
                else
                    wst[j * 892 + i] = vonk * pow(as * (usta * usta * usta) - cs * zehat, 1. / 3.);
            }
        }
    }
}

/* vertical grid  initialize_geography_mod.F90:43-74 (no vgrid file) */
void orc_build_grid(int nz, double dmax, int l_stretchgrid, double dscale, double *zm, double *hm, double *dm)
{
    double sumh = 0.0, hsum, dfac, sk;
    int i;
    if (l_stretchgrid) {
        sumh = 0.0;
        dfac = 1.0 - exp(-dscale);
        for (i = 1; i <= nz; i++) {
            sk = -((double)i - 0.5) / (double)nz;
            hm[i - 1] = dmax * dfac / (double)nz / dscale / (1.0 + sk * dfac);
            sumh = sumh + hm[i - 1];
        }
    }
    hsum = 0.0;
    for (i = 1; i <= nz; i++) {
        if (l_stretchgrid)
            hm[i - 1] = hm[i - 1] * dmax / sumh;
        else
            hm[i - 1] = dmax / (double)nz;
        zm[i - 1] = 0.0 - (hsum + 0.5 * hm[i - 1]);
        hsum = hsum + hm[i - 1];
        dm[i] = hsum;
    }
    dm[0] = 0.0;
    hm[nz] = 1.e-10;
    zm[nz] = -dmax;
}

/* Coriolis  initialize_geography_mod.F90:78-88 ; twopi = 8*atan(1.) (namelist_mod.F90:94) */
void orc_coriolis(int npts, const double *dlat, double *f)
{
    double twopi = 8 * atan(1.);
    int ipt;
    for (ipt = 0; ipt < npts; ipt++) {
        if (fabs(dlat[ipt]) < 2.5)
            f[ipt] = 2. * (twopi / 86164.) * sin(2.5 * twopi / 360.) * f_sign(1., dlat[ipt]);
        else
            f[ipt] = 2. * (twopi / 86164.) * sin(dlat[ipt] * twopi / 360.);
    }
}

/* forcing map of mckpp_fluxes  fluxes_mod.F90:56-72 (l_rest = .FALSE.) */
void orc_fluxes_map(int npts, int nsflxs, double flsn, double el, double *taux, const double *tauy,
                    const double *swf, const double *lwf, const double *lhf, const double *shf,
                    const double *rain, const double *snow, const int32_t *l_ocean, double *sflux)
{
    int ipt;
#define SF_(ipt, i) sflux[(size_t)(ipt) + (size_t)npts * ((size_t)((i)-1) + (size_t)nsflxs * (size_t)4)]
    for (ipt = 0; ipt < npts; ipt++) {
        if (l_ocean[ipt]) {
            if ((taux[ipt] == 0.0) && (tauy[ipt] == 0.0)) taux[ipt] = 1.e-10;
            SF_(ipt, 1) = taux[ipt];
            SF_(ipt, 2) = tauy[ipt];
            SF_(ipt, 3) = swf[ipt];
            SF_(ipt, 4) = lwf[ipt] + lhf[ipt] + shf[ipt] - snow[ipt] * flsn;
            SF_(ipt, 5) = 1e-10;
            SF_(ipt, 6) = rain[ipt] + snow[ipt] + (lhf[ipt] / el);
        }
    }
#undef SF_
}

const char *orc_3d_member_names(void)
{
    return "U,X,Rig,dbloc,Shsq,hmixd,Us,Xs,rho,cp,buoy,ocdepth,f,swfrac,swdk_opt,difm,difs,dift,"
           "wU,wX,wXNT,ghat,relax_sst,fcorr,SST0,fcorr_twod,tinc_fcorr,sinc_fcorr,fcorr_withz,"
           "sfcorr_withz,advection,relax_sal,scorr,relax_ocnT,ocnTcorr,sal_clim,ocnT_clim,hmix,kmix,"
           "Tref,uref,vref,Ssurf,Sref,SSref,sflux,freeze_flag,reset_flag,dampu_flag,dampv_flag,U_init,"
           "bottom_temp,l_ocean,l_initflag,run_physics,old,new_,jerlov,nmodeadv,modeadv,"
           "diag_iter,diag_nreint,diag_status,diag_talpha,diag_sbeta";
}

const char *orc_const_member_names(void)
{
    return "nz,nzp1,nztmax,nzp1tmax,npts,nsflxs,njdt,maxmodeadv,itermax,iso_bot,dt_uvdamp,"
           "LKPP,LRI,LDD,L_SSref,L_RELAX_SST,L_RELAX_CALCONLY,L_FCORR,L_FCORR_WITHZ,"
           "L_SFCORR,L_SFCORR_WITHZ,L_RELAX_SAL,L_RELAX_OCNT,L_NO_FREEZE,L_NO_ISOTHERM,L_DAMP_CURR,"
           "L_VARY_BOTTOM_TEMP,have_ocnT_file,have_sal_file,pad0,hmixtolfrac,dto,grav,vonk,sice,"
           "iso_thresh,zm,hm,dm,tri,wmt,wst";
}

/* ================================================================== */
/* Second-reading support (oracle/second_reading.py, oracle/AUDIT.md): */
/* one column driven from Python, so that an independently written     */
/* restatement of mckpp_physics_ocnstep's control flow and of bldepth  */
/* can be run against this file's physics routines and compared bit    */
/* for bit.  TEST INFRASTRUCTURE ONLY.                                 */
/* ================================================================== */
void *orc_col_new(const orc_const *c)
{
    col1d *p = (col1d *)malloc(sizeof(col1d));
    col1d_alloc(p, c);
    return p;
}
void orc_col_free(void *h)
{
    col1d *p = (col1d *)h;
    col1d_free(p);
    free(p);
}
/* mckpp_fields_3dto1d for `point` (1-based), with the module variable ntime */
void orc_col_load(void *h, const orc_const *c, const orc_3d *s, int point, int ntime)
{
    col1d *p = (col1d *)h;
    p->status = 0;
    p->ntime = ntime;
    fields_3dto1d(s, point, p, c);
}
/* mckpp_fields_1dto3d + the oracle-only diagnostics orc_physics_driver stores */
void orc_col_store(void *h, const orc_const *c, orc_3d *s, int point, int iter_final, int nreint)
{
    col1d *p = (col1d *)h;
    fields_1dto3d(p, point, s, c);
    if (s->diag_iter) s->diag_iter[point - 1] = iter_final;
    if (s->diag_nreint) s->diag_nreint[point - 1] = nreint;
    if (s->diag_status) s->diag_status[point - 1] = p->status;
}
void orc_col_vmix(void *h, const orc_const *c, double *hmix, int *kmix)
{
    verticalmixing((col1d *)h, c, hmix, kmix);
}
/* Uo(nzp1,2), Xo(nzp1,2): the entry state ocnstep passes down (column-major) */
void orc_col_ocnint(void *h, const orc_const *c, int kmix, const double *Uo, const double *Xo)
{
    col1d *p = (col1d *)h;
    int k, l;
    for (l = 1; l <= 2; l++)
        for (k = 1; k <= c->nzp1; k++) {
            UO_(p, k, l) = Uo[(l - 1) * c->nzp1 + k - 1];
            XO_(p, k, l) = Xo[(l - 1) * c->nzp1 + k - 1];
        }
    ocnint(p, c, 1, kmix);
}
void orc_col_check_profile(void *h, const orc_const *c) { check_profile((col1d *)h, c); }
/* array members: index 0 of the returned pointer is Fortran index `*lb`, `*n` elements */
double *orc_col_array(void *h, const char *name, int *lb, int *n)
{
    col1d *p = (col1d *)h;
    const int n1 = p->nzp1 + 1, nt = p->nztmax + 1;
#define ARR(nm, ptr, lb_, n_) if (!strcmp(name, nm)) { *lb = (lb_); *n = (n_); return (ptr); }
    /* 1-based (k,l) members carry an unused slot 0 per component: stride n1 */
    ARR("U", p->U, 0, 2 * n1) ARR("X", p->X, 0, 2 * n1) ARR("Us", p->Us, 0, 4 * n1) ARR("Xs", p->Xs, 0, 4 * n1)
    ARR("hmixd", p->hmixd, 0, 2) ARR("difm", p->difm, 0, nt) ARR("difs", p->difs, 0, nt) ARR("dift", p->dift, 0, nt)
    ARR("ghat", p->ghat, 0, nt) ARR("wU", p->wU, 0, 3 * nt) ARR("wX", p->wX, 0, 3 * nt)
    ARR("talpha", p->talpha, 0, p->nzp1tmax + 1) ARR("sbeta", p->sbeta, 0, p->nzp1tmax + 1)
    ARR("dVsq", p->dVsq, 0, n1) ARR("Ritop", p->Ritop, 0, n1) ARR("dbloc", p->dbloc, 0, p->nz + 1)
    ARR("swfrac", p->swfrac, 0, n1) ARR("sflux", p->sflux, 1, p->nsflxs * 5 * (p->njdt + 1))
    /* inputs and write-only results of ocnint (second reading of ocnint / the solvers) */
    ARR("wXNT", p->wXNT, 0, 2 * nt) ARR("rho", p->rho, 0, p->nzp1tmax + 1) ARR("cp", p->cp, 0, p->nzp1tmax + 1)
    ARR("tinc_fcorr", p->tinc_fcorr, 0, n1) ARR("sinc_fcorr", p->sinc_fcorr, 0, n1) ARR("ocnTcorr", p->ocnTcorr, 0, n1)
    ARR("scorr", p->scorr, 0, n1) ARR("fcorr_withz", p->fcorr_withz, 0, n1) ARR("sfcorr_withz", p->sfcorr_withz, 0, n1)
    ARR("ocnT_clim", p->ocnT_clim, 0, n1) ARR("sal_clim", p->sal_clim, 0, n1)
    ARR("advection", p->advection, 1, p->maxmodeadv * 2)
    /* inputs and results of the second half of vmix (second reading of kppmix and what it calls) */
    ARR("alphaDT", p->alphaDT, 0, n1) ARR("betaDS", p->betaDS, 0, n1) ARR("Shsq", p->Shsq, 0, n1) ARR("Rig", p->Rig, 0, n1)
    ARR("buoy", p->buoy, 0, p->nzp1tmax + 1) ARR("swdk_opt", p->swdk_opt, 0, p->nz + 1) ARR("U_init", p->U_init, 0, 2 * n1)
#undef ARR
    *lb = 0; *n = 0;
    return NULL;
}
double orc_col_get(void *h, const char *name)
{
    col1d *p = (col1d *)h;
#define SC(nm, v) if (!strcmp(name, nm)) return (double)(v);
    SC("f", p->f) SC("old", p->old) SC("new", p->new_) SC("reset_flag", p->reset_flag) SC("comp_flag", p->comp_flag)
    SC("status", p->status) SC("SSref", p->SSref) SC("Sref", p->Sref) SC("ocdepth", p->ocdepth) SC("jerlov", p->jerlov)
    SC("l_initflag", p->l_initflag) SC("ntime", p->ntime)
    SC("relax_sst", p->relax_sst) SC("SST0", p->SST0) SC("fcorr", p->fcorr) SC("fcorr_twod", p->fcorr_twod)
    SC("relax_sal", p->relax_sal) SC("relax_ocnT", p->relax_ocnT) SC("nmodeadv2", p->nmodeadv[2])
    SC("Ssurf", p->Ssurf) SC("rhoh2o", p->rhoh2o) SC("l_ocean", p->l_ocean) SC("freeze_flag", p->freeze_flag)
    SC("dbg_ustar", p->dbg_ustar) SC("dbg_Bo", p->dbg_Bo) SC("dbg_Bosol", p->dbg_Bosol) SC("dbg_hbl", p->dbg_hbl)
    SC("dbg_bfsfc", p->dbg_bfsfc) SC("dbg_stable", p->dbg_stable) SC("dbg_caseA", p->dbg_caseA) SC("dbg_kbl", p->dbg_kbl)
#undef SC
    return 0.0 / 0.0;
}
/* modeadv(j, i), j = 1..maxmodeadv, i = 1..2 */
int orc_col_modeadv(void *h, int j, int i) { col1d *p = (col1d *)h; return MODEADV_(p, j, i); }
void orc_col_set(void *h, const char *name, double v)
{
    col1d *p = (col1d *)h;
#define SS(nm, lhs, T) if (!strcmp(name, nm)) { lhs = (T)v; return; }
    SS("f", p->f, double) SS("old", p->old, int) SS("new", p->new_, int) SS("reset_flag", p->reset_flag, double)
    SS("comp_flag", p->comp_flag, int) SS("status", p->status, int) SS("dampu_flag", p->dampu_flag, double)
    SS("dampv_flag", p->dampv_flag, double) SS("hmix", p->hmix, double) SS("kmix", p->kmix, double)
    SS("uref", p->uref, double) SS("vref", p->vref, double) SS("Tref", p->Tref, double) SS("Ssurf", p->Ssurf, double)
    SS("l_initflag", p->l_initflag, int) SS("ocdepth", p->ocdepth, double) SS("jerlov", p->jerlov, int)
    SS("fcorr", p->fcorr, double) SS("freeze_flag", p->freeze_flag, double)
#undef SS
}
