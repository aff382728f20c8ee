// kpp_api.cu -- the C ABI of include/kpp_gpu.h: handle, device-resident state,
// host<->device movement of kpp_3d_fields members, step / init launches, report.
//
// The handle owns a device mirror of the members of the reference's
// `kpp_3d_fields` (src/mckpp_data_fields.F90:8-101) that the column physics
// touches.  Host arrays are column-fastest already (first Fortran extent is
// npts), so every transfer is a pitched 2-D copy of whole rows; the device
// leading dimension is npts rounded up to 32.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>

#include "../../include/kpp_gpu.h"
#include "kpp_dev.h"

extern "C" {
cudaError_t kpp_launch_step_strict(const KppDevArgs *, KppReportDev *, int, cudaStream_t);
cudaError_t kpp_launch_step_fast(const KppDevArgs *, KppReportDev *, int, cudaStream_t);
cudaError_t kpp_launch_init_strict(const KppDevArgs *, cudaStream_t);
cudaError_t kpp_launch_init_fast(const KppDevArgs *, cudaStream_t);
cudaError_t kpp_launch_test_eos_strict(int, const double *, const double *, const double *, double *, double *,
                                       double *, double *, cudaStream_t);
cudaError_t kpp_launch_test_eos_fast(int, const double *, const double *, const double *, double *, double *,
                                     double *, double *, cudaStream_t);
cudaError_t kpp_launch_test_wscale_strict(const KppDevArgs *, int, const double *, const double *, const double *,
                                          const double *, double *, double *, cudaStream_t);
cudaError_t kpp_launch_test_wscale_fast(const KppDevArgs *, int, const double *, const double *, const double *,
                                        const double *, double *, double *, cudaStream_t);
cudaError_t kpp_launch_test_swfrac_strict(int, const double *, const int *, double *, cudaStream_t);
cudaError_t kpp_launch_test_swfrac_fast(int, const double *, const int *, double *, cudaStream_t);
cudaError_t kpp_launch_fluxmap_strict(int, int, const double *, const int *, double, double, double *, cudaStream_t);
cudaError_t kpp_launch_main_strict(const KppDevArgs *, cudaStream_t);
cudaError_t kpp_launch_main_fast(const KppDevArgs *, cudaStream_t);
cudaError_t kpp_launch_coop_strict(const KppDevArgs *, cudaStream_t);
cudaError_t kpp_launch_coop_fast(const KppDevArgs *, cudaStream_t);
cudaError_t kpp_launch_report_strict(const KppDevArgs *, KppReportDev *, cudaStream_t);
cudaError_t kpp_launch_report_fast(const KppDevArgs *, KppReportDev *, cudaStream_t);
int kpp_exp_is_host_libm_strict(void);
int kpp_exp_is_host_libm_fast(void);
int kpp_coop_fits_strict(int);
cudaError_t kpp_launch_pack_rows_strict(int, int, const void *, int, long, int, double *, long, const double *, cudaStream_t);
cudaError_t kpp_launch_blend_strict(size_t, const double *, const double *, double, double, double *, cudaStream_t);
cudaError_t kpp_launch_test_div_strict(int, const double *, const double *, double *, cudaStream_t);
cudaError_t kpp_launch_test_div_fast(int, const double *, const double *, double *, cudaStream_t);
}

namespace {

// how one member of kpp_3d_fields maps onto its device mirror
struct FieldMap {
    const char *name;
    int elem;            // bytes per element (8 = REAL, 4 = INTEGER/LOGICAL)
    long host_rows;      // rows (of npts) of the whole host array
    int ncomp;           // components moved
    long host_comp_rows; // host rows between components
    long host_row0;      // first host row moved (of component 0)
    long rows;           // rows moved per component
    long dev_row0;       // device row of the first moved row (component 0)
    long dev_comp_rows;  // device rows between components
    long dev_rows;       // device rows allocated
    void **dev;          // where the device pointer lives
};

}  // namespace

struct kpp_handle {
    kpp_dims d;
    kpp_consts k;
    int device;
    int ld;
    cudaStream_t stream;
    cudaEvent_t ev0, ev1;
    KppDevArgs a;
    std::vector<FieldMap> map;
    std::vector<void *> allocs;
    KppReportDev *rep_dev;
    KppReportDev *rep_host;     // pinned
    int last_ntime;
    bool stepped;
    std::string err;
    // device buffers that KppDevArgs declares const
    double *U_init, *Sref, *SSref, *f, *ocdepth, *sflux, *relax_sst, *SST0, *fcorr_twod, *relax_sal, *relax_ocnT,
        *sal_clim, *ocnT_clim, *fcorr_withz, *sfcorr_withz, *bottom_temp, *advection;
    int *jerlov, *l_ocean, *run_physics, *nmodeadv, *modeadv;
    std::vector<double *> slots;
    long long launches;
    int pass_budget_req;
    double *stage;              // (npts, 2*nzp1) packing area of kpp_gpu_pack_output
    double *clim_rec[2][2];     // [ocnT, sal][prev, next] resident climatology records (ld x nzp1)
    double *rawflux;    // 8 rows x ld: staging of the raw flux fields (kpp_gpu_upload_fluxes)
    bool guard;                                       // KPP_GUARD=1: canary zones around every device array
    std::vector<std::pair<char *, size_t>> guards;    // (allocation base, payload bytes)
    // Where this handle's columns sit in the HOST arrays: a plain handle owns all of them
    // (host_npts == d.npts, host_col0 == 0); a part of a multi-GPU group owns the contiguous block
    // [host_col0, host_col0 + d.npts) of host arrays whose first extent is host_npts.
    int host_npts, host_col0;
    // Multi-GPU group (kpp_gpu_create_multi): the group handle owns no device state; every call
    // fans out to parts[i] (one per device, own stream), whose transfers land in / come from
    // disjoint column slices of the host's own arrays -- no collective anywhere.
    std::vector<kpp_handle *> parts;
    // host copies for the deferred 'mode out of range' check (solvers.F90:320-324)
    std::vector<int32_t> nmodeadv_host, modeadv_host;
    bool modeadv_dirty;
    // asynchronous stragglers (kpp_gpu_set_async_stragglers), see step_lagged()
    struct Lag {
        bool on, pending;            // enabled / steps queued since the last join
        long k;                      // steps since the last join
        cudaStream_t sB, sC;         // B: finishes a step's hand-overs, C: the lane's own steps
        cudaEvent_t ev_main[2], ev_fin[2], ev_zero[3], ev_lane, ev_reset;
        int *cont_list2[2], *cont_count2, *lane_list[3], *lane_count, *in_lane;
        int *next_fin, *next_lane;   // fetch counters of the two side-stream cooperative launches
        int *last_count;             // hand-over count of the last step (for the report)
        std::vector<cudaEvent_t> trace;   // KPP_LAG_TRACE=1: six timing events per step (A, C, B: begin/end)
        int *trace_counts;                // ... and the lengths of the hand-over and lane lists (pinned)
    } lag;
    // asynchronous output ring (kpp_gpu_output_ring_*)
    struct Ring {
        std::vector<int> ids;
        std::vector<size_t> dev_off, host_off;   // element offsets of each block in a device / host slot
        std::vector<int> rows;
        size_t dev_elems, host_elems;
        int depth, next;
        std::vector<double *> dev;                // [depth] device staging (dense npts x rows blocks)
        std::vector<double *> host;               // [depth] pinned host slots (owned by the group or the plain handle)
        std::vector<cudaEvent_t> ev_packed, ev_done;
        std::vector<char> submitted;
        cudaStream_t io;
        bool owns_host;
    } *ring;
};

namespace {

thread_local std::string g_err;

int fail(kpp_handle *h, int code, const std::string &msg)
{
    if (h) h->err = msg;
    g_err = msg;
    return code;
}

#define CU(call)                                                                                          \
    do {                                                                                                  \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess)                                                                            \
            return fail(h, KPP_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));              \
    } while (0)

// KPP_GUARD=1 (debugging aid; compute-sanitizer is not available on every pool): every device array gets a
// guard zone filled with a canary on either side; kpp_gpu_debug_check_guards reports any byte a kernel wrote
// outside its arrays.
constexpr size_t GUARD_BYTES = 64 * 1024;
constexpr unsigned char GUARD_BYTE = 0xA5;

template <class T>
int dev_alloc(kpp_handle *h, T **p, size_t n)
{
    const size_t g = h->guard ? GUARD_BYTES : 0;
    void *q = nullptr;
    cudaError_t e = cudaMalloc(&q, n * sizeof(T) + 2 * g);
    if (e != cudaSuccess) return fail(h, KPP_E_NOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(e));
    if (g) {
        e = cudaMemsetAsync(q, GUARD_BYTE, n * sizeof(T) + 2 * g, h->stream);
        if (e != cudaSuccess) return fail(h, KPP_E_CUDA, std::string("cudaMemset: ") + cudaGetErrorString(e));
        h->guards.push_back({(char *)q, n * sizeof(T)});
    }
    e = cudaMemsetAsync((char *)q + g, 0, n * sizeof(T), h->stream);
    if (e != cudaSuccess) return fail(h, KPP_E_CUDA, std::string("cudaMemset: ") + cudaGetErrorString(e));
    h->allocs.push_back(q);
    *p = (T *)((char *)q + g);
    return 0;
}

template <class T>
int dev_upload(kpp_handle *h, const T **dst, const std::vector<T> &src)
{
    T *q = nullptr;
    int rc = dev_alloc(h, &q, src.size());
    if (rc) return rc;
    cudaError_t e = cudaMemcpyAsync(q, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice, h->stream);
    if (e != cudaSuccess) return fail(h, KPP_E_CUDA, std::string("cudaMemcpy: ") + cudaGetErrorString(e));
    // the source vector dies with the caller: finish the copy now
    e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) return fail(h, KPP_E_CUDA, std::string("sync: ") + cudaGetErrorString(e));
    *dst = q;
    return 0;
}

// Jerlov water types (swfrac_mod.F90:28-34, fluxes_mod.F90:128-132)
const double jw_rfac[5] = {0.58, 0.62, 0.67, 0.77, 0.78};
const double jw_a1[5] = {0.35, 0.6, 1.0, 1.5, 1.4};
const double jw_a2[5] = {23.0, 20.0, 17.0, 14.0, 7.9};

// Per-level tables derived from the grid on the host.  Each entry is computed with
// the very expression the reference evaluates per column and per pass, so the
// device reads the same IEEE double it would have computed.
int build_tables(kpp_handle *h, const double *zm_, const double *hm_, const double *dm_, const double *tri_,
                 const double *wmt, const double *wst)
{
    const int nz = h->d.nz, nzp1 = nz + 1, nzt = h->d.nztmax;
    const double dto = h->k.dto;
    auto zm = [&](int k) { return zm_[k - 1]; };
    auto hm = [&](int k) { return hm_[k - 1]; };
    std::vector<double> v_zm(nzp1 + 1, 0.0), v_hm(nzp1 + 2, 0.0), v_dm(nz + 1, 0.0), v_tri0(nz + 1, 0.0),
        v_tri1(nz + 1, 0.0), v_p0(nzp1 + 1, 0.0), v_dzb(nz + 1, 0.0), v_dtoh(nzp1 + 1, 0.0), v_deltaz(nz + 1, 0.0),
        v_zint(nz + 1, 0.0), v_zref(nz + 1, 0.0), v_wz0(nz + 1, 0.0), v_zrmz(nz + 1, 0.0);
    for (int k = 1; k <= nzp1; k++) {
        v_zm[k] = zm(k);
        v_hm[k] = hm(k);
        v_p0[k] = (-zm(k)) / 10.0;
        v_dtoh[k] = dto / hm(k);
    }
    for (int k = 0; k <= nz; k++) {
        v_dm[k] = dm_[k];
        v_tri0[k] = tri_[0 * (nzt + 1) + k];
        v_tri1[k] = tri_[1 * (nzt + 1) + k];
    }
    const double epsilon = 0.1;
    std::vector<int> refoff(nz + 2, 0);
    std::vector<double> refwz, refdel;
    for (int n = 1; n <= nz; n++) {
        v_dzb[n] = zm(n) - zm(n + 1);
        v_deltaz[n] = 0.5 * (hm(n) + hm(n + 1));
        v_zint[n] = -zm(n) + 0.5 * hm(n);
        const double zref = epsilon * zm(n);
        v_zref[n] = zref;
        v_wz0[n] = fmax(zm(1), zref);
        v_zrmz[n] = zref - zm(n);
        refoff[n] = (int)refwz.size();
        for (int kl = 1; kl <= nz; kl++) {
            if (zref >= zm(kl)) break;
            const double wz = fmin(zm(kl) - zm(kl + 1), zm(kl) - zref);
            const double del = 0.5 * wz / (zm(kl) - zm(kl + 1));
            refwz.push_back(wz);
            refdel.push_back(del);
        }
    }
    refoff[nz + 1] = (int)refwz.size();
    if (refwz.empty()) { refwz.push_back(0.0); refdel.push_back(0.0); }

    // Jerlov tables: MCKPP_PHYSICS_SWFRAC_OPT(hbf=1.0) and mckpp_fluxes_swdk(-dm(k)),
    // filled with the host libm exp like the reference's ntime<=1 fill
    std::vector<double> swfrac_tab(5 * (nzp1 + 1), 0.0), swdk_tab(5 * (nz + 1), 0.0);
    for (int j = 0; j < 5; j++) {
        const double rmin = -80., fact = 1.0;
        for (int l = 1; l <= nzp1; l++) {
            const double r1 = fmax(zm(l) * fact / jw_a1[j], rmin);
            const double r2 = fmax(zm(l) * fact / jw_a2[j], rmin);
            swfrac_tab[j * (nzp1 + 1) + l] = jw_rfac[j] * exp(r1) + (1. - jw_rfac[j]) * exp(r2);
        }
        for (int k = 0; k <= nz; k++) {
            const double z = -dm_[k];
            swdk_tab[j * (nz + 1) + k] = jw_rfac[j] * exp(z / jw_a1[j]) + (1.0 - jw_rfac[j]) * exp(z / jw_a2[j]);
        }
    }
    std::vector<double2> wtab(892 * 50);
    for (int j = 0; j < 50; j++)
        for (int i = 0; i < 892; i++) wtab[j * 892 + i] = make_double2(wmt[j * 892 + i], wst[j * 892 + i]);

    KppDevArgs &a = h->a;
    int rc = 0;
    if ((rc = dev_upload(h, &a.zm, v_zm))) return rc;
    if ((rc = dev_upload(h, &a.hm, v_hm))) return rc;
    if ((rc = dev_upload(h, &a.dm, v_dm))) return rc;
    if ((rc = dev_upload(h, &a.tri0, v_tri0))) return rc;
    if ((rc = dev_upload(h, &a.tri1, v_tri1))) return rc;
    if ((rc = dev_upload(h, &a.p0, v_p0))) return rc;
    if ((rc = dev_upload(h, &a.dzb, v_dzb))) return rc;
    if ((rc = dev_upload(h, &a.dtoh, v_dtoh))) return rc;
    if ((rc = dev_upload(h, &a.deltaz, v_deltaz))) return rc;
    if ((rc = dev_upload(h, &a.zint, v_zint))) return rc;
    if ((rc = dev_upload(h, &a.zref, v_zref))) return rc;
    if ((rc = dev_upload(h, &a.wz0, v_wz0))) return rc;
    if ((rc = dev_upload(h, &a.zrmz, v_zrmz))) return rc;
    if ((rc = dev_upload(h, &a.refoff, refoff))) return rc;
    if ((rc = dev_upload(h, &a.refwz, refwz))) return rc;
    if ((rc = dev_upload(h, &a.refdel, refdel))) return rc;
    if ((rc = dev_upload(h, &a.swfrac_tab, swfrac_tab))) return rc;
    if ((rc = dev_upload(h, &a.swdk_tab, swdk_tab))) return rc;
    if ((rc = dev_upload(h, &a.wtab, wtab))) return rc;
    a.dmNZ = dm_[nz];
    return 0;
}

const char *const kFieldNames[KPP_F__COUNT] = {
    "U", "X", "Us", "Xs", "hmixd", "old", "new", "hmix", "kmix", "Tref", "uref", "vref", "Ssurf", "Sref", "SSref",
    "f", "ocdepth", "jerlov", "l_ocean", "run_physics", "sflux", "U_init", "relax_sst", "SST0", "fcorr_twod",
    "fcorr", "relax_sal", "relax_ocnT", "sal_clim", "ocnT_clim", "fcorr_withz", "sfcorr_withz", "bottom_temp",
    "nmodeadv", "modeadv", "advection", "freeze_flag", "reset_flag", "dampu_flag", "dampv_flag", "rho", "cp",
    "buoy", "Rig", "dbloc", "Shsq", "difm", "difs", "dift", "ghat", "wU", "wX", "wXNT", "tinc_fcorr", "sinc_fcorr",
    "ocnTcorr", "scorr", "swfrac", "swdk_opt", "diag_iter", "diag_nreint", "diag_status", "diag_talpha",
    "diag_sbeta"};

int build_field_map(kpp_handle *h)
{
    const long nz = h->d.nz, nzp1 = nz + 1, nzt = h->d.nztmax, nztt = nzt + 1, mm = h->d.maxmodeadv;
    KppDevArgs &a = h->a;
    h->map.assign(KPP_F__COUNT, FieldMap{});
    auto set = [&](int id, int elem, long host_rows, int ncomp, long hcr, long hr0, long rows, long dr0, long dcr,
                   long drows, void **dev) {
        h->map[id] = FieldMap{kFieldNames[id], elem, host_rows, ncomp, hcr, hr0, rows, dr0, dcr, drows, dev};
    };
    auto whole = [&](int id, int elem, long rows, void **dev) { set(id, elem, rows, 1, 0, 0, rows, 0, 0, rows, dev); };
#define P(x) ((void **)&(x))
    whole(KPP_F_U, 8, 2 * nzp1, P(a.U));
    whole(KPP_F_X, 8, 2 * nzp1, P(a.X));
    whole(KPP_F_US, 8, 4 * nzp1, P(a.Us));
    whole(KPP_F_XS, 8, 4 * nzp1, P(a.Xs));
    whole(KPP_F_HMIXD, 8, 2, P(a.hmixd));
    whole(KPP_F_OLD, 4, 1, P(a.old_));
    whole(KPP_F_NEW, 4, 1, P(a.new_));
    whole(KPP_F_HMIX, 8, 1, P(a.hmix));
    whole(KPP_F_KMIX, 8, 1, P(a.kmix));
    whole(KPP_F_TREF, 8, 1, P(a.Tref));
    whole(KPP_F_UREF, 8, 1, P(a.uref));
    whole(KPP_F_VREF, 8, 1, P(a.vref));
    whole(KPP_F_SSURF, 8, 1, P(a.Ssurf));
    whole(KPP_F_SREF, 8, 1, P(h->Sref));
    whole(KPP_F_SSREF, 8, 1, P(h->SSref));
    whole(KPP_F_F, 8, 1, P(h->f));
    whole(KPP_F_OCDEPTH, 8, 1, P(h->ocdepth));
    whole(KPP_F_JERLOV, 4, 1, P(h->jerlov));
    whole(KPP_F_L_OCEAN, 4, 1, P(h->l_ocean));
    whole(KPP_F_RUN_PHYSICS, 4, 1, P(h->run_physics));
    // sflux(npts,nsflxs,5,0:njdt): rows (1:6,5,0) -> host row (0*5+4)*nsflxs
    set(KPP_F_SFLUX, 8, (long)h->d.nsflxs * 5 * (h->d.njdt + 1), 1, 0, 4L * h->d.nsflxs, 6, 0, 0, 6, P(h->sflux));
    whole(KPP_F_U_INIT, 8, 2 * nzp1, P(h->U_init));
    whole(KPP_F_RELAX_SST, 8, 1, P(h->relax_sst));
    whole(KPP_F_SST0, 8, 1, P(h->SST0));
    whole(KPP_F_FCORR_TWOD, 8, 1, P(h->fcorr_twod));
    whole(KPP_F_FCORR, 8, 1, P(a.fcorr));
    whole(KPP_F_RELAX_SAL, 8, 1, P(h->relax_sal));
    whole(KPP_F_RELAX_OCNT, 8, 1, P(h->relax_ocnT));
    whole(KPP_F_SAL_CLIM, 8, nzp1, P(h->sal_clim));
    whole(KPP_F_OCNT_CLIM, 8, nzp1, P(h->ocnT_clim));
    whole(KPP_F_FCORR_WITHZ, 8, nzp1, P(h->fcorr_withz));
    whole(KPP_F_SFCORR_WITHZ, 8, nzp1, P(h->sfcorr_withz));
    whole(KPP_F_BOTTOM_TEMP, 8, 1, P(h->bottom_temp));
    set(KPP_F_NMODEADV, 4, 2, 1, 0, 1, 1, 0, 0, 1, P(h->nmodeadv));              // (:,2)
    set(KPP_F_MODEADV, 4, 2 * mm, 1, 0, mm, mm, 0, 0, mm, P(h->modeadv));         // (:,:,2)
    set(KPP_F_ADVECTION, 8, 2 * mm, 1, 0, mm, mm, 0, 0, mm, P(h->advection));     // (:,:,2)
    whole(KPP_F_FREEZE_FLAG, 8, 1, P(a.freeze_flag));
    whole(KPP_F_RESET_FLAG, 8, 1, P(a.reset_flag));
    whole(KPP_F_DAMPU_FLAG, 8, 1, P(a.dampu_flag));
    whole(KPP_F_DAMPV_FLAG, 8, 1, P(a.dampv_flag));
    set(KPP_F_RHO, 8, nztt + 1, 1, 0, 0, nzp1 + 1, 0, 0, nzp1 + 1, P(a.rho));    // (0:nzp1tmax) rows 0:nzp1
    set(KPP_F_CP, 8, nztt + 1, 1, 0, 0, nzp1 + 1, 0, 0, nzp1 + 1, P(a.cp));
    set(KPP_F_BUOY, 8, nztt, 1, 0, 0, nzp1, 0, 0, nzp1, P(a.buoy));              // (nzp1tmax) rows 1:nzp1
    set(KPP_F_RIG, 8, nzp1, 1, 0, 0, nz, 0, 0, nzp1, P(a.Rig));                  // rows 1:nz
    whole(KPP_F_DBLOC, 8, nz, P(a.dbloc));
    set(KPP_F_SHSQ, 8, nzp1, 1, 0, 0, nz, 0, 0, nzp1, P(a.Shsq));
    set(KPP_F_DIFM, 8, nzt + 1, 1, 0, 0, nzp1 + 1, 0, 0, nzp1 + 1, P(a.difm));   // (0:nztmax) rows 0:nzp1
    set(KPP_F_DIFS, 8, nzt + 1, 1, 0, 0, nzp1 + 1, 0, 0, nzp1 + 1, P(a.difs));
    set(KPP_F_DIFT, 8, nzt + 1, 1, 0, 0, nzp1 + 1, 0, 0, nzp1 + 1, P(a.dift));
    set(KPP_F_GHAT, 8, nzt, 1, 0, 0, nz, 0, 0, nz, P(a.ghat));                   // (nztmax) rows 1:nz
    set(KPP_F_WU, 8, 3 * (nzt + 1), 2, nzt + 1, 0, nz + 1, 0, nz + 1, 2 * (nz + 1), P(a.wU));
    set(KPP_F_WX, 8, 3 * (nzt + 1), 3, nzt + 1, 0, nz + 1, 0, nz + 1, 3 * (nz + 1), P(a.wX));
    set(KPP_F_WXNT, 8, 2 * (nzt + 1), 1, nzt + 1, 0, nz + 1, 0, nz + 1, nz + 1, P(a.wXNT));
    whole(KPP_F_TINC_FCORR, 8, nzp1, P(a.tinc_fcorr));
    whole(KPP_F_SINC_FCORR, 8, nzp1, P(a.sinc_fcorr));
    whole(KPP_F_OCNTCORR, 8, nzp1, P(a.ocnTcorr));
    whole(KPP_F_SCORR, 8, nzp1, P(a.scorr));
    whole(KPP_F_SWFRAC, 8, nzp1, P(a.swfrac));
    whole(KPP_F_SWDK_OPT, 8, nz + 1, P(a.swdk_opt));
    whole(KPP_F_DIAG_ITER, 4, 1, P(a.diag_iter));
    whole(KPP_F_DIAG_NREINT, 4, 1, P(a.diag_nreint));
    whole(KPP_F_DIAG_STATUS, 4, 1, P(a.diag_status));
    whole(KPP_F_DIAG_TALPHA, 8, nzp1 + 1, P(a.talpha));
    whole(KPP_F_DIAG_SBETA, 8, nzp1 + 1, P(a.sbeta));
#undef P
    for (int id = 0; id < KPP_F__COUNT; id++) {
        FieldMap &m = h->map[id];
        if (!m.dev) return fail(h, KPP_E_INVALID, std::string("internal: unmapped field ") + kFieldNames[id]);
        char *p = nullptr;
        int rc = dev_alloc(h, &p, (size_t)m.dev_rows * (size_t)h->ld * (size_t)m.elem);
        if (rc) return rc;
        *m.dev = p;
    }
    return 0;
}

void link_const_args(kpp_handle *h)
{
    KppDevArgs &a = h->a;
    a.U_init = h->U_init; a.Sref = h->Sref; a.SSref = h->SSref; a.f = h->f; a.ocdepth = h->ocdepth;
    a.sflux = h->sflux; a.relax_sst = h->relax_sst; a.SST0 = h->SST0; a.fcorr_twod = h->fcorr_twod;
    a.relax_sal = h->relax_sal; a.relax_ocnT = h->relax_ocnT; a.sal_clim = h->sal_clim; a.ocnT_clim = h->ocnT_clim;
    a.fcorr_withz = h->fcorr_withz; a.sfcorr_withz = h->sfcorr_withz; a.bottom_temp = h->bottom_temp;
    a.advection = h->advection; a.jerlov = h->jerlov; a.l_ocean = h->l_ocean; a.run_physics = h->run_physics;
    a.nmodeadv = h->nmodeadv; a.modeadv = h->modeadv;
}

// a multi-GPU group handle forwards `call` (written in terms of a handle `p`) to every part
#define FANOUT(h, call)                                                  \
    if ((h) && !(h)->parts.empty()) {                                    \
        for (kpp_handle * p : (h)->parts) {                              \
            const int rc_ = (call);                                      \
            if (rc_) { (h)->err = p->err; return rc_; }                  \
        }                                                                \
        return KPP_OK;                                                   \
    }

int lag_join(kpp_handle *h);

// wait for everything queued on the handle's stream(s); no report, no status side effects
int wait_streams(kpp_handle *h)
{
    if (!h->parts.empty()) {
        for (kpp_handle *p : h->parts) {
            const int rc = wait_streams(p);
            if (rc) { h->err = p->err; return rc; }
        }
        return KPP_OK;
    }
    CU(cudaSetDevice(h->device));
    if (h->lag.pending) { const int rc = lag_join(h); if (rc) return rc; }
    CU(cudaStreamSynchronize(h->stream));
    return KPP_OK;
}

// ---------------------------------------------------------------- asynchronous stragglers
// A few columns of a large run stop converging and iterate to itermax = 200 passes; the cooperative
// kernel needs ~6 ms for such a column, serially, whatever the size of the domain.  Synchronously
// that is +6 ms on every step that has one (on a GPU that owns 1/8 of the grid more than the step
// itself).  A column's step n+1 only depends on its OWN step n, so with this mode on
//   stream A  runs the step kernel of step n for every column that is not in the lane;
//   stream B  finishes step n for the columns the step kernel handed over (they join the lane);
//   stream C  runs step n in the cooperative kernel, from pass 0, for the columns in the lane
// and step n+1 starts on A while B and C are still busy.  Every column does exactly the same
// arithmetic in the same order as before: results do not change by a bit.  Whatever reads or writes
// device state through this API (downloads, uploads, outputs, kpp_gpu_sync) first JOINS: A waits for
// B and C, the report of the last step is taken, and the lane is emptied (its columns go back to the
// step kernel).  A host that syncs after every step therefore sees today's behaviour; one that queues
// steps (device-resident forcing slots, outputs through the ring once a day) hides the stragglers.
int lag_join(kpp_handle *h)
{
    kpp_handle::Lag &L = h->lag;
    if (!L.pending) return KPP_OK;
    CU(cudaSetDevice(h->device));
    const int pl = (int)((L.k - 1) & 1);
    CU(cudaStreamWaitEvent(h->stream, L.ev_fin[pl], 0));
    CU(cudaStreamWaitEvent(h->stream, L.ev_lane, 0));
    KppDevArgs a = h->a;
    a.cont_count = L.last_count;
    cudaError_t e = h->k.numerics ? kpp_launch_report_fast(&a, h->rep_dev, h->stream) : kpp_launch_report_strict(&a, h->rep_dev, h->stream);
    if (e != cudaSuccess) return fail(h, KPP_E_CUDA, std::string("report launch: ") + cudaGetErrorString(e));
    h->launches += 1;
    CU(cudaEventRecord(h->ev1, h->stream));
    CU(cudaMemcpyAsync(h->rep_host, h->rep_dev, sizeof(KppReportDev), cudaMemcpyDeviceToHost, h->stream));
    // empty the lane: every column goes back to the step kernel
    CU(cudaMemsetAsync(L.in_lane, 0, (size_t)h->ld * sizeof(int), h->stream));
    CU(cudaMemsetAsync(L.lane_count, 0, 3 * sizeof(int), h->stream));
    CU(cudaEventRecord(L.ev_reset, h->stream));
    L.pending = false;
    L.k = 0;
    if (!L.trace.empty()) {     // KPP_LAG_TRACE: when each stream's kernel of each step became eligible / finished
        CU(cudaStreamSynchronize(h->stream));
        CU(cudaStreamSynchronize(L.sB));
        CU(cudaStreamSynchronize(L.sC));
        fprintf(stderr, "lag trace (ms): step  main begin end | lane begin end | finish begin end | lane columns, handed over\n");
        for (size_t i = 0; i + 5 < L.trace.size(); i += 6) {
            float t[6];
            for (int j = 0; j < 6; j++) cudaEventElapsedTime(&t[j], L.trace[0], L.trace[i + j]);
            fprintf(stderr, "lag trace: %3zu  %8.3f %8.3f | %8.3f %8.3f | %8.3f %8.3f | %d %d\n", i / 6, t[0], t[1], t[2], t[3], t[4], t[5],
                    L.trace_counts[2 * (i / 6) + 1], L.trace_counts[2 * (i / 6)]);
        }
        for (cudaEvent_t e : L.trace) cudaEventDestroy(e);
        L.trace.clear();
    }
    return KPP_OK;
}

int step_lagged(kpp_handle *h, int ntime)
{
    kpp_handle::Lag &L = h->lag;
    const bool fast = h->k.numerics != 0;
    const long k = L.k;
    const int p = (int)(k & 1), l = (int)(k % 3), ln = (int)((k + 1) % 3);
    cudaStream_t A = h->stream, B = L.sB, C = L.sC;
    h->a.ntime = ntime;
    h->last_ntime = ntime;
    KppDevArgs a = h->a;
    a.in_lane = L.in_lane;
    // ---- A: the step kernel (the hand-over buffers of parity p were last read by the finish of step k-2)
    if (k == 0) CU(cudaEventRecord(h->ev0, A));
    if (k >= 2) CU(cudaStreamWaitEvent(A, L.ev_fin[p], 0));
    a.cont_list = L.cont_list2[p];
    a.cont_count = L.cont_count2 + p;
    a.lane_out_list = nullptr; a.lane_out_count = nullptr;
    CU(cudaMemsetAsync(a.cont_count, 0, sizeof(int), A));
    static const bool trace = getenv("KPP_LAG_TRACE") != nullptr;
    cudaEvent_t tr[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    int *tc = nullptr;
    if (trace && L.trace.size() < 6 * 256) {
        if (!L.trace_counts) CU(cudaHostAlloc((void **)&L.trace_counts, 2 * 256 * sizeof(int), cudaHostAllocDefault));
        tc = L.trace_counts + 2 * (L.trace.size() / 6);
        for (int i = 0; i < 6; i++) { CU(cudaEventCreate(&tr[i])); L.trace.push_back(tr[i]); }
    }
    if (tr[0]) CU(cudaEventRecord(tr[0], A));
    cudaError_t e = fast ? kpp_launch_main_fast(&a, A) : kpp_launch_main_strict(&a, A);
    if (e != cudaSuccess) return fail(h, KPP_E_CUDA, std::string("step launch: ") + cudaGetErrorString(e));
    CU(cudaEventRecord(L.ev_main[p], A));
    if (tr[1]) CU(cudaEventRecord(tr[1], A));
    // ---- C: the lane's own step (list l: appended by the lane of step k-1 and the finish of step k-1)
    if (k == 0) CU(cudaStreamWaitEvent(C, L.ev_reset, 0));
    CU(cudaMemsetAsync(L.lane_count + ln, 0, sizeof(int), C));
    CU(cudaEventRecord(L.ev_zero[l], C));
    if (k >= 1) CU(cudaStreamWaitEvent(C, L.ev_fin[(k - 1) & 1], 0));
    KppDevArgs aL = a;
    aL.cont_list = L.lane_list[l]; aL.cont_count = L.lane_count + l;
    aL.lane_out_list = L.lane_list[ln]; aL.lane_out_count = L.lane_count + ln;
    aL.cont_next = L.next_lane;
    if (tr[2]) CU(cudaEventRecord(tr[2], C));
    e = fast ? kpp_launch_coop_fast(&aL, C) : kpp_launch_coop_strict(&aL, C);
    if (e != cudaSuccess) return fail(h, KPP_E_CUDA, std::string("lane launch: ") + cudaGetErrorString(e));
    CU(cudaEventRecord(L.ev_lane, C));
    if (tr[3]) CU(cudaEventRecord(tr[3], C));
    if (tc) CU(cudaMemcpyAsync(tc + 1, L.lane_count + l, sizeof(int), cudaMemcpyDeviceToHost, C));
    // ---- B: finish this step for the columns the step kernel handed over; they join the lane of step k+1
    if (k == 0) CU(cudaStreamWaitEvent(B, L.ev_reset, 0));
    CU(cudaStreamWaitEvent(B, L.ev_main[p], 0));
    CU(cudaStreamWaitEvent(B, L.ev_zero[l], 0));
    KppDevArgs aF = a;
    aF.lane_out_list = L.lane_list[ln]; aF.lane_out_count = L.lane_count + ln;
    aF.cont_next = L.next_fin;
    if (tr[4]) CU(cudaEventRecord(tr[4], B));
    e = fast ? kpp_launch_coop_fast(&aF, B) : kpp_launch_coop_strict(&aF, B);
    if (e != cudaSuccess) return fail(h, KPP_E_CUDA, std::string("finish launch: ") + cudaGetErrorString(e));
    CU(cudaEventRecord(L.ev_fin[p], B));
    if (tr[5]) CU(cudaEventRecord(tr[5], B));
    if (tc) CU(cudaMemcpyAsync(tc, a.cont_count, sizeof(int), cudaMemcpyDeviceToHost, B));
    L.last_count = a.cont_count;
    h->launches += 3;
    L.k = k + 1;
    L.pending = true;
    h->stepped = true;
    return KPP_OK;
}

bool lag_usable(const kpp_handle *h)
{
    return h->lag.on && h->a.pass_budget > 0 && !h->k.L_VARY_BOTTOM_TEMP;
}

// a multi-GPU group handle forwards `call` first; a plain handle joins its straggler lane
#define JOIN(h)                                               \
    if ((h) && (h)->parts.empty() && (h)->lag.pending) {      \
        const int rcj_ = lag_join(h);                         \
        if (rcj_) return rcj_;                                \
    }

int check_field(kpp_handle *h, int id, size_t bytes)
{
    if (!h) return fail(nullptr, KPP_E_INVALID, "null handle");
    if (id < 0 || id >= KPP_F__COUNT) return fail(h, KPP_E_INVALID, "bad field id");
    const FieldMap &m = h->map[id];
    const size_t want = (size_t)m.host_rows * (size_t)h->host_npts * (size_t)m.elem;
    if (bytes != want) {
        char buf[256];
        snprintf(buf, sizeof buf, "field %s: host buffer is %zu bytes, expected %zu (whole Fortran array)", m.name,
                 bytes, want);
        return fail(h, KPP_E_INVALID, buf);
    }
    return 0;
}

}  // namespace

extern "C" {

int kpp_gpu_abi_version(void) { return KPP_GPU_ABI_VERSION; }

int kpp_gpu_exp_is_host_libm(int numerics) { return numerics ? kpp_exp_is_host_libm_fast() : kpp_exp_is_host_libm_strict(); }

int kpp_gpu_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

const char *kpp_gpu_strerror(int code)
{
    switch (code) {
    case KPP_OK: return "ok";
    case KPP_E_INVALID: return "invalid argument";
    case KPP_E_CUDA: return "CUDA runtime error";
    case KPP_E_NODEVICE: return "no CUDA device (this library has no CPU fallback)";
    case KPP_E_PIVOT_ZERO: return "tridiagonal solver hit a zero pivot (reference: MCKPP_ABORT)";
    case KPP_E_NOMEM: return "out of device memory";
    default: return "unknown error";
    }
}

const char *kpp_gpu_last_error(const kpp_handle *h) { return h ? h->err.c_str() : g_err.c_str(); }

const char *kpp_gpu_field_name(int id) { return (id >= 0 && id < KPP_F__COUNT) ? kFieldNames[id] : nullptr; }

int kpp_gpu_create(const kpp_dims *dims, const kpp_consts *consts, const double *zm, const double *hm,
                   const double *dm, const double *tri, const double *wmt, const double *wst, int device,
                   kpp_handle **out)
{
    if (!dims || !consts || !zm || !hm || !dm || !tri || !wmt || !wst || !out)
        return fail(nullptr, KPP_E_INVALID, "null argument");
    *out = nullptr;
    if (dims->npts <= 0 || dims->nz < 3) return fail(nullptr, KPP_E_INVALID, "npts must be > 0 and nz >= 3");
    if (dims->nztmax < dims->nz + 1)
        return fail(nullptr, KPP_E_INVALID, "nztmax must be >= nz+1 (ocnint_mod.F90:33,153; kppmix_mod.F90:82)");
    if (dims->nsflxs < 6 || dims->njdt < 0 || dims->maxmodeadv < 1 || dims->maxmodeadv > 6)
        return fail(nullptr, KPP_E_INVALID, "nsflxs >= 6, njdt >= 0, 1 <= maxmodeadv <= 6 required");
    if (!consts->LKPP)
        return fail(nullptr, KPP_E_INVALID, "LKPP=.FALSE. is not supported: the reference leaves hmix/kmix undefined");
    if (!(consts->dto > 0)) return fail(nullptr, KPP_E_INVALID, "dto must be > 0");
    if (consts->numerics != 0 && consts->numerics != 1) return fail(nullptr, KPP_E_INVALID, "numerics must be 0 or 1");
    for (int k = 0; k <= dims->nz; k++) {
        if (!(zm[k] < 0.0) || !(hm[k] > 0.0))
            return fail(nullptr, KPP_E_INVALID, "grid: zm(k) < 0 and hm(k) > 0 required (pressure P=-zm(k) > 0)");
        if (k > 0 && !(zm[k] < zm[k - 1])) return fail(nullptr, KPP_E_INVALID, "grid: zm must decrease with k");
    }
    if (consts->L_NO_ISOTHERM && (consts->iso_bot < 2 || consts->iso_bot > dims->nz + 1))
        return fail(nullptr, KPP_E_INVALID, "iso_bot out of range");
    if (dims->nz > 1200)
        return fail(nullptr, KPP_E_INVALID, "nz > 1200: the per-level grid tables no longer fit in shared memory");
    {
        // the kernels index every field with 32-bit element offsets
        const long long ldl = ((long long)dims->npts + 31) / 32 * 32;
        if ((4LL * (dims->nz + 1) + 8) * ldl >= (1LL << 31))
            return fail(nullptr, KPP_E_INVALID, "npts*(4*nzp1) must stay below 2^31 elements per handle: split the columns over more handles");
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        cudaGetLastError();
        return fail(nullptr, KPP_E_NODEVICE, "no CUDA device; this library has no CPU fallback");
    }
    if (device < 0 || device >= ndev) return fail(nullptr, KPP_E_INVALID, "device index out of range");

    kpp_handle *h = new kpp_handle();
    h->d = *dims;
    h->k = *consts;
    h->device = device;
    h->ld = ((dims->npts + 31) / 32) * 32;
    h->rep_dev = nullptr;
    h->rep_host = nullptr;
    h->last_ntime = 0;
    h->stepped = false;
    h->launches = 0;
    h->rawflux = nullptr;
    h->stage = nullptr;
    h->guard = getenv("KPP_GUARD") && atoi(getenv("KPP_GUARD")) != 0;
    h->host_npts = dims->npts;
    h->host_col0 = 0;
    h->modeadv_dirty = false;
    h->ring = nullptr;
    memset(&h->lag, 0, sizeof(h->lag));
    for (auto &r : h->clim_rec) r[0] = r[1] = nullptr;
    memset(&h->a, 0, sizeof(h->a));
    h->stream = nullptr;
    h->ev0 = h->ev1 = nullptr;
#define CUC(call)                                                                                          \
    do {                                                                                                   \
        cudaError_t e_ = (call);                                                                           \
        if (e_ != cudaSuccess) {                                                                           \
            int rc_ = fail(h, KPP_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));            \
            g_err = h->err;                                                                                \
            kpp_gpu_destroy(h);                                                                            \
            return rc_;                                                                                    \
        }                                                                                                  \
    } while (0)
    CUC(cudaSetDevice(device));
    CUC(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    CUC(cudaEventCreate(&h->ev0));
    CUC(cudaEventCreate(&h->ev1));
    CUC(cudaMalloc((void **)&h->rep_dev, sizeof(KppReportDev)));
    CUC(cudaMemsetAsync(h->rep_dev, 0, sizeof(KppReportDev), h->stream));
    CUC(cudaMallocHost((void **)&h->rep_host, sizeof(KppReportDev)));
    memset(h->rep_host, 0, sizeof(KppReportDev));

    KppDevArgs &a = h->a;
    a.npts = dims->npts; a.ld = h->ld; a.nz = dims->nz; a.nzp1 = dims->nz + 1; a.ntime = 0;
    a.itermax = consts->itermax; a.iso_bot = consts->iso_bot; a.maxmodeadv = dims->maxmodeadv;
    a.LRI = consts->LRI; a.LDD = consts->LDD; a.L_SSref = consts->L_SSref; a.L_RELAX_SST = consts->L_RELAX_SST;
    a.L_RELAX_CALCONLY = consts->L_RELAX_CALCONLY; a.L_FCORR = consts->L_FCORR; a.L_FCORR_WITHZ = consts->L_FCORR_WITHZ;
    a.L_SFCORR = consts->L_SFCORR; a.L_SFCORR_WITHZ = consts->L_SFCORR_WITHZ; a.L_RELAX_SAL = consts->L_RELAX_SAL;
    a.L_RELAX_OCNT = consts->L_RELAX_OCNT; a.L_NO_FREEZE = consts->L_NO_FREEZE; a.L_NO_ISOTHERM = consts->L_NO_ISOTHERM;
    a.L_DAMP_CURR = consts->L_DAMP_CURR;
    a.have_clim_files = (consts->have_ocnT_file && consts->have_sal_file) ? 1 : 0;
    a.dto = consts->dto; a.grav = consts->grav; a.vonk = consts->vonk; a.sice = consts->sice;
    a.hmixtolfrac = consts->hmixtolfrac; a.iso_thresh = consts->iso_thresh;
    {
        // blmix_mod.F90:62 and bldepth_mod.F90:91, evaluated once with the host libm
        const double cstar = 5.0, cs = 98.96, epsilon = 0.1, cv = 1.6, Ricr = 0.30;
        a.cg = cstar * consts->vonk * pow(cs * consts->vonk * epsilon, 1. / 3.);
        a.Vtc = cv * sqrt(0.2 / cs / epsilon) / (consts->vonk * consts->vonk) / Ricr;
        a.uvdamp = consts->dt_uvdamp * (86400. / consts->dto);
    }
    int rc = build_tables(h, zm, hm, dm, tri, wmt, wst);
    if (!rc) rc = build_field_map(h);
    // scratch that never crosses the ABI: tile-major level records (see kpp_kernels.cu)
    if (!rc) rc = dev_alloc(h, &a.scr, (size_t)(h->ld / 32) * (size_t)(a.nzp1 + 1) * KPP_NF * 32);
    // straggler hand-over (kpp_step_kernel -> kpp_coop_kernel)
    if (!rc) rc = dev_alloc(h, &a.cont, (size_t)h->ld);
    if (!rc) rc = dev_alloc(h, &a.cont_list, (size_t)h->ld);
    if (!rc) rc = dev_alloc(h, &a.cont_count, (size_t)1);
    if (!rc) rc = dev_alloc(h, &a.cont_next, (size_t)1);
    if (!rc) rc = dev_alloc(h, &a.tile_counter, (size_t)1);
    if (rc) { g_err = h->err; kpp_gpu_destroy(h); return rc; }
    {
        // default: hand stragglers over after 6 passes; domains too small to fill the GPU with one
        // thread per column run entirely in the cooperative kernel (-1)
        int budget = dims->npts <= KPP_SMALL_DOMAIN_COLUMNS ? -1 : 6;
        if (const char *e = getenv("KPP_PASS_BUDGET")) budget = atoi(e);
        h->pass_budget_req = budget < -1 ? -1 : budget;
        // buoyancy is stored down to (expected kbl + margin); KPP_BUOY_MARGIN is for tests of the
        // recompute path (a large negative value makes the scan recompute every level)
        a.buoy_margin = 6;
        if (const char *e = getenv("KPP_BUOY_MARGIN")) a.buoy_margin = atoi(e);
        a.pass_budget = kpp_coop_fits_strict(a.nz) ? h->pass_budget_req : 0;
    }
    link_const_args(h);
    a.pivot_sticky = &h->rep_dev->pivot_sticky;
    // defaults of mckpp_allocate/initialize: jerlov = 3, l_ocean = run_physics = .TRUE., ocdepth = -10000
    {
        std::vector<int> ones(h->ld, 1), threes(h->ld, 3);
        std::vector<double> dep(h->ld, -10000.0);
        CUC(cudaMemcpyAsync(h->l_ocean, ones.data(), h->ld * 4, cudaMemcpyHostToDevice, h->stream));
        CUC(cudaMemcpyAsync(h->run_physics, ones.data(), h->ld * 4, cudaMemcpyHostToDevice, h->stream));
        CUC(cudaMemcpyAsync(h->jerlov, threes.data(), h->ld * 4, cudaMemcpyHostToDevice, h->stream));
        CUC(cudaMemcpyAsync(h->ocdepth, dep.data(), h->ld * 8, cudaMemcpyHostToDevice, h->stream));
        CUC(cudaStreamSynchronize(h->stream));
    }
#undef CUC
    *out = h;
    return KPP_OK;
}

int kpp_gpu_create_multi(const kpp_dims *dims, const kpp_consts *consts, const double *zm, const double *hm,
                         const double *dm, const double *tri, const double *wmt, const double *wst, int ngpus,
                         const int *devices, kpp_handle **out)
{
    if (!dims || !out) return fail(nullptr, KPP_E_INVALID, "null argument");
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        cudaGetLastError();
        return fail(nullptr, KPP_E_NODEVICE, "no CUDA device; this library has no CPU fallback");
    }
    if (ngpus <= 0) ngpus = ndev;
    if (!devices && ngpus > ndev) return fail(nullptr, KPP_E_INVALID, "ngpus exceeds the visible devices");
    if (dims->npts <= 0) return fail(nullptr, KPP_E_INVALID, "npts must be > 0");
    // contiguous blocks of ceil(npts/ngpus) columns, rounded up to whole 32-column tiles
    const int block = ((dims->npts + ngpus - 1) / ngpus + 31) / 32 * 32;
    kpp_handle *g = new kpp_handle();
    g->d = *dims;
    if (consts) g->k = *consts;
    g->device = -1; g->ld = 0; g->stream = nullptr; g->ev0 = g->ev1 = nullptr; g->rep_dev = nullptr; g->rep_host = nullptr;
    g->last_ntime = 0; g->stepped = false; g->launches = 0; g->rawflux = nullptr; g->stage = nullptr;
    g->host_npts = dims->npts; g->host_col0 = 0; g->modeadv_dirty = false; g->ring = nullptr; g->pass_budget_req = 0;
    g->guard = false;
    memset(&g->lag, 0, sizeof(g->lag));
    for (auto &r : g->clim_rec) r[0] = r[1] = nullptr;
    memset(&g->a, 0, sizeof(g->a));
    for (int i = 0, c0 = 0; i < ngpus && c0 < dims->npts; i++, c0 += block) {
        kpp_dims d = *dims;
        d.npts = (dims->npts - c0 < block) ? dims->npts - c0 : block;
        kpp_handle *p = nullptr;
        const int rc = kpp_gpu_create(&d, consts, zm, hm, dm, tri, wmt, wst, devices ? devices[i] : i, &p);
        if (rc) {
            const std::string e = g_err;
            kpp_gpu_destroy(g);
            return fail(nullptr, rc, e);
        }
        p->host_npts = dims->npts;
        p->host_col0 = c0;
        g->parts.push_back(p);
    }
    *out = g;
    return KPP_OK;
}

int kpp_gpu_num_parts(const kpp_handle *h) { return h ? (h->parts.empty() ? 1 : (int)h->parts.size()) : 0; }

int kpp_gpu_part_columns(const kpp_handle *h, int part, int *device, int *col0, int *ncols)
{
    if (!h) return KPP_E_INVALID;
    const kpp_handle *p = h->parts.empty() ? (part == 0 ? h : nullptr) : (part >= 0 && part < (int)h->parts.size() ? h->parts[part] : nullptr);
    if (!p) return KPP_E_INVALID;
    if (device) *device = p->device;
    if (col0) *col0 = p->host_col0;
    if (ncols) *ncols = p->d.npts;
    return KPP_OK;
}

int kpp_gpu_destroy(kpp_handle *h)
{
    if (!h) return KPP_OK;
    kpp_gpu_output_ring_destroy(h);
    if (!h->parts.empty()) {
        for (kpp_handle *p : h->parts) kpp_gpu_destroy(p);
        delete h;
        return KPP_OK;
    }
    cudaSetDevice(h->device);
    if (h->lag.sB) cudaStreamSynchronize(h->lag.sB);
    if (h->lag.sC) cudaStreamSynchronize(h->lag.sC);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->lag.sB) {
        for (int i = 0; i < 2; i++) { cudaEventDestroy(h->lag.ev_main[i]); cudaEventDestroy(h->lag.ev_fin[i]); }
        for (int i = 0; i < 3; i++) cudaEventDestroy(h->lag.ev_zero[i]);
        cudaEventDestroy(h->lag.ev_lane); cudaEventDestroy(h->lag.ev_reset);
        cudaStreamDestroy(h->lag.sB); cudaStreamDestroy(h->lag.sC);
    }
    for (void *p : h->allocs) cudaFree(p);
    if (h->rep_dev) cudaFree(h->rep_dev);
    if (h->rep_host) cudaFreeHost(h->rep_host);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->stream) cudaStreamDestroy(h->stream);
    cudaGetLastError();
    delete h;
    return KPP_OK;
}

size_t kpp_gpu_field_host_bytes(const kpp_handle *h, int id)
{
    if (!h || id < 0 || id >= KPP_F__COUNT) return 0;
    if (!h->parts.empty()) return kpp_gpu_field_host_bytes(h->parts[0], id);
    const FieldMap &m = h->map[id];
    return (size_t)m.host_rows * (size_t)h->host_npts * (size_t)m.elem;
}

static int move_field(kpp_handle *h, int id, void *host, bool to_device)
{
    const FieldMap &m = h->map[id];
    const size_t wbytes = (size_t)h->d.npts * m.elem, hpitch = (size_t)h->host_npts * m.elem, dpitch = (size_t)h->ld * m.elem;
    CU(cudaSetDevice(h->device));
    for (int cidx = 0; cidx < m.ncomp; cidx++) {
        char *hp = (char *)host + (size_t)(m.host_row0 + cidx * m.host_comp_rows) * hpitch + (size_t)h->host_col0 * m.elem;
        char *dp = (char *)(*m.dev) + (size_t)(m.dev_row0 + cidx * m.dev_comp_rows) * dpitch;
        if (to_device)
            CU(cudaMemcpy2DAsync(dp, dpitch, hp, hpitch, wbytes, (size_t)m.rows, cudaMemcpyHostToDevice, h->stream));
        else
            CU(cudaMemcpy2DAsync(hp, hpitch, dp, dpitch, wbytes, (size_t)m.rows, cudaMemcpyDeviceToHost, h->stream));
    }
    return KPP_OK;
}

int kpp_gpu_upload_field(kpp_handle *h, int id, const void *host, size_t bytes)
{
    FANOUT(h, kpp_gpu_upload_field(p, id, host, bytes));
    JOIN(h);
    int rc = check_field(h, id, bytes);
    if (rc) return rc;
    if (!host) return fail(h, KPP_E_INVALID, "null host buffer");
    const size_t n = (size_t)h->d.npts, hn = (size_t)h->host_npts, c0 = (size_t)h->host_col0;
    if (id == KPP_F_JERLOV) {
        // jerlov indexes the five water types of swfrac / swdk (swfrac_mod.F90:28-34): anything else
        // would read outside the reference's rfac/a1/a2 arrays (and this library's tables)
        const int32_t *jp = (const int32_t *)host + c0;
        for (size_t i = 0; i < n; i++)
            if (jp[i] < 1 || jp[i] > 5) return fail(h, KPP_E_INVALID, "jerlov: water type must be 1..5 (swfrac_mod.F90:28-34)");
    }
    if (id == KPP_F_NMODEADV || id == KPP_F_MODEADV) {
        // 'mode out of range' is fatal in the reference (solvers.F90:320-324) -- but only for the
        // nmodeadv(:,2) entries rhsmod actually visits; the slots beyond are ALLOCATEd and never
        // initialised.  Keep host copies and check at step time, when both members are known.
        const int mm = h->d.maxmodeadv;
        if (id == KPP_F_NMODEADV) {
            const int32_t *np_ = (const int32_t *)host + hn + c0;                 // (:,2)
            h->nmodeadv_host.assign(np_, np_ + n);
        } else {
            h->modeadv_host.resize((size_t)mm * n);
            for (int im = 0; im < mm; im++) {
                const int32_t *mp = (const int32_t *)host + ((size_t)mm + im) * hn + c0;   // (:,im,2)
                memcpy(&h->modeadv_host[(size_t)im * n], mp, n * sizeof(int32_t));
            }
        }
        h->modeadv_dirty = true;
    }
    rc = move_field(h, id, (void *)host, true);
    if (rc) return rc;
    // pageable host memory: the runtime has staged the data when the call returns
    return KPP_OK;
}

int kpp_gpu_download_field(kpp_handle *h, int id, void *host, size_t bytes)
{
    if (h && !h->parts.empty()) {
        // enqueue on every device first, then wait: the parts' copies run concurrently
        int rc_ = kpp_gpu_download_field_async(h, id, host, bytes);
        if (rc_) return rc_;
        return wait_streams(h);
    }
    JOIN(h);
    int rc = check_field(h, id, bytes);
    if (rc) return rc;
    if (!host) return fail(h, KPP_E_INVALID, "null host buffer");
    rc = move_field(h, id, host, false);
    if (rc) return rc;
    CU(cudaStreamSynchronize(h->stream));
    return KPP_OK;
}

int kpp_gpu_download_field_async(kpp_handle *h, int id, void *host, size_t bytes)
{
    FANOUT(h, kpp_gpu_download_field_async(p, id, host, bytes));
    JOIN(h);
    int rc = check_field(h, id, bytes);
    if (rc) return rc;
    if (!host) return fail(h, KPP_E_INVALID, "null host buffer");
    return move_field(h, id, host, false);
}

int kpp_gpu_upload_forcing(kpp_handle *h, const double *sflux6)
{
    if (!h || !sflux6) return fail(h, KPP_E_INVALID, "null argument");
    FANOUT(h, kpp_gpu_upload_forcing(p, sflux6));
    JOIN(h);
    CU(cudaSetDevice(h->device));
    const size_t wbytes = (size_t)h->d.npts * 8;
    CU(cudaMemcpy2DAsync(h->sflux, (size_t)h->ld * 8, sflux6 + h->host_col0, (size_t)h->host_npts * 8, wbytes, 6,
                         cudaMemcpyHostToDevice, h->stream));
    return KPP_OK;
}

int kpp_gpu_upload_fluxes(kpp_handle *h, const double *taux, const double *tauy, const double *swf, const double *lwf,
                          const double *lhf, const double *shf, const double *rain, const double *snow, double flsn,
                          double el)
{
    if (!h || !taux || !tauy || !swf || !lwf || !lhf || !shf || !rain || !snow) return fail(h, KPP_E_INVALID, "null argument");
    FANOUT(h, kpp_gpu_upload_fluxes(p, taux, tauy, swf, lwf, lhf, shf, rain, snow, flsn, el));
    JOIN(h);
    CU(cudaSetDevice(h->device));
    if (!h->rawflux) {
        int rc = dev_alloc(h, &h->rawflux, (size_t)8 * h->ld);
        if (rc) return rc;
    }
    const double *src[8] = {taux, tauy, swf, lwf, lhf, shf, rain, snow};
    for (int i = 0; i < 8; i++)
        CU(cudaMemcpyAsync(h->rawflux + (size_t)i * h->ld, src[i] + h->host_col0, (size_t)h->d.npts * 8, cudaMemcpyHostToDevice, h->stream));
    // the map itself is IEEE +,-,*,/ only: one variant serves both numerics (no contraction: -fmad=false TU)
    cudaError_t e = kpp_launch_fluxmap_strict(h->d.npts, h->ld, h->rawflux, h->l_ocean, flsn, el, h->sflux, h->stream);
    if (e != cudaSuccess) return fail(h, KPP_E_CUDA, std::string("fluxmap launch: ") + cudaGetErrorString(e));
    h->launches += 1;
    h->a.sflux = h->sflux;
    return KPP_OK;
}

int kpp_gpu_reserve_forcing_slots(kpp_handle *h, int nslots)
{
    if (!h || nslots < 0) return fail(h, KPP_E_INVALID, "bad argument");
    FANOUT(h, kpp_gpu_reserve_forcing_slots(p, nslots));
    CU(cudaSetDevice(h->device));
    while ((int)h->slots.size() < nslots) {
        double *p = nullptr;
        int rc = dev_alloc(h, &p, (size_t)6 * h->ld);
        if (rc) return rc;
        h->slots.push_back(p);
    }
    return KPP_OK;
}

int kpp_gpu_upload_forcing_slot(kpp_handle *h, int slot, const double *sflux6)
{
    if (!h || !sflux6) return fail(h, KPP_E_INVALID, "null argument");
    FANOUT(h, kpp_gpu_upload_forcing_slot(p, slot, sflux6));
    JOIN(h);
    if (slot < 0 || slot >= (int)h->slots.size()) return fail(h, KPP_E_INVALID, "bad slot");
    CU(cudaSetDevice(h->device));
    const size_t wbytes = (size_t)h->d.npts * 8;
    CU(cudaMemcpy2DAsync(h->slots[slot], (size_t)h->ld * 8, sflux6 + h->host_col0, (size_t)h->host_npts * 8, wbytes, 6,
                         cudaMemcpyHostToDevice, h->stream));
    return KPP_OK;
}

int kpp_gpu_select_forcing_slot(kpp_handle *h, int slot)
{
    if (!h) return fail(h, KPP_E_INVALID, "null handle");
    FANOUT(h, kpp_gpu_select_forcing_slot(p, slot));
    if (slot < -1 || slot >= (int)h->slots.size()) return fail(h, KPP_E_INVALID, "bad slot");
    h->a.sflux = (slot < 0) ? h->sflux : h->slots[slot];
    return KPP_OK;
}

long long kpp_gpu_launch_count(const kpp_handle *h)
{
    if (!h) return 0;
    long long n = h->launches;
    for (const kpp_handle *p : h->parts) n += p->launches;
    return n;
}

int kpp_gpu_init_vmix(kpp_handle *h)
{
    if (!h) return fail(h, KPP_E_INVALID, "null handle");
    FANOUT(h, kpp_gpu_init_vmix(p));
    JOIN(h);
    CU(cudaSetDevice(h->device));
    h->a.ntime = 0;
    cudaError_t e = h->k.numerics ? kpp_launch_init_fast(&h->a, h->stream) : kpp_launch_init_strict(&h->a, h->stream);
    if (e != cudaSuccess) return fail(h, KPP_E_CUDA, std::string("init launch: ") + cudaGetErrorString(e));
    h->launches += 1;
    return KPP_OK;
}

int kpp_gpu_step(kpp_handle *h, int ntime)
{
    if (!h) return fail(h, KPP_E_INVALID, "null handle");
    FANOUT(h, kpp_gpu_step(p, ntime));
    if (h->modeadv_dirty) {
        // solvers.F90:320-324: an advection mode outside 1..7 is fatal -- for the entries rhsmod visits
        const size_t n = (size_t)h->d.npts;
        if (h->nmodeadv_host.size() == n && h->modeadv_host.size() == n * (size_t)h->d.maxmodeadv)
            for (size_t c = 0; c < n; c++)
                for (int im = 0; im < h->nmodeadv_host[c] && im < h->d.maxmodeadv; im++)
                    if (h->modeadv_host[(size_t)im * n + c] > 7)
                        return fail(h, KPP_E_INVALID, "modeadv: mode out of range (solvers.F90:320)");
        h->modeadv_dirty = false;
    }
    CU(cudaSetDevice(h->device));
    if (lag_usable(h)) return step_lagged(h, ntime);
    JOIN(h);
    h->a.ntime = ntime;
    h->last_ntime = ntime;
    CU(cudaEventRecord(h->ev0, h->stream));
    cudaError_t e = h->k.numerics ? kpp_launch_step_fast(&h->a, h->rep_dev, h->k.L_VARY_BOTTOM_TEMP, h->stream)
                                  : kpp_launch_step_strict(&h->a, h->rep_dev, h->k.L_VARY_BOTTOM_TEMP, h->stream);
    if (e != cudaSuccess) return fail(h, KPP_E_CUDA, std::string("step launch: ") + cudaGetErrorString(e));
    h->launches += 2 + (h->k.L_VARY_BOTTOM_TEMP ? 1 : 0) + (h->a.pass_budget != 0 ? 1 : 0);
    CU(cudaEventRecord(h->ev1, h->stream));
    CU(cudaMemcpyAsync(h->rep_host, h->rep_dev, sizeof(KppReportDev), cudaMemcpyDeviceToHost, h->stream));
    h->stepped = true;
    return KPP_OK;
}

}  // extern "C"

// ---------------------------------------------------------------- output packing
namespace {
const char *const kOutNames[KPP_OUT__COUNT] = {
    "u", "v", "T", "S", "B", "wu", "wv", "wT", "wS", "wB", "wTnt", "difm", "dift", "difs", "rho", "cp", "scorr", "Rig",
    "dbloc", "Shsq", "tinc_fcorr", "fcorr_z", "sinc_fcorr", "hmix", "fcorr", "taux_in", "tauy_in", "solar_in",
    "nsolar_in", "PminusE_in", "freeze_flag", "comp_flag", "dampu_flag", "dampv_flag", "uvel", "vvel", "T", "S", "CP",
    "rho", "hmix", "kmix", "Sref", "SSref", "Ssurf", "Tref", "old", "new", "Us", "Vs", "Ts", "Ss", "hmixd"};

struct OutSeg {
    const void *src;
    int is_int;
    long src_row0, nrows, dst_row0;
};
struct OutDesc {
    long rows;            // rows of the packed block; rows no segment covers are zero
    int nseg;
    OutSeg seg[2];
    const double *addvec;
};

// where every output comes from in the device layout (row conventions: build_field_map)
bool out_desc(const kpp_handle *h, int id, OutDesc &o)
{
    const KppDevArgs &a = h->a;
    const long nz = h->d.nz, nzp1 = nz + 1;
    o = OutDesc{};
    auto one = [&](const void *src, long row0, long nrows, long rows, long dst0 = 0, int is_int = 0) {
        o.rows = rows; o.nseg = 1; o.seg[0] = OutSeg{src, is_int, row0, nrows, dst0};
    };
    auto two_levels = [&](const double *src, long comp) {   // (:,:,comp,0:1) of Us/Xs
        o.rows = 2 * nzp1; o.nseg = 2;
        o.seg[0] = OutSeg{src, 0, (0 * 2 + comp) * nzp1, nzp1, 0};
        o.seg[1] = OutSeg{src, 0, (1 * 2 + comp) * nzp1, nzp1, nzp1};
    };
    switch (id) {
    case KPP_OUT_U: case KPP_OUT_R_UVEL: one(a.U, 0, nzp1, nzp1); break;
    case KPP_OUT_V: case KPP_OUT_R_VVEL: one(a.U, nzp1, nzp1, nzp1); break;
    case KPP_OUT_T: case KPP_OUT_R_T: one(a.X, 0, nzp1, nzp1); break;
    case KPP_OUT_S: one(a.X, nzp1, nzp1, nzp1); o.addvec = h->Sref; break;
    case KPP_OUT_R_S: one(a.X, nzp1, nzp1, nzp1); break;
    case KPP_OUT_B: one(a.buoy, 0, nzp1, nzp1); break;
    case KPP_OUT_WU: one(a.wU, 0, nz + 1, nzp1); break;
    case KPP_OUT_WV: one(a.wU, nz + 1, nz + 1, nzp1); break;
    case KPP_OUT_WT: one(a.wX, 0, nz + 1, nzp1); break;
    case KPP_OUT_WS: one(a.wX, nz + 1, nz + 1, nzp1); break;
    case KPP_OUT_WB: one(a.wX, 2 * (nz + 1), nz + 1, nzp1); break;
    case KPP_OUT_WTNT: one(a.wXNT, 0, nz + 1, nzp1); break;
    case KPP_OUT_DIFM: one(a.difm, 1, nz, nzp1, 1); break;
    case KPP_OUT_DIFT: one(a.dift, 1, nz, nzp1, 1); break;
    case KPP_OUT_DIFS: one(a.difs, 1, nz, nzp1, 1); break;
    case KPP_OUT_RHO: case KPP_OUT_R_RHO: one(a.rho, 1, nzp1, nzp1); break;
    case KPP_OUT_CP: case KPP_OUT_R_CP: one(a.cp, 1, nzp1, nzp1); break;
    case KPP_OUT_SCORR: one(a.scorr, 0, nzp1, nzp1); break;
    case KPP_OUT_RIG: one(a.Rig, 0, nz, nzp1); break;
    case KPP_OUT_DBLOC: one(a.dbloc, 0, nz, nzp1); break;
    case KPP_OUT_SHSQ: one(a.Shsq, 0, nz, nzp1); break;
    case KPP_OUT_TINC_FCORR: one(a.tinc_fcorr, 0, nzp1, nzp1); break;
    case KPP_OUT_FCORR_Z: one(a.ocnTcorr, 0, nzp1, nzp1); break;
    case KPP_OUT_SINC_FCORR: one(a.sinc_fcorr, 0, nzp1, nzp1); break;
    case KPP_OUT_HMIX: case KPP_OUT_R_HMIX: one(a.hmix, 0, 1, 1); break;
    case KPP_OUT_FCORR: one(a.fcorr, 0, 1, 1); break;
    case KPP_OUT_TAUX_IN: one(a.sflux, 0, 1, 1); break;
    case KPP_OUT_TAUY_IN: one(a.sflux, 1, 1, 1); break;
    case KPP_OUT_SOLAR_IN: one(a.sflux, 2, 1, 1); break;
    case KPP_OUT_NSOLAR_IN: one(a.sflux, 3, 1, 1); break;
    case KPP_OUT_PMINUSE_IN: one(a.sflux, 5, 1, 1); break;
    case KPP_OUT_FREEZE_FLAG: one(a.freeze_flag, 0, 1, 1); break;
    case KPP_OUT_COMP_FLAG: one(a.reset_flag, 0, 1, 1); break;
    case KPP_OUT_DAMPU_FLAG: one(a.dampu_flag, 0, 1, 1); break;
    case KPP_OUT_DAMPV_FLAG: one(a.dampv_flag, 0, 1, 1); break;
    case KPP_OUT_R_KMIX: one(a.kmix, 0, 1, 1); break;
    case KPP_OUT_R_SREF: one(h->Sref, 0, 1, 1); break;
    case KPP_OUT_R_SSREF: one(h->SSref, 0, 1, 1); break;
    case KPP_OUT_R_SSURF: one(a.Ssurf, 0, 1, 1); break;
    case KPP_OUT_R_TREF: one(a.Tref, 0, 1, 1); break;
    case KPP_OUT_R_OLD: one(a.old_, 0, 1, 1, 0, 1); break;
    case KPP_OUT_R_NEW: one(a.new_, 0, 1, 1, 0, 1); break;
    case KPP_OUT_R_US: two_levels(a.Us, 0); break;
    case KPP_OUT_R_VS: two_levels(a.Us, 1); break;
    case KPP_OUT_R_TS: two_levels(a.Xs, 0); break;
    case KPP_OUT_R_SS: two_levels(a.Xs, 1); break;
    case KPP_OUT_R_HMIXD: one(a.hmixd, 0, 2, 2); break;
    default: return false;
    }
    return true;
}
}  // namespace

extern "C" {

const char *kpp_gpu_output_name(int out_id)
{
    return (out_id >= 0 && out_id < KPP_OUT__COUNT) ? kOutNames[out_id] : "?";
}

int kpp_gpu_output_rows(const kpp_handle *h, int out_id)
{
    OutDesc o;
    if (h && !h->parts.empty()) h = h->parts[0];
    if (!h || !out_desc(h, out_id, o)) return KPP_E_INVALID;
    return (int)o.rows;
}

// pack output `out_id` of this (plain or part) handle into `dst` (dense npts x rows) on the handle's stream
static int pack_block(kpp_handle *h, int out_id, double *dst)
{
    OutDesc o;
    if (!out_desc(h, out_id, o)) return fail(h, KPP_E_INVALID, "unknown output id");
    long covered = 0;
    for (int s = 0; s < o.nseg; s++) covered += o.seg[s].nrows;
    if (covered < o.rows) CU(cudaMemsetAsync(dst, 0, (size_t)h->d.npts * (size_t)o.rows * 8, h->stream));
    for (int s = 0; s < o.nseg; s++) {
        const OutSeg &g = o.seg[s];
        cudaError_t e = kpp_launch_pack_rows_strict(h->d.npts, h->ld, g.src, g.is_int, g.src_row0, (int)g.nrows, dst,
                                                    g.dst_row0, o.addvec, h->stream);
        if (e != cudaSuccess) return fail(h, KPP_E_CUDA, std::string("pack launch: ") + cudaGetErrorString(e));
        h->launches += 1;
    }
    return KPP_OK;
}

// device block (dense npts x rows) -> this handle's column slice of the host block (host_npts x rows)
static int copy_block_to_host(kpp_handle *h, double *host, const double *dev, long rows, cudaStream_t st)
{
    const size_t w = (size_t)h->d.npts * 8;
    if (h->host_npts == h->d.npts)
        CU(cudaMemcpyAsync(host, dev, w * (size_t)rows, cudaMemcpyDeviceToHost, st));
    else
        CU(cudaMemcpy2DAsync(host + h->host_col0, (size_t)h->host_npts * 8, dev, w, w, (size_t)rows, cudaMemcpyDeviceToHost, st));
    return KPP_OK;
}

int kpp_gpu_pack_output_async(kpp_handle *h, int out_id, double *host, size_t bytes)
{
    if (!h || !host) return fail(h, KPP_E_INVALID, "null argument");
    FANOUT(h, kpp_gpu_pack_output_async(p, out_id, host, bytes));
    JOIN(h);
    OutDesc o;
    if (!out_desc(h, out_id, o)) return fail(h, KPP_E_INVALID, "unknown output id");
    const size_t want = (size_t)h->host_npts * (size_t)o.rows * 8;
    if (bytes != want)
        return fail(h, KPP_E_INVALID, std::string("output ") + kOutNames[out_id] + ": expected " + std::to_string(want) +
                                          " bytes, got " + std::to_string(bytes));
    CU(cudaSetDevice(h->device));
    if (!h->stage) {
        int rc = dev_alloc(h, &h->stage, (size_t)h->d.npts * 2 * (size_t)(h->d.nz + 1));
        if (rc) return rc;
    }
    // stream order makes the one staging block safe to reuse: the previous output's copy has
    // finished before this one's kernels start
    int rc = pack_block(h, out_id, h->stage);
    if (rc) return rc;
    return copy_block_to_host(h, host, h->stage, o.rows, h->stream);
}

int kpp_gpu_pack_output(kpp_handle *h, int out_id, double *host, size_t bytes)
{
    int rc = kpp_gpu_pack_output_async(h, out_id, host, bytes);
    if (rc) return rc;
    return wait_streams(h);
}

// ---------------------------------------------------------------- asynchronous output ring
// The host I/O layer of the reference sends its output set after EVERY physics step
// (mckpp_ocean_model_3D.F90:62 -> mckpp_xios_control.F90:52-57 -> mckpp_xios_io.F90:72-207); XIOS
// decides what is written.  Pulling that set synchronously costs a PCIe transfer per step that is
// far longer than the step itself.  The ring packs the chosen blocks into device staging on the
// step's stream (a device-to-device copy) and moves them to pinned host slots on a SECOND stream:
// the device->host copy of step n overlaps the kernels of step n+1, and a slot stays valid for
// the host until it is submitted again `depth` submits later.
static void ring_free(kpp_handle *h)
{
    kpp_handle::Ring *r = h->ring;
    if (!r) return;
    cudaSetDevice(h->device);
    if (r->io) cudaStreamSynchronize(r->io);
    for (double *d : r->dev) if (d) cudaFree(d);
    if (r->owns_host) for (double *q : r->host) if (q) cudaFreeHost(q);
    for (cudaEvent_t e : r->ev_packed) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : r->ev_done) if (e) cudaEventDestroy(e);
    if (r->io) cudaStreamDestroy(r->io);
    delete r;
    h->ring = nullptr;
}

static int ring_setup(kpp_handle *h, const int32_t *ids, int n, int depth, double *const *host_slots)
{
    ring_free(h);
    kpp_handle::Ring *r = new kpp_handle::Ring();
    h->ring = r;
    r->depth = depth; r->next = 0; r->io = nullptr; r->owns_host = (host_slots == nullptr);
    r->dev_elems = r->host_elems = 0;
    for (int i = 0; i < n; i++) {
        OutDesc o;
        if (!out_desc(h, ids[i], o)) return fail(h, KPP_E_INVALID, "output ring: unknown output id");
        r->ids.push_back(ids[i]);
        r->rows.push_back((int)o.rows);
        r->dev_off.push_back(r->dev_elems);
        r->host_off.push_back(r->host_elems);
        r->dev_elems += (size_t)h->d.npts * (size_t)o.rows;
        r->host_elems += (size_t)h->host_npts * (size_t)o.rows;
    }
    CU(cudaSetDevice(h->device));
    CU(cudaStreamCreateWithFlags(&r->io, cudaStreamNonBlocking));
    r->dev.assign(depth, nullptr); r->host.assign(depth, nullptr);
    r->ev_packed.assign(depth, nullptr); r->ev_done.assign(depth, nullptr); r->submitted.assign(depth, 0);
    for (int s = 0; s < depth; s++) {
        if (cudaMalloc((void **)&r->dev[s], (r->dev_elems ? r->dev_elems : 1) * 8) != cudaSuccess) {
            cudaGetLastError();
            return fail(h, KPP_E_NOMEM, "output ring: out of device memory for the staging slots");
        }
        if (host_slots) r->host[s] = host_slots[s];
        else if (cudaHostAlloc((void **)&r->host[s], (r->host_elems ? r->host_elems : 1) * 8, cudaHostAllocPortable) != cudaSuccess) {
            cudaGetLastError();
            return fail(h, KPP_E_NOMEM, "output ring: cannot pin the host slots");
        }
        CU(cudaEventCreateWithFlags(&r->ev_packed[s], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&r->ev_done[s], cudaEventDisableTiming));
    }
    return KPP_OK;
}

int kpp_gpu_output_ring_create(kpp_handle *h, const int32_t *out_ids, int n_ids, int depth)
{
    if (!h || !out_ids || n_ids <= 0 || depth < 1 || depth > 16) return fail(h, KPP_E_INVALID, "output ring: bad argument");
    if (h->parts.empty()) return ring_setup(h, out_ids, n_ids, depth, nullptr);
    // group: the pinned host slots hold the blocks of ALL columns; every part copies its own column slice
    int rc = ring_setup(h->parts[0], out_ids, n_ids, depth, nullptr);
    if (rc) { h->err = h->parts[0]->err; return rc; }
    for (size_t i = 1; i < h->parts.size(); i++) {
        rc = ring_setup(h->parts[i], out_ids, n_ids, depth, h->parts[0]->ring->host.data());
        if (rc) { h->err = h->parts[i]->err; return rc; }
    }
    return KPP_OK;
}

size_t kpp_gpu_output_ring_slot_bytes(const kpp_handle *h)
{
    if (h && !h->parts.empty()) h = h->parts[0];
    return (h && h->ring) ? h->ring->host_elems * 8 : 0;
}

size_t kpp_gpu_output_ring_offset(const kpp_handle *h, int index)
{
    if (h && !h->parts.empty()) h = h->parts[0];
    if (!h || !h->ring || index < 0 || index >= (int)h->ring->ids.size()) return (size_t)-1;
    return h->ring->host_off[index] * 8;
}

static int ring_submit_one(kpp_handle *h, int *slot_out)
{
    kpp_handle::Ring *r = h->ring;
    if (!r) return fail(h, KPP_E_INVALID, "output ring: not created");
    CU(cudaSetDevice(h->device));
    JOIN(h);
    const int s = r->next;
    r->next = (s + 1) % r->depth;
    // the staging slot may still be on its way to the host from `depth` submits ago
    if (r->submitted[s]) CU(cudaStreamWaitEvent(h->stream, r->ev_done[s], 0));
    for (size_t i = 0; i < r->ids.size(); i++) {
        int rc = pack_block(h, r->ids[i], r->dev[s] + r->dev_off[i]);
        if (rc) return rc;
    }
    CU(cudaEventRecord(r->ev_packed[s], h->stream));
    CU(cudaStreamWaitEvent(r->io, r->ev_packed[s], 0));
    for (size_t i = 0; i < r->ids.size(); i++) {
        int rc = copy_block_to_host(h, r->host[s] + r->host_off[i], r->dev[s] + r->dev_off[i], r->rows[i], r->io);
        if (rc) return rc;
    }
    CU(cudaEventRecord(r->ev_done[s], r->io));
    r->submitted[s] = 1;
    *slot_out = s;
    return KPP_OK;
}

int kpp_gpu_output_ring_submit(kpp_handle *h, int *slot)
{
    if (!h || !slot) return fail(h, KPP_E_INVALID, "null argument");
    if (h->parts.empty()) return ring_submit_one(h, slot);
    for (kpp_handle *p : h->parts) {
        int rc = ring_submit_one(p, slot);      // all parts advance in lock step: same slot
        if (rc) { h->err = p->err; return rc; }
    }
    return KPP_OK;
}

int kpp_gpu_output_ring_wait(kpp_handle *h, int slot, double **host)
{
    if (!h) return fail(h, KPP_E_INVALID, "null handle");
    kpp_handle *first = h->parts.empty() ? h : h->parts[0];
    if (!first->ring || slot < 0 || slot >= first->ring->depth) return fail(h, KPP_E_INVALID, "output ring: bad slot");
    const size_t np_ = h->parts.empty() ? 1 : h->parts.size();
    for (size_t i = 0; i < np_; i++) {
        kpp_handle *p = h->parts.empty() ? h : h->parts[i];
        if (!p->ring->submitted[slot]) return fail(h, KPP_E_INVALID, "output ring: slot was never submitted");
        cudaError_t e = cudaEventSynchronize(p->ring->ev_done[slot]);
        if (e != cudaSuccess) return fail(h, KPP_E_CUDA, std::string("output ring wait: ") + cudaGetErrorString(e));
    }
    if (host) *host = first->ring->host[slot];
    return KPP_OK;
}

int kpp_gpu_output_ring_destroy(kpp_handle *h)
{
    if (!h) return KPP_OK;
    // parts 1.. borrow part 0's host slots: free them first
    for (size_t i = h->parts.size(); i-- > 0;) ring_free(h->parts[i]);
    ring_free(h);
    return KPP_OK;
}

// ---------------------------------------------------------------- climatology blend
int kpp_gpu_upload_clim_record(kpp_handle *h, int id, int which, const double *record, size_t bytes)
{
    if (!h || !record) return fail(h, KPP_E_INVALID, "null argument");
    if ((id != KPP_F_OCNT_CLIM && id != KPP_F_SAL_CLIM) || which < 0 || which > 1)
        return fail(h, KPP_E_INVALID, "climatology record: id must be KPP_F_OCNT_CLIM or KPP_F_SAL_CLIM, which 0 or 1");
    FANOUT(h, kpp_gpu_upload_clim_record(p, id, which, record, bytes));
    JOIN(h);
    const size_t nzp1 = (size_t)h->d.nz + 1, npts = (size_t)h->d.npts, hn = (size_t)h->host_npts;
    if (bytes != hn * nzp1 * 8) return fail(h, KPP_E_INVALID, "climatology record: size mismatch");
    CU(cudaSetDevice(h->device));
    double *&rec = h->clim_rec[id == KPP_F_SAL_CLIM ? 1 : 0][which];
    if (!rec) {
        int rc = dev_alloc(h, &rec, (size_t)h->ld * nzp1);
        if (rc) return rc;
        CU(cudaMemsetAsync(rec, 0, (size_t)h->ld * nzp1 * 8, h->stream));   // the pad columns
    }
    CU(cudaMemcpy2DAsync(rec, (size_t)h->ld * 8, record + h->host_col0, hn * 8, npts * 8, nzp1, cudaMemcpyHostToDevice, h->stream));
    return KPP_OK;
}

int kpp_gpu_blend_clim(kpp_handle *h, int id, double prev_weight, double next_weight)
{
    if (!h) return fail(h, KPP_E_INVALID, "null handle");
    if (id != KPP_F_OCNT_CLIM && id != KPP_F_SAL_CLIM) return fail(h, KPP_E_INVALID, "blend: not a climatology field");
    FANOUT(h, kpp_gpu_blend_clim(p, id, prev_weight, next_weight));
    JOIN(h);
    const int w = id == KPP_F_SAL_CLIM ? 1 : 0;
    if (!h->clim_rec[w][0] || !h->clim_rec[w][1]) return fail(h, KPP_E_INVALID, "blend: upload both records first");
    CU(cudaSetDevice(h->device));
    double *dst = id == KPP_F_SAL_CLIM ? h->sal_clim : h->ocnT_clim;
    cudaError_t e = kpp_launch_blend_strict((size_t)h->ld * (size_t)(h->d.nz + 1), h->clim_rec[w][0], h->clim_rec[w][1],
                                            prev_weight, next_weight, dst, h->stream);
    if (e != cudaSuccess) return fail(h, KPP_E_CUDA, std::string("blend launch: ") + cudaGetErrorString(e));
    h->launches += 1;
    return KPP_OK;
}

int kpp_gpu_set_pass_budget(kpp_handle *h, int budget)
{
    if (!h) return fail(h, KPP_E_INVALID, "null handle");
    if (budget < -1) return fail(h, KPP_E_INVALID, "pass budget must be >= -1");
    FANOUT(h, kpp_gpu_set_pass_budget(p, budget));
    JOIN(h);
    h->pass_budget_req = budget;
    // columns deeper than the cooperative kernel's shared memory can hold stay with the per-thread kernel
    h->a.pass_budget = kpp_coop_fits_strict(h->a.nz) ? budget : 0;
    return KPP_OK;
}

int kpp_gpu_debug_check_guards(kpp_handle *h)
{
    if (!h) return KPP_E_INVALID;
    if (!h->parts.empty()) {
        int bad = 0;
        for (kpp_handle *p : h->parts) {
            const int b = kpp_gpu_debug_check_guards(p);
            if (b < 0) return b;
            bad += b;
        }
        return bad;
    }
    if (!h->guard) return fail(h, KPP_E_INVALID, "guard zones are off: set KPP_GUARD=1 before kpp_gpu_create");
    int rc = wait_streams(h);
    if (rc) return rc;
    std::vector<unsigned char> buf(GUARD_BYTES);
    int bad = 0;
    for (auto &g : h->guards)
        for (int side = 0; side < 2; side++) {
            CU(cudaMemcpy(buf.data(), g.first + (side ? GUARD_BYTES + g.second : 0), GUARD_BYTES, cudaMemcpyDeviceToHost));
            for (size_t i = 0; i < GUARD_BYTES; i++)
                if (buf[i] != GUARD_BYTE) { bad++; break; }
        }
    return bad;
}

int kpp_gpu_set_async_stragglers(kpp_handle *h, int on)
{
    if (!h) return fail(h, KPP_E_INVALID, "null handle");
    FANOUT(h, kpp_gpu_set_async_stragglers(p, on));
    JOIN(h);
    kpp_handle::Lag &L = h->lag;
    CU(cudaSetDevice(h->device));
    if (on && !L.sB) {
        // high priority: when SMs free up at the end of a step the few cooperative CTAs are placed before the next
        // step kernel's; that one is persistent with dynamically fetched tiles, so its CTAs simply share out the
        // rest of the machine
        int pr_lo = 0, pr_hi = 0;
        CU(cudaDeviceGetStreamPriorityRange(&pr_lo, &pr_hi));
        CU(cudaStreamCreateWithPriority(&L.sB, cudaStreamNonBlocking, pr_hi));
        CU(cudaStreamCreateWithPriority(&L.sC, cudaStreamNonBlocking, pr_hi));
        for (int i = 0; i < 2; i++) {
            CU(cudaEventCreateWithFlags(&L.ev_main[i], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&L.ev_fin[i], cudaEventDisableTiming));
        }
        for (int i = 0; i < 3; i++) CU(cudaEventCreateWithFlags(&L.ev_zero[i], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&L.ev_lane, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&L.ev_reset, cudaEventDisableTiming));
        int rc = 0;
        for (int i = 0; i < 2 && !rc; i++) rc = dev_alloc(h, &L.cont_list2[i], (size_t)h->ld);
        for (int i = 0; i < 3 && !rc; i++) rc = dev_alloc(h, &L.lane_list[i], (size_t)h->ld);
        if (!rc) rc = dev_alloc(h, &L.cont_count2, (size_t)2);
        if (!rc) rc = dev_alloc(h, &L.lane_count, (size_t)3);
        if (!rc) rc = dev_alloc(h, &L.in_lane, (size_t)h->ld);
        if (!rc) rc = dev_alloc(h, &L.next_fin, (size_t)1);
        if (!rc) rc = dev_alloc(h, &L.next_lane, (size_t)1);
        if (rc) return rc;
        CU(cudaEventRecord(L.ev_reset, h->stream));     // orders the allocations' memsets before B and C start
    }
    L.on = on != 0;
    return KPP_OK;
}

int kpp_gpu_sync(kpp_handle *h, kpp_step_report *report)
{
    if (!h) return fail(h, KPP_E_INVALID, "null handle");
    if (!h->parts.empty()) {
        // every part waits for its own device; counts add up, the device time is the slowest part's
        if (report) memset(report, 0, sizeof(*report));
        int rc_all = KPP_OK;
        for (kpp_handle *p : h->parts) {
            kpp_step_report r;
            const int rc = kpp_gpu_sync(p, &r);
            if (rc && !rc_all) { rc_all = rc; h->err = p->err; }
            if (report) {
                report->ntime = r.ntime;
                report->n_active += r.n_active; report->n_long_iter += r.n_long_iter; report->n_reint += r.n_reint;
                report->n_reint_fail += r.n_reint_fail; report->n_reset += r.n_reset;
                report->n_pivot_zero += r.n_pivot_zero; report->n_iter_cap += r.n_iter_cap;
                report->n_handed_over += r.n_handed_over; report->sum_iter += r.sum_iter;
                if (r.max_iter > report->max_iter) report->max_iter = r.max_iter;
                if (r.kernel_ms > report->kernel_ms) report->kernel_ms = r.kernel_ms;
            }
        }
        return rc_all;
    }
    CU(cudaSetDevice(h->device));
    JOIN(h);
    CU(cudaStreamSynchronize(h->stream));
    if (report) memset(report, 0, sizeof(*report));
    if (!h->stepped) return KPP_OK;
    const KppReportDev &r = *h->rep_host;
    h->a.coop_expect = r.n_handed_over;      // sizes the next cooperative launch (kpp_launch_coop)
    if (report) {
        report->ntime = h->last_ntime;
        report->n_active = r.n_active;
        report->n_long_iter = r.n_long_iter;
        report->n_reint = r.n_reint;
        report->n_reint_fail = r.n_reint_fail;
        report->n_reset = r.n_reset;
        report->n_pivot_zero = r.n_pivot_zero;
        report->n_iter_cap = r.n_iter_cap;
        report->max_iter = r.max_iter;
        report->n_handed_over = r.n_handed_over;
        report->sum_iter = r.sum_iter;
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, h->ev0, h->ev1) == cudaSuccess) report->kernel_ms = ms;
        else cudaGetLastError();
    }
    if (r.pivot_sticky > 0) {
        // sticky across steps that were queued without a sync in between; reported once
        CU(cudaMemsetAsync(&h->rep_dev->pivot_sticky, 0, sizeof(int), h->stream));
        h->rep_host->pivot_sticky = 0;
        return fail(h, KPP_E_PIVOT_ZERO, "Algorithm for solving tridiag matrix failed (bet = 0)");
    }
    return KPP_OK;
}

int kpp_gpu_get_status(kpp_handle *h, int32_t *status)
{
    if (!h || !status) return fail(h, KPP_E_INVALID, "null argument");
    const int total = h->parts.empty() ? h->host_npts : h->parts[0]->host_npts;
    return kpp_gpu_download_field(h, KPP_F_DIAG_STATUS, status, (size_t)total * 4);
}

int kpp_gpu_host_alloc(void **ptr, size_t bytes)
{
    kpp_handle *h = nullptr;
    if (!ptr) return fail(h, KPP_E_INVALID, "null argument");
    CU(cudaMallocHost(ptr, bytes ? bytes : 1));
    return KPP_OK;
}

int kpp_gpu_host_free(void *ptr)
{
    kpp_handle *h = nullptr;
    if (ptr) CU(cudaFreeHost(ptr));
    return KPP_OK;
}

// ---- unit-test entry points ------------------------------------------------
static int test_setup(int device)
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        cudaGetLastError();
        return fail(nullptr, KPP_E_NODEVICE, "no CUDA device; this library has no CPU fallback");
    }
    if (device < 0 || device >= ndev) return fail(nullptr, KPP_E_INVALID, "device index out of range");
    if (cudaSetDevice(device) != cudaSuccess) return fail(nullptr, KPP_E_CUDA, "cudaSetDevice failed");
    return 0;
}

int kpp_gpu_test_eos(int device, int numerics, int n, const double *S, const double *T, const double *P,
                     double *sig0, double *alpha, double *beta, double *cp)
{
    kpp_handle *h = nullptr;
    int rc = test_setup(device);
    if (rc) return rc;
    double *d = nullptr;
    const size_t nb = (size_t)n * 8;
    CU(cudaMalloc((void **)&d, 7 * nb));
    CU(cudaMemcpy(d, S, nb, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(d + n, T, nb, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(d + 2 * (size_t)n, P, nb, cudaMemcpyHostToDevice));
    cudaError_t e = numerics ? kpp_launch_test_eos_fast(n, d, d + n, d + 2 * (size_t)n, d + 3 * (size_t)n,
                                                        d + 4 * (size_t)n, d + 5 * (size_t)n, d + 6 * (size_t)n, 0)
                             : kpp_launch_test_eos_strict(n, d, d + n, d + 2 * (size_t)n, d + 3 * (size_t)n,
                                                          d + 4 * (size_t)n, d + 5 * (size_t)n, d + 6 * (size_t)n, 0);
    if (e != cudaSuccess) { cudaFree(d); return fail(h, KPP_E_CUDA, cudaGetErrorString(e)); }
    CU(cudaMemcpy(sig0, d + 3 * (size_t)n, nb, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(alpha, d + 4 * (size_t)n, nb, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(beta, d + 5 * (size_t)n, nb, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(cp, d + 6 * (size_t)n, nb, cudaMemcpyDeviceToHost));
    cudaFree(d);
    return KPP_OK;
}

int kpp_gpu_test_wscale(kpp_handle *h, int n, const double *sigma, const double *hbl, const double *ustar,
                        const double *bfsfc, double *wm, double *ws)
{
    if (!h) return fail(h, KPP_E_INVALID, "null handle");
    if (!h->parts.empty()) h = h->parts[0];
    CU(cudaSetDevice(h->device));
    double *d = nullptr;
    const size_t nb = (size_t)n * 8;
    CU(cudaMalloc((void **)&d, 6 * nb));
    CU(cudaMemcpy(d, sigma, nb, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(d + n, hbl, nb, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(d + 2 * (size_t)n, ustar, nb, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(d + 3 * (size_t)n, bfsfc, nb, cudaMemcpyHostToDevice));
    cudaError_t e = h->k.numerics
                        ? kpp_launch_test_wscale_fast(&h->a, n, d, d + n, d + 2 * (size_t)n, d + 3 * (size_t)n,
                                                      d + 4 * (size_t)n, d + 5 * (size_t)n, h->stream)
                        : kpp_launch_test_wscale_strict(&h->a, n, d, d + n, d + 2 * (size_t)n, d + 3 * (size_t)n,
                                                        d + 4 * (size_t)n, d + 5 * (size_t)n, h->stream);
    if (e != cudaSuccess) { cudaFree(d); return fail(h, KPP_E_CUDA, cudaGetErrorString(e)); }
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaMemcpy(wm, d + 4 * (size_t)n, nb, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(ws, d + 5 * (size_t)n, nb, cudaMemcpyDeviceToHost));
    cudaFree(d);
    return KPP_OK;
}

int kpp_gpu_test_swfrac(int device, int numerics, int n, const double *z, const int32_t *jerlov, double *out)
{
    kpp_handle *h = nullptr;
    int rc = test_setup(device);
    if (rc) return rc;
    double *dz = nullptr, *dout = nullptr;
    int *dj = nullptr;
    CU(cudaMalloc((void **)&dz, (size_t)n * 8));
    CU(cudaMalloc((void **)&dout, (size_t)n * 8));
    CU(cudaMalloc((void **)&dj, (size_t)n * 4));
    CU(cudaMemcpy(dz, z, (size_t)n * 8, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(dj, jerlov, (size_t)n * 4, cudaMemcpyHostToDevice));
    cudaError_t e = numerics ? kpp_launch_test_swfrac_fast(n, dz, dj, dout, 0) : kpp_launch_test_swfrac_strict(n, dz, dj, dout, 0);
    if (e != cudaSuccess) return fail(h, KPP_E_CUDA, cudaGetErrorString(e));
    CU(cudaMemcpy(out, dout, (size_t)n * 8, cudaMemcpyDeviceToHost));
    cudaFree(dz); cudaFree(dout); cudaFree(dj);
    return KPP_OK;
}

int kpp_gpu_test_div(int device, int numerics, int n, const double *a, const double *b, double *plain, double *split,
                     double *ok)
{
    kpp_handle *h = nullptr;
    int rc = test_setup(device);
    if (rc) return rc;
    double *d = nullptr;
    CU(cudaMalloc((void **)&d, (size_t)n * 8 * 5));
    CU(cudaMemcpy(d, a, (size_t)n * 8, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(d + n, b, (size_t)n * 8, cudaMemcpyHostToDevice));
    cudaError_t e = numerics ? kpp_launch_test_div_fast(n, d, d + n, d + 2 * (size_t)n, 0)
                             : kpp_launch_test_div_strict(n, d, d + n, d + 2 * (size_t)n, 0);
    if (e != cudaSuccess) return fail(h, KPP_E_CUDA, cudaGetErrorString(e));
    CU(cudaMemcpy(plain, d + 2 * (size_t)n, (size_t)n * 8, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(split, d + 3 * (size_t)n, (size_t)n * 8, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(ok, d + 4 * (size_t)n, (size_t)n * 8, cudaMemcpyDeviceToHost));
    cudaFree(d);
    return KPP_OK;
}

}  // extern "C"
