// mckpp_host.hpp -- C++ host-side mirror of the reference's driver interface over the
// C ABI (include/kpp_gpu.h).  Where a Fortran host would hold the module globals
// kpp_3d_fields / kpp_const_fields (src/mckpp_data_fields.F90:348-349), this host holds the
// same arrays as std::vector with the Fortran (column-major, npts-first) element order,
// and exposes the reference's entry points with the same names and meaning:
//
//   mckpp_initialize_ocean_model()   src/mckpp_initialize_ocean.F90:18
//   mckpp_physics_driver()           src/mckpp_physics_driver_mod.F90:15
//
// Errors follow the reference: warnings to stderr, fatal conditions (tridiagonal zero
// pivot, no device) throw -- the reference calls MCKPP_ABORT.  No CPU fallback.
#pragma once
#include <cstdint>
#include <cstdio>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/kpp_gpu.h"

namespace mckpp {

struct Field {
    int id;
    bool is_int;
    std::vector<double> r;      // REAL members
    std::vector<int32_t> i;     // INTEGER / LOGICAL members
    void *data() { return is_int ? (void *)i.data() : (void *)r.data(); }
    size_t bytes() const { return is_int ? i.size() * 4 : r.size() * 8; }
};

class Kpp3dFields {
  public:
    std::map<std::string, Field> f;
    Field &operator[](const std::string &n) { return f.at(n); }
    // allocate every member the path touches with the whole-array size the ABI expects
    void allocate(kpp_handle *h)
    {
        for (int id = 0; id < KPP_F__COUNT; id++) {
            const std::string name = kpp_gpu_field_name(id);
            Field fl;
            fl.id = id;
            fl.is_int = (name == "old" || name == "new" || name == "jerlov" || name == "l_ocean" ||
                         name == "run_physics" || name == "nmodeadv" || name == "modeadv" || name == "diag_iter" ||
                         name == "diag_nreint" || name == "diag_status");
            const size_t nb = kpp_gpu_field_host_bytes(h, id);
            if (fl.is_int) fl.i.assign(nb / 4, 0); else fl.r.assign(nb / 8, 0.0);
            f[name] = fl;
        }
    }
};

class PhysicsDriver {
  public:
    kpp_handle *h = nullptr;
    kpp_dims dims;
    Kpp3dFields kpp_3d_fields;
    kpp_step_report last{};
    bool verbose = true;

    PhysicsDriver(const kpp_dims &d, const kpp_consts &k, const double *zm, const double *hm, const double *dm,
                  const double *tri, const double *wmt, const double *wst, int device = 0)
        : dims(d)
    {
        const int rc = kpp_gpu_create(&d, &k, zm, hm, dm, tri, wmt, wst, device, &h);
        if (rc != KPP_OK) throw std::runtime_error(std::string("kpp_gpu_create: ") + kpp_gpu_last_error(nullptr));
        kpp_3d_fields.allocate(h);
    }
    ~PhysicsDriver() { if (h) kpp_gpu_destroy(h); }
    PhysicsDriver(const PhysicsDriver &) = delete;
    PhysicsDriver &operator=(const PhysicsDriver &) = delete;

    void check(int rc, const char *what)
    {
        if (rc != KPP_OK) throw std::runtime_error(std::string(what) + ": " + kpp_gpu_last_error(h));
    }
    void push(const std::string &name)
    {
        Field &fl = kpp_3d_fields[name];
        check(kpp_gpu_upload_field(h, fl.id, fl.data(), fl.bytes()), name.c_str());
    }
    void pull(const std::string &name)
    {
        Field &fl = kpp_3d_fields[name];
        check(kpp_gpu_download_field(h, fl.id, fl.data(), fl.bytes()), name.c_str());
    }
    void push_inputs()
    {
        static const char *names[] = {"U", "X", "Us", "Xs", "hmixd", "old", "new", "hmix", "kmix", "Tref", "uref", "vref",
                                      "Ssurf", "Sref", "SSref", "f", "ocdepth", "jerlov", "l_ocean", "run_physics",
                                      "sflux", "U_init", "relax_sst", "SST0", "fcorr_twod", "fcorr", "relax_sal",
                                      "relax_ocnT", "sal_clim", "ocnT_clim", "fcorr_withz", "sfcorr_withz",
                                      "bottom_temp", "nmodeadv", "modeadv", "advection", "freeze_flag"};
        for (const char *n : names) push(n);
    }
    void pull_scalars()
    {
        static const char *names[] = {"hmix", "kmix", "Tref", "uref", "vref", "Ssurf", "old", "new", "reset_flag",
                                      "dampu_flag", "dampv_flag", "freeze_flag", "fcorr"};
        for (const char *n : names) pull(n);
    }
    // per-column loop of MCKPP_INITIALIZE_OCEAN_MODEL
    void mckpp_initialize_ocean_model()
    {
        check(kpp_gpu_init_vmix(h), "kpp_gpu_init_vmix");
        check(kpp_gpu_sync(h, nullptr), "kpp_gpu_sync");
        for (const char *n : {"hmix", "kmix", "Tref", "uref", "vref", "old", "new", "hmixd", "Us", "Xs"}) pull(n);
    }
    // mckpp_physics_driver() for timestep ntime; the forcing is read from kpp_3d_fields["sflux"]
    void mckpp_physics_driver(int ntime)
    {
        Field &sf = kpp_3d_fields["sflux"];
        const size_t n = (size_t)dims.npts;
        // sflux(:,1:6,5,0): rows (0*5+4)*nsflxs .. +5 of the (npts,nsflxs,5,0:njdt) array
        check(kpp_gpu_upload_forcing(h, sf.r.data() + (size_t)4 * dims.nsflxs * n), "kpp_gpu_upload_forcing");
        check(kpp_gpu_step(h, ntime), "kpp_gpu_step");
        const int rc = kpp_gpu_sync(h, &last);
        if (verbose) {
            if (last.n_long_iter)
                fprintf(stderr, "MCKPP_PHYSICS_OCNSTEP: long iteration at timestep %d on %d points\n", ntime, last.n_long_iter);
            if (last.n_reint_fail)
                fprintf(stderr, "MCKPP_PHYSICS_OCNSTEP: Failed to find a reasonable solution in the semi-implicit "
                                "integration after 10 iterations on %d points\n", last.n_reint_fail);
            if (last.n_reset)
                fprintf(stderr, "MCKPP_PHSYICS_OVERRIDE_CHECK_PROFILE: Resetting %d points\n", last.n_reset);
        }
        if (rc == KPP_E_PIVOT_ZERO) {
            fprintf(stderr, "MCKPP_PHSYICS_SOLVER_TRIDMAT: Algorithm for solving tridiag matrix failed.\n");
            throw std::runtime_error("MCKPP_ABORT");
        }
        check(rc, "kpp_gpu_sync");
        pull_scalars();
    }
    // one xios_send_field block of mckpp_xios_diagnostic_output / _restart_output (xios_io.F90:72-207,
    // 406-431), packed on the device: (npts, rows) doubles, column index fastest
    std::vector<double> xios_block(int out_id)
    {
        const int rows = kpp_gpu_output_rows(h, out_id);
        if (rows < 0) throw std::runtime_error("unknown output id");
        std::vector<double> block((size_t)dims.npts * rows);
        check(kpp_gpu_pack_output(h, out_id, block.data(), block.size() * sizeof(double)), kpp_gpu_output_name(out_id));
        return block;
    }
    // MCKPP_BOUNDARY_INTERPOLATE_TEMP/_SAL (boundary_interpolate.F90:60,115) on the device; records
    // (npts, nzp1) are handed over only when the time bracket moved
    void boundary_interpolate(int field_id, double prev_weight, double next_weight, const double *prev_rec = nullptr,
                              const double *next_rec = nullptr)
    {
        const size_t bytes = (size_t)dims.npts * (dims.nz + 1) * sizeof(double);
        if (prev_rec) check(kpp_gpu_upload_clim_record(h, field_id, 0, prev_rec, bytes), "kpp_gpu_upload_clim_record");
        if (next_rec) check(kpp_gpu_upload_clim_record(h, field_id, 1, next_rec, bytes), "kpp_gpu_upload_clim_record");
        check(kpp_gpu_blend_clim(h, field_id, prev_weight, next_weight), "kpp_gpu_blend_clim");
    }
};

}  // namespace mckpp
