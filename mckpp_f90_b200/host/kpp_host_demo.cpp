// kpp_host_demo.cpp -- a compiled host above the C ABI, standing in for the Fortran main
// loop (src/mckpp_ocean_model_3D.F90:38-70) where no Fortran compiler exists.
//
//   kpp_host_demo <case.bin> <out.bin>
//
// case.bin (little endian; written by tests/test_gpu_host_cpp.py):
//   int32 npts,nz,nztmax,nsflxs,njdt,maxmodeadv,nsteps,nfields ; kpp_consts ;
//   zm(nzp1) hm(nzp1) dm(0:nz) tri(0:nztmax,0:1) wmt wst (892*50 each) ;
//   nfields x { int32 id ; int64 bytes ; data }   members of kpp_3d_fields to push ;
//   nsteps x 6*npts doubles                          sflux(:,1:6,5,0) of every step
// out.bin: X, U, hmix, kmix after the last step, then the packed XIOS blocks "S" and "difm".
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "mckpp_host.hpp"

template <class T> static void rd(FILE *fp, T *p, size_t n)
{
    if (fread(p, sizeof(T), n, fp) != n) { fprintf(stderr, "short read\n"); exit(2); }
}

int main(int argc, char **argv)
{
    if (argc < 3) { fprintf(stderr, "usage: %s case.bin out.bin\n", argv[0]); return 2; }
    FILE *fp = fopen(argv[1], "rb");
    if (!fp) { perror(argv[1]); return 2; }
    int32_t hd[8];
    rd(fp, hd, 8);
    kpp_dims d{hd[0], hd[1], hd[2], hd[3], hd[4], hd[5]};
    const int nsteps = hd[6], nfields = hd[7];
    kpp_consts k;
    rd(fp, &k, 1);
    const int nzp1 = d.nz + 1;
    std::vector<double> zm(nzp1), hm(nzp1), dm(d.nz + 1), tri((size_t)(d.nztmax + 1) * 2), wmt(892 * 50), wst(892 * 50);
    rd(fp, zm.data(), zm.size()); rd(fp, hm.data(), hm.size()); rd(fp, dm.data(), dm.size());
    rd(fp, tri.data(), tri.size()); rd(fp, wmt.data(), wmt.size()); rd(fp, wst.data(), wst.size());
    try {
        mckpp::PhysicsDriver drv(d, k, zm.data(), hm.data(), dm.data(), tri.data(), wmt.data(), wst.data(), 0);
        drv.verbose = false;
        for (int i = 0; i < nfields; i++) {
            int32_t id; int64_t nb;
            rd(fp, &id, 1); rd(fp, &nb, 1);
            mckpp::Field &fl = drv.kpp_3d_fields[kpp_gpu_field_name(id)];
            if ((size_t)nb != fl.bytes()) { fprintf(stderr, "size mismatch for field %d\n", id); return 2; }
            rd(fp, (char *)fl.data(), (size_t)nb);
        }
        drv.push_inputs();
        drv.mckpp_initialize_ocean_model();
        mckpp::Field &sf = drv.kpp_3d_fields["sflux"];
        for (int nt = 1; nt <= nsteps; nt++) {
            rd(fp, sf.r.data() + (size_t)4 * d.nsflxs * d.npts, (size_t)6 * d.npts);   // mckpp_fluxes
            drv.mckpp_physics_driver(nt);
        }
        drv.pull("X"); drv.pull("U");
        FILE *fo = fopen(argv[2], "wb");
        for (const char *n : {"X", "U", "hmix", "kmix"}) {
            mckpp::Field &fl = drv.kpp_3d_fields[n];
            fwrite(fl.data(), 1, fl.bytes(), fo);
        }
        // two blocks of the XIOS diagnostic set as the reference sends them: "S" and "difm"
        for (int id : {(int)KPP_OUT_S, (int)KPP_OUT_DIFM}) {
            const std::vector<double> blk = drv.xios_block(id);
            fwrite(blk.data(), sizeof(double), blk.size(), fo);
        }
        fclose(fo);
        printf("kpp_host_demo: %d columns x %d steps, last step kernel %.3f ms, max iter %d\n", d.npts, nsteps,
               drv.last.kernel_ms, drv.last.max_iter);
    } catch (const std::exception &e) {
        fprintf(stderr, "fatal: %s\n", e.what());
        return 1;
    }
    fclose(fp);
    return 0;
}
