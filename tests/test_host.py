"""CPU tests of the host-side logic: data model, synthetic inputs, partitioning."""
import numpy as np

from mckpp_f90_b200 import synth, hostinit
from mckpp_f90_b200.fields import KppDims, allocate_3d_fields, field_shapes


def test_memory_image_matches_the_reference_allocators():
    d = KppDims(npts=7, nz=100)
    assert d.nzp1 == 101 and d.nztmax == 114 and d.nzp1tmax == 115
    f = allocate_3d_fields(d)
    # src/mckpp_data_fields.F90:355-446
    assert f["U"].shape == (7, 101, 2) and f["Us"].shape == (7, 101, 2, 2) and f["rho"].shape == (7, 116)
    assert f["buoy"].shape == (7, 115) and f["difm"].shape == (7, 115) and f["ghat"].shape == (7, 114)
    assert f["wU"].shape == (7, 115, 3) and f["wXNT"].shape == (7, 115, 2) and f["sflux"].shape == (7, 9, 5, 2)
    assert f["swdk_opt"].shape == (7, 101) and f["dbloc"].shape == (7, 100) and f["hmixd"].shape == (7, 2)
    for k, v in f.items():
        assert v.flags.f_contiguous and v.dtype in (np.float64, np.int32), k
    assert np.all(f["sflux"][:, :, 4, 0] == 1e-20)         # fluxes_mod.F90:27
    assert np.all(f["jerlov"] == 3)                          # initialize_optics_mod.F90:43


def test_synthetic_inputs_are_deterministic_and_partitionable():
    cfg = synth.scaled(synth.CONFIGS["cfg2"], 10, 6)
    cf, f, r = synth.make_case(cfg)
    cf2, f2, r2 = synth.make_case(cfg)
    assert all(np.array_equal(f[k], f2[k]) for k in f)
    # a rank's contiguous block equals the same columns of the whole domain
    cfb, fb, rb = synth.make_case(cfg, col_offset=20, ncols=25)
    for k in ("X", "f", "Sref", "dlat", "ocdepth"):
        assert np.array_equal(fb[k], f[k][20:45]), k
    s_all = synth.apply_forcing(cfg, cf, f, r, 5)
    s_blk = synth.apply_forcing(cfg, cfb, fb, rb, 5)
    assert np.array_equal(s_blk, s_all[:, 20:45])
    assert s_all.shape == (6, 60) and s_all.flags.c_contiguous


def test_configs_match_baseline_shapes():
    c = synth.CONFIGS
    assert c["cfg1"].npts == 16 and c["cfg1"].dto == 10800.0 and synth.nsteps(c["cfg1"]) == 8
    assert c["cfg2"].npts == 60000 and c["cfg2"].nz == 100 and synth.nsteps(c["cfg2"]) == 30 * 72
    assert c["cfg3"].npts == 44000 and c["cfg4"].npts == 700000 and c["cfg4"].LDD
    assert c["cfg5"].nz == 250 and c["cfg5"].stretch and c["cfg5"].corrections
    k = synth.make_consts(c["cfg5"])
    # legal together (initialize_namelist_mod.F90:251-265)
    assert k.L_FCORR_WITHZ and k.L_RELAX_OCNT and not k.L_FCORR and not k.L_RELAX_SST


def test_stretched_grid_reference_integral_trip_counts():
    # trips of the vmix reference-integral loop (verticalmixing_mod.F90:118-128)
    def trips(nz, stretch, dscale):
        zm, hm, dm = hostinit.build_grid(nz, 1000.0, stretch, dscale)
        tot = 0
        for n in range(1, nz + 1):
            zref = 0.1 * zm[n - 1]
            for kl in range(1, nz + 1):
                if zref >= zm[kl - 1]:
                    break
                tot += 1
        return tot
    assert trips(100, False, 0.0) == 500
    assert trips(250, True, 4.0) == 5426
