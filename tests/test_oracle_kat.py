"""CPU tests of the oracle: the reference's own known-answer values (the only golden
vectors in the reference tree), its input builders against the host-side builders,
size-independent properties, and the committed golden fixtures."""
import ctypes as C
import glob
import os

import numpy as np
import pytest

import oracle_lib
from mckpp_f90_b200 import synth, hostinit
from mckpp_f90_b200.fields import KppDims, KppConsts, copy_fields

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---- the reference's check values (src/mckpp_physics_state_equations.F90:24-25,105-111)
def test_cpsw_check_value():
    # "CHECK VALUE: CPSW = 3849.500 J/(KG DEG. C) FOR S = 40, T = 40 DEG C, P0= 10000 DECIBARS"
    # the polynomial as written gives 3849.49948: equal to the documented value at its printed precision
    assert abs(oracle_lib.cpsw(40.0, 40.0, 10000.0) - 3849.500) < 1e-3


def test_abk80_check_values():
    # "S=35,T=15(degC),P=0(dbar)-->Alpha=2.14136e-4, Beta=7.51638e-4, Kappa=4.32576e-5"
    a, b, k, s0, s = oracle_lib.abk80(35.0, 15.0, 0.0, want_kappa=True)
    assert abs(a - 2.14136e-4) < 1e-9 and abs(b - 7.51638e-4) < 1e-9 and abs(k - 4.32576e-5) < 1e-10  # +-1 in the last printed digit
    assert s0 == s            # P = 0: sigma = sigma0
    # "S=40,T=0(degC),P=10,000(dbar)-->Alpha=2.69822e-4, Beta=6.88317e-4, Kappa=3.55271e-5"
    a, b, k, s0, s = oracle_lib.abk80(40.0, 0.0, 10000.0, want_kappa=True)
    assert abs(a - 2.69822e-4) < 1e-9 and abs(b - 6.88317e-4) < 1e-9 and abs(k - 3.55271e-5) < 1e-10
    # UNESCO 1980 published density check: S=40, T=0... sigma(S=35,T=5,P=10000 dbar) is 1069.48914 in
    # Fofonoff & Millard; the reference documents no density value, so only sanity here
    assert 32.0 < s0 < 32.3 and 74.0 < s < 75.0


def test_abk80_temperature_clamp():
    # T clamped at -2 (state_equations.F90:143-144)
    assert oracle_lib.abk80(35.0, -5.0, 100.0) == oracle_lib.abk80(35.0, -2.0, 100.0)
    assert oracle_lib.cpsw(35.0, -5.0, 100.0) == oracle_lib.cpsw(35.0, -2.0, 100.0)


# ---- input builders: oracle C vs host numpy (both restate the same Fortran)
@pytest.mark.parametrize("nz,stretch,dscale", [(100, False, 0.0), (250, True, 4.0), (69, True, 2.0)])
def test_grid_builders_agree(nz, stretch, dscale):
    L = oracle_lib.lib()
    zm, hm, dm = np.zeros(nz + 1), np.zeros(nz + 1), np.zeros(nz + 1)
    L.orc_build_grid(C.c_int(nz), C.c_double(1000.0), C.c_int(int(stretch)), C.c_double(dscale),
                     zm.ctypes.data_as(C.c_void_p), hm.ctypes.data_as(C.c_void_p), dm.ctypes.data_as(C.c_void_p))
    z2, h2, d2 = hostinit.build_grid(nz, 1000.0, stretch, dscale)
    assert np.array_equal(zm, z2) and np.array_equal(hm, h2) and np.array_equal(dm, d2)
    assert hm[nz] == 1e-10 and zm[nz] == -1000.0 and dm[0] == 0.0
    assert abs(dm[nz] - 1000.0) < 1e-9 and np.all(np.diff(zm) < 0)
    dims = KppDims(npts=1, nz=nz)
    tri = np.zeros((dims.nztmax + 1, 2, 1), order="F")
    L.orc_build_tri(C.c_int(nz), C.c_int(dims.nztmax), C.c_double(1200.0), zm.ctypes.data_as(C.c_void_p),
                    hm.ctypes.data_as(C.c_void_p), tri.ctypes.data_as(C.c_void_p))
    assert np.array_equal(tri, hostinit.build_tri(dims, 1200.0, z2, h2))


def test_lookup_builders_agree():
    L = oracle_lib.lib()
    wmt, wst = np.zeros((892, 50), order="F"), np.zeros((892, 50), order="F")
    L.orc_build_lookup(C.c_double(0.4), wmt.ctypes.data_as(C.c_void_p), wst.ctypes.data_as(C.c_void_p))
    w2, s2 = hostinit.build_lookup(0.4)
    assert np.array_equal(wmt, w2) and np.array_equal(wst, s2)     # same libm pow on both sides
    assert np.all(np.isfinite(wmt)) and np.all(wmt >= 0) and np.all(wst >= 0)
    # neutral limit zehat = 0 (i = 891): wm = ws = vonk*ustar
    u = np.arange(50) * (0.04 / 49)
    assert np.allclose(wmt[891, :], 0.4 * u, rtol=1e-12, atol=1e-300)


def test_coriolis_and_forcing_map_agree():
    L = oracle_lib.lib()
    lat = np.array([-70.0, -2.4, 0.0, 1.0, 2.5, 45.0])
    f = np.zeros_like(lat)
    L.orc_coriolis(C.c_int(lat.size), lat.ctypes.data_as(C.c_void_p), f.ctypes.data_as(C.c_void_p))
    assert np.array_equal(f, hostinit.coriolis(lat))
    assert f[1] < 0 and f[2] > 0 and abs(f[3]) == abs(f[1]) and f[5] > f[4] > 0   # +-2.5 deg clamp
    cfg = synth.scaled(synth.CONFIGS["cfg2"], 5, 3)
    cf, fld, r = synth.make_case(cfg)
    synth.apply_forcing(cfg, cf, fld, r, 7)
    # same raw fluxes through the oracle's C map
    n = cfg.npts
    t_day = (6 * cf.consts.dto / 86400.0) % 1.0
    import math
    sun = max(0.0, 900.0 * math.sin(2.0 * math.pi * (t_day - 0.25)))
    raw = dict(taux=0.15 * (2 * r[:, 4] - 1), tauy=0.15 * (2 * r[:, 5] - 1), swf=sun * (0.6 + 0.4 * r[:, 3]),
               lwf=np.full(n, -60.0), lhf=-50.0 - 250.0 * r[:, 6], shf=np.full(n, -10.0), rain=2e-4 * r[:, 7] ** 2,
               snow=np.zeros(n))
    sfl = np.zeros_like(fld["sflux"])
    args = [np.ascontiguousarray(raw[k]) for k in ("taux", "tauy", "swf", "lwf", "lhf", "shf", "rain", "snow")]
    L.orc_fluxes_map(C.c_int(n), C.c_int(9), C.c_double(cf.consts.FLSN), C.c_double(cf.consts.EL),
                     *[a.ctypes.data_as(C.c_void_p) for a in args], fld["l_ocean"].ctypes.data_as(C.c_void_p),
                     sfl.ctypes.data_as(C.c_void_p))
    assert np.array_equal(sfl[:, 0:6, 4, 0], fld["sflux"][:, 0:6, 4, 0])


def test_tridmat_solves_the_system():
    L = oracle_lib.lib()
    rng = np.random.default_rng(3)
    n = 40
    cu = -rng.random(n); cl = -rng.random(n); cu[0] = 0.0; cl[-1] = 0.0
    cc = 1.0 - cu - cl + rng.random(n)
    x = rng.standard_normal(n)
    A = np.diag(cc) + np.diag(cl[:-1], 1) + np.diag(cu[1:], -1)
    rhs = A @ x
    yo = np.zeros(n + 1); yo[n] = 7.0
    yn = np.zeros(n + 1)
    pz = C.c_int(0)
    L.orc_tridmat(*[a.ctypes.data_as(C.c_void_p) for a in (cu, cc, cl, rhs, yo)], C.c_int(n),
                  yn.ctypes.data_as(C.c_void_p), C.c_int(n + 5), C.byref(pz))
    assert pz.value == 0 and np.allclose(yn[:n], x, rtol=1e-10) and yn[n] == 7.0   # yn(nzi+1) = yo(nzi+1)


# ---- size-independent properties of the column step
def _run(cfg, nsteps, nthreads=1, realloc=False, gidx=None):
    cf, f, r = synth.make_case(cfg, gidx=gidx)
    orc = oracle_lib.Oracle(cf, f, nthreads=nthreads)
    synth.apply_forcing(cfg, cf, f, r, 1)
    orc.initialize_ocean_model()
    for nt in range(1, nsteps + 1):
        synth.apply_forcing(cfg, cf, f, r, nt)
        assert orc.physics_driver(nt, realloc_1d=realloc) == 0
    return f, orc


def test_columns_are_independent_and_threads_do_not_matter():
    cfg = synth.scaled(synth.CONFIGS["cfg2"], 6, 4)
    f1, o1 = _run(cfg, 4, nthreads=1)
    f2, o2 = _run(cfg, 4, nthreads=4, realloc=True)
    for k in ("X", "U", "hmix", "kmix", "difm", "wX"):
        assert np.array_equal(f1[k], f2[k]), k
    # a subset of columns computed alone gives the same answer (no horizontal coupling)
    sel = np.array([3, 7, 20])
    f3, _ = _run(cfg, 4, gidx=sel)
    assert np.array_equal(f3["X"], f1["X"][sel]) and np.array_equal(f3["hmix"], f1["hmix"][sel])
    assert np.all(o1.diag["iter"] >= 6)          # 3 compulsory + 3 converged passes (ocnstep_mod.F90:122,170)


def test_land_points_are_untouched_and_rest_state_is_steady():
    cfg = synth.scaled(synth.CONFIGS["cfg2"], 6, 4)
    cf, f, r = synth.make_case(cfg)
    f["run_physics"][::3] = 0
    f["l_ocean"][::3] = 0
    before = copy_fields(f)
    orc = oracle_lib.Oracle(cf, f, nthreads=2)
    synth.apply_forcing(cfg, cf, f, r, 1)
    before["sflux"][...] = f["sflux"]
    orc.initialize_ocean_model()
    orc.physics_driver(1)
    land = f["run_physics"] == 0
    for k in ("U", "X", "Us", "Xs", "hmix", "difm", "rho"):
        assert np.array_equal(f[k][land], before[k][land]), k
    assert not np.array_equal(f["X"][~land], before["X"][~land])


def test_freeze_clamp_and_flag():
    cfg = synth.scaled(synth.CONFIGS["cfg5"], 8, 4)
    f, orc = _run(cfg, 3)
    assert f["X"][:, :, 0].min() >= -1.8          # overrides.F90:87-90
    assert f["freeze_flag"].max() > 0


# ---- golden fixtures (oracle outputs, tools/make_golden.py): guards the oracle against drift
@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.npz"))))
def test_oracle_reproduces_golden(path):
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(ROOT, "tools", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    name = os.path.basename(path)[:-4]
    cfg, n = mg.CASES[name]
    out = mg.run_case(cfg, n)
    gold = np.load(path)
    for k in gold.files:
        assert np.array_equal(out[k], gold[k]), (name, k)
