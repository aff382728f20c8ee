"""Independent pin of the oracle: the C oracle's ocnstep control flow
and bldepth against a second, differently structured reading of the same Fortran
(oracle/second_reading.py).  Both must produce the same bits on every BASELINE configuration
(scaled), on the itermax / instability-trap / damping / isothermal branches, and on every single
bldepth call made along the way.  CPU only."""
import numpy as np
import pytest

import oracle_lib
import second_reading as sr
from mckpp_f90_b200 import synth
from mckpp_f90_b200.fields import copy_fields

SMALL = {
    "cfg1": (synth.CONFIGS["cfg1"], 8),
    "cfg2": (synth.scaled(synth.CONFIGS["cfg2"], 6, 4), 40),
    "cfg3": (synth.scaled(synth.CONFIGS["cfg3"], 5, 4), 12),
    "cfg4": (synth.scaled(synth.CONFIGS["cfg4"], 6, 4), 16),
    "cfg5": (synth.scaled(synth.CONFIGS["cfg5"], 4, 3), 8),
}
CHECK = ["U", "X", "Us", "Xs", "hmixd", "hmix", "kmix", "Tref", "uref", "vref", "Ssurf", "old", "new", "reset_flag",
         "dampu_flag", "dampv_flag", "freeze_flag", "rho", "cp", "buoy", "Rig", "dbloc", "Shsq", "difm", "difs", "dift",
         "ghat", "wU", "wX", "wXNT", "tinc_fcorr", "sinc_fcorr", "ocnTcorr", "scorr", "fcorr"]


def _pair(cfg, consts=None, setup=None):
    cf, fa, r = synth.make_case(cfg)
    for k, v in (consts or {}).items():
        setattr(cf.consts, k, v)
    if setup:
        setup(cf, fa, r)
    fb = copy_fields(fa)
    return cf, r, fa, oracle_lib.Oracle(cf, fa, nthreads=1), fb, oracle_lib.Oracle(cf, fb, nthreads=1)


def _run(cfg, nsteps, consts=None, setup=None, stress=None, second_ocnint=False, tail_log=None, head_log=None):
    cf, r, fa, oa, fb, ob = _pair(cfg, consts, setup)
    log = []
    seen = {"nreint": 0, "status": 0, "iter": 0}
    probe = sr.bldepth_probe(cf, log)
    if tail_log is not None:
        p1, p2 = probe, sr.vmix_tail_probe(cf, tail_log)

        def probe(col):
            p1(col); p2(col)
    if head_log is not None:
        p3, p4 = probe, sr.vmix_head_probe(cf, head_log)

        def probe(col):
            p3(col); p4(col)
    synth.apply_forcing(cfg, cf, fa, r, 1)
    fb["sflux"][...] = fa["sflux"]
    oa.initialize_ocean_model()
    if second_ocnint:       # the second reading of the initialisation loop too (its vmix is probed like every other)
        sr.initialize_ocean_model2(ob, cf, fb, probe)
    else:
        ob.initialize_ocean_model()
    for name in CHECK:
        assert np.array_equal(fa[name], fb[name], equal_nan=True), ("init", name)
    for nt in range(1, nsteps + 1):
        synth.apply_forcing(cfg, cf, fa, r, nt)
        if stress:
            stress(fa, nt)
        fb["sflux"][...] = fa["sflux"]
        oa.physics_driver(nt)
        sr.physics_driver(ob, cf, fb, nt, probe, second_ocnint)
        for name in CHECK:
            assert np.array_equal(fa[name], fb[name], equal_nan=True), (nt, name)
        for d in ("iter", "nreint", "status"):
            assert np.array_equal(oa.diag[d], ob.diag[d]), (nt, d, oa.diag[d], ob.diag[d])
        seen["nreint"] = max(seen["nreint"], int(oa.diag["nreint"].max()))
        seen["iter"] = max(seen["iter"], int(oa.diag["iter"].max()))
        seen["status"] |= int(np.bitwise_or.reduce(oa.diag["status"]))
    oa.seen = seen
    return log, oa


def _check_bldepth_log(log):
    assert log
    for ours, theirs in log:
        assert ours[1] == theirs[1], ("kbl", ours, theirs)
        for a, b in zip(ours, theirs):
            assert a == b or (np.isnan(a) and np.isnan(b)), (ours, theirs)


@pytest.mark.parametrize("name", list(SMALL))
def test_second_reading_of_ocnstep_and_bldepth_is_bitwise_the_oracle(name):
    cfg, nsteps = SMALL[name]
    log, oa = _run(cfg, nsteps)
    _check_bldepth_log(log)
    assert oa.diag["iter"].min() >= 6


def test_second_reading_storm_deepens_the_boundary_layer():
    """Strong wind and cooling: the boundary layer deepens through many levels, so the scan, the hmix
    convergence test and the iteration count take many different values."""
    cfg = synth.scaled(synth.CONFIGS["cfg2"], 6, 4)

    def stress(f, nt):
        f["sflux"][:, 0, 4, 0] *= 6.0
        f["sflux"][:, 1, 4, 0] *= 6.0
        f["sflux"][:, 3, 4, 0] = -900.0

    log, oa = _run(cfg, 60, stress=stress)
    _check_bldepth_log(log)
    assert len({t[1] for _, t in log}) >= 5


def test_second_reading_itermax_branches():
    """itermax = 4: `iter .lt. itermax` fails while iconv < 3, so both `hmixn > hmixe` outcomes of
    ocnstep_mod.F90:175-181 (goto 45 past itermax / fall through) and the 'long iteration' status occur."""
    cfg = synth.scaled(synth.CONFIGS["cfg2"], 8, 6)
    log, oa = _run(cfg, 30, consts=dict(itermax=4))
    _check_bldepth_log(log)
    assert oa.seen["iter"] > 5 and oa.seen["status"] & 1      # ran past itermax + 1: 'long iteration'


@pytest.mark.parametrize("second", [False, True])
def test_second_reading_trap_damping_isothermal(second):
    # second: ocnint, the solvers and check_profile (reset to climatology / U_init, isothermal reset) from the second reading too
    cfg = synth.scaled(synth.CONFIGS["cfg2"], 6, 4)

    def setup(cf, f, r):
        f["U_init"][:, :, 0] = 0.01
        f["ocnT_clim"][:] = f["X"][:, :, 0]
        f["sal_clim"][:] = f["X"][:, :, 1]
        f["X"][::5, :, 0] = 5.0

    def stress(f, nt):
        if nt == 2:
            f["sflux"][::7, 0, 4, 0] = 4000.0       # |U| >= 10: trap, 11 integrations, reset

    log, oa = _run(cfg, 3, consts=dict(L_DAMP_CURR=True, L_NO_ISOTHERM=True, iso_bot=30, iso_thresh=0.05,
                                       have_ocnT_file=True, have_sal_file=True, L_VARY_BOTTOM_TEMP=True),
                   setup=setup, stress=stress, second_ocnint=second)
    _check_bldepth_log(log)
    # 11 integrations, 'failed to find a reasonable solution', reset, isothermal reset
    assert oa.seen["nreint"] == 11 and oa.seen["status"] & 2 and oa.seen["status"] & 4 and oa.seen["status"] & 32


def test_second_reading_bldepth_on_adversarial_inputs():
    """bldepth alone on random inputs that reach the branches a smooth run rarely does: Monin-Obukhov
    and Ekman limits, shallow ocdepth, the l_initflag exception, negative Ritop, scan reaching km."""
    import ctypes as C
    cfg = synth.scaled(synth.CONFIGS["cfg5"], 2, 2)
    cf, f, r = synth.make_case(cfg)
    orc = oracle_lib.Oracle(cf, f, nthreads=1)
    col = sr.Column(orc)
    L = col.L
    n1 = cf.dims.nzp1 + 1
    zm = np.concatenate([[0.0], cf.zm]); hm = np.concatenate([[0.0], cf.hm])
    rng = np.random.default_rng(11)
    kbls = set()
    for trial in range(400):
        L.orc_col_load(col.h, C.byref(orc.c), C.byref(orc.s), 1, 5)
        # a random but physically shaped state: the C oracle's vmix then calls its bldepth on it
        nz = cf.dims.nz
        depth = rng.uniform(5.0, 600.0)
        col.X[0, :] = 2.0 + rng.uniform(5, 25) * np.exp(cf.zm / depth) + 0.01 * rng.standard_normal(nz + 1)
        col.X[1, :] = 0.3 * rng.standard_normal() * np.exp(cf.zm / 200.0)
        col.U[0, :] = rng.uniform(-1, 1) * np.exp(cf.zm / rng.uniform(5, 100))
        col.U[1, :] = rng.uniform(-1, 1) * np.exp(cf.zm / rng.uniform(5, 100))
        sfl = col._arr("sflux").reshape(-1, cf.dims.nsflxs)       # [(time level, j), i]: sflux(i,5,0) = sfl[4, i-1]
        sfl[4, 0:6] = [rng.normal(0, 0.1), rng.normal(0, 0.1), rng.choice([0.0, rng.uniform(0, 900)]),
                       rng.uniform(-400, 200), rng.choice([1e-10, -1e-5]), rng.normal(0, 1e-4)]
        col.set("f", rng.choice([1e-4, -7e-5, 1e-6, 3e-9]))
        col.set("ocdepth", rng.choice([-10000.0, -rng.uniform(5.0, 900.0)]))
        col.set("jerlov", int(rng.integers(1, 6)))
        col.set("l_initflag", float(trial % 5 == 0))
        h, k = col.vmix()
        g = col.get
        ours = sr.bldepth(zm, hm, cf.consts.vonk, np.asarray(cf.wmt), np.asarray(cf.wst), col._arr("dVsq")[:n1],
                          col._arr("Ritop")[:n1], col._arr("dbloc"), col._arr("swfrac")[:n1], g("dbg_ustar"),
                          g("dbg_Bo"), g("dbg_Bosol"), g("f"), g("ocdepth"), int(g("jerlov")), bool(g("l_initflag")))
        theirs = (g("dbg_hbl"), int(g("dbg_kbl")), g("dbg_bfsfc"), g("dbg_stable"), g("dbg_caseA"))
        assert ours == theirs, (trial, ours, theirs)
        kbls.add(theirs[1])
    col.close()
    assert len(kbls) >= 10


# --------------------------------------------------------------------------- ocnint and the solvers, second reading
def _setup_relax_sst(cf, f, r):
    f["relax_sst"][:] = 1.0 / (10 * 86400.0)
    f["relax_sst"][::5] = 0.0
    f["SST0"][:] = f["X"][:, 0, 0] + 0.5


def _setup_fcorr(cf, f, r):
    f["fcorr_twod"][:] = 40.0 * (r[:, 0] - 0.5)


def _setup_advection(cf, f, r):
    n = f["nmodeadv"].shape[0]
    f["nmodeadv"][:, 1] = (np.arange(n) % 4)
    for m in range(6):
        f["modeadv"][:, m, 1] = 1 + (np.arange(n) + m) % 7
        f["advection"][:, m, 1] = 1e-6 * (r[:, m] - 0.5)


@pytest.mark.parametrize("name", list(SMALL))
def test_second_reading_of_ocnint_and_solvers_is_bitwise_the_oracle(name):
    """ocnint, tridcof, tridrhs, tridmat (ocnint_mod.F90:19-221, solvers.F90:14-161) restated in numpy, array-at-a-time,
    take the place of the C oracle's in every pass of every column: all fields 1dto3d writes stay bit-identical
    (cfg5 brings the flux corrections at depth, the relaxations and the freeze clamp)."""
    cfg, nsteps = SMALL[name]
    log, oa = _run(cfg, min(nsteps, 12), second_ocnint=True)
    _check_bldepth_log(log)


@pytest.mark.parametrize("case", ["relax_sst", "relax_sst_calconly", "fcorr_twod", "advection", "ldd"])
def test_second_reading_of_ocnint_switches(case):
    """The surface relaxation / two-dimensional flux correction branches (ocnint_mod.F90:98-123), all seven
    advection modes of rhsmod (solvers.F90:229-334; three salinity modes per column, every mode on some column)
    and separate T and S matrices (double diffusion)."""
    consts, setup = {"relax_sst": (dict(L_RELAX_SST=True), _setup_relax_sst),
                     "relax_sst_calconly": (dict(L_RELAX_SST=True, L_RELAX_CALCONLY=True), _setup_relax_sst),
                     "fcorr_twod": (dict(L_FCORR=True), _setup_fcorr),
                     "advection": (None, _setup_advection),
                     "ldd": (dict(LDD=True), None)}[case]
    cfg = synth.scaled(synth.CONFIGS["cfg2"], 7, 4)
    log, oa = _run(cfg, 8, consts=consts, setup=setup, second_ocnint=True)
    _check_bldepth_log(log)


def test_second_reading_of_tridmat_zero_pivot():
    """bet == 0 (solvers.F90:140-150): flagged, replaced by 1.E-12, and the sweep goes on -- in both readings."""
    cfg = synth.scaled(synth.CONFIGS["cfg2"], 3, 2)

    def setup(cf, f, r):
        cf.tri[60, 0, 0] = 0.0          # cu(60) = 0 and cc(60) = 1 + tri(60,1)*diff(60)
        cf.tri[60, 1, 0] = -1.0e4       # ... = 0 for diff(60) = 1e-4 (the background viscosity of the quiet interior)

    log, oa = _run(cfg, 2, setup=setup, second_ocnint=True)
    assert oa.seen["status"] & 8


# --------------------------------------------------------------------------- second half of vmix, second reading
def _check_tail_log(tail):
    assert tail
    names = set()
    for name, ours, theirs in tail:
        names.add(name)
        assert np.array_equal(ours, theirs, equal_nan=True), (name, np.nonzero(ours != theirs)[0][:5], ours[ours != theirs][:3], theirs[ours != theirs][:3])
    return names


@pytest.mark.parametrize("name", list(SMALL))
def test_second_reading_of_kppmix_rimix_ddmix_blmix_enhance(name):
    """After every vmix of the C oracle the numpy reading of verticalmixing_mod.F90:102-159, kppmix, rimix, z121, ddmix,
    blmix and enhance recomputes alphaDT, betaDS, Ritop, dVsq, dbloc, Shsq, Rig, difm, difs, dift and ghat from the
    iterate and the EOS results: bitwise equal on every level of every pass (cfg4: double diffusion on)."""
    cfg, nsteps = SMALL[name]
    tail = []
    log, oa = _run(cfg, min(nsteps, 6), tail_log=tail)
    names = _check_tail_log(tail)
    assert {"difm", "difs", "dift", "ghat", "Ritop", "dVsq", "Rig"} <= names


def test_second_reading_of_ddmix_both_branches_and_no_ri():
    cfg = synth.scaled(synth.CONFIGS["cfg2"], 6, 4)

    def setup(cf, f, r):        # cold and fresh over warm and salty (diffusive), salty over fresh (fingering)
        zm = cf.zm
        even = np.arange(f["X"].shape[0]) % 2 == 0
        f["X"][even, :, 0] = -1.0 + 4.0 * (1.0 - np.exp(zm[None, :] / 150.0))
        S = np.where(even[:, None], 34.0 + 0.2 * (1.0 - np.exp(zm[None, :] / 150.0)), 35.0 + 1.0 * np.exp(zm[None, :] / 300.0))
        f["X"][:, :, 1] = S - f["Sref"][:, None]

    tail = []
    _run(cfg, 4, consts=dict(LDD=True), setup=setup, tail_log=tail)
    _check_tail_log(tail)
    tail = []
    _run(cfg, 4, consts=dict(LRI=False), tail_log=tail)
    _check_tail_log(tail)


def test_second_reading_of_kppmix_under_cooling_and_wind():
    """Strong cooling and wind: unstable forcing (the nonlocal term ghat is non-zero), boundary layers many levels
    deep, both caseA outcomes in enhance -- the branches a quiet start never takes."""
    cfg = synth.scaled(synth.CONFIGS["cfg2"], 4, 3)

    def stress(f, nt):
        f["sflux"][:, 0, 4, 0] *= 6.0
        f["sflux"][:, 1, 4, 0] *= 6.0
        f["sflux"][:, 3, 4, 0] = -900.0

    tail = []
    log, oa = _run(cfg, 30, stress=stress, tail_log=tail)
    _check_tail_log(tail)
    ghat = [t[1] for t in tail if t[0] == "ghat"]
    assert max(float(g.max()) for g in ghat) > 0.0
    assert len({t[1] for _, t in log}) >= 4                   # kbl takes several values
    assert {t[4] for _, t in log} == {0.0, 1.0}               # caseA: both


# --------------------------------------------------------------------------- first half of vmix, second reading
@pytest.mark.parametrize("name", list(SMALL))
def test_second_reading_of_eos_cpsw_ntflux_surface_fluxes(name):
    """ABK80 (Sig80, Bet80, Alf80), CPSW, ntflux/swdk and the kinematic surface fluxes of
    verticalmixing_mod.F90:47-100, restated profile-at-a-time in numpy, against what the C oracle's vmix left in the
    column: rho, cp, talpha, sbeta, buoy, swdk_opt, wXNT, rhoh2o, wU(0,:), wX(0,:), ustar, B0, B0sol -- bitwise, every
    level, every pass (cfg5: temperatures at the -2 degC clamp, NZ=250 pressures)."""
    cfg, nsteps = SMALL[name]
    head = []
    log, oa = _run(cfg, min(nsteps, 6), head_log=head)
    names = _check_tail_log(head)
    assert {"rho", "cp", "talpha", "sbeta", "buoy", "wXNT", "ustar", "Bo", "Bosol", "wX0", "wU0", "rhoh2o", "swfrac", "swdk_opt"} <= names


def test_second_reading_eos_check_values():
    """The reference's own check values (state_equations.F90:24-26, 107-113) through the second reading."""
    a, b, s0 = sr.abk80(np.array([40.0]), np.array([0.0]), np.array([10000.0]))
    assert abs(a[0] - 2.69822e-4) < 5e-10 and abs(b[0] - 6.88317e-4) < 5e-10
    assert abs(float(sr.cpsw(np.array([40.0]), np.array([40.0]), np.array([10000.0]))[0]) - 3849.500) < 1e-3
