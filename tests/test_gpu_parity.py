"""GPU parity tests (run on the B200 box with -m gpu): the CUDA path, called through the
C ABI, against the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star):
  * integer outputs (kmix = kbl, old/new, final `iter`, number of re-integrations,
    reset_flag, status bits): bit-exact; any flip is enumerated in the failure message;
  * strict numerics, teacher-forced single step (both sides start the step from the
    oracle's state): every field 1dto3d writes agrees to <= 1e-12 relative -- in
    practice bit-identical, the only source of difference being exp() (CUDA libdevice
    vs glibc, <= 1 ulp each);
  * strict numerics, free running for N days: T,S <= 1e-8, U,V,hmix,diffusivities <= 1e-6;
  * fast numerics (FMA contraction + shared reciprocals): teacher-forced <= 1e-9,
    free running same N-day tolerances, integer flips allowed on <= 0.5 % of columns and
    enumerated.
"""
import glob
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import oracle_lib
import parity
from mckpp_f90_b200 import capi, driver, synth
from mckpp_f90_b200.fields import copy_fields

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

TOL_TEACHER_STRICT = 1e-12
TOL_TEACHER_FAST = 1e-9
TOL_FREE_TS = 1e-8
TOL_FREE_OTHER = 1e-6


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if capi.load().kpp_gpu_device_count() < 1:
        pytest.fail("no CUDA device: the product path has no CPU fallback and these tests need the B200")


def _assert_ints_exact(P, what):
    im = P.int_mismatches()
    bad = {k: v for k, v in im.items() if not k.endswith("_count") and v}
    assert not bad, f"{what}: integer outputs differ (column, gpu, oracle): {bad}"


def _assert_bitwise(P, what):
    run = P.f_orc["run_physics"] != 0
    for fld in parity.FLOAT_FIELDS:
        assert np.array_equal(P.f_gpu[fld][run], P.f_orc[fld][run]), (what, fld)


def _worst(P, fields=None):
    c = P.compare(fields)
    k = max(c, key=lambda n: c[n][1])
    return k, c[k][1], c


# --------------------------------------------------------------------------- unit parity
def test_eos_matches_oracle_bitwise_and_reference_check_values():
    rng = np.random.default_rng(1)
    n = 4096
    S = rng.uniform(0.0, 41.0, n); T = rng.uniform(-3.0, 40.0, n); P = rng.uniform(0.5, 10000.0, n)
    S[:3] = [35.0, 40.0, 40.0]; T[:3] = [15.0, 0.0, 40.0]; P[:3] = [1e-6, 10000.0, 10000.0]
    sig0, alpha, beta, cp = capi.test_eos(S, T, P, numerics=0)
    ref = np.array([oracle_lib.abk80(s, t, p) for s, t, p in zip(S, T, P)])
    cpr = np.array([oracle_lib.cpsw(s, t, p) for s, t, p in zip(S, T, P)])
    assert np.array_equal(alpha, ref[:, 0]) and np.array_equal(beta, ref[:, 1]) and np.array_equal(sig0, ref[:, 3])
    assert np.array_equal(cp, cpr)
    # the reference's own check values (state_equations.F90:24-25,109-111)
    assert abs(alpha[1] - 2.69822e-4) < 1e-9 and abs(beta[1] - 6.88317e-4) < 1e-9
    assert abs(cp[2] - 3849.500) < 1e-3
    # fast variant: tolerance only
    s2, a2, b2, c2 = capi.test_eos(S, T, P, numerics=1)
    assert parity.rel_err(a2, alpha) < 1e-13 and parity.rel_err(b2, beta) < 1e-13
    assert parity.rel_err(s2, sig0) < 1e-12 and parity.rel_err(c2, cp) < 1e-14


def test_exp_path_bitwise_when_host_libm_emulation_is_active():
    rng = np.random.default_rng(2)
    z = rng.uniform(0.5, 900.0, 5000)
    j = rng.integers(1, 6, 5000).astype(np.int32)
    got = capi.test_swfrac(z, j)
    import ctypes as C
    L = oracle_lib.lib()
    ref = np.empty_like(z)
    out = C.c_double()
    for i in range(z.size):
        L.orc_swfrac(C.c_double(-1.0), C.c_double(z[i]), C.c_int(int(j[i])), C.byref(out))
        ref[i] = out.value
    if capi.load().kpp_gpu_exp_is_host_libm(0):
        # strict variant evaluates glibc's exp algorithm with the host libm's table: same bits
        assert np.array_equal(got, ref)
    else:
        assert parity.rel_err(got, ref) < 5e-16
    fast = capi.test_swfrac(z, j, numerics=1)
    assert parity.rel_err(fast, ref) < 1e-15


def test_split_division_equals_ieee_division_bitwise():
    """div_recip/div_with (kpp_kernels.cu) take the reciprocal off the Thomas dependency chain in
    the cooperative kernel; wherever they claim validity they must be the IEEE quotient, and
    a/b on the device must be the host's a/b."""
    rng = np.random.default_rng(7)
    n = 1 << 20
    a = rng.standard_normal(n) * 10.0 ** rng.uniform(-12, 12, n)
    b = rng.standard_normal(n) * 10.0 ** rng.uniform(-12, 12, n)
    # extremes: zeros of both signs, denormals, huge, tiny, inf, nan, exact quotients, near-ties
    ext = np.array([0.0, -0.0, 5e-324, -5e-324, 2.2250738585072014e-308, 1e-300, 1e-200, 1e-100, 1e-30, 1.0, -1.0, 3.0,
                    1.0 / 3.0, 1e30, 1e100, 1e200, 1e300, 1.7976931348623157e308, np.inf, -np.inf, np.nan,
                    1.0 + 2.0 ** -52, 1.0 - 2.0 ** -53, 2.0 ** -969, 2.0 ** -970, 2.0 ** 1000])
    ea, eb = np.meshgrid(ext, ext)
    a = np.concatenate([a, ea.ravel(), rng.integers(1, 1 << 53, 4096).astype(np.float64)])
    b = np.concatenate([b, eb.ravel(), rng.integers(1, 1 << 53, 4096).astype(np.float64)])
    for numerics in (0, 1):
        plain, split, ok = capi.test_div(a, b, numerics=numerics)
        with np.errstate(all="ignore"):
            host = a / b
        assert np.array_equal(plain.view(np.uint64)[~np.isnan(host)], host.view(np.uint64)[~np.isnan(host)])
        assert np.array_equal(np.isnan(plain), np.isnan(host))
        assert np.array_equal(split.view(np.uint64)[ok & ~np.isnan(host)], plain.view(np.uint64)[ok & ~np.isnan(host)])
        assert not np.isnan(split[ok]).any()
        # the guard is not vacuous: ordinary operands pass it
        assert ok[: 1 << 20].mean() > 0.999


def test_wscale_matches_oracle_bitwise():
    cfg = synth.scaled(synth.CONFIGS["cfg1"], 2, 2)
    cf, f, r = synth.make_case(cfg)
    g = capi.KppGpu(cf)
    rng = np.random.default_rng(4)
    n = 20000
    sigma = rng.uniform(0.0, 1.0, n); hbl = rng.uniform(1.0, 900.0, n); ustar = rng.uniform(0.0, 0.06, n)
    bfsfc = rng.normal(0.0, 3e-7, n)
    bfsfc[:200] = rng.normal(0.0, 1e-4, 200)          # beyond the table: clamped index, extrapolated fraction
    wm, ws = capi.test_wscale(g, sigma, hbl, ustar, bfsfc)
    import ctypes as C
    orc = oracle_lib.Oracle(cf, f)
    rm, rs = np.empty(n), np.empty(n)
    a, b = C.c_double(), C.c_double()
    for i in range(n):
        orc.L.orc_wscale(C.byref(orc.c), C.c_double(sigma[i]), C.c_double(hbl[i]), C.c_double(ustar[i]),
                         C.c_double(bfsfc[i]), C.byref(a), C.byref(b))
        rm[i], rs[i] = a.value, b.value
    assert np.array_equal(wm, rm) and np.array_equal(ws, rs)
    g.close()


# --------------------------------------------------------------------------- init + teacher-forced steps
SMALL = {
    "cfg1": synth.CONFIGS["cfg1"],
    "cfg2": synth.scaled(synth.CONFIGS["cfg2"], 24, 16),
    "cfg3": synth.scaled(synth.CONFIGS["cfg3"], 20, 10),
    "cfg4": synth.scaled(synth.CONFIGS["cfg4"], 24, 16),
    "cfg5": synth.scaled(synth.CONFIGS["cfg5"], 16, 12),
}


@pytest.mark.parametrize("name", list(SMALL))
def test_initial_vmix_matches_oracle(name):
    P = parity.Pair(SMALL[name], numerics=0)
    P.init()
    k, w, c = _worst(P)
    assert w <= TOL_TEACHER_STRICT, (k, w)
    assert np.array_equal(P.f_gpu["kmix"], P.f_orc["kmix"])
    assert np.array_equal(P.f_gpu["old"], P.f_orc["old"]) and np.array_equal(P.f_gpu["new"], P.f_orc["new"])
    P.close()


@pytest.mark.parametrize("name", list(SMALL))
@pytest.mark.parametrize("numerics", [0, 1])
def test_teacher_forced_steps(name, numerics):
    """Each step starts from the oracle's state on both sides (pyramid iii)."""
    P = parity.Pair(SMALL[name], numerics=numerics)
    P.init()
    tol = TOL_TEACHER_STRICT if numerics == 0 else TOL_TEACHER_FAST
    nflip = 0
    for nt in range(1, 7):
        rc, rep = P.step(nt, teacher_forced=True)
        assert rc == 0
        im = P.int_mismatches()
        if numerics == 0:
            _assert_ints_exact(P, f"{name} nt={nt}")
            k, w, c = _worst(P)
            assert w <= tol, (name, nt, k, w)
        else:
            # columns whose integer outputs flipped are enumerated and excluded from the float check
            flipped = set()
            for key in ("kmix", "iter", "nreint"):
                flipped |= {i for i, _, _ in im[key]}
            nflip += len(flipped)
            keep = np.ones(P.cfg.npts, bool)
            keep[list(flipped)] = False
            for fld in parity.FLOAT_FIELDS:
                e = parity.scaled_err(P.f_gpu[fld][keep], P.f_orc[fld][keep])
                assert e <= tol, (name, nt, fld, e, sorted(flipped))
    assert nflip <= max(1, int(0.005 * 6 * P.cfg.npts)), f"too many integer flips in fast mode: {nflip}"
    P.close()


# --------------------------------------------------------------------------- free running N days
@pytest.mark.parametrize("name,nsteps", [("cfg1", 8), ("cfg2", 72), ("cfg4", 36), ("cfg5", 36)])
def test_free_running_strict(name, nsteps):
    P = parity.Pair(SMALL[name], numerics=0, nthreads=0)
    P.init()
    for nt in range(1, nsteps + 1):
        rc, rep = P.step(nt)
        assert rc == 0
    c = P.compare()
    _assert_ints_exact(P, f"{name} after {nsteps} steps")
    for fld in ("X", "Xs"):
        assert c[fld][1] <= TOL_FREE_TS, (fld, c[fld])
    for fld in ("U", "Us", "hmix", "difm", "difs", "dift", "ghat", "rho", "cp", "wX", "wU"):
        assert c[fld][1] <= TOL_FREE_OTHER, (fld, c[fld])
    P.close()


# --------------------------------------------------------------------------- straggler hand-over
@pytest.mark.parametrize("name,nsteps,budget", [("cfg1", 8, 1), ("cfg2", 40, 1), ("cfg2", 40, 4), ("cfg4", 20, 2),
                                                ("cfg5", 24, 1), ("cfg5", 24, 5), ("cfg1", 8, -1), ("cfg2", 40, -1),
                                                ("cfg4", 20, -1), ("cfg5", 24, -1)])
def test_handover_to_cooperative_kernel_is_bitwise_neutral(name, nsteps, budget):
    """kpp_gpu_set_pass_budget only moves work: with a budget of 1 every column leaves the
    per-thread kernel after its first pass and the cooperative kernel (one CTA per column) does
    the rest of the step -- compulsory passes, convergence loop, instability trap, results and
    check_profile.  The strict variant must stay bit-identical to the oracle, field by field."""
    P = parity.Pair(SMALL[name], numerics=0, nthreads=0, budget=budget)
    P.init()
    handed = 0
    for nt in range(1, nsteps + 1):
        rc, rep = P.step(nt)
        assert rc == 0
        handed += rep.n_handed_over
    if budget <= 2:
        # every stepped column needs at least three passes (and -1 skips the per-thread kernel altogether)
        assert handed == nsteps * int((P.f_orc["run_physics"] != 0).sum())
    _assert_ints_exact(P, f"{name} budget {budget}")
    _assert_bitwise(P, f"{name} budget {budget}")
    P.close()


@pytest.mark.parametrize("budget", [1, 3, -1])
def test_handover_covers_switches_trap_and_nonconvergence(budget):
    """Hand-over with every switch/branch case of test_switches_and_branches (relaxation, flux
    corrections, advection modes, bottom temperature, land mask, isothermal reset, itermax reached
    without convergence) and with the instability trap re-integrating inside the cooperative
    kernel: bit-identical to the oracle."""
    cfg = synth.scaled(synth.CONFIGS["cfg2"], 16, 8)
    for case, (consts, setup) in CASES.items():
        P = parity.Pair(cfg, numerics=0, consts=consts, setup=setup, budget=budget)
        P.init()
        for nt in range(1, 4):
            rc, rep = P.step(nt)
            assert rc == 0
            assert rep.n_handed_over > 0
        _assert_ints_exact(P, f"{case} budget {budget}")
        _assert_bitwise(P, f"{case} budget {budget}")
        P.close()
    # instability trap: 11 integrations, the later ones started by the cooperative kernel
    P = parity.Pair(synth.scaled(synth.CONFIGS["cfg2"], 12, 6), numerics=0, setup=_setup_trap, budget=budget)
    P.init()
    P.forcing(1)
    for f in (P.f_orc, P.f_gpu):
        f["sflux"][::7, 0, 4, 0] = 4000.0
    sf = np.ascontiguousarray(P.f_orc["sflux"][:, 0:6, 4, 0].T)
    P.orc.physics_driver(1)
    P.gpu.gpu.upload_forcing(sf)
    P.gpu.gpu.step(1)
    rep = P.gpu.gpu.sync()
    P.gpu.pull(driver.ALL_OUTPUTS)
    P.gpu.pull_diag()
    assert rep.n_reint > 0 and rep.n_reint_fail > 0 and rep.n_reset > 0 and rep.n_handed_over > 0
    assert P.gpu.diag["nreint"].max() == 11
    _assert_ints_exact(P, f"trap budget {budget}")
    _assert_bitwise(P, f"trap budget {budget}")
    P.close()


@pytest.mark.parametrize("margin", ["-1000", "0", "1"])
def test_buoyancy_recompute_path_is_bitwise_neutral(margin, monkeypatch):
    """The step kernel stores buoyancy only down to (expected kbl + margin) and recomputes it where the
    scan goes deeper.  With a hugely negative margin every buoyancy the scan reads is recomputed, with
    0 and 1 the boundary cases (kbl + 1 is always read) are hit: same bits as the oracle."""
    monkeypatch.setenv("KPP_BUOY_MARGIN", margin)
    for name, nsteps in (("cfg2", 30), ("cfg5", 12)):
        P = parity.Pair(SMALL[name], numerics=0, nthreads=0)
        P.init()
        for nt in range(1, nsteps + 1):
            rc, rep = P.step(nt)
            assert rc == 0
        _assert_ints_exact(P, f"{name} buoy margin {margin}")
        _assert_bitwise(P, f"{name} buoy margin {margin}")
        P.close()


def test_free_running_fast_one_day():
    P = parity.Pair(SMALL["cfg2"], numerics=1)
    P.init()
    for nt in range(1, 73):
        P.step(nt)
    im = P.int_mismatches()
    flipped = {i for key in ("kmix", "iter") for i, _, _ in im[key]}
    assert len(flipped) <= max(2, int(0.01 * P.cfg.npts)), f"integer flips (column, gpu, oracle): {im}"
    keep = np.ones(P.cfg.npts, bool)
    keep[list(flipped)] = False
    assert parity.scaled_err(P.f_gpu["X"][keep], P.f_orc["X"][keep]) <= TOL_FREE_TS
    assert parity.scaled_err(P.f_gpu["U"][keep], P.f_orc["U"][keep]) <= TOL_FREE_OTHER
    assert parity.scaled_err(P.f_gpu["hmix"][keep], P.f_orc["hmix"][keep]) <= TOL_FREE_OTHER
    P.close()


# --------------------------------------------------------------------------- switches / branches
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


def _setup_bottom(cf, f, r):
    f["bottom_temp"][:] = f["X"][:, -1, 0] + 0.01


def _setup_land(cf, f, r):
    f["run_physics"][::4] = 0
    f["l_ocean"][::4] = 0


def _setup_trap(cf, f, r):
    # absurd wind stress on a few columns: |U| >= 10 trips the instability trap, the step is
    # re-integrated with f*1.01 up to 11 times, then check_profile resets U to U_init
    f["U_init"][:, :, 0] = 0.01


def _setup_iso(cf, f, r):
    f["ocnT_clim"][:] = f["X"][:, :, 0]
    f["sal_clim"][:] = f["X"][:, :, 1]
    f["X"][::3, :, 0] = 5.0           # isothermal columns -> reset to climatology


def _setup_dd_regimes(cf, f, r):
    # Both branches of ddmix (ddmix_mod.F90:30-50), which the prescribed synthetic profiles never reach:
    # even columns cold and fresh over warm and salty (alphaDT < betaDS < 0: the diffusive branch with its two
    # exp calls, on every level), odd columns salty over fresh (salt fingering where 1 < Rrho < 1.9)
    zm = cf.zm
    even = np.arange(f["X"].shape[0]) % 2 == 0
    f["X"][even, :, 0] = -1.0 + 4.0 * (1.0 - np.exp(zm[None, :] / 150.0))
    S = np.where(even[:, None], 34.0 + 0.2 * (1.0 - np.exp(zm[None, :] / 150.0)), 35.0 + 1.0 * np.exp(zm[None, :] / 300.0))
    f["X"][:, :, 1] = S - f["Sref"][:, None]


CASES = {
    "dd_regimes": (dict(LDD=True), _setup_dd_regimes),
    "damp_curr": (dict(L_DAMP_CURR=True), None),
    "relax_sst": (dict(L_RELAX_SST=True), _setup_relax_sst),
    "relax_sst_calconly": (dict(L_RELAX_SST=True, L_RELAX_CALCONLY=True), _setup_relax_sst),
    "fcorr_twod": (dict(L_FCORR=True), _setup_fcorr),
    "no_ssref": (dict(L_SSref=False), None),
    "no_ri": (dict(LRI=False), None),
    "ldd": (dict(LDD=True), None),
    "advection_modes": (dict(), _setup_advection),
    "vary_bottom_temp": (dict(L_VARY_BOTTOM_TEMP=True), _setup_bottom),
    "land_mask": (dict(), _setup_land),
    "no_isotherm": (dict(L_NO_ISOTHERM=True, iso_bot=30, iso_thresh=0.05, have_ocnT_file=True, have_sal_file=True),
                    _setup_iso),
    "itermax_small": (dict(itermax=4), None),
}


@pytest.mark.parametrize("case", list(CASES))
def test_switches_and_branches(case):
    consts, setup = CASES[case]
    cfg = synth.scaled(synth.CONFIGS["cfg2"], 16, 8)
    P = parity.Pair(cfg, numerics=0, consts=consts, setup=setup)
    P.init()
    seen_status = 0
    min_reset = 0.0
    for nt in range(1, 5):
        rc, rep = P.step(nt)
        assert rc == 0
        _assert_ints_exact(P, f"{case} nt={nt}")
        seen_status |= int(np.bitwise_or.reduce(P.gpu.diag["status"]))
        min_reset = min(min_reset, float(P.f_gpu["reset_flag"].min()))
    k, w, c = _worst(P)
    assert w <= 1e-11, (case, k, w)
    if case == "land_mask":
        land = P.f_orc["run_physics"] == 0
        cf, f0, r = synth.make_case(cfg)
        assert np.array_equal(P.f_gpu["X"][land], f0["X"][land])      # untouched
        assert rep.n_active == int((~land).sum())
    if case == "damp_curr":
        assert P.f_gpu["dampu_flag"].max() > 0
    if case == "dd_regimes":
        d = P.f_orc["dift"][:, 1:cfg.nz] != P.f_orc["difs"][:, 1:cfg.nz]
        assert d[0::2].sum() > 1000 and d[1::2].sum() > 100        # diffusive / salt-fingering levels did occur
    if case == "no_isotherm":
        assert seen_status & capi.ST_ISO_RESET          # the isothermal columns were reset at step 1
        assert min_reset < 0                             # reset_flag = -(number of integrations) (overrides.F90:119)
    P.close()


def test_instability_trap_reintegration_and_reset():
    cfg = synth.scaled(synth.CONFIGS["cfg2"], 12, 6)
    P = parity.Pair(cfg, numerics=0, setup=_setup_trap)
    P.init()
    P.forcing(1)
    # inject the absurd stress directly at sflux level on both sides
    for f in (P.f_orc, P.f_gpu):
        f["sflux"][::7, 0, 4, 0] = 4000.0
    sf = np.ascontiguousarray(P.f_orc["sflux"][:, 0:6, 4, 0].T)
    rc = P.orc.physics_driver(1)
    P.gpu.gpu.upload_forcing(sf)
    P.gpu.gpu.step(1)
    rep = P.gpu.gpu.sync()
    P.gpu.pull(driver.ALL_OUTPUTS)
    P.gpu.pull_diag()
    assert rep.n_reint > 0 and rep.n_reint_fail > 0 and rep.n_reset > 0
    assert np.array_equal(P.gpu.diag["nreint"], P.orc.diag["nreint"])
    assert np.array_equal(P.gpu.diag["status"], P.orc.diag["status"])
    assert P.gpu.diag["nreint"].max() == 11                       # comp_iter_max + 1 (ocnstep_mod.F90:89,228)
    hit = P.gpu.diag["nreint"] == 11
    assert np.array_equal(P.f_gpu["U"][hit], P.f_gpu["U_init"][hit])   # overrides.F90:76
    k, w, c = _worst(P)
    assert w <= 1e-11, (k, w)
    P.close()


# --------------------------------------------------------------------------- golden fixtures
@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.npz"))))
def test_gpu_reproduces_golden(path):
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(ROOT, "tools", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    name = os.path.basename(path)[:-4]
    cfg, n = mg.CASES[name]
    gold = np.load(path)
    cf, f, r = synth.make_case(cfg)
    m = driver.MckppPhysics(cf, f, numerics=0, sync_mode="full")
    synth.apply_forcing(cfg, cf, f, r, 1)
    m.push_inputs()
    m.mckpp_initialize_ocean_model()
    for nt in range(1, n + 1):
        synth.apply_forcing(cfg, cf, f, r, nt)
        m.mckpp_physics_driver(nt)
        m.pull_diag()
        assert np.array_equal(m.diag["iter"], gold["iter"][nt - 1]), (name, nt)
    assert np.array_equal(f["kmix"], gold["kmix"]) and np.array_equal(f["old"], gold["old"])
    for k in ("X", "U", "hmix", "difm", "difs", "dift", "rho", "cp", "wX"):
        assert parity.scaled_err(f[k], gold[k]) <= 1e-10, (name, k)
    m.close()


# --------------------------------------------------------------------------- full-size properties
def _run_gpu(cfg, nsteps, col_offset=0, ncols=None, numerics=0, gidx=None, budget=None):
    cf, f, r = synth.make_case(cfg, col_offset=col_offset, ncols=ncols, gidx=gidx)
    m = driver.MckppPhysics(cf, f, numerics=numerics, pull=driver.SCALAR_OUTPUTS + ["X", "U"])
    if budget is not None:
        m.gpu.set_pass_budget(budget)
    synth.apply_forcing(cfg, cf, f, r, 1)
    m.push_inputs()
    m.mckpp_initialize_ocean_model()
    reps = []
    for nt in range(1, nsteps + 1):
        synth.apply_forcing(cfg, cf, f, r, nt)
        reps.append(m.mckpp_physics_driver(nt).as_dict())
    m.pull_diag()
    it = m.diag["iter"].copy()
    m.close()
    return f, reps, it


def test_full_size_partition_invariance_and_determinism():
    """cfg2 at BASELINE size (300x200 = 60,000 columns): two runs are bit-identical, and a
    2-way block partition (what 2 GPUs would own) reproduces the single-handle run bit for bit."""
    cfg = synth.CONFIGS["cfg2"]
    fa, ra, ia = _run_gpu(cfg, 3)
    fb, rb, ib = _run_gpu(cfg, 3)
    assert np.array_equal(fa["X"], fb["X"]) and np.array_equal(fa["hmix"], fb["hmix"]) and np.array_equal(ia, ib)
    half = cfg.npts // 2
    f0, _, i0 = _run_gpu(cfg, 3, col_offset=0, ncols=half)
    f1, _, i1 = _run_gpu(cfg, 3, col_offset=half, ncols=cfg.npts - half)
    assert np.array_equal(np.concatenate([f0["X"], f1["X"]]), fa["X"])
    assert np.array_equal(np.concatenate([f0["U"], f1["U"]]), fa["U"])
    assert np.array_equal(np.concatenate([f0["kmix"], f1["kmix"]]), fa["kmix"])
    assert np.array_equal(np.concatenate([i0, i1]), ia)
    assert ra[-1]["n_active"] == cfg.npts and ra[-1]["max_iter"] >= 6 and ra[-1]["n_pivot_zero"] == 0
    # sampled oracle check at full size: 64 strided columns computed alone by the oracle
    sel = np.arange(0, cfg.npts, cfg.npts // 64)[:64]
    cf, f, r = synth.make_case(cfg, gidx=sel)
    orc = oracle_lib.Oracle(cf, f)
    synth.apply_forcing(cfg, cf, f, r, 1)
    orc.initialize_ocean_model()
    for nt in range(1, 4):
        synth.apply_forcing(cfg, cf, f, r, nt)
        orc.physics_driver(nt)
    assert parity.scaled_err(fa["X"][sel], f["X"]) <= 1e-12
    assert np.array_equal(fa["kmix"][sel], f["kmix"]) and np.array_equal(ia[sel], orc.diag["iter"])


def test_full_size_every_column_through_the_cooperative_kernel():
    """cfg2 at BASELINE size with a pass budget of 1: all 60,000 columns are finished by
    kpp_coop_kernel, ~600 CTAs at a time each looping over its share of the hand-over list.
    Same bits as the default schedule, run to run as well (a missing barrier would show here)."""
    cfg = synth.CONFIGS["cfg2"]
    fa, ra, ia = _run_gpu(cfg, 2, budget=6)
    fb, rb, ib = _run_gpu(cfg, 2, budget=1)
    fc, rc_, ic = _run_gpu(cfg, 2, budget=1)
    assert rb[-1]["n_handed_over"] == cfg.npts and ra[-1]["n_handed_over"] == 0
    for fx, ix in ((fb, ib), (fc, ic)):
        for name in ("X", "U", "hmix", "kmix", "Tref", "Ssurf"):
            assert np.array_equal(fa[name], fx[name]), name
        assert np.array_equal(ia, ix)


def test_full_size_day_two_stragglers_match_the_oracle_bitwise():
    """cfg2 at BASELINE size into model day 2 (100 steps), where a few of the 60,000 columns stop
    converging and run to itermax = 200 (+1) passes.  Those columns go through the hand-over and the
    cooperative kernel; they, and a strided sample of ordinary ones, are recomputed alone by the
    oracle for all 100 steps and must agree bit for bit (state, mixed-layer depth, pass count)."""
    cfg = synth.CONFIGS["cfg2"]
    nsteps = 100
    fa, ra, ia = _run_gpu(cfg, nsteps, budget=6)
    assert max(r["max_iter"] for r in ra) >= 200, "this workload is expected to produce non-converging columns by step 100"
    assert sum(r["n_handed_over"] for r in ra) > 0
    slow = np.nonzero(ia > 6)[0]
    assert slow.size > 0 and ia.max() >= 200
    sel = np.unique(np.concatenate([slow, np.arange(0, cfg.npts, cfg.npts // 24)[:24]]))
    cf, f, r = synth.make_case(cfg, gidx=sel)
    orc = oracle_lib.Oracle(cf, f)
    synth.apply_forcing(cfg, cf, f, r, 1)
    orc.initialize_ocean_model()
    for nt in range(1, nsteps + 1):
        synth.apply_forcing(cfg, cf, f, r, nt)
        orc.physics_driver(nt)
    assert np.array_equal(ia[sel], orc.diag["iter"])
    for name in ("X", "U", "hmix", "kmix", "Tref", "Ssurf"):
        assert np.array_equal(fa[name][sel], f[name]), name


def test_columns_too_deep_for_the_cooperative_kernel_stay_with_the_step_kernel():
    """nz = 450: one column's level records no longer fit the cooperative kernel's shared memory,
    the library then never hands over (budget silently 0) and results are unchanged."""
    from dataclasses import replace
    cfg = replace(synth.scaled(synth.CONFIGS["cfg2"], 4, 3), nz=450)
    P = parity.Pair(cfg, numerics=0, budget=1)
    P.init()
    for nt in range(1, 3):
        rc, rep = P.step(nt)
        assert rep.n_handed_over == 0
    _assert_ints_exact(P, "nz=450")
    _assert_bitwise(P, "nz=450")
    P.close()


def test_rest_state_is_steady_and_bounded():
    """No forcing, uniform T/S, zero currents: nothing to mix, the state must not move
    (idempotence), whatever the column count."""
    cfg = synth.scaled(synth.CONFIGS["cfg2"], 64, 32)
    cf, f, r = synth.make_case(cfg)
    f["X"][:, :, 0] = 10.0
    f["X"][:, :, 1] = 0.0
    f["Sref"][:] = 35.0; f["SSref"][:] = 35.0; f["Ssurf"][:] = 35.0
    m = driver.MckppPhysics(cf, f, numerics=0, pull=driver.SCALAR_OUTPUTS + ["X", "U"])
    f["sflux"][:, 0:6, 4, 0] = 0.0
    f["sflux"][:, 0, 4, 0] = 1e-10
    m.push_inputs()
    m.mckpp_initialize_ocean_model()
    for nt in range(1, 4):
        m.mckpp_physics_driver(nt)
    assert np.all(np.abs(f["X"][:, :, 0] - 10.0) < 1e-9) and np.all(np.abs(f["X"][:, :, 1]) < 1e-9)
    assert np.all(np.abs(f["U"]) < 1e-6)
    m.close()


# --------------------------------------------------------------------------- error behaviour
def test_error_behaviour():
    cfg = synth.scaled(synth.CONFIGS["cfg1"], 2, 2)
    cf, f, r = synth.make_case(cfg)
    g = capi.KppGpu(cf)
    with pytest.raises(capi.KppError) as e:
        g.upload("U", np.zeros((3, 3), order="F"))            # wrong size: must be the whole Fortran array
    assert e.value.code == capi.KPP_E_INVALID
    g.close()
    cf.dims.nztmax = cf.dims.nz                               # nztmax >= nz+1 is required
    with pytest.raises(capi.KppError):
        capi.KppGpu(cf)


# --------------------------------------------------------------------------- forcing map on the device
def test_forcing_map_on_device_matches_host_map():
    cfg = synth.scaled(synth.CONFIGS["cfg2"], 20, 10)
    cf, f, r = synth.make_case(cfg)
    f["l_ocean"][::6] = 0
    f["run_physics"][:] = f["l_ocean"]
    m = driver.MckppPhysics(cf, f, numerics=0)
    m.push_inputs()
    n = cfg.npts
    rng = np.random.default_rng(7)
    raw = dict(taux=rng.normal(0, 0.1, n), tauy=rng.normal(0, 0.1, n), swf=rng.uniform(0, 900, n),
               lwf=rng.uniform(-80, 0, n), lhf=rng.uniform(-300, 0, n), shf=rng.uniform(-30, 10, n),
               rain=rng.uniform(0, 2e-4, n), snow=rng.uniform(0, 1e-5, n))
    raw["taux"][:5] = 0.0; raw["tauy"][:5] = 0.0           # taux = 1e-10 guard (fluxes_mod.F90:58-59)
    before = f["sflux"].copy(order="F")
    m.mckpp_fluxes(**raw)
    m.pull(["sflux"])
    from mckpp_f90_b200 import hostinit
    ref = {"sflux": before.copy(order="F"), "l_ocean": f["l_ocean"]}
    hostinit.fluxes_map(ref, cf.consts, **raw)
    assert np.array_equal(f["sflux"][:, 0:6, 4, 0], ref["sflux"][:, 0:6, 4, 0])
    # and the step runs from the device-side forcing
    m.mckpp_initialize_ocean_model()
    rep = m.mckpp_physics_driver(1, forcing_changed=False)
    assert rep.n_active == int((f["l_ocean"] != 0).sum())
    m.close()


def test_odd_level_count_and_async_download():
    """nz = 33 (odd, not a multiple of the pipeline depth), 35 columns (partial tile)."""
    from dataclasses import replace
    cfg = replace(synth.scaled(synth.CONFIGS["cfg2"], 7, 5), nz=33)
    P = parity.Pair(cfg, numerics=0)
    P.init()
    for nt in range(1, 5):
        rc, rep = P.step(nt)
    _assert_ints_exact(P, "nz=33")
    k, w, c = _worst(P)
    assert w <= 1e-12, (k, w)
    a = np.zeros_like(P.f_gpu["hmix"])
    P.gpu.gpu.download_async("hmix", a)
    P.gpu.gpu.sync()
    assert np.array_equal(a, P.f_orc["hmix"])
    P.close()


def test_free_running_strict_five_days_bitwise():
    """N days free running (cfg2 physics, 16x12 columns, 360 steps): the strict variant stays
    bit-identical to the oracle on T, S, U, V, hmix, kmix -- the north-star's 'after N days' bar
    with tolerance zero."""
    cfg = synth.scaled(synth.CONFIGS["cfg2"], 16, 12)
    P = parity.Pair(cfg, numerics=0, nthreads=0)
    P.init()
    for nt in range(1, 361):
        rc, rep = P.step(nt)
        assert rc == 0
    _assert_ints_exact(P, "cfg2 after 5 days")
    for fld in ("X", "U", "Xs", "Us", "hmix", "hmixd", "difm", "difs", "dift", "ghat", "rho", "cp", "wX", "wU"):
        assert np.array_equal(P.f_gpu[fld], P.f_orc[fld]), fld
    # the mixed layer did evolve (this is not a trivial steady state)
    assert P.f_gpu["hmix"].max() > 20.0 and np.abs(P.f_gpu["U"]).max() > 0.1
    P.close()


@pytest.mark.parametrize("name", ["cfg3", "cfg4", "cfg5"])
def test_full_size_other_configs(name):
    """BASELINE sizes of config 3 (44,000 columns), config 4 (700,000 columns, double diffusion: the
    north-star's scaling grid) and config 5 (60,000 columns, NZ=250 stretched, corrections + freeze
    clamp): every column steps, results deterministic, and a strided sample recomputed by the oracle
    alone agrees bit for bit."""
    cfg = synth.CONFIGS[name]
    fa, ra, ia = _run_gpu(cfg, 2)
    assert ra[-1]["n_active"] == cfg.npts and ra[-1]["n_pivot_zero"] == 0 and ra[-1]["max_iter"] >= 6
    assert np.isfinite(fa["X"]).all() and np.isfinite(fa["U"]).all()
    sel = np.arange(0, cfg.npts, cfg.npts // 48)[:48]
    cf, f, r = synth.make_case(cfg, gidx=sel)
    orc = oracle_lib.Oracle(cf, f)
    synth.apply_forcing(cfg, cf, f, r, 1)
    orc.initialize_ocean_model()
    for nt in range(1, 3):
        synth.apply_forcing(cfg, cf, f, r, nt)
        orc.physics_driver(nt)
    assert np.array_equal(fa["X"][sel], f["X"]) and np.array_equal(fa["U"][sel], f["U"])
    assert np.array_equal(fa["kmix"][sel], f["kmix"]) and np.array_equal(ia[sel], orc.diag["iter"])
    if name == "cfg5":
        assert fa["X"][:, :, 0].min() >= -1.8 and fa["freeze_flag"].max() > 0


def test_full_size_cfg4_into_the_straggler_regime():
    """cfg4 at BASELINE size (700,000 columns, LDD) far enough for columns to stop converging (they
    run to itermax and go through the hand-over to the cooperative kernel): the slow columns and a
    strided sample are recomputed alone by the oracle for every step and agree bit for bit."""
    cfg = synth.CONFIGS["cfg4"]
    nsteps = 100           # the first columns run to itermax at step 96 (profiles/r2_iter_scan_cfg4.txt)
    fa, ra, ia = _run_gpu(cfg, nsteps)
    assert ra[-1]["n_active"] == cfg.npts and all(r["n_pivot_zero"] == 0 for r in ra)
    slow_any = max(r["max_iter"] for r in ra)
    slow = np.nonzero(ia > 6)[0][:40]
    sel = np.unique(np.concatenate([slow, np.arange(0, cfg.npts, cfg.npts // 40)[:40]]))
    cf, f, r = synth.make_case(cfg, gidx=sel)
    orc = oracle_lib.Oracle(cf, f)
    synth.apply_forcing(cfg, cf, f, r, 1)
    orc.initialize_ocean_model()
    for nt in range(1, nsteps + 1):
        synth.apply_forcing(cfg, cf, f, r, nt)
        orc.physics_driver(nt)
    assert np.array_equal(ia[sel], orc.diag["iter"])
    for name in ("X", "U", "hmix", "kmix", "Tref", "Ssurf"):
        assert np.array_equal(fa[name][sel], f[name]), name
    # with 700,000 columns some always need more than the minimum of six passes by now
    assert slow_any > 6 and sum(r["n_handed_over"] for r in ra) > 0


# --------------------------------------------------------------------------- fatal path: tridiagonal zero pivot
def _setup_zero_pivot(level):
    def setup(cf, f, r):
        # solvers.F90:137-149: bet = cc(i) - cu(i)*gam(i) == 0.  Below the mixed layer of a column at
        # rest Rig >> Riinfty, so difm(i) is exactly difmiw = 1e-4 (rimix_mod.F90:96); with
        # tri(i,0) = 0 and tri(i,1) = -1e4: cu = -0, cc = 1 + (-1e4 * 1e-4) = 1 - 1 = 0 -> bet = 0
        # in the momentum matrix (U and V), while 1e-5 * -1e4 leaves the T and S matrices regular.
        cf.tri[level, 0, 0] = 0.0
        cf.tri[level, 1, 0] = -1.0e4
    return setup


@pytest.mark.parametrize("budget", [6, 1, -1])
def test_tridiagonal_zero_pivot_is_fatal_and_bitwise_the_oracle(budget):
    """The reference prints 'Algorithm for solving tridiag matrix failed' and calls MCKPP_ABORT
    (solvers.F90:140-149).  Here the column continues with bet = 1e-12 (the statement after the abort),
    carries KPP_ST_PIVOT_ZERO, and kpp_gpu_sync returns KPP_E_PIVOT_ZERO so that the host aborts -- from
    the per-thread kernel (budget 6) and from the cooperative kernel (1, -1) alike, with the same bits
    as the oracle, which flags the same columns."""
    cfg = synth.scaled(synth.CONFIGS["cfg2"], 8, 5)
    P = parity.Pair(cfg, numerics=0, setup=_setup_zero_pivot(60), budget=budget)
    P.init()
    P.forcing(1)
    rc = P.orc.physics_driver(1)
    assert rc == -1                                            # the oracle's "would have aborted"
    P.gpu.gpu.upload_forcing(np.ascontiguousarray(P.f_orc["sflux"][:, 0:6, 4, 0].T))
    P.gpu.gpu.step(1)
    with pytest.raises(capi.KppError) as e:
        P.gpu.gpu.sync()
    assert e.value.code == capi.KPP_E_PIVOT_ZERO
    rep = P.gpu.gpu.last_report
    assert rep.n_pivot_zero == cfg.npts
    P.gpu.pull(driver.ALL_OUTPUTS)
    P.gpu.pull_diag()
    assert (P.gpu.diag["status"] & capi.ST_PIVOT_ZERO).all() and (P.orc.diag["status"] & capi.ST_PIVOT_ZERO).all()
    _assert_ints_exact(P, f"zero pivot budget {budget}")
    _assert_bitwise(P, f"zero pivot budget {budget}")
    P.close()


def test_zero_pivot_of_an_unsynced_earlier_step_is_not_lost():
    """kpp_gpu_step is asynchronous: with two steps queued and one sync the report is the second
    step's, but a zero pivot in the first must still make that sync fail (sticky flag)."""
    cfg = synth.scaled(synth.CONFIGS["cfg2"], 8, 5)
    cf, f, r = synth.make_case(cfg)
    _setup_zero_pivot(60)(cf, f, r)
    g = capi.KppGpu(cf)
    sf = synth.apply_forcing(cfg, cf, f, r, 1)
    for name in driver.INPUT_FIELDS:
        g.upload(name, f[name])
    g.init_vmix()
    g.upload_forcing(sf)
    g.step(1)
    # repair the matrix on the device side is impossible (tri is fixed at creation): instead make the
    # second step inactive for every column, so its own report carries no pivot at all
    off = np.zeros(cfg.npts, np.int32)
    g.upload("run_physics", off)
    g.step(2)
    with pytest.raises(capi.KppError) as e:
        g.sync()
    assert e.value.code == capi.KPP_E_PIVOT_ZERO
    assert g.last_report.n_pivot_zero == 0 and g.last_report.n_active == 0      # the report is step 2's
    g.sync()                                                                     # reported once
    g.close()


def test_input_validation_follows_the_reference():
    cfg = synth.scaled(synth.CONFIGS["cfg1"], 4, 4)
    cf, f, r = synth.make_case(cfg)
    g = capi.KppGpu(cf)
    bad = f["jerlov"].copy(); bad[3] = 6
    with pytest.raises(capi.KppError):
        g.upload("jerlov", bad)                  # indexes the five water types (swfrac_mod.F90:28-34)
    bad[3] = 0
    with pytest.raises(capi.KppError):
        g.upload("jerlov", bad)
    # 'mode out of range' (solvers.F90:320) only for the entries rhsmod visits: slots beyond nmodeadv(:,2)
    # are never initialised in the reference and must not matter
    for name in driver.INPUT_FIELDS:
        g.upload(name, f[name])
    g.init_vmix()
    mode = f["modeadv"].copy(order="F"); nmode = f["nmodeadv"].copy(order="F")
    mode[:, :, 1] = 99                            # garbage everywhere ...
    nmode[:, 1] = 0                               # ... but no entry is active
    g.upload("modeadv", mode); g.upload("nmodeadv", nmode)
    g.upload_forcing(synth.apply_forcing(cfg, cf, f, r, 1))
    g.step(1); g.sync()
    nmode[5, 1] = 2; mode[5, 0, 1] = 3; mode[5, 1, 1] = 9        # an active out-of-range mode
    g.upload("nmodeadv", nmode); g.upload("modeadv", mode)
    with pytest.raises(capi.KppError) as e:
        g.step(2)
    assert e.value.code == capi.KPP_E_INVALID
    mode[5, 1, 1] = 7
    g.upload("modeadv", mode)
    g.step(2); g.sync()
    g.close()
