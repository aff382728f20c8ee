"""Multi-GPU inside the library (kpp_gpu_create_multi) and the asynchronous output ring, through the
C ABI on the B200 box.  A group handle block-partitions the columns over several devices; on a box
with one GPU the same code path is exercised with several parts on device 0 (own streams, own
mirrors).  The bar: bit-identical to a single-device handle and to the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import oracle_lib
import parity
from mckpp_f90_b200 import capi, driver, synth
from mckpp_f90_b200.fields import copy_fields


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if capi.load().kpp_gpu_device_count() < 1:
        pytest.fail("no CUDA device: the product path has no CPU fallback and these tests need the B200")


def _devices(n):
    nd = capi.load().kpp_gpu_device_count()
    return [i % nd for i in range(n)]


def _run(cfg, nsteps, devices=None, consts=None, setup=None, ring=None, budget=6):
    cf, f, r = synth.make_case(cfg)
    for k, v in (consts or {}).items():
        setattr(cf.consts, k, v)
    if setup:
        setup(cf, f, r)
    m = driver.MckppPhysics(cf, f, numerics=0, sync_mode="full", devices=devices)
    m.gpu.set_pass_budget(budget)      # domains this small default to the cooperative kernel: pin the per-thread one
    synth.apply_forcing(cfg, cf, f, r, 1)
    m.push_inputs()
    m.mckpp_initialize_ocean_model()
    reps, ringed = [], []
    if ring is not None:
        m.gpu.output_ring_create(ring, depth=2)
    for nt in range(1, nsteps + 1):
        synth.apply_forcing(cfg, cf, f, r, nt)
        if ring is None:
            reps.append(m.mckpp_physics_driver(nt).as_dict())
        else:
            # asynchronous protocol: step n+1 is launched before the blocks of step n are read
            m.gpu.upload_forcing(np.ascontiguousarray(f["sflux"][:, 0:6, 4, 0].T))
            m.gpu.step(nt)
            slot = m.gpu.output_ring_submit()
            if nt > 1:
                got = m.gpu.output_ring_wait(prev_slot)
                ringed.append({k: v.copy() for k, v in got.items()})
            prev_slot = slot
            reps.append(m.gpu.sync().as_dict())
    if ring is not None:
        got = m.gpu.output_ring_wait(prev_slot)
        ringed.append({k: v.copy() for k, v in got.items()})
        m.pull(driver.ALL_OUTPUTS)
    m.pull_diag()
    diag = {k: v.copy() for k, v in m.diag.items()}
    parts = m.gpu.parts()
    packed = m.mckpp_xios_diagnostic_output()
    restart = m.mckpp_xios_restart_output()
    m.close()
    return f, reps, diag, parts, packed, restart, ringed


@pytest.mark.parametrize("name,nparts,nx,ny", [("cfg2", 2, 25, 13), ("cfg4", 3, 20, 11), ("cfg5", 4, 13, 7), ("cfg1", 8, 4, 4)])
def test_group_handle_is_bitwise_the_single_handle_and_the_oracle(name, nparts, nx, ny):
    """Column counts that are not multiples of 32 per part, more parts than tiles (cfg1: 16 columns on 8
    'devices' collapse to one part), every field, every packed output block, reports added up."""
    cfg = synth.scaled(synth.CONFIGS[name], nx, ny)
    nsteps = 6
    fa, ra, da, pa, oa, sa, _ = _run(cfg, nsteps)
    fb, rb, db, pb, ob, sb, _ = _run(cfg, nsteps, devices=_devices(nparts))
    assert len(pa) == 1 and sum(n for _, _, n in pb) == cfg.npts
    assert [c0 for _, c0, _ in pb] == list(np.cumsum([0] + [n for _, _, n in pb[:-1]]))
    if cfg.npts > 32 * nparts:
        assert len(pb) == nparts
    for fld in parity.FLOAT_FIELDS + parity.INT_FIELDS:
        assert np.array_equal(fa[fld], fb[fld]), fld
    for k in da:
        assert np.array_equal(da[k], db[k]), k
    for k in oa:
        assert np.array_equal(oa[k], ob[k]), ("diagnostic output", k)
    for k in sa:
        assert np.array_equal(sa[k], sb[k]), ("restart output", k)
    for x, y in zip(ra, rb):
        for key in ("n_active", "n_long_iter", "n_reint", "n_reset", "n_pivot_zero", "max_iter", "sum_iter"):
            assert x[key] == y[key], key
    # and against the oracle
    cf, f, r = synth.make_case(cfg)
    orc = oracle_lib.Oracle(cf, f)
    synth.apply_forcing(cfg, cf, f, r, 1)
    orc.initialize_ocean_model()
    for nt in range(1, nsteps + 1):
        synth.apply_forcing(cfg, cf, f, r, nt)
        orc.physics_driver(nt)
    for fld in ("X", "U", "Xs", "Us", "hmix", "kmix", "difm", "difs", "wX", "rho"):
        assert np.array_equal(fb[fld], f[fld]), fld
    assert np.array_equal(db["iter"], orc.diag["iter"])


def test_group_handle_switch_cases_and_forcing_map():
    """Land mask, advection modes, bottom temperature and the device-side forcing map through a 3-part group."""
    cfg = synth.scaled(synth.CONFIGS["cfg2"], 19, 7)

    def setup(cf, f, r):
        n = f["nmodeadv"].shape[0]
        f["run_physics"][::4] = 0; f["l_ocean"][::4] = 0
        f["nmodeadv"][:, 1] = (np.arange(n) % 4)
        for m_ in range(6):
            f["modeadv"][:, m_, 1] = 1 + (np.arange(n) + m_) % 7
            f["advection"][:, m_, 1] = 1e-6 * (r[:, m_] - 0.5)
        f["bottom_temp"][:] = f["X"][:, -1, 0] + 0.01

    consts = dict(L_VARY_BOTTOM_TEMP=True, L_DAMP_CURR=True)
    fa, ra, da, *_ = _run(cfg, 4, consts=consts, setup=setup)
    fb, rb, db, *_ = _run(cfg, 4, devices=_devices(3), consts=consts, setup=setup)
    for fld in parity.FLOAT_FIELDS + parity.INT_FIELDS:
        assert np.array_equal(fa[fld], fb[fld]), fld
    assert np.array_equal(da["status"], db["status"]) and ra[-1]["n_active"] == rb[-1]["n_active"] < cfg.npts
    # forcing map on the device (kpp_gpu_upload_fluxes) through the group
    cf, f, r = synth.make_case(cfg)
    m = driver.MckppPhysics(cf, f, numerics=0, devices=_devices(3))
    m.push_inputs()
    rng = np.random.default_rng(3)
    n = cfg.npts
    raw = dict(taux=rng.normal(0, 0.1, n), tauy=rng.normal(0, 0.1, n), swf=rng.uniform(0, 900, n),
               lwf=rng.uniform(-80, 0, n), lhf=rng.uniform(-300, 0, n), shf=rng.uniform(-30, 10, n),
               rain=rng.uniform(0, 2e-4, n), snow=rng.uniform(0, 1e-5, n))
    before = f["sflux"].copy(order="F")
    m.mckpp_fluxes(**raw)
    m.pull(["sflux"])
    from mckpp_f90_b200 import hostinit
    ref = {"sflux": before, "l_ocean": f["l_ocean"]}
    hostinit.fluxes_map(ref, cf.consts, **raw)
    assert np.array_equal(f["sflux"][:, 0:6, 4, 0], ref["sflux"][:, 0:6, 4, 0])
    m.close()


@pytest.mark.parametrize("devices", [None, "3parts"])
def test_output_ring_delivers_the_same_blocks_one_step_later(devices):
    """The asynchronous ring against the synchronous packing: every block of every step identical, while the
    host reads step n only after step n+1 has been launched (depth 2)."""
    cfg = synth.scaled(synth.CONFIGS["cfg4"], 21, 9)
    ids = list(range(capi.out_ids()["KPP_OUT_R_UVEL"]))          # the 34 diagnostic blocks
    dev = None if devices is None else _devices(3)
    nsteps = 5
    *_, ringed = _run(cfg, nsteps, devices=dev, ring=ids)
    assert len(ringed) == nsteps
    # reference: synchronous packing after each step on a plain handle
    cf, f, r = synth.make_case(cfg)
    m = driver.MckppPhysics(cf, f, numerics=0)
    synth.apply_forcing(cfg, cf, f, r, 1)
    m.push_inputs()
    m.mckpp_initialize_ocean_model()
    names = m.gpu.output_ids(restart=False)
    for nt in range(1, nsteps + 1):
        synth.apply_forcing(cfg, cf, f, r, nt)
        m.mckpp_physics_driver(nt)
        want = m.mckpp_xios_diagnostic_output()
        for oid, name in names.items():
            assert np.array_equal(ringed[nt - 1][oid], want[name]), (nt, name)
    m.close()


def test_real_multi_gpu_full_size_when_the_box_has_more_than_one():
    """On a multi-GPU box: cfg3 at BASELINE size split over all devices inside the library equals the
    single-device run bit for bit (on a one-GPU box the partition runs as parts on device 0)."""
    nd = capi.load().kpp_gpu_device_count()
    cfg = synth.CONFIGS["cfg3"]
    nparts = max(2, min(nd, 8))

    def run(devices):
        cf, f, r = synth.make_case(cfg)
        m = driver.MckppPhysics(cf, f, numerics=0, pull=driver.SCALAR_OUTPUTS + ["X", "U"], devices=devices)
        synth.apply_forcing(cfg, cf, f, r, 1)
        m.push_inputs()
        m.mckpp_initialize_ocean_model()
        for nt in range(1, 4):
            synth.apply_forcing(cfg, cf, f, r, nt)
            rep = m.mckpp_physics_driver(nt)
        m.pull_diag()
        it = m.diag["iter"].copy()
        parts = m.gpu.parts()
        m.close()
        return f, it, rep, parts

    fa, ia, ra, _ = run(None)
    fb, ib, rb, parts = run(_devices(nparts))
    assert len(parts) == nparts and len({d for d, _, _ in parts}) == min(nd, nparts)
    for fld in ("X", "U", "hmix", "kmix", "Tref", "Ssurf", "old", "new"):
        assert np.array_equal(fa[fld], fb[fld]), fld
    assert np.array_equal(ia, ib) and ra.n_active == rb.n_active == cfg.npts and ra.sum_iter == rb.sum_iter


# --------------------------------------------------------------------------- asynchronous stragglers
def _queued_run(cfg, nsteps, async_on, budget, sync_every, devices=None, consts=None):
    """Steps queued on device-resident forcing slots; kpp_gpu_sync only every `sync_every` steps."""
    cf, f, r = synth.make_case(cfg)
    for k, v in (consts or {}).items():
        setattr(cf.consts, k, v)
    m = driver.MckppPhysics(cf, f, numerics=0, sync_mode="full", devices=devices)
    g = m.gpu
    g.set_pass_budget(budget)
    g.set_async_stragglers(async_on)
    synth.apply_forcing(cfg, cf, f, r, 1)
    m.push_inputs()
    m.mckpp_initialize_ocean_model()
    g.reserve_forcing_slots(nsteps)
    for nt in range(1, nsteps + 1):
        g.upload_forcing_slot(nt - 1, synth.apply_forcing(cfg, cf, f, r, nt))
    handed, max_iter = 0, 0
    for nt in range(1, nsteps + 1):
        g.select_forcing_slot(nt - 1)
        g.step(nt)
        if nt % sync_every == 0 or nt == nsteps:
            rep = g.sync()
            handed += rep.n_handed_over
            max_iter = max(max_iter, rep.max_iter)
    m.pull(driver.ALL_OUTPUTS)
    m.pull_diag()
    diag = {k: v.copy() for k, v in m.diag.items()}
    m.close()
    return f, diag, handed, max_iter


@pytest.mark.parametrize("name,nx,ny,nsteps,budget,sync_every", [("cfg2", 24, 16, 30, 1, 1000), ("cfg2", 24, 16, 30, 1, 4),
                                                                  ("cfg4", 20, 11, 20, 2, 7), ("cfg5", 13, 7, 12, 1, 5),
                                                                  ("cfg2", 25, 13, 40, 6, 1000)])
def test_async_stragglers_are_bitwise_neutral(name, nx, ny, nsteps, budget, sync_every):
    """kpp_gpu_set_async_stragglers moves the hand-over continuation to a second stream and the following
    steps of those columns to a third (the lane), without a sync in between.  With a pass budget of 1 EVERY
    column is handed over at its first step and then lives in the lane until the next join: all of the
    machinery (finish on B, lane steps on C from pass 0, joins that empty the lane) carries the whole run.
    Bit-identical to the oracle and to the synchronous schedule, for every field."""
    cfg = synth.scaled(synth.CONFIGS[name], nx, ny)
    fa, da, ha, _ = _queued_run(cfg, nsteps, True, budget, sync_every)
    fb, db, hb, _ = _queued_run(cfg, nsteps, False, budget, 1)
    for fld in parity.FLOAT_FIELDS + parity.INT_FIELDS:
        assert np.array_equal(fa[fld], fb[fld]), fld
    for k in da:
        assert np.array_equal(da[k], db[k]), k
    cf, f, r = synth.make_case(cfg)
    orc = oracle_lib.Oracle(cf, f)
    synth.apply_forcing(cfg, cf, f, r, 1)
    orc.initialize_ocean_model()
    for nt in range(1, nsteps + 1):
        synth.apply_forcing(cfg, cf, f, r, nt)
        orc.physics_driver(nt)
    for fld in parity.FLOAT_FIELDS + parity.INT_FIELDS:
        assert np.array_equal(fa[fld], f[fld]), fld
    assert np.array_equal(da["iter"], orc.diag["iter"]) and np.array_equal(da["status"], orc.diag["status"])


def test_async_stragglers_group_handle_and_trap():
    """The same through a 3-part group handle, with the instability trap re-integrating inside the lane."""
    cfg = synth.scaled(synth.CONFIGS["cfg2"], 19, 7)
    fa, da, ha, _ = _queued_run(cfg, 12, True, 1, 5, devices=_devices(3), consts=dict(L_DAMP_CURR=True))
    fb, db, hb, _ = _queued_run(cfg, 12, False, 6, 1, consts=dict(L_DAMP_CURR=True))
    for fld in parity.FLOAT_FIELDS + parity.INT_FIELDS:
        assert np.array_equal(fa[fld], fb[fld]), fld
    assert np.array_equal(da["iter"], db["iter"])


def test_async_stragglers_full_size_day_two():
    """cfg2 at BASELINE size, 110 queued steps with ONE sync at the end: by then columns have been running to
    itermax for a dozen steps and live in the lane.  Every column equals the synchronous schedule; the slow
    ones and a strided sample equal the oracle."""
    cfg = synth.CONFIGS["cfg2"]
    nsteps = 110
    fa, da, ha, mxa = _queued_run(cfg, nsteps, True, 6, 10 ** 6)
    fb, db, hb, mxb = _queued_run(cfg, nsteps, False, 6, 1)
    assert mxb >= 200 and hb > 0
    for fld in ("X", "U", "Xs", "Us", "hmix", "kmix", "Tref", "Ssurf", "old", "new", "difm", "difs", "wX"):
        assert np.array_equal(fa[fld], fb[fld]), fld
    assert np.array_equal(da["iter"], db["iter"]) and np.array_equal(da["status"], db["status"])
    slow = np.nonzero(da["iter"] > 6)[0]
    sel = np.unique(np.concatenate([slow, np.arange(0, cfg.npts, cfg.npts // 16)[:16]]))
    cf, f, r = synth.make_case(cfg, gidx=sel)
    orc = oracle_lib.Oracle(cf, f)
    synth.apply_forcing(cfg, cf, f, r, 1)
    orc.initialize_ocean_model()
    for nt in range(1, nsteps + 1):
        synth.apply_forcing(cfg, cf, f, r, nt)
        orc.physics_driver(nt)
    assert np.array_equal(da["iter"][sel], orc.diag["iter"])
    for fld in ("X", "U", "hmix", "kmix"):
        assert np.array_equal(fa[fld][sel], f[fld]), fld


# --------------------------------------------------------------------------- stray writes (stand-in for memcheck)
@pytest.mark.parametrize("name,nx,ny,nz,budget,async_on,parts", [
    ("cfg2", 7, 5, 33, 6, False, None),          # 35 columns (partial tile), odd level count
    ("cfg2", 8, 6, 100, 1, False, None),         # every column through the cooperative kernel
    ("cfg5", 6, 4, 250, 6, False, None),         # flux corrections staged through the pipeline
    ("cfg4", 13, 7, 100, 2, True, [0, 0, 0]),    # asynchronous lane, 3-part group of 91 columns, LDD
])
def test_no_kernel_writes_outside_its_arrays(name, nx, ny, nz, budget, async_on, parts, monkeypatch):
    """compute-sanitizer is closed on this pool, so the library carries its own check: with KPP_GUARD=1 every
    device array sits between 64 KB canary zones; after steps, hand-overs, the output ring and the packed output
    sets no zone may have changed."""
    from dataclasses import replace
    monkeypatch.setenv("KPP_GUARD", "1")
    cfg = replace(synth.scaled(synth.CONFIGS[name], nx, ny), nz=nz)
    cf, f, r = synth.make_case(cfg)
    m = driver.MckppPhysics(cf, f, numerics=0, devices=parts)
    g = m.gpu
    g.set_pass_budget(budget)
    g.set_async_stragglers(async_on)
    synth.apply_forcing(cfg, cf, f, r, 1)
    m.push_inputs()
    m.mckpp_initialize_ocean_model()
    nst = 4
    g.reserve_forcing_slots(nst)
    for nt in range(1, nst + 1):
        g.upload_forcing_slot(nt - 1, synth.apply_forcing(cfg, cf, f, r, nt))
    g.output_ring_create(list(range(capi.out_ids()["KPP_OUT_R_UVEL"])), depth=2)
    for nt in range(1, nst + 1):
        g.select_forcing_slot(nt - 1)
        g.step(nt)
    g.sync()
    g.output_ring_wait(g.output_ring_submit())
    m.mckpp_xios_diagnostic_output()
    m.mckpp_xios_restart_output()
    m.pull(driver.ALL_OUTPUTS)
    assert g.check_guards() == 0
    m.close()
