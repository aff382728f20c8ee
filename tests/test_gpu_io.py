"""GPU parity of the stages SURVEY.md 8(f) ranks next, through the C ABI:
(f2) the output sets of mckpp_xios_diagnostic_output / _restart_output packed on the device,
(f4) climatology time interpolation on the device.  Bit-exact against oracle/io_oracle.py."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import io_oracle
import parity
from mckpp_f90_b200 import capi, hostinit, synth

pytestmark = pytest.mark.gpu


def _stepped_pair(nsteps=5, **kw):
    cfg = synth.scaled(synth.CONFIGS["cfg5"], 12, 5)       # corrections on: tinc_fcorr, scorr ... are non-trivial
    P = parity.Pair(cfg, numerics=0, **kw)
    P.init()
    for nt in range(1, nsteps + 1):
        P.step(nt)
    return P


def test_diagnostic_and_restart_sets_match_reference_packing_bitwise():
    P = _stepped_pair()
    nz = P.cfg.nz
    want = io_oracle.xios_diagnostic_output(P.f_orc, nz)
    got = P.gpu.mckpp_xios_diagnostic_output()
    assert list(got) == list(want)
    for name in want:
        assert got[name].shape == want[name].shape, name
        assert np.array_equal(got[name], want[name]), name
    # the reshuffles are really there
    assert np.all(got["difm"][:, 0] == 0) and np.abs(got["difm"][:, 1:]).max() > 0
    assert np.all(got["dbloc"][:, nz] == 0) and np.abs(got["dbloc"][:, :nz]).max() > 0
    assert np.abs(got["S"] - P.f_orc["X"][:, :, 1]).min() > 30.0          # Sref added
    want_r = io_oracle.xios_restart_output(P.f_orc, nz)
    got_r = P.gpu.mckpp_xios_restart_output()
    assert list(got_r) == list(want_r)
    for name in want_r:
        g = got_r[name].reshape(want_r[name].shape, order="F")           # (npts, 2*nzp1) block == (npts,nzp1,2)
        assert np.array_equal(g, want_r[name]), name
    assert set(np.unique(got_r["old"])) <= {0.0, 1.0} and np.array_equal(got_r["old"] + got_r["new"], np.ones(P.cfg.npts))
    P.close()


def test_pack_output_single_calls_and_errors():
    P = _stepped_pair(nsteps=2)
    g = P.gpu.gpu
    ids = capi.out_ids()
    s = g.pack_output(ids["KPP_OUT_S"])
    assert np.array_equal(s, P.f_orc["X"][:, :, 1] + P.f_orc["Sref"][:, None])
    pinned = capi.pinned_empty((P.cfg.npts, P.cfg.nz + 1))
    g.pack_output(ids["KPP_OUT_T"], host=pinned, sync=False)
    g.sync()
    assert np.array_equal(np.asarray(pinned), P.f_orc["X"][:, :, 0])
    with pytest.raises(capi.KppError):
        g.pack_output(ids["KPP_OUT__COUNT"])
    with pytest.raises(capi.KppError):
        g.pack_output(ids["KPP_OUT_HMIX"], host=np.empty(P.cfg.npts + 1))
    assert g.L.kpp_gpu_output_rows(g.h, ids["KPP_OUT_R_US"]) == 2 * (P.cfg.nz + 1)
    assert g.L.kpp_gpu_output_rows(g.h, ids["KPP_OUT_HMIX"]) == 1
    P.close()


def test_climatology_interpolation_on_device_and_its_use_by_the_step():
    P = _stepped_pair(nsteps=1)
    npts, nzp1 = P.cfg.npts, P.cfg.nz + 1
    rng = np.random.default_rng(11)
    base_t, base_s = P.f_orc["ocnT_clim"].copy(), P.f_orc["sal_clim"].copy()
    prev_t = np.asfortranarray(base_t + 0.3 * rng.standard_normal((npts, nzp1)))
    next_t = np.asfortranarray(base_t + 0.3 * rng.standard_normal((npts, nzp1)))
    prev_s = np.asfortranarray(base_s + 0.05 * rng.standard_normal((npts, nzp1)))
    next_s = np.asfortranarray(base_s + 0.05 * rng.standard_normal((npts, nzp1)))
    with pytest.raises(capi.KppError):
        P.gpu.gpu.blend_clim("ocnT_clim", 0.5, 0.5)                      # no records yet
    # 30-day records, model day 20.9: weights with the reference's INTEGER truncations
    pt, nt, pw, nw = hostinit.boundary_interp_weights(20.9, 2160, 1200.0, 86400.0, 360)
    assert (pt, nt) == (15, 45)
    P.gpu.mckpp_boundary_interpolate("ocnT_clim", pw, nw, prev_t, next_t)
    P.gpu.mckpp_boundary_interpolate("sal_clim", pw, nw, prev_s, next_s)
    got_t = np.zeros((npts, nzp1), order="F")
    got_s = np.zeros((npts, nzp1), order="F")
    P.gpu.gpu.download("ocnT_clim", got_t)
    P.gpu.gpu.download("sal_clim", got_s)
    assert np.array_equal(got_t, io_oracle.boundary_interpolate(prev_t, next_t, pw, nw))
    assert np.array_equal(got_s, io_oracle.boundary_interpolate(prev_s, next_s, pw, nw))
    # later in the same bracket: re-blend the resident records, no upload
    pt2, nt2, pw2, nw2 = hostinit.boundary_interp_weights(33.0, 2160, 1200.0, 86400.0, 360)
    assert (pt2, nt2) == (15, 45) and pw2 != pw
    P.gpu.mckpp_boundary_interpolate("ocnT_clim", pw2, nw2)
    P.gpu.gpu.download("ocnT_clim", got_t)
    want_t = io_oracle.boundary_interpolate(prev_t, next_t, pw2, nw2)
    assert np.array_equal(got_t, want_t)
    # and the physics relaxes towards it: same climatology on the oracle side -> same bits after steps
    P.f_orc["ocnT_clim"][...] = want_t
    P.f_orc["sal_clim"][...] = io_oracle.boundary_interpolate(prev_s, next_s, pw, nw)
    P.f_gpu["ocnT_clim"][...] = P.f_orc["ocnT_clim"]
    P.f_gpu["sal_clim"][...] = P.f_orc["sal_clim"]
    for ntime in range(2, 6):
        P.step(ntime)
    run = P.f_orc["run_physics"] != 0
    for fld in parity.FLOAT_FIELDS:
        assert np.array_equal(P.f_gpu[fld][run], P.f_orc[fld][run]), fld
    assert np.abs(P.f_gpu["tinc_fcorr"]).max() > 0
    P.close()
