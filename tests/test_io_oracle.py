"""CPU tests of the oracle for the stages either side of the step: the numpy
restatement against hand-computed cases, and the host mirror of the weight computation."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, ROOT)
import io_oracle
from mckpp_f90_b200 import capi, hostinit, synth
from mckpp_f90_b200.fields import allocate_3d_fields


def _fields(nz=6, npts=5, seed=3):
    cfg = synth.scaled(synth.CONFIGS["cfg1"], npts, 1)
    cf, f, r = synth.make_case(cfg)
    rng = np.random.default_rng(seed)
    for name, arr in f.items():
        if arr.dtype == np.float64:
            arr[...] = rng.standard_normal(arr.shape)
    f["old"][:] = rng.integers(0, 2, f["old"].shape)
    f["new"][:] = 1 - f["old"]
    return cf, f


def test_diagnostic_set_shapes_and_reshuffles():
    cf, f = _fields()
    nz, nzp1, npts = cf.dims.nz, cf.dims.nz + 1, cf.dims.npts
    out = io_oracle.xios_diagnostic_output(f, nz)
    assert len(out) == 34
    for name in ("u", "S", "B", "wu", "wB", "wTnt", "difm", "rho", "cp", "Rig", "dbloc", "Shsq", "fcorr_z"):
        assert out[name].shape == (npts, nzp1), name
    # S = X(:,k,2)+Sref (xios_io.F90:94-97)
    assert np.array_equal(out["S"][:, 3], f["X"][:, 3, 1] + f["Sref"])
    # diffusivities: zero top row, difm(:,1:NZ) below it (:120-123); index 0 of the 0-based array is dropped
    assert np.all(out["difm"][:, 0] == 0.0)
    assert np.array_equal(out["difm"][:, 1:], f["difm"][:, 1:nz + 1])
    assert np.array_equal(out["dift"][:, nz], f["dift"][:, nz])
    # dbloc padded with a zero bottom row (:148-150)
    assert np.array_equal(out["dbloc"][:, :nz], f["dbloc"]) and np.all(out["dbloc"][:, nz] == 0.0)
    # rho(:,1:NZP1) of a 0-based array; wX(:,0:NZ,NSP1)
    assert np.array_equal(out["rho"], f["rho"][:, 1:nzp1 + 1])
    assert np.array_equal(out["wB"], f["wX"][:, :nz + 1, 2])
    assert np.array_equal(out["PminusE_in"], f["sflux"][:, 5, 4, 0])
    assert np.array_equal(out["comp_flag"], f["reset_flag"])


def test_restart_set():
    cf, f = _fields()
    nz, nzp1, npts = cf.dims.nz, cf.dims.nz + 1, cf.dims.npts
    out = io_oracle.xios_restart_output(f, nz)
    assert len(out) == 19
    assert np.array_equal(out["S"], f["X"][:, :, 1])                 # no Sref in the restart
    assert out["old"].dtype == np.float64 and np.array_equal(out["old"], f["old"])
    assert out["Vs"].shape == (npts, nzp1, 2) and np.array_equal(out["Vs"][:, :, 1], f["Us"][:, :, 1, 1])
    assert np.array_equal(out["Ts"][:, 2, 0], f["Xs"][:, 2, 0, 0])
    assert out["hmixd"].shape == (npts, 2)


def test_interpolation_weights_hand_cases():
    # 30-day records (ndtupd*dto/spd = 30): records valid at 15, 45, 75 ... days
    w = io_oracle.boundary_interp_weights
    # day 20: bracket 15..45, 5 days past prev
    assert w(20.0, 2160, 1200.0, 86400.0, 360) == (15, 45, (30 - 5) / 30, 1 - (30 - 5) / 30)
    # day 20.9 truncates to 20 (INTEGER true_time)
    assert w(20.9, 2160, 1200.0, 86400.0, 360) == w(20.0, 2160, 1200.0, 86400.0, 360)
    # day 3: prev would be -15 -> wraps by the period, weight from the unwrapped distance (:32-34)
    pt, nt, pw, nw = w(3.0, 2160, 1200.0, 86400.0, 360)
    assert (pt, nt) == (345, 375) and pw == (30 - 18) / 30 and nw == 1 - pw
    # exactly on a record: weight 1 on prev
    assert w(45.0, 2160, 1200.0, 86400.0, 360)[2:] == (1.0, 0.0)
    # 5-day records: ndays_upd/2 = 2.5 is not an integer, prev_time truncates (7.5 -> 7)
    pt, nt, pw, nw = w(9.0, 360, 1200.0, 86400.0, 360)
    assert (pt, nt) == (7, 12) and pw == (5 - 2) / 5
    # the host mirror is the same function
    for args in ((20.0, 2160, 1200.0, 86400.0, 360), (3.0, 2160, 1200.0, 86400.0, 360), (9.0, 360, 1200.0, 86400.0, 360),
                 (123.4, 72, 1200.0, 86400.0, 360)):
        assert hostinit.boundary_interp_weights(*args) == w(*args)


def test_interpolate_is_next_times_weight_plus_prev_times_weight():
    rng = np.random.default_rng(0)
    p, n = rng.standard_normal((7, 5)), rng.standard_normal((7, 5))
    out = io_oracle.boundary_interpolate(p, n, 0.3, 0.7)
    assert np.array_equal(out, n * 0.7 + p * 0.3)


def test_output_enum_matches_names_the_reference_sends():
    ids = capi.out_ids()
    assert ids["KPP_OUT__COUNT"] == 34 + 19
    assert ids["KPP_OUT_R_UVEL"] == 34
    cf, f = _fields()
    diag = list(io_oracle.xios_diagnostic_output(f, cf.dims.nz))
    rest = list(io_oracle.xios_restart_output(f, cf.dims.nz))
    # the library's name table (kpp_gpu_output_name) follows the same order as the reference's sends
    L = capi.load()
    assert [L.kpp_gpu_output_name(i).decode() for i in range(34)] == diag
    assert [L.kpp_gpu_output_name(i).decode() for i in range(34, 53)] == rest
