"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy) of the two host-side stages either side of
the column step that SURVEY.md 8(f2)/(f4) rank next.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline leg may import this; the product never does.

  xios_diagnostic_output / xios_restart_output   src/mckpp_xios_io.F90:72-207, 406-431
  boundary_interp_weights / boundary_interpolate src/mckpp_boundary_interpolate.F90:14-123

Parity pinning: the reference holds no golden vectors for these stages ("parity unpinned" by the
reference; pinned by this literal restatement and the hand-computed cases in tests/test_io_oracle.py).
Arrays are the reference's own (Fortran shapes, first extent npts; 0-based Fortran lower bounds are
noted where they shift the numpy index)."""
import math

import numpy as np


def xios_diagnostic_output(f, nz):
    """mckpp_xios_diagnostic_output (xios_io.F90:72-207): {xios field id: array as sent}.
    `cplwght` (:186-192) is a regridding of a host-only array and is left to the host."""
    nzp1 = nz + 1
    npts = f["U"].shape[0]
    out = {}
    out["u"] = f["U"][:, :, 0].copy()                         # :84
    out["v"] = f["U"][:, :, 1].copy()                         # :87
    out["T"] = f["X"][:, :, 0].copy()                         # :90
    temp_2d = np.zeros((npts, nzp1), order="F")
    for k in range(nzp1):                                     # :93-96
        temp_2d[:, k] = f["X"][:, k, 1] + f["Sref"][:]
    out["S"] = temp_2d.copy()
    out["B"] = f["buoy"][:, 0:nzp1].copy()                    # buoy(:,1:NZP1) :100
    out["wu"] = f["wU"][:, 0:nz + 1, 0].copy()                # wU(:,0:NZ,1) :103
    out["wv"] = f["wU"][:, 0:nz + 1, 1].copy()
    out["wT"] = f["wX"][:, 0:nz + 1, 0].copy()
    out["wS"] = f["wX"][:, 0:nz + 1, 1].copy()
    out["wB"] = f["wX"][:, 0:nz + 1, 2].copy()                # NSP1 = 3 :115
    out["wTnt"] = f["wXNT"][:, 0:nz + 1, 0].copy()            # :118
    for name in ("difm", "dift", "difs"):                     # :120-133
        temp_2d = np.zeros((npts, nzp1), order="F")
        temp_2d[:, 0] = 0.0
        temp_2d[:, 1:nzp1] = f[name][:, 1:nz + 1]             # (0:nztmax): Fortran index 1:NZ
        out[name] = temp_2d
    out["rho"] = f["rho"][:, 1:nzp1 + 1].copy()               # rho(:,1:NZP1) of (0:nzp1tmax) :136
    out["cp"] = f["cp"][:, 1:nzp1 + 1].copy()
    out["scorr"] = f["scorr"].copy()
    out["Rig"] = f["Rig"].copy()
    temp_2d = np.zeros((npts, nzp1), order="F")               # :148-150
    temp_2d[:, 0:nz] = f["dbloc"][:, 0:nz]
    temp_2d[:, nzp1 - 1] = 0.0
    out["dbloc"] = temp_2d
    out["Shsq"] = f["Shsq"].copy()
    out["tinc_fcorr"] = f["tinc_fcorr"].copy()
    out["fcorr_z"] = f["ocnTcorr"].copy()
    out["sinc_fcorr"] = f["sinc_fcorr"].copy()
    out["hmix"] = f["hmix"].copy()                            # :168
    out["fcorr"] = f["fcorr"].copy()
    out["taux_in"] = f["sflux"][:, 0, 4, 0].copy()            # sflux(:,1,5,0) :174
    out["tauy_in"] = f["sflux"][:, 1, 4, 0].copy()
    out["solar_in"] = f["sflux"][:, 2, 4, 0].copy()
    out["nsolar_in"] = f["sflux"][:, 3, 4, 0].copy()
    out["PminusE_in"] = f["sflux"][:, 5, 4, 0].copy()         # sflux(:,6,5,0) :186
    out["freeze_flag"] = f["freeze_flag"].copy()
    out["comp_flag"] = f["reset_flag"].copy()                 # :201
    out["dampu_flag"] = f["dampu_flag"].copy()
    out["dampv_flag"] = f["dampv_flag"].copy()
    return out


def xios_restart_output(f, nz):
    """mckpp_xios_restart_output (xios_io.F90:406-431), without the scalar "time"."""
    nzp1 = nz + 1
    out = {}
    out["uvel"] = f["U"][:, :, 0].copy()
    out["vvel"] = f["U"][:, :, 1].copy()
    out["T"] = f["X"][:, :, 0].copy()
    out["S"] = f["X"][:, :, 1].copy()
    out["CP"] = f["cp"][:, 1:nzp1 + 1].copy()
    out["rho"] = f["rho"][:, 1:nzp1 + 1].copy()
    for name in ("hmix", "kmix", "Sref", "SSref", "Ssurf", "Tref"):
        out[name] = f[name].copy()
    out["old"] = f["old"].astype(np.float64)                  # REAL(old) :425
    out["new"] = f["new"].astype(np.float64)
    out["Us"] = f["Us"][:, :, 0, 0:2].copy()                  # Us(:,:,1,0:1) :427
    out["Vs"] = f["Us"][:, :, 1, 0:2].copy()
    out["Ts"] = f["Xs"][:, :, 0, 0:2].copy()
    out["Ss"] = f["Xs"][:, :, 1, 0:2].copy()
    out["hmixd"] = f["hmixd"][:, 0:2].copy()
    return out


def _fint(x):
    """Fortran assignment REAL -> INTEGER: truncation toward zero."""
    return int(math.trunc(x))


def boundary_interp_weights(time, ndtupd, dto, spd, period):
    """MCKPP_BOUNDARY_INTERPOLATE_TEMP/_SAL (boundary_interpolate.F90:27-52 / 82-107):
    (prev_time, next_time, prev_weight, next_weight).  prev_time, next_time, true_time are
    INTEGER in the reference, so every assignment to them truncates."""
    true_time = _fint(time)                                           # :27
    ndays_upd = ndtupd * dto / spd                                    # :28
    prev_time = _fint(math.floor((true_time + ndays_upd / 2) / ndays_upd) * ndays_upd - ndays_upd * 0.5)   # :31
    if prev_time < 0:                                                 # :32-34
        prev_weight = (ndays_upd - abs(true_time - prev_time)) / ndays_upd
        prev_time = prev_time + period
    else:                                                             # :36
        prev_weight = (ndays_upd - (true_time - prev_time)) / ndays_upd
    next_time = _fint(prev_time + ndays_upd)                          # :51
    next_weight = 1 - prev_weight                                     # :52
    return prev_time, next_time, prev_weight, next_weight


def boundary_interpolate(prev_rec, next_rec, prev_weight, next_weight):
    """kpp_3d_fields%ocnT_clim = next_ocnT*next_weight + prev_ocnT*prev_weight (:60, :115)."""
    return next_rec * next_weight + prev_rec * prev_weight
