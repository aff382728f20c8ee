"""Host-side builders of the hot path's *inputs* (SURVEY §8 row a20).

In the reference these are Fortran init routines that stay on the host; the
GPU library only consumes their results through ``kpp_gpu_create``.  They are
restated here (numpy / Python floats = IEEE double) so that tests and bench.py
can produce the same inputs without the reference:

* vertical grid ``zm, hm, dm``   src/mckpp_initialize_geography_mod.F90:43-74
* Coriolis ``f``                 src/mckpp_initialize_geography_mod.F90:78-88
* ``tri``                        src/mckpp_initialize_ocean.F90:34-43
* ``wmt, wst`` lookup tables     src/mckpp_physics_lookup_mod.F90:42-64
* forcing map -> ``sflux``       src/mckpp_fluxes_mod.F90:56-72
"""
from __future__ import annotations

import math
import numpy as np

from .fields import KppDims, KppConsts, KppConstFields


def build_grid(nz: int, dmax: float, l_stretchgrid: bool = False, dscale: float = 0.0):
    """initialize_geography_mod.F90:43-74 (no vgrid file).  Scalar loop on purpose:
    ``hsum`` is a serial accumulation and its rounding defines the grid."""
    hm = np.zeros(nz + 1)
    zm = np.zeros(nz + 1)
    dm = np.zeros(nz + 1)
    sumh = 0.0
    if l_stretchgrid:
        dfac = 1.0 - math.exp(-dscale)
        for i in range(1, nz + 1):
            sk = -(float(i) - 0.5) / float(nz)
            hm[i - 1] = dmax * dfac / float(nz) / dscale / (1.0 + sk * dfac)
            sumh = sumh + hm[i - 1]
    hsum = 0.0
    for i in range(1, nz + 1):
        if l_stretchgrid:
            hm[i - 1] = hm[i - 1] * dmax / sumh
        else:
            hm[i - 1] = dmax / float(nz)
        zm[i - 1] = 0.0 - (hsum + 0.5 * hm[i - 1])
        hsum = hsum + hm[i - 1]
        dm[i] = hsum
    dm[0] = 0.0
    hm[nz] = 1.0e-10
    zm[nz] = -dmax
    return zm, hm, dm


def coriolis(dlat: np.ndarray) -> np.ndarray:
    """initialize_geography_mod.F90:78-88; twopi = 8*atan(1.) (namelist_mod.F90:94)."""
    twopi = 8.0 * math.atan(1.0)
    dlat = np.asarray(dlat, dtype=np.float64)
    f = np.empty_like(dlat)
    for i, la in enumerate(dlat):
        if abs(la) < 2.5:
            f[i] = 2.0 * (twopi / 86164.0) * math.sin(2.5 * twopi / 360.0) * math.copysign(1.0, la)
        else:
            f[i] = 2.0 * (twopi / 86164.0) * math.sin(la * twopi / 360.0)
    return f


def build_tri(dims: KppDims, dto: float, zm: np.ndarray, hm: np.ndarray) -> np.ndarray:
    """initialize_ocean.F90:34-43.  tri(0:nztmax,0:1,1) in Fortran order."""
    nz = dims.nz
    tri = np.zeros((dims.nztmax + 1, 2, 1), order="F")
    dzb = np.zeros(nz + 1)
    for k in range(1, nz + 1):
        dzb[k] = zm[k - 1] - zm[k]
    tri[0, 1, 0] = dto / hm[0]
    tri[1, 1, 0] = dto / hm[0] / dzb[1]
    for k in range(2, nz + 1):
        tri[k, 1, 0] = dto / hm[k - 1] / dzb[k]
        tri[k, 0, 0] = dto / hm[k - 1] / dzb[k - 1]
    return tri


def build_lookup(vonk: float):
    """physics_lookup_mod.F90:42-64.  wmt/wst(0:891,0:49), Fortran order.
    Uses libm ``pow`` through Python floats, like the Fortran ``**`` with a
    real exponent."""
    ni, nj = 890, 48
    epsln = 1.0e-20
    c1 = 5.0
    zmin, zmax, umin, umax = -4.0e-7, 0.0, 0.0, 0.04
    am, cm, c2, zetam = 1.257, 8.380, 16.0, -0.2
    as_, cs, c3, zetas = -28.86, 98.96, 16.0, -1.0
    deltaz = (zmax - zmin) / (ni + 1)
    deltau = (umax - umin) / (nj + 1)
    wmt = np.zeros((892, 50), order="F")
    wst = np.zeros((892, 50), order="F")
    for i in range(0, ni + 2):
        zehat = deltaz * i + zmin
        for j in range(0, nj + 2):
            usta = deltau * j + umin
            zeta = zehat / (usta * usta * usta + epsln)
            if zehat >= 0.0:
                wmt[i, j] = vonk * usta / (1.0 + c1 * zeta)
                wst[i, j] = wmt[i, j]
            else:
                if zeta > zetam:
                    wmt[i, j] = vonk * usta * math.pow(1.0 - c2 * zeta, 1.0 / 4.0)
                else:
                    wmt[i, j] = vonk * math.pow(am * (usta * usta * usta) - cm * zehat, 1.0 / 3.0)
                if zeta > zetas:
                    wst[i, j] = vonk * usta * math.pow(1.0 - c3 * zeta, 1.0 / 2.0)
                else:
                    wst[i, j] = vonk * math.pow(as_ * (usta * usta * usta) - cs * zehat, 1.0 / 3.0)
    return wmt, wst


_LOOKUP_CACHE = {}


def build_const_fields(dims: KppDims, consts: KppConsts, dmax: float = 1000.0,
                       l_stretchgrid: bool = False, dscale: float = 0.0) -> KppConstFields:
    zm, hm, dm = build_grid(dims.nz, dmax, l_stretchgrid, dscale)
    tri = build_tri(dims, consts.dto, zm, hm)
    key = float(consts.vonk)
    if key not in _LOOKUP_CACHE:
        _LOOKUP_CACHE[key] = build_lookup(consts.vonk)
    wmt, wst = _LOOKUP_CACHE[key]
    return KppConstFields(dims=dims, consts=consts, zm=zm, hm=hm, dm=dm, tri=tri, wmt=wmt, wst=wst)


def fluxes_map(fields: dict, consts: KppConsts, taux, tauy, swf, lwf, lhf, shf, rain, snow):
    """mckpp_fluxes forcing map (fluxes_mod.F90:56-72, l_rest=.FALSE.) into
    ``sflux(:,1:6,5,0)``.  Vectorised; every expression is elementwise so the
    rounding equals the scalar loop's."""
    sflux = fields["sflux"]
    oc = fields["l_ocean"] != 0
    taux = np.array(taux, dtype=np.float64, copy=True)
    tauy = np.asarray(tauy, dtype=np.float64)
    zero = (taux == 0.0) & (tauy == 0.0) & oc
    taux[zero] = 1.0e-10
    sflux[oc, 0, 4, 0] = taux[oc]
    sflux[oc, 1, 4, 0] = tauy[oc]
    sflux[oc, 2, 4, 0] = np.asarray(swf)[oc]
    sflux[oc, 3, 4, 0] = ((np.asarray(lwf) + np.asarray(lhf)) + np.asarray(shf) - np.asarray(snow) * consts.FLSN)[oc]
    sflux[oc, 4, 4, 0] = 1e-10
    sflux[oc, 5, 4, 0] = ((np.asarray(rain) + np.asarray(snow)) + (np.asarray(lhf) / consts.EL))[oc]


def boundary_interp_weights(time, ndtupd, dto, spd, period):
    """Time bracket and weights of MCKPP_BOUNDARY_INTERPOLATE_TEMP / _SAL
    (src/mckpp_boundary_interpolate.F90:27-52, 82-107): (prev_time, next_time, prev_weight,
    next_weight).  `prev_time`, `next_time` and `true_time` are INTEGER in the reference, so each
    assignment truncates toward zero; the weights are REAL(8)."""
    import math
    true_time = int(math.trunc(time))
    ndays_upd = ndtupd * dto / spd
    prev_time = int(math.trunc(math.floor((true_time + ndays_upd / 2) / ndays_upd) * ndays_upd - ndays_upd * 0.5))
    if prev_time < 0:
        prev_weight = (ndays_upd - abs(true_time - prev_time)) / ndays_upd
        prev_time = prev_time + period
    else:
        prev_weight = (ndays_upd - (true_time - prev_time)) / ndays_upd
    next_time = int(math.trunc(prev_time + ndays_upd))
    next_weight = 1 - prev_weight
    return prev_time, next_time, prev_weight, next_weight
