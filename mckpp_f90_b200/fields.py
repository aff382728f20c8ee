"""Host memory image of the reference's data model.

Mirrors ``TYPE kpp_3D_type`` / ``kpp_const_type`` of the reference
(src/mckpp_data_fields.F90:8-101, 187-346) and their allocators
(``mckpp_allocate_3d_fields`` :353-447, namelist_mod.F90:87-89): every array
is a numpy array in **Fortran order with the Fortran shape**, REAL = float64
(``-fdefault-real-8``), INTEGER / LOGICAL = int32, so ``arr[ipt, k, ...]``
indexes like the Fortran ``arr(ipt+1, k+lb, ...)`` and the bytes are exactly
what a Fortran host would pass through ISO_C_BINDING.

Only the members the hot path (3dto1d / ocnstep / check_profile / 1dto3d /
bottomtemp; src/mckpp_types_transfer.F90) touches are modelled; I/O-only members
(taux..snow, sst, iceconc, ...) stay with the host and are out of scope.
"""
from __future__ import annotations

from dataclasses import dataclass, field, asdict
import numpy as np

F8 = np.float64
I4 = np.int32


@dataclass
class KppDims:
    """mckpp_parameters (src/mckpp_parameters.F90:4-59) subset used by the path."""
    npts: int
    nz: int
    nztmax: int = 0          # must be >= nz+1 (ocnint_mod.F90:33,153; kppmix_mod.F90:82)
    nsflxs: int = 9          # initialize_namelist_mod.F90:37
    njdt: int = 1            # :38
    maxmodeadv: int = 6      # :40

    def __post_init__(self):
        if self.nztmax <= 0:
            self.nztmax = self.nz + 14     # same offset as run/3D_ocn.nml:2,4 (nz=69, nztmax=83)
        if self.nztmax < self.nz + 1:
            raise ValueError("nztmax must be >= nz+1")

    @property
    def nzp1(self):
        return self.nz + 1

    @property
    def nzp1tmax(self):
        return self.nztmax + 1


@dataclass
class KppConsts:
    """kpp_const_type scalars / switches read by the hot path, with the
    reference's namelist defaults (src/mckpp_initialize_namelist_mod.F90:27-47,
    92-119)."""
    dto: float = 1200.0
    grav: float = 9.816
    vonk: float = 0.4
    sice: float = 4.0
    EL: float = 2.50e6
    FLSN: float = 334000.0
    hmixtolfrac: float = 0.1
    itermax: int = 200
    iso_thresh: float = 0.002
    iso_bot: int = 2
    dt_uvdamp: int = 360
    LKPP: bool = True
    LRI: bool = True
    LDD: bool = False
    L_SSref: bool = True
    L_RELAX_SST: bool = False
    L_RELAX_CALCONLY: bool = False
    L_FCORR: bool = False
    L_FCORR_WITHZ: bool = False
    L_SFCORR: bool = False
    L_SFCORR_WITHZ: bool = False
    L_RELAX_SAL: bool = False
    L_RELAX_OCNT: bool = False
    L_NO_FREEZE: bool = False
    L_NO_ISOTHERM: bool = False
    L_DAMP_CURR: bool = False
    L_VARY_BOTTOM_TEMP: bool = False
    have_ocnT_file: bool = False     # ocnT_file .ne. 'none'  (overrides.F90:57)
    have_sal_file: bool = False      # sal_file  .ne. 'none'


# name -> (dtype, shape(dims) in Fortran extents, lower bounds per dim for documentation)
def _shapes(d: KppDims):
    n, nzp1, nz, nzt, nztt = d.npts, d.nzp1, d.nz, d.nztmax, d.nzp1tmax
    return {
        "U": (F8, (n, nzp1, 2)),
        "X": (F8, (n, nzp1, 2)),
        "Rig": (F8, (n, nzp1)),
        "dbloc": (F8, (n, nz)),
        "Shsq": (F8, (n, nzp1)),
        "hmixd": (F8, (n, 2)),                 # (npts,0:1)
        "Us": (F8, (n, nzp1, 2, 2)),           # (npts,nzp1,nvel,0:1)
        "Xs": (F8, (n, nzp1, 2, 2)),
        "rho": (F8, (n, nztt + 1)),            # (npts,0:nzp1tmax)
        "cp": (F8, (n, nztt + 1)),
        "buoy": (F8, (n, nztt)),               # (npts,nzp1tmax)
        "ocdepth": (F8, (n,)),
        "f": (F8, (n,)),
        "swfrac": (F8, (n, nzp1)),
        "swdk_opt": (F8, (n, nz + 1)),         # (npts,0:nz)
        "difm": (F8, (n, nzt + 1)),            # (npts,0:nztmax)
        "difs": (F8, (n, nzt + 1)),
        "dift": (F8, (n, nzt + 1)),
        "wU": (F8, (n, nzt + 1, 3)),           # (npts,0:nztmax,nvp1)
        "wX": (F8, (n, nzt + 1, 3)),           # (npts,0:nztmax,nsp1)
        "wXNT": (F8, (n, nzt + 1, 2)),         # (npts,0:nztmax,nsclr)
        "ghat": (F8, (n, nzt)),                # (npts,nztmax)
        "relax_sst": (F8, (n,)),
        "fcorr": (F8, (n,)),
        "SST0": (F8, (n,)),
        "fcorr_twod": (F8, (n,)),
        "tinc_fcorr": (F8, (n, nzp1)),
        "sinc_fcorr": (F8, (n, nzp1)),
        "fcorr_withz": (F8, (n, nzp1)),
        "sfcorr_withz": (F8, (n, nzp1)),
        "advection": (F8, (n, d.maxmodeadv, 2)),
        "relax_sal": (F8, (n,)),
        "scorr": (F8, (n, nzp1)),
        "relax_ocnT": (F8, (n,)),
        "ocnTcorr": (F8, (n, nzp1)),
        "sal_clim": (F8, (n, nzp1)),
        "ocnT_clim": (F8, (n, nzp1)),
        "hmix": (F8, (n,)),
        "kmix": (F8, (n,)),                    # REAL in the reference
        "Tref": (F8, (n,)),
        "uref": (F8, (n,)),
        "vref": (F8, (n,)),
        "Ssurf": (F8, (n,)),
        "Sref": (F8, (n,)),
        "SSref": (F8, (n,)),
        "sflux": (F8, (n, d.nsflxs, 5, d.njdt + 1)),   # (npts,nsflxs,5,0:njdt)
        "freeze_flag": (F8, (n,)),
        "reset_flag": (F8, (n,)),
        "dampu_flag": (F8, (n,)),
        "dampv_flag": (F8, (n,)),
        "U_init": (F8, (n, nzp1, 2)),
        "bottom_temp": (F8, (n,)),
        "dlat": (F8, (n,)),
        "dlon": (F8, (n,)),
        "l_ocean": (I4, (n,)),
        "l_initflag": (I4, (n,)),
        "run_physics": (I4, (n,)),
        "old": (I4, (n,)),
        "new": (I4, (n,)),
        "jerlov": (I4, (n,)),
        "nmodeadv": (I4, (n, 2)),
        "modeadv": (I4, (n, d.maxmodeadv, 2)),
    }


def field_shapes(d: KppDims):
    return _shapes(d)


def allocate_3d_fields(d: KppDims) -> dict:
    """mckpp_allocate_3d_fields (src/mckpp_data_fields.F90:353-447): zero-filled
    Fortran-ordered arrays.  sflux(:,:,5,0) gets the 1e-20 of
    mckpp_initialize_fluxes (src/mckpp_fluxes_mod.F90:23-27)."""
    out = {}
    for name, (dt, shp) in _shapes(d).items():
        out[name] = np.zeros(shp, dtype=dt, order="F")
    out["sflux"][:, :, 4, 0] = 1e-20
    out["jerlov"][:] = 3            # initialize_optics_mod.F90:43
    out["l_ocean"][:] = 1
    out["run_physics"][:] = 1
    out["ocdepth"][:] = -10000.0    # initialize_landsea_mod.F90:83 (no land-sea file)
    return out


def copy_fields(f: dict) -> dict:
    return {k: np.array(v, order="F", copy=True) for k, v in f.items()}


@dataclass
class KppConstFields:
    """kpp_const_fields: scalars + grid arrays + tables."""
    dims: KppDims
    consts: KppConsts
    zm: np.ndarray = None     # (nzp1)
    hm: np.ndarray = None     # (nzp1)
    dm: np.ndarray = None     # (0:nz)
    tri: np.ndarray = None    # (0:nztmax,0:1,1)
    wmt: np.ndarray = None    # (0:891,0:49)
    wst: np.ndarray = None
