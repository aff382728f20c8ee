"""Synthetic inputs of the named shapes (BASELINE.json configs 1-5).

There is no network and no netCDF: initial profiles, masks and forcing are
generated here, deterministically (SplitMix64, seed 0x4B5050 + column), as
plain numpy arrays in the reference's memory image (see fields.py).  The
oracle and the GPU library are both fed these same arrays, so inputs are
bit-identical on both sides by construction.

Forcing level "A" is the reference's built-in constant forcing
(src/mckpp_fluxes_mod.F90:41-49).  Level "B" is a diurnal / spatially varying
stress test, mapped to ``sflux`` exactly as fluxes_mod.F90:59-70 does.
"""
from __future__ import annotations

from dataclasses import dataclass, replace
import math
import numpy as np

from .fields import KppDims, KppConsts, KppConstFields, allocate_3d_fields
from . import hostinit

_MASK = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64_draws(seeds: np.ndarray, ndraw: int) -> np.ndarray:
    """ndraw uniform doubles in [0,1) per seed (vectorised SplitMix64)."""
    x = seeds.astype(np.uint64).copy()
    out = np.empty((seeds.shape[0], ndraw), dtype=np.float64)
    with np.errstate(over="ignore"):
        for d in range(ndraw):
            x = (x + np.uint64(0x9E3779B97F4A7C15)) & _MASK
            z = x.copy()
            z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _MASK
            z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _MASK
            z = z ^ (z >> np.uint64(31))
            out[:, d] = (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
    return out


@dataclass
class SynthConfig:
    name: str
    nx: int
    ny: int
    nz: int = 100
    dto: float = 1200.0
    dmax: float = 1000.0
    stretch: bool = False
    dscale: float = 0.0
    forcing: str = "B"          # "A" reference built-in constant, "B" diurnal stress
    LDD: bool = False
    corrections: bool = False   # config 5: fcorr_withz, sfcorr_withz, relax_ocnT, relax_sal, L_NO_FREEZE
    shallow_frac: float = 0.0   # fraction of columns with ocdepth in [-900,-50]
    ice_frac: float = 0.0       # fraction of "ice" columns (SST -1.9, sflux(5) = -1e-5)
    land_frac: float = 0.0      # fraction of land points (run_physics = .FALSE.)
    # fraction of columns with salt ABOVE fresher water: with the prescribed profile
    # S = 35 - exp(z/300) alphaDT and betaDS never have the same sign and ddmix (ddmix_mod.F90:30-50) would
    # add nothing anywhere -- config 4 switches double diffusion ON, so a quarter of its columns carry a
    # profile where the salt-fingering branch really runs (density ratio between 1 and 1.9 at depth)
    dd_frac: float = 0.0
    ndays: float = 1.0

    @property
    def npts(self):
        return self.nx * self.ny


# BASELINE.json configs (sizes per SURVEY §8d)
CONFIGS = {
    "cfg1": SynthConfig("cfg1: 16-column 4x4, NZ=100, 3-hour step, 1 day", 4, 4, dto=10800.0, forcing="A", ndays=1.0),
    "cfg2": SynthConfig("cfg2: regional 300x200, NZ=100, synthetic fluxes, 30 days", 300, 200, ndays=30.0),
    "cfg3": SynthConfig("cfg3: global 1deg ~44k ocean columns, NZ=100", 220, 200),
    "cfg4": SynthConfig("cfg4: global 0.25deg ~700k ocean columns, LDD", 1000, 700, LDD=True, dd_frac=0.25),
    "cfg5": SynthConfig("cfg5: NZ=250 stretched, corrections + freeze clamp", 300, 200, nz=250, stretch=True,
                        dscale=4.0, corrections=True, shallow_frac=0.10, ice_frac=0.05),
}


def scaled(cfg: SynthConfig, nx: int, ny: int) -> SynthConfig:
    """Same physics/flags on a smaller (or larger) column count."""
    return replace(cfg, nx=nx, ny=ny)


def make_consts(cfg: SynthConfig) -> KppConsts:
    c = KppConsts(dto=cfg.dto, LDD=cfg.LDD)
    if cfg.corrections:
        c.L_FCORR_WITHZ = True
        c.L_SFCORR_WITHZ = True
        c.L_RELAX_OCNT = True
        c.L_RELAX_SAL = True
        c.L_NO_FREEZE = True
        c.have_ocnT_file = True
        c.have_sal_file = True
    return c


def make_case(cfg: SynthConfig, col_offset: int = 0, ncols: int | None = None, gidx=None):
    """Returns (const_fields: KppConstFields, fields: dict, draws: ndarray).

    ``col_offset``/``ncols`` select a contiguous block of the global compact
    column list (multi-GPU partitioning): block k of a run is bit-identical to
    the same columns of the full run because every column depends only on its
    global index.  ``gidx`` selects an arbitrary list of global column indices
    instead (bounded CPU-baseline samples of the same workload)."""
    ntot = cfg.npts
    if gidx is not None:
        gidx = np.asarray(gidx, dtype=np.int64)
        n = int(gidx.size)
    else:
        n = ntot - col_offset if ncols is None else ncols
    dims = KppDims(npts=n, nz=cfg.nz)
    consts = make_consts(cfg)
    cf = hostinit.build_const_fields(dims, consts, cfg.dmax, cfg.stretch, cfg.dscale)
    f = allocate_3d_fields(dims)

    if gidx is None:
        gidx = np.arange(col_offset, col_offset + n, dtype=np.int64)
    r = splitmix64_draws((np.uint64(0x4B5050) + gidx.astype(np.uint64)), 10)
    row = gidx // cfg.nx
    colx = gidx % cfg.nx
    lat = -70.0 + 140.0 * (row + 0.5) / cfg.ny
    lon = 360.0 * (colx + 0.5) / cfg.nx
    f["dlat"][:] = lat
    f["dlon"][:] = lon
    f["f"][:] = hostinit.coriolis(lat)

    zm = cf.zm
    sst = 28.0 - 26.0 * np.sin(np.deg2rad(lat)) ** 2 + (r[:, 0] - 0.5)
    ice = r[:, 8] < cfg.ice_frac
    sst = np.where(ice, -1.9, sst)
    hscale = 120.0 + 80.0 * r[:, 1]
    T = 2.0 + (sst[:, None] - 2.0) * np.exp(zm[None, :] / hscale[:, None])
    S = 35.0 - 1.0 * np.exp(zm[None, :] / 300.0) + 0.2 * (r[:, 2, None] - 0.5)
    if cfg.dd_frac > 0:
        # salt above fresher water with the temperature profile's own decay scale and an amplitude tied to
        # the column's thermal contrast: the density ratio alpha*dT/(beta*dS) stays above 1 everywhere
        # (statically stable) and falls into ddmix's salt-fingering range (1, 1.9) in the cold water at depth
        fing = r[:, 9] > 1.0 - cfg.dd_frac
        Sf = 35.0 + 0.08 * (sst[:, None] - 2.0) * np.exp(zm[None, :] / hscale[:, None]) + 0.2 * (r[:, 2, None] - 0.5)
        S = np.where(fing[:, None], Sf, S)
    # reference salinity removed (initialize_ocean_profiles_mod.F90:104-109)
    sref = (S[:, 0] + S[:, -1]) / 2.0
    f["Sref"][:] = sref
    f["SSref"][:] = sref
    f["X"][:, :, 0] = T
    f["X"][:, :, 1] = S - sref[:, None]
    f["U"][:] = 0.0
    f["U_init"][:] = f["U"]
    f["Tref"][:] = f["X"][:, 0, 0]
    f["Ssurf"][:] = f["SSref"] if consts.L_SSref else f["X"][:, 0, 1] + f["Sref"]
    f["jerlov"][:] = 3

    if cfg.shallow_frac > 0:
        sh = r[:, 9] < cfg.shallow_frac
        f["ocdepth"][:] = np.where(sh, -(50.0 + 850.0 * r[:, 3]), -10000.0)
    if cfg.land_frac > 0:
        land = r[:, 9] > 1.0 - cfg.land_frac
        f["l_ocean"][:] = np.where(land, 0, 1)
        f["run_physics"][:] = f["l_ocean"]

    if cfg.corrections:
        f["fcorr_withz"][:] = 5.0 * np.exp(zm[None, :] / 100.0)
        f["sfcorr_withz"][:] = 0.0
        f["ocnT_clim"][:] = f["X"][:, :, 0]
        f["sal_clim"][:] = f["X"][:, :, 1]
        f["relax_ocnT"][:] = 1.0 / (30.0 * 86400.0)
        f["relax_sal"][:] = 1.0 / (30.0 * 86400.0)
    return cf, f, r


def apply_forcing(cfg: SynthConfig, cf: KppConstFields, f: dict, r: np.ndarray, nt: int):
    """Host forcing for timestep nt (1-based), written into sflux(:,1:6,5,0) the way
    mckpp_fluxes does (fluxes_mod.F90:59-70).  Returns the (6,npts) block."""
    n = f["sflux"].shape[0]
    c = cf.consts
    if cfg.forcing == "A":
        taux = np.full(n, 0.01); tauy = np.zeros(n); swf = np.full(n, 200.0); lwf = np.zeros(n)
        lhf = np.full(n, -150.0); shf = np.zeros(n); rain = np.full(n, 6e-5); snow = np.zeros(n)
    else:
        t_day = ((nt - 1) * c.dto / 86400.0) % 1.0
        sun = max(0.0, 900.0 * math.sin(2.0 * math.pi * (t_day - 0.25)))
        swf = sun * (0.6 + 0.4 * r[:, 3])
        taux = 0.15 * (2.0 * r[:, 4] - 1.0)
        tauy = 0.15 * (2.0 * r[:, 5] - 1.0)
        lhf = -50.0 - 250.0 * r[:, 6]
        lwf = np.full(n, -60.0)
        shf = np.full(n, -10.0)
        rain = 2e-4 * r[:, 7] ** 2
        snow = np.zeros(n)
    hostinit.fluxes_map(f, c, taux, tauy, swf, lwf, lhf, shf, rain, snow)
    if cfg.ice_frac > 0:
        ice = r[:, 8] < cfg.ice_frac
        # ice-melt freshwater term used in verticalmixing_mod.F90:91-93; injected at
        # sflux level because the reference's map hard-codes 1e-10 (fluxes_mod.F90:68)
        f["sflux"][ice, 4, 4, 0] = -1.0e-5
    return np.ascontiguousarray(f["sflux"][:, 0:6, 4, 0].T)


def block_partition(npts: int, world: int, rank: int):
    """(first column, columns) of rank's block when npts columns are split over `world` GPUs: contiguous blocks of
    ceil(npts/world) columns rounded up to whole 32-column tiles -- the rule of kpp_gpu_create_multi, which
    bench.py applies to its ranks.  A trailing rank may get fewer columns, or none."""
    block = ((npts + world - 1) // world + 31) // 32 * 32
    col0 = min(rank * block, npts)
    return col0, max(0, min(block, npts - col0))


def nsteps(cfg: SynthConfig) -> int:
    return int(round(cfg.ndays * 86400.0 / cfg.dto))
