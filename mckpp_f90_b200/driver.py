"""Host-side mirror of the reference's driver interface for the column step.

Same names and argument meaning as the Fortran the GPU path replaces:

* ``kpp_3d_fields`` / ``kpp_const_fields``   module globals of mckpp_data_fields
  (src/mckpp_data_fields.F90:348-349) -> attributes of :class:`MckppPhysics`
* ``mckpp_initialize_ocean_model()``         src/mckpp_initialize_ocean.F90:18
* ``mckpp_physics_driver()``                 src/mckpp_physics_driver_mod.F90:15

The host keeps owning ``kpp_3d_fields`` (numpy arrays with the Fortran memory
image, see fields.py); the device mirror lives in the library handle.  A
"dirty" protocol moves only what changed: the host pushes its inputs once and
the per-step forcing ``sflux(:,1:6,5,0)`` every step; it pulls the fields in
``pull`` after each step (default: the per-column scalars 1dto3d writes) and
anything else on demand with :meth:`pull`.  ``sync_mode="full"`` reproduces
the reference's behaviour of refreshing every member after every step.

Warnings and fatal errors follow the reference: the 'long iteration' /
'Failed to find a reasonable solution' / 'Resetting point' warnings
(ocnstep_mod.F90:184-191,229-236; overrides.F90:59-75) are reported from the
per-column status word; a tridiagonal zero pivot raises (the reference calls
MCKPP_ABORT, solvers.F90:140-149).  There is no CPU fallback.
"""
from __future__ import annotations

import sys
import numpy as np

from . import capi
from .fields import KppConstFields

# members the physics reads (pushed by push_inputs)
INPUT_FIELDS = ["U", "X", "Us", "Xs", "hmixd", "old", "new", "hmix", "kmix", "Tref", "uref", "vref", "Ssurf", "Sref",
                "SSref", "f", "ocdepth", "jerlov", "l_ocean", "run_physics", "sflux", "U_init", "relax_sst", "SST0",
                "fcorr_twod", "fcorr", "relax_sal", "relax_ocnT", "sal_clim", "ocnT_clim", "fcorr_withz",
                "sfcorr_withz", "bottom_temp", "nmodeadv", "modeadv", "advection", "freeze_flag", "reset_flag",
                "dampu_flag", "dampv_flag", "tinc_fcorr", "wXNT"]
# per-column scalars that mckpp_fields_1dto3d writes back (types_transfer.F90:296-326)
SCALAR_OUTPUTS = ["hmix", "kmix", "Tref", "uref", "vref", "Ssurf", "old", "new", "reset_flag", "dampu_flag",
                  "dampv_flag", "freeze_flag", "fcorr"]
# prognostic state
STATE_OUTPUTS = ["U", "X", "Us", "Xs", "hmixd"]
# profile diagnostics that 1dto3d writes back (types_transfer.F90:207-294)
DIAG_OUTPUTS = ["rho", "cp", "buoy", "Rig", "dbloc", "Shsq", "difm", "difs", "dift", "ghat", "wU", "wX", "wXNT",
                "tinc_fcorr", "sinc_fcorr", "ocnTcorr", "scorr", "swfrac", "swdk_opt"]
ALL_OUTPUTS = SCALAR_OUTPUTS + STATE_OUTPUTS + DIAG_OUTPUTS


class MckppPhysics:
    def __init__(self, kpp_const_fields: KppConstFields, kpp_3d_fields: dict, device: int = 0, numerics: int = 0,
                 sync_mode: str = "lazy", pull=None, verbose: bool = False, ngpus: int = None, devices=None):
        """ngpus / devices: partition the columns over several GPUs of the node inside the library
        (one host process, the reference's single-rank layout); everything else is unchanged."""
        self.kpp_const_fields = kpp_const_fields
        self.kpp_3d_fields = kpp_3d_fields
        self.gpu = capi.KppGpu(kpp_const_fields, device=device, numerics=numerics, ngpus=ngpus, devices=devices)
        self.sync_mode = sync_mode
        self.pull_after_step = list(ALL_OUTPUTS if sync_mode == "full" else (SCALAR_OUTPUTS if pull is None else pull))
        self.verbose = verbose
        self.last_report = None
        n = kpp_const_fields.dims.npts
        self.diag = {"iter": np.zeros(n, np.int32), "nreint": np.zeros(n, np.int32), "status": np.zeros(n, np.int32)}
        self._sflux6 = np.zeros((6, n))

    # ---- data movement -------------------------------------------------
    def push_inputs(self, names=None):
        for name in (INPUT_FIELDS if names is None else names):
            self.gpu.upload(name, self.kpp_3d_fields[name])

    def pull(self, names):
        names = list(names)
        for name in names:
            self.gpu.download_async(name, self.kpp_3d_fields[name])
        if names:
            self.gpu.L.kpp_gpu_sync(self.gpu.h, None)

    def pull_diag(self):
        self.gpu.download("diag_iter", self.diag["iter"])
        self.gpu.download("diag_nreint", self.diag["nreint"])
        self.gpu.download("diag_status", self.diag["status"])
        return self.diag

    # ---- the reference's entry points -----------------------------------
    def mckpp_initialize_ocean_model(self):
        """Per-column loop of MCKPP_INITIALIZE_OCEAN_MODEL (initialize_ocean.F90:54-104)."""
        self.gpu.init_vmix()
        self.gpu.sync()
        self.pull(["hmix", "kmix", "Tref", "uref", "vref", "old", "new", "hmixd", "Us", "Xs"] +
                  (DIAG_OUTPUTS if self.sync_mode == "full" else []))

    def mckpp_physics_driver(self, ntime: int, forcing_changed: bool = True):
        """One call of mckpp_physics_driver() (physics_driver_mod.F90:15-73) for timestep ntime."""
        f = self.kpp_3d_fields
        if forcing_changed:
            # sflux(:,1:6,5,0) is what mckpp_fluxes just filled (fluxes_mod.F90:63-70)
            np.copyto(self._sflux6, f["sflux"][:, 0:6, 4, 0].T)
            self.gpu.upload_forcing(self._sflux6)
        self.gpu.step(ntime)
        try:
            rep = self.gpu.sync()
        except capi.KppError as e:
            if e.code == capi.KPP_E_PIVOT_ZERO:
                # MCKPP_PHSYICS_SOLVER_TRIDMAT: the reference prints and calls MCKPP_ABORT
                sys.stderr.write("MCKPP_PHSYICS_SOLVER_TRIDMAT: Algorithm for solving tridiag matrix failed.\n")
            raise
        self.last_report = rep
        self._warn(rep, ntime)
        self.pull(self.pull_after_step)
        return rep

    def mckpp_fluxes(self, taux, tauy, swf, lwf, lhf, shf, rain, snow):
        """mckpp_fluxes() with the forcing map on the device (fluxes_mod.F90:56-72):
        the host hands over the eight raw flux fields, sflux(:,1:6,5,0) is filled in HBM and the
        next mckpp_physics_driver(..., forcing_changed=False) uses it."""
        k = self.kpp_const_fields.consts
        self.gpu.upload_fluxes(taux, tauy, swf, lwf, lhf, shf, rain, snow, k.FLSN, k.EL)

    # ---- what the I/O layer sends, packed on the device
    def _packed_set(self, restart):
        out, ids = {}, self.gpu.output_ids(restart=restart)
        for oid, name in ids.items():
            out[name] = self.gpu.pack_output(oid, sync=False)
        self.gpu.L.kpp_gpu_sync(self.gpu.h, None)
        return out

    def mckpp_xios_diagnostic_output(self) -> dict:
        """mckpp_xios_diagnostic_output (xios_io.F90:72-207): {xios field id: array as sent}, with the
        temp_2d reshuffles (S = X2 + Sref, shifted diffusivities, padded dbloc) done on the device.
        `cplwght` is host data and stays with the host."""
        return self._packed_set(restart=False)

    def mckpp_xios_restart_output(self) -> dict:
        """mckpp_xios_restart_output (xios_io.F90:406-431) without the scalar "time"."""
        return self._packed_set(restart=True)

    # ---- climatology interpolation on the device
    def mckpp_boundary_interpolate(self, name, prev_weight, next_weight, prev_rec=None, next_rec=None):
        """MCKPP_BOUNDARY_INTERPOLATE_TEMP (name="ocnT_clim") / _SAL ("sal_clim")
        (boundary_interpolate.F90:14-123): clim = next*next_weight + prev*prev_weight in HBM.
        Pass the two records only when the bracket moved (hostinit.boundary_interp_weights gives
        prev_time/next_time and the weights); otherwise the resident records are re-blended."""
        if prev_rec is not None:
            self.gpu.upload_clim_record(name, 0, prev_rec)
        if next_rec is not None:
            self.gpu.upload_clim_record(name, 1, next_rec)
        self.gpu.blend_clim(name, prev_weight, next_weight)

    def _warn(self, rep, ntime):
        if not self.verbose:
            return
        if rep.n_long_iter:
            sys.stderr.write(f"MCKPP_PHYSICS_OCNSTEP: long iteration at timestep {ntime} on {rep.n_long_iter} points\n")
        if rep.n_reint_fail:
            sys.stderr.write("MCKPP_PHYSICS_OCNSTEP: Failed to find a reasonable solution in the semi-implicit "
                             f"integration after 10 iterations on {rep.n_reint_fail} points\n")
        if rep.n_reset:
            sys.stderr.write(f"MCKPP_PHSYICS_OVERRIDE_CHECK_PROFILE: Resetting {rep.n_reset} points\n")

    def close(self):
        self.gpu.close()
