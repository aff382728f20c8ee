"""ctypes binding of the CPU oracle (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs may import this module.  The product package
(mckpp_f90_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_CONST_INT = ["nz", "nzp1", "nztmax", "nzp1tmax", "npts", "nsflxs", "njdt", "maxmodeadv", "itermax",
              "iso_bot", "dt_uvdamp", "LKPP", "LRI", "LDD", "L_SSref", "L_RELAX_SST", "L_RELAX_CALCONLY",
              "L_FCORR", "L_FCORR_WITHZ", "L_SFCORR", "L_SFCORR_WITHZ", "L_RELAX_SAL", "L_RELAX_OCNT",
              "L_NO_FREEZE", "L_NO_ISOTHERM", "L_DAMP_CURR", "L_VARY_BOTTOM_TEMP", "have_ocnT_file",
              "have_sal_file", "pad0"]
_CONST_DBL = ["hmixtolfrac", "dto", "grav", "vonk", "sice", "iso_thresh"]
_CONST_PTR = ["zm", "hm", "dm", "tri", "wmt", "wst"]


class OrcConst(C.Structure):
    _fields_ = ([(n, C.c_int32) for n in _CONST_INT] + [(n, C.c_double) for n in _CONST_DBL]
                + [(n, C.c_void_p) for n in _CONST_PTR])


_3D_DBL = ["U", "X", "Rig", "dbloc", "Shsq", "hmixd", "Us", "Xs", "rho", "cp", "buoy", "ocdepth", "f", "swfrac",
           "swdk_opt", "difm", "difs", "dift", "wU", "wX", "wXNT", "ghat", "relax_sst", "fcorr", "SST0",
           "fcorr_twod", "tinc_fcorr", "sinc_fcorr", "fcorr_withz", "sfcorr_withz", "advection", "relax_sal",
           "scorr", "relax_ocnT", "ocnTcorr", "sal_clim", "ocnT_clim", "hmix", "kmix", "Tref", "uref", "vref",
           "Ssurf", "Sref", "SSref", "sflux", "freeze_flag", "reset_flag", "dampu_flag", "dampv_flag", "U_init",
           "bottom_temp"]
_3D_INT = ["l_ocean", "l_initflag", "run_physics", "old", "new_", "jerlov", "nmodeadv", "modeadv"]
_3D_DIAG = ["diag_iter", "diag_nreint", "diag_status", "diag_talpha", "diag_sbeta"]


class Orc3d(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in _3D_DBL + _3D_INT + _3D_DIAG]


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libmckpp_oracle.so")
    src = os.path.join(_HERE, "mckpp_oracle.c")
    hdr = os.path.join(_HERE, "mckpp_oracle.h")
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["make", "-s", "-C", _HERE, "libmckpp_oracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = build()
        L = C.CDLL(so)
        L.orc_3d_member_names.restype = C.c_char_p
        L.orc_const_member_names.restype = C.c_char_p
        assert L.orc_3d_member_names().decode().split(",") == _3D_DBL + _3D_INT + _3D_DIAG
        assert L.orc_const_member_names().decode().split(",") == _CONST_INT + _CONST_DBL + _CONST_PTR
        L.orc_physics_driver.restype = C.c_int
        L.orc_physics_driver.argtypes = [C.POINTER(OrcConst), C.POINTER(Orc3d), C.c_int, C.c_int, C.c_int]
        L.orc_initialize_ocean_model.restype = C.c_int
        L.orc_initialize_ocean_model.argtypes = [C.POINTER(OrcConst), C.POINTER(Orc3d), C.c_int]
        L.orc_cpsw.restype = C.c_double
        L.orc_cpsw.argtypes = [C.c_double] * 3
        L.orc_swdk.restype = C.c_double
        L.orc_swdk.argtypes = [C.c_double, C.c_int]
        _LIB = L
    return _LIB


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Oracle:
    """Drives the oracle on a host memory image (dict of Fortran-ordered numpy
    arrays, see mckpp_f90_b200.fields) -- the same image the GPU library is fed."""

    def __init__(self, const_fields, fields: dict, nthreads: int = 0):
        self.L = lib()
        self.cf = const_fields
        self.f = fields
        self.nthreads = nthreads
        d, k = const_fields.dims, const_fields.consts
        c = OrcConst()
        c.nz, c.nzp1, c.nztmax, c.nzp1tmax, c.npts = d.nz, d.nzp1, d.nztmax, d.nzp1tmax, d.npts
        c.nsflxs, c.njdt, c.maxmodeadv = d.nsflxs, d.njdt, d.maxmodeadv
        c.itermax, c.iso_bot, c.dt_uvdamp = k.itermax, k.iso_bot, k.dt_uvdamp
        for n in ["LKPP", "LRI", "LDD", "L_SSref", "L_RELAX_SST", "L_RELAX_CALCONLY", "L_FCORR", "L_FCORR_WITHZ",
                  "L_SFCORR", "L_SFCORR_WITHZ", "L_RELAX_SAL", "L_RELAX_OCNT", "L_NO_FREEZE", "L_NO_ISOTHERM",
                  "L_DAMP_CURR", "L_VARY_BOTTOM_TEMP", "have_ocnT_file", "have_sal_file"]:
            setattr(c, n, int(bool(getattr(k, n))))
        c.hmixtolfrac, c.dto, c.grav, c.vonk, c.sice, c.iso_thresh = (k.hmixtolfrac, k.dto, k.grav, k.vonk,
                                                                       k.sice, k.iso_thresh)
        self._keep = []
        for n in _CONST_PTR:
            a = np.asfortranarray(getattr(const_fields, n), dtype=np.float64)
            self._keep.append(a)
            setattr(c, n, _ptr(a))
        self.c = c
        n = d.npts
        self.diag = {
            "iter": np.zeros(n, np.int32), "nreint": np.zeros(n, np.int32), "status": np.zeros(n, np.int32),
            "talpha": np.zeros((n, d.nzp1 + 1), order="F"), "sbeta": np.zeros((n, d.nzp1 + 1), order="F"),
        }
        s = Orc3d()
        for name in _3D_DBL:
            a = fields[name]
            assert a.dtype == np.float64 and a.flags.f_contiguous, name
            setattr(s, name, _ptr(a))
        for name in _3D_INT:
            a = fields["new" if name == "new_" else name]
            assert a.dtype == np.int32 and a.flags.f_contiguous, name
            setattr(s, name, _ptr(a))
        s.diag_iter = _ptr(self.diag["iter"])
        s.diag_nreint = _ptr(self.diag["nreint"])
        s.diag_status = _ptr(self.diag["status"])
        s.diag_talpha = _ptr(self.diag["talpha"])
        s.diag_sbeta = _ptr(self.diag["sbeta"])
        self.s = s

    def initialize_ocean_model(self):
        return self.L.orc_initialize_ocean_model(C.byref(self.c), C.byref(self.s), self.nthreads)

    def physics_driver(self, ntime: int, realloc_1d: bool = False):
        return self.L.orc_physics_driver(C.byref(self.c), C.byref(self.s), int(ntime), self.nthreads,
                                         int(realloc_1d))


def abk80(S, T, P, want_kappa=False):
    L = lib()
    a, b, k = C.c_double(1.0), C.c_double(1.0), C.c_double(1.0 if want_kappa else 0.0)
    s0, s = C.c_double(0.0), C.c_double(0.0)
    L.orc_abk80(C.c_double(S), C.c_double(T), C.c_double(P), C.byref(a), C.byref(b), C.byref(k), C.byref(s0),
                C.byref(s))
    return a.value, b.value, k.value, s0.value, s.value


def cpsw(S, T, P):
    return lib().orc_cpsw(S, T, P)
