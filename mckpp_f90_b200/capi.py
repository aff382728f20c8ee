"""ctypes binding of the C ABI in include/kpp_gpu.h (libkpp_gpu.so).

This is the same binding a Fortran ISO_C_BINDING host performs (see
INTEGRATION.md): plain pointers, sizes, int32 and double.  There is no CPU
fallback -- if the CUDA library is missing or no device is present the calls
raise.
"""
from __future__ import annotations

import ctypes as C
import os
import re

import numpy as np

from .fields import KppDims, KppConsts, KppConstFields

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("KPP_LIB_PATH") or os.path.join(_HERE, "libkpp_gpu.so")   # override: kernel A/B experiments
HEADER_PATH = os.path.join(_HERE, "..", "include", "kpp_gpu.h")

KPP_OK, KPP_E_INVALID, KPP_E_CUDA, KPP_E_NODEVICE, KPP_E_PIVOT_ZERO, KPP_E_NOMEM = 0, -1, -2, -3, -4, -5

ST_LONG_ITER, ST_REINT_FAIL, ST_RESET, ST_PIVOT_ZERO, ST_ITER_CAP, ST_ISO_RESET, ST_BAD_OLDNEW = 1, 2, 4, 8, 16, 32, 64


class KppError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"kpp_gpu error {code}: {msg}")
        self.code = code


class CDims(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("npts", "nz", "nztmax", "nsflxs", "njdt", "maxmodeadv")]


_CONST_D = ["dto", "grav", "vonk", "sice", "hmixtolfrac", "iso_thresh"]
_CONST_I = ["itermax", "iso_bot", "dt_uvdamp", "LKPP", "LRI", "LDD", "L_SSref", "L_RELAX_SST", "L_RELAX_CALCONLY",
            "L_FCORR", "L_FCORR_WITHZ", "L_SFCORR", "L_SFCORR_WITHZ", "L_RELAX_SAL", "L_RELAX_OCNT", "L_NO_FREEZE",
            "L_NO_ISOTHERM", "L_DAMP_CURR", "L_VARY_BOTTOM_TEMP", "have_ocnT_file", "have_sal_file", "numerics",
            "reserved"]


class CConsts(C.Structure):
    _fields_ = [(n, C.c_double) for n in _CONST_D] + [(n, C.c_int32) for n in _CONST_I]


class StepReport(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("ntime", "n_active", "n_long_iter", "n_reint", "n_reint_fail", "n_reset",
                                          "n_pivot_zero", "n_iter_cap", "max_iter", "n_handed_over")] + \
               [("sum_iter", C.c_int64), ("kernel_ms", C.c_float), ("reserved2", C.c_float)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_ if not n.startswith("reserved")}


def header_field_ids() -> dict:
    """Parse ``enum kpp_field_id`` from the header so Python never drifts from the ABI."""
    txt = open(HEADER_PATH).read()
    body = txt[txt.index("typedef enum kpp_field_id"):txt.index("} kpp_field_id;")]
    names = re.findall(r"^\s*(KPP_F_\w+)", body, flags=re.M)
    return {n: i for i, n in enumerate(names)}


def out_ids() -> dict:
    """Parse ``enum kpp_out_id`` (SURVEY 8(f2) output sets) from the header."""
    txt = open(HEADER_PATH).read()
    body = txt[txt.index("typedef enum kpp_out_id"):txt.index("} kpp_out_id;")]
    names = re.findall(r"^\s*(KPP_OUT_\w+)", body, flags=re.M)
    return {n: i for i, n in enumerate(names)}


FIELD_IDS = header_field_ids()
# python-side names of kpp_3d_fields members -> field id
FIELD_BY_NAME = {}

_EXPORTS = ["kpp_gpu_abi_version", "kpp_gpu_device_count", "kpp_gpu_strerror", "kpp_gpu_last_error", "kpp_gpu_create",
            "kpp_gpu_destroy", "kpp_gpu_upload_field", "kpp_gpu_download_field", "kpp_gpu_field_host_bytes",
            "kpp_gpu_field_name", "kpp_gpu_upload_forcing", "kpp_gpu_init_vmix", "kpp_gpu_step", "kpp_gpu_set_pass_budget", "kpp_gpu_sync", "kpp_gpu_output_name", "kpp_gpu_output_rows",
            "kpp_gpu_pack_output", "kpp_gpu_pack_output_async", "kpp_gpu_upload_clim_record", "kpp_gpu_blend_clim",
            "kpp_gpu_get_status", "kpp_gpu_host_alloc", "kpp_gpu_host_free", "kpp_gpu_test_eos",
            "kpp_gpu_test_wscale", "kpp_gpu_test_swfrac", "kpp_gpu_test_div"]

_LIB = None


def header_exports() -> list:
    """Function names declared in include/kpp_gpu.h."""
    txt = open(HEADER_PATH).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(kpp_gpu_\w+)\s*\(", txt)))


def load():
    """Load libkpp_gpu.so; fail loudly when it is missing (no fallback)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise KppError(KPP_E_NODEVICE, f"{LIB_PATH} not built: run `python -m mckpp_f90_b200.build` "
                                       "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, i32, dp = C.c_void_p, C.c_int, C.POINTER(C.c_double)
    L.kpp_gpu_abi_version.restype = i32
    L.kpp_gpu_device_count.restype = i32
    L.kpp_gpu_exp_is_host_libm.restype = i32
    L.kpp_gpu_exp_is_host_libm.argtypes = [i32]
    L.kpp_gpu_strerror.restype = C.c_char_p
    L.kpp_gpu_strerror.argtypes = [i32]
    L.kpp_gpu_last_error.restype = C.c_char_p
    L.kpp_gpu_last_error.argtypes = [vp]
    L.kpp_gpu_create.restype = i32
    L.kpp_gpu_create.argtypes = [C.POINTER(CDims), C.POINTER(CConsts), vp, vp, vp, vp, vp, vp, i32, C.POINTER(vp)]
    L.kpp_gpu_destroy.restype = i32
    L.kpp_gpu_destroy.argtypes = [vp]
    L.kpp_gpu_create_multi.restype = i32
    L.kpp_gpu_create_multi.argtypes = [C.POINTER(CDims), C.POINTER(CConsts), vp, vp, vp, vp, vp, vp, i32, vp, C.POINTER(vp)]
    L.kpp_gpu_num_parts.restype = i32
    L.kpp_gpu_num_parts.argtypes = [vp]
    L.kpp_gpu_part_columns.restype = i32
    L.kpp_gpu_part_columns.argtypes = [vp, i32, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]
    L.kpp_gpu_output_ring_create.restype = i32
    L.kpp_gpu_output_ring_create.argtypes = [vp, vp, i32, i32]
    L.kpp_gpu_output_ring_slot_bytes.restype = C.c_size_t
    L.kpp_gpu_output_ring_slot_bytes.argtypes = [vp]
    L.kpp_gpu_output_ring_offset.restype = C.c_size_t
    L.kpp_gpu_output_ring_offset.argtypes = [vp, i32]
    L.kpp_gpu_output_ring_submit.restype = i32
    L.kpp_gpu_output_ring_submit.argtypes = [vp, C.POINTER(i32)]
    L.kpp_gpu_output_ring_wait.restype = i32
    L.kpp_gpu_output_ring_wait.argtypes = [vp, i32, C.POINTER(vp)]
    L.kpp_gpu_output_ring_destroy.restype = i32
    L.kpp_gpu_output_ring_destroy.argtypes = [vp]
    for fn in (L.kpp_gpu_upload_field, L.kpp_gpu_download_field, L.kpp_gpu_download_field_async):
        fn.restype = i32
        fn.argtypes = [vp, i32, vp, C.c_size_t]
    L.kpp_gpu_field_host_bytes.restype = C.c_size_t
    L.kpp_gpu_field_host_bytes.argtypes = [vp, i32]
    L.kpp_gpu_field_name.restype = C.c_char_p
    L.kpp_gpu_field_name.argtypes = [i32]
    L.kpp_gpu_upload_forcing.restype = i32
    L.kpp_gpu_upload_forcing.argtypes = [vp, vp]
    L.kpp_gpu_init_vmix.restype = i32
    L.kpp_gpu_init_vmix.argtypes = [vp]
    L.kpp_gpu_upload_fluxes.restype = i32
    L.kpp_gpu_upload_fluxes.argtypes = [vp] + [vp] * 8 + [C.c_double, C.c_double]
    L.kpp_gpu_reserve_forcing_slots.restype = i32
    L.kpp_gpu_reserve_forcing_slots.argtypes = [vp, i32]
    L.kpp_gpu_upload_forcing_slot.restype = i32
    L.kpp_gpu_upload_forcing_slot.argtypes = [vp, i32, vp]
    L.kpp_gpu_select_forcing_slot.restype = i32
    L.kpp_gpu_select_forcing_slot.argtypes = [vp, i32]
    L.kpp_gpu_launch_count.restype = C.c_longlong
    L.kpp_gpu_launch_count.argtypes = [vp]
    L.kpp_gpu_step.restype = i32
    L.kpp_gpu_step.argtypes = [vp, i32]
    L.kpp_gpu_output_name.restype = C.c_char_p
    L.kpp_gpu_output_name.argtypes = [i32]
    L.kpp_gpu_output_rows.restype = i32
    L.kpp_gpu_output_rows.argtypes = [vp, i32]
    for fn in (L.kpp_gpu_pack_output, L.kpp_gpu_pack_output_async):
        fn.restype = i32
        fn.argtypes = [vp, i32, vp, C.c_size_t]
    L.kpp_gpu_upload_clim_record.restype = i32
    L.kpp_gpu_upload_clim_record.argtypes = [vp, i32, i32, vp, C.c_size_t]
    L.kpp_gpu_blend_clim.restype = i32
    L.kpp_gpu_blend_clim.argtypes = [vp, i32, C.c_double, C.c_double]
    L.kpp_gpu_set_pass_budget.restype = i32
    L.kpp_gpu_set_pass_budget.argtypes = [vp, i32]
    L.kpp_gpu_sync.restype = i32
    L.kpp_gpu_sync.argtypes = [vp, C.POINTER(StepReport)]
    L.kpp_gpu_debug_check_guards.restype = i32
    L.kpp_gpu_debug_check_guards.argtypes = [vp]
    L.kpp_gpu_set_async_stragglers.restype = i32
    L.kpp_gpu_set_async_stragglers.argtypes = [vp, i32]
    L.kpp_gpu_get_status.restype = i32
    L.kpp_gpu_get_status.argtypes = [vp, vp]
    L.kpp_gpu_host_alloc.restype = i32
    L.kpp_gpu_host_alloc.argtypes = [C.POINTER(vp), C.c_size_t]
    L.kpp_gpu_host_free.restype = i32
    L.kpp_gpu_host_free.argtypes = [vp]
    L.kpp_gpu_test_eos.restype = i32
    L.kpp_gpu_test_eos.argtypes = [i32, i32, i32, vp, vp, vp, vp, vp, vp, vp]
    L.kpp_gpu_test_wscale.restype = i32
    L.kpp_gpu_test_wscale.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp]
    L.kpp_gpu_test_swfrac.restype = i32
    L.kpp_gpu_test_swfrac.argtypes = [i32, i32, i32, vp, vp, vp]
    L.kpp_gpu_test_div.restype = i32
    L.kpp_gpu_test_div.argtypes = [i32, i32, i32, vp, vp, vp, vp, vp]
    _LIB = L
    # map member names (as in fields.py) to ids using the library's own names
    for name, fid in FIELD_IDS.items():
        if name == "KPP_F__COUNT":
            continue
        FIELD_BY_NAME[L.kpp_gpu_field_name(fid).decode()] = fid
    return L


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def pinned_empty(shape, dtype=np.float64, order="F") -> np.ndarray:
    """numpy array backed by page-locked host memory (kept for the life of the process)."""
    L = load()
    dtype = np.dtype(dtype)
    count = int(np.prod(shape))
    n = count * dtype.itemsize
    ptr = C.c_void_p()
    rc = L.kpp_gpu_host_alloc(C.byref(ptr), n)
    if rc != 0:
        raise KppError(rc, L.kpp_gpu_last_error(None).decode())
    buf = (C.c_char * max(n, 1)).from_address(ptr.value)
    _PINNED.append((ptr, buf))
    return np.frombuffer(buf, dtype=dtype, count=count).reshape(shape, order=order)


_PINNED = []


class KppGpu:
    """One handle: all columns on one GPU, or -- with ``ngpus``/``devices`` -- block-partitioned over
    several GPUs of the node inside the library (kpp_gpu_create_multi); every method is the same."""

    def __init__(self, cf: KppConstFields, device: int = 0, numerics: int = 0, ngpus: int = None, devices=None):
        self.L = load()
        d, k = cf.dims, cf.consts
        self.dims = d
        cd = CDims(d.npts, d.nz, d.nztmax, d.nsflxs, d.njdt, d.maxmodeadv)
        cc = CConsts()
        for n in _CONST_D:
            setattr(cc, n, float(getattr(k, n)))
        for n in _CONST_I:
            if n in ("numerics", "reserved"):
                continue
            setattr(cc, n, int(getattr(k, n)))
        cc.numerics = int(numerics)
        arrs = [np.ascontiguousarray(np.asarray(getattr(cf, n), dtype=np.float64).ravel(order="F"))
                for n in ("zm", "hm", "dm", "tri", "wmt", "wst")]
        h = C.c_void_p()
        if ngpus is None and devices is None:
            rc = self.L.kpp_gpu_create(C.byref(cd), C.byref(cc), *[_p(a) for a in arrs], int(device), C.byref(h))
        else:
            dv = None if devices is None else np.ascontiguousarray(devices, dtype=np.int32)
            n = int(ngpus) if ngpus is not None else int(dv.size)
            rc = self.L.kpp_gpu_create_multi(C.byref(cd), C.byref(cc), *[_p(a) for a in arrs], n,
                                             None if dv is None else _p(dv), C.byref(h))
        if rc != 0:
            raise KppError(rc, self.L.kpp_gpu_last_error(None).decode())
        self.h = h
        self.numerics = numerics

    def _check(self, rc):
        if rc != 0:
            raise KppError(rc, self.L.kpp_gpu_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.L.kpp_gpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def parts(self):
        """[(device, first column, columns)] of the devices behind the handle."""
        out = []
        for i in range(self.L.kpp_gpu_num_parts(self.h)):
            d, c0, n = C.c_int(), C.c_int(), C.c_int()
            self._check(self.L.kpp_gpu_part_columns(self.h, i, C.byref(d), C.byref(c0), C.byref(n)))
            out.append((d.value, c0.value, n.value))
        return out

    # ---- asynchronous output ring: device->host copies of the output set overlap the next step
    def output_ring_create(self, ids, depth: int = 2):
        ids = np.ascontiguousarray(list(ids), dtype=np.int32)
        self._check(self.L.kpp_gpu_output_ring_create(self.h, _p(ids), int(ids.size), int(depth)))
        self._ring_ids = [int(i) for i in ids]
        self._ring_rows = [self.L.kpp_gpu_output_rows(self.h, i) for i in self._ring_ids]
        self._ring_off = [self.L.kpp_gpu_output_ring_offset(self.h, k) for k in range(ids.size)]
        return self.L.kpp_gpu_output_ring_slot_bytes(self.h)

    def output_ring_submit(self) -> int:
        slot = C.c_int(-1)
        self._check(self.L.kpp_gpu_output_ring_submit(self.h, C.byref(slot)))
        return slot.value

    def output_ring_wait(self, slot: int) -> dict:
        """{out_id: array(npts[, rows]) viewing the pinned slot} -- valid until the slot is resubmitted."""
        ptr = C.c_void_p()
        self._check(self.L.kpp_gpu_output_ring_wait(self.h, int(slot), C.byref(ptr)))
        out = {}
        npts = self.dims.npts
        for oid, rows, off in zip(self._ring_ids, self._ring_rows, self._ring_off):
            buf = (C.c_double * (npts * rows)).from_address(ptr.value + off)
            a = np.frombuffer(buf, dtype=np.float64)
            out[oid] = a.reshape((npts, rows), order="F") if rows > 1 else a
        return out

    def output_ring_destroy(self):
        self._check(self.L.kpp_gpu_output_ring_destroy(self.h))

    def upload(self, name: str, arr: np.ndarray):
        fid = FIELD_BY_NAME[name]
        assert arr.flags.f_contiguous or arr.ndim == 1, name
        self._check(self.L.kpp_gpu_upload_field(self.h, fid, _p(arr), arr.nbytes))

    def download(self, name: str, arr: np.ndarray):
        fid = FIELD_BY_NAME[name]
        assert arr.flags.f_contiguous or arr.ndim == 1, name
        self._check(self.L.kpp_gpu_download_field(self.h, fid, _p(arr), arr.nbytes))

    def download_async(self, name: str, arr: np.ndarray):
        """Enqueue only; valid after the next sync()/download()."""
        fid = FIELD_BY_NAME[name]
        assert arr.flags.f_contiguous or arr.ndim == 1, name
        self._check(self.L.kpp_gpu_download_field_async(self.h, fid, _p(arr), arr.nbytes))

    def upload_forcing(self, sflux6: np.ndarray):
        assert sflux6.dtype == np.float64 and sflux6.flags.c_contiguous and sflux6.shape == (6, self.dims.npts)
        self._check(self.L.kpp_gpu_upload_forcing(self.h, _p(sflux6)))

    def upload_fluxes(self, taux, tauy, swf, lwf, lhf, shf, rain, snow, flsn: float, el: float):
        """mckpp_fluxes' forcing map on the device (fluxes_mod.F90:56-72)."""
        arrs = [np.ascontiguousarray(x, dtype=np.float64) for x in (taux, tauy, swf, lwf, lhf, shf, rain, snow)]
        assert all(a.shape == (self.dims.npts,) for a in arrs)
        self._check(self.L.kpp_gpu_upload_fluxes(self.h, *[_p(a) for a in arrs], float(flsn), float(el)))

    def reserve_forcing_slots(self, n: int):
        self._check(self.L.kpp_gpu_reserve_forcing_slots(self.h, int(n)))

    def upload_forcing_slot(self, slot: int, sflux6: np.ndarray):
        assert sflux6.dtype == np.float64 and sflux6.flags.c_contiguous and sflux6.shape == (6, self.dims.npts)
        self._check(self.L.kpp_gpu_upload_forcing_slot(self.h, int(slot), _p(sflux6)))

    def select_forcing_slot(self, slot: int):
        self._check(self.L.kpp_gpu_select_forcing_slot(self.h, int(slot)))

    def launch_count(self) -> int:
        return int(self.L.kpp_gpu_launch_count(self.h))

    # ---- output sets packed on the device
    def output_ids(self, restart=False):
        """{out_id: xios field id} of the diagnostic (default) or restart set."""
        first_restart = out_ids()["KPP_OUT_R_UVEL"]
        n = out_ids()["KPP_OUT__COUNT"]
        rng = range(first_restart, n) if restart else range(0, first_restart)
        return {i: self.L.kpp_gpu_output_name(i).decode() for i in rng}

    def pack_output(self, out_id: int, host: np.ndarray = None, sync=True) -> np.ndarray:
        rows = self.L.kpp_gpu_output_rows(self.h, int(out_id))
        if rows < 0:
            raise KppError(rows, "unknown output id")
        if host is None:
            host = np.empty((self.dims.npts, rows) if rows > 1 else (self.dims.npts,), dtype=np.float64, order="F")
        fn = self.L.kpp_gpu_pack_output if sync else self.L.kpp_gpu_pack_output_async
        self._check(fn(self.h, int(out_id), _p(host), host.nbytes))
        return host

    # ---- climatology records resident on the device, blended there
    def upload_clim_record(self, name: str, which: int, record: np.ndarray):
        rec = np.asfortranarray(record, dtype=np.float64)
        self._check(self.L.kpp_gpu_upload_clim_record(self.h, FIELD_BY_NAME[name], int(which), _p(rec), rec.nbytes))

    def blend_clim(self, name: str, prev_weight: float, next_weight: float):
        self._check(self.L.kpp_gpu_blend_clim(self.h, FIELD_BY_NAME[name], float(prev_weight), float(next_weight)))

    def set_pass_budget(self, budget: int):
        """Scheduling knob (kpp_gpu_set_pass_budget): passes a column iterates in the per-thread
        kernel before the cooperative kernel takes it over; 0 = never.  No effect on results."""
        self._check(self.L.kpp_gpu_set_pass_budget(self.h, int(budget)))

    def check_guards(self) -> int:
        """KPP_GUARD=1 debugging aid: canary zones around the device arrays that a kernel wrote into (0 = clean)."""
        rc = self.L.kpp_gpu_debug_check_guards(self.h)
        if rc < 0:
            self._check(rc)
        return rc

    def set_async_stragglers(self, on: bool = True):
        """Scheduling knob (kpp_gpu_set_async_stragglers): hand-overs finish on a second stream while the next
        step of the other columns runs; every call that touches device state joins first.  No effect on results."""
        self._check(self.L.kpp_gpu_set_async_stragglers(self.h, int(bool(on))))

    def init_vmix(self):
        self._check(self.L.kpp_gpu_init_vmix(self.h))

    def step(self, ntime: int):
        self._check(self.L.kpp_gpu_step(self.h, int(ntime)))

    def sync(self) -> StepReport:
        rep = StepReport()
        rc = self.L.kpp_gpu_sync(self.h, C.byref(rep))
        self.last_report = rep
        self._check(rc)
        return rep

    def status(self) -> np.ndarray:
        st = np.zeros(self.dims.npts, np.int32)
        self._check(self.L.kpp_gpu_get_status(self.h, _p(st)))
        return st


def test_eos(S, T, P, numerics=0, device=0):
    L = load()
    S, T, P = (np.ascontiguousarray(x, dtype=np.float64) for x in (S, T, P))
    out = [np.empty_like(S) for _ in range(4)]
    rc = L.kpp_gpu_test_eos(device, numerics, S.size, _p(S), _p(T), _p(P), *[_p(o) for o in out])
    if rc != 0:
        raise KppError(rc, L.kpp_gpu_last_error(None).decode())
    return out   # sig0, alpha, beta, cp


def test_swfrac(z, jerlov, numerics=0, device=0):
    L = load()
    z = np.ascontiguousarray(z, dtype=np.float64)
    j = np.ascontiguousarray(jerlov, dtype=np.int32)
    out = np.empty_like(z)
    rc = L.kpp_gpu_test_swfrac(device, numerics, z.size, _p(z), _p(j), _p(out))
    if rc != 0:
        raise KppError(rc, L.kpp_gpu_last_error(None).decode())
    return out


def test_div(a, b, numerics=0, device=0):
    L = load()
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    plain, split, ok = np.empty_like(a), np.empty_like(a), np.empty_like(a)
    rc = L.kpp_gpu_test_div(device, numerics, a.size, _p(a), _p(b), _p(plain), _p(split), _p(ok))
    if rc != 0:
        raise KppError(rc, L.kpp_gpu_last_error(None).decode())
    return plain, split, ok != 0


def test_wscale(gpu: KppGpu, sigma, hbl, ustar, bfsfc):
    a = [np.ascontiguousarray(x, dtype=np.float64) for x in (sigma, hbl, ustar, bfsfc)]
    wm, ws = np.empty_like(a[0]), np.empty_like(a[0])
    gpu._check(gpu.L.kpp_gpu_test_wscale(gpu.h, a[0].size, *[_p(x) for x in a], _p(wm), _p(ws)))
    return wm, ws
