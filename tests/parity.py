"""Shared helpers of the parity tests: run the oracle and the CUDA path on the
same host memory image and compare every field 1dto3d writes back."""
from __future__ import annotations

import os

import numpy as np

import oracle_lib
from mckpp_f90_b200 import synth, driver
from mckpp_f90_b200.fields import copy_fields

FLOAT_FIELDS = ["U", "X", "Us", "Xs", "hmixd", "hmix", "Tref", "uref", "vref", "Ssurf", "rho", "cp", "buoy", "Rig",
                "dbloc", "Shsq", "difm", "difs", "dift", "ghat", "wU", "wX", "wXNT", "tinc_fcorr", "sinc_fcorr",
                "ocnTcorr", "scorr", "swfrac", "swdk_opt", "freeze_flag", "dampu_flag", "dampv_flag", "fcorr"]
INT_FIELDS = ["kmix", "old", "new", "reset_flag"]


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.maximum(np.abs(a), np.abs(b))
    diff = np.abs(a - b)
    with np.errstate(invalid="ignore", divide="ignore"):
        r = np.where(den > 0, diff / den, 0.0)
    r = np.where(np.isnan(a) | np.isnan(b), np.where(np.isnan(a) & np.isnan(b), 0.0, np.inf), r)
    return float(r.max()) if r.size else 0.0


def scaled_err(a, b):
    """max |a-b| / max|b| over the whole field: robust for fields that cross zero."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    s = float(np.max(np.abs(b))) if b.size else 0.0
    if s == 0.0:
        return float(np.max(np.abs(a))) if a.size else 0.0
    return float(np.max(np.abs(a - b))) / s


class Pair:
    """Oracle and GPU side by side on identical inputs."""

    def __init__(self, cfg, numerics=0, device=0, nthreads=0, consts=None, setup=None, itermax=None, budget=None):
        """consts: dict of KppConsts overrides; setup(cf, fields, r): mutate the inputs (both sides get a copy).
        budget: kpp_gpu_set_pass_budget.  The library's default sends domains this small straight to the
        cooperative kernel; the parity tests pin 6 (stragglers only) unless told otherwise, so that the
        one-thread-per-column kernel is what they exercise (KPP_TEST_PASS_BUDGET overrides for a whole run)."""
        self.cfg = cfg
        self.cf, self.f_orc, self.r = synth.make_case(cfg)
        for k, v in (consts or {}).items():
            assert hasattr(self.cf.consts, k), k
            setattr(self.cf.consts, k, v)
        if setup is not None:
            setup(self.cf, self.f_orc, self.r)
        self.f_gpu = copy_fields(self.f_orc)
        self.orc = oracle_lib.Oracle(self.cf, self.f_orc, nthreads=nthreads)
        self.gpu = driver.MckppPhysics(self.cf, self.f_gpu, device=device, numerics=numerics, sync_mode="full")
        self.gpu.gpu.set_pass_budget(int(os.environ.get("KPP_TEST_PASS_BUDGET", "6")) if budget is None else budget)
        self.gpu.push_inputs()

    def forcing(self, nt):
        synth.apply_forcing(self.cfg, self.cf, self.f_orc, self.r, nt)
        self.f_gpu["sflux"][...] = self.f_orc["sflux"]

    def init(self):
        self.forcing(1)
        self.gpu.push_inputs(["sflux"])
        self.orc.initialize_ocean_model()
        self.gpu.mckpp_initialize_ocean_model()
        self.gpu.pull(driver.ALL_OUTPUTS)

    def step(self, nt, teacher_forced=False):
        self.forcing(nt)
        if teacher_forced:
            # both sides start the step from the oracle's state
            for name in driver.INPUT_FIELDS:
                self.f_gpu[name][...] = self.f_orc[name]
            self.gpu.push_inputs()
        rc = self.orc.physics_driver(nt)
        rep = self.gpu.mckpp_physics_driver(nt)
        self.gpu.pull_diag()
        return rc, rep

    def compare(self, fields=None):
        out = {}
        run = self.f_orc["run_physics"] != 0
        for name in (FLOAT_FIELDS if fields is None else fields):
            a, b = self.f_gpu[name][run], self.f_orc[name][run]
            out[name] = (rel_err(a, b), scaled_err(a, b))
        return out

    def int_mismatches(self):
        """Enumerates integer-output differences: {name: [(column, gpu, oracle), ...]}"""
        out = {}
        run = self.f_orc["run_physics"] != 0
        pairs = [(n, self.f_gpu[n], self.f_orc[n]) for n in INT_FIELDS]
        pairs += [("iter", self.gpu.diag["iter"], self.orc.diag["iter"]),
                  ("nreint", self.gpu.diag["nreint"], self.orc.diag["nreint"]),
                  ("status", self.gpu.diag["status"], self.orc.diag["status"])]
        for name, a, b in pairs:
            bad = np.nonzero((np.asarray(a) != np.asarray(b)) & run)[0]
            out[name] = [(int(i), float(a[i]), float(b[i])) for i in bad[:50]]
            out[name + "_count"] = int(bad.size)
        return out

    def close(self):
        self.gpu.close()
