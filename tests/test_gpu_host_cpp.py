"""The compiled C++ host (mckpp_f90_b200/host) over the C ABI reproduces the oracle:
the same boundary a Fortran ISO_C_BINDING host would use, exercised from compiled code."""
import ctypes as C
import os
import struct
import subprocess

import numpy as np
import pytest

import oracle_lib
from mckpp_f90_b200 import capi, driver, synth, build as kbuild

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_host_matches_oracle(tmp_path):
    exe = os.path.join(ROOT, "mckpp_f90_b200", "host", "kpp_host_demo")
    if not os.path.exists(exe):
        kbuild.build_host_demo()
    cfg = synth.scaled(synth.CONFIGS["cfg2"], 12, 8)
    nsteps = 5
    cf, f, r = synth.make_case(cfg)
    synth.apply_forcing(cfg, cf, f, r, 1)
    L = capi.load()
    d, k = cf.dims, cf.consts
    cc = capi.CConsts()
    for n in capi._CONST_D:
        setattr(cc, n, float(getattr(k, n)))
    for n in capi._CONST_I:
        if n not in ("numerics", "reserved"):
            setattr(cc, n, int(getattr(k, n)))
    case = tmp_path / "case.bin"
    names = [n for n in driver.INPUT_FIELDS if n not in ("tinc_fcorr", "wXNT", "reset_flag", "dampu_flag", "dampv_flag")]
    with open(case, "wb") as fp:
        fp.write(struct.pack("<8i", d.npts, d.nz, d.nztmax, d.nsflxs, d.njdt, d.maxmodeadv, nsteps, len(names)))
        fp.write(bytes(cc))
        for a in (cf.zm, cf.hm, cf.dm, cf.tri, cf.wmt, cf.wst):
            fp.write(np.asarray(a, dtype=np.float64).tobytes(order="F"))
        for n in names:
            a = f[n]
            fp.write(struct.pack("<iq", capi.FIELD_BY_NAME[n], a.nbytes))
            fp.write(a.tobytes(order="F"))
        forc = []
        for nt in range(1, nsteps + 1):
            forc.append(synth.apply_forcing(cfg, cf, f, r, nt))
            fp.write(forc[-1].tobytes())
    out = tmp_path / "out.bin"
    res = subprocess.run([exe, str(case), str(out)], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    # oracle on the same inputs
    cf2, f2, r2 = synth.make_case(cfg)
    orc = oracle_lib.Oracle(cf2, f2)
    synth.apply_forcing(cfg, cf2, f2, r2, 1)
    orc.initialize_ocean_model()
    for nt in range(1, nsteps + 1):
        synth.apply_forcing(cfg, cf2, f2, r2, nt)
        orc.physics_driver(nt)
    raw = np.fromfile(out, dtype=np.float64)
    n, nzp1 = d.npts, d.nzp1
    X = raw[:2 * nzp1 * n].reshape((n, nzp1, 2), order="F")
    U = raw[2 * nzp1 * n:4 * nzp1 * n].reshape((n, nzp1, 2), order="F")
    hmix = raw[4 * nzp1 * n:4 * nzp1 * n + n]
    kmix = raw[4 * nzp1 * n + n:4 * nzp1 * n + 2 * n]
    assert np.array_equal(kmix, f2["kmix"])
    assert np.allclose(X, f2["X"], rtol=1e-12, atol=0) and np.allclose(hmix, f2["hmix"], rtol=1e-12)
    assert np.allclose(U, f2["U"], rtol=1e-10, atol=1e-300)
    # packed XIOS blocks from compiled code: S = X2 + Sref, difm shifted under a zero row
    o = 4 * nzp1 * n + 2 * n
    S = raw[o:o + nzp1 * n].reshape((n, nzp1), order="F")
    difm = raw[o + nzp1 * n:o + 2 * nzp1 * n].reshape((n, nzp1), order="F")
    assert np.array_equal(S, X[:, :, 1] + f2["Sref"][:, None])
    assert np.all(difm[:, 0] == 0.0) and np.allclose(difm[:, 1:], f2["difm"][:, 1:nzp1], rtol=1e-10)
