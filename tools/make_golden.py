"""Generates tests/golden/*.npz from the CPU oracle (the reference itself cannot be
built in this image -- no Fortran compiler -- so these are ORACLE outputs: they pin
the oracle against drift and give the GPU tests fixtures that need no oracle run).

    python tools/make_golden.py
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import numpy as np
import oracle_lib
from mckpp_f90_b200 import synth

KEEP = ["U", "X", "hmix", "kmix", "Tref", "difm", "difs", "dift", "ghat", "rho", "cp", "wX", "wU", "Us", "Xs", "old",
        "new", "freeze_flag", "tinc_fcorr"]
CASES = {
    "cfg1_16col_8steps": (synth.CONFIGS["cfg1"], 8),
    "cfg2_4x4_12steps": (synth.scaled(synth.CONFIGS["cfg2"], 4, 4), 12),
    "cfg4_4x4_6steps": (synth.scaled(synth.CONFIGS["cfg4"], 4, 4), 6),
    "cfg5_4x4_6steps": (synth.scaled(synth.CONFIGS["cfg5"], 4, 4), 6),
}


def run_case(cfg, nsteps):
    cf, f, r = synth.make_case(cfg)
    orc = oracle_lib.Oracle(cf, f, nthreads=1)
    synth.apply_forcing(cfg, cf, f, r, 1)
    orc.initialize_ocean_model()
    iters = []
    for nt in range(1, nsteps + 1):
        synth.apply_forcing(cfg, cf, f, r, nt)
        assert orc.physics_driver(nt) == 0
        iters.append(orc.diag["iter"].copy())
    out = {k: f[k] for k in KEEP}
    out["iter"] = np.stack(iters)
    return out


if __name__ == "__main__":
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    for name, (cfg, n) in CASES.items():
        out = run_case(cfg, n)
        path = os.path.join(ROOT, "tests", "golden", name + ".npz")
        np.savez_compressed(path, **out)
        print(name, os.path.getsize(path), "bytes")
