"""Step latency of small domains: one thread per column (budget 6) against the whole integration
in the cooperative kernel (budget -1).  python tools/small_domain.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from mckpp_f90_b200 import synth, driver
for nx, ny in ((4, 4), (16, 8), (32, 32), (64, 32), (64, 64), (128, 64), (128, 128)):
    row = []
    for budget in (6, -1):
        cfg = synth.scaled(synth.CONFIGS["cfg2"], nx, ny)
        cf, f, r = synth.make_case(cfg)
        m = driver.MckppPhysics(cf, f, numerics=0)
        m.gpu.set_pass_budget(budget)
        synth.apply_forcing(cfg, cf, f, r, 1)
        m.push_inputs(); m.mckpp_initialize_ocean_model()
        ms = []
        for nt in range(1, 13):
            synth.apply_forcing(cfg, cf, f, r, nt)
            ms.append(m.mckpp_physics_driver(nt).kernel_ms)
        row.append(float(np.median(ms[2:])))
        m.close()
    print(f"npts={nx*ny:6d}  per-thread {row[0]:7.3f} ms   cooperative {row[1]:7.3f} ms   ratio {row[0]/row[1]:5.2f}")
