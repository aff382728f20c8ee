"""Small strict run for compute-sanitizer: every column finishes in the cooperative kernel (budget 1),
then a few output blocks are packed.  python tools/sanitize_small.py [budget]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from mckpp_f90_b200 import synth, driver, capi
budget = int(sys.argv[1]) if len(sys.argv) > 1 else 1
for name, nx, ny, nst in (("cfg2", 8, 6, 3), ("cfg5", 6, 4, 2), ("cfg4", 8, 5, 2)):
    cfg = synth.scaled(synth.CONFIGS[name], nx, ny)
    cf, f, r = synth.make_case(cfg)
    m = driver.MckppPhysics(cf, f, numerics=0)
    m.gpu.set_pass_budget(budget)
    synth.apply_forcing(cfg, cf, f, r, 1)
    m.push_inputs(); m.mckpp_initialize_ocean_model()
    for nt in range(1, nst + 1):
        synth.apply_forcing(cfg, cf, f, r, nt)
        rep = m.mckpp_physics_driver(nt)
    out = m.mckpp_xios_diagnostic_output()
    print(name, "handed over", rep.n_handed_over, "max_iter", rep.max_iter, "hmix", float(f["hmix"].mean()), "S", float(out["S"].mean()))
    m.close()
