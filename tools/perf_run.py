"""GPU-only timing of the step on a scaled config (no oracle): python tools/perf_run.py cfg2 300 200 10 [numerics]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from mckpp_f90_b200 import synth, driver

name, nx, ny, nst = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
numerics = int(sys.argv[5]) if len(sys.argv) > 5 else 0
cfg = synth.scaled(synth.CONFIGS[name], nx, ny)
cf, f, r = synth.make_case(cfg)
m = driver.MckppPhysics(cf, f, numerics=numerics)
synth.apply_forcing(cfg, cf, f, r, 1)
m.push_inputs()
m.mckpp_initialize_ocean_model()
ms = []
for nt in range(1, nst + 1):
    synth.apply_forcing(cfg, cf, f, r, nt)
    rep = m.mckpp_physics_driver(nt)
    ms.append(rep.kernel_ms)
    if nt <= 3 or nt == nst:
        print(f"nt={nt} kernel_ms={rep.kernel_ms:.3f} mean_iter={rep.sum_iter/max(rep.n_active,1):.2f} max_iter={rep.max_iter} hmix={f['hmix'].mean():.2f}")
ms = np.array(ms[2:])
print(f"{name} npts={cfg.npts} nz={cfg.nz} numerics={numerics} block={os.environ.get('KPP_BLOCK','128')}: median {np.median(ms):.3f} ms/step -> {cfg.npts/np.median(ms)*1e3/1e6:.3f} M col-steps/s")
