"""Per-step kernel time and iteration statistics over many steps (tail analysis)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from mckpp_f90_b200 import synth, driver
name, nx, ny, nst = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
cfg = synth.scaled(synth.CONFIGS[name], nx, ny)
cf, f, r = synth.make_case(cfg)
m = driver.MckppPhysics(cf, f, numerics=0)
synth.apply_forcing(cfg, cf, f, r, 1)
m.push_inputs(); m.mckpp_initialize_ocean_model()
for nt in range(1, nst + 1):
    synth.apply_forcing(cfg, cf, f, r, nt)
    rep = m.mckpp_physics_driver(nt)
    if rep.max_iter > 6 or nt % 12 == 0:
        m.pull_diag()
        it = m.diag["iter"]
        print(f"nt={nt} ms={rep.kernel_ms:.2f} mean_iter={rep.sum_iter/rep.n_active:.3f} max_iter={rep.max_iter} n>6={(it>6).sum()} n>12={(it>12).sum()} n>50={(it>50).sum()} long={rep.n_long_iter} hmix_max={f['hmix'].max():.1f}")
