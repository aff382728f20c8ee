"""Device time of packing the 34 diagnostic blocks into the output ring's staging (the part that sits on the step's
stream; the device->host copy runs on the ring's own stream).  python tools/pack_timing.py   [KPP_NO_BULK_PACK=1]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from mckpp_f90_b200 import synth, driver, capi
cfg = synth.CONFIGS["cfg2"]
cf, f, r = synth.make_case(cfg)
m = driver.MckppPhysics(cf, f, numerics=0)
synth.apply_forcing(cfg, cf, f, r, 1)
m.push_inputs(); m.mckpp_initialize_ocean_model()
m.mckpp_physics_driver(1)
g = m.gpu
ids = list(range(capi.out_ids()["KPP_OUT_R_UVEL"]))
nbytes = g.output_ring_create(ids, depth=4)
ts = []
for rep in range(3):
    g.sync()
    t0 = time.perf_counter()
    slot = g.output_ring_submit()
    g.sync()                      # the step stream: packing only
    ts.append(time.perf_counter() - t0)
    g.output_ring_wait(slot)
dt = min(ts)
print(f"pack of {len(ids)} blocks ({nbytes/1e6:.0f} MB) into device staging: {dt*1e3:.3f} ms = {2*nbytes/dt/1e12:.2f} TB/s read+write "
      f"({'plain kernel' if os.environ.get('KPP_NO_BULK_PACK') else 'cp.async.bulk (TMA) where the rows allow it'})")
m.close()
