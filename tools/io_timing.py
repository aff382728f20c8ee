"""Timing of the SURVEY 8(f2)/(f4) stages at cfg2 size (60,000 columns): packing + D2H of the XIOS
diagnostic and restart sets into pinned host memory, and the climatology blend.  python tools/io_timing.py"""
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
for nt in range(1, 4):
    synth.apply_forcing(cfg, cf, f, r, nt)
    m.mckpp_physics_driver(nt)
g = m.gpu
for restart in (False, True):
    ids = g.output_ids(restart=restart)
    bufs = {}
    for oid in ids:
        rows = g.L.kpp_gpu_output_rows(g.h, oid)
        bufs[oid] = capi.pinned_empty((cfg.npts, rows) if rows > 1 else (cfg.npts,))
    nbytes = sum(b.nbytes for b in bufs.values())
    for rep in range(3):
        g.sync()
        t0 = time.perf_counter()
        for oid, b in bufs.items():
            g.pack_output(oid, host=b, sync=False)
        g.sync()
        dt = time.perf_counter() - t0
    print(f"{'restart' if restart else 'diagnostic'} set: {len(ids)} blocks, {nbytes/1e6:.0f} MB to pinned host memory in {dt*1e3:.1f} ms = {nbytes/dt/1e9:.1f} GB/s")
nzp1 = cfg.nz + 1
prev = np.asfortranarray(np.random.default_rng(0).standard_normal((cfg.npts, nzp1)))
nxt = np.asfortranarray(np.random.default_rng(1).standard_normal((cfg.npts, nzp1)))
g.upload_clim_record("ocnT_clim", 0, prev); g.upload_clim_record("ocnT_clim", 1, nxt); g.sync()
for rep in range(3):
    t0 = time.perf_counter()
    for i in range(20):
        g.blend_clim("ocnT_clim", 0.25 + 0.01 * i, 0.75 - 0.01 * i)
    g.sync()
    dt = (time.perf_counter() - t0) / 20
print(f"climatology blend: {3*cfg.npts*nzp1*8/1e6:.0f} MB of HBM traffic in {dt*1e6:.0f} us = {3*cfg.npts*nzp1*8/dt/1e12:.2f} TB/s")
m.close()
