"""Sustained step time with and without asynchronous stragglers: steps queued on device-resident forcing,
one kpp_gpu_sync per `group` steps.  python tools/async_timing.py cfg2 300 200 [spinup=80] [timed=50] [group=10] [only modes containing this word]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from mckpp_f90_b200 import synth, driver

name, nx, ny = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
spin = int(sys.argv[4]) if len(sys.argv) > 4 else 80
timed = int(sys.argv[5]) if len(sys.argv) > 5 else 50
group = int(sys.argv[6]) if len(sys.argv) > 6 else 10
only = sys.argv[7] if len(sys.argv) > 7 else ""
cfg = synth.scaled(synth.CONFIGS[name], nx, ny)
for mode in ("sync every step", "queued, synchronous hand-over", "queued, asynchronous stragglers"):
    if only and only not in mode:
        continue
    cf, f, r = synth.make_case(cfg)
    m = driver.MckppPhysics(cf, f, numerics=0)
    g = m.gpu
    g.set_async_stragglers(mode.endswith("stragglers"))
    synth.apply_forcing(cfg, cf, f, r, 1)
    m.push_inputs(); m.mckpp_initialize_ocean_model()
    g.reserve_forcing_slots(72)
    for i in range(72):
        g.upload_forcing_slot(i, synth.apply_forcing(cfg, cf, f, r, i + 1))
    nt = 0
    for _ in range(spin):
        nt += 1; g.select_forcing_slot((nt - 1) % 72); g.step(nt)
    g.sync()
    t0 = time.perf_counter()
    handed = 0; mx = 0
    for i in range(timed):
        nt += 1; g.select_forcing_slot((nt - 1) % 72); g.step(nt)
        if mode.startswith("sync") or (i + 1) % group == 0 or i + 1 == timed:
            rep = g.sync(); handed += rep.n_handed_over; mx = max(mx, rep.max_iter)
    dt = time.perf_counter() - t0
    m.pull(["hmix"])
    print(f"{name} npts={cfg.npts} steps {spin+1}..{spin+timed}: {mode:34s} {1e3*dt/timed:7.3f} ms/step = {cfg.npts*timed/dt/1e6:6.2f} M/s  (max_iter {mx}, hmix sum {f['hmix'].sum():.6f})", flush=True)
    m.close()
