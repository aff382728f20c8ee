"""Step time of one configuration at several domain sizes and CTA sizes (launch-shape experiments):
python tools/size_sweep.py cfg4 "350x250,500x350" "0,384,512" [nsteps]   (block 0 = the library's own choice)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from mckpp_f90_b200 import synth, driver

name, sizes, blocks = sys.argv[1], sys.argv[2].split(","), [int(b) for b in sys.argv[3].split(",")]
nst = int(sys.argv[4]) if len(sys.argv) > 4 else 9
for sz in sizes:
    nx, ny = (int(v) for v in sz.split("x"))
    cfg = synth.scaled(synth.CONFIGS[name], nx, ny)
    cf, f, r = synth.make_case(cfg)
    row = []
    for b in blocks:
        if b:
            os.environ["KPP_BLOCK"] = str(b)
        else:
            os.environ.pop("KPP_BLOCK", None)
        m = driver.MckppPhysics(cf, {k: v.copy(order="F") for k, v in f.items()}, numerics=0)
        synth.apply_forcing(cfg, cf, m.kpp_3d_fields, r, 1)
        m.push_inputs(); m.mckpp_initialize_ocean_model()
        ms = []
        for nt in range(1, nst + 1):
            synth.apply_forcing(cfg, cf, m.kpp_3d_fields, r, nt)
            ms.append(m.mckpp_physics_driver(nt).kernel_ms)
        m.close()
        row.append(float(np.median(ms[2:])))
    print(f"{name} npts={cfg.npts:7d} nz={cfg.nz} " + "  ".join(f"block {b or 'auto'}: {t:7.3f} ms = {cfg.npts / t / 1e3:6.2f} M/s" for b, t in zip(blocks, row)), flush=True)
