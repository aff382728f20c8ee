"""Manual parity run: oracle vs CUDA on a scaled config.  Usage:
   python tools/parity_run.py cfg2 40 40 72 [numerics] [teacher]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
from mckpp_f90_b200 import synth
import parity

name, nx, ny, nst = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
numerics = int(sys.argv[5]) if len(sys.argv) > 5 else 0
teacher = len(sys.argv) > 6 and sys.argv[6] == "teacher"
cfg = synth.scaled(synth.CONFIGS[name], nx, ny)
P = parity.Pair(cfg, numerics=numerics, nthreads=8)
P.init()
c = P.compare()
print("INIT worst:", sorted(((v[0], k) for k, v in c.items()), reverse=True)[:6])
print("INIT ints:", {k: v for k, v in P.int_mismatches().items() if k.endswith("_count")})
tg = 0.0
for nt in range(1, nst + 1):
    rc, rep = P.step(nt, teacher_forced=teacher)
    tg += rep.kernel_ms
    if nt <= 3 or nt % 12 == 0 or nt == nst:
        c = P.compare()
        worst = sorted(((v[1], k) for k, v in c.items()), reverse=True)[:5]
        im = P.int_mismatches()
        print(f"nt={nt} kernel_ms={rep.kernel_ms:.3f} max_iter={rep.max_iter} sum_iter={rep.sum_iter} worst(scaled)={worst}")
        print("     ints:", {k: v for k, v in im.items() if k.endswith('_count') and v})
print("total kernel ms", tg, "col-steps/s", cfg.npts * nst / (tg * 1e-3))
