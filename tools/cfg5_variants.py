import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mckpp_f90_b200 import synth, driver
base = synth.CONFIGS["cfg5"]
variants = {
 "all (cfg5)": {},
 "no corrections at all": dict(L_FCORR_WITHZ=False, L_SFCORR_WITHZ=False, L_RELAX_OCNT=False, L_RELAX_SAL=False),
 "only fcorr_withz": dict(L_SFCORR_WITHZ=False, L_RELAX_OCNT=False, L_RELAX_SAL=False),
 "only relax_ocnT+relax_sal": dict(L_FCORR_WITHZ=False, L_SFCORR_WITHZ=False),
 "no freeze clamp": dict(L_NO_FREEZE=False),
}
for name, ov in variants.items():
    cf, f, r = synth.make_case(base)
    for k, v in ov.items():
        setattr(cf.consts, k, v)
    m = driver.MckppPhysics(cf, f, numerics=0)
    synth.apply_forcing(base, cf, m.kpp_3d_fields, r, 1)
    m.push_inputs(); m.mckpp_initialize_ocean_model()
    ms = []
    for nt in range(1, 8):
        synth.apply_forcing(base, cf, m.kpp_3d_fields, r, nt)
        ms.append(m.mckpp_physics_driver(nt).kernel_ms)
    m.close()
    print(f"cfg5 variant {name:28s}: {np.median(ms[2:]):7.3f} ms", flush=True)
