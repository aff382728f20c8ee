"""Small strict runs for compute-sanitizer: the per-thread step kernel (budget 6), every column finished by the
cooperative kernel (budget 1), the asynchronous straggler lane (three streams), a 3-part group handle, the
output ring with its bulk-copy (TMA) packing.   compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from mckpp_f90_b200 import synth, driver, capi

for name, nx, ny, nst, budget, async_on, parts in (("cfg2", 8, 6, 3, 6, False, None), ("cfg2", 8, 6, 3, 1, False, None),
                                                   ("cfg5", 6, 4, 2, 1, False, None), ("cfg4", 8, 5, 3, 2, True, None),
                                                   ("cfg2", 10, 8, 4, 1, True, [0, 0, 0])):
    cfg = synth.scaled(synth.CONFIGS[name], nx, ny)
    cf, f, r = synth.make_case(cfg)
    m = driver.MckppPhysics(cf, f, numerics=0, devices=parts)
    g = m.gpu
    g.set_pass_budget(budget)
    g.set_async_stragglers(async_on)
    synth.apply_forcing(cfg, cf, f, r, 1)
    m.push_inputs(); m.mckpp_initialize_ocean_model()
    g.reserve_forcing_slots(nst)
    for nt in range(1, nst + 1):
        g.upload_forcing_slot(nt - 1, synth.apply_forcing(cfg, cf, f, r, nt))
    ids = list(range(capi.out_ids()["KPP_OUT_R_UVEL"]))
    g.output_ring_create(ids, depth=2)
    for nt in range(1, nst + 1):
        g.select_forcing_slot(nt - 1)
        g.step(nt)                      # queued: no sync in between when the stragglers are asynchronous
    rep = g.sync()
    slot = g.output_ring_submit()
    blocks = g.output_ring_wait(slot)
    m.pull(driver.ALL_OUTPUTS)
    out = m.mckpp_xios_diagnostic_output()
    assert np.array_equal(blocks[3], out["S"])
    print(name, cfg.npts, "columns, budget", budget, "async" if async_on else "sync", "parts", len(g.parts()),
          "max_iter", rep.max_iter, "hmix", float(f["hmix"].mean()), "S", float(out["S"].mean()), flush=True)
    g.output_ring_destroy()
    m.close()
