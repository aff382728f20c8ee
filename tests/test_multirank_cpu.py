"""world_size-2 gloo tests (CPU) of the N>1 path: the columns of the domain are
block-partitioned over ranks with no data-path collective; torch.distributed only
brackets the timed region and reduces the timing (MAX).  The per-rank compute is the
oracle here (no GPU in this container); on the GPU box tests/test_gpu_parity.py checks
the same partition invariance through the CUDA library."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, nsteps, out_dir):
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        sys.path.insert(0, p)
    import oracle_lib
    from mckpp_f90_b200 import synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    base = synth.scaled(synth.CONFIGS["cfg2"], 6, 4)
    cfg = synth.scaled(base, base.nx, base.ny * world)          # weak scaling: bench.py's partition
    ncols = base.npts
    cf, f, r = synth.make_case(cfg, col_offset=rank * ncols, ncols=ncols)
    orc = oracle_lib.Oracle(cf, f, nthreads=1)
    synth.apply_forcing(cfg, cf, f, r, 1)
    orc.initialize_ocean_model()
    dist.barrier()
    for nt in range(1, nsteps + 1):
        synth.apply_forcing(cfg, cf, f, r, nt)
        orc.physics_driver(nt)
    dist.barrier()
    # timing reduction as in bench.py: MAX over ranks
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert t.item() == float(world)
    # gather of an output field to rank 0 (diagnostics only; never on the step)
    x = torch.from_numpy(np.ascontiguousarray(f["X"][:, :, 0]))
    parts = [torch.empty_like(x) for _ in range(world)] if rank == 0 else None
    dist.gather(x, parts, dst=0)
    if rank == 0:
        np.save(os.path.join(out_dir, "gathered.npy"), torch.cat(parts, 0).numpy())
    dist.destroy_process_group()


def test_two_rank_partition_equals_single_domain(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_lib
    from mckpp_f90_b200 import synth
    oracle_lib.build()
    world, nsteps = 2, 3
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, nsteps, str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "gathered.npy")
    base = synth.scaled(synth.CONFIGS["cfg2"], 6, 4)
    cfg = synth.scaled(base, base.nx, base.ny * world)
    cf, f, r = synth.make_case(cfg)
    orc = oracle_lib.Oracle(cf, f, nthreads=2)
    synth.apply_forcing(cfg, cf, f, r, 1)
    orc.initialize_ocean_model()
    for nt in range(1, nsteps + 1):
        synth.apply_forcing(cfg, cf, f, r, nt)
        orc.physics_driver(nt)
    # bit-identical: partitions are independent
    assert np.array_equal(got, f["X"][:, :, 0])
