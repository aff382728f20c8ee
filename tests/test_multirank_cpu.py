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
    cfg = synth.scaled(synth.CONFIGS["cfg4"], 11, 7)            # 77 columns: blocks of 64 and 13 (strong scaling,
    col0, ncols = synth.block_partition(cfg.npts, world, rank)  # bench.py's and kpp_gpu_create_multi's partition)
    cf, f, r = synth.make_case(cfg, col_offset=col0, ncols=ncols)
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
    parts = [None] * world if rank == 0 else None
    dist.gather_object((col0, np.ascontiguousarray(f["X"][:, :, 0])), parts, dst=0)
    if rank == 0:
        assert [p[0] for p in parts] == [0, 64] and [p[1].shape[0] for p in parts] == [64, 13]
        np.save(os.path.join(out_dir, "gathered.npy"), np.concatenate([p[1] for p in parts], 0))
    dist.destroy_process_group()


def test_block_partition_covers_every_column_once():
    from mckpp_f90_b200 import synth
    for npts in (1, 16, 77, 44000, 60000, 700000):
        for world in (1, 2, 3, 4, 8):
            blocks = [synth.block_partition(npts, world, r) for r in range(world)]
            assert sum(n for _, n in blocks) == npts
            pos = 0
            for c0, n in blocks:
                assert c0 == pos or n == 0
                assert c0 % 32 == 0 or n == 0
                pos += n


def test_two_rank_partition_equals_single_domain(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_lib
    from mckpp_f90_b200 import synth
    oracle_lib.build()
    world, nsteps = 2, 3
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, nsteps, str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "gathered.npy")
    cfg = synth.scaled(synth.CONFIGS["cfg4"], 11, 7)
    cf, f, r = synth.make_case(cfg)
    orc = oracle_lib.Oracle(cf, f, nthreads=2)
    synth.apply_forcing(cfg, cf, f, r, 1)
    orc.initialize_ocean_model()
    for nt in range(1, nsteps + 1):
        synth.apply_forcing(cfg, cf, f, r, nt)
        orc.physics_driver(nt)
    # bit-identical: partitions are independent
    assert np.array_equal(got, f["X"][:, :, 0])
