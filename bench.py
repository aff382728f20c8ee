#!/usr/bin/env python
"""bench.py -- ocean column-timesteps/s of the MC-KPP column-physics step on B200.

Contract (see the task statement):  python bench.py --gpus N --steps K --warmup W
prints ONE JSON line on rank 0.  A "step" is one call of the hot path
(mckpp_physics_driver: every ocean column advanced by one timestep) on synthetic forcing.

Workload (default --config cfg4): BASELINE.json configs[3], the global 0.25-degree grid --
700,000 ocean columns, NZ=100, double diffusion on, dto=1200 s, diurnal forcing -- the largest
configuration that fits one B200 (36 GB of 180) and the grid the north-star scales on.  The columns
are block-partitioned over the N GPUs (STRONG scaling: the total is fixed, every rank owns
ceil(700000/N) columns); columns never communicate, so there is no collective on the step --
torch.distributed (NCCL) only brackets the timed region and takes the max over ranks.
--config cfg2 (regional 300x200, round 1's workload), cfg3, cfg5 select the other shapes;
--scaling weak gives every rank a full copy of the configuration instead.
Launched directly with --gpus N > 1 (no torchrun) the N devices are driven by ONE process through
the library's own multi-GPU handle (kpp_gpu_create_multi) -- the reference's single-rank host.

  value      SUSTAINED device-resident throughput: K steps timed after --spinup steps (default 80 =
             into model day 2, where some columns stop converging and iterate to itermax = 200, as they
             do in the reference).  Forcing of every step is already in HBM.
  honeymoon  the same K steps timed right after the W warm-up steps from rest (every column converges
             in the minimum of 6 passes): round 1's headline, kept for comparison.
  e2e        the metric through the host API with HOST buffers: per step sflux(:,1:6,5,0) host->device
             from pinned memory, the step, the 13 per-column outputs of 1dto3d device->host.
  e2e_output / e2e_output_every_72
             the same plus the 34 blocks mckpp_xios_diagnostic_output sends (state, fluxes, diffusivities,
             ...: the whole set an UNCHANGED host reads after every step) through the asynchronous output
             ring, every step / once per model day (72 steps).
  roofline   algorithmic state bytes per column-step / kernel time vs the measured HBM
             bandwidth (MEASURED_PEAKS.json); traffic and fp64 pipe share from the ncu capture of THIS
             kernel source (profiles/traffic.json is stamped with the source hash; stale -> null).
  cpu_baseline  the CPU oracle (C restatement of the reference's OpenMP column loop; the Fortran itself
             cannot be built in this image) on a bounded sample, plus a second row with the reference's
             per-column gather/scatter + 34 allocations (mckpp_types_transfer.F90).

--impl reference  times that CPU oracle alone with all host threads (kind = "port").
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np

PERIOD = 72      # forcing level B is diurnal: 86400 s / dto = 72 steps


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--spinup", type=int, default=80, help="untimed steps before the sustained region (0: none)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="cfg4", help="cfg1..cfg5 (BASELINE.json configs)")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--numerics", type=int, default=int(os.environ.get("KPP_NUMERICS", "0")))
    ap.add_argument("--cpu-sample-cols", type=int, default=20000)
    ap.add_argument("--cpu-sample-steps", type=int, default=40)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-output-e2e", action="store_true", help="skip the e2e_output legs")
    ap.add_argument("--no-other-shapes", action="store_true", help="skip the short cfg2/cfg3/cfg5 measurements (1 GPU only)")
    return ap.parse_args()


def algorithmic_bytes(nz: int) -> int:
    """read U,X (4p) + Us,Xs both levels (8p); write U,X (4p) + Us,Xs(new) (4p); scalars 196 B."""
    return 20 * (nz + 1) * 8 + 196


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def kernel_source_sha() -> str:
    h = hashlib.sha256()
    for f in ("kpp_kernels.cu", "kpp_dev.h"):
        h.update(open(os.path.join(ROOT, "mckpp_f90_b200", "csrc", f), "rb").read())
    return h.hexdigest()[:16]


def profiled_counters(workload: str, cols_gpu: int = 0):
    """ncu counters of the dominant kernel, only if they were captured from THIS kernel source on this
    workload (tools/ncu_extract.py stamps profiles/traffic.json).  Several captures of one configuration at
    different domain sizes may exist (workload labels sharing the "cfgN" prefix): the one closest to the number of
    columns a launch covers here is taken."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        d = json.load(open(p))
    except Exception:
        return None, "no ncu capture committed"
    tag = workload.split(":")[0].split(" ")[0]
    cands = [e for e in d.get("captures", []) if e.get("kernel_sha") == kernel_source_sha()
             and (e.get("workload") == workload or str(e.get("workload", "")).split(":")[0].split(" ")[0] == tag)]
    if not cands:
        return None, "profiles/traffic.json holds no capture of this kernel source on this workload (stale): null"
    e = min(cands, key=lambda x: abs(np.log(max(int(x.get("columns", 1)), 1) / float(cols_gpu))) if cols_gpu else (x.get("workload") != workload))
    return e, f"ncu --set full, {e.get('source')}"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []
        self.first = 0

    def mark(self):
        """the timed region starts here: only samples taken from now on count"""
        self.first = len(self.lines)

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines[self.first:]:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 8:
                continue
            try:
                sm.append(float(p[0])); mx.append(float(p[1]))
            except ValueError:
                continue
            for nm, val in zip(names, p[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def workload_config(args, base, world, ncols_rank, single_process):
    """The `config` object: identical for the GPU arm and the reference arm."""
    return {"workload": base.name, "columns_total": base.npts if args.scaling == "strong" else base.npts * world,
            "nz": base.nz, "dto_s": base.dto, "LDD": bool(base.LDD),
            "forcing": "B (diurnal, per-column random amplitudes)" if base.forcing == "B" else "A (reference built-in constants)",
            "numerics": "strict" if args.numerics == 0 else "fast",
            "scaling": args.scaling}


def cpu_oracle_throughput(cfg, ncols, nsteps, nthreads, warm=1, realloc_1d=False, first_step=2):
    """Times the CPU oracle on a strided sample of the workload's columns: `warm` untimed steps, then `nsteps`."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_lib
    from mckpp_f90_b200 import synth
    stride = max(1, cfg.npts // ncols)
    gidx = np.arange(0, cfg.npts, stride)[:ncols]
    cf, f, r = synth.make_case(cfg, gidx=gidx)
    orc = oracle_lib.Oracle(cf, f, nthreads=nthreads)
    synth.apply_forcing(cfg, cf, f, r, 1)
    orc.initialize_ocean_model()
    nt = 0
    for _ in range(max(1, warm)):      # untimed (page-in, ntime<=1 table fills)
        nt += 1
        synth.apply_forcing(cfg, cf, f, r, nt)
        orc.physics_driver(nt, realloc_1d=realloc_1d)
    t0 = time.perf_counter()
    for _ in range(nsteps):
        nt += 1
        synth.apply_forcing(cfg, cf, f, r, nt)
        orc.physics_driver(nt, realloc_1d=realloc_1d)
    dt = time.perf_counter() - t0
    niter = float(orc.diag["iter"].mean())
    return gidx.size * nsteps / dt, dt, gidx.size, niter


def other_shape(name, device, numerics, steps=24, warm=4):
    """Short device-resident measurement of another BASELINE shape, per-step sync: column-steps/s over the first
    steps from rest (every column converges in ~6 passes) and over the later ones (config 5 hands hundreds of
    non-converging columns per step to the cooperative kernel from step 11 on), and the contract's roofline fraction."""
    from mckpp_f90_b200 import synth, driver
    cfg = synth.CONFIGS[name]
    cf, f, r = synth.make_case(cfg)
    m = driver.MckppPhysics(cf, f, device=device, numerics=numerics)
    g = m.gpu
    synth.apply_forcing(cfg, cf, f, r, 1)
    m.push_inputs()
    m.mckpp_initialize_ocean_model()
    g.reserve_forcing_slots(warm + steps)
    for i in range(warm + steps):
        g.upload_forcing_slot(i, synth.apply_forcing(cfg, cf, f, r, i + 1))
    ms, mx = [], []
    for i in range(warm + steps):
        g.select_forcing_slot(i)
        g.step(i + 1)
        rep = g.sync()
        ms.append(rep.kernel_ms); mx.append(rep.max_iter)
    m.close()
    t = float(np.median(ms[warm:warm + 6])) * 1e-3
    t_late = float(np.mean(ms[warm + 8:])) * 1e-3
    peak, _ = measured_peaks()
    return {"workload": cfg.name, "columns": cfg.npts, "nz": cfg.nz, "value": cfg.npts / t, "unit": "column-steps/s",
            "kernel_ms_per_step": 1e3 * t, "roofline_frac": cfg.npts / t * algorithmic_bytes(cfg.nz) / 1e9 / peak,
            "later_steps": {"value": cfg.npts / t_late, "kernel_ms_per_step": 1e3 * t_late, "max_iter": int(max(mx[warm + 8:])),
                            "what": f"mean of steps {warm + 9}..{warm + steps}"},
            "what": f"median of steps {warm + 1}..{warm + 6} from rest, device-resident forcing, sync after every step"}


def run_reference(args, rank, world):
    """Reference arm: the reference's CPU implementation of the path on the host cores.
    The Fortran cannot be compiled in this image (no gfortran/MPI/netCDF/XIOS), so the literal C
    restatement under oracle/ stands in (kind = "port").  One "step" here is one timestep of a bounded
    strided SAMPLE of the workload's columns (the full 700,000-column step takes ~10 s on 16 cores);
    W warm-up steps, then exactly K timed steps; ms_per_step is the measured time of one sample step."""
    if rank != 0:
        return
    from mckpp_f90_b200 import synth
    base = synth.CONFIGS[args.config]
    cores = os.cpu_count() or 1
    ncols = min(args.cpu_sample_cols, base.npts)
    K, W = max(1, args.steps), max(1, args.warmup)
    val, dt, n, niter = cpu_oracle_throughput(base, ncols, K, cores, warm=W)
    line = {
        "impl": "reference", "metric": "ocean column-timesteps/sec", "value": val, "unit": "column-steps/s",
        "n_gpus": args.gpus, "steps": K, "warmup": W,
        "ms_per_step": 1e3 * dt / K, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, base, world, ncols, False),
        "ms_per_step_is": f"one timestep of the {n}-column sample (measured, not extrapolated); the whole "
                          f"{base.npts}-column step would take {1e3 * base.npts / val:.0f} ms at this rate",
        "cpu_baseline": {"value": val, "unit": "column-steps/s", "cores": cores, "kind": "port",
                         "sample": f"{n} strided columns of the workload x {K} steps after {W} warm-up steps ({dt:.1f} s), "
                                   f"mean iter {niter:.2f}; C restatement of the reference's OpenMP column loop "
                                   "(oracle/); the reference's own Fortran/MPI/XIOS build is not buildable in this image"},
        "e2e": {"value": val, "unit": "column-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from mckpp_f90_b200 import synth, driver, capi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    # one process, several devices: the library's own multi-GPU handle
    single_process = world == 1 and args.gpus > 1
    ngroups = args.gpus if single_process else world

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    base = synth.CONFIGS[args.config]
    if args.scaling == "strong":
        # contiguous blocks of ceil(npts/world) columns rounded up to whole 32-column tiles (what
        # kpp_gpu_create_multi does inside the library)
        cfg = base
        col0, ncols = synth.block_partition(base.npts, world, rank)
        total_cols = base.npts
    else:
        cfg = synth.scaled(base, base.nx, base.ny * world)
        ncols, col0 = base.npts, rank * base.npts
        total_cols = base.npts * world
    if ncols == 0:
        raise SystemExit("more ranks than 32-column tiles")
    cf, f, r = synth.make_case(cfg, col_offset=col0, ncols=ncols)
    K, W, S = args.steps, args.warmup, args.spinup
    model = driver.MckppPhysics(cf, f, device=local_rank, numerics=args.numerics,
                                ngpus=args.gpus if single_process else None)
    gpu = model.gpu
    synth.apply_forcing(cfg, cf, f, r, 1)
    model.push_inputs()
    model.mckpp_initialize_ocean_model()

    # ---- one model day of forcing (the diurnal forcing has a period of 72 steps): pinned host copies
    # and device-resident slots
    nper = PERIOD if cfg.forcing == "B" and abs(cfg.dto * PERIOD - 86400.0) < 1e-9 else None
    nslots = nper if nper else (W + S + 2 * K + 8)
    forc = capi.pinned_empty((nslots, 6, ncols), order="C")
    for i in range(nslots):
        forc[i] = synth.apply_forcing(cfg, cf, f, r, i + 1)
    gpu.reserve_forcing_slots(nslots)
    for i in range(nslots):
        gpu.upload_forcing_slot(i, forc[i])
    gpu.sync()
    slot_of = (lambda nt: (nt - 1) % nslots)

    nt = 0

    # Stragglers asynchronously (kpp_gpu_set_async_stragglers): the steps of a region are QUEUED on the
    # device-resident forcing and synced once; columns that stop converging are finished on side streams
    # while the next steps of the others run.  The e2e regions below sync (and so join) after every step.
    gpu.set_async_stragglers(True)

    def resident_steps(n, timed):
        """n queued steps with device-resident forcing, one sync at the end; returns (device ms of the whole
        queue by CUDA events on the step stream, sum_iter / handed-over / max_iter of the LAST step)."""
        nonlocal nt
        for _ in range(n):
            nt += 1
            gpu.select_forcing_slot(slot_of(nt))
            gpu.step(nt)
        rep = gpu.sync()
        return rep.kernel_ms, rep.sum_iter, rep.n_handed_over, rep.max_iter

    # ---- warm-up from rest
    resident_steps(W, False)

    # ---- timed region 0: "honeymoon" (every column converges in 6 passes)
    barrier()
    t0 = time.perf_counter()
    hm_kernel_ms, hm_iter, hm_ho, _ = resident_steps(K, True)
    barrier()
    wall_honeymoon = time.perf_counter() - t0

    # ---- spin-up into the sustained regime (untimed); nvidia-smi is started here so that it is already
    # emitting when the timed regions begin (a 60,000-column region is over in 70 ms)
    sampler = ClockSampler(local_rank)
    sampler.start()
    if S > nt:
        resident_steps(S - nt, False)
    spun = nt

    # ---- timed region 1: K steps, inputs resident in HBM (the reported value)
    sampler.mark()
    launches0 = gpu.launch_count()
    barrier()
    t0 = time.perf_counter()
    kernel_ms, sum_iter, handed_over, max_iter = resident_steps(K, True)
    barrier()
    wall_resident = time.perf_counter() - t0
    launches = gpu.launch_count() - launches0

    # ---- timed region 2: e2e through the host API (pinned host buffers, H2D + D2H per step)
    gpu.select_forcing_slot(-1)
    outs = {n: capi.pinned_empty(f[n].shape, f[n].dtype) for n in driver.SCALAR_OUTPUTS}
    h2d = 6 * ncols * 8
    d2h = sum(a.nbytes for a in outs.values())

    def e2e_steps(n, ring_every=0):
        nonlocal nt
        pending = None
        for i in range(n):
            nt += 1
            gpu.upload_forcing(forc[slot_of(nt)])
            gpu.step(nt)
            for nm, a in outs.items():
                gpu.download_async(nm, a)
            if ring_every and i % ring_every == 0:
                if pending is not None:
                    gpu.output_ring_wait(pending)       # the host has consumed the previous output step
                pending = gpu.output_ring_submit()
            gpu.sync()                 # one wait for the step and its 13 result copies
        if pending is not None:
            gpu.output_ring_wait(pending)

    barrier()
    t0 = time.perf_counter()
    e2e_steps(K)
    barrier()
    wall_e2e = time.perf_counter() - t0
    clocks = sampler.stop()        # sampled over the sustained and the e2e regions (both timed)

    # ---- timed regions 3, 4: the output set of an unchanged host through the asynchronous ring
    out_info = None
    wall_out_every, wall_out_day, k_out = 0.0, 0.0, 0
    if not args.no_output_e2e:
        try:
            ids = list(range(capi.out_ids()["KPP_OUT_R_UVEL"]))          # the 34 blocks of mckpp_xios_diagnostic_output
            slot_bytes = gpu.output_ring_create(ids, depth=2)
            k_out = max(2, min(K, 6))
            barrier()
            t0 = time.perf_counter()
            e2e_steps(k_out, ring_every=1)
            barrier()
            wall_out_every = time.perf_counter() - t0
            barrier()
            t0 = time.perf_counter()
            e2e_steps(PERIOD, ring_every=PERIOD)
            barrier()
            wall_out_day = time.perf_counter() - t0
            gpu.output_ring_destroy()
            out_info = {"blocks": len(ids), "d2h_bytes_per_output_step": int(slot_bytes)}
        except capi.KppError as e:
            out_info = {"error": str(e)}

    # ---- max over ranks (device-event kernel time and wall time)
    t = torch.tensor([kernel_ms * 1e-3, wall_resident, wall_e2e, wall_honeymoon, hm_kernel_ms * 1e-3, wall_out_every,
                      wall_out_day], dtype=torch.float64, device="cuda")
    cnt = torch.tensor([float(sum_iter), float(handed_over), float(max_iter)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        cmax = cnt.clone()
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        dist.all_reduce(cmax, op=dist.ReduceOp.MAX)
        cnt[2] = cmax[2]
    t_kernel, t_res, t_e2e, t_hm, t_hm_kernel, t_out_every, t_out_day = [float(x) for x in t.tolist()]
    sum_iter_all, handed_all, max_iter_all = [float(x) for x in cnt.tolist()]

    if rank == 0:
        peak, peak_src = measured_peaks()
        balg = algorithmic_bytes(cfg.nz)
        value = total_cols * K / t_res
        # per-GPU kernel-only rate of the slowest rank's block (dominant kernel = kpp_step_kernel)
        cols_gpu = ncols if not single_process else -(-ncols // args.gpus)
        kern_val = cols_gpu * K / t_kernel
        achieved = kern_val * balg / 1e9
        prof, prof_src = profiled_counters(base.name, cols_gpu)
        if prof and cols_gpu != prof.get("columns"):
            # the capture is of the whole domain on one GPU; a launch here covers cols_gpu columns
            prof = dict(prof)
            prof["bytes_per_launch"] = prof["bytes_per_column_step"] * cols_gpu
            prof_src += (f" (captured at {prof['columns']} columns per launch; scaled by the measured bytes per "
                         f"column-step to the {cols_gpu} columns of one GPU's launch; dram/fp64 fractions are the capture's)")
        conf = workload_config(args, base, world, ncols, single_process)
        conf.update({"columns_per_gpu": cols_gpu,
                     "spinup_steps": spun,
                     "l2": "state+scratch per GPU >> 126 MB L2 (inputs larger than L2)",
                     "parallelism": (f"columns block-partitioned over {ngroups} GPU(s), no collective on the step; "
                                     + ("one process, in-library multi-GPU handle" if single_process else
                                        "one process per GPU" if world > 1 else "one GPU"))})
        line = {
            "metric": "ocean column-timesteps/sec", "value": value, "unit": "column-steps/s",
            "n_gpus": ngroups, "steps": K, "warmup": W, "ms_per_step": 1e3 * t_res / K, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": conf,
            "value_is": f"sustained: {K} queued steps (one sync at the end, stragglers finished asynchronously) timed "
                        f"after {spun} steps of spin-up (model day {spun * cfg.dto / 86400.0:.2f})",
            "honeymoon": {"value": total_cols * K / t_hm, "unit": "column-steps/s", "ms_per_step": 1e3 * t_hm / K,
                          "kernel_ms_per_step": 1e3 * t_hm_kernel / K, "mean_iter_last_step": hm_iter / float(ncols),
                          "what": f"the same {K} steps timed right after {W} warm-up steps from rest (round 1's headline)"},
            "e2e": {"value": total_cols * K / t_e2e, "unit": "column-steps/s", "h2d_bytes_per_step": h2d * (1 if single_process else world),
                    "d2h_bytes_per_step": d2h * (1 if single_process else world),
                    "what": "per step: sflux(:,1:6,5,0) host->device from pinned memory, step, the 13 per-column "
                            "outputs of 1dto3d device->host"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": (prof or {}).get("bytes_per_launch"), "traffic_source": prof_src,
                         "fp64_frac": (prof or {}).get("fp64_pipe_frac"),
                         "dram_frac_of_measured_peak": (prof or {}).get("dram_frac"),
                         "peak_source": peak_src, "kernel": "kpp_step_kernel",
                         "algorithmic_bytes_per_column_step": balg,
                         "kernel_ms_per_step": 1e3 * t_kernel / K,
                         "kernel_column_steps_per_s_per_gpu": kern_val,
                         "note": "achieved = algorithmic state bytes (SURVEY 8d) x columns / the step's kernel time "
                                 "(CUDA events on the handle's stream, slowest GPU); the kernel itself moves several "
                                 "times that as per-pass scratch and is bound by that traffic (traffic / "
                                 "dram_frac_of_measured_peak), the fp64 pipe is the next limit (fp64_frac): "
                                 "DESIGN.md 4/7 and profiles/"},
            # the LAST timed step (the steps are queued and synced once): mean / max passes per column, and the
            # columns the step kernel handed to the cooperative kernel in that step
            "mean_iter": sum_iter_all / float(total_cols),
            "max_iter": int(max_iter_all),
            "handed_over_columns": int(handed_all),
        }
        if out_info is not None and "error" not in out_info:
            nb = out_info["d2h_bytes_per_output_step"] * (1 if single_process else world)
            line["e2e_output"] = {"value": total_cols * k_out / t_out_every, "unit": "column-steps/s", "steps": k_out,
                                  "h2d_bytes_per_step": line["e2e"]["h2d_bytes_per_step"],
                                  "d2h_bytes_per_step": line["e2e"]["d2h_bytes_per_step"] + nb,
                                  "what": "e2e plus, EVERY step, the 34 blocks of mckpp_xios_diagnostic_output (what an "
                                          "unchanged host reads after each step) through the asynchronous output ring "
                                          "(depth 2: the copy of step n overlaps the kernels of step n+1)"}
            line["e2e_output_every_72"] = {"value": total_cols * PERIOD / t_out_day, "unit": "column-steps/s", "steps": PERIOD,
                                           "d2h_bytes_per_output_step": nb,
                                           "what": "e2e plus the same 34 blocks once per model day (72 steps)"}
        elif out_info is not None:
            line["e2e_output"] = out_info
        if not args.no_cpu_baseline and ngroups == 1:
            cores = os.cpu_count() or 1
            val, dt, n, niter = cpu_oracle_throughput(base, min(args.cpu_sample_cols, base.npts), args.cpu_sample_steps, cores)
            line["cpu_baseline"] = {"value": val, "unit": "column-steps/s", "cores": cores, "kind": "port",
                                    "sample": f"{n} strided columns of the workload x {args.cpu_sample_steps} steps "
                                              f"({dt:.1f} s), mean iter {niter:.2f}; C restatement (oracle/) of the "
                                              "reference's OpenMP column loop"}
            v2, dt2, n2, _ = cpu_oracle_throughput(base, min(args.cpu_sample_cols, base.npts) // 2,
                                                   max(4, args.cpu_sample_steps // 4), cores, realloc_1d=True)
            line["cpu_baseline_realloc"] = {"value": v2, "unit": "column-steps/s", "cores": cores, "kind": "port",
                                            "sample": f"{n2} columns x {max(4, args.cpu_sample_steps // 4)} steps ({dt2:.1f} s) with "
                                                      "the reference's per-column 3dto1d/1dto3d gather/scatter AND its 34 "
                                                      "ALLOCATEs per column per step (mckpp_types_transfer.F90:15-327)"}
        if ngroups == 1 and not args.no_other_shapes:
            model.close()
            model = None
            line["other_shapes"] = {n: other_shape(n, local_rank, args.numerics) for n in ("cfg2", "cfg3", "cfg5") if n != args.config}
        print(json.dumps(line), flush=True)
    if model is not None:
        model.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
