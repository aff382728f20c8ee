"""Builds libkpp_gpu.so (sm_100a) in-tree with nvcc.

The kernel translation unit is compiled twice -- strict (``-fmad=false``) and
fast (``-fmad=true``) numerics -- and linked with the C-ABI layer.  No torch,
no JIT cache: the .so lives next to the sources so it travels with the repo.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# KPP_BUILD_TAG=x builds libkpp_gpu_x.so from build_x/ (kernel A/B experiments: select with KPP_LIB_PATH)
_TAG = os.environ.get("KPP_BUILD_TAG", "")
LIB = os.path.join(HERE, f"libkpp_gpu_{_TAG}.so" if _TAG else "libkpp_gpu.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-O2", "-Xcompiler", "-ffp-contract=off"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _sources():
    return [os.path.join(CSRC, f) for f in ("kpp_kernels.cu", "kpp_api.cu", "kpp_dev.h")] + [
        os.path.join(HERE, "..", "include", "kpp_gpu.h"), os.path.join(HERE, "gen_exp_table.py"),
        os.path.join(HERE, "host", "kpp_host_demo.cpp"), os.path.join(HERE, "host", "mckpp_host.hpp")]


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in _sources())


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    nvcc = _nvcc()
    env = dict(os.environ)
    # strict variant: reproduce the host libm's exp() bit for bit (table read from libm, validated)
    from . import gen_exp_table
    gen_exp_table.generate(verbose=verbose)
    # the image exports CC/CXX pointing at a gcc without OpenMP specs; nvcc only needs a host g++
    ccbin = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else None
    base = [nvcc] + ARCH + COMMON
    if ccbin:
        base += ["-ccbin", ccbin]
    if verbose:
        base += ["-Xptxas", "-v"]
    for var in ("KPP_STEP_MIN_BLOCKS", "KPP_STEP_BLOCK", "KPP_PIPE_D", "KPP_COOP_PROF", "KPP_DEEP_ROOMY", "KPP_SHARE_RCP", "KPP_EOS_SHARE_RCP", "KPP_EXP_A", "KPP_EXP_B"):
        if os.environ.get(var):
            base += [f"-D{var}=" + os.environ[var]]
    objs = []
    jobs = [
        ("kpp_kernels_strict.o", "kpp_kernels.cu", ["-DKPP_VARIANT_STRICT", "-fmad=false", "-prec-div=true", "-prec-sqrt=true"]),
        ("kpp_kernels_fast.o", "kpp_kernels.cu", ["-DKPP_VARIANT_FAST", "-fmad=true", "-prec-div=true", "-prec-sqrt=true"]),
        ("kpp_api.o", "kpp_api.cu", []),
    ]
    bdir = os.path.join(HERE, f"build_{_TAG}" if _TAG else "build")
    os.makedirs(bdir, exist_ok=True)
    procs = []
    for obj, src, extra in jobs:
        out = os.path.join(bdir, obj)
        cmd = base + extra + ["-c", os.path.join(CSRC, src), "-o", out]
        procs.append((cmd, subprocess.Popen(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(out)
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    link = [nvcc] + ARCH + ["-shared", "-o", LIB] + objs + (["-ccbin", ccbin] if ccbin else [])
    subprocess.check_call(link, env=env)
    if not _TAG:
        build_host_demo()
    return LIB


def build_host_demo() -> str:
    """The compiled (C++) host above the C ABI: host/kpp_host_demo."""
    exe = os.path.join(HERE, "host", "kpp_host_demo")
    cmd = ["/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++", "-O2", "-std=c++17",
           os.path.join(HERE, "host", "kpp_host_demo.cpp"), "-o", exe, "-L", HERE, "-lkpp_gpu",
           "-Wl,-rpath," + HERE, "-Wl,-rpath,$ORIGIN/.."]
    subprocess.check_call(cmd)
    return exe


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
