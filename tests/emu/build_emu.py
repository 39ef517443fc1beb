"""Build tests/emu/_build/libpluto_gpu_emu.so (TEST INFRASTRUCTURE ONLY).

The product's CUDA sources are compiled UNCHANGED by g++ against the shim
tests/emu/include/cuda_runtime.h; the only source transformation is textual and done
here: ``kernel<<<grid, block, smem, stream>>>(args);`` becomes a call of the lock-step
interpreter (tests/emu/emu_runtime.cpp) and ``extern __shared__ T name[];`` a pointer to
the interpreter's per-block buffer.  EXACT objects are compiled without FMA contraction
(what -fmad=false gives on the GPU), FAST objects with it (-mfma -ffp-contract=fast) and
with PG_HOST_EMU (IEEE division / square root in place of the MUFU-seeded iterations).

The library lets the CPU-only test suite run the kernels' index logic and arithmetic
against the oracle before GPU time is spent.  It is not shipped, not imported by
pluto_b200, and never used for a measurement.
"""
from __future__ import annotations

import concurrent.futures
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "pluto_b200", "csrc")
BUILD = os.path.join(HERE, "_build")
LIB = os.path.join(BUILD, "libpluto_gpu_emu.so")

LAUNCH = re.compile(r"([A-Za-z_][\w:]*(?:<[^<>;{}]*?>)?)\s*<<<(.*?)>>>\s*\((.*?)\);", re.S)
DYN_SMEM = re.compile(r"extern\s+__shared__\s+(?:__align__\s*\(\s*\d+\s*\)\s+)?(\w+)\s+(\w+)\s*\[\s*\]\s*;")


def transform(text: str) -> str:
    text = LAUNCH.sub(lambda m: f"pg_emu::launch ([&]{{ {m.group(1)} ({m.group(3)}); }}, {m.group(2)});", text)
    return DYN_SMEM.sub(lambda m: f"{m.group(1)} *{m.group(2)} = ({m.group(1)} *)pg_emu::dyn_smem ();", text)


def newer(src_files, target):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in src_files)


def build(verbose: bool = False) -> str:
    """Build (or reuse) the interpreter library; concurrent callers (pytest-xdist workers) take turns on a lock file."""
    import fcntl
    os.makedirs(BUILD, exist_ok=True)
    with open(os.path.join(BUILD, ".lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            return _build(verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build(verbose: bool = False) -> str:
    src_dir = os.path.join(BUILD, "src")
    os.makedirs(src_dir, exist_ok=True)
    sources = sorted(f for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh")))
    deps = [os.path.join(CSRC, f) for f in sources] + [os.path.join(HERE, "emu_runtime.cpp"),
                                                         os.path.join(HERE, "include", "cuda_runtime.h"),
                                                         os.path.join(ROOT, "include", "pluto_gpu.h"), __file__]
    if not newer(deps, LIB):
        return LIB
    for f in sources:
        with open(os.path.join(CSRC, f)) as fh:
            text = transform(fh.read())
        out = os.path.join(src_dir, f)
        if not os.path.exists(out) or open(out).read() != text:
            with open(out, "w") as fh:
                fh.write(text)
    # "../../include" from _build/src is tests/emu/include: give it the header as well
    link2 = os.path.join(HERE, "include", "pluto_gpu.h")
    if os.path.lexists(link2):
        os.remove(link2)
    os.symlink(os.path.join(ROOT, "include", "pluto_gpu.h"), link2)

    common = ["g++", "-x", "c++", "-std=c++17", "-O2", "-fPIC", "-w", "-I", os.path.join(HERE, "include"), "-I", src_dir]
    exact = common + ["-ffp-contract=off", "-DPG_NS=pg_exact"]
    fast = common + ["-mfma", "-ffp-contract=fast", "-DPG_NS=pg_fast", "-DPG_FAST=1", "-DPG_HOST_EMU=1"]
    jobs = [(common + ["-ffp-contract=off"], "pluto_gpu.cu", "pluto_gpu.o"),
            (exact, "ct_kernels.cu", "ct_exact.o"), (fast, "ct_kernels.cu", "ct_fast.o")]
    # the Roe units (solver 2) of the FAST library keep IEEE arithmetic (pluto_b200/csrc/Makefile: -fmad=false)
    fast_roe = common + ["-ffp-contract=off", "-DPG_NS=pg_fast", "-DPG_FAST=1", "-DPG_HOST_EMU=1"]
    for s in range(5):
        fs = fast_roe if s == 2 else fast
        jobs.append((exact + [f"-DPG_SOLVER={s}"], "sweep_inst.cu", f"sweep_exact_{s}.o"))
        jobs.append((fs + [f"-DPG_SOLVER={s}"], "sweep_inst.cu", f"sweep_fast_{s}.o"))
        if os.path.exists(os.path.join(CSRC, "ctu_inst.cu")):
            jobs.append((exact + [f"-DPG_SOLVER={s}"], "ctu_inst.cu", f"ctu_exact_{s}.o"))
            jobs.append((fs + [f"-DPG_SOLVER={s}"], "ctu_inst.cu", f"ctu_fast_{s}.o"))
    objs = []

    def cc(job):
        flags, src, obj = job
        o = os.path.join(BUILD, obj)
        cmd = flags + ["-c", os.path.join(src_dir, src), "-o", o]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"{' '.join(cmd)}\n{r.stderr[-6000:]}")
        return o

    with concurrent.futures.ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        objs = list(ex.map(cc, jobs))
    rt = os.path.join(BUILD, "emu_runtime.o")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-c", os.path.join(HERE, "emu_runtime.cpp"), "-o", rt])
    subprocess.check_call(["g++", "-shared", "-o", LIB] + objs + [rt])
    if verbose:
        print("built", LIB)
    return LIB


if __name__ == "__main__":
    build(verbose=True)
    sys.exit(0)
