"""Builds libb200lap.so (the only native artefact of the product) with nvcc for sm_100a.

    python build.py            # rebuild when a source is newer than the library
    python build.py --force

Flags: -gencode arch=compute_100a,code=sm_100a (B200 only, no other arch, no PTX fallback),
-lineinfo (ncu source pages), -fmad=false (the reference solver is built without FMA contraction;
the MLP asks for its FMAs explicitly with fmaf()).  The library links cudart statically and has no
torch dependency.
"""
from __future__ import annotations

import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "b200lap", "libb200lap.so")
OUT_PROF = os.path.join(HERE, "b200lap", "libb200lap_prof.so")   # same sources + cycle counters in the solver trace (tools/ only)

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-std=c++17", "-O3", "-lineinfo", "-fmad=false",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O2", "-shared",
    "-diag-suppress", "177",
]


def nvcc_path() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libb200lap.so can only be built with the CUDA toolkit")


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.cpp")) +
                  glob.glob(os.path.join(CSRC, "*.inc")) + [os.path.join(HERE, "..", "include", "b200lap.h")])


def stale(out: str = OUT) -> bool:
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(s) > t for s in sources())


def build(force: bool = False, verbose: bool = False, profile: bool = False) -> str:
    """profile=True builds the measurement variant libb200lap_prof.so (-DB200LAP_SOLVER_PROFILE: SM-cycle counters
    in the solver trace); the tools load it with B200LAP_PROFILE_LIB=1.  The product library never carries them."""
    out = OUT_PROF if profile else OUT
    if not force and not stale(out):
        return out
    prof = ["-DB200LAP_SOLVER_PROFILE"] if profile else []
    if os.environ.get("B200LAP_KMAX"):       # experiment builds (tools/): scans per batched relax step
        prof += ["-DB200LAP_KMAX=" + os.environ["B200LAP_KMAX"]]
        out = out.replace(".so", "_k" + os.environ["B200LAP_KMAX"] + ".so")
    cmd = [nvcc_path()] + NVCC_FLAGS + prof + (["-Xptxas", "-v"] if verbose else []) + ["-o", out, os.path.join(CSRC, "api.cu"), os.path.join(CSRC, "host_narrow.cpp")]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, profile="--profile" in sys.argv))
