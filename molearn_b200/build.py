"""In-tree build of ``libmolearn_b200.so`` (nvcc, sm_100a) and staging of the Amber parameter files.

``python -m molearn_b200.build`` or ``__graft_entry__.build()``.  The shared library is written to
``molearn_b200/_C/`` (git-ignored, shipped to the GPU box with the working tree).

The Amber ff14SB text files (``amino12.lib``, ``parm10.dat``, ``frcmod.ff14SB``; AmberTools, GPL)
are data of the reference installation and are *read at run time*; they are not part of this
repository's history.  ``stage_params()`` copies them from a molearn checkout / install
(``$MOLEARN_REF`` or ``$MOLEARN_PARAMS``) into ``molearn_b200/_params/`` (git-ignored) so that
machines without the reference tree (the GPU box) can build topologies.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "_C")
LIB = os.path.join(OUT_DIR, "libmolearn_b200.so")
PARAMS = os.path.join(HERE, "_params")
PARAM_FILES = ("amino12.lib", "parm10.dat", "frcmod.ff14SB")
SOURCES = ("energy.cu", "topology.cpp")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA extension cannot be built (there is no CPU fallback)")


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "molearn_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    if force or _stale():
        extra = os.environ.get("MLP_NVCC_DEFS", "").split()          # e.g. "-DNB_MINB=4" when tuning
        cmd = [_nvcc(), *NVCC_FLAGS, *extra, "-o", LIB, *[os.path.join(CSRC, s) for s in SOURCES]]
        if verbose:
            cmd.insert(1, "-Xptxas")
            cmd.insert(2, "-v")
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    return LIB


def find_reference_params():
    cands = []
    if os.environ.get("MOLEARN_PARAMS"):
        cands.append(os.environ["MOLEARN_PARAMS"])
    ref = os.environ.get("MOLEARN_REF", "/root/reference")
    cands.append(os.path.join(ref, "src", "molearn", "loss_functions", "parameters"))
    try:  # an installed molearn
        import importlib.util
        spec = importlib.util.find_spec("molearn")
        if spec and spec.submodule_search_locations:
            for loc in spec.submodule_search_locations:
                cands.append(os.path.join(loc, "loss_functions", "parameters"))
                cands.append(os.path.join(loc, "parameters"))
    except Exception:
        pass
    for c in cands:
        if all(os.path.isfile(os.path.join(c, f)) for f in PARAM_FILES):
            return c
    return None


def stage_params() -> str:
    """Make ``molearn_b200/_params`` hold the three Amber files; returns the directory."""
    if all(os.path.isfile(os.path.join(PARAMS, f)) for f in PARAM_FILES):
        return PARAMS
    src = find_reference_params()
    if src is None:
        raise FileNotFoundError(
            "Amber parameter files (amino12.lib, parm10.dat, frcmod.ff14SB) not found: set $MOLEARN_PARAMS "
            "to molearn's loss_functions/parameters directory")
    os.makedirs(PARAMS, exist_ok=True)
    for f in PARAM_FILES:
        shutil.copyfile(os.path.join(src, f), os.path.join(PARAMS, f))
    return PARAMS


def param_dir() -> str:
    if os.environ.get("MOLEARN_PARAMS"):
        return os.environ["MOLEARN_PARAMS"]
    return stage_params()


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(stage_params())
