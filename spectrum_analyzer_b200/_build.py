"""In-tree build of libspectrum_analyzer_cuda.so with nvcc for sm_100a.

`build()` is what ``__graft_entry__.build()`` calls; it cross-compiles without a GPU.
The library lands in ``spectrum_analyzer_b200/lib/`` (git-ignored, travels with gpurun).
"""
from __future__ import annotations

import os
import shutil
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_PKG, "csrc")
LIB_DIR = os.path.join(_PKG, "lib")
# SA_LIB_PATH: load another build of the same library (timing experiments, profiles/microbench/)
LIB_PATH = os.environ.get("SA_LIB_PATH") or os.path.join(LIB_DIR, "libspectrum_analyzer_cuda.so")
HEADER = os.path.join(os.path.dirname(_PKG), "include", "spectrum_analyzer_cuda.h")

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC,-ffp-contract=off", "-diag-suppress", "177", "-shared",
]


def _sources() -> list[str]:
    return sorted(os.path.join(_CSRC, f) for f in os.listdir(_CSRC))


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libspectrum_analyzer_cuda.so cannot be built")


def is_stale() -> bool:
    if os.environ.get("SA_LIB_PATH"):
        # an explicitly chosen build (experiment variant) is never rebuilt over: it either exists or the load fails
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"SA_LIB_PATH={LIB_PATH} does not exist")
        return False
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(s) > t for s in _sources() + [HEADER])


def build(force: bool = False, verbose: bool = False, extra_flags: list[str] | None = None) -> str:
    """Compile csrc/sa_lib.cu -> lib/libspectrum_analyzer_cuda.so (sm_100a only)."""
    if not force and not is_stale():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [_nvcc(), *NVCC_FLAGS, *(extra_flags or []), "-o", LIB_PATH, os.path.join(_CSRC, "sa_lib.cu")]
    # the environment's CC/CXX may point at a wrapper; let nvcc pick the system g++
    env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
    res = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if verbose or res.returncode != 0:
        print(" ".join(cmd))
        print(res.stdout)
        print(res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libspectrum_analyzer_cuda.so")
    return LIB_PATH
