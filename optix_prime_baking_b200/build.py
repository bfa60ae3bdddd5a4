"""In-tree build of libaobake.so (nvcc, sm_100a only) — no JIT cache, no torch extension:
the C-ABI library is what a reference maintainer would link."""
from __future__ import annotations

import os
import shutil
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libaobake.so")

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC,-ffp-contract=off", "-shared",
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libaobake.so cannot be built (there is no CPU fallback)")


def sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh"))] + \
           [os.path.join(ROOT, "include", "aobake.h")]


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(s) > t for s in sources())


def build(force: bool = False, verbose: bool = False, out: str | None = None, extra_flags=()) -> str:
    """Builds libaobake.so.  `out` + `extra_flags` build a variant elsewhere (kernel A/B runs load it
    through the AOBAKE_LIB environment variable; see profiles/ab_variants.py)."""
    if out is None and not force and not needs_build():
        return LIB_PATH
    out = out or LIB_PATH
    host_cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [_nvcc(), *NVCC_FLAGS, *extra_flags, "-ccbin", host_cxx, "-I", os.path.join(ROOT, "include"), "-I", CSRC,
           "-o", out, os.path.join(CSRC, "aobake.cu")]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return out


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
