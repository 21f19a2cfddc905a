"""Compile the reference's OWN max-pool C++ into oracle/_ref/ (TEST INFRASTRUCTURE).

Source is compiled from where it lies (/root/reference/src/layers/CPU_kernels/
maxpooling_backend.cpp) -- nothing is copied into this repo.  Output:
oracle/_ref/_MaxPooling2DBackend_cpp<EXT_SUFFIX> (git-ignored, NOT gpurun-ignored,
so it travels to the GPU box like our own .so files).

The reference's Conv2D.cpp is NOT buildable here: it includes <Eigen/Core>, an
un-vendored third-party dependency with no pinned version (CmakeLists.txt:21,
.gitignore drops eigen/); see DESIGN.md.  Its two oneDNN files and the "GPU"
file do not compile at all.

OpenMP is deliberately NOT enabled: neither of the reference's CMake files adds
OpenMP flags, so the as-built kernels are single-threaded.
"""
from __future__ import annotations

import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/src/layers/CPU_kernels/maxpooling_backend.cpp"
OUT_DIR = os.path.join(HERE, "_ref")


def target_path() -> str:
    return os.path.join(OUT_DIR, "_MaxPooling2DBackend_cpp" + sysconfig.get_config_var("EXT_SUFFIX"))


def build(force: bool = False) -> str | None:
    """Build if the reference tree is present; return the .so path (or None)."""
    out = target_path()
    if not os.path.exists(REF_SRC):
        return out if os.path.exists(out) else None
    if os.path.exists(out) and not force and os.path.getmtime(out) >= os.path.getmtime(REF_SRC):
        return out
    import pybind11
    os.makedirs(OUT_DIR, exist_ok=True)
    # <omp.h> is included by the source but no OpenMP runtime call is made; the
    # pragmas are ignored without -fopenmp (as in the reference's own build).
    cmd = ["g++", "-O2", "-shared", "-fPIC", "-std=c++17", "-Wno-unknown-pragmas",
           "-I" + sysconfig.get_paths()["include"], "-I" + pybind11.get_include(),
           REF_SRC, "-o", out]
    subprocess.run(cmd, check=True)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
