"""Compile the reference's OWN C++ kernels into oracle/_ref/ (TEST INFRASTRUCTURE).

Sources are compiled from where they lie under /root/reference/src/layers/CPU_kernels/ -- nothing is copied into this
repo.  Outputs (git-ignored, NOT gpurun-ignored, so they travel to the GPU box like our own .so files):

* oracle/_ref/_MaxPooling2DBackend_cpp<EXT_SUFFIX>  from maxpooling_backend.cpp: self-contained.
* oracle/_ref/_Conv2DBackend_cpp<EXT_SUFFIX>        from Conv2D.cpp.  That file includes <Eigen/Core>, an un-vendored
  third-party dependency with no pinned version (CmakeLists.txt:21, .gitignore drops eigen/), and this image has no
  Eigen: it is compiled against oracle/eigen_standin/Eigen/Core, a stand-in written for this repository that implements
  only the five operations the file uses with Eigen's documented semantics (column-major MatrixXf / Map, products,
  transpose, setZero, +=).  The reference's im2col / col2im / batching / pybind code runs unmodified.  As SURVEY.md Q6
  says the result is only right for ONE output channel (the kernel and the output are row-major NumPy memory mapped as
  column-major matrices, Conv2D.cpp:114, :137): tests/test_oracle_ref_conv.py pins the oracle's conv restatement
  against it channel by channel and documents the multi-channel defect.  It is a CHECKER of the oracle only -- never
  the CPU baseline's conv (wrong for OC > 1) and never on a product path.

The two oneDNN files and the "GPU" file of the reference do not compile at all.

OpenMP is deliberately NOT enabled: neither of the reference's CMake files adds OpenMP flags, so the as-built kernels are
single-threaded.
"""
from __future__ import annotations

import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = "/root/reference/src/layers/CPU_kernels"
REF_SRC = os.path.join(REF_DIR, "maxpooling_backend.cpp")
REF_CONV_SRC = os.path.join(REF_DIR, "Conv2D.cpp")
EIGEN_STANDIN = os.path.join(HERE, "eigen_standin")
OUT_DIR = os.path.join(HERE, "_ref")


def target_path() -> str:
    return os.path.join(OUT_DIR, "_MaxPooling2DBackend_cpp" + sysconfig.get_config_var("EXT_SUFFIX"))


def conv_target_path() -> str:
    return os.path.join(OUT_DIR, "_Conv2DBackend_cpp" + sysconfig.get_config_var("EXT_SUFFIX"))


def _compile(src, out, deps=(), includes=(), force=False):
    if not os.path.exists(src):
        return out if os.path.exists(out) else None
    newest = max(os.path.getmtime(p) for p in (src,) + tuple(deps))
    if os.path.exists(out) and not force and os.path.getmtime(out) >= newest:
        return out
    import pybind11
    os.makedirs(OUT_DIR, exist_ok=True)
    # <omp.h> is included by the sources but no OpenMP runtime call is made; the
    # pragmas are ignored without -fopenmp (as in the reference's own build).
    cmd = ["g++", "-O2", "-shared", "-fPIC", "-std=c++17", "-Wno-unknown-pragmas",
           "-I" + sysconfig.get_paths()["include"], "-I" + pybind11.get_include()] + ["-I" + i for i in includes] + [src, "-o", out]
    subprocess.run(cmd, check=True)
    return out


def build(force: bool = False) -> str | None:
    """Build both modules if the reference tree is present; return the max-pool .so path (or None)."""
    _compile(REF_CONV_SRC, conv_target_path(), deps=(os.path.join(EIGEN_STANDIN, "Eigen", "Core"),),
             includes=(EIGEN_STANDIN,), force=force)
    return _compile(REF_SRC, target_path(), force=force)


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
    print(conv_target_path() if os.path.exists(conv_target_path()) else None)
