"""Load oracle/_ref/_MaxPooling2DBackend_cpp (the reference's own compiled C++), if built."""
from __future__ import annotations

import importlib.util
import os

from .build_ref import conv_target_path, target_path


def load_ref_maxpool():
    path = target_path()
    if not os.path.exists(path):
        return None
    spec = importlib.util.spec_from_file_location("_MaxPooling2DBackend_cpp", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_ref_conv():
    """The reference's Conv2D.cpp compiled against oracle/eigen_standin (see build_ref.py): a checker for the conv oracle,
    correct for one output channel only."""
    path = conv_target_path()
    if not os.path.exists(path):
        return None
    spec = importlib.util.spec_from_file_location("_Conv2DBackend_cpp", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod

