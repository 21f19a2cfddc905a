"""neuromah_b200 -- B200-native training hot path behind Neuromah's layer/optimizer/model API.

Import paths mirror the reference's (``Neuromah.src.layers`` -> ``neuromah_b200.src.layers``);
``install_alias()`` additionally registers the package under the name ``Neuromah`` so the
reference's example scripts run unchanged.
"""
import sys as _sys

__version__ = "0.1.0"


def install_alias(name: str = "Neuromah"):
    """Make ``import Neuromah.src...`` resolve to this package (examples/test_*_MNIST.py:2-9) at any depth: every module
    of the package is registered under the alias name as THE SAME module object, so that a deep import like the reference
    notebook's ``from Neuromah.src.losses.categorical_crossentropy import Loss_CategoricalCrossentropy``
    (examples/test.ipynb cell 1) yields the class Model.finalize tests with ``isinstance`` -- not a second execution of
    the file under the alias name, with its own copy of the library stream and launch state."""
    import importlib
    import pkgutil
    me = _sys.modules[__name__]
    _sys.modules.setdefault(name, me)
    for info in pkgutil.walk_packages(me.__path__, prefix=__name__ + "."):
        if info.name.startswith((__name__ + ".compat", __name__ + ".csrc")):
            continue                      # the reference-side shim is not part of the mirror
        spec = info.module_finder.find_spec(info.name)
        if spec is None or not (info.ispkg or str(spec.origin).endswith(".py")):
            continue                      # libneuromah_b200*.so sit in the package directory: not Python modules
        mod = importlib.import_module(info.name)
        _sys.modules.setdefault(name + info.name[len(__name__):], mod)
    return me
