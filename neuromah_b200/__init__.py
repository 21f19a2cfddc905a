"""neuromah_b200 -- B200-native training hot path behind Neuromah's layer/optimizer/model API.

Import paths mirror the reference's (``Neuromah.src.layers`` -> ``neuromah_b200.src.layers``);
``install_alias()`` additionally registers the package under the name ``Neuromah`` so the
reference's example scripts run unchanged.
"""
import sys as _sys

__version__ = "0.1.0"


def install_alias(name: str = "Neuromah"):
    """Make ``import Neuromah.src...`` resolve to this package (examples/test_*_MNIST.py:2-9)."""
    import importlib
    me = _sys.modules[__name__]
    _sys.modules.setdefault(name, me)
    for sub in ("src", "src.layers", "src.activations", "src.losses", "src.optimizers", "src.core",
                "src.metrics", "src.initializers", "src.utils", "Federated"):
        mod = importlib.import_module(f"{__name__}.{sub}")
        _sys.modules.setdefault(f"{name}.{sub}", mod)
    return me
