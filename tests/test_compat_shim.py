"""INTEGRATION.md section 2 as CODE: neuromah_b200/compat/compiled/__init__.py is the file a maintainer copies to the
reference's ``src/layers/compiled/`` seam (Conv_wrapepr.py:3, :91; MaxPooling2D.py:2, :53-80).

* CPU (authoring container, where /root/reference exists): the UNMODIFIED reference ``Layer_Conv2D`` /
  ``Layer_MaxPooling2D`` import against the shim and bind its four functions (no compute without a GPU).
* GPU: the shim's functions through the reference's own max-pool KATs (tests/layers/test_maxpool2D.py:14-57), the conv
  shape cases of tests/layers/test_Conv2dBackend.py:6-43, and against the oracle on the wrapper's call pattern
  (np.pad outside, Conv_wrapepr.py:78-84).  The GPU box has no /root/reference, so the wrappers' few lines are
  restated in the test rather than imported.
"""
import os
import sys
import tempfile
import types

import numpy as np
import pytest

from oracle import np_ref as R

REF_ROOT = "/root/reference"


def _shim():
    from neuromah_b200.compat import compiled
    return compiled


def test_shim_exports_the_seam_the_reference_imports():
    c = _shim()
    assert callable(c._Conv2DBackend_cpp.conv2d_cpu) and callable(c._Conv2DBackend_cpp.conv2d_backward_cpu)
    assert callable(c._MaxPooling2DBackend_cpp.maxpool2d_forward_cpu)
    assert callable(c._MaxPooling2DBackend_cpp.maxpool2d_backward_cpu)
    # argument validation happens before any device call (Conv2D.cpp:86-87 -> RuntimeError)
    with pytest.raises(RuntimeError):
        c.conv2d_cpu(np.zeros((3, 3)), np.zeros((1, 1, 3, 3)))
    with pytest.raises(RuntimeError):
        c.conv2d_cpu(np.zeros((1, 2, 5, 5)), np.zeros((4, 3, 3, 3)))


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF_ROOT, "src")), reason="reference tree only in the authoring container")
def test_unmodified_reference_layers_bind_the_shim():
    pkg = "NeuromahShimmed"
    tmp = tempfile.mkdtemp(prefix="nmshim_")
    os.symlink(REF_ROOT, os.path.join(tmp, pkg))
    sys.path.insert(0, tmp)
    try:
        shim = _shim()
        sys.modules[f"{pkg}.src.layers.compiled"] = shim           # what copying the file to src/layers/compiled/ does
        rnn = types.ModuleType(f"{pkg}.src.layers.RNN")            # the reference imports a file it does not ship
        rnn.Layer_RNN = type("Layer_RNN", (), {})
        sys.modules[rnn.__name__] = rnn
        import importlib
        layers = importlib.import_module(f"{pkg}.src.layers")
        conv_mod = importlib.import_module(f"{pkg}.src.layers.Conv_wrapepr")
        pool_mod = importlib.import_module(f"{pkg}.src.layers.MaxPooling2D")
        assert conv_mod._Conv2DBackend_cpp is shim._Conv2DBackend_cpp
        assert pool_mod._MaxPooling2DBackend_cpp is shim._MaxPooling2DBackend_cpp
        conv = layers.Layer_Conv2D(1, 4, (3, 3), padding="same")
        pool = layers.Layer_MaxPooling2D((2, 2))
        assert conv.weights.shape == (4, 1, 3, 3) and pool.pool_size == (2, 2)
    finally:
        sys.path.remove(tmp)
        for k in [k for k in sys.modules if k.startswith(pkg)]:
            del sys.modules[k]


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF_ROOT, "tests")), reason="reference tree only in the authoring container")
def test_reference_own_adam_unit_tests_pass_against_the_mirror():
    """The reference's tests/optimizers/test_Adam.py, UNMODIFIED, loaded from where it lies and run with `Neuromah`
    aliased to the mirror package: its module-path import (`from Neuromah.src.optimizers.Adam import Optimizer_Adam`,
    :4) and its three cases (defaults :13-20, custom values :22-29, ValueError on bad betas :31-35)."""
    import importlib.util
    import unittest
    import neuromah_b200
    neuromah_b200.install_alias()
    spec = importlib.util.spec_from_file_location("ref_test_Adam", os.path.join(REF_ROOT, "tests", "optimizers", "test_Adam.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    from neuromah_b200.src.optimizers import Optimizer_Adam
    assert mod.Optimizer_Adam is Optimizer_Adam
    suite = unittest.defaultTestLoader.loadTestsFromModule(mod)
    result = unittest.TextTestRunner(stream=open(os.devnull, "w")).run(suite)
    assert result.testsRun == 3 and result.wasSuccessful(), result.failures + result.errors


# ------------------------------------------------------------------ GPU: values through the seam ----
@pytest.mark.gpu
def test_shim_maxpool_reference_kats():
    """tests/layers/test_maxpool2D.py:14-57, value for value, through Layer_MaxPooling2D's call pair."""
    b = _shim()._MaxPooling2DBackend_cpp
    x = np.arange(1, 17, dtype=np.float32).reshape(1, 1, 4, 4)
    out, idx = b.maxpool2d_forward_cpu(x, 2, 2, 2, 2, "valid")                       # MaxPooling2D.py:53-58
    np.testing.assert_array_equal(out, np.array([[[[6, 8], [14, 16]]]], np.float32))
    assert out.dtype == np.float32 and idx.dtype == np.int32 and idx.shape == (1, 1, 2, 2, 2)
    dv = np.array([[[[1, 2], [3, 4]]]], np.float32)
    dx = b.maxpool2d_backward_cpu(dv, idx, x, 2, 2, 2, 2, "valid")                   # MaxPooling2D.py:73-80
    np.testing.assert_array_equal(dx, np.array([[[[0, 0, 0, 0], [0, 1, 0, 2], [0, 0, 0, 0], [0, 3, 0, 4]]]], np.float32))


@pytest.mark.gpu
@pytest.mark.parametrize("padding,ph,sh", [("valid", 2, 2), ("same", 3, 2), ("valid", 3, 1)])
def test_shim_maxpool_matches_oracle(padding, ph, sh):
    b = _shim()._MaxPooling2DBackend_cpp
    rng = np.random.default_rng(7)
    x = rng.standard_normal((3, 5, 9, 11)).astype(np.float32)
    out, idx = b.maxpool2d_forward_cpu(x, ph, ph, sh, sh, padding)
    ro, ri = R.maxpool2d_forward_cpu(x, ph, ph, sh, sh, padding)
    np.testing.assert_array_equal(out, ro)
    np.testing.assert_array_equal(idx, ri)
    dv = rng.standard_normal(out.shape).astype(np.float32)
    np.testing.assert_allclose(b.maxpool2d_backward_cpu(dv, idx, x, ph, ph, sh, sh, padding),
                               R.maxpool2d_backward_cpu(dv, ri, x, ph, ph, sh, sh, padding), rtol=1e-6, atol=1e-6)


@pytest.mark.gpu
def test_shim_conv_reference_shape_cases_and_oracle_values():
    """tests/layers/test_Conv2dBackend.py:6-43 pins shapes only; values against the oracle restatement of
    Conv2D.cpp:83-240 on Layer_Conv2D.forward's call pattern (np.pad outside the native call)."""
    b = _shim()._Conv2DBackend_cpp
    rng = np.random.default_rng(11)
    x = rng.standard_normal((2, 3, 8, 8)).astype(np.float32)
    w = rng.standard_normal((5, 3, 3, 3)).astype(np.float32)
    xp = np.pad(x, ((0, 0), (0, 0), (1, 1), (1, 1)), mode="constant")               # Conv_wrapepr.py:78-84 ('same', k = 3)
    y = b.conv2d_cpu(xp, w)
    assert y.shape == (2, 5, 8, 8) and y.dtype == np.float32
    np.testing.assert_allclose(y, R.conv2d_cpu(xp, w), rtol=1e-4, atol=1e-5)
    assert b.conv2d_cpu(x, w).shape == (2, 5, 6, 6)                                  # 'valid'
    dy = rng.standard_normal(y.shape).astype(np.float32)
    dx, dw = b.conv2d_backward_cpu(xp, w, dy)
    rdx, rdw = R.conv2d_backward_cpu(xp, w, dy)
    assert dx.shape == xp.shape and dw.shape == w.shape
    np.testing.assert_allclose(dx, rdx, rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(dw, rdw, rtol=1e-4, atol=1e-4)


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF_ROOT, "examples")), reason="reference tree only in the authoring container")
def test_every_neuromah_import_of_the_reference_examples_resolves_to_the_mirror():
    """Each `from Neuromah... import ...` line of the reference's example scripts and notebook, executed as written with
    `Neuromah` aliased to the mirror: the names resolve to the mirror's own objects (module-path imports included, e.g.
    examples/test_Dense_MNIST.py:4, :7).  The only lines that may fail are the federated HTTP client / server ones
    (`Neuromah.Federated.Client` / `.Auth`: FastAPI / JWT wire path, out of scope -- DESIGN.md section 7)."""
    import json
    import re
    import neuromah_b200
    neuromah_b200.install_alias()
    ex = os.path.join(REF_ROOT, "examples")
    lines = []
    for name in sorted(os.listdir(ex)):
        path = os.path.join(ex, name)
        if name.endswith(".py"):
            src = open(path, encoding="utf-8").read().splitlines()
        elif name.endswith(".ipynb"):
            nb = json.load(open(path, encoding="utf-8"))
            src = [l for c in nb["cells"] if c["cell_type"] == "code" for l in "".join(c["source"]).splitlines()]
        else:
            continue
        lines += [(name, l.split("#")[0].strip()) for l in src if re.match(r"\s*from\s+Neuromah[.\w]*\s+import\s+\w", l)]
    assert len(lines) >= 20
    resolved = 0
    for name, line in sorted(set(lines)):
        ns = {}
        if re.search(r"Neuromah\.Federated\.(Client|Server|Auth)\b", line):
            continue
        exec(line, ns)                                        # noqa: S102 - the reference's own import statement
        for k, v in ns.items():
            if k != "__builtins__":
                assert getattr(v, "__module__", "").startswith("neuromah_b200."), f"{name}: {line}: {k} -> {v!r}"
                resolved += 1
    assert resolved >= 30
