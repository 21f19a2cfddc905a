"""CPU tests: the C-ABI library loads and exports every symbol include/neuromah_b200.h declares
(no compute calls without a GPU), plus the host-side logic of the mirror API."""
import ctypes
import os

import numpy as np
import pytest

import neuromah_b200
from neuromah_b200 import ops, parallel
from neuromah_b200._lib import LIB_PATH, _SIGS, declared_symbols, lib


def test_library_exports_every_declared_symbol():
    assert os.path.exists(LIB_PATH), "build the library first: python -c 'import __graft_entry__ as g; g.build()'"
    dll = ctypes.CDLL(LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 50
    missing = [n for n in names if not hasattr(dll, n)]
    assert not missing, f"declared in the header but not exported: {missing}"
    assert set(names) == set(_SIGS), "ctypes signature table out of sync with the header"


def test_version_and_error_string_without_gpu():
    dll = lib.load()
    assert dll.nm_version() == 1
    assert isinstance(dll.nm_last_error(), bytes)


def test_no_cpu_fallback_exists():
    """The product package must not import the oracle (or fall back to NumPy compute)."""
    root = os.path.dirname(neuromah_b200.__file__)
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f


def test_same_padding_rule():
    assert ops.same_padding(3) == (1, 1) and ops.same_padding(5) == (2, 2)
    assert ops.same_padding(2) == (0, 1) and ops.same_padding(4) == (1, 2)      # Conv_wrapepr.py:134-137
    assert ops.same_padding(1) == (0, 0)


def test_pool_geometry_matches_reference_rule():
    assert ops.pool_geometry(28, 28, 2, 2, 2, 2, "valid") == (14, 14, 0, 0)
    assert ops.pool_geometry(7, 9, 3, 3, 2, 2, "same") == (4, 5, 1, 1)
    assert ops.pool_geometry(7, 9, 2, 3, 1, 2, "same") == (7, 5, 0, 1)


def test_shard_bounds():
    assert [parallel.shard_bounds(4096, r, 8) for r in (0, 7)] == [(0, 512), (3584, 4096)]
    with pytest.raises(ValueError):
        parallel.shard_bounds(10, 0, 4)


def test_constructor_validation_matches_reference_errors():
    """tests/layers/test_Conv2dBackend.py:64-81 and tests/optimizers/test_Adam.py:13-35 (no device needed)."""
    from neuromah_b200.src.layers import Layer_Conv2D, Layer_Dense, Layer_MaxPooling2D
    from neuromah_b200.src.optimizers import Optimizer_Adam, Optimizer_SGD
    with pytest.raises(ValueError):
        Layer_Conv2D(in_channels=3, out_channels=2, kernel_size=3, padding="invalid")
    with pytest.raises(ValueError):
        Layer_Conv2D(in_channels=3, out_channels=2, kernel_size=(3, -1))
    with pytest.raises(ValueError):
        Layer_Conv2D(in_channels=3, out_channels=2, kernel_size=3, stride=(1, -1))
    with pytest.raises(ValueError):
        Layer_Conv2D(in_channels=-1, out_channels=2, kernel_size=3)
    with pytest.raises(ValueError):
        Layer_Conv2D(in_channels=3, out_channels=-2, kernel_size=3)
    with pytest.raises(ValueError):
        Layer_Dense(0, 3)
    with pytest.raises(ValueError):
        Layer_MaxPooling2D((1, 2, 3))
    opt = Optimizer_Adam()
    assert (opt.learning_rate, opt.decay, opt.beta_1, opt.beta_2, opt.epsilon, opt.iterations) == \
           (0.001, 0.0, 0.9, 0.999, 1e-7, 0)
    opt = Optimizer_Adam(learning_rate=0.2, epsilon=1e-5, beta_1=0.8, beta_2=0.888, decay=0.5)
    assert (opt.learning_rate, opt.beta_1, opt.beta_2, opt.epsilon, opt.decay) == (0.2, 0.8, 0.888, 1e-5, 0.5)
    with pytest.raises(ValueError):
        Optimizer_Adam(beta_1=1.2, beta_2=0.5)
    with pytest.raises(ValueError):
        Optimizer_Adam(beta_1=0.8, beta_2=-0.5)
    with pytest.raises(ValueError):
        Optimizer_SGD(learning_rate=-1)
    with pytest.raises(TypeError):
        Optimizer_SGD(learning_rate="x")


def test_initializers_draw_like_the_reference():
    from neuromah_b200.src.initializers import HeInitializer, RandomNormalInitializer, XavierInitializer
    np.random.seed(0)
    a = RandomNormalInitializer().initialize((3, 4))
    np.random.seed(0)
    np.testing.assert_array_equal(a, np.random.randn(3, 4) * 0.01)          # Initializer.py:96
    np.random.seed(1)
    b = HeInitializer().initialize((8, 3, 3, 3))
    np.random.seed(1)
    np.testing.assert_array_equal(b, np.random.normal(0, np.sqrt(2.0 / 27), size=(8, 3, 3, 3)))
    np.random.seed(2)
    c = XavierInitializer().initialize((6, 5))
    lim = np.sqrt(6.0 / 11)
    assert np.all(np.abs(c) <= lim)


def test_alias_makes_reference_imports_resolve():
    neuromah_b200.install_alias()
    from Neuromah.src.layers import Layer_Conv2D, Layer_MaxPooling2D, Layer_Dense, Layer_Flatten   # noqa: F401
    from Neuromah.src.optimizers import Optimizer_Adam          # noqa: F401
    from Neuromah.src.losses import Loss_CategoricalCrossentropy  # noqa: F401
    from Neuromah.src.core import Model                         # noqa: F401
    from Neuromah.src.activations import Activation_ReLU, Activation_Softmax   # noqa: F401
    from Neuromah.src.metrics import Accuracy_Categorical       # noqa: F401
    # the rest of the reference's package surface around Model.train (activations/__init__.py:1-4, losses/__init__.py:1-4,
    # metrics/__init__.py:1-2, core/__init__.py:1-5)
    from Neuromah.src.activations import Activation_Linear, Activation_Sigmoid
    from Neuromah.src.losses import Loss_BinaryCrossentropy, Loss_MeanAbsoluteError, Loss_MeanSquaredError  # noqa: F401
    from Neuromah.src.metrics import Accuracy_Regression
    from Neuromah.src.core import Accuracy, Activation, Loss
    assert issubclass(Activation_Sigmoid, Activation) and issubclass(Activation_Linear, Activation)
    assert issubclass(Loss_MeanSquaredError, Loss) and issubclass(Accuracy_Regression, Accuracy)


def test_alias_deep_imports_share_module_and_class_identity():
    """The reference notebook imports classes by their module path (examples/test.ipynb cell 1:
    `from Neuromah.src.losses.categorical_crossentropy import Loss_CategoricalCrossentropy`,
    `from Neuromah.src.metrics.Accuracy_Categorical import Accuracy_Categorical`).  Under the alias these must be the
    SAME objects as the mirror package's -- Model.finalize selects the fused softmax / cross-entropy path with
    isinstance -- and the names the package __init__s rebind from submodule to class must stay classes."""
    import sys
    neuromah_b200.install_alias()
    from Neuromah.src.losses.categorical_crossentropy import Loss_CategoricalCrossentropy as A
    from Neuromah.src.metrics.Accuracy_Categorical import Accuracy_Categorical as C
    from Neuromah.src.core.Model import Model as M
    import Neuromah.device as dev_alias
    import neuromah_b200.device as dev_real
    from neuromah_b200.src.losses import Loss_CategoricalCrossentropy
    from neuromah_b200.src.metrics import Accuracy_Categorical
    from neuromah_b200.src.core import Model
    assert A is Loss_CategoricalCrossentropy and C is Accuracy_Categorical and M is Model
    assert isinstance(Model, type) and isinstance(Accuracy_Categorical, type)
    assert dev_alias is dev_real                          # one library stream / launch counter, not a copy per name
    assert sys.modules["Neuromah.src.layers.Dense"] is sys.modules["neuromah_b200.src.layers.Dense"]
    assert not any(m.startswith("Neuromah.compat") for m in sys.modules)
    with pytest.raises(ImportError):
        import Neuromah.src.layers.RNN                    # noqa: F401  (absent from the reference as well)


def test_regression_accuracy_and_element_counting_on_the_host():
    """Accuracy_Regression (Accuracy_Regression.py:10-17) and Accuracy.calculate (Accuracy.py:19-23) are host arithmetic."""
    from neuromah_b200.src.metrics import Accuracy_Categorical, Accuracy_Regression
    y = np.array([[0.0, 1.0], [2.0, 3.0], [4.0, 5.0]])
    acc = Accuracy_Regression(model=None)
    acc.init(y)
    assert acc.precision == np.std(y) / 250
    acc.init(y * 10)                                   # kept unless reinit
    assert acc.precision == np.std(y) / 250
    acc.init(y * 10, reinit=True)
    assert acc.precision == np.std(y * 10) / 250
    acc.new_pass()
    pred = y + np.array([[0.0, 1.0], [0.0, 0.0], [1.0, 0.5]])
    assert acc.calculate(pred, y) == 0.5               # mean over the six compared elements
    assert (acc.accumulated_sum, acc.accumulated_count) == (3, 3)      # element hits, but samples counted
    assert acc.calculate_accumulated() == 1.0
    binary = Accuracy_Categorical(binary=True, model=None)
    binary.new_pass()
    assert binary.calculate(np.array([[1, 0], [1, 1]]), np.array([[1.0, 1.0], [1.0, 1.0]])) == 0.75
    with pytest.raises(NotImplementedError):
        from neuromah_b200.src.core import Activation
        Activation().forward(None, False)


def test_federated_average_host_mirror():
    from neuromah_b200.Federated import federated_average
    out = federated_average([[1.0, 2.0], [3.0, 6.0]], [100, 300])
    assert out == [2.5, 5.0]


def test_tensor_monitor_json_shape(tmp_path):
    """TensorMonitor keeps the reference's hooks and JSON layout (src/utils/TensorMonitor.py:7-136)."""
    import json
    import types
    from neuromah_b200.src.utils import TensorMonitor
    Layer_Conv2D = type("Layer_Conv2D", (), {})
    layer = Layer_Conv2D()
    layer.__dict__.update(weights=np.arange(12.0).reshape(3, 4), biases=np.zeros((1, 4)),
                          weight_gradients=np.ones((3, 4)), bias_gradients=np.ones((1, 4)), kernel_size=(3, 3))
    model = types.SimpleNamespace(layers=[layer, object()], optimizer=types.SimpleNamespace(learning_rate=0.1),
                                  loss=object(), accuracy=object())
    tm = TensorMonitor(log_dir=str(tmp_path))
    tm.start_run(model)
    tm.start_epoch(1)
    tm.log_scalar("loss/epoch_training", 0.5)
    tm.log_layer_parameters(layer)
    tm.end_epoch()
    tm.save_logs()
    log = json.load(open(tm.log_file_path))
    assert set(log) == {"metadata", "epochs"}
    assert log["metadata"]["layers"][0]["name"] == "Layer_Conv2D"
    assert log["metadata"]["layers"][0]["config"]["weights"] == {"shape": [3, 4], "dtype": "float64"}
    ep = log["epochs"][0]
    assert ep["epoch"] == 1 and ep["metrics"]["scalars"]["loss/epoch_training"] == 0.5 and ep["time_elapsed"] >= 0
    hist = ep["parameters"]["histograms"]
    assert set(hist) == {"Layer_Conv2D/weights", "Layer_Conv2D/dweights", "Layer_Conv2D/biases", "Layer_Conv2D/dbiases"}
    assert sum(hist["Layer_Conv2D/weights"]["histogram"]) == 12 and len(hist["Layer_Conv2D/weights"]["bin_edges"]) == 11


def test_next_row_optimizers_and_dropout_host_side():
    """SURVEY 8f row 3 on the host: constructor validation, attribute names and defaults as in the reference
    (src/optimizers/RMSprop.py:21-40, Adagrad.py, Nadam.py; src/layers/Dropout.py:12) -- no device needed."""
    from neuromah_b200.src.optimizers import Optimizer_Adagrad, Optimizer_Nadam, Optimizer_RMSprop
    from neuromah_b200.src.layers import Layer_Dropout
    r = Optimizer_RMSprop()
    assert (r.learning_rate, r.current_learning_rate, r.decay, r.iterations, r.epsilon, r.rho) == (0.001, 0.001, 0., 0, 1e-7, 0.9)
    a = Optimizer_Adagrad()
    assert (a.learning_rate, a.decay, a.epsilon, a.iterations) == (1., 0., 1e-7, 0)
    n = Optimizer_Nadam()
    assert (n.learning_rate, n.beta1, n.beta2, n.epsilon) == (1e-3, 0.9, 0.999, 1e-8)
    for opt in (r, a, n):
        assert opt.caches == {}
    for bad in (dict(learning_rate=0), dict(learning_rate="x"), dict(decay=-1), dict(epsilon=0), dict(rho=1.0), dict(rho=-0.1)):
        with pytest.raises(ValueError):
            Optimizer_RMSprop(**bad)
    for bad in (dict(learning_rate=-1.0), dict(decay=-0.5), dict(epsilon=-1e-7)):
        with pytest.raises(ValueError):
            Optimizer_Adagrad(**bad)
    with pytest.raises(TypeError):
        Optimizer_Nadam(learning_rate="fast")
    with pytest.raises(TypeError):
        Optimizer_Nadam(beta1=1)                       # the reference insists on float betas
    with pytest.raises(ValueError):
        Optimizer_Nadam(beta2=1.0)
    with pytest.raises(ValueError):
        Optimizer_Nadam(epsilon=0.0)
    # decayed learning rate: lr / (1 + decay * iterations), evaluated in pre_update_params
    r = Optimizer_RMSprop(learning_rate=0.1, decay=0.5)
    r.iterations = 4
    r.pre_update_params()
    assert r.current_learning_rate == pytest.approx(0.1 / 3.0)
    r.post_update_params()
    assert r.iterations == 5
    # Layer_Dropout keeps the KEEP probability in .rate, like the reference
    d = Layer_Dropout(0.3)
    assert d.rate == pytest.approx(0.7) and d.binary_mask is None
    with pytest.raises(ValueError):
        Layer_Dropout(1.0)
    assert Layer_Dropout(0.1).seed != Layer_Dropout(0.1).seed           # distinct default streams per layer
    for layer_with_no_params in (d,):
        assert layer_with_no_params.get_parameters() == {}


def test_checkpoint_format_is_guarded(tmp_path):
    """Model.load refuses files that are not neuromah_b200 model checkpoints (host side only)."""
    import pickle
    from neuromah_b200 import checkpoint
    p = tmp_path / "not_a_model.pkl"
    with open(p, "wb") as f:
        pickle.dump({"format": "something else"}, f)
    with pytest.raises(ValueError):
        checkpoint.load(str(p))


def _header_prototypes():
    """name -> (return kind, [argument kinds]) parsed from include/neuromah_b200.h; kinds: ptr, int, uint, float, double,
    size_t, llong, ullong."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(neuromah_b200.__file__)))
    text = open(os.path.join(root, "include", "neuromah_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    text = re.sub(r"//[^\n]*", " ", text)

    def kind(decl):
        d = " ".join(decl.replace("const", " ").split())
        if "*" in d:
            return "ptr"
        words = d.split()
        words = words[:-1] if len(words) > 1 and words[-1] not in ("int", "long", "float", "double", "size_t", "unsigned") else words
        t = " ".join(words)
        return {"int": "int", "unsigned": "uint", "unsigned int": "uint", "float": "float", "double": "double", "size_t": "size_t",
                "long long": "llong", "unsigned long long": "ullong"}[t]
    out = {}
    for m in re.finditer(r"([A-Za-z_][\w \*]*?)\b(nm_\w+)\s*\(([^)]*)\)\s*;", text):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        arglist = [] if args in ("", "void") else [kind(a) for a in args.split(",")]
        out[name] = ("ptr" if "*" in ret else kind(ret + " x"), arglist)
    return out


def test_ctypes_signatures_match_the_header_prototypes():
    """Every prototype of include/neuromah_b200.h against the hand-written ctypes table: same argument count, and the same
    kind of C type in every position (a drifted signature corrupts the call silently)."""
    import ctypes as C
    protos = _header_prototypes()
    assert set(protos) == set(_SIGS)

    def ckind(t):
        if t in (C.c_void_p, C.c_char_p) or (isinstance(t, type) and issubclass(t, C._Pointer)):
            return "ptr"
        return {C.c_int: "int", C.c_uint: "uint", C.c_float: "float", C.c_double: "double", C.c_size_t: "size_t",
                C.c_longlong: "llong", C.c_ulonglong: "ullong"}[t]
    bad = []
    same = {"size_t": "u64", "ullong": "u64", "llong": "i64"}        # ctypes aliases the 64-bit integer types on LP64
    norm = lambda k: same.get(k, k)
    for name, (ret, args) in protos.items():
        res, argtypes = _SIGS[name]
        got = (norm(ckind(res)), [norm(ckind(a)) for a in argtypes])
        ret, args = norm(ret), [norm(a) for a in args]
        if got != (ret, args):
            bad.append((name, (ret, args), got))
    assert not bad, bad
