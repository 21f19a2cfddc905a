"""GPU parity for SURVEY.md 8f row 3: Optimizer_RMSprop / Adagrad / Nadam (and the vectorised SGD) slab kernels and
Layer_Dropout, against the oracle restatement and the goldens written by the unmodified reference.
FP32-exact bar: rtol 1e-4 / atol 1e-5."""
import numpy as np
import pytest

from conftest import golden, synth
from oracle import np_ref as R
from oracle import ref_net

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-4, 1e-5


def _dev(a):
    from neuromah_b200.device import DeviceArray
    return DeviceArray.from_numpy(np.ascontiguousarray(a, dtype=np.float32))


@pytest.mark.parametrize("n,n_apply", [(1, 1), (7, 2), (1024, 1), (4099, 2), (50186, 2)])
def test_slab_optimizer_kernels_match_oracle(n, n_apply):
    from neuromah_b200._lib import lib
    from neuromah_b200.device import stream
    rng = np.random.default_rng(n)
    p0 = rng.standard_normal(n)
    s = stream()
    for kind in ("sgd", "sgd_mom", "rmsprop", "adagrad", "nadam"):
        p, a, b = p0.copy(), np.abs(rng.standard_normal(n)) * 0.1, np.abs(rng.standard_normal(n)) * 0.1
        if kind == "sgd_mom":
            a = rng.standard_normal(n) * 0.1
        dp, da, db = _dev(p), _dev(a), _dev(b)
        p, a, b = (np.float32(v).astype(np.float64) for v in (p, a, b))
        for t in range(3):
            g = rng.standard_normal(n) * (10.0 ** -t)
            dg = _dev(g)
            g = np.float32(g).astype(np.float64)
            for _ in range(n_apply):
                if kind == "sgd":
                    p, _m = R.sgd_update(p, g, None, lr=0.05)
                elif kind == "sgd_mom":
                    p, a = R.sgd_update(p, g, a, lr=0.05, momentum=0.9)
                elif kind == "rmsprop":
                    p, a = R.rmsprop_update(p, g, a, lr=0.01, rho=0.9, eps=1e-7)
                elif kind == "adagrad":
                    p, a = R.adagrad_update(p, g, a, lr=0.1, eps=1e-7)
                else:
                    p, a, b = R.nadam_update(p, g, a, b, lr=0.01, beta1=0.9, beta2=0.999, eps=1e-8, t=t)
            if kind == "sgd":
                lib.nm_sgd_step_f32(dp.p, dg.p, None, n, 0.05, 0.0, n_apply, None, s)
            elif kind == "sgd_mom":
                lib.nm_sgd_step_f32(dp.p, dg.p, da.p, n, 0.05, 0.9, n_apply, None, s)
            elif kind == "rmsprop":
                lib.nm_rmsprop_step_f32(dp.p, dg.p, da.p, n, 0.01, 0.9, 1e-7, n_apply, None, s)
            elif kind == "adagrad":
                lib.nm_adagrad_step_f32(dp.p, dg.p, da.p, n, 0.1, 1e-7, n_apply, None, s)
            else:
                lib.nm_nadam_step_f32(dp.p, dg.p, da.p, db.p, n, 0.01, 0.9, 0.999, 1e-8, 1 - 0.9 ** (t + 1), 1 - 0.999 ** (t + 1),
                                      n_apply, s)
            np.testing.assert_allclose(dp.numpy(), p, rtol=RTOL, atol=ATOL, err_msg=f"{kind} step {t}")
        if kind != "sgd":
            np.testing.assert_allclose(da.numpy(), a, rtol=RTOL, atol=ATOL, err_msg=kind)


def test_unaligned_slab_views_take_the_scalar_path():
    """A view that starts 4 bytes into a slab is not 16-byte aligned: same result as the vector path."""
    from neuromah_b200._lib import lib
    from neuromah_b200.device import stream
    import ctypes as C
    rng = np.random.default_rng(3)
    n = 1001
    p, g, c = rng.standard_normal(n + 1), rng.standard_normal(n + 1), np.abs(rng.standard_normal(n + 1))
    dp, dg, dc = _dev(p), _dev(g), _dev(c)
    off = lambda d: C.c_void_p(d.ptr + 4)
    lib.nm_rmsprop_step_f32(off(dp), off(dg), off(dc), n, 0.01, 0.9, 1e-7, 1, None, stream())
    pe, ce = R.rmsprop_update(np.float32(p[1:]).astype(np.float64), np.float32(g[1:]).astype(np.float64),
                              np.float32(c[1:]).astype(np.float64), lr=0.01)
    np.testing.assert_allclose(dp.numpy()[1:], pe, rtol=RTOL, atol=ATOL)
    assert dp.numpy()[0] == np.float32(p[0])


def _softmax_regression(kind, **kw):
    from neuromah_b200.src.core import Model
    from neuromah_b200.src.layers import Layer_Dense
    from neuromah_b200.src.activations import Activation_Softmax
    from neuromah_b200.src.losses import Loss_CategoricalCrossentropy
    from neuromah_b200.src import optimizers as O
    from neuromah_b200.src.metrics import Accuracy_Categorical
    spec = [{"kind": "dense", "n_in": 64, "n_out": 10, "relu": False}, {"kind": "softmax"}]
    params = ref_net.init_params(spec, np.random.RandomState(31))
    model = Model()
    d = Layer_Dense(64, 10)
    d.set_parameters(params[0]["weights"], params[0]["biases"])
    model.add(d)
    model.add(Activation_Softmax())
    opt = {"rmsprop": O.Optimizer_RMSprop, "adagrad": O.Optimizer_Adagrad, "nadam": O.Optimizer_Nadam}[kind](**kw)
    model.set(loss=Loss_CategoricalCrossentropy, optimizer=opt, accuracy=Accuracy_Categorical)
    model.finalize()
    return model


@pytest.mark.parametrize("kind,kw", [("rmsprop", dict(learning_rate=0.01, decay=1e-3)), ("adagrad", dict(learning_rate=0.05)),
                                     ("nadam", dict(learning_rate=0.01, decay=1e-2))])
@pytest.mark.parametrize("api", ["per_layer", "train_step"])
def test_single_layer_optimizer_goldens(kind, kw, api):
    """20 steps of softmax regression against the unmodified reference's optimizers (with the Q1 double application),
    through the reference's call triple (update_params per registered layer) and through Model.train_step (one slab launch)."""
    from test_gpu_model import reference_style_step
    g = golden("optimizers_single_layer_20steps.npz")
    model = _softmax_regression(kind, **kw)
    model.loss.new_pass(); model.accuracy.new_pass()
    x, y = synth(32, 32 * 20, (64,))
    losses = []
    for s in range(20):
        xb, yb = x[s * 32:(s + 1) * 32], y[s * 32:(s + 1) * 32]
        if api == "per_layer":
            losses.append(reference_style_step(model, xb, yb)[0])
        else:
            model.train_step(xb, yb)
    if api == "per_layer":
        np.testing.assert_allclose(losses, g[f"{kind}_losses"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(model.arena.host_params(), g[f"{kind}_params"], rtol=RTOL, atol=ATOL)
    assert model.optimizer.iterations == 20


def test_name_keyed_state_fails_like_the_reference_on_two_shapes():
    """RMSprop.py:99-101 keys the cache by 'weights' / 'biases' only: a second, differently shaped layer cannot broadcast."""
    from neuromah_b200.src.layers import Layer_Dense
    from neuromah_b200.src.optimizers import Optimizer_RMSprop
    a, b = Layer_Dense(8, 4), Layer_Dense(4, 3)
    for l in (a, b):
        l.forward(np.ones((2, l.weights.shape[0]), np.float32), True)
        l.backward(np.ones((2, l.weights.shape[1]), np.float32))
    opt = Optimizer_RMSprop()
    opt.update_params(a)
    with pytest.raises(ValueError):
        opt.update_params(b)


def test_nonfinite_gradient_raises_value_error():
    from neuromah_b200.src.layers import Layer_Dense
    from neuromah_b200.src.optimizers import Optimizer_Adagrad
    l = Layer_Dense(4, 3)
    l.forward(np.ones((2, 4), np.float32), True)
    l.backward(np.array([[np.nan, 0, 0], [0, 0, 0]], np.float32))
    with pytest.raises(ValueError):
        Optimizer_Adagrad().update_params(l)


@pytest.mark.parametrize("opt_name", ["sgd", "rmsprop", "adagrad"])
def test_nonfinite_gradient_raises_through_model_train(opt_name):
    """The arena path Model.train uses (update_arena) keeps the reference's ValueError on NaN/Inf gradients
    (SGD.py:80-83, RMSprop.py:93-96): the device flag is read once per epoch beside the loss sums."""
    from neuromah_b200.src.optimizers import Optimizer_SGD, Optimizer_RMSprop, Optimizer_Adagrad
    from test_gpu_model import build
    spec = ref_net.mlp_spec(16, 8, 4)
    model = build(spec, ref_net.init_params(spec, np.random.RandomState(0)), optimizer="sgd", learning_rate=0.1)
    opt = {"sgd": Optimizer_SGD(learning_rate=0.1), "rmsprop": Optimizer_RMSprop(), "adagrad": Optimizer_Adagrad()}[opt_name]
    model.optimizer = opt
    opt.bind_arena(model.arena)
    x = np.random.default_rng(0).random((32, 16), dtype=np.float32)
    y = np.eye(4)[np.random.default_rng(1).integers(0, 4, 32)]
    model.train(x, y, epochs=1, batch_size=16, verbose=0)          # finite gradients: no error
    x[3, 5] = np.nan
    with pytest.raises(ValueError, match="Gradient"):
        model.train(x, y, epochs=1, batch_size=16, verbose=0)


def test_constructor_validation_matches_reference():
    from neuromah_b200.src.optimizers import Optimizer_RMSprop, Optimizer_Adagrad, Optimizer_Nadam
    for bad in (dict(learning_rate=0), dict(decay=-1), dict(epsilon=0), dict(rho=1.0)):
        with pytest.raises(ValueError):
            Optimizer_RMSprop(**bad)
    with pytest.raises(ValueError):
        Optimizer_Adagrad(learning_rate=-1.0)
    with pytest.raises(TypeError):
        Optimizer_Nadam(beta1=1)
    with pytest.raises(ValueError):
        Optimizer_Nadam(beta2=1.5)


def _dropout_model(fixed=None):
    from neuromah_b200.src.core import Model
    from neuromah_b200.src.layers import Layer_Dense, Layer_Dropout
    from neuromah_b200.src.activations import Activation_ReLU, Activation_Softmax
    from neuromah_b200.src.losses import Loss_CategoricalCrossentropy
    from neuromah_b200.src.optimizers import Optimizer_Adam
    from neuromah_b200.src.metrics import Accuracy_Categorical
    spec = [{"kind": "dense", "n_in": 64, "n_out": 32, "relu": True}, {"kind": "dropout", "rate": 0.3},
            {"kind": "dense", "n_in": 32, "n_out": 10, "relu": False}, {"kind": "softmax"}]
    params = ref_net.init_params(spec, np.random.RandomState(41))
    model = Model()
    d1 = Layer_Dense(64, 32, activation=Activation_ReLU()); d1.set_parameters(params[0]["weights"], params[0]["biases"])
    drop = Layer_Dropout(0.3)
    d2 = Layer_Dense(32, 10); d2.set_parameters(params[2]["weights"], params[2]["biases"])
    for l in (d1, drop, d2, Activation_Softmax()):
        model.add(l)
    model.set(loss=Loss_CategoricalCrossentropy, optimizer=Optimizer_Adam(learning_rate=0.01), accuracy=Accuracy_Categorical)
    model.finalize()
    return model, drop


def test_dropout_mlp_golden_with_the_reference_masks():
    from test_gpu_model import reference_style_step
    g = golden("dropout_mlp_10steps.npz")
    model, drop = _dropout_model()
    model.loss.new_pass(); model.accuracy.new_pass()
    x, y = synth(42, 16 * 10, (64,))
    losses = []
    for s in range(10):
        drop.fixed_mask = g["masks"][s]
        losses.append(reference_style_step(model, x[s * 16:(s + 1) * 16], y[s * 16:(s + 1) * 16])[0])
    np.testing.assert_allclose(losses, g["losses"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(model.arena.host_params(), g["final_params"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(np.asarray(model.forward(x[:16], training=False)), g["infer"], rtol=RTOL, atol=ATOL)


def test_dropout_device_mask_properties():
    from neuromah_b200.src.layers import Layer_Dropout
    rng = np.random.default_rng(0)
    x = rng.standard_normal((257, 1031)).astype(np.float32)            # odd sizes: vector body + scalar tail
    n = x.size
    for rate in (0.1, 0.5, 0.8):
        keep = 1 - rate
        lay = Layer_Dropout(rate, seed=77)
        out = np.asarray(lay.forward(x, True))
        m = np.asarray(lay.binary_mask)
        assert set(np.unique(m)) == {np.float32(0), np.float32(1.0 / np.float32(keep))}
        np.testing.assert_allclose(out, x * m, rtol=1e-6, atol=0)
        frac = np.mean(m > 0)
        assert abs(frac - keep) < 5 * np.sqrt(keep * (1 - keep) / n)          # Bernoulli(keep), 5 sigma
        # no structure along rows / columns
        assert np.abs(np.mean(m > 0, axis=0) - keep).max() < 6 * np.sqrt(keep * (1 - keep) / x.shape[0])
        dv = rng.standard_normal(x.shape).astype(np.float32)
        np.testing.assert_allclose(np.asarray(lay.backward(dv)), dv * m, rtol=1e-6, atol=0)     # same mask (Dropout.py:40)
        first = m.copy()
        lay.forward(x, True)                                                   # next call: fresh mask
        second = np.asarray(lay.binary_mask)
        assert 0.2 < np.mean((first > 0) != (second > 0)) / (2 * keep * (1 - keep)) < 1.8
        again = Layer_Dropout(rate, seed=77)
        again.forward(x, True)
        np.testing.assert_array_equal(np.asarray(again.binary_mask), first)   # same seed + call index => same mask
        inf = np.asarray(lay.forward(x, False))
        np.testing.assert_array_equal(inf, x)                                  # inference: copy (Dropout.py:23-25)


def test_dense_example_stack_as_written_in_the_reference():
    """examples/test_Dense_MNIST.py:38-49 call for call (reduced widths): Dropout layers, a Dense with an embedded Softmax
    followed by Activation_Softmax (the doubled softmax, SURVEY Q10: its Jacobian backward runs on the device), loss and
    accuracy handed to Model.set as instances.  Goldens from the unmodified reference, dropout masks replayed."""
    from test_gpu_model import reference_style_step
    from neuromah_b200.src.core import Model
    from neuromah_b200.src.layers import Layer_Dense, Layer_Dropout
    from neuromah_b200.src.activations import Activation_ReLU, Activation_Softmax
    from neuromah_b200.src.losses import Loss_CategoricalCrossentropy
    from neuromah_b200.src.optimizers import Optimizer_Adam
    from neuromah_b200.src.metrics import Accuracy_Categorical
    g = golden("dense_example_stack_6steps.npz")
    model = Model()
    d1 = Layer_Dense(n_inputs=64, n_neurons=32, activation=Activation_ReLU())
    d2 = Layer_Dense(n_inputs=32, n_neurons=16, activation=Activation_ReLU())
    d3 = Layer_Dense(n_inputs=16, n_neurons=10, activation=Activation_Softmax())
    for d, k in zip((d1, d2, d3), ("w0", "w1", "w2")):
        d.set_parameters(g[k], np.zeros((1, g[k].shape[1])))
    drop1, drop2 = Layer_Dropout(rate=0.25), Layer_Dropout(rate=0.25)
    for l in (d1, drop1, d2, drop2, d3, Activation_Softmax()):
        model.add(l)
    model.set(loss=Loss_CategoricalCrossentropy(model=model), optimizer=Optimizer_Adam(learning_rate=0.01),
              accuracy=Accuracy_Categorical(model=model))
    model.finalize()
    model.loss.new_pass(); model.accuracy.new_pass()
    x, y = synth(62, 32 * 6, (64,))
    losses = []
    for s in range(6):
        drop1.fixed_mask, drop2.fixed_mask = g["masks1"][s], g["masks2"][s]
        losses.append(reference_style_step(model, x[s * 32:(s + 1) * 32], y[s * 32:(s + 1) * 32])[0])
    np.testing.assert_allclose(losses, g["losses"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(model.arena.host_params(), g["final_params"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(np.asarray(model.forward(x[:32], training=False)), g["infer"], rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("n,bins", [(288, 10), (50186, 10), (1 << 20, 64), (7, 3)])
def test_device_histogram_equals_numpy(n, bins):
    """TensorMonitor.log_histogram on a device tensor (TensorMonitor.py:93-104): counts and edges equal np.histogram's on the
    float64 values the reference holds its parameters in."""
    from neuromah_b200.device import DeviceArray
    from neuromah_b200.src.utils.TensorMonitor import device_histogram
    rng = np.random.default_rng(n)
    x = (rng.standard_normal(n) * 0.3).astype(np.float32)
    x[::5] = np.round(x[::5], 1)                                   # repeated values, some exactly on bin edges
    hist, edges = device_histogram(DeviceArray.from_numpy(x), bins)
    ref_h, ref_e = np.histogram(x.astype(np.float64), bins=bins)
    np.testing.assert_array_equal(hist, ref_h)
    np.testing.assert_array_equal(edges, ref_e)
    const = DeviceArray.from_numpy(np.full((100,), 0.25, np.float32))
    h2, e2 = device_histogram(const, 10)
    r2, re2 = np.histogram(np.full((100,), 0.25, np.float64), bins=10)
    np.testing.assert_array_equal(h2, r2)
    np.testing.assert_array_equal(e2, re2)
