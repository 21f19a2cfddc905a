"""GPU parity, op level: every FP32-exact kernel through the C ABI (via the mirror layer classes)
against the oracle on identical seeded inputs.  Tolerance = the north star's FP32-exact bar:
rtol 1e-4 / atol 1e-5; integer / index results (argmax indices, correct flags) bit-exact."""
import numpy as np
import pytest

from conftest import golden
from oracle import np_ref as R

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-4, 1e-5


def close(a, b, rtol=RTOL, atol=ATOL):
    np.testing.assert_allclose(np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64), rtol=rtol, atol=atol)


CONV_CASES = [
    # N, C, H, W, OC, (kh, kw), padding, relu
    (1, 3, 5, 5, 2, (3, 3), "valid", False),      # tests/layers/test_Conv2dBackend.py:6-24
    (1, 3, 5, 5, 2, (3, 3), "same", False),
    (4, 1, 28, 28, 32, (3, 3), "same", True),     # MNIST conv1
    (3, 32, 14, 14, 64, (3, 3), "same", True),    # MNIST conv2
    (2, 5, 9, 7, 7, (2, 4), "same", True),        # even kernel, ragged tiles
    (2, 6, 8, 8, 5, (1, 1), "valid", False),
    (2, 3, 11, 10, 70, (5, 3), "valid", True),    # OC > one N tile
    (5, 17, 6, 6, 130, (3, 3), "same", False),    # C*k*k not a multiple of BK
]


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv2d_forward_backward(case):
    from neuromah_b200.src.layers import Layer_Conv2D
    from neuromah_b200.src.activations import Activation_ReLU
    n, c, h, w, oc, k, padding, relu = case
    rng = np.random.default_rng(hash(case) % 2**32)
    x = rng.standard_normal((n, c, h, w)).astype(np.float32)
    np.random.seed(1)
    layer = Layer_Conv2D(c, oc, k, padding=padding, activation=Activation_ReLU() if relu else None)
    wt = (rng.standard_normal((oc, c) + k) * 0.3).astype(np.float32)
    layer.set_parameters(wt, np.zeros((oc, 1), np.float32))
    layer.forward(x, training=True)
    ref_out, ref_pre = R.conv_layer_forward(x, wt, None, padding=padding, relu=relu)
    assert layer.output.shape == ref_out.shape
    close(layer.output, ref_out)
    dv = rng.standard_normal(ref_out.shape).astype(np.float32)
    layer.backward(dv)
    dx, dw, db = R.conv_layer_backward(x, wt, ref_pre, dv, padding=padding, relu=relu, mode="wired")
    assert layer.weight_gradients.shape == wt.shape and layer.bias_gradients.shape == (oc, 1)
    close(layer.dinputs, dx, atol=1e-4)
    close(layer.weight_gradients, dw, atol=1e-4)
    close(layer.bias_gradients, db, atol=1e-4)


def test_conv2d_stub_backward_mode_matches_shipped_reference():
    """as shipped, dweights = dinputs = 0 and only dbiases is real."""
    from neuromah_b200 import config
    from neuromah_b200.src.layers import Layer_Conv2D
    rng = np.random.default_rng(3)
    x = rng.standard_normal((2, 3, 6, 6)).astype(np.float32)
    layer = Layer_Conv2D(3, 4, 3, padding="same")
    layer.forward(x, True)
    dv = rng.standard_normal((2, 4, 6, 6)).astype(np.float32)
    config.conv_backward = "stub"
    try:
        layer.backward(dv)
    finally:
        config.conv_backward = "wired"
    assert not np.any(np.asarray(layer.weight_gradients)) and not np.any(np.asarray(layer.dinputs))
    close(layer.bias_gradients, dv.sum(axis=(0, 2, 3)).reshape(4, 1))


def test_conv2d_bias_switch():
    """the reference never adds the conv bias; config.conv_bias_in_forward turns it on."""
    from neuromah_b200 import config
    from neuromah_b200.src.layers import Layer_Conv2D
    rng = np.random.default_rng(8)
    x = rng.standard_normal((2, 2, 5, 5)).astype(np.float32)
    layer = Layer_Conv2D(2, 3, 3, padding="valid")
    wt = np.asarray(layer.weights)
    b = rng.standard_normal((3, 1)).astype(np.float32)
    layer.set_parameters(wt, b)
    layer.forward(x, True)
    close(layer.output, R.conv2d_cpu(x, wt))
    config.conv_bias_in_forward = True
    try:
        layer.forward(x, True)
    finally:
        config.conv_bias_in_forward = False
    close(layer.output, R.conv2d_cpu(x, wt) + b.reshape(1, 3, 1, 1))


def test_reference_conv_unit_tests_ported():
    """tests/layers/test_Conv2dBackend.py:6-62 against the device layer."""
    from neuromah_b200.src.layers import Layer_Conv2D
    layer = Layer_Conv2D(in_channels=3, out_channels=2, kernel_size=3, padding="valid")
    assert layer.weights.shape == (2, 3, 3, 3) and abs(float(np.mean(np.asarray(layer.weights)))) < 0.1
    inputs = np.random.randn(1, 3, 5, 5)
    layer.forward(inputs, training=True)
    assert layer.output.shape == (1, 2, 3, 3)
    layer.backward(np.random.randn(1, 2, 3, 3))
    assert layer.weight_gradients.shape == (2, 3, 3, 3) and layer.bias_gradients.shape == (2, 1)
    layer = Layer_Conv2D(in_channels=3, out_channels=2, kernel_size=3, padding="same")
    layer.forward(inputs, training=True)
    assert layer.output.shape == (1, 2, 5, 5)

    class MockActivation:                      # test_Conv2dBackend.py:46-51 (host activation object)
        def forward(self, inputs):
            self.output = np.maximum(0, np.asarray(inputs))

        def backward(self, dvalues):
            self.dinputs = np.asarray(dvalues) * (self.output > 0)

    layer = Layer_Conv2D(in_channels=3, out_channels=2, kernel_size=3, padding="valid", activation=MockActivation())
    layer.forward(inputs, training=True)
    assert np.all(np.asarray(layer.output) >= 0)


POOL_CASES = [((2, 2), (2, 2), "valid"), ((3, 3), (2, 2), "same"), ((2, 3), (1, 2), "same"),
              ((3, 3), (1, 1), "valid"), ((2, 2), (3, 3), "same")]


@pytest.mark.parametrize("ps,st,pad", POOL_CASES)
def test_maxpool_forward_backward(ps, st, pad):
    from neuromah_b200.src.layers import Layer_MaxPooling2D
    rng = np.random.default_rng(0)
    x = np.round(rng.standard_normal((2, 3, 9, 8)) * 2).astype(np.float32)        # ties on purpose
    layer = Layer_MaxPooling2D(ps, strides=st, padding=pad)
    layer.forward(x, training=True)
    out, idx = R.maxpool2d_forward_cpu(x, *ps, *st, pad)
    np.testing.assert_array_equal(np.asarray(layer.output), out)
    np.testing.assert_array_equal(np.asarray(layer.max_indices), idx)           # index work: bit-exact
    dv = rng.standard_normal(out.shape).astype(np.float32)
    layer.backward(dv)
    close(layer.dinputs, R.maxpool2d_backward_cpu(dv, idx, x, *ps, *st, pad), atol=1e-6)


def test_reference_maxpool_unit_tests_ported():
    """tests/layers/test_maxpool2D.py:14-57 -- the reference's only numeric KATs."""
    from neuromah_b200.src.layers import Layer_MaxPooling2D
    layer = Layer_MaxPooling2D(pool_size=(2, 2), strides=(2, 2), padding="valid")
    inputs = np.arange(1, 17, dtype=np.float32).reshape(1, 1, 4, 4)
    layer.forward(inputs, training=True)
    np.testing.assert_array_equal(np.asarray(layer.output), [[[[6, 8], [14, 16]]]])
    layer.backward(np.array([[[[1, 2], [3, 4]]]], dtype=np.float32))
    np.testing.assert_array_equal(np.asarray(layer.dinputs),
                                  [[[[0, 0, 0, 0], [0, 1, 0, 2], [0, 0, 0, 0], [0, 3, 0, 4]]]])
    layer.forward(np.zeros((1, 1, 4, 4), np.float32))                             # Q8 tie rule
    np.testing.assert_array_equal(np.asarray(layer.max_indices)[0, 0].reshape(4, 2), [[0, 0], [0, 2], [2, 0], [2, 2]])


def test_dense_golden():
    from neuromah_b200.src.layers import Layer_Dense
    from neuromah_b200.src.activations import Activation_ReLU
    g = golden("ops.npz")
    d = Layer_Dense(20, 12, activation=Activation_ReLU(), weight_regularizer_l1=0.01, weight_regularizer_l2=0.02)
    d.set_parameters(g["dense_w"], np.zeros((1, 12)))
    d.forward(g["dense_x"], True)
    close(d.output, g["dense_out"])
    d.backward(g["dense_dv"])
    close(d.dweights, g["dense_dw"]); close(d.dbiases, g["dense_db"]); close(d.dinputs, g["dense_dx"])


@pytest.mark.parametrize("b,n_in,n_out,relu", [(128, 784, 128, True), (128, 128, 10, False), (33, 3136, 10, False),
                                               (7, 19, 5, True), (1, 1, 1, False)])
def test_dense_forward_backward(b, n_in, n_out, relu):
    from neuromah_b200.src.layers import Layer_Dense
    from neuromah_b200.src.activations import Activation_ReLU
    rng = np.random.default_rng(b * 1000 + n_in)
    x = rng.standard_normal((b, n_in)).astype(np.float32)
    w = (rng.standard_normal((n_in, n_out)) * 0.1).astype(np.float32)
    bias = rng.standard_normal((1, n_out)).astype(np.float32)
    d = Layer_Dense(n_in, n_out, activation=Activation_ReLU() if relu else None)
    d.set_parameters(w, bias)
    d.forward(x, True)
    pre = R.dense_forward(x, w.astype(np.float64), bias.astype(np.float64))
    close(d.output, R.relu_forward(pre) if relu else pre, atol=1e-4)
    dv = rng.standard_normal((b, n_out))
    d.backward(dv)
    dx, dw, db = R.dense_backward(x, w.astype(np.float64), R.relu_backward(dv, pre) if relu else dv)
    close(d.dweights, dw, atol=1e-4); close(d.dbiases, db); close(d.dinputs, dx, atol=1e-4)


def test_dense_input_validation():
    from neuromah_b200.src.layers import Layer_Dense
    d = Layer_Dense(4, 3)
    with pytest.raises(TypeError):
        d.forward([[1, 2, 3, 4]], True)
    with pytest.raises(ValueError):
        d.forward(np.zeros((2, 5), np.float32), True)
    with pytest.raises(ValueError):
        d.forward(np.array([[np.nan, 0, 0, 0]], np.float32), True)
    with pytest.raises(ValueError):
        d.set_parameters(np.zeros((4, 4)), np.zeros((1, 3)))


def test_softmax_cce_golden():
    from neuromah_b200.src.activations import Activation_Softmax, Activation_ReLU
    from neuromah_b200.src.losses import Loss_CategoricalCrossentropy
    from neuromah_b200.src.losses.Activation_Softmax_Loss_CategoricalCrossentropy import \
        Activation_Softmax_Loss_CategoricalCrossentropy
    g = golden("ops.npz")
    sm = Activation_Softmax()
    sm.forward(g["sm_in"].astype(np.float32))
    close(sm.output, g["sm_out"])
    assert np.array_equal(sm.predictions(sm.output), np.argmax(g["sm_out"], axis=1))
    loss = Loss_CategoricalCrossentropy(model=None)
    close(loss.forward(sm.output, g["labels"]), g["cce"])
    close(loss.forward(sm.output, np.eye(10)[g["labels"]]), g["cce_onehot"])
    fused = Activation_Softmax_Loss_CategoricalCrossentropy()
    fused.backward(sm.output, np.eye(10)[g["labels"]])
    close(fused.dinputs, g["fused_grad"])
    loss.backward(sm.output, g["labels"])
    close(loss.dinputs, g["cce_bwd"], rtol=1e-3)
    sm.backward(g["sm_dv"])
    close(sm.dinputs, g["sm_bwd"])
    one = Activation_Softmax()
    one.forward(np.eye(10, dtype=np.float32)[[3]])                   # notebook fingerprint (test.ipynb cell 18)
    p = np.asarray(one.output)
    assert abs(p[0, 3] - 0.23196932) < 1e-6 and abs(p[0, 0] - 0.08533674) < 1e-6
    r = Activation_ReLU()
    xr = np.array([[-1.0, 0.0, 2.0, -0.0, 5.0]], np.float32)
    r.forward(xr); r.backward(np.ones_like(xr))
    np.testing.assert_array_equal(np.asarray(r.output), [[0, 0, 2, 0, 5]])
    np.testing.assert_array_equal(np.asarray(r.dinputs), [[0, 0, 1, 0, 1]])


def test_cce_clip_and_accuracy_flags():
    from neuromah_b200 import ops
    from neuromah_b200.device import DeviceArray, labels_to_device
    p = np.array([[1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.2, 0.4, 0.4], [0.5, 0.5, 0.0]], np.float32)
    y = np.array([0, 0, 2, 1])
    accum = DeviceArray.zeros((2,), np.float64)
    losses, correct, _ = ops.softmax_ce(DeviceArray.from_numpy(p), labels_to_device(y), want_grad=False, accum=accum)
    close(losses, R.cce_forward(p.astype(np.float64), y), rtol=1e-5)
    np.testing.assert_array_equal(np.asarray(correct), (np.argmax(p, axis=1) == y).astype(np.int32))  # first-max ties
    a = np.asarray(accum)
    close(a[0], R.cce_forward(p.astype(np.float64), y).sum(), rtol=1e-5)
    assert a[1] == float((np.argmax(p, axis=1) == y).sum())


def test_adam_double_application_golden():
    """Adam.py:86-100 applied twice with the same t, decay 0.1, three iterations."""
    from neuromah_b200.device import DeviceArray
    from neuromah_b200.src.optimizers import Optimizer_Adam
    g = golden("ops.npz")

    class L:
        def __init__(s):
            s.w = DeviceArray.from_numpy(g["adam_w0"], np.float32); s.g = DeviceArray.zeros((5, 4))

        def get_parameters(s):
            return {"weights": (s.w, s.g)}
    lay = L()
    opt = Optimizer_Adam(learning_rate=0.01, decay=0.1)
    for it in range(3):
        lay.g.copy_from(g["adam_g"][it])
        opt.pre_update_params(); opt.update_params(lay); opt.update_params(lay); opt.post_update_params()
        close(lay.w, g[f"adam_w{it + 1}"])
    assert opt.iterations == 3 and abs(opt.current_learning_rate - 0.01 / 1.2) < 1e-12
    # fused n_apply=2 in one launch == two launches
    from neuromah_b200 import ops
    w2 = DeviceArray.from_numpy(g["adam_w0"], np.float32); m = DeviceArray.zeros((5, 4)); v = DeviceArray.zeros((5, 4))
    for it in range(3):
        lay.g.copy_from(g["adam_g"][it])
        ops.adam_step(w2, lay.g, m, v, lr=0.01 / (1 + 0.1 * it), beta_1=0.9, beta_2=0.999, eps=1e-7, t=it, n_apply=2)
        close(w2, g[f"adam_w{it + 1}"])


def test_adam_large_slab_vectorised_and_tail():
    from neuromah_b200 import ops
    from neuromah_b200.device import DeviceArray
    rng = np.random.default_rng(0)
    n = 1_000_003                                            # not a multiple of 4: exercises the scalar tail
    w = rng.standard_normal(n).astype(np.float32); g = (rng.standard_normal(n) * 1e-3).astype(np.float32)
    dw, dg = DeviceArray.from_numpy(w), DeviceArray.from_numpy(g)
    m, v = DeviceArray.zeros((n,)), DeviceArray.zeros((n,))
    ops.adam_step(dw, dg, m, v, lr=1e-3, beta_1=0.9, beta_2=0.999, eps=1e-7, t=0, n_apply=2)
    rw, rm, rv = w.astype(np.float64), 0.0, 0.0
    for _ in range(2):
        rw, rm, rv = R.adam_update(rw, g.astype(np.float64), rm, rv, lr=1e-3, t=0)
    close(dw, rw); close(m, rm, atol=1e-7); close(v, rv, atol=1e-9)


def test_sgd_plain_momentum_and_nonfinite():
    from neuromah_b200.device import DeviceArray
    from neuromah_b200.src.optimizers import Optimizer_SGD
    rng = np.random.default_rng(1)
    w0 = rng.standard_normal((6, 3)); gs = rng.standard_normal((3, 6, 3))

    class L:
        def __init__(s):
            s.w = DeviceArray.from_numpy(w0, np.float32); s.g = DeviceArray.zeros((6, 3))

        def get_parameters(s):
            return {"weights": (s.w, s.g)}
    for mu in (0.0, 0.9):
        lay, opt = L(), Optimizer_SGD(learning_rate=0.1, decay=0.05, momentum=mu)
        rw, rm = w0.copy(), np.zeros_like(w0)
        for it in range(3):
            lay.g.copy_from(gs[it])
            opt.pre_update_params(); opt.update_params(lay); opt.post_update_params()
            rw, rm = R.sgd_update(rw, gs[it], rm, lr=R.decayed_lr(0.1, 0.05, it), momentum=mu)
            close(lay.w, rw)
    lay, opt = L(), Optimizer_SGD(learning_rate=0.1)
    bad = gs[0].copy(); bad[2, 1] = np.nan
    lay.g.copy_from(bad)
    with pytest.raises(ValueError):                           # SGD.py:80-83
        opt.update_params(lay)


def test_bad_arguments_surface_as_python_exceptions():
    from neuromah_b200._lib import lib
    from neuromah_b200.device import DeviceArray, stream
    a = DeviceArray.zeros((4,))
    with pytest.raises(ValueError):
        lib.nm_adam_step_f32(a.p, a.p, a.p, a.p, 4, 0.1, 0.9, 0.999, 1e-7, 0.1, 0.001, 0, stream())   # n_apply = 0
    with pytest.raises(ValueError):
        lib.nm_conv2d_fprop_f32(a.p, a.p, None, a.p, 1, 1, 2, 2, 1, 3, 3, 0, 0, 0, 0, 0, stream())   # kernel > input
