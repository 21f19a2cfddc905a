"""GPU parity, model level: the mirror Model / layers / optimizers stepped exactly like the
reference's Model.train (Model.py:154-174) against golden vectors written by the UNMODIFIED
reference and against the oracle.  FP32-exact bar: rtol 1e-4 / atol 1e-5."""
import numpy as np
import pytest

from conftest import golden, synth
from oracle import ref_net

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-4, 1e-5


def build(spec, params, optimizer="adam", mode=None, **opt_kw):
    """Build a neuromah_b200 Model from an oracle spec with the given initial parameters."""
    from neuromah_b200.src.core import Model
    from neuromah_b200.src.layers import Layer_Conv2D, Layer_MaxPooling2D, Layer_Dense, Layer_Flatten
    from neuromah_b200.src.activations import Activation_ReLU, Activation_Softmax
    from neuromah_b200.src.losses import Loss_CategoricalCrossentropy
    from neuromah_b200.src.optimizers import Optimizer_Adam, Optimizer_SGD
    from neuromah_b200.src.metrics import Accuracy_Categorical
    model = Model()
    for L, p in zip(spec, params):
        k = L["kind"]
        if k == "conv":
            layer = Layer_Conv2D(L["ic"], L["oc"], L["k"], padding=L["padding"],
                                 activation=Activation_ReLU() if L.get("relu") else None)
        elif k == "pool":
            layer = Layer_MaxPooling2D(L["size"], strides=L.get("stride"), padding=L.get("padding", "valid"))
        elif k == "flatten":
            layer = Layer_Flatten()
        elif k == "dense":
            layer = Layer_Dense(L["n_in"], L["n_out"], activation=Activation_ReLU() if L.get("relu") else None,
                                weight_regularizer_l1=L.get("l1", 0), weight_regularizer_l2=L.get("l2", 0))
        else:
            layer = Activation_Softmax()
        if p is not None:
            layer.set_parameters(p["weights"], p["biases"])
        model.add(layer)
    opt = Optimizer_Adam(**opt_kw) if optimizer == "adam" else Optimizer_SGD(**opt_kw)
    model.set(loss=Loss_CategoricalCrossentropy, optimizer=opt, accuracy=Accuracy_Categorical)
    model.finalize(mode=mode)
    return model


def reference_style_step(model, x, y):
    """Model.train's loop body, call for call (Model.py:154-174), through the public API."""
    output = model.forward(x, training=True)
    losses = model.loss.calculate(output, y, include_regularization=True)
    predictions = model.output_layer_activation.predictions(output)
    acc = model.accuracy.calculate(predictions, y)
    model.backward(output, y)
    model.optimizer.pre_update_params()
    for layer in model.trainable_layers:
        model.optimizer.update_params(layer)
    model.optimizer.post_update_params()
    return float(losses["data_loss"]), float(acc)


def test_finalize_registers_trainable_layers_twice():
    spec = ref_net.mnist_cnn_spec()
    model = build(spec, ref_net.init_params(spec, np.random.RandomState(0)))
    assert len(model.trainable_layers) == 6 and len({id(l) for l in model.trainable_layers}) == 3   # Q1
    assert model.softmax_classifier_output is not None
    assert model.arena.n_params == 50186


def test_mlp_200_step_loss_curve_golden():
    """BASELINE.json configs[0]: 784-128-10 MLP, batch 128, Adam 1e-3, 200 steps, reference-style loop."""
    g = golden("mlp_200steps.npz")
    spec = ref_net.mlp_spec()
    model = build(spec, ref_net.init_params(spec, np.random.RandomState(0)), learning_rate=0.001)
    model.loss.new_pass(); model.accuracy.new_pass()
    x, y = synth(1, 128 * 200, (784,))
    losses, accs = zip(*[reference_style_step(model, x[s * 128:(s + 1) * 128], y[s * 128:(s + 1) * 128])
                         for s in range(200)])
    np.testing.assert_allclose(losses, g["losses"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(accs, g["accs"], atol=2.0 / 128)
    np.testing.assert_allclose(model.arena.host_params(), g["final_params"], rtol=RTOL, atol=ATOL)


def test_mlp_train_api_fused_path_matches_golden():
    """Same run through model.train (device-resident data, arena optimizer, device loss sums)."""
    g = golden("mlp_200steps.npz")
    spec = ref_net.mlp_spec()
    model = build(spec, ref_net.init_params(spec, np.random.RandomState(0)), learning_rate=0.001)
    x, y = synth(1, 128 * 200, (784,))
    model.train(x, y, epochs=1, batch_size=128, verbose=0)
    np.testing.assert_allclose(model.arena.host_params(), g["final_params"], rtol=RTOL, atol=ATOL)
    assert abs(model.last_epoch["data_loss"] - float(np.mean(g["losses"]))) < 1e-4
    assert model.optimizer.iterations == 200


def test_mlp_sgd_golden():
    g = golden("mlp_sgd_20steps.npz")
    spec = ref_net.mlp_spec(64, 32, 10)
    model = build(spec, ref_net.init_params(spec, np.random.RandomState(5)), optimizer="sgd",
                  learning_rate=0.5, decay=0.01)
    model.loss.new_pass(); model.accuracy.new_pass()
    x, y = synth(6, 32 * 20, (64,))
    losses = [reference_style_step(model, x[s * 32:(s + 1) * 32], y[s * 32:(s + 1) * 32])[0] for s in range(20)]
    np.testing.assert_allclose(losses, g["losses"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(model.arena.host_params(), g["final_params"], rtol=RTOL, atol=ATOL)


def test_cnn_per_layer_golden():
    """MNIST CNN (examples/test_Conv2D_MNIST.py:47-53), batch 4, 3 Adam steps: every layer's output,
    every dinputs, the gradient slab and the post-update weights per step."""
    g = golden("cnn_small_3steps.npz")
    spec = ref_net.mnist_cnn_spec()
    model = build(spec, ref_net.init_params(spec, np.random.RandomState(2)), learning_rate=0.001)
    model.layers[0].compute_dinputs = True                  # the golden run recorded the first dgrad too
    model.loss.new_pass(); model.accuracy.new_pass()
    x, y = synth(3, 12, (1, 28, 28))
    for s in range(3):
        loss, _ = reference_style_step(model, x[s * 4:(s + 1) * 4], y[s * 4:(s + 1) * 4])
        assert abs(loss - float(g[f"loss{s}"])) < 1e-5
        if s == 0:
            for i, layer in enumerate(model.layers):
                np.testing.assert_allclose(np.asarray(layer.output), g[f"out{i}"], rtol=RTOL, atol=ATOL, err_msg=f"out{i}")
                if f"din{i}" in g.files:
                    np.testing.assert_allclose(np.asarray(layer.dinputs), g[f"din{i}"], rtol=RTOL, atol=ATOL,
                                               err_msg=f"din{i}")
            np.testing.assert_allclose(model.arena.host_grads(), g["grads0"], rtol=RTOL, atol=1e-6)
        np.testing.assert_allclose(model.arena.host_params(), g[f"params{s}"], rtol=RTOL, atol=ATOL, err_msg=f"step {s}")


def test_cnn_stub_mode_golden():
    """The reference AS SHIPPED (conv backward stub): conv weights never move."""
    from neuromah_b200 import config
    g = golden("cnn_stub_2steps.npz")
    spec = ref_net.mnist_cnn_spec()
    params = ref_net.init_params(spec, np.random.RandomState(2))
    config.conv_backward = "stub"
    try:
        model = build(spec, params, learning_rate=0.001)
        model.loss.new_pass(); model.accuracy.new_pass()
        x, y = synth(3, 8, (1, 28, 28))
        for s in range(2):
            loss, _ = reference_style_step(model, x[s * 4:(s + 1) * 4], y[s * 4:(s + 1) * 4])
            assert abs(loss - float(g[f"loss{s}"])) < 1e-5
    finally:
        config.conv_backward = "wired"
    np.testing.assert_allclose(model.arena.host_params(), g["final_params"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(np.asarray(model.layers[0].weights), params[0]["weights"], rtol=0, atol=1e-7)


def _resync_oracle(net, model):
    """Copy the device state (weights + Adam moments + step count) into the oracle so that the next step
    starts from IDENTICAL state on both sides: per-step parity without the chaotic divergence that ReLU /
    arg-max flips and Adam's sign-like update (g/(|g|+1e-7)) produce once two float32 runs drift apart."""
    net.set_flat(model.arena.host_params().astype(np.float64))
    m = model.arena.state[0].numpy().astype(np.float64)
    v = model.arena.state[1].numpy().astype(np.float64)
    for layer, name, off, size, shape in model.arena.table:
        i = [id(l) for l in model.layers].index(id(layer))
        net.state[i][name] = (m[off:off + size].reshape(shape), v[off:off + size].reshape(shape))
    net.iterations = model.optimizer.iterations


def test_cnn_per_step_parity_batch_256_with_state_resync():
    """Four consecutive Adam steps at batch 256: before each step the oracle takes the device's state,
    then both step once; loss, gradient slab and post-update weights must agree per step."""
    spec = ref_net.mnist_cnn_spec()
    params = ref_net.init_params(spec, np.random.RandomState(4))
    model = build(spec, params, learning_rate=0.001)
    net = ref_net.RefNet(spec, params)
    x, y = synth(5, 1024, (1, 28, 28))
    for s in range(4):
        _resync_oracle(net, model)
        xb, yb = x[s * 256:(s + 1) * 256], y[s * 256:(s + 1) * 256]
        out = model.train_step(xb, yb)
        r = net.step(xb, yb)
        np.testing.assert_allclose(np.asarray(out), r["output"], rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(model.arena.host_grads(), net.flat("grads"), rtol=1e-3, atol=1e-7, err_msg=f"grads step {s}")
        np.testing.assert_allclose(model.arena.host_params(), net.flat(), rtol=RTOL, atol=ATOL, err_msg=f"weights step {s}")


def test_cnn_train_api_vs_oracle_batch_256():
    """model.train over 2 epochs x 2 steps of 256 vs the oracle driver: loss sums, accuracy, predict/evaluate.
    Free-running weights are compared with a drift budget (see _resync_oracle for why): almost all elements
    within the FP32-exact bar, none further than a few Adam steps (lr = 1e-3)."""
    spec = ref_net.mnist_cnn_spec()
    params = ref_net.init_params(spec, np.random.RandomState(4))
    model = build(spec, params, learning_rate=0.001)
    net = ref_net.RefNet(spec, params)
    x, y = synth(5, 512, (1, 28, 28))
    xv, yv = synth(6, 100, (1, 28, 28))
    model.train(x, y, epochs=2, batch_size=256, verbose=0, validation_data=(xv, yv))
    for _ in range(2):
        ls = [net.step(x[s * 256:(s + 1) * 256], y[s * 256:(s + 1) * 256])["loss_sum"] for s in range(2)]
    diff = np.abs(model.arena.host_params() - net.flat())
    assert np.mean(diff > ATOL + RTOL * np.abs(net.flat())) < 0.05 and diff.max() < 2e-3
    assert abs(model.last_epoch["data_loss"] - sum(ls) / 512) < 1e-4
    out = net.forward(x, training=False)
    assert abs(model.last_epoch["acc"] - float(np.mean(np.argmax(out, 1) == np.argmax(y, 1)))) <= 4 / 512
    # predict / evaluate against the oracle evaluated at the DEVICE's weights
    net.set_flat(model.arena.host_params().astype(np.float64))
    pred = model.predict(xv, batch_size=64)
    ref_pred = net.forward(xv, training=False)
    np.testing.assert_allclose(pred, ref_pred, rtol=RTOL, atol=ATOL)
    ev = model.evaluate(xv, yv, batch_size=32, verbose=0)
    assert abs(ev["loss"] - float(np.mean(-np.log(np.clip(np.sum(ref_pred * yv, 1), 1e-7, 1 - 1e-7))))) < 1e-4
    assert abs(ev["acc"] - float(np.mean(np.argmax(ref_pred, 1) == np.argmax(yv, 1)))) <= 1 / 100


def test_virtual_shards_sum_to_full_batch_gradient():
    """SURVEY 8e without NCCL: G slices with the GLOBAL batch as divisor, gradient slabs summed on the
    host, equal the un-split batch-B gradient (linearity property; full config-2 batch 4096)."""
    from neuromah_b200 import parallel
    spec = ref_net.mnist_cnn_spec()
    params = ref_net.init_params(spec, np.random.RandomState(7))
    B, G = 4096, 8
    x, y = synth(8, B, (1, 28, 28))
    full = build(spec, params)
    out = full.forward(x, training=True); full.backward(out, y)
    g_full = full.arena.host_grads().astype(np.float64)
    shard = build(spec, params)
    parallel.set_global_batch(shard, B)
    acc = np.zeros_like(g_full)
    for r in range(G):
        lo, hi = parallel.shard_bounds(B, r, G)
        o = shard.forward(x[lo:hi], training=True); shard.backward(o, y[lo:hi])
        acc += shard.arena.host_grads()
    np.testing.assert_allclose(acc, g_full, rtol=1e-3, atol=1e-6)


def test_full_batch_4096_step_properties():
    """Config-2 size (batch 4096): post-ReLU outputs >= 0, pooled = max of its window, softmax rows sum to 1,
    and the bias-gradient identity dbias = sum(dvalues) holds through the fused ones-column."""
    spec = ref_net.mnist_cnn_spec()
    model = build(spec, ref_net.init_params(spec, np.random.RandomState(1)))
    x, y = synth(9, 4096, (1, 28, 28))
    out = model.forward(x, training=True)
    p = np.asarray(out)
    np.testing.assert_allclose(p.sum(axis=1), 1.0, atol=1e-5)
    c1, p1 = np.asarray(model.layers[0].output), np.asarray(model.layers[1].output)
    assert c1.min() >= 0
    np.testing.assert_array_equal(p1, c1.reshape(4096, 32, 14, 2, 14, 2).max(axis=(3, 5)))
    model.backward(out, y)
    d2 = np.asarray(model.layers[2].activation.dinputs, dtype=np.float64)
    np.testing.assert_allclose(np.asarray(model.layers[2].bias_gradients)[:, 0], d2.sum(axis=(0, 2, 3)), rtol=1e-3, atol=1e-6)


def test_save_and_load_parameters_roundtrip(tmp_path):
    spec = ref_net.mlp_spec(16, 8, 4)
    m1 = build(spec, ref_net.init_params(spec, np.random.RandomState(0)))
    m2 = build(spec, ref_net.init_params(spec, np.random.RandomState(1)))
    path = str(tmp_path / "params.pkl")
    m1.save_parameters(path)
    m2.load_parameters(path)
    np.testing.assert_array_equal(m1.arena.host_params(), m2.arena.host_params())


@pytest.mark.parametrize("case", ["mlp_dropout_rmsprop", "cnn_bf16_adam", "cnn_bf16x3_adam"])
def test_save_load_whole_model_resumes_training_identically(tmp_path, case):
    """Model.save / Model.load (Model.py:396-417): the reloaded model carries parameters, optimizer state and iteration
    count, so continuing training on both gives the same parameters."""
    from neuromah_b200.src.core import Model
    from neuromah_b200.src.layers import Layer_Dense, Layer_Dropout
    from neuromah_b200.src.activations import Activation_ReLU, Activation_Softmax
    from neuromah_b200.src.losses import Loss_CategoricalCrossentropy
    from neuromah_b200.src.optimizers import Optimizer_RMSprop
    from neuromah_b200.src.metrics import Accuracy_Categorical
    if case == "mlp_dropout_rmsprop":
        m1 = Model()
        for l in (Layer_Dense(20, 16, activation=Activation_ReLU(), weight_regularizer_l2=1e-3), Layer_Dropout(0.25, seed=9),
                  Layer_Dense(16, 10), Activation_Softmax()):
            m1.add(l)
        m1.set(loss=Loss_CategoricalCrossentropy, optimizer=Optimizer_RMSprop(learning_rate=0.01, decay=1e-2),
               accuracy=Accuracy_Categorical)
        m1.finalize()
        x, y = synth(50, 32 * 6, (20,))
        B = 32
    else:
        spec = ref_net.mnist_cnn_spec()
        m1 = build(spec, ref_net.init_params(spec, np.random.RandomState(12)), mode="bf16x3" if "x3" in case else "bf16",
                   learning_rate=0.002, decay=1e-3)
        x, y = synth(51, 16 * 6, (1, 28, 28))
        B = 16
    for s in range(3):
        m1.train_step(x[s * B:(s + 1) * B], y[s * B:(s + 1) * B])
    path = str(tmp_path / "model.pkl")
    m1.save(path)
    m2 = Model.load(path)
    assert m2.mode == m1.mode and m2.optimizer.iterations == 3 and type(m2.optimizer) is type(m1.optimizer)
    assert [type(l).__name__ for l in m2.layers] == [type(l).__name__ for l in m1.layers]
    np.testing.assert_array_equal(m1.arena.host_params(), m2.arena.host_params())
    for a, b in zip(m1.arena.state, m2.arena.state):
        np.testing.assert_array_equal(np.asarray(a), np.asarray(b))
    for s in range(3, 6):
        m1.train_step(x[s * B:(s + 1) * B], y[s * B:(s + 1) * B])
        m2.train_step(x[s * B:(s + 1) * B], y[s * B:(s + 1) * B])
    np.testing.assert_allclose(m1.arena.host_params(), m2.arena.host_params(), rtol=1e-6, atol=1e-7)
    assert m2.optimizer.current_learning_rate == pytest.approx(m1.optimizer.current_learning_rate)


def test_vgg_style_stack_per_step_parity():
    """BASELINE.json config 4 (VGG-style Conv2D stack on CIFAR-shaped 32x32x3 data, He init), reduced in width so the
    NumPy oracle finishes in seconds: C(3->8) C(8->8) P C(8->16) C(16->16) P Flatten Dense(1024->10) Softmax, FP32-exact
    kernels, per-step outputs / gradients / post-Adam weights within the north-star tolerance."""
    spec = []
    ic = 3
    for oc in (8, 16):
        spec += [{"kind": "conv", "ic": ic, "oc": oc, "k": 3, "padding": "same", "relu": True},
                 {"kind": "conv", "ic": oc, "oc": oc, "k": 3, "padding": "same", "relu": True},
                 {"kind": "pool", "size": 2, "stride": 2, "padding": "valid"}]
        ic = oc
    spec += [{"kind": "flatten"}, {"kind": "dense", "n_in": 16 * 8 * 8, "n_out": 10, "relu": False}, {"kind": "softmax"}]
    params = ref_net.init_params(spec, np.random.RandomState(31), init="he")
    model = build(spec, params, learning_rate=0.001)
    net = ref_net.RefNet(spec, params)
    x, y = synth(32, 3 * 48, (3, 32, 32))
    for s in range(3):
        _resync_oracle(net, model)
        xb, yb = x[s * 48:(s + 1) * 48], y[s * 48:(s + 1) * 48]
        out = model.train_step(xb, yb)
        ref = net.step(xb, yb)
        np.testing.assert_allclose(np.asarray(out), ref["output"], rtol=1e-4, atol=1e-5)
        g_ref = net.flat("grads")
        np.testing.assert_allclose(model.arena.host_grads(), g_ref, rtol=1e-4, atol=1e-5 * max(1.0, np.abs(g_ref).max()))
        np.testing.assert_allclose(model.arena.host_params(), net.flat(), rtol=1e-4, atol=1e-5)
