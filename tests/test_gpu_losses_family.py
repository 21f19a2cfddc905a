"""GPU parity for the other output activations / losses a Model.train step can end in -- Activation_Sigmoid /
Activation_Linear, Loss_MeanSquaredError / MeanAbsoluteError / BinaryCrossentropy, Accuracy_Regression /
Accuracy_Categorical(binary=True) -- through the C ABI (nm_sigmoid_*_f32, nm_loss_*_f32) against vectors the UNMODIFIED
reference produced (tests/golden/losses_family.npz) and against the oracle on larger / ragged shapes.
Tolerance: the FP32-exact bar, rtol 1e-4 / atol 1e-5."""
import numpy as np
import pytest

from conftest import LOSSES_FAMILY_CASES, golden, losses_family_case, losses_family_inputs
from oracle import np_ref as R

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-4, 1e-5


def close(a, b, rtol=RTOL, atol=ATOL):
    np.testing.assert_allclose(np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64), rtol=rtol, atol=atol)


def _loss_classes():
    from neuromah_b200.src.losses import Loss_BinaryCrossentropy, Loss_MeanAbsoluteError, Loss_MeanSquaredError
    return {"mse": Loss_MeanSquaredError, "mae": Loss_MeanAbsoluteError, "bce": Loss_BinaryCrossentropy}


def test_sigmoid_and_linear_against_the_reference_vectors():
    from neuromah_b200.src.activations import Activation_Linear, Activation_Sigmoid
    g, v = golden("losses_family.npz"), losses_family_inputs()
    sg = Activation_Sigmoid()
    sg.forward(v["x"], training=True)
    sg.backward(v["dv"])
    close(sg.output, g["sigmoid_out"])
    close(sg.dinputs, g["sigmoid_dx"])
    np.testing.assert_array_equal(sg.predictions(sg.output), g["sigmoid_pred"])
    lin = Activation_Linear()
    lin.forward(v["x"], training=True)
    lin.backward(v["dv"])
    np.testing.assert_array_equal(np.asarray(lin.output), v["x"])
    np.testing.assert_array_equal(np.asarray(lin.dinputs), v["dv"])
    np.testing.assert_array_equal(np.asarray(lin.predictions(lin.output)), v["x"])


def test_sigmoid_saturates_like_numpy_and_handles_ragged_sizes():
    from neuromah_b200.src.activations import Activation_Sigmoid
    rng = np.random.default_rng(3)
    x = (rng.standard_normal((257, 1031)) * 8).astype(np.float32)       # odd sizes, more elements than one grid pass
    x[0, :4] = [-200.0, 200.0, 0.0, -88.0]
    dv = rng.standard_normal(x.shape).astype(np.float32)
    sg = Activation_Sigmoid()
    sg.forward(x, training=False)
    with np.errstate(over="ignore"):
        ref = R.sigmoid_forward(x.astype(np.float64))
    close(sg.output, ref)
    out = np.asarray(sg.output)
    assert out[0, 0] == 0.0 and out[0, 1] == 1.0 and out[0, 2] == 0.5 and np.all(np.isfinite(out))
    sg.backward(dv)
    close(sg.dinputs, R.sigmoid_backward(out.astype(np.float64), dv.astype(np.float64)))


@pytest.mark.parametrize("name,target", [("mse", "t_reg"), ("mae", "t_reg"), ("bce", "t_bin")])
def test_losses_against_the_reference_vectors(name, target):
    g, v = golden("losses_family.npz"), losses_family_inputs()
    loss = _loss_classes()[name](model=None)
    close(loss.forward(v["pred"], v[target]), g[f"{name}_fwd"])
    loss.backward(v["pred"], v[target])
    close(loss.dinputs, g[f"{name}_bwd"])
    out4 = loss.forward(v["pred4"], v["t4"])                  # every non-batch axis is averaged
    assert out4.shape == (6,)
    close(out4, g[f"{name}_fwd4"])
    loss.backward(v["pred4"], v["t4"])
    assert loss.dinputs.shape == (6, 2, 3, 4)
    close(loss.dinputs, g[f"{name}_bwd4"])


@pytest.mark.parametrize("name", ["mse", "mae", "bce"])
@pytest.mark.parametrize("shape", [(1, 1), (5, 1), (3, 33), (130, 7), (1000, 100), (2, 4097)])
def test_losses_against_the_oracle_on_ragged_shapes(name, shape):
    rng = np.random.default_rng(shape[0] * 7919 + shape[1] + 104729 * ["mse", "mae", "bce"].index(name))
    pred = rng.random(shape, dtype=np.float32)
    t = (rng.random(shape) > 0.5).astype(np.float32) if name == "bce" else rng.standard_normal(shape).astype(np.float32)
    if name == "mae" and pred.size > 2:
        t.flat[1] = pred.flat[1]                               # np.sign(0) = 0
    loss = _loss_classes()[name](model=None)
    p64, t64 = pred.astype(np.float64), t.astype(np.float64)
    close(loss.forward(pred, t), getattr(R, f"{name}_forward")(p64, t64))
    loss.backward(pred, t)
    close(loss.dinputs, getattr(R, f"{name}_backward")(p64, t64))


def test_bce_clips_saturated_predictions():
    from neuromah_b200.src.losses import Loss_BinaryCrossentropy
    pred = np.array([[0.0, 1.0], [1.0, 0.0], [0.5, 0.25]], dtype=np.float32)
    t = np.array([[1.0, 0.0], [1.0, 0.0], [1.0, 0.0]], dtype=np.float32)
    loss = Loss_BinaryCrossentropy(model=None)
    p64, t64 = pred.astype(np.float64), t.astype(np.float64)
    fwd = np.asarray(loss.forward(pred, t))
    assert np.all(np.isfinite(fwd))
    # the clip bound 1 - 1e-7 is not a float32 number: the saturated 'right answer' row may differ by 2e-8 absolute
    close(fwd, R.bce_forward(p64, t64), rtol=1e-4, atol=1e-7)
    loss.backward(pred, t)
    close(loss.dinputs, R.bce_backward(p64, t64), rtol=1e-4, atol=1e-5)


def test_loss_shape_mismatch_raises():
    from neuromah_b200.src.losses import Loss_MeanSquaredError
    loss = Loss_MeanSquaredError(model=None)
    with pytest.raises(ValueError):
        loss.forward(np.zeros((4, 3), np.float32), np.zeros((4, 2), np.float32))
    with pytest.raises(ValueError):
        loss.backward(np.zeros((4, 3), np.float32), np.zeros((5, 3), np.float32))


def _model(name):
    from neuromah_b200.src.core import Model
    from neuromah_b200.src.layers import Layer_Dense
    from neuromah_b200.src.activations import Activation_Linear, Activation_ReLU, Activation_Sigmoid
    from neuromah_b200.src.optimizers import Optimizer_Adam, Optimizer_SGD
    from neuromah_b200.src.metrics import Accuracy_Categorical, Accuracy_Regression
    c = LOSSES_FAMILY_CASES[name]
    w, x, y = losses_family_case(name)
    model = Model()
    d1 = Layer_Dense(8, 16, activation=Activation_ReLU())
    d2 = Layer_Dense(16, c["n_out"])
    for d, wi in zip((d1, d2), w):
        d.set_parameters(wi, np.zeros((1, wi.shape[1])))
    for l in (d1, d2, Activation_Sigmoid() if c["act"] == "sigmoid" else Activation_Linear()):
        model.add(l)
    opt = Optimizer_Adam(learning_rate=c["lr"]) if c["opt"] == "adam" else Optimizer_SGD(learning_rate=c["lr"])
    acc = Accuracy_Categorical(binary=True, model=model) if c["acc"] == "binary" else Accuracy_Regression(model=model)
    model.set(loss=_loss_classes()[c["loss"]](model=model), optimizer=opt, accuracy=acc)
    model.finalize()
    return model, c, x, y


@pytest.mark.parametrize("name", sorted(LOSSES_FAMILY_CASES))
def test_model_steps_reproduce_the_reference_runs(name):
    """Dense(8,16,ReLU) -> Dense(16,n) -> Linear | Sigmoid with MSE / MAE / BCE, stepped through Model.train_step: loss per
    step, post-update parameters and the inference output of the unmodified reference."""
    g = golden("losses_family.npz")
    model, c, x, y = _model(name)
    assert model.softmax_classifier_output is None                       # the generic loss.backward path (Model.py:362-366)
    model.accuracy.init(y)
    if c["acc"] == "regression":
        assert model.accuracy.precision == pytest.approx(float(g[f"{name}_precision"]), rel=1e-12)
    model.loss.new_pass(); model.accuracy.new_pass()
    losses, accs = [], []
    for s in range(c["steps"]):
        model.train_step(x[s * 32:(s + 1) * 32], y[s * 32:(s + 1) * 32])
        losses.append(model.last_step["loss"]); accs.append(model.last_step["acc"])
    np.testing.assert_allclose(losses, g[f"{name}_losses"], rtol=RTOL, atol=ATOL)
    # one hit more or less out of 32 x n_out comparisons when a prediction lies within FP32 rounding of the threshold
    np.testing.assert_allclose(accs, g[f"{name}_accs"], atol=1.01 / (32 * c["n_out"]))
    np.testing.assert_allclose(model.arena.host_params(), g[f"{name}_final_params"], rtol=RTOL, atol=2 * ATOL)
    close(model.forward(x[:32], training=False), g[f"{name}_infer"], rtol=RTOL, atol=2 * ATOL)
    assert model.accuracy.accumulated_count == g[f"{name}_acc_running"][1]          # samples, not elements
    assert abs(model.accuracy.accumulated_sum - g[f"{name}_acc_running"][0]) <= c["steps"]


def test_train_evaluate_predict_with_a_regression_head():
    """Model.train / evaluate / predict take the generic loop for these losses (Model.py:114-317): the MSE falls."""
    model, c, x, y = _model("mse")
    model.accuracy.init(y)            # evaluate() before train(): the reference's evaluate does not initialise the metric
    first = model.evaluate(x, y, batch_size=64, verbose=0)["loss"]
    model.train(x, y, epochs=30, batch_size=32, verbose=0, validation_data=(x[:64], y[:64]))
    last = model.evaluate(x, y, batch_size=64, verbose=0)
    assert last["loss"] < 0.5 * first
    pred = model.predict(x[:40], batch_size=16)
    assert pred.shape == (40, 3)
    close(np.mean((pred - y[:40]) ** 2, axis=1), np.asarray(model.loss.forward(pred.astype(np.float32), y[:40])))


@pytest.mark.parametrize("name", ["mse", "bce"])
def test_save_load_resumes_training_identically(tmp_path, name):
    """Model.save / Model.load with a Linear / Sigmoid head: layers, loss and metric classes, the regression precision,
    parameters and optimizer state survive; both models then take the same steps."""
    from neuromah_b200.src.core import Model
    m1, c, x, y = _model(name)
    m1.accuracy.init(y)
    m1.loss.new_pass(); m1.accuracy.new_pass()
    for s in range(3):
        m1.train_step(x[s * 32:(s + 1) * 32], y[s * 32:(s + 1) * 32])
    path = str(tmp_path / "model.pkl")
    m1.save(path)
    m2 = Model.load(path)
    assert [type(l).__name__ for l in m2.layers] == [type(l).__name__ for l in m1.layers]
    assert type(m2.loss) is type(m1.loss) and type(m2.accuracy) is type(m1.accuracy)
    assert m2.optimizer.iterations == 3 and m2.softmax_classifier_output is None
    if name == "mse":
        assert m2.accuracy.precision == m1.accuracy.precision
    else:
        assert m2.accuracy.binary is True
    np.testing.assert_array_equal(m1.arena.host_params(), m2.arena.host_params())
    m2.loss.new_pass(); m2.accuracy.new_pass()
    for s in range(3, 6):
        m1.train_step(x[s * 32:(s + 1) * 32], y[s * 32:(s + 1) * 32])
        m2.train_step(x[s * 32:(s + 1) * 32], y[s * 32:(s + 1) * 32])
        assert m2.last_step["loss"] == pytest.approx(m1.last_step["loss"], rel=1e-6)
    np.testing.assert_allclose(m1.arena.host_params(), m2.arena.host_params(), rtol=1e-6, atol=1e-7)


def test_dense_with_an_embedded_sigmoid():
    """Layer_Dense(activation=Activation_Sigmoid()) (Dense.py:82-85, :90-93): the embedded activation's forward / backward
    wrap the GEMMs exactly as in the reference."""
    from neuromah_b200.src.layers import Layer_Dense
    from neuromah_b200.src.activations import Activation_Sigmoid
    rng = np.random.default_rng(17)
    x = rng.standard_normal((37, 19)).astype(np.float32)
    w = (rng.standard_normal((19, 11)) * 0.4).astype(np.float32)
    b = rng.standard_normal((1, 11)).astype(np.float32)
    dv = rng.standard_normal((37, 11)).astype(np.float32)
    layer = Layer_Dense(19, 11, activation=Activation_Sigmoid())
    layer.set_parameters(w, b)
    layer.forward(x, training=True)
    x64, w64, b64, dv64 = (a.astype(np.float64) for a in (x, w, b, dv))
    out = R.sigmoid_forward(R.dense_forward(x64, w64, b64))
    close(layer.output, out)
    layer.backward(dv)
    dx, dw, db = R.dense_backward(x64, w64, R.sigmoid_backward(out, dv64))
    close(layer.dinputs, dx)
    close(layer.dweights, dw)
    close(layer.dbiases, db)


def test_tensor_monitor_hooks_on_the_generic_training_loop(tmp_path, capsys):
    """Model.train with a TensorMonitor and a non-softmax head: the per-epoch hooks of Model.py:203-227 run on the generic
    loop too (scalars, parameter histograms computed on the device, one JSON file at the end)."""
    import json
    from neuromah_b200.src.core import Model
    from neuromah_b200.src.layers import Layer_Dense
    from neuromah_b200.src.activations import Activation_ReLU, Activation_Sigmoid
    from neuromah_b200.src.losses import Loss_BinaryCrossentropy
    from neuromah_b200.src.optimizers import Optimizer_Adam
    from neuromah_b200.src.metrics import Accuracy_Categorical
    from neuromah_b200.src.utils import TensorMonitor
    _, x, y = losses_family_case("bce")
    np.random.seed(3)
    model = Model()
    for l in (Layer_Dense(8, 16, activation=Activation_ReLU()), Layer_Dense(16, 2), Activation_Sigmoid()):
        model.add(l)
    tm = TensorMonitor(log_dir=str(tmp_path))
    model.set(loss=Loss_BinaryCrossentropy, optimizer=Optimizer_Adam(learning_rate=0.01),     # loss passed as a CLASS
              accuracy=Accuracy_Categorical(binary=True, model=model), tensorMonitor=tm)
    model.finalize()
    model.train(x, y, epochs=2, batch_size=64, verbose=1)
    assert "data_loss" in capsys.readouterr().out
    assert model.last_epoch["loss"] > 0 and 0 <= model.last_epoch["acc"] <= 1
    log = json.load(open(tm.log_file_path))
    assert log["metadata"]["loss_function"] == "Loss_BinaryCrossentropy" and len(log["epochs"]) == 2
    ep = log["epochs"][1]
    assert ep["metrics"]["scalars"]["loss/epoch_training"] == pytest.approx(model.last_epoch["loss"])
    hist = ep["parameters"]["histograms"]["Layer_Dense/weights"]
    assert sum(hist["histogram"]) in (8 * 16, 16 * 2) and len(hist["bin_edges"]) == 11
