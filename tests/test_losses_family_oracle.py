"""CPU: the oracle's sigmoid / MSE / MAE / binary-cross-entropy / regression-accuracy restatements against vectors the
UNMODIFIED reference produced (tests/golden/losses_family.npz, written by tests/golden/make_golden.py::losses_family), and
the oracle step driver on the three small models of that fixture."""
import numpy as np
import pytest

from conftest import LOSSES_FAMILY_CASES, golden, losses_family_case, losses_family_inputs
from oracle import np_ref as R
from oracle import ref_net


@pytest.fixture(scope="module")
def G():
    return golden("losses_family.npz")


def test_sigmoid_matches_the_reference(G):
    v = {k: a.astype(np.float64) for k, a in losses_family_inputs().items()}
    out = R.sigmoid_forward(v["x"])
    np.testing.assert_allclose(out, G["sigmoid_out"], rtol=1e-14)
    np.testing.assert_allclose(R.sigmoid_backward(out, v["dv"]), G["sigmoid_dx"], rtol=1e-14)
    np.testing.assert_array_equal(R.sigmoid_predictions(out), G["sigmoid_pred"])
    assert abs(R.sigmoid_forward(np.array([1.0]))[0] - np.e / (np.e + 1)) < 1e-15


@pytest.mark.parametrize("name,target", [("mse", "t_reg"), ("mae", "t_reg"), ("bce", "t_bin")])
def test_losses_match_the_reference(G, name, target):
    v = {k: a.astype(np.float64) for k, a in losses_family_inputs().items()}
    fwd, bwd = getattr(R, f"{name}_forward"), getattr(R, f"{name}_backward")
    np.testing.assert_allclose(fwd(v["pred"], v[target]), G[f"{name}_fwd"], rtol=1e-13)
    np.testing.assert_allclose(bwd(v["pred"], v[target]), G[f"{name}_bwd"], rtol=1e-13)
    # every non-batch axis is averaged; the gradient is divided by the elements per sample, not by the batch
    np.testing.assert_allclose(fwd(v["pred4"], v["t4"]), G[f"{name}_fwd4"], rtol=1e-13)
    np.testing.assert_allclose(bwd(v["pred4"], v["t4"]), G[f"{name}_bwd4"], rtol=1e-13)
    assert fwd(v["pred4"], v["t4"]).shape == (6,)


def test_loss_gradients_are_the_derivative_of_the_per_sample_mean():
    """Finite differences in float64: MSE and BCE gradients are d(sample loss)/d(pred); MAE's is its NEGATIVE (the
    reference ships sign(y - p), MAE.py:23 -- reproduced, not corrected)."""
    rng = np.random.default_rng(5)
    p, t = rng.random((4, 3)) * 0.8 + 0.1, (rng.random((4, 3)) > 0.5) * 1.0
    for name, sign in (("mse", 1.0), ("bce", 1.0), ("mae", -1.0)):
        fwd, bwd = getattr(R, f"{name}_forward"), getattr(R, f"{name}_backward")
        g = bwd(p, t)
        for (i, j) in ((0, 0), (2, 1), (3, 2)):
            e = np.zeros_like(p); e[i, j] = 1e-6
            fd = (fwd(p + e, t)[i] - fwd(p - e, t)[i]) / 2e-6
            assert abs(sign * g[i, j] - fd) < 1e-6 * max(1.0, abs(fd))


def test_bce_clips_saturated_predictions():
    p, t = np.array([[0.0, 1.0]]), np.array([[1.0, 0.0]])
    np.testing.assert_allclose(R.bce_forward(p, t), [-np.log(1e-7)], rtol=1e-9)
    np.testing.assert_allclose(R.bce_backward(p, t), [[-1 / 1e-7 / 2, 1 / (1 - (1 - 1e-7)) / 2]], rtol=1e-6)


def _spec(case):
    return [{"kind": "dense", "n_in": 8, "n_out": 16, "relu": True}, {"kind": "dense", "n_in": 16, "n_out": case["n_out"], "relu": False},
            {"kind": case["act"]}]


@pytest.mark.parametrize("name", sorted(LOSSES_FAMILY_CASES))
def test_oracle_step_driver_reproduces_the_reference_models(G, name):
    c = LOSSES_FAMILY_CASES[name]
    w, x, y = losses_family_case(name)
    params = [{"weights": w[0], "biases": np.zeros((1, 16))}, {"weights": w[1], "biases": np.zeros((1, c["n_out"]))}, None]
    net = ref_net.RefNet(_spec(c), params, optimizer=c["opt"], lr=c["lr"], loss=c["loss"])
    net.init_accuracy(y)
    if c["acc"] == "regression":
        assert net.precision == pytest.approx(float(G[f"{name}_precision"]), rel=1e-13)
    for s in range(c["steps"]):
        r = net.step(x[s * 32:(s + 1) * 32], y[s * 32:(s + 1) * 32])
        assert r["loss"] == pytest.approx(float(G[f"{name}_losses"][s]), rel=1e-10)
        assert r["acc"] == pytest.approx(float(G[f"{name}_accs"][s]), abs=1e-12)
    np.testing.assert_allclose(net.flat(), G[f"{name}_final_params"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(net.forward(x[:32], training=False), G[f"{name}_infer"], rtol=1e-9, atol=1e-12)


def test_accuracy_counts_elements_in_the_sum_and_samples_in_the_count(G):
    """Accuracy.calculate (Accuracy.py:19-23) with several outputs per sample: the binary model's running 'accuracy'
    of the reference run exceeds 1 (348 element hits over 256 samples)."""
    s, n = G["bce_acc_running"]
    assert (s, n) == (348.0, 256.0)
    mean, total, count = R.accuracy_elementwise(np.array([[True, False], [True, True], [False, False]]))
    assert (mean, total, count) == (0.5, 3.0, 3)
