"""CPU tests: the oracle (oracle/np_ref.py, oracle/ref_net.py) against
  * the reference's own known-answer tests (tests/layers/test_maxpool2D.py:14-57,
    tests/optimizers/test_Adam.py:13-35, examples/test.ipynb cell 18 softmax fingerprint),
  * golden vectors produced by the unmodified reference (tests/golden/make_golden.py),
  * the live reference when /root/reference is present,
  * float64 finite differences for the conv restatement (the reference pins no conv values).
"""
import numpy as np
import pytest

from conftest import golden, synth
from oracle import np_ref as R
from oracle import ref_net


# ---------------------------------------------------------------- max-pool KATs ----
RAMP = np.arange(1, 17, dtype=np.float32).reshape(1, 1, 4, 4)


def _pool_backends():
    out = [("numpy", R)]
    if ref_net._REF_POOL is not None:
        out.append(("ref_cpp", ref_net._REF_POOL))
    return out


@pytest.mark.parametrize("name,be", _pool_backends())
def test_maxpool_reference_kat(name, be):
    out, idx = be.maxpool2d_forward_cpu(RAMP, 2, 2, 2, 2, "valid")
    np.testing.assert_array_equal(out, [[[[6, 8], [14, 16]]]])
    dv = np.array([[[[1, 2], [3, 4]]]], dtype=np.float32)
    dx = be.maxpool2d_backward_cpu(dv, idx, RAMP, 2, 2, 2, 2, "valid")
    np.testing.assert_array_equal(dx, [[[[0, 0, 0, 0], [0, 1, 0, 2], [0, 0, 0, 0], [0, 3, 0, 4]]]])


def test_maxpool_tie_rule_all_zero():
    """ties keep the first element of the row-major scan."""
    _, idx = R.maxpool2d_forward_cpu(np.zeros((1, 1, 4, 4), np.float32), 2, 2, 2, 2, "valid")
    np.testing.assert_array_equal(idx[0, 0].reshape(4, 2), [[0, 0], [0, 2], [2, 0], [2, 2]])


@pytest.mark.skipif(ref_net._REF_POOL is None, reason="oracle/_ref not built")
@pytest.mark.parametrize("ps,st,pad", [((2, 2), (2, 2), "valid"), ((3, 3), (2, 2), "same"), ((2, 3), (1, 2), "same"),
                                        ((3, 3), (1, 1), "valid"), ((2, 2), (3, 3), "same")])
def test_maxpool_restatement_matches_reference_cpp(ps, st, pad):
    rng = np.random.default_rng(0)
    x = np.round(rng.standard_normal((2, 3, 9, 8)) * 2).astype(np.float32)      # ties on purpose
    o1, i1 = R.maxpool2d_forward_cpu(x, *ps, *st, pad)
    o2, i2 = ref_net._REF_POOL.maxpool2d_forward_cpu(x, *ps, *st, pad)
    np.testing.assert_array_equal(o1, o2)
    np.testing.assert_array_equal(i1, i2)
    dv = rng.standard_normal(o1.shape).astype(np.float32)
    d1 = R.maxpool2d_backward_cpu(dv, i1, x, *ps, *st, pad)
    d2 = ref_net._REF_POOL.maxpool2d_backward_cpu(dv, i2, x, *ps, *st, pad)
    np.testing.assert_allclose(d1, d2, rtol=1e-6, atol=1e-6)


# ------------------------------------------------------------------- softmax KAT ----
def test_softmax_notebook_fingerprint():
    """examples/test.ipynb cell 18: softmax of a one-hot row = e/(e+9), 1/(e+9)."""
    p = R.softmax_forward(np.eye(10)[[3]])
    assert abs(p[0, 3] - 0.23196932) < 1e-8 and abs(p[0, 0] - 0.08533674) < 1e-8


# -------------------------------------------------------------- golden: op level ----
def test_ops_golden():
    g = golden("ops.npz")
    pre = R.dense_forward(g["dense_x"], g["dense_w"], np.zeros((1, 12)))
    np.testing.assert_allclose(R.relu_forward(pre), g["dense_out"], rtol=0, atol=1e-12)
    dx, dw, db = R.dense_backward(g["dense_x"], g["dense_w"], R.relu_backward(g["dense_dv"], pre), l1=0.01, l2=0.02)
    np.testing.assert_allclose(dw, g["dense_dw"], atol=1e-12)
    np.testing.assert_allclose(db, g["dense_db"], atol=1e-12)
    np.testing.assert_allclose(dx, g["dense_dx"], atol=1e-12)
    sm = R.softmax_forward(g["sm_in"])
    np.testing.assert_allclose(sm, g["sm_out"], atol=1e-15)
    np.testing.assert_allclose(R.cce_forward(sm, g["labels"]), g["cce"], atol=1e-15)
    np.testing.assert_allclose(R.cce_forward(sm, np.eye(10)[g["labels"]]), g["cce_onehot"], atol=1e-15)
    np.testing.assert_allclose(R.softmax_cce_backward(sm, np.eye(10)[g["labels"]]), g["fused_grad"], atol=1e-15)
    np.testing.assert_allclose(R.cce_backward(sm, g["labels"]), g["cce_bwd"], atol=1e-12)
    np.testing.assert_allclose(R.softmax_backward(sm, g["sm_dv"]), g["sm_bwd"], atol=1e-12)
    for tag, (ps, st, pad) in {"a": ((2, 2), (2, 2), "valid"), "b": ((3, 3), (2, 2), "same"),
                               "c": ((2, 3), (1, 2), "same")}.items():
        o, i = R.maxpool2d_forward_cpu(g["pool_x"], *ps, *st, pad)
        np.testing.assert_array_equal(o, g[f"pool_{tag}_out"])
        np.testing.assert_array_equal(i, g[f"pool_{tag}_idx"])
        np.testing.assert_allclose(R.maxpool2d_backward_cpu(g[f"pool_{tag}_dv"], i, g["pool_x"], *ps, *st, pad),
                                   g[f"pool_{tag}_dx"], atol=1e-6)
    # Adam, applied twice per iteration with the same t (Q1), decay 0.1
    w, m, v = g["adam_w0"].copy(), 0.0, 0.0
    for it in range(3):
        lr = R.decayed_lr(0.01, 0.1, it)
        for _ in range(2):
            w, m, v = R.adam_update(w, g["adam_g"][it], m, v, lr=lr, t=it)
        np.testing.assert_allclose(w, g[f"adam_w{it + 1}"], rtol=0, atol=1e-15)


# ------------------------------------------------------------ golden: model level ----
def test_mlp_200_steps_golden():
    g = golden("mlp_200steps.npz")
    spec = ref_net.mlp_spec()
    net = ref_net.RefNet(spec, ref_net.init_params(spec, np.random.RandomState(0)))
    x, y = synth(1, 128 * 200, (784,))
    losses = [net.step(x[s * 128:(s + 1) * 128], y[s * 128:(s + 1) * 128])["loss"] for s in range(200)]
    np.testing.assert_allclose(losses, g["losses"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(net.flat(), g["final_params"], rtol=1e-6, atol=1e-7)


def test_mlp_sgd_golden():
    g = golden("mlp_sgd_20steps.npz")
    spec = ref_net.mlp_spec(64, 32, 10)
    net = ref_net.RefNet(spec, ref_net.init_params(spec, np.random.RandomState(5)), optimizer="sgd", lr=0.5, decay=0.01)
    x, y = synth(6, 32 * 20, (64,))
    losses = [net.step(x[s * 32:(s + 1) * 32], y[s * 32:(s + 1) * 32])["loss"] for s in range(20)]
    np.testing.assert_allclose(losses, g["losses"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(net.flat(), g["final_params"], rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("kind,lr,decay,eps", [("rmsprop", 0.01, 1e-3, 1e-7), ("adagrad", 0.05, 0.0, 1e-7), ("nadam", 0.01, 1e-2, 1e-8)])
def test_single_layer_optimizer_goldens(kind, lr, decay, eps):
    """RMSprop / Adagrad / Nadam restatements against the unmodified reference (softmax regression, 20 steps, Q1 doubling)."""
    g = golden("optimizers_single_layer_20steps.npz")
    spec = [{"kind": "dense", "n_in": 64, "n_out": 10, "relu": False}, {"kind": "softmax"}]
    net = ref_net.RefNet(spec, ref_net.init_params(spec, np.random.RandomState(31)), optimizer=kind, lr=lr, decay=decay, eps=eps)
    x, y = synth(32, 32 * 20, (64,))
    losses = [net.step(x[s * 32:(s + 1) * 32], y[s * 32:(s + 1) * 32])["loss"] for s in range(20)]
    np.testing.assert_allclose(losses, g[f"{kind}_losses"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(net.flat(), g[f"{kind}_params"], rtol=1e-9, atol=1e-12)


def test_dropout_mlp_golden():
    """Layer_Dropout fwd/bwd restatement with the masks the reference drew (Dropout.py:27), Adam, 10 steps + inference."""
    g = golden("dropout_mlp_10steps.npz")
    spec = [{"kind": "dense", "n_in": 64, "n_out": 32, "relu": True}, {"kind": "dropout", "rate": 0.3},
            {"kind": "dense", "n_in": 32, "n_out": 10, "relu": False}, {"kind": "softmax"}]
    params = ref_net.init_params(spec, np.random.RandomState(41))
    net = ref_net.RefNet(spec, params, optimizer="adam", lr=0.01)
    x, y = synth(42, 16 * 10, (64,))
    assert set(np.unique(g["masks"])) == {0.0, np.float32(1 / 0.7)}
    losses = []
    for s in range(10):
        net.dropout_masks[1] = g["masks"][s].astype(np.float64)
        losses.append(net.step(x[s * 16:(s + 1) * 16], y[s * 16:(s + 1) * 16])["loss"])
    np.testing.assert_allclose(losses, g["losses"], rtol=1e-6, atol=1e-8)       # masks are stored as float32
    np.testing.assert_allclose(net.flat(), g["final_params"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(net.forward(x[:16], training=False), g["infer"], rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("mode,fixture,steps", [("wired", "cnn_small_3steps.npz", 3), ("stub", "cnn_stub_2steps.npz", 2)])
def test_cnn_golden(mode, fixture, steps):
    g = golden(fixture)
    spec = ref_net.mnist_cnn_spec()
    net = ref_net.RefNet(spec, ref_net.init_params(spec, np.random.RandomState(2)), conv_backward=mode)
    x, y = synth(3, 4 * steps, (1, 28, 28))
    for s in range(steps):
        r = net.step(x[s * 4:(s + 1) * 4], y[s * 4:(s + 1) * 4])
        assert abs(r["loss"] - float(g[f"loss{s}"])) < 1e-12
        if mode == "wired":
            if s == 0:
                for i, o in enumerate(net.outs):
                    np.testing.assert_allclose(o, g[f"out{i}"], rtol=1e-6, atol=1e-7)
                np.testing.assert_allclose(net.flat("grads"), g["grads0"], rtol=1e-9, atol=1e-12)
            np.testing.assert_allclose(net.flat(), g[f"params{s}"], rtol=1e-6, atol=1e-7)
    if mode == "stub":
        np.testing.assert_allclose(net.flat(), g["final_params"], rtol=1e-6, atol=1e-7)


def test_vgg_small_golden():
    """The conv -> conv -> pool stack of BASELINE.json config 4 (narrowed 8x): the oracle's step reproduces the run of
    the unmodified reference Python (tests/golden/make_golden.py::vgg_small) -- losses, step-0 probabilities and
    gradients, parameters after each of the two Adam steps."""
    g = golden("vgg_small_2steps.npz")
    spec, ic = [], 3
    for oc in (8, 16):
        spec += [{"kind": "conv", "ic": ic, "oc": oc, "k": 3, "padding": "same", "relu": True},
                 {"kind": "conv", "ic": oc, "oc": oc, "k": 3, "padding": "same", "relu": True},
                 {"kind": "pool", "size": 2, "stride": 2, "padding": "valid"}]
        ic = oc
    spec += [{"kind": "flatten"}, {"kind": "dense", "n_in": 16 * 8 * 8, "n_out": 10, "relu": False}, {"kind": "softmax"}]
    net = ref_net.RefNet(spec, ref_net.init_params(spec, np.random.RandomState(31), init="he"))
    x, y = synth(32, 6 * 2, (3, 32, 32))
    for s in range(2):
        r = net.step(x[s * 6:(s + 1) * 6], y[s * 6:(s + 1) * 6])
        assert abs(r["loss"] - float(g[f"loss{s}"])) < 1e-9
        if s == 0:
            np.testing.assert_allclose(r["output"], g["probs0"], rtol=1e-5, atol=1e-7)
            np.testing.assert_allclose(net.flat("grads"), g["grads0"], rtol=1e-4, atol=1e-9)
        np.testing.assert_allclose(net.flat(), g[f"params{s}"], rtol=1e-5, atol=1e-6)


# -------------------------------------------------------- live reference (here) ----
def test_oracle_matches_live_reference(have_reference):
    if not have_reference:
        pytest.skip("reference tree not present on this machine")
    from oracle import ref_shim
    spec = ref_net.mnist_cnn_spec()
    params = ref_net.init_params(spec, np.random.RandomState(9))
    model = ref_shim.build_model(spec, params)
    assert len(model.trainable_layers) == 6          # Q1: 3 layers registered twice
    net = ref_net.RefNet(spec, params)
    x, y = synth(10, 8 * 2, (1, 28, 28))
    for s in range(2):
        a = ref_shim.train_step(model, x[s * 8:(s + 1) * 8], y[s * 8:(s + 1) * 8])
        b = net.step(x[s * 8:(s + 1) * 8], y[s * 8:(s + 1) * 8])
        assert a["loss"] == b["loss"]
    ref_flat = np.concatenate([np.concatenate([l.weights.ravel(), l.biases.ravel()])
                               for l in model.trainable_layers[:3]])
    np.testing.assert_array_equal(net.flat(), ref_flat)


# ------------------------------------------------ conv: finite differences (f64) ----
@pytest.mark.parametrize("padding,k", [("valid", (3, 3)), ("same", (3, 3)), ("same", (2, 4)), ("valid", (1, 1))])
def test_conv_gradcheck(padding, k):
    rng = np.random.default_rng(4)
    x = rng.standard_normal((2, 3, 6, 7))
    w = rng.standard_normal((4, 3) + k)
    dv = rng.standard_normal(R.conv_layer_forward(x, w, None, padding=padding, dtype=np.float64)[0].shape)

    def f(xx, ww):
        return float(np.sum(R.conv_layer_forward(xx, ww, None, padding=padding, dtype=np.float64)[0] * dv))

    dx, dw, db = R.conv_layer_backward(x, w, None, dv, padding=padding, mode="wired", dtype=np.float64)
    eps = 1e-6
    for arr, grad in ((x, dx), (w, dw)):
        for _ in range(12):
            i = tuple(rng.integers(0, s) for s in arr.shape)
            old = arr[i]
            arr[i] = old + eps; fp = f(x, w)
            arr[i] = old - eps; fm = f(x, w)
            arr[i] = old
            assert abs((fp - fm) / (2 * eps) - grad[i]) < 1e-6
    np.testing.assert_allclose(db[:, 0], dv.sum(axis=(0, 2, 3)), rtol=1e-12)


def test_conv_matches_direct_definition():
    """im2col-GEMM restatement == direct cross-correlation sum (documented intent)."""
    rng = np.random.default_rng(5)
    x = rng.standard_normal((2, 3, 5, 6)).astype(np.float32)
    w = rng.standard_normal((4, 3, 3, 2)).astype(np.float32)
    out = R.conv2d_cpu(x, w)
    ref = np.zeros_like(out)
    for p in range(3):
        for q in range(2):
            ref += np.einsum("nchw,oc->nohw", x[:, :, p:p + 3, q:q + 5], w[:, :, p, q])
    np.testing.assert_allclose(out, ref, rtol=1e-5, atol=1e-5)


def test_federated_average():
    u = [np.array([1.0, 2.0]), np.array([3.0, 6.0])]
    np.testing.assert_allclose(R.federated_average(u, [100, 300]), [2.5, 5.0])
