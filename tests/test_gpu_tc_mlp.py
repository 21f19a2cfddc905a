"""GPU parity of the general tcgen05 Dense kernels (nm_tc_gemm.cu) and of the MLP tensor-core plan
(neuromah_b200/tc_plan_mlp.py) -- BASELINE.json config 1 (784-128-10 MLP, batch 128) and the hidden-layer shapes of
examples/test_Dense_MNIST.py:39-44 (784-128-64-10) -- against the oracle (Dense.py:64-108, Relu.py:7-14).

Stated tolerance of the mode (bf16 operands and stored activations, FP32 accumulation, FP32 master weights):
  op level     bit-exact on integer-valued data (products and sums exactly representable); real-valued data within
               bf16 rounding of the output (2^-8 relative) + FP32 summation error
  plan level   probabilities |err| <= 2e-3 + 1e-2 |ref|; gradient tensors relative L2 <= 3e-2 for the head and the layer
               below it, <= 6e-2 deeper (the convention of tests/test_gpu_vgg_tc_plan.py); the 200-step loss curve of
               config 1 within 2e-2 of the reference's own golden curve (tests/golden/mlp_200steps.npz)
"""
import numpy as np
import pytest

from conftest import golden, synth
from oracle import np_ref as R
from oracle import ref_net
from test_gpu_model import _resync_oracle, build

pytestmark = pytest.mark.gpu


def bf16_round(a):
    i = np.ascontiguousarray(a, np.float32).view(np.uint32).astype(np.uint64)
    return (((i + 0x7FFF + ((i >> 16) & 1)) >> 16) << 16).astype(np.uint32).view(np.float32)


def to_bits(a):
    return (np.ascontiguousarray(bf16_round(a)).view(np.uint32) >> 16).astype(np.uint16)


def from_bits(b):
    return (b.astype(np.uint32) << 16).view(np.float32)


def rel_l2(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


CASES = [(128, 784, 128), (128, 128, 64), (300, 784, 128), (4096, 784, 128), (64, 64, 64), (1, 8, 64), (257, 200, 192)]


@pytest.mark.parametrize("B,n_in,n_out", CASES)
@pytest.mark.parametrize("integer", [True, False])
def test_dense_gen_fprop_dgrad_wgrad(B, n_in, n_out, integer):
    from neuromah_b200._lib import lib
    from neuromah_b200.device import DeviceArray, stream
    rng = np.random.default_rng(B + n_in + n_out)
    if integer:
        x = rng.integers(-2, 3, (B, n_in)).astype(np.float32)
        w = rng.integers(-1, 2, (n_in, n_out)).astype(np.float32)
        b = rng.integers(-3, 4, (1, n_out)).astype(np.float32)
        dy = rng.integers(-2, 3, (B, n_out)).astype(np.float32)
    else:
        x = bf16_round(rng.standard_normal((B, n_in)).astype(np.float32))
        w = (rng.standard_normal((n_in, n_out)) * 0.05).astype(np.float32)
        b = rng.standard_normal((1, n_out)).astype(np.float32)
        dy = bf16_round(rng.standard_normal((B, n_out)).astype(np.float32))
    wb = bf16_round(w)
    w_dev = DeviceArray.from_numpy(w)
    w_kn, w_nk = DeviceArray((n_in, n_out), np.uint16), DeviceArray((n_out, n_in), np.uint16)
    lib.nm_tc_pack_dense_gen_bf16(w_dev.p, w_kn.p, w_nk.p, n_in, n_out, stream())
    np.testing.assert_array_equal(w_kn.numpy(), to_bits(w))
    np.testing.assert_array_equal(w_nk.numpy(), to_bits(w).T)
    # ---- fprop (+ bias + ReLU); float32 -> bf16 conversion of the input on the device
    xf = DeviceArray.from_numpy(x)
    xb = DeviceArray.from_numpy(np.full((B, n_in), 0x7FC0, np.uint16))
    lib.nm_tc_to_bf16(xf.p, xb.p, x.size, stream())
    np.testing.assert_array_equal(xb.numpy(), to_bits(x))
    b_dev = DeviceArray.from_numpy(b)
    y = DeviceArray.from_numpy(np.full((B, n_out), 0x7FC0, np.uint16))
    lib.nm_tc_dense_fprop_gen_bf16(xb.p, w_nk.p, b_dev.p, y.p, B, n_in, n_out, 1, 1.0, 0, 0, None, stream())
    ref = np.maximum(x.astype(np.float64) @ wb.astype(np.float64) + b, 0)
    got = from_bits(y.numpy())
    if integer:
        np.testing.assert_array_equal(got, bf16_round(ref.astype(np.float32)))
    else:
        np.testing.assert_allclose(got, ref, rtol=2 ** -7, atol=2e-3)
    # ---- dgrad with the ReLU mask of the layer below
    act_below = np.maximum(rng.integers(-2, 3, (B, n_in)), 0).astype(np.float32)
    dyb, mask = DeviceArray.from_numpy(to_bits(dy)), DeviceArray.from_numpy(to_bits(act_below))
    dx = DeviceArray.from_numpy(np.full((B, n_in), 0x7FC0, np.uint16))
    lib.nm_tc_dense_dgrad_gen_bf16(dyb.p, w_kn.p, mask.p, 1.0, dx.p, B, n_in, n_out, stream())
    refd = (dy.astype(np.float64) @ wb.astype(np.float64).T) * (act_below > 0)
    gotd = from_bits(dx.numpy())
    if integer:
        np.testing.assert_array_equal(gotd, bf16_round(refd.astype(np.float32)))
    else:
        np.testing.assert_allclose(gotd, refd, rtol=2 ** -7, atol=2e-3)
    lib.nm_tc_dense_dgrad_gen_bf16(dyb.p, w_kn.p, None, 1.0, dx.p, B, n_in, n_out, stream())
    refd = dy.astype(np.float64) @ wb.astype(np.float64).T
    if integer:
        np.testing.assert_array_equal(from_bits(dx.numpy()), bf16_round(refd.astype(np.float32)))
    # ---- wgrad (+ bias gradient, L1 / L2 terms), Dense.py:96-105 with scale = 1 / B
    dw = DeviceArray.from_numpy(np.full((n_in, n_out), np.nan, np.float32))
    db = DeviceArray.from_numpy(np.full((1, n_out), np.nan, np.float32))
    ws = DeviceArray((max(lib.nm_tc_dense_wgrad_gen_workspace(B, n_in, n_out), 16),), np.uint8)
    scale, l1, l2 = (1.0, 0.0, 0.0) if integer else (1.0 / B, 1e-3, 2e-3)
    lib.nm_tc_dense_wgrad_gen_bf16(xb.p, dyb.p, w_dev.p, dw.p, db.p, ws.p, ws.nbytes, B, n_in, n_out, scale, l1, l2, stream())
    refw = scale * (x.astype(np.float64).T @ dy.astype(np.float64)) + l1 * np.sign(w) + 2 * l2 * w
    refb = scale * dy.astype(np.float64).sum(axis=0, keepdims=True)
    if integer:
        np.testing.assert_array_equal(dw.numpy(), refw.astype(np.float32))
        np.testing.assert_array_equal(db.numpy(), refb.astype(np.float32))
    else:
        np.testing.assert_allclose(dw.numpy(), refw, rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(db.numpy(), refb, rtol=1e-4, atol=1e-5)
    dw2 = DeviceArray.zeros((n_in, n_out))
    lib.nm_tc_dense_wgrad_gen_bf16(xb.p, dyb.p, w_dev.p, dw2.p, None, ws.p, ws.nbytes, B, n_in, n_out, scale, l1, l2, stream())
    np.testing.assert_array_equal(dw.numpy(), dw2.numpy())           # bitwise repeatable


def mlp3_spec():
    """Hidden-layer shapes of examples/test_Dense_MNIST.py:39-44 (without its Dropout layers / doubled softmax)."""
    return [{"kind": "dense", "n_in": 784, "n_out": 128, "relu": True}, {"kind": "dense", "n_in": 128, "n_out": 64, "relu": True},
            {"kind": "dense", "n_in": 64, "n_out": 10, "relu": False}, {"kind": "softmax"}]


@pytest.mark.parametrize("spec_fn,batch", [(ref_net.mlp_spec, 128), (mlp3_spec, 96)])
def test_tc_mlp_plan_step_gradients_from_identical_state(spec_fn, batch):
    from neuromah_b200.tc_plan_mlp import TcMlpPlan
    spec = spec_fn()
    params = ref_net.init_params(spec, np.random.RandomState(5), init="he")
    model = build(spec, params, mode="bf16", learning_rate=0.001)
    assert isinstance(model.plan, TcMlpPlan)
    net = ref_net.RefNet(spec, params)
    x, y = synth(6, 3 * batch, (784,))
    for s in range(3):
        _resync_oracle(net, model)
        xb, yb = x[s * batch:(s + 1) * batch], y[s * batch:(s + 1) * batch]
        model._accum.fill_zero()
        out = model.train_step(xb, yb)
        r = net.step(xb, yb)
        np.testing.assert_allclose(np.asarray(out), r["output"], rtol=1e-2, atol=2e-3)
        loss_sum, _ = model._read_accum()
        assert abs(loss_sum - r["loss_sum"]) <= 2e-3 * batch
        g, gr = model.arena.host_grads(), net.flat("grads")
        off, errs = 0, []
        trainable = [l for l in model.layers if hasattr(l, "weights")]
        for (layer, name, _, size, _) in model.arena.table:
            depth = len(trainable) - 1 - [id(t) for t in trainable].index(id(layer))       # 0 = the head
            errs.append((rel_l2(g[off:off + size], gr[off:off + size]), 3e-2 if depth == 0 else 6e-2))
            off += size
        print("step", s, " ".join(f"{e:.4f}" for e, _ in errs))
        assert all(e <= tol for e, tol in errs), errs
        for li, layer in enumerate(model.layers[:-2]):                   # hidden activations, host-convertible views
            a = np.asarray(layer.output)
            assert a.shape == net.outs[li].shape and np.abs(a - net.outs[li]).max() <= 2e-2 * np.abs(net.outs[li]).max()
    assert model.optimizer.iterations == 3


def test_tc_mlp_200_step_loss_curve_golden():
    """BASELINE.json configs[0] in tensor-core mode: the reference's own 200-step loss curve (mlp_200steps.npz)."""
    g = golden("mlp_200steps.npz")
    spec = ref_net.mlp_spec()
    model = build(spec, ref_net.init_params(spec, np.random.RandomState(0)), mode="bf16", learning_rate=0.001)
    x, y = synth(1, 128 * 200, (784,))
    losses = []
    for s in range(200):
        model._accum.fill_zero()
        model.train_step(x[s * 128:(s + 1) * 128], y[s * 128:(s + 1) * 128])
        losses.append(model._read_accum()[0] / 128)
    diff = np.abs(np.array(losses) - g["losses"])
    print(f"max diff {diff.max():.4g} at step {diff.argmax()}, mean {diff.mean():.4g}")
    assert diff.max() <= 2e-2 and diff[:20].max() <= 5e-3


def test_tc_mlp_train_api_graph_replay_uint8_and_determinism():
    """Model.train in bf16 mode on an MLP: graph-replayed steps, raw uint8 inputs == host-normalised floats, repeatable."""
    spec = mlp3_spec()
    params = ref_net.init_params(spec, np.random.RandomState(9), init="he")
    rng = np.random.default_rng(3)
    xu = rng.integers(0, 256, (6 * 128, 784)).astype(np.uint8)
    lab = rng.integers(0, 10, len(xu))
    xu[np.arange(len(xu)), lab * 70] = 255                               # a learnable signal
    y = np.eye(10)[lab]
    runs = []
    for x in (xu.astype(np.float32) / 255.0, xu, xu):
        m = build(spec, params, mode="bf16", learning_rate=0.002)
        m.train(x, y, epochs=3, batch_size=128, verbose=0)
        assert len(m._graphs) == 1
        runs.append((m.arena.host_params(), m.last_epoch["loss"]))
    np.testing.assert_array_equal(runs[0][0], runs[1][0])
    np.testing.assert_array_equal(runs[1][0], runs[2][0])
    eager = build(spec, params, mode="bf16", learning_rate=0.002)
    eager.graph_steps = False
    eager.train(xu, y, epochs=3, batch_size=128, verbose=0)
    np.testing.assert_allclose(runs[1][0], eager.arena.host_params(), rtol=1e-3, atol=1e-4)
    assert np.isfinite(runs[0][1]) and runs[0][1] < 2.4


def _dense_example_model(mode, seeds=(71, 72), lr=0.001):
    """The stack of examples/test_Dense_MNIST.py:39-44 (Dense+ReLU, Dropout(0.25), Dense+ReLU, Dropout(0.25), Dense, Softmax)
    with a single softmax."""
    from neuromah_b200.src.core import Model
    from neuromah_b200.src.layers import Layer_Dense, Layer_Dropout
    from neuromah_b200.src.activations import Activation_ReLU, Activation_Softmax
    from neuromah_b200.src.losses import Loss_CategoricalCrossentropy
    from neuromah_b200.src.optimizers import Optimizer_Adam
    from neuromah_b200.src.metrics import Accuracy_Categorical
    spec = [{"kind": "dense", "n_in": 784, "n_out": 128, "relu": True}, {"kind": "dropout", "rate": 0.25},
            {"kind": "dense", "n_in": 128, "n_out": 64, "relu": True}, {"kind": "dropout", "rate": 0.25},
            {"kind": "dense", "n_in": 64, "n_out": 10, "relu": False}, {"kind": "softmax"}]
    params = ref_net.init_params(spec, np.random.RandomState(17), init="he")
    model = Model()
    layers = [Layer_Dense(784, 128, activation=Activation_ReLU()), Layer_Dropout(0.25, seed=seeds[0]),
              Layer_Dense(128, 64, activation=Activation_ReLU()), Layer_Dropout(0.25, seed=seeds[1]),
              Layer_Dense(64, 10), Activation_Softmax()]
    for layer, p in zip(layers, params):
        if p is not None:
            layer.set_parameters(p["weights"], p["biases"])
        model.add(layer)
    model.set(loss=Loss_CategoricalCrossentropy, optimizer=Optimizer_Adam(learning_rate=lr), accuracy=Accuracy_Categorical)
    model.finalize(mode=mode)
    return model, spec, params


def _philox_mask(shape, keep, seed, offset):
    """The mask Layer_Dropout's FP32 kernel draws for (seed, offset): the fused epilogue must draw the same one."""
    from neuromah_b200._lib import lib
    from neuromah_b200.device import DeviceArray, stream
    ones = DeviceArray.from_numpy(np.ones(shape, np.float32))
    out, mask = DeviceArray(shape, np.float32), DeviceArray(shape, np.float32)
    lib.nm_dropout_fwd_f32(ones.p, out.p, mask.p, ones.size, float(keep), seed, offset, stream())
    return mask.numpy().astype(np.float64)


def test_tc_mlp_plan_with_fused_dropout_vs_oracle():
    """Dense example stack in tensor-core mode: the Dropout layers are fused into the Dense epilogues.  Step s draws the mask
    of Philox stream offset s (the optimizer's step count, read on the device), identical to the FP32 layer's mask for that
    (seed, offset); with those masks the oracle's step (Dropout.py:21-40) gives the same outputs and gradients."""
    from neuromah_b200.tc_plan_mlp import TcMlpPlan
    model, spec, params = _dense_example_model("bf16")
    assert isinstance(model.plan, TcMlpPlan) and len(model.plan.drops) == 2
    net = ref_net.RefNet(spec, params)
    B = 64
    x, y = synth(81, 3 * B, (784,))
    for s in range(3):
        _resync_oracle(net, model)
        xb, yb = x[s * B:(s + 1) * B], y[s * B:(s + 1) * B]
        net.dropout_masks = {1: _philox_mask((B, 128), 0.75, 71, s), 3: _philox_mask((B, 64), 0.75, 72, s)}
        model._accum.fill_zero()
        out = model.train_step(xb, yb)
        r = net.step(xb, yb)
        np.testing.assert_allclose(np.asarray(out), r["output"], rtol=1e-2, atol=2e-3)
        h1 = np.asarray(model.layers[0].output)                       # the dropped activation
        alive = net.outs[0] > 1e-2                                    # clearly positive before the dropout (ReLU near-ties excluded)
        assert np.array_equal((h1 == 0)[alive], (net.dropout_masks[1] == 0)[alive]), "dropout pattern differs from the FP32 layer's mask"
        g, gr = model.arena.host_grads(), net.flat("grads")
        off, errs = 0, []
        for (_, name, _, size, _) in model.arena.table:
            errs.append(rel_l2(g[off:off + size], gr[off:off + size]))
            off += size
        print("step", s, " ".join(f"{e:.4f}" for e in errs))
        assert errs[-1] <= 3e-2 and errs[-2] <= 3e-2 and max(errs) <= 8e-2, errs      # measured <= 0.058 (two bf16 dropout-scaled layers)
    # inference: dropout is the identity (Dropout.py:23-25)
    p_eval = np.asarray(model.forward(x[:B], training=False))
    net.set_flat(model.arena.host_params().astype(np.float64))
    np.testing.assert_allclose(p_eval, net.forward(x[:B], training=False), rtol=1e-2, atol=2e-3)


def test_tc_mlp_dropout_graph_replay_draws_a_fresh_mask_every_step():
    """Model.train replays ONE captured graph; the dropout offset is the on-device step counter, so the replayed steps equal
    eager steps (which pass the host's step count) -- a mask baked into the graph would not."""
    rng = np.random.default_rng(5)
    x = rng.random((4 * 128, 784), dtype=np.float32)
    y = np.eye(10)[rng.integers(0, 10, len(x))]
    runs = []
    for graph in (True, False):
        model, _, _ = _dense_example_model("bf16", lr=0.002)
        model.graph_steps = graph
        model.train(x, y, epochs=3, batch_size=128, verbose=0)
        runs.append(model.arena.host_params())
        assert len(getattr(model, "_graphs", {})) == (1 if graph else 0)
    np.testing.assert_allclose(runs[0], runs[1], rtol=1e-3, atol=1e-4)
