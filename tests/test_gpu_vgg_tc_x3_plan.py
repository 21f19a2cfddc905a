"""GPU parity of the FP32-exact tensor-core mode for conv stacks (``Model.finalize(mode="bf16x3")`` -> TcStackPlan(split=True),
neuromah_b200/tc_plan_gen.py): three bf16 planes per tensor, six plane-pair products per contraction.

Stated tolerance: the FP32-exact one, rtol 1e-4 / atol 1e-5 against the reference's NumPy path -- checked here on the SAME
goldens of the unmodified reference that pin the CUDA-core FP32 mode (tests/test_gpu_model.py): the MNIST CNN's per-layer
outputs, gradient slab and post-Adam weights over 3 steps (cnn_small_3steps.npz) and the narrowed VGG-style stack of config 4
(vgg_small_2steps.npz, widths 8 / 16: zero-padded to 64 channels inside the plan), plus per-tensor gradient errors against the
float64 oracle from identical state on 64 / 128-channel layers (measured ~1e-6 relative L2).
"""
import numpy as np
import pytest

from conftest import golden, synth
from oracle import ref_net
from test_gpu_model import ATOL, RTOL, _resync_oracle, build, reference_style_step
from test_gpu_vgg_tc_plan import narrow_spec, rel_l2, small_vgg_spec

pytestmark = pytest.mark.gpu


def test_x3_mnist_cnn_per_layer_golden():
    """examples/test_Conv2D_MNIST.py:47-53 stack, batch 4, 3 Adam steps through the public API: every layer's output, the
    gradient slab and the post-update weights of the unmodified reference's run."""
    from neuromah_b200.tc_plan_gen import TcStackPlan
    g = golden("cnn_small_3steps.npz")
    spec = ref_net.mnist_cnn_spec()
    model = build(spec, ref_net.init_params(spec, np.random.RandomState(2)), mode="bf16x3", learning_rate=0.001)
    assert isinstance(model.plan, TcStackPlan) and model.plan.split
    model.loss.new_pass(); model.accuracy.new_pass()
    x, y = synth(3, 12, (1, 28, 28))
    for s in range(3):
        loss, _ = reference_style_step(model, x[s * 4:(s + 1) * 4], y[s * 4:(s + 1) * 4])
        assert abs(loss - float(g[f"loss{s}"])) < 1e-5
        if s == 0:
            for i, layer in enumerate(model.layers):
                np.testing.assert_allclose(np.asarray(layer.output).reshape(g[f"out{i}"].shape), g[f"out{i}"], rtol=RTOL, atol=ATOL,
                                           err_msg=f"out{i}")
            np.testing.assert_allclose(model.arena.host_grads(), g["grads0"], rtol=RTOL, atol=1e-6)
        np.testing.assert_allclose(model.arena.host_params(), g[f"params{s}"], rtol=RTOL, atol=ATOL, err_msg=f"step {s}")


def test_x3_vgg_small_golden():
    """The conv-conv-pool stack of config 4 narrowed 8x (3->8->8 P 8->16->16 P Dense), the run of the unmodified reference."""
    g = golden("vgg_small_2steps.npz")
    spec, ic = [], 3
    for oc in (8, 16):
        spec += [{"kind": "conv", "ic": ic, "oc": oc, "k": 3, "padding": "same", "relu": True},
                 {"kind": "conv", "ic": oc, "oc": oc, "k": 3, "padding": "same", "relu": True},
                 {"kind": "pool", "size": 2, "stride": 2, "padding": "valid"}]
        ic = oc
    spec += [{"kind": "flatten"}, {"kind": "dense", "n_in": 16 * 8 * 8, "n_out": 10, "relu": False}, {"kind": "softmax"}]
    model = build(spec, ref_net.init_params(spec, np.random.RandomState(31), init="he"), mode="bf16x3", learning_rate=0.001)
    x, y = synth(32, 6 * 2, (3, 32, 32))
    for s in range(2):
        model._accum.fill_zero()
        out = model.train_step(x[s * 6:(s + 1) * 6], y[s * 6:(s + 1) * 6])
        assert abs(model._read_accum()[0] / 6 - float(g[f"loss{s}"])) < 1e-5
        if s == 0:
            np.testing.assert_allclose(np.asarray(out), g["probs0"], rtol=RTOL, atol=ATOL)
            np.testing.assert_allclose(model.arena.host_grads(), g["grads0"], rtol=RTOL, atol=1e-6)
        np.testing.assert_allclose(model.arena.host_params(), g[f"params{s}"], rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("spec_fn", [small_vgg_spec, narrow_spec])
def test_x3_stack_step_vs_oracle_per_tensor(spec_fn):
    """64 / 128-channel layers (K loops over two chunks, N tiles of 64 and 128) and the zero-padded narrow widths, from
    identical state: probabilities, activations and every gradient tensor against the float64 oracle."""
    spec = spec_fn()
    params = ref_net.init_params(spec, np.random.RandomState(5), init="he")
    model = build(spec, params, mode="bf16x3", learning_rate=0.001)
    net = ref_net.RefNet(spec, params)
    x, y = synth(6, 3 * 32, (3, 16, 16))
    n_feat = len(spec) - 3                                            # conv / pool layers
    for s in range(3):
        _resync_oracle(net, model)
        xb, yb = x[s * 32:(s + 1) * 32], y[s * 32:(s + 1) * 32]
        model._accum.fill_zero()
        out = model.train_step(xb, yb)
        r = net.step(xb, yb)
        np.testing.assert_allclose(np.asarray(out), r["output"], rtol=RTOL, atol=ATOL)
        loss_sum, _ = model._read_accum()
        assert abs(loss_sum - r["loss_sum"]) <= 1e-5 * 32
        for li in range(n_feat):
            np.testing.assert_allclose(np.asarray(model.layers[li].output), net.outs[li], rtol=RTOL, atol=ATOL, err_msg=f"layer {li}")
        g, gr = model.arena.host_grads(), net.flat("grads")
        off, errs = 0, []
        for (_, name, _, size, _) in model.arena.table:
            errs.append(rel_l2(g[off:off + size], gr[off:off + size]))
            off += size
        print("step", s, " ".join(f"{e:.2e}" for e in errs))
        assert max(errs) <= 2e-5, errs
    assert model.optimizer.iterations == 3


def test_x3_stack_train_api_graph_replay_uint8_and_determinism():
    """Model.train in bf16x3 mode: one CUDA graph, bitwise repeatable, raw uint8 pixels == host-normalised float32, and the run
    follows the CUDA-core FP32 mode's weights closely (both are FP32-exact; their summation orders differ)."""
    spec = narrow_spec()
    params = ref_net.init_params(spec, np.random.RandomState(13), init="he")
    rng = np.random.default_rng(3)
    xu = rng.integers(0, 256, (4 * 32, 3, 16, 16)).astype(np.uint8)
    y = np.eye(10)[rng.integers(0, 10, len(xu))]
    runs = []
    for xin in (xu.astype(np.float32) / np.float32(255.0), xu, xu):
        m = build(spec, params, mode="bf16x3", learning_rate=0.001)
        m.train(xin, y, epochs=3, batch_size=32, verbose=0)
        assert len(m._graphs) == 1
        runs.append(m.arena.host_params().copy())
    np.testing.assert_array_equal(runs[0], runs[1])
    np.testing.assert_array_equal(runs[1], runs[2])
    ref = build(spec, params, mode=None, learning_rate=0.001)
    ref.train(xu.astype(np.float32) / np.float32(255.0), y, epochs=1, batch_size=32, verbose=0)
    one = build(spec, params, mode="bf16x3", learning_rate=0.001)
    one.train(xu, y, epochs=1, batch_size=32, verbose=0)
    np.testing.assert_allclose(one.arena.host_params(), ref.arena.host_params(), rtol=1e-3, atol=2e-5)


def test_x3_virtual_shards_global_batch_rule_and_public_api():
    """SURVEY 8e in bf16x3 mode through the public forward / backward: G shards with the global batch as divisor sum to the
    un-split gradient (the data-parallel rule), to FP32 summation order."""
    from neuromah_b200 import parallel
    spec = ref_net.mnist_cnn_spec()
    params = ref_net.init_params(spec, np.random.RandomState(7), init="he")
    B, G = 256, 4
    x, y = synth(8, B, (1, 28, 28))
    full = build(spec, params, mode="bf16x3")
    out = full.forward(x, training=True); full.backward(out, y)
    g_full = full.arena.host_grads().astype(np.float64)
    shard = build(spec, params, mode="bf16x3")
    parallel.set_global_batch(shard, B)
    acc = np.zeros_like(g_full)
    for r in range(G):
        lo, hi = parallel.shard_bounds(B, r, G)
        o = shard.forward(x[lo:hi], training=True); shard.backward(o, y[lo:hi])
        acc += shard.arena.host_grads()
    assert rel_l2(acc, g_full) <= 2e-6
