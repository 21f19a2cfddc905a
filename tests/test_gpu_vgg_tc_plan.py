"""GPU parity of the general bf16 tensor-core plan (neuromah_b200/tc_plan_gen.py) -- the VGG-style stack of
BASELINE.json config 4 -- against the oracle's float64 step on identical inputs and state.

Stated tolerance of the mode (bf16 operands and stored activations, FP32 accumulation), the same figures as the MNIST
plan's tests (tests/test_gpu_tc_model.py): probabilities 2e-3 + 1e-2 |ref|, every gradient tensor <= 3e-2 relative L2.
The stack is narrowed / shortened so that the NumPy oracle finishes in seconds; the kernels take the same code paths
as the full config (K loop over 64-channel chunks, N tiles of 64 and 128, padded first-layer channels)."""
import numpy as np
import pytest

from conftest import synth
from oracle import ref_net
from test_gpu_model import _resync_oracle, build

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-12))


def small_vgg_spec(hw=16):
    """C(3->64) C(64->64) P C(64->128) C(128->128) P Flatten Dense(->10) Softmax on hw x hw inputs."""
    spec, ic = [], 3
    for oc in (64, 128):
        spec += [{"kind": "conv", "ic": ic, "oc": oc, "k": 3, "padding": "same", "relu": True},
                 {"kind": "conv", "ic": oc, "oc": oc, "k": 3, "padding": "same", "relu": True},
                 {"kind": "pool", "size": 2, "stride": 2, "padding": "valid"}]
        ic = oc
    spec += [{"kind": "flatten"}, {"kind": "dense", "n_in": 128 * (hw // 4) ** 2, "n_out": 10, "relu": False}, {"kind": "softmax"}]
    return spec


def test_tc_stack_plan_is_selected_and_rejects_other_nets():
    from neuromah_b200.tc_plan_gen import TcStackPlan
    spec = small_vgg_spec()
    model = build(spec, ref_net.init_params(spec, np.random.RandomState(1), init="he"), mode="bf16")
    assert isinstance(model.plan, TcStackPlan)
    bad = small_vgg_spec()
    bad[0]["oc"] = bad[1]["ic"] = 48                  # out_channels not a multiple of 64
    with pytest.raises(ValueError):
        build(bad, ref_net.init_params(bad, np.random.RandomState(1), init="he"), mode="bf16")


def test_tc_stack_forward_activations_and_probs():
    spec = small_vgg_spec()
    params = ref_net.init_params(spec, np.random.RandomState(3), init="he")
    model = build(spec, params, mode="bf16")
    net = ref_net.RefNet(spec, params)
    x, _ = synth(4, 24, (3, 16, 16))
    out = np.asarray(model.forward(x, training=False))
    ref = net.forward(x)
    np.testing.assert_allclose(out, ref, rtol=1e-2, atol=2e-3)
    for li in range(6):                                # conv outputs (post-ReLU) and pooled activations
        a, r = np.asarray(model.layers[li].output), net.outs[li]
        assert a.shape == r.shape, li
        assert np.abs(a - r).max() <= 2e-2 * np.abs(r).max(), li


def test_tc_stack_step_gradients_from_identical_state():
    spec = small_vgg_spec()
    params = ref_net.init_params(spec, np.random.RandomState(5), init="he")
    model = build(spec, params, mode="bf16", learning_rate=0.001)
    net = ref_net.RefNet(spec, params)
    x, y = synth(6, 3 * 32, (3, 16, 16))
    for s in range(3):
        _resync_oracle(net, model)
        xb, yb = x[s * 32:(s + 1) * 32], y[s * 32:(s + 1) * 32]
        model._accum.fill_zero()
        out = model.train_step(xb, yb)
        r = net.step(xb, yb)
        np.testing.assert_allclose(np.asarray(out), r["output"], rtol=1e-2, atol=2e-3)
        loss_sum, _ = model._read_accum()
        assert abs(loss_sum - r["loss_sum"]) <= 2e-3 * 32
        g, gr = model.arena.host_grads(), net.flat("grads")
        off = 0
        for (_, name, _, size, _) in model.arena.table:
            assert rel_l2(g[off:off + size], gr[off:off + size]) <= 3e-2, (s, name, off)
            off += size
    assert model.optimizer.iterations == 3


def test_tc_stack_bitwise_deterministic_and_graph_replay():
    spec = small_vgg_spec()
    params = ref_net.init_params(spec, np.random.RandomState(9), init="he")
    x, y = synth(10, 128, (3, 16, 16))
    runs = []
    for _ in range(2):
        m = build(spec, params, mode="bf16")
        m.train(x, y, epochs=3, batch_size=32, verbose=0)
        runs.append(m.arena.host_params().copy())
    np.testing.assert_array_equal(runs[0], runs[1])
