"""GPU parity of the FP32-exact tensor-core Dense mode (``Model.finalize(mode="bf16x3")``, nm_tc_dense_*_gen_x3 in
nm_tc_gemm.cu): every operand is three bf16 planes h + m + l (24 mantissa bits), every contraction the six plane-pair
products above 2^-24 accumulated in one FP32 TMEM tile.

Stated tolerance of the mode: the FP32-exact one -- rtol 1e-4 / atol 1e-5 against the reference's NumPy path (north_star),
checked on the reference's own 200-step golden run of BASELINE.json configs[0] (tests/golden/mlp_200steps.npz: loss curve
and post-update weights) exactly as tests/test_gpu_model.py checks the CUDA-core FP32 kernels.  Op level: 4e-6 (fprop; 2e-6 dgrad / wgrad) of the
largest output magnitude against float64 NumPy on float32 inputs (measured <= 1.0e-6 / 3.3e-7 / 2.8e-7: FP32 accumulation order).
Reference arithmetic: Dense.py:64-108, Relu.py:7-14, Dropout.py:21-40.
"""
import numpy as np
import pytest

from conftest import golden, synth
from oracle import ref_net
from test_gpu_model import ATOL, RTOL, _resync_oracle, build, reference_style_step
from test_gpu_tc_mlp import _dense_example_model, _philox_mask, from_bits, mlp3_spec, rel_l2, to_bits

pytestmark = pytest.mark.gpu


def planes_to_f64(p):
    p = from_bits(p).astype(np.float64)
    return p[0] + p[1] + p[2]


def split_dev(a, cap_rows=None):
    """float32 [rows][cols] -> device planes [3][cap_rows][cols] through nm_tc_split3_f32."""
    from neuromah_b200._lib import lib
    from neuromah_b200.device import DeviceArray, stream
    rows, cols = a.shape
    cap = cap_rows or rows
    out = DeviceArray.from_numpy(np.full((3, cap, cols), 0x7FC0, np.uint16))          # NaN-filled: rows >= `rows` must never be read
    src = DeviceArray.from_numpy(np.ascontiguousarray(a, np.float32))
    lib.nm_tc_split3_f32(src.p, None, 1.0, out.p, a.size, cap * cols, stream())
    return out


def test_split3_planes_carry_24_bits_and_mask():
    from neuromah_b200._lib import lib
    from neuromah_b200.device import DeviceArray, stream
    rng = np.random.default_rng(0)
    x = (rng.standard_normal((37, 24)) * np.exp(rng.uniform(-20, 20, (37, 24)))).astype(np.float32)
    x[0, :4] = [0.0, -0.0, 1.0, -3.5]
    p = split_dev(x).numpy()
    err = np.abs(planes_to_f64(p) - x.astype(np.float64))
    assert (err <= 2.0 ** -24 * np.abs(x)).all()
    np.testing.assert_array_equal(p[0], to_bits(x))          # h is the round-to-nearest bf16 of x (what the ReLU masks read)
    # ReLU mask + scale
    mask = rng.standard_normal(x.shape).astype(np.float32)
    mask[1, :3] = [0.0, -0.0, 1e-30]
    md = DeviceArray.from_numpy(to_bits(mask))
    out = DeviceArray.from_numpy(np.zeros((3,) + x.shape, np.uint16))
    lib.nm_tc_split3_f32(DeviceArray.from_numpy(x).p, md.p, 1.25, out.p, x.size, x.size, stream())
    want = np.where(from_bits(to_bits(mask)) > 0, x * np.float32(1.25), np.float32(0))
    got = planes_to_f64(out.numpy())
    assert (np.abs(got - want) <= 2.0 ** -24 * np.abs(want)).all()
    # raw uint8 pixels: float32(u) / 255, the division of the FP32 path (nm_u8_to_f32_div)
    u = rng.integers(0, 256, (5, 16)).astype(np.uint8)
    outu = DeviceArray.from_numpy(np.zeros((3, 5, 16), np.uint16))
    lib.nm_tc_split3_u8(DeviceArray.from_numpy(u).p, 255.0, outu.p, u.size, u.size, stream())
    wantu = u.astype(np.float32) / np.float32(255.0)
    assert (np.abs(planes_to_f64(outu.numpy()) - wantu) <= 2.0 ** -24 * wantu).all()


CASES = [(128, 784, 128), (128, 128, 64), (300, 784, 128), (2048, 784, 128), (64, 64, 64), (1, 8, 64), (257, 200, 192)]


@pytest.mark.parametrize("B,n_in,n_out", CASES)
def test_dense_gen_x3_fprop_dgrad_wgrad_fp32_exact(B, n_in, n_out):
    from neuromah_b200._lib import lib
    from neuromah_b200.device import DeviceArray, stream
    rng = np.random.default_rng(B + n_in + n_out)
    cap = B + 5                                                          # planes wider than the batch: capacity != current batch
    x = rng.standard_normal((B, n_in)).astype(np.float32)
    w = (rng.standard_normal((n_in, n_out)) * 0.05).astype(np.float32)
    b = rng.standard_normal((1, n_out)).astype(np.float32)
    dy = rng.standard_normal((B, n_out)).astype(np.float32)
    w_dev = DeviceArray.from_numpy(w)
    w_kn, w_nk = DeviceArray((3, n_in, n_out), np.uint16), DeviceArray((3, n_out, n_in), np.uint16)
    lib.nm_tc_pack_dense_gen_x3(w_dev.p, w_kn.p, w_nk.p, n_in, n_out, stream())
    assert np.abs(planes_to_f64(w_kn.numpy()) - w).max() <= 2.0 ** -24 * np.abs(w).max()
    np.testing.assert_array_equal(w_nk.numpy(), w_kn.numpy().transpose(0, 2, 1))
    x3 = split_dev(x, cap)
    # ---- fprop + bias + ReLU, planes and the FP32 copy
    y3 = DeviceArray.from_numpy(np.full((3, cap, n_out), 0x7FC0, np.uint16))
    yf = DeviceArray.from_numpy(np.full((B, n_out), np.nan, np.float32))
    lib.nm_tc_dense_fprop_gen_x3(x3.p, cap * n_in, w_nk.p, DeviceArray.from_numpy(b).p, y3.p, cap * n_out, yf.p, B, n_in, n_out, 1,
                                 1.0, 0, 0, None, stream())
    ref = np.maximum(x.astype(np.float64) @ w.astype(np.float64) + b, 0)
    scale = np.abs(ref).max()
    got = yf.numpy()
    print("fprop max err / scale", np.abs(got - ref).max() / scale)
    assert np.abs(got - ref).max() <= 4e-6 * scale
    gp = planes_to_f64(y3.numpy()[:, :B])
    assert np.abs(gp - got).max() <= 2.0 ** -23 * scale                # the planes carry the FP32 result
    assert (y3.numpy()[:, B:] == 0x7FC0).all()                           # rows past the batch untouched
    # ---- dgrad with the ReLU mask of the layer below (its h plane) and a dropout scale
    dy3 = split_dev(dy, cap)
    act = rng.standard_normal((B, n_in)).astype(np.float32)
    act3 = split_dev(act, cap)
    dx3 = DeviceArray.from_numpy(np.full((3, cap, n_in), 0x7FC0, np.uint16))
    lib.nm_tc_dense_dgrad_gen_x3(dy3.p, cap * n_out, w_kn.p, act3.p, 1.0 / 0.75, dx3.p, cap * n_in, B, n_in, n_out, stream())
    live = from_bits(act3.numpy()[0, :B]) > 0
    ref = np.where(live, (dy.astype(np.float64) @ w.astype(np.float64).T) * np.float64(np.float32(1.0 / 0.75)), 0.0)
    got = planes_to_f64(dx3.numpy()[:, :B])
    print("dgrad max err / scale", np.abs(got - ref).max() / np.abs(ref).max())
    assert np.abs(got - ref).max() <= 2e-6 * np.abs(ref).max()
    # ---- wgrad + bias gradient + L2 term: contraction over the batch, rows past B of the planes are NaN and must not be read
    dw, db = DeviceArray((n_in, n_out), np.float32), DeviceArray((1, n_out), np.float32)
    ws = DeviceArray((max(lib.nm_tc_dense_wgrad_gen_workspace(B, n_in, n_out), 256),), np.uint8)
    sc = 1.0 / B
    lib.nm_tc_dense_wgrad_gen_x3(x3.p, cap * n_in, dy3.p, cap * n_out, w_dev.p, dw.p, db.p, ws.p, ws.nbytes, B, n_in, n_out, sc, 0.0, 0.01,
                                 stream())
    ref = (x.astype(np.float64).T @ dy.astype(np.float64)) * np.float64(np.float32(sc)) + 2 * np.float64(np.float32(0.01)) * w
    got = dw.numpy()
    print("wgrad max err / scale", np.abs(got - ref).max() / np.abs(ref).max())
    assert np.abs(got - ref).max() <= 2e-6 * np.abs(ref).max()
    refb = dy.astype(np.float64).sum(0) * np.float64(np.float32(sc))
    assert np.abs(db.numpy().ravel() - refb).max() <= 2e-6 * max(np.abs(refb).max(), np.abs(dy).max() * np.sqrt(B) * sc)


def test_x3_mode_rejects_unsupported_stacks():
    spec = [{"kind": "dense", "n_in": 20, "n_out": 10, "relu": False}, {"kind": "softmax"}]       # no hidden layer: neither plan
    with pytest.raises(ValueError, match="bf16x3"):
        build(spec, ref_net.init_params(spec, np.random.RandomState(0)), mode="bf16x3")


def test_x3_mlp_200_step_golden_fp32_exact_tolerance():
    """BASELINE.json configs[0], the reference's own 200-step run: loss curve, accuracies and post-update weights within
    the FP32-exact tolerance -- the test of the CUDA-core FP32 mode (tests/test_gpu_model.py) on the tensor cores."""
    from neuromah_b200.tc_plan_mlp import TcMlpPlan
    g = golden("mlp_200steps.npz")
    spec = ref_net.mlp_spec()
    model = build(spec, ref_net.init_params(spec, np.random.RandomState(0)), mode="bf16x3", learning_rate=0.001)
    assert isinstance(model.plan, TcMlpPlan) and model.plan.split
    model.loss.new_pass(); model.accuracy.new_pass()
    x, y = synth(1, 128 * 200, (784,))
    losses, accs = zip(*[reference_style_step(model, x[s * 128:(s + 1) * 128], y[s * 128:(s + 1) * 128]) for s in range(200)])
    np.testing.assert_allclose(losses, g["losses"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(accs, g["accs"], atol=2.0 / 128)
    np.testing.assert_allclose(model.arena.host_params(), g["final_params"], rtol=RTOL, atol=ATOL)


def test_x3_mlp_train_api_graph_replay_matches_golden_and_uint8():
    """The same run through Model.train (one CUDA graph, arena Adam, device loss sums), then raw uint8 pixels == the
    host-normalised float32 input bit for bit (the device divides by 255 like the FP32 path)."""
    g = golden("mlp_200steps.npz")
    spec = ref_net.mlp_spec()
    model = build(spec, ref_net.init_params(spec, np.random.RandomState(0)), mode="bf16x3", learning_rate=0.001)
    x, y = synth(1, 128 * 200, (784,))
    model.train(x, y, epochs=1, batch_size=128, verbose=0)
    assert len(model._graphs) == 1
    np.testing.assert_allclose(model.arena.host_params(), g["final_params"], rtol=RTOL, atol=ATOL)
    assert abs(model.last_epoch["data_loss"] - float(np.mean(g["losses"]))) < 1e-4
    assert model.optimizer.iterations == 200
    spec3 = mlp3_spec()
    params = ref_net.init_params(spec3, np.random.RandomState(9), init="he")
    rng = np.random.default_rng(3)
    xu = rng.integers(0, 256, (4 * 128, 784)).astype(np.uint8)
    y3 = np.eye(10)[rng.integers(0, 10, len(xu))]
    runs = []
    for xin in (xu.astype(np.float32) / np.float32(255.0), xu):
        m = build(spec3, params, mode="bf16x3", learning_rate=0.002)
        m.train(xin, y3, epochs=2, batch_size=128, verbose=0)
        runs.append(m.arena.host_params())
    np.testing.assert_array_equal(runs[0], runs[1])


def test_x3_three_layer_step_vs_oracle_per_tensor():
    """784-128-64-10 from identical state: outputs, hidden activations and every gradient tensor against the float64 oracle."""
    spec = mlp3_spec()
    params = ref_net.init_params(spec, np.random.RandomState(5), init="he")
    model = build(spec, params, mode="bf16x3", learning_rate=0.001)
    net = ref_net.RefNet(spec, params)
    batch = 96
    x, y = synth(6, 3 * batch, (784,))
    for s in range(3):
        _resync_oracle(net, model)
        xb, yb = x[s * batch:(s + 1) * batch], y[s * batch:(s + 1) * batch]
        model._accum.fill_zero()
        out = model.train_step(xb, yb)
        r = net.step(xb, yb)
        np.testing.assert_allclose(np.asarray(out), r["output"], rtol=RTOL, atol=ATOL)
        loss_sum, _ = model._read_accum()
        assert abs(loss_sum - r["loss_sum"]) <= 1e-5 * batch
        g, gr = model.arena.host_grads(), net.flat("grads")
        off = 0
        for (_, name, _, size, _) in model.arena.table:
            e = rel_l2(g[off:off + size], gr[off:off + size])
            assert e <= 2e-6, (name, e)
            off += size
        for li, layer in enumerate(model.layers[:-2]):
            np.testing.assert_allclose(np.asarray(layer.output), net.outs[li], rtol=RTOL, atol=ATOL)


def test_x3_dense_example_stack_with_fused_dropout_vs_oracle():
    """examples/test_Dense_MNIST.py:39-44 stack (single softmax): Dropout fused in the split epilogue draws the FP32 layer's
    Philox mask for (seed, step); with those masks the oracle's step agrees to the FP32-exact tolerance."""
    model, spec, params = _dense_example_model("bf16x3")
    assert model.plan.split and len(model.plan.drops) == 2
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
        np.testing.assert_allclose(np.asarray(out), r["output"], rtol=RTOL, atol=ATOL)
        g, gr = model.arena.host_grads(), net.flat("grads")
        off = 0
        for (_, name, _, size, _) in model.arena.table:
            e = rel_l2(g[off:off + size], gr[off:off + size])
            assert e <= 5e-6, (name, e)
            off += size
    p_eval = np.asarray(model.forward(x[:B], training=False))
    net.set_flat(model.arena.host_params().astype(np.float64))
    np.testing.assert_allclose(p_eval, net.forward(x[:B], training=False), rtol=RTOL, atol=ATOL)
