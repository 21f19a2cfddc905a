"""CPU: the conv oracle (oracle/np_ref.py) against the reference's OWN conv source, src/layers/CPU_kernels/Conv2D.cpp,
compiled from where it lies by oracle/build_ref.py against oracle/eigen_standin (the file needs <Eigen/Core>, which the
reference neither vendors nor pins and this image does not have; the stand-in implements the five Eigen operations the
file uses with Eigen's documented column-major semantics -- see its header).

The reference's im2col / col2im / batching / pybind code runs unmodified.  What it pins:

* ONE output channel: forward, dgrad and wgrad equal the oracle's to FP32 rounding -- im2col row order (c, p, q),
  cross-correlation (no kernel flip), stride 1 / no padding, wgrad summed over the batch with no 1/B, dgrad as col2im with
  accumulation over overlapping windows (Conv2D.cpp:19-74, :113-134, :207-222).
* ANY number of output channels: the intended math is independent per output channel, so the oracle's multi-channel result
  equals OC calls of the compiled reference with one filter each (dgrad = the sum of the per-filter results).
* SURVEY Q6, the reference's multi-channel defect: the row-major NumPy kernel and output are mapped as COLUMN-major
  matrices (Conv2D.cpp:114, :134-137), so a direct OC > 1 call returns a storage-order scramble of the intended result.
  The test reproduces the scramble from the oracle (kernel read as flat[o + k*OC], output written as flat[o + p*OC]), which
  is why the mirror implements the intended math and not the as-shipped values.
"""
import numpy as np
import pytest

from oracle import np_ref as R
from oracle._ref_loader import load_ref_conv

REF = load_ref_conv()
pytestmark = pytest.mark.skipif(REF is None, reason="oracle/_ref/_Conv2DBackend_cpp not built (needs /root/reference once)")

RTOL, ATOL = 2e-5, 2e-5          # float32 sums of up to a few hundred products, two summation orders

SHAPES = [
    # N, C, H, W, kH, kW
    (2, 3, 7, 6, 3, 3),
    (3, 1, 5, 5, 2, 4),
    (2, 4, 6, 6, 1, 1),
    (1, 2, 9, 8, 5, 3),
    (4, 5, 4, 4, 4, 4),          # one output pixel
    (2, 1, 28, 28, 3, 3),        # MNIST conv1 geometry (valid)
]


def _data(shape, oc, seed):
    n, c, h, w, kh, kw = shape
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, c, h, w)).astype(np.float32)
    k = rng.standard_normal((oc, c, kh, kw)).astype(np.float32)
    d = rng.standard_normal((n, oc, h - kh + 1, w - kw + 1)).astype(np.float32)
    return x, k, d


@pytest.mark.parametrize("shape", SHAPES)
def test_single_output_channel_matches_the_compiled_reference(shape):
    x, k, d = _data(shape, 1, 11)
    out = REF.conv2d_cpu(x, k)
    assert out.dtype == np.float32 and out.shape == d.shape
    np.testing.assert_allclose(out, R.conv2d_cpu(x, k), rtol=RTOL, atol=ATOL)
    dx, dw = REF.conv2d_backward_cpu(x, k, d)
    odx, odw = R.conv2d_backward_cpu(x, k, d)
    assert dx.shape == x.shape and dw.shape == k.shape
    np.testing.assert_allclose(dx, odx, rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(dw, odw, rtol=RTOL, atol=ATOL * shape[0])         # summed over the batch, no 1/B


@pytest.mark.parametrize("shape,oc", [(SHAPES[0], 4), (SHAPES[3], 3), (SHAPES[5], 32), ((2, 32, 14, 14, 3, 3), 5)])
def test_multi_channel_oracle_equals_the_compiled_reference_filter_by_filter(shape, oc):
    x, k, d = _data(shape, oc, 12)
    out, (dx, dw) = R.conv2d_cpu(x, k), R.conv2d_backward_cpu(x, k, d)
    dx_sum = np.zeros_like(x, dtype=np.float64)
    for o in range(oc):
        ko, do = np.ascontiguousarray(k[o:o + 1]), np.ascontiguousarray(d[:, o:o + 1])
        np.testing.assert_allclose(out[:, o:o + 1], REF.conv2d_cpu(x, ko), rtol=RTOL, atol=ATOL)
        rdx, rdw = REF.conv2d_backward_cpu(x, ko, do)
        np.testing.assert_allclose(dw[o:o + 1], rdw, rtol=RTOL, atol=ATOL * shape[0] * 4)
        dx_sum += rdx
    np.testing.assert_allclose(dx, dx_sum, rtol=RTOL, atol=ATOL * oc)


@pytest.mark.parametrize("shape,oc", [(SHAPES[0], 4), (SHAPES[1], 3), (SHAPES[3], 2)])
def test_direct_multi_channel_call_shows_the_storage_order_defect(shape, oc):
    """what the as-shipped reference returns for OC > 1, predicted from the oracle."""
    x, k, _ = _data(shape, oc, 13)
    n, c, h, w, kh, kw = shape
    out = REF.conv2d_cpu(x, k)
    intended = R.conv2d_cpu(x, k)
    assert np.abs(out - intended).max() > 0.5                                     # not a rounding difference
    K, P = c * kh * kw, (h - kh + 1) * (w - kw + 1)
    k_as_read = k.reshape(K, oc).T.reshape(oc, c, kh, kw)                         # kernelMat(o, k) = flat[o + k*OC]
    y = R.conv2d_cpu(x, np.ascontiguousarray(k_as_read))                          # the product the GEMM really forms
    as_written = np.stack([y[b].reshape(oc, P).T.reshape(oc, h - kh + 1, w - kw + 1) for b in range(n)])   # flat[o + p*OC]
    np.testing.assert_allclose(out, as_written, rtol=RTOL, atol=ATOL)


def test_error_cases_of_the_compiled_reference():
    """Conv2D.cpp:86-87, :101-102 -> RuntimeError (pybind translates std::runtime_error); the mirror raises the same type."""
    with pytest.raises(RuntimeError):
        REF.conv2d_cpu(np.zeros((3, 3), np.float32), np.zeros((1, 1, 3, 3), np.float32))
    with pytest.raises(RuntimeError):
        REF.conv2d_cpu(np.zeros((1, 2, 5, 5), np.float32), np.zeros((4, 3, 3, 3), np.float32))
    with pytest.raises(RuntimeError):
        R.conv2d_cpu(np.zeros((3, 3), np.float32), np.zeros((1, 1, 3, 3), np.float32))
    with pytest.raises(RuntimeError):
        R.conv2d_cpu(np.zeros((1, 2, 5, 5), np.float32), np.zeros((4, 3, 3, 3), np.float32))


def test_same_padding_through_the_wrapper_call_pattern():
    """Conv_wrapepr.py:78-91: np.pad outside, the backend on the padded input -- with the compiled reference as the
    backend (one filter) against the oracle's layer-level forward."""
    rng = np.random.default_rng(14)
    x = rng.standard_normal((3, 2, 8, 8)).astype(np.float32)
    for k in ((3, 3), (2, 2), (4, 3), (5, 5)):
        w = rng.standard_normal((1, 2) + k).astype(np.float32)
        (pt, pb), (pl, pr) = R.same_padding(*k)
        xp = np.pad(x, ((0, 0), (0, 0), (pt, pb), (pl, pr)))
        out = REF.conv2d_cpu(xp, w)
        assert out.shape == (3, 1, 8, 8)
        ref, _ = R.conv_layer_forward(x, w, np.zeros((1, 1)), padding="same", relu=False)
        np.testing.assert_allclose(out, ref, rtol=RTOL, atol=ATOL)


def test_oracle_step_driver_equals_the_reference_model_on_its_own_compiled_kernels():
    """A one-filter CNN -- Conv(3->1, 3x3, same, ReLU) -> MaxPool 2 -> Flatten -> Dense(16->4) -> Softmax, Adam -- stepped
    three times by the UNMODIFIED reference Model with BOTH of its own compiled kernels at the `compiled` seam (Conv2D.cpp
    through the Eigen stand-in, maxpooling_backend.cpp), against the oracle's step driver from the same initial weights.
    The only non-reference line in the reference run is the wiring of Layer_Conv2D.backward to the backend's own
    conv2d_backward_cpu (the shipped backward is a stub)."""
    from oracle import ref_net, ref_shim
    if not ref_shim.available():
        pytest.skip("reference tree only in the authoring container")
    ref_shim.load()
    from NeuromahRef.src.layers.compiled import _Conv2DBackend_cpp as seam
    spec = [{"kind": "conv", "ic": 3, "oc": 1, "k": (3, 3), "padding": "same", "relu": True},
            {"kind": "pool", "size": (2, 2)}, {"kind": "flatten"},
            {"kind": "dense", "n_in": 16, "n_out": 4, "relu": False}, {"kind": "softmax"}]
    rs = np.random.RandomState(21)
    params = [{"weights": rs.randn(1, 3, 3, 3) * 0.4, "biases": np.zeros((1, 1))}, None, None,
              {"weights": rs.randn(16, 4) * 0.3, "biases": np.zeros((1, 4))}, None]
    rng = np.random.default_rng(22)
    x = rng.random((8 * 3, 3, 8, 8), dtype=np.float32)
    y = np.eye(4)[rng.integers(0, 4, 8 * 3)]
    saved = (seam.conv2d_cpu, seam.conv2d_backward_cpu)
    try:
        seam.conv2d_cpu, seam.conv2d_backward_cpu = REF.conv2d_cpu, REF.conv2d_backward_cpu
        model = ref_shim.build_model(spec, params, optimizer="adam", lr=0.01)
        ref_losses = [ref_shim.train_step(model, x[s * 8:(s + 1) * 8], y[s * 8:(s + 1) * 8])["loss"] for s in range(3)]
        conv, dense = model.layers[0], model.layers[3]
        ref_flat = np.concatenate([conv.weights.ravel(), conv.biases.ravel(), dense.weights.ravel(), dense.biases.ravel()])
    finally:
        seam.conv2d_cpu, seam.conv2d_backward_cpu = saved
    net = ref_net.RefNet(spec, params, optimizer="adam", lr=0.01)
    losses = [net.step(x[s * 8:(s + 1) * 8], y[s * 8:(s + 1) * 8])["loss"] for s in range(3)]
    np.testing.assert_allclose(losses, ref_losses, rtol=1e-6)
    np.testing.assert_allclose(net.flat(), ref_flat, rtol=1e-4, atol=1e-6)
